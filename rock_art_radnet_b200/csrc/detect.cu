// detect.cu - detection post-processing that follows the RoI pool and the classifier head
// (SURVEY.md 8(f), rows f1 and f2).
//
//   K5 classify_decode  per-RoI class decision + scalar apply_regr of
//                       RADNet.apply_spatial_pyramid_pooling (reference faster_rcnn/RADNet.py:123-150,
//                       rpn.py:346-378)
//   K6 class_nms        per-class greedy NMS of one tile's labelled boxes (rpn.py:380-456 as called at
//                       RADNet.py:574/639/698) + get_real_coordinates (RADNet.py:44-51) + tile offset
//                       (RADNet.py:582-600); K5 and K6 also exist fused in one launch
//   K7 final_nms        cluster-and-average merge of the tiles of one image (RADNet.py:156-240)
//
// Everything here is tiny (a few hundred boxes per tile) and latency-bound, so each kernel keeps its
// whole working set in shared memory:
//   * K5/K6 use one CTA per tile.  Classes are independent NMS problems, so entries are ranked inside
//     their class by counting, the "earlier box suppresses later box" relation of every class is
//     evaluated once into a bit matrix (one warp per row, one ballot per 32 pairs), and then ONE WARP
//     PER CLASS walks its row list with the removed-set held one word per lane.
//   * K7 uses one CTA per (image, class): bitonic sort of (score, position) keys, then one block-wide
//     pass per cluster; a dedicated finaliser warp averages cluster k while the other warps already
//     search cluster k+1 (double-buffered member bitmaps, one __syncthreads per cluster).  The last
//     CTA of an image packs the per-class results into the output record.
//
// Exactness: boxes are integers, so inter and union are exact; the predicate is evaluated as the
// reference does, `inter / ((area_i + area_j - inter) + 1e-6) > thr`, with one IEEE float64 divide.
// Mixed float32 / Python-float operations of the reference follow the NumPy >= 2 (NEP 50) rule, i.e.
// they are float32 operations (oracle/detect_oracle.py states the same).
#include "common.cuh"

namespace radnet {

constexpr int kDetThreads = 1024;
constexpr int kDetWarps = kDetThreads / 32;
constexpr int kMaxClasses = 32;
constexpr int kMatrixCap = 1024;        // entries per segment handled by the bit-matrix kernel
constexpr int kClusterCap = 4096;       // entries per (segment, class) handled by the cluster kernel
constexpr int kRecHeader = 288;         // int32 hdr[8] + order[32] + count[32]
constexpr int kCoordLimit = 1 << 25;    // |coordinate| bound that keeps areas exact in float64

struct __align__(16) DetEntry {
    int cls;
    float prob;
    int x1, y1, x2, y2;
    int src;    // K5: RoI index; K6: index in the concatenated input; K7: input index of the cluster's top box
    int aux;    // K7: number of members averaged; else 0
};
static_assert(sizeof(DetEntry) == 32, "DetEntry layout is part of the ABI");

enum { H_NDET = 0, H_NIN, H_NDEGEN, H_NTIES, H_NCLASSES, H_NFALLBACK, H_NNEARTIE, H_NRANGE };

__device__ __forceinline__ int32_t *rec_hdr(unsigned char *r) { return reinterpret_cast<int32_t *>(r); }
__device__ __forceinline__ int32_t *rec_order(unsigned char *r) { return reinterpret_cast<int32_t *>(r) + 8; }
__device__ __forceinline__ int32_t *rec_count(unsigned char *r) { return reinterpret_cast<int32_t *>(r) + 40; }
__device__ __forceinline__ DetEntry *rec_entries(unsigned char *r) { return reinterpret_cast<DetEntry *>(r + kRecHeader); }
__device__ __forceinline__ const int32_t *rec_hdr(const unsigned char *r) { return reinterpret_cast<const int32_t *>(r); }
__device__ __forceinline__ const DetEntry *rec_entries(const unsigned char *r) {
    return reinterpret_cast<const DetEntry *>(r + kRecHeader);
}

// `inter/(union+1e-6) > thr` on integer boxes, float64 as in rpn.py:429-447 / RADNet.py:206-223
__device__ __forceinline__ bool overlap_gt(const int4 &a, const int4 &b, double thr) {
    const int iw = min(a.z, b.z) - max(a.x, b.x);
    const int ih = min(a.w, b.w) - max(a.y, b.y);
    if (iw <= 0 || ih <= 0) return 0.0 > thr;
    const double inter = (double)((long long)iw * ih);
    const double aa = (double)((long long)(a.z - a.x) * (a.w - a.y));
    const double ab = (double)((long long)(b.z - b.x) * (b.w - b.y));
    const double uni = __dsub_rn(__dadd_rn(aa, ab), inter);
    return __ddiv_rn(inter, __dadd_rn(uni, 1e-6)) > thr;
}

// block-wide exclusive scan of one int per thread; *total = grand total (same in every thread).
// s_scan: 34 ints.  Contains two __syncthreads.
__device__ __forceinline__ int det_block_exscan(int v, int *s_scan, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    __syncthreads();                    // previous users of s_scan are done
    if (lane == 31) s_scan[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int t = (lane < (int)(blockDim.x >> 5)) ? s_scan[lane] : 0;
        int ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += n;
        }
        s_scan[lane] = ti - t;
        if (lane == 31) s_scan[32] = ti;
    }
    __syncthreads();
    *total = s_scan[32];
    return s_scan[w] + inc - v;
}

// Python / NumPy float floor division a // b (npy_divmod), then int(round()) - RADNet.py:46-49
__device__ __forceinline__ int real_coordinate(int v, double ratio) {
    const double a = (double)v;
    const double mod = fmod(a, ratio);
    double div = __ddiv_rn(__dsub_rn(a, mod), ratio);
    if (mod != 0.0 && ((ratio < 0.0) != (mod < 0.0))) div = __dsub_rn(div, 1.0);
    double fl = 0.0;
    if (div != 0.0) {
        fl = floor(div);
        if (__dsub_rn(div, fl) > 0.5) fl = __dadd_rn(fl, 1.0);
    }
    return __double2int_rn(fl);
}

__device__ __forceinline__ int near_half_flag(double v) {
    const double f = v - floor(v);
    return fabs(f - 0.5) < 1e-9 ? 2 : 0;
}

struct RoiDecision {
    int cls;        // -1 = skipped (below bbox_threshold or 'bg')
    float prob;
    int4 box;       // x1,y1,x2,y2 in resized-image pixels
    int flags;      // 1 regression fell back to the RoI, 2 round near-tie, 4 coordinate out of range
};

// RADNet.py:123-150 for one RoI (x,y,w,h in feature cells)
__device__ RoiDecision decide_roi(const float *pc, const float *pr, int n_cls, int4 roi, float thr,
                                  const float *std4, int stride) {
    RoiDecision o;
    o.cls = -1; o.prob = 0.f; o.box = make_int4(0, 0, 0, 0); o.flags = 0;
    // np.max / np.argmax: first maximum; a NaN wins and the first NaN is the argmax
    float m = pc[0];
    int am = 0;
    bool isnan_ = (m != m);
    for (int k = 1; k < n_cls; ++k) {
        const float v = pc[k];
        if (!isnan_) {
            if (v != v) { isnan_ = true; m = v; am = k; }
            else if (v > m) { m = v; am = k; }
        }
    }
    if (m < thr || am == n_cls - 1) return o;                         // RADNet.py:126
    const float tx = __fdiv_rn(pr[4 * am + 0], std4[0]);               // float32 divides, RADNet.py:140-143
    const float ty = __fdiv_rn(pr[4 * am + 1], std4[1]);
    const float tw = __fdiv_rn(pr[4 * am + 2], std4[2]);
    const float th = __fdiv_rn(pr[4 * am + 3], std4[3]);
    const double x = roi.x, y = roi.y, w = roi.z, h = roi.w;
    const double cx = __dadd_rn(x, __dmul_rn(w, 0.5));                 // rpn.py:352-359
    const double cy = __dadd_rn(y, __dmul_rn(h, 0.5));
    const double cx1 = __dadd_rn(__dmul_rn((double)tx, w), cx);
    const double cy1 = __dadd_rn(__dmul_rn((double)ty, h), cy);
    const double w1 = __dmul_rn(exp((double)tw), w);
    const double h1 = __dmul_rn(exp((double)th), h);
    const double x1 = __dsub_rn(cx1, __dmul_rn(w1, 0.5));
    const double y1 = __dsub_rn(cy1, __dmul_rn(h1, 0.5));
    double rx = x, ry = y, rw = w, rh = h;
    const bool finite = isfinite(x1) && isfinite(y1) && isfinite(w1) && isfinite(h1);
    if (finite) {                                                      // rpn.py:360-363
        o.flags |= near_half_flag(x1) | near_half_flag(y1) | near_half_flag(w1) | near_half_flag(h1);
        rx = rint(x1); ry = rint(y1); rw = rint(w1); rh = rint(h1);
    } else {
        o.flags |= 1;                                                  // ValueError / OverflowError path, rpn.py:366-372
    }
    const double lim = (double)(kCoordLimit / (stride > 0 ? stride : 1) / 2);
    if (!(fabs(rx) <= lim && fabs(ry) <= lim && fabs(rw) <= lim && fabs(rh) <= lim)) {
        o.flags |= 4;
        rx = fmax(-lim, fmin(lim, rx)); ry = fmax(-lim, fmin(lim, ry));
        rw = fmax(-lim, fmin(lim, rw)); rh = fmax(-lim, fmin(lim, rh));
    }
    const int ix = (int)rx, iy = (int)ry, iw = (int)rw, ih = (int)rh;
    o.cls = am;
    o.prob = m;
    o.box = make_int4(stride * ix, stride * iy, stride * (ix + iw), stride * (iy + ih));   // RADNet.py:149
    return o;
}

// ------------------------------------------------------------------------------------------------
// K5 / K6: one CTA per tile
// ------------------------------------------------------------------------------------------------
enum { kModeDecode = 0, kModeHeadNms = 1, kModeRecNms = 2 };

struct ClassNmsParams {
    // head source (kModeDecode, kModeHeadNms)
    const float *p_cls, *p_regr;
    int R, n_cls;
    const unsigned char *det;
    size_t det_stride;
    int det_max_boxes;
    const int32_t *rois, *roi_count;
    float bbox_thr;
    float std4[4];
    int stride;
    // record source (kModeRecNms)
    const unsigned char *rec_in;
    size_t in_stride;
    int n_in;
    const int32_t *in_count;
    // NMS + coordinate transform
    double thr;
    int max_boxes;
    const double *ratio;
    const int32_t *origin;
    // output
    unsigned char *rec_out;
    size_t out_stride;
    int out_max_det;
    int cap, words;
};

template <int kMode>
__global__ void __launch_bounds__(kDetThreads, 1) class_nms_kernel(ClassNmsParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    int4 *s_box = reinterpret_cast<int4 *>(smem);                                  // [cap]
    uint32_t *s_key = reinterpret_cast<uint32_t *>(s_box + p.cap);                 // [cap] score key
    int *s_cls = reinterpret_cast<int *>(s_key + p.cap);                           // [cap]
    int *s_src = s_cls + p.cap;                                                    // [cap]
    int *s_sorted = s_src + p.cap;                                                 // [cap] sorted position -> entry
    int *s_slot = s_sorted + p.cap;                                                // [cap] sorted position -> pick ordinal or -1
    uint32_t *s_mat = reinterpret_cast<uint32_t *>(s_slot + p.cap);                // [cap][words]
    __shared__ int s_scan[34];
    __shared__ int s_first[kMaxClasses], s_ccount[kMaxClasses], s_cbase[kMaxClasses], s_ckept[kMaxClasses],
        s_obase[kMaxClasses], s_order[kMaxClasses];
    __shared__ int s_stat[8];
    __shared__ int s_n;

    const int seg = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned char *out = p.rec_out + (size_t)seg * p.out_stride;
    if (threadIdx.x < kMaxClasses) {
        s_first[threadIdx.x] = 0x7fffffff;
        s_ccount[threadIdx.x] = 0;
        s_ckept[threadIdx.x] = 0;
        s_order[threadIdx.x] = -1;
    }
    if (threadIdx.x < 8) s_stat[threadIdx.x] = 0;
    __syncthreads();

    // ---- 1. entries -> shared memory, input order ---------------------------------------------
    int n = 0;
    bool fault = false;
    if (kMode != kModeRecNms) {
        int count = p.R;
        const unsigned char *det = p.det ? p.det + (size_t)seg * p.det_stride : nullptr;
        if (det) count = min(max(reinterpret_cast<const int32_t *>(det)[0], 0), min(p.R, p.det_max_boxes));
        else if (p.roi_count) count = min(max(p.roi_count[seg], 0), p.R);
        const int r = threadIdx.x;
        RoiDecision d;
        d.cls = -1; d.flags = 0;
        if (r < count) {
            int4 roi;
            if (det) {
                const int4 b = reinterpret_cast<const int4 *>(det + 16)[r];         // x1,y1,x2,y2
                roi = make_int4(b.x, b.y, b.z - b.x, b.w - b.y);                    // RADNet.py:564-565
            } else {
                roi = reinterpret_cast<const int4 *>(p.rois)[(size_t)seg * p.R + r];
            }
            const size_t row = (size_t)seg * p.R + r;
            d = decide_roi(p.p_cls + row * p.n_cls, p.p_regr + row * 4 * (p.n_cls - 1), p.n_cls, roi,
                           p.bbox_thr, p.std4, p.stride);
        }
        const bool keep = d.cls >= 0;
        if (keep) {
            if (d.flags & 1) atomicAdd(&s_stat[H_NFALLBACK], 1);
            if (d.flags & 2) atomicAdd(&s_stat[H_NNEARTIE], 1);
            if (d.flags & 4) atomicAdd(&s_stat[H_NRANGE], 1);
        }
        const int slot = det_block_exscan(keep ? 1 : 0, s_scan, &n);
        if (keep) {
            s_box[slot] = d.box;
            s_key[slot] = score_to_key(d.prob);
            s_cls[slot] = d.cls;
            s_src[slot] = r;
        }
        if (threadIdx.x == 0) s_stat[H_NIN] = count;
    } else {
        // concatenate the entries of n_in records
        int n_rec = p.n_in;
        if (p.in_count) n_rec = min(max(p.in_count[seg], 0), p.n_in);
        const unsigned char *base = p.rec_in + (size_t)seg * p.n_in * p.in_stride;
        int cnt = 0;
        if ((int)threadIdx.x < n_rec) {
            cnt = rec_hdr(base + (size_t)threadIdx.x * p.in_stride)[H_NDET];
            if (cnt < 0) { atomicExch(&s_stat[7], 1); cnt = 0; }
        }
        const int off = det_block_exscan(cnt, s_scan, &n);
        // the exclusive offsets are parked in s_slot (n_in <= cap is checked by the host)
        if ((int)threadIdx.x < n_rec) s_slot[threadIdx.x] = off;
        __syncthreads();
        fault = s_stat[7] != 0;
        if (n <= p.cap) {
            for (int j = w; j < n_rec; j += kDetWarps) {
                const unsigned char *rj = base + (size_t)j * p.in_stride;
                const int cj = max(rec_hdr(rj)[H_NDET], 0);
                const int oj = s_slot[j];
                const DetEntry *e = rec_entries(rj);
                for (int i = lane; i < cj; i += 32) {
                    const DetEntry v = e[i];
                    s_box[oj + i] = make_int4(v.x1, v.y1, v.x2, v.y2);
                    s_key[oj + i] = score_to_key(v.prob);
                    s_cls[oj + i] = v.cls;
                    s_src[oj + i] = oj + i;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_stat[H_NIN] = n; s_stat[7] = 0; }
    }
    __syncthreads();
    if (n > p.cap || fault) {          // capacity exceeded (-2) or an input record carried a fault (-1)
        if (threadIdx.x == 0) {
            int32_t *h = rec_hdr(out);
            for (int i = 0; i < 8; ++i) h[i] = 0;
            h[H_NDET] = fault ? -1 : -2;
            h[H_NIN] = n;
        }
        return;
    }

    // ---- 2. classes: counts, first appearance, order --------------------------------------------
    for (int i = threadIdx.x; i < n; i += kDetThreads) {
        const int c = s_cls[i];
        const int4 b = s_box[i];
        atomicAdd(&s_ccount[c], 1);
        atomicMin(&s_first[c], i);
        if (b.x >= b.z || b.y >= b.w) atomicAdd(&s_stat[H_NDEGEN], 1);             // rpn.py:400-401 would assert
        if (max(max(abs(b.x), abs(b.y)), max(abs(b.z), abs(b.w))) > kCoordLimit) atomicAdd(&s_stat[H_NRANGE], 1);
    }
    __syncthreads();
    if (threadIdx.x < kMaxClasses) {
        const int c = threadIdx.x;
        if (s_ccount[c] > 0) {
            int rank = 0, basec = 0;
            for (int o = 0; o < kMaxClasses; ++o)
                if (s_ccount[o] > 0 && s_first[o] < s_first[c]) { ++rank; basec += s_ccount[o]; }
            s_order[rank] = c;
            s_cbase[c] = basec;
            atomicAdd(&s_stat[H_NCLASSES], 1);
        } else {
            s_cbase[c] = 0;
        }
    }
    __syncthreads();

    if constexpr (kMode == kModeDecode) {
        // apply_spatial_pyramid_pooling returns the boxes in RoI order (RADNet.py:130-150)
        DetEntry *e = rec_entries(out);
        for (int i = threadIdx.x; i < p.out_max_det; i += kDetThreads) {
            DetEntry v;
            if (i < n) {
                const int4 b = s_box[i];
                v.cls = s_cls[i]; v.prob = key_to_score(s_key[i]);
                v.x1 = b.x; v.y1 = b.y; v.x2 = b.z; v.y2 = b.w; v.src = s_src[i]; v.aux = 0;
            } else {
                v.cls = 0; v.prob = 0.f; v.x1 = v.y1 = v.x2 = v.y2 = 0; v.src = 0; v.aux = 0;
            }
            e[i] = v;
        }
        if (threadIdx.x < kMaxClasses) {
            rec_order(out)[threadIdx.x] = s_order[threadIdx.x];
            rec_count(out)[threadIdx.x] = s_ccount[threadIdx.x];
        }
        if (threadIdx.x < 8) rec_hdr(out)[threadIdx.x] = (threadIdx.x == H_NDET) ? (n > p.out_max_det ? -2 : n) : s_stat[threadIdx.x];
        return;
    }
    if constexpr (kMode != kModeDecode) {

    // ---- 3. rank inside the class: descending score, higher input index first on ties --------
    for (int i = threadIdx.x; i < n; i += kDetThreads) {
        const int c = s_cls[i];
        const uint32_t k = s_key[i];
        int rank = 0;
        bool tie = false;
        for (int j = 0; j < n; ++j) {
            const bool same = s_cls[j] == c;
            const uint32_t kj = s_key[j];
            rank += (same && (kj > k || (kj == k && j > i))) ? 1 : 0;
            tie |= same && kj == k && j != i;
        }
        const int pos = s_cbase[c] + rank;
        s_sorted[pos] = i;
        s_slot[pos] = -1;
        if (tie) atomicAdd(&s_stat[H_NTIES], 1);
    }
    __syncthreads();

    // ---- 4. bit matrix: row a, bit (b - class base): "a suppresses b", b ranked after a ------
    for (int a = w; a < n; a += kDetWarps) {
        const int ea = s_sorted[a];
        const int c = s_cls[ea];
        const int cb = s_cbase[c], ce = cb + s_ccount[c];
        const int4 ba = s_box[ea];
        uint32_t *row = s_mat + (size_t)a * p.words;
        for (int wd = (a - cb) >> 5; wd <= (ce - 1 - cb) >> 5; ++wd) {
            const int b = cb + (wd << 5) + lane;
            bool hit = false;
            if (b > a && b < ce) hit = overlap_gt(ba, s_box[s_sorted[b]], p.thr);
            const uint32_t bits = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) row[wd] = bits;
        }
    }
    __syncthreads();

    // ---- 5. greedy pass: one warp per class, removed set one word per lane --------------------
    for (int oc = w; oc < kMaxClasses; oc += kDetWarps) {
        const int c = s_order[oc];
        if (c < 0) continue;
        const int cb = s_cbase[c], cn = s_ccount[c];
        const int nw = (cn + 31) >> 5;
        uint32_t removed = 0;              // lane l: word l of the class-relative removed set
        int kept = 0;
        for (int wd = 0; wd < nw && kept < p.max_boxes; ++wd) {
            const int hi = min(32, cn - (wd << 5));
            const uint32_t in_range = (hi == 32) ? 0xffffffffu : ((1u << hi) - 1u);
            uint32_t visited = 0;
            while (kept < p.max_boxes) {
                const uint32_t cur = __shfl_sync(0xffffffffu, removed, wd);
                const uint32_t cand = ~cur & in_range & ~visited;
                if (!cand) break;
                const int bit = __ffs(cand) - 1;
                visited |= (bit == 31) ? 0xffffffffu : ((2u << bit) - 1u);
                const int a = cb + (wd << 5) + bit;
                if (lane == 0) s_slot[a] = kept;
                ++kept;
                if (lane >= wd && lane < nw) removed |= s_mat[(size_t)a * p.words + lane];
            }
        }
        if (lane == 0) s_ckept[c] = kept;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int o = 0;
        for (int r = 0; r < kMaxClasses; ++r) {
            const int c = s_order[r];
            if (c < 0) break;
            s_obase[c] = o;
            o += s_ckept[c];
        }
        s_n = o;
    }
    __syncthreads();

    // ---- 6. record: classes in first-appearance order, boxes in pick order --------------------
    const int n_out = min(s_n, p.out_max_det);
    DetEntry *e = rec_entries(out);
    const bool xf = p.ratio != nullptr;
    const double ratio = xf ? p.ratio[seg] : 1.0;
    const int ox = p.origin ? p.origin[2 * seg] : 0, oy = p.origin ? p.origin[2 * seg + 1] : 0;
    for (int a = threadIdx.x; a < n; a += kDetThreads) {
        const int k = s_slot[a];
        if (k < 0) continue;
        const int i = s_sorted[a];
        const int o = s_obase[s_cls[i]] + k;
        if (o >= p.out_max_det) continue;
        int4 b = s_box[i];
        if (xf) {                                                                   // RADNet.py:46-49, 582-600
            b.x = real_coordinate(b.x, ratio); b.y = real_coordinate(b.y, ratio);
            b.z = real_coordinate(b.z, ratio); b.w = real_coordinate(b.w, ratio);
        }
        DetEntry v;
        v.cls = s_cls[i]; v.prob = key_to_score(s_key[i]);
        v.x1 = b.x + ox; v.y1 = b.y + oy; v.x2 = b.z + ox; v.y2 = b.w + oy;
        v.src = s_src[i]; v.aux = 0;
        e[o] = v;
    }
    for (int i = n_out + threadIdx.x; i < p.out_max_det; i += kDetThreads) {
        DetEntry v;
        v.cls = 0; v.prob = 0.f; v.x1 = v.y1 = v.x2 = v.y2 = 0; v.src = 0; v.aux = 0;
        e[i] = v;
    }
    if (threadIdx.x < kMaxClasses) {
        rec_order(out)[threadIdx.x] = s_order[threadIdx.x];
        rec_count(out)[threadIdx.x] = s_ckept[threadIdx.x];
    }
    if (threadIdx.x < 8) rec_hdr(out)[threadIdx.x] = (threadIdx.x == H_NDET) ? (s_n > p.out_max_det ? -2 : n_out) : s_stat[threadIdx.x];
    }   // kMode != kModeDecode
}

// ------------------------------------------------------------------------------------------------
// K7 (and the large-segment form of K6): one CTA per (segment, class), one pass per cluster
// ------------------------------------------------------------------------------------------------
struct ClusterParams {
    const unsigned char *rec_in;
    size_t in_stride;
    int n_in;
    const int32_t *in_count;
    int n_cls;
    int average;            // 1 = final_nms (cluster + average), 0 = plain greedy NMS
    double thr;
    float conf_thr;
    int n_obj_avg;
    int max_boxes;          // plain NMS only
    const double *ratio;    // plain NMS only (optional)
    const int32_t *origin;
    unsigned char *rec_out;
    size_t out_stride;
    int out_max_det;
    unsigned char *ws;      // int counter[S] (padded), then [S][n_cls] { int hdr[8]; DetEntry e[cap] }
    size_t ws_counters, ws_seg_stride, ws_cls_stride;
    int cap;                // entries per (segment, class) in shared memory (power of two)
    int skip_if_done;       // plain NMS: leave segments whose record was already written by the bit-matrix kernel
};

enum { C_COUNT = 0, C_FIRST, C_TIES, C_DEGEN, C_EMPTY, C_NIN, C_FAULT, C_RANGE };

// NumPy's float32 add-reduce order for a contiguous array (pairwise, blocks of 128, eight partial sums)
__device__ float pairwise_sum_f32(const float *a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(pairwise_sum_f32(a, n2), pairwise_sum_f32(a + n2, n - n2));
}

__global__ void __launch_bounds__(kDetThreads, 1) cluster_kernel(ClusterParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(smem);      // [cap] (score key << 32) | position
    int4 *s_box = reinterpret_cast<int4 *>(s_keys + p.cap);                         // [cap] by class position
    int *s_src = reinterpret_cast<int *>(s_box + p.cap);                            // [cap] index in the concatenated input
    uint32_t *s_alive = reinterpret_cast<uint32_t *>(s_src + p.cap);                // [cap/32]
    uint32_t *s_member = s_alive + p.cap / 32;                                      // [2][cap/32]
    int *s_list = reinterpret_cast<int *>(s_member + 2 * (p.cap / 32));             // [cap] finaliser: ordered members
    float *s_vals = reinterpret_cast<float *>(s_list + p.cap);                      // [cap] finaliser: probs to average
    int *s_off = reinterpret_cast<int *>(s_vals + p.cap);                           // [n_in + 1]
    __shared__ int s_scan[34];
    __shared__ int s_wmax[2][kDetWarps];
    __shared__ int s_stat[8];
    __shared__ int s_last;

    const int c = blockIdx.x, seg = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int32_t *seg_counter = reinterpret_cast<int32_t *>(p.ws) + seg;
    unsigned char *ws_seg = p.ws + p.ws_counters + (size_t)seg * p.ws_seg_stride;
    unsigned char *ws_cls = ws_seg + (size_t)c * p.ws_cls_stride;
    int32_t *chdr = reinterpret_cast<int32_t *>(ws_cls);
    DetEntry *cout = reinterpret_cast<DetEntry *>(ws_cls + 32);
    if (p.skip_if_done &&
        reinterpret_cast<const volatile int32_t *>(p.rec_out + (size_t)seg * p.out_stride)[H_NDET] != -2)
        return;               // few enough boxes: the bit-matrix kernel launched before this one has done the segment
    if (threadIdx.x < 8) s_stat[threadIdx.x] = (threadIdx.x == C_FIRST) ? 0x7fffffff : 0;
    __syncthreads();

    // ---- 1. offsets of the input records ------------------------------------------------------
    int n_rec = p.n_in;
    if (p.in_count) n_rec = min(max(p.in_count[seg], 0), p.n_in);
    const unsigned char *base = p.rec_in + (size_t)seg * p.n_in * p.in_stride;
    int total = 0;
    for (int j0 = 0; j0 < n_rec; j0 += kDetThreads) {        // n_in may exceed the block size
        const int j = j0 + threadIdx.x;
        int cnt = 0;
        if (j < n_rec) {
            cnt = rec_hdr(base + (size_t)j * p.in_stride)[H_NDET];
            if (cnt < 0) { s_stat[C_FAULT] = 1; cnt = 0; }
        }
        int chunk_total;
        const int off = det_block_exscan(cnt, s_scan, &chunk_total);
        if (j < n_rec) s_off[j] = total + off;
        total += chunk_total;
    }
    if (threadIdx.x == 0) s_off[n_rec] = total;
    __syncthreads();

    // ---- 2. ordered compaction of this class's entries ----------------------------------------
    int m = 0;
    for (int g0 = 0; g0 < total; g0 += kDetThreads) {
        const int g = g0 + threadIdx.x;
        bool mine = false;
        DetEntry v;
        if (g < total) {
            int lo = 0, hi = n_rec;                           // last record with s_off[j] <= g
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= g) lo = mid; else hi = mid;
            }
            v = rec_entries(base + (size_t)lo * p.in_stride)[g - s_off[lo]];
            mine = v.cls == c;
        }
        int got;
        const int slot = m + det_block_exscan(mine ? 1 : 0, s_scan, &got);
        if (mine && slot < p.cap) {
            s_box[slot] = make_int4(v.x1, v.y1, v.x2, v.y2);
            s_keys[slot] = ((unsigned long long)score_to_key(v.prob) << 32) | (unsigned)slot;
            s_src[slot] = g;
            if (slot == 0) s_stat[C_FIRST] = g;
            if (v.x1 >= v.x2 || v.y1 >= v.y2) atomicAdd(&s_stat[C_DEGEN], 1);      // RADNet.py:176-177 would assert
            if (max(max(abs(v.x1), abs(v.y1)), max(abs(v.x2), abs(v.y2))) > kCoordLimit) atomicAdd(&s_stat[C_RANGE], 1);
        }
        m += got;
    }
    __syncthreads();
    const bool overflow = m > p.cap;
    int n_out = 0;
    if (!overflow && m > 0) {
        // ---- 3. ascending bitonic sort of (score key, position); padding sorts last ------------
        int P = 1;
        while (P < m) P <<= 1;
        for (int i = m + threadIdx.x; i < P; i += kDetThreads) s_keys[i] = ~0ull;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < P; i += kDetThreads) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = s_keys[i], b = s_keys[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { s_keys[i] = b; s_keys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        // score ties (reported)
        {
            int t = 0;
            for (int i = threadIdx.x; i < m; i += kDetThreads) {
                const uint32_t k = (uint32_t)(s_keys[i] >> 32);
                const bool tie = (i > 0 && (uint32_t)(s_keys[i - 1] >> 32) == k) ||
                                 (i + 1 < m && (uint32_t)(s_keys[i + 1] >> 32) == k);
                t += tie ? 1 : 0;
            }
            t = __reduce_add_sync(0xffffffffu, t);
            if (lane == 0 && t) atomicAdd(&s_stat[C_TIES], t);
        }
        const int nwords = (m + 31) >> 5;
        for (int i = threadIdx.x; i < nwords; i += kDetThreads) {
            const int hi = min(32, m - (i << 5));
            s_alive[i] = (hi == 32) ? 0xffffffffu : ((1u << hi) - 1u);
        }
        __syncthreads();

        // ---- 4. one pass per cluster ------------------------------------------------------------
        // Warps 0..30 search; warp 31 finalises the previous cluster meanwhile.
        const float thr_c = p.conf_thr;
        int top = m - 1;
        int it = 0;
        int prev_top = -1;
        const int n_search = kDetWarps - 1;
        while (true) {
            const int buf = it & 1;
            if (w < n_search) {
                int wmax = -1;
                if (top >= 0) {
                    const int4 tb = s_box[(uint32_t)s_keys[top]];
                    for (int wd = w; wd < nwords; wd += n_search) {
                        const int r = (wd << 5) + lane;
                        const uint32_t aw = s_alive[wd];
                        bool hit = false;
                        if (r < top && ((aw >> lane) & 1u)) hit = overlap_gt(tb, s_box[(uint32_t)s_keys[r]], p.thr);
                        uint32_t mw = __ballot_sync(0xffffffffu, hit);
                        if ((top >> 5) == wd) mw |= 1u << (top & 31);
                        const uint32_t left = aw & ~mw;
                        if (lane == 0) {
                            s_member[buf * (p.cap / 32) + wd] = mw;
                            s_alive[wd] = left;
                        }
                        if (left) wmax = max(wmax, (wd << 5) + 31 - __clz(left));
                    }
                }
                if (lane == 0) s_wmax[buf][w] = wmax;
            } else if (prev_top >= 0) {
                // finalise the cluster found in the previous pass (member bitmap buf ^ 1)
                const uint32_t *mem = s_member + (buf ^ 1) * (p.cap / 32);
                DetEntry v;
                v.cls = c; v.aux = 0;
                const int tpos = (int)(uint32_t)s_keys[prev_top];
                const float ptop = key_to_score((uint32_t)(s_keys[prev_top] >> 32));
                v.src = s_src[tpos];
                if (!p.average) {
                    const int4 b = s_box[tpos];
                    v.prob = ptop; v.x1 = b.x; v.y1 = b.y; v.x2 = b.z; v.y2 = b.w;
                } else {
                    // ordered member list: lane l owns a contiguous run of words
                    const int per = (nwords + 31) >> 5;
                    int cnt = 0;
                    for (int q = 0; q < per; ++q) {
                        const int wd = lane * per + q;
                        if (wd < nwords) cnt += __popc(mem[wd]);
                    }
                    int inc = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, inc, d);
                        if (lane >= d) inc += t;
                    }
                    const int n_mem = __shfl_sync(0xffffffffu, inc, 31);
                    int o = inc - cnt;
                    for (int q = 0; q < per; ++q) {
                        const int wd = lane * per + q;
                        if (wd < nwords) {
                            uint32_t bits = mem[wd];
                            while (bits) {
                                const int b = __ffs(bits) - 1;
                                bits &= bits - 1;
                                s_list[o++] = (wd << 5) + b;
                            }
                        }
                    }
                    __syncwarp();
                    // representatives (RADNet.py:226-231): the n_obj_avg best members when even the top is
                    // below the confidence threshold, else the members above it; order = ascending score
                    int n_rep = 0;
                    long long sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0;
                    const bool low = ptop < thr_c;
                    const int from = low ? max(0, n_mem - p.n_obj_avg) : 0;
                    for (int q0 = from; q0 < n_mem; q0 += 32) {
                        const int q = q0 + lane;
                        bool rep = false;
                        int pos = 0;
                        float pr = 0.f;
                        if (q < n_mem) {
                            const unsigned long long kk = s_keys[s_list[q]];
                            pos = (int)(uint32_t)kk;
                            pr = key_to_score((uint32_t)(kk >> 32));
                            rep = low || pr > thr_c;
                        }
                        const uint32_t rm = __ballot_sync(0xffffffffu, rep);
                        if (rep) {
                            s_vals[n_rep + __popc(rm & lanemask_lt())] = pr;
                            const int4 b = s_box[pos];
                            sx1 += b.x; sy1 += b.y; sx2 += b.z; sy2 += b.w;
                        }
                        n_rep += __popc(rm);
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        sx1 += __shfl_xor_sync(0xffffffffu, sx1, d);
                        sy1 += __shfl_xor_sync(0xffffffffu, sy1, d);
                        sx2 += __shfl_xor_sync(0xffffffffu, sx2, d);
                        sy2 += __shfl_xor_sync(0xffffffffu, sy2, d);
                    }
                    __syncwarp();
                    v.aux = n_rep;
                    if (n_rep > 0) {
                        const double dn = (double)n_rep;
                        v.x1 = __double2int_rn(rint(__ddiv_rn((double)sx1, dn)));     // RADNet.py:238
                        v.y1 = __double2int_rn(rint(__ddiv_rn((double)sy1, dn)));
                        v.x2 = __double2int_rn(rint(__ddiv_rn((double)sx2, dn)));
                        v.y2 = __double2int_rn(rint(__ddiv_rn((double)sy2, dn)));
                        float s = 0.f;
                        if (lane == 0) s = pairwise_sum_f32(s_vals, n_rep);
                        // float32 sum / intp count is a float64 divide rounded to float32 (numpy _mean)
                        v.prob = (float)__ddiv_rn((double)s, dn);                    // RADNet.py:239
                    } else {
                        // mean of an empty slice (top score == threshold exactly): NaN, as NumPy
                        v.x1 = v.y1 = v.x2 = v.y2 = (int)0x80000000;
                        v.prob = __int_as_float(0x7fc00000);
                        if (lane == 0) atomicAdd(&s_stat[C_EMPTY], 1);
                    }
                }
                if (lane == 0 && it - 1 < p.cap) cout[it - 1] = v;
            }
            __syncthreads();
            if (top < 0) break;
            prev_top = top;
            // next top = highest rank still alive
            int nt = (lane < n_search) ? s_wmax[buf][lane] : -1;
            nt = __reduce_max_sync(0xffffffffu, nt);
            top = nt;
            ++it;
            if (!p.average && it >= p.max_boxes) top = -1;        // rpn.py:449-450: stop after max_boxes picks
        }
        n_out = it;
    }

    // ---- 5. per-class result -> scratch; the last CTA of the segment packs the record ----------
    __syncthreads();
    if (threadIdx.x == 0) {
        chdr[C_COUNT] = overflow ? -2 : n_out;
        chdr[C_FIRST] = s_stat[C_FIRST];
        chdr[C_TIES] = s_stat[C_TIES];
        chdr[C_DEGEN] = s_stat[C_DEGEN];
        chdr[C_EMPTY] = s_stat[C_EMPTY];
        chdr[C_NIN] = m;
        chdr[C_FAULT] = s_stat[C_FAULT];
        chdr[C_RANGE] = s_stat[C_RANGE];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(seg_counter, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // pack: classes in first-appearance order
    __shared__ int s_cnt[kMaxClasses], s_fst[kMaxClasses], s_base[kMaxClasses], s_ord[kMaxClasses];
    __shared__ int s_hdr[8];
    if (threadIdx.x < 8) s_hdr[threadIdx.x] = 0;
    if (threadIdx.x < kMaxClasses) {
        s_cnt[threadIdx.x] = 0;
        s_fst[threadIdx.x] = 0x7fffffff;
        s_ord[threadIdx.x] = -1;
    }
    __syncthreads();
    if ((int)threadIdx.x < p.n_cls) {
        const volatile int32_t *h = reinterpret_cast<const volatile int32_t *>(ws_seg + (size_t)threadIdx.x * p.ws_cls_stride);
        const int cnt = h[C_COUNT];
        if (cnt == -2) atomicExch(&s_hdr[H_NDET], -2);
        if (h[C_FAULT]) atomicExch(&s_hdr[5], 1);
        s_cnt[threadIdx.x] = max(cnt, 0);
        s_fst[threadIdx.x] = h[C_NIN] > 0 ? h[C_FIRST] : 0x7fffffff;
        atomicAdd(&s_hdr[H_NDEGEN], h[C_DEGEN]);
        atomicAdd(&s_hdr[H_NTIES], h[C_TIES]);
        atomicAdd(&s_hdr[6], h[C_EMPTY]);
        atomicAdd(&s_hdr[H_NRANGE], h[C_RANGE]);
        if (h[C_NIN] > 0) atomicAdd(&s_hdr[H_NCLASSES], 1);
    }
    __syncthreads();
    if ((int)threadIdx.x < p.n_cls && s_fst[threadIdx.x] != 0x7fffffff) {
        int rank = 0, b = 0;
        for (int o = 0; o < p.n_cls; ++o)
            if (s_fst[o] < s_fst[threadIdx.x]) { ++rank; b += s_cnt[o]; }
        s_ord[rank] = threadIdx.x;
        s_base[threadIdx.x] = b;
    }
    __syncthreads();
    unsigned char *out = p.rec_out + (size_t)seg * p.out_stride;
    DetEntry *e = rec_entries(out);
    int n_total = 0;
    for (int o = 0; o < p.n_cls; ++o) n_total += s_cnt[o];
    const bool bad = s_hdr[H_NDET] == -2 || s_hdr[5];
    const int n_emit = bad ? 0 : min(n_total, p.out_max_det);
    const bool xf = p.ratio != nullptr;
    const double ratio = xf ? p.ratio[seg] : 1.0;
    const int ox = p.origin ? p.origin[2 * seg] : 0, oy = p.origin ? p.origin[2 * seg + 1] : 0;
    for (int cc = 0; cc < p.n_cls; ++cc) {
        if (s_fst[cc] == 0x7fffffff || bad) continue;
        const int4 *src = reinterpret_cast<const int4 *>(ws_seg + (size_t)cc * p.ws_cls_stride + 32);
        for (int i = threadIdx.x; i < s_cnt[cc]; i += kDetThreads) {
            const int o = s_base[cc] + i;
            if (o >= p.out_max_det) continue;
            // written by another CTA of this launch: read through L2
            const int4 lo4 = __ldcg(src + 2 * i), hi4 = __ldcg(src + 2 * i + 1);
            DetEntry v;
            v.cls = lo4.x; v.prob = __int_as_float(lo4.y); v.x1 = lo4.z; v.y1 = lo4.w;
            v.x2 = hi4.x; v.y2 = hi4.y; v.src = hi4.z; v.aux = hi4.w;
            if (xf) {
                v.x1 = real_coordinate(v.x1, ratio); v.y1 = real_coordinate(v.y1, ratio);
                v.x2 = real_coordinate(v.x2, ratio); v.y2 = real_coordinate(v.y2, ratio);
            }
            v.x1 += ox; v.y1 += oy; v.x2 += ox; v.y2 += oy;
            e[o] = v;
        }
    }
    for (int i = n_emit + threadIdx.x; i < p.out_max_det; i += kDetThreads) {
        DetEntry v;
        v.cls = 0; v.prob = 0.f; v.x1 = v.y1 = v.x2 = v.y2 = 0; v.src = 0; v.aux = 0;
        e[i] = v;
    }
    if (threadIdx.x < kMaxClasses) {
        rec_order(out)[threadIdx.x] = bad ? -1 : s_ord[threadIdx.x];
        rec_count(out)[threadIdx.x] = bad ? 0 : s_cnt[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        int32_t *h = rec_hdr(out);
        h[H_NDET] = s_hdr[5] ? -1 : ((s_hdr[H_NDET] == -2 || n_total > p.out_max_det) ? -2 : n_emit);
        h[H_NIN] = total;
        h[H_NDEGEN] = s_hdr[H_NDEGEN];
        h[H_NTIES] = s_hdr[H_NTIES];
        h[H_NCLASSES] = s_hdr[H_NCLASSES];
        h[H_NFALLBACK] = s_hdr[6];          // K7: clusters whose representative set was empty
        h[H_NNEARTIE] = 0;
        h[H_NRANGE] = s_hdr[H_NRANGE];
        *seg_counter = 0;                   // leave the workspace ready for the next launch
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
__global__ void real_coordinates_kernel(const int32_t *__restrict__ v, long long n, double ratio,
                                        int32_t *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = real_coordinate(v[i], ratio);
}

static int optin_smem() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 232448;
    const int v = device_smem_optin(dev);
    return v > 0 ? v : 232448;
}

static size_t matrix_smem_bytes(int cap, int words) {
    return (size_t)cap * (16 + 4 * 5) + (size_t)cap * words * 4;
}

static int round_pow2(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

static size_t cluster_smem_bytes(int cap, int n_in) {
    return (size_t)cap * (8 + 16 + 4 + 4 + 4) + (size_t)(cap / 32) * 4 * 3 + (size_t)(n_in + 1) * 4 + 16;
}

template <int kMode>
static int launch_class_nms(ClassNmsParams &p, int S, int cap, cudaStream_t st) {
    p.cap = cap < 32 ? 32 : cap;
    p.words = (kMode == kModeDecode) ? 0 : (p.cap + 31) / 32;      // decode only: no suppression matrix
    const size_t smem = matrix_smem_bytes(p.cap, p.words);
    if (smem > (size_t)optin_smem()) {
        set_error("class_nms: %d entries need %zu B of shared memory", p.cap, smem);
        return RADNET_E_UNSUPPORTED;
    }
    {
        int dev = 0;
        RADNET_CUDA(cudaGetDevice(&dev));
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(class_nms_kernel<kMode>), dev, smem)) return rc;
    }
    class_nms_kernel<kMode><<<S, kDetThreads, smem, st>>>(p);
    return check_launch("class_nms_kernel");
}

static int fill_head(ClassNmsParams &p, const float *p_cls, const float *p_regr, int B, int R, int n_cls,
                     const void *det, int det_max_boxes, const int32_t *rois, const int32_t *roi_count,
                     double bbox_threshold, const double *h_regr_std4, int rpn_stride, void *rec_out,
                     int rec_max_det, const char *who) {
    RADNET_CHECK_ARG(p_cls && p_regr && rec_out && h_regr_std4, "%s: null pointer", who);
    RADNET_CHECK_ARG((det != nullptr) != (rois != nullptr), "%s: give either detection records or an RoI array", who);
    RADNET_CHECK_ARG(B >= 1 && R >= 1 && R <= kMatrixCap && n_cls >= 2 && n_cls <= kMaxClasses,
                     "%s: bad sizes B=%d R=%d (<=%d) n_cls=%d (<=%d)", who, B, R, kMatrixCap, n_cls, kMaxClasses);
    RADNET_CHECK_ARG(rec_max_det >= 1 && rpn_stride >= 1 && rpn_stride <= 4096, "%s: bad rec_max_det=%d or rpn_stride=%d",
                     who, rec_max_det, rpn_stride);
    RADNET_CHECK_ARG(!det || det_max_boxes >= 1, "%s: det_max_boxes=%d", who, det_max_boxes);
    for (int k = 0; k < 4; ++k) {
        RADNET_CHECK_ARG((float)h_regr_std4[k] != 0.f, "%s: classifier_regr_std[%d] == 0", who, k);
        p.std4[k] = (float)h_regr_std4[k];
    }
    p.p_cls = p_cls; p.p_regr = p_regr; p.R = R; p.n_cls = n_cls;
    p.det = reinterpret_cast<const unsigned char *>(det);
    p.det_stride = det ? radnet_det_record_bytes(det_max_boxes) : 0;
    p.det_max_boxes = det_max_boxes;
    p.rois = rois; p.roi_count = roi_count;
    p.bbox_thr = (float)bbox_threshold;
    p.stride = rpn_stride;
    p.rec_out = reinterpret_cast<unsigned char *>(rec_out);
    p.out_stride = radnet_cls_record_bytes(rec_max_det);
    p.out_max_det = rec_max_det;
    return RADNET_OK;
}

static int launch_cluster(const void *rec_in, int in_max_det, int S, int n_in, const int32_t *in_count, int n_cls,
                          int average, double thr, double conf_thr, int n_obj_avg, int max_boxes,
                          const double *ratio, const int32_t *origin, void *rec_out, int out_max_det, void *ws,
                          size_t ws_bytes, cudaStream_t st, const char *who, int skip_if_done = 0) {
    long long total_cap = (long long)n_in * in_max_det;
    int cap = round_pow2((int)(total_cap < kClusterCap ? total_cap : kClusterCap));
    const size_t need = radnet_final_nms_workspace_bytes(S, n_in, in_max_det, n_cls);
    if (ws_bytes < need) {
        set_error("%s: workspace %zu < %zu", who, ws_bytes, need);
        return RADNET_E_WORKSPACE;
    }
    ClusterParams p{};
    p.rec_in = reinterpret_cast<const unsigned char *>(rec_in);
    p.in_stride = radnet_cls_record_bytes(in_max_det);
    p.n_in = n_in; p.in_count = in_count; p.n_cls = n_cls; p.average = average;
    p.thr = thr; p.conf_thr = (float)conf_thr; p.n_obj_avg = n_obj_avg; p.max_boxes = max_boxes;
    p.ratio = ratio; p.origin = origin;
    p.rec_out = reinterpret_cast<unsigned char *>(rec_out);
    p.out_stride = radnet_cls_record_bytes(out_max_det);
    p.out_max_det = out_max_det;
    p.ws = reinterpret_cast<unsigned char *>(ws);
    p.cap = cap;
    p.skip_if_done = skip_if_done;
    p.ws_counters = align_up((size_t)S * 4, 256);
    p.ws_cls_stride = 32 + (size_t)cap * sizeof(DetEntry);
    p.ws_seg_stride = (size_t)n_cls * p.ws_cls_stride;
    const size_t smem = cluster_smem_bytes(cap, n_in);
    if (smem > (size_t)optin_smem()) {
        set_error("%s: %zu B of shared memory needed", who, smem);
        return RADNET_E_UNSUPPORTED;
    }
    {
        int dev = 0;
        RADNET_CUDA(cudaGetDevice(&dev));
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(cluster_kernel), dev, smem)) return rc;
    }
    // segment counters start at zero (the packing CTA also resets its own)
    RADNET_CUDA(cudaMemsetAsync(p.ws, 0, p.ws_counters, st));
    cluster_kernel<<<dim3(n_cls, S), kDetThreads, smem, st>>>(p);
    return check_launch("cluster_kernel");
}

}  // namespace radnet

using namespace radnet;

extern "C" size_t radnet_cls_record_bytes(int max_det) {
    if (max_det < 0) return 0;
    return (size_t)kRecHeader + (size_t)max_det * sizeof(DetEntry);
}

extern "C" int radnet_classify_decode(const float *p_cls, const float *p_regr, int B, int R, int n_cls,
                                      const void *det, int det_max_boxes, const int32_t *rois,
                                      const int32_t *roi_count, double bbox_threshold,
                                      const double *h_regr_std4, int rpn_stride, void *rec_out,
                                      int rec_max_det, void *stream) {
    ClassNmsParams p{};
    int rc = fill_head(p, p_cls, p_regr, B, R, n_cls, det, det_max_boxes, rois, roi_count, bbox_threshold,
                       h_regr_std4, rpn_stride, rec_out, rec_max_det, "classify_decode");
    if (rc) return rc;
    return launch_class_nms<kModeDecode>(p, B, R, (cudaStream_t)stream);
}

extern "C" int radnet_classify_nms(const float *p_cls, const float *p_regr, int B, int R, int n_cls,
                                   const void *det, int det_max_boxes, const int32_t *rois,
                                   const int32_t *roi_count, double bbox_threshold,
                                   const double *h_regr_std4, int rpn_stride, double nms_thr, int max_boxes,
                                   const double *ratio, const int32_t *origin, void *rec_out,
                                   int rec_max_det, void *stream) {
    ClassNmsParams p{};
    int rc = fill_head(p, p_cls, p_regr, B, R, n_cls, det, det_max_boxes, rois, roi_count, bbox_threshold,
                       h_regr_std4, rpn_stride, rec_out, rec_max_det, "classify_nms");
    if (rc) return rc;
    RADNET_CHECK_ARG(max_boxes >= 1, "classify_nms: max_boxes=%d", max_boxes);
    p.thr = nms_thr; p.max_boxes = max_boxes; p.ratio = ratio; p.origin = origin;
    return launch_class_nms<kModeHeadNms>(p, B, R, (cudaStream_t)stream);
}

extern "C" size_t radnet_final_nms_workspace_bytes(int S, int n_in, int in_max_det, int n_cls) {
    if (S < 1 || n_in < 1 || in_max_det < 1 || n_cls < 1) return 0;
    long long total_cap = (long long)n_in * in_max_det;
    int cap = round_pow2((int)(total_cap < kClusterCap ? total_cap : kClusterCap));
    return align_up((size_t)S * 4, 256) + (size_t)S * n_cls * (32 + (size_t)cap * sizeof(DetEntry));
}

extern "C" size_t radnet_class_nms_workspace_bytes(int S, int n_in, int in_max_det, int n_cls) {
    if ((long long)n_in * in_max_det <= kMatrixCap) return 16;
    return radnet_final_nms_workspace_bytes(S, n_in, in_max_det, n_cls);
}

extern "C" int radnet_class_nms(const void *rec_in, int in_max_det, int S, int n_in, const int32_t *in_count,
                                int n_cls, double thr, int max_boxes, const double *ratio,
                                const int32_t *origin, void *rec_out, int out_max_det, void *ws,
                                size_t ws_bytes, void *stream) {
    RADNET_CHECK_ARG(rec_in && rec_out, "class_nms: null pointer");
    RADNET_CHECK_ARG(S >= 1 && n_in >= 1 && in_max_det >= 1 && out_max_det >= 1 && max_boxes >= 1 && n_cls >= 1 &&
                         n_cls <= kMaxClasses,
                     "class_nms: bad sizes S=%d n_in=%d in_max_det=%d out_max_det=%d max_boxes=%d n_cls=%d", S, n_in,
                     in_max_det, out_max_det, max_boxes, n_cls);
    if ((long long)n_in * in_max_det <= kMatrixCap) {
        ClassNmsParams p{};
        p.rec_in = reinterpret_cast<const unsigned char *>(rec_in);
        p.in_stride = radnet_cls_record_bytes(in_max_det);
        p.n_in = n_in; p.in_count = in_count; p.n_cls = n_cls;
        p.thr = thr; p.max_boxes = max_boxes; p.ratio = ratio; p.origin = origin;
        p.rec_out = reinterpret_cast<unsigned char *>(rec_out);
        p.out_stride = radnet_cls_record_bytes(out_max_det);
        p.out_max_det = out_max_det;
        return launch_class_nms<kModeRecNms>(p, S, n_in * in_max_det, (cudaStream_t)stream);
    }
    RADNET_CHECK_ARG(ws, "class_nms: null workspace");
    // The records can hold more than the bit-matrix kernel takes, but they rarely do: it goes first with its
    // maximum capacity and marks the segments it could not hold (-2); the iterative kernel then only works on those.
    int skip = 0;
    if (n_in <= kMatrixCap) {
        ClassNmsParams p{};
        p.rec_in = reinterpret_cast<const unsigned char *>(rec_in);
        p.in_stride = radnet_cls_record_bytes(in_max_det);
        p.n_in = n_in; p.in_count = in_count; p.n_cls = n_cls;
        p.thr = thr; p.max_boxes = max_boxes; p.ratio = ratio; p.origin = origin;
        p.rec_out = reinterpret_cast<unsigned char *>(rec_out);
        p.out_stride = radnet_cls_record_bytes(out_max_det);
        p.out_max_det = out_max_det;
        const int rc = launch_class_nms<kModeRecNms>(p, S, kMatrixCap, (cudaStream_t)stream);
        if (rc) return rc;
        skip = 1;
    }
    return launch_cluster(rec_in, in_max_det, S, n_in, in_count, n_cls, 0, thr, 0.0, 0, max_boxes, ratio, origin,
                          rec_out, out_max_det, ws, ws_bytes, (cudaStream_t)stream, "class_nms", skip);
}

extern "C" int radnet_final_nms(const void *rec_in, int in_max_det, int S, int n_in, const int32_t *in_count,
                                int n_cls, double avg_thr, double conf_thr, int n_obj_avg, void *rec_out,
                                int out_max_det, void *ws, size_t ws_bytes, void *stream) {
    RADNET_CHECK_ARG(rec_in && rec_out && ws, "final_nms: null pointer");
    RADNET_CHECK_ARG(S >= 1 && S <= 65535 && n_in >= 1 && in_max_det >= 1 && out_max_det >= 1 && n_obj_avg >= 1 &&
                         n_cls >= 1 && n_cls <= kMaxClasses,
                     "final_nms: bad sizes S=%d n_in=%d in_max_det=%d out_max_det=%d n_obj_avg=%d n_cls=%d", S, n_in,
                     in_max_det, out_max_det, n_obj_avg, n_cls);
    return launch_cluster(rec_in, in_max_det, S, n_in, in_count, n_cls, 1, avg_thr, conf_thr, n_obj_avg, 0x7fffffff,
                          nullptr, nullptr, rec_out, out_max_det, ws, ws_bytes, (cudaStream_t)stream, "final_nms");
}

extern "C" int radnet_real_coordinates(const int32_t *v, long long n, double ratio, int32_t *out, void *stream) {
    RADNET_CHECK_ARG(v && out && n >= 0, "real_coordinates: bad arguments");
    RADNET_CHECK_ARG(ratio > 0.0 && ratio < 1e300, "real_coordinates: ratio must be positive and finite");
    if (n == 0) return RADNET_OK;
    const long long blocks = (n + 255) / 256;
    RADNET_CHECK_ARG(blocks <= 0x7fffffffLL, "real_coordinates: n too large");
    real_coordinates_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(v, n, ratio, out);
    return check_launch("real_coordinates_kernel");
}


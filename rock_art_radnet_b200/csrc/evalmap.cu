// evalmap.cu - f4: the mAP evaluation of the reference (test.py:48-173).
//
//   radnet_match_detections   get_objects (test.py:48-113): detections in descending score order (ties: higher index
//                             first = the reversed stable ascending argsort) are matched greedily to the FIRST not yet
//                             matched figure of the same class with utils.iou >= threshold.  Like the reference - which
//                             concatenates the detections and figures of the whole test set before matching
//                             (test.py:221-227) - there is one pool of figures, not one per image.
//                             The O(n_det * n_gt) IoU tests run in parallel into a bit matrix (CTA per detection);
//                             only the greedy walk over it is sequential (one warp: candidates AND NOT matched,
//                             lowest set bit).
//   radnet_class_ap           calc_class_ap (test.py:117-173): sort by score, running TP / FP, precision / recall,
//                             precision envelope from the right, AP = sum p[i+1] * (r[i+1] - r[i]).
//
// Sorting uses cub::DeviceRadixSort (stable, ascending on an order-preserving image of the float64 score), read from
// the end - a library call for a plain sort; everything else is hand-written.
#include <cub/device/device_radix_sort.cuh>

#include "iou.cuh"

namespace radnet {

__global__ void score_keys_kernel(const double *score, int n, uint64_t *keys, int32_t *idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // ascending order of the keys = ascending order of the scores (NaN last, like np.argsort)
    keys[i] = score_to_key64(score[i]);
    idx[i] = i;
}

// visiting order: rank r -> detection order[n-1-r]
struct MatchParams {
    const double *det_box;     // [n_det][4] x1,y1,x2,y2
    const int32_t *det_cls;    // [n_det]
    const double *gt_box;      // [n_gt][4]
    const int32_t *gt_cls;     // [n_gt]
    int n_det, n_gt, words;
    double thr;
    const int32_t *sorted_idx; // [n_det] ascending by score (stable)
    uint32_t *cand;            // [n_det][words] bit g of row r: figure g can be matched by the r-th visited detection
    int32_t *visit;            // [n_det] detection index visited at rank r
    int32_t *match;            // [n_det] matched figure or -1, by rank
    uint32_t *gt_matched;      // [words]
};

__global__ void __launch_bounds__(256) match_candidates_kernel(MatchParams p) {
    const int r = blockIdx.x;
    const int d = p.sorted_idx[p.n_det - 1 - r];
    if (threadIdx.x == 0) p.visit[r] = d;
    const double x1 = p.det_box[4 * d], y1 = p.det_box[4 * d + 1], x2 = p.det_box[4 * d + 2], y2 = p.det_box[4 * d + 3];
    const int c = p.det_cls[d];
    for (int g0 = 0; g0 < p.words * 32; g0 += 256) {
        const int g = g0 + threadIdx.x;
        bool ok = false;
        if (g < p.n_gt && p.gt_cls[g] == c) {
            const double *q = p.gt_box + 4 * g;
            ok = ref_iou(x1, y1, x2, y2, q[0], q[1], q[2], q[3]) >= p.thr;          // test.py:89-91
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0 && g0 / 32 + (threadIdx.x >> 5) < p.words)
            p.cand[(size_t)r * p.words + g0 / 32 + (threadIdx.x >> 5)] = m;
    }
}

// one warp: the sequential part of get_objects
__global__ void __launch_bounds__(32) match_resolve_kernel(MatchParams p) {
    const int lane = threadIdx.x;
    for (int w = lane; w < p.words; w += 32) p.gt_matched[w] = 0u;
    __syncwarp();
    for (int r = 0; r < p.n_det; ++r) {
        int found = -1;
        for (int w0 = 0; w0 < p.words && found < 0; w0 += 32) {
            const int w = w0 + lane;
            const uint32_t free_cand = w < p.words ? (p.cand[(size_t)r * p.words + w] & ~p.gt_matched[w]) : 0u;
            const unsigned any = __ballot_sync(0xffffffffu, free_cand != 0u);
            if (any) {
                const int src = __ffs(any) - 1;                       // lowest word = lowest figure index
                const uint32_t bits = __shfl_sync(0xffffffffu, free_cand, src);
                const int bit = __ffs(bits) - 1;
                found = (w0 + src) * 32 + bit;
                if (lane == src) p.gt_matched[w] |= 1u << bit;
            }
        }
        __syncwarp();
        if (lane == 0) p.match[r] = found;
    }
}

// ---- calc_class_ap: one CTA, tiles of 1024 entries -------------------------------------------------------------------
struct ApParams {
    const int32_t *y_true;     // [n]
    const double *y_pred;      // [n]
    const int32_t *sorted_idx; // [n] ascending by score (stable)
    int n;
    double *precision, *recall, *iprec, *irec;    // [n] each
    double *ap;                // [1]
};

__global__ void __launch_bounds__(1024) class_ap_kernel(ApParams p) {
    __shared__ int s_w[2][33];
    __shared__ int s_carry[2];
    __shared__ long long s_ngt;
    __shared__ double s_red[32];
    __shared__ double s_max;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = p.n;
    // n_gt = sum(y_true)
    long long loc = 0;
    for (int i = threadIdx.x; i < n; i += 1024) loc += p.y_true[i];
    for (int d = 16; d > 0; d >>= 1) loc += __shfl_down_sync(0xffffffffu, loc, d);
    if (threadIdx.x == 0) { s_ngt = 0; s_carry[0] = 0; s_carry[1] = 0; }
    __syncthreads();
    if (lane == 0) atomicAdd((unsigned long long *)&s_ngt, (unsigned long long)loc);
    __syncthreads();
    const double n_gt = (double)s_ngt;
    // running TP / FP in visiting order (rank r -> entry sorted_idx[n-1-r]), precision and recall (test.py:131-147)
    for (int r0 = 0; r0 < n; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        int tp = 0, fp = 0;
        if (r < n) {
            const int e = p.sorted_idx[n - 1 - r];
            const int t = p.y_true[e];
            const double s = p.y_pred[e];
            tp = (t > 0 && s > 0.0) ? 1 : 0;
            fp = (t == 0 && s > 0.0) ? 1 : 0;
        }
        int itp = tp, ifp = fp;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, itp, d), b = __shfl_up_sync(0xffffffffu, ifp, d);
            if (lane >= d) { itp += a; ifp += b; }
        }
        if (lane == 31) { s_w[0][w] = itp; s_w[1][w] = ifp; }
        __syncthreads();
        if (w < 2) {
            const int v = s_w[w][lane];
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int a = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += a;
            }
            s_w[w][lane] = inc - v;
            if (lane == 31) s_w[w][32] = inc;
        }
        __syncthreads();
        if (r < n) {
            const int ctp = s_carry[0] + s_w[0][w] + itp, cfp = s_carry[1] + s_w[1][w] + ifp;
            p.precision[r] = (ctp + cfp == 0) ? 0.0 : (double)ctp / (double)(ctp + cfp);
            p.recall[r] = n_gt != 0.0 ? (double)ctp / n_gt : 0.0;
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] += s_w[0][32]; s_carry[1] += s_w[1][32]; }
        __syncthreads();
    }
    // precision envelope from the right (test.py:154-163): running maximum, tiles walked backwards
    if (threadIdx.x == 0) s_max = 0.0;
    __syncthreads();
    for (int hi = n; hi > 0; hi -= 1024) {
        const int r = hi - 1 - threadIdx.x;                // thread 0 takes the rightmost entry of the tile
        double v = r >= 0 ? p.precision[r] : 0.0;
        // inclusive max-scan over the tile in thread order
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double a = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v = fmax(v, a);
        }
        if (lane == 31) s_red[w] = v;
        __syncthreads();
        double pre = s_max;
        for (int k = 0; k < w; ++k) pre = fmax(pre, s_red[k]);
        v = fmax(v, pre);
        if (r >= 0) { p.iprec[r] = v; p.irec[r] = p.recall[r]; }
        __syncthreads();
        if (threadIdx.x == 1023) s_max = v;
        __syncthreads();
    }
    // AP = sum_i iprec[i+1] * (irec[i+1] - irec[i])   (test.py:167-169; the reference adds left to right)
    double acc = 0.0;
    for (int i = threadIdx.x; i + 1 < n; i += 1024) acc += p.iprec[i + 1] * (p.irec[i + 1] - p.irec[i]);
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    __syncthreads();
    if (lane == 0) s_red[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 32; ++k) t += s_red[k];
        *p.ap = t;
    }
}

static size_t sort_temp_bytes(int n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, n);
    return bytes;
}

// stable ascending argsort of n float64 scores into ws: returns the sorted index array
static int argsort_scores(const double *score, int n, unsigned char *ws, size_t ws_bytes, cudaStream_t st,
                          int32_t **sorted_idx, size_t *used) {
    const size_t nn = align_up((size_t)n, 64);
    uint64_t *k_in = reinterpret_cast<uint64_t *>(ws), *k_out = k_in + nn;
    int32_t *i_in = reinterpret_cast<int32_t *>(k_out + nn), *i_out = i_in + nn;
    unsigned char *temp = reinterpret_cast<unsigned char *>(i_out + nn);
    size_t temp_bytes = sort_temp_bytes(n);
    *used = (size_t)(temp - ws) + align_up(temp_bytes, 256);
    if (*used > ws_bytes) {
        set_error("evaluation: workspace %zu < %zu", ws_bytes, *used);
        return RADNET_E_WORKSPACE;
    }
    score_keys_kernel<<<(n + 255) / 256, 256, 0, st>>>(score, n, k_in, i_in);
    RADNET_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k_in, k_out, i_in, i_out, n, 0, 64, st));
    *sorted_idx = i_out;
    return RADNET_OK;
}

static size_t argsort_bytes(int n) {
    const size_t nn = align_up((size_t)(n > 0 ? n : 1), 64);
    return nn * (8 + 8 + 4 + 4) + align_up(sort_temp_bytes(n > 0 ? n : 1), 256) + 256;
}

}  // namespace radnet

using namespace radnet;

extern "C" size_t radnet_match_detections_workspace_bytes(int n_det, int n_gt) {
    if (n_det < 0 || n_gt < 0) return 0;
    const size_t words = ((size_t)(n_gt > 0 ? n_gt : 1) + 31) / 32;
    return argsort_bytes(n_det) + align_up((size_t)(n_det > 0 ? n_det : 1) * words * 4, 256) + align_up(words * 4, 256);
}

extern "C" int radnet_match_detections(const double *det_box, const int32_t *det_cls, const double *det_prob, int n_det,
                                       const double *gt_box, const int32_t *gt_cls, int n_gt, double thr,
                                       int32_t *visit, int32_t *match, void *ws, size_t ws_bytes, void *stream) {
    RADNET_CHECK_ARG(n_det >= 0 && n_gt >= 0 && ws, "match_detections: bad arguments");
    if (n_det == 0) return RADNET_OK;
    RADNET_CHECK_ARG(det_box && det_cls && det_prob && visit && match && (n_gt == 0 || (gt_box && gt_cls)),
                     "match_detections: null pointer");
    if (ws_bytes < radnet_match_detections_workspace_bytes(n_det, n_gt)) {
        set_error("match_detections: workspace %zu < %zu", ws_bytes, radnet_match_detections_workspace_bytes(n_det, n_gt));
        return RADNET_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *w8 = reinterpret_cast<unsigned char *>(ws);
    int32_t *sorted_idx = nullptr;
    size_t used = 0;
    if (int rc = argsort_scores(det_prob, n_det, w8, ws_bytes, st, &sorted_idx, &used)) return rc;
    MatchParams p{};
    p.det_box = det_box; p.det_cls = det_cls; p.gt_box = gt_box; p.gt_cls = gt_cls;
    p.n_det = n_det; p.n_gt = n_gt; p.words = (int)(((size_t)(n_gt > 0 ? n_gt : 1) + 31) / 32); p.thr = thr;
    p.sorted_idx = sorted_idx;
    p.cand = reinterpret_cast<uint32_t *>(w8 + argsort_bytes(n_det));
    p.gt_matched = reinterpret_cast<uint32_t *>(w8 + argsort_bytes(n_det) + align_up((size_t)n_det * p.words * 4, 256));
    p.visit = visit; p.match = match;
    match_candidates_kernel<<<n_det, 256, 0, st>>>(p);
    if (int rc = check_launch("match_candidates_kernel")) return rc;
    match_resolve_kernel<<<1, 32, 0, st>>>(p);
    return check_launch("match_resolve_kernel");
}

extern "C" size_t radnet_class_ap_workspace_bytes(int n) { return n < 0 ? 0 : argsort_bytes(n); }

extern "C" int radnet_class_ap(const int32_t *y_true, const double *y_pred, int n, double *precision, double *recall,
                               double *interp_precision, double *interp_recall, double *ap, void *ws,
                               size_t ws_bytes, void *stream) {
    RADNET_CHECK_ARG(n >= 0 && ap && ws, "class_ap: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        RADNET_CUDA(cudaMemsetAsync(ap, 0, sizeof(double), st));
        return RADNET_OK;
    }
    RADNET_CHECK_ARG(y_true && y_pred && precision && recall && interp_precision && interp_recall, "class_ap: null pointer");
    int32_t *sorted_idx = nullptr;
    size_t used = 0;
    if (int rc = argsort_scores(y_pred, n, reinterpret_cast<unsigned char *>(ws), ws_bytes, st, &sorted_idx, &used)) return rc;
    ApParams p{};
    p.y_true = y_true; p.y_pred = y_pred; p.sorted_idx = sorted_idx; p.n = n;
    p.precision = precision; p.recall = recall; p.iprec = interp_precision; p.irec = interp_recall; p.ap = ap;
    class_ap_kernel<<<1, 1024, 0, st>>>(p);
    return check_launch("class_ap_kernel");
}

// rpn_targets.cu - K3: RPN anchor target assignment (reference faster_rcnn/utils.py:554-775, 815-816;
// upstream name calc_rpn), batched over panels, ONE launch.
//
// K3 is ONE persistent launch of one CTA (1024 threads) per SM, and the SMs take two roles:
//   * fill CTAs (the lower block indices).  The regression tensor is zero and the label tensor holds only the
//     "anchor lies inside the image" flags everywhere except at the few positive anchors, so 10*A*H*W*8 bytes per
//     panel are known at kernel start.  Panels are filled in groups of one compute round; inside a group every fill
//     CTA streams its contiguous share with 16-byte stores from the first microsecond, and per panel segment
//     thread 0 alone fences and adds the segment's size to the panel's counter (release) while the other warps
//     already store the next segment;
//   * compute CTAs: one panel at a time, all anchor shapes against the figures of the panel in shared memory -
//     exact float64 IoU only where it can matter (see below); the positives and their regression targets are
//     parked in shared memory.  When the panel's counter says it is completely filled (acquire) they are stored
//     on top, a positive is forced for figures without one, and best_anchor / n_hits are written.
// Why roles per SM: global stores that wait for L2 credits stall the SM's whole load/store pipe, shared-memory
// traffic included - with fill warps and compute warps inside one CTA every compute phase ran 2-3x longer
// (first form of this round).  Why not fewer compute SMs: a bare fill of 66.6 MB needs >= ~110 SMs storing to
// reach the 4.06 TB/s such a short burst gets on B200 (tools/exp/fill_bw.cu, profiles/r02_fill_bw.log), but a
// panel's compute is a 12-15 us chain of short phases; forms with two 512-thread compute CTAs per SM on 22 % of
// the SMs and pull-based fill units measured 36-41 us for 64 panels against 27.6 us for this one (133-145 us
// against 160 us for 512 panels).
// Every counter in the workspace is left at zero by the launch that used it: no memsets, no second kernel.
// Both output layouts (reference channel-first; NHWC with the regression half scaled by std_scaling,
// utils.py:475-478) only differ in address arithmetic.
//
// Order-dependent reference semantics and how they are kept without a serial loop:
//   * best anchor per GT = first anchor, in the reference's loop order
//     size -> ratio -> ix -> jy (utils.py:616-632), whose float32-rounded IoU is the
//     maximum (float32 accumulator utils.py:603; NumPy>=2 compares in float32, see
//     SURVEY.md row a3').  Realised as a 64-bit max of
//     (float32 bits of IoU) << 32 | (0xFFFFFFFF - loop_order).
//   * per-anchor best GT: strict '>' from 0.0, first GT wins ties (utils.py:710-713).
//   * forced positives (utils.py:741-766) are applied in GT order (last writer wins),
//     with the float32-rounded targets (utils.py:605,766), after the regular positives.
#include "iou.cuh"

namespace radnet {

constexpr int kTgtThreads = 1024;
constexpr int kTgtWarps = kTgtThreads / 32;

struct RpnTargetParams {
    const double *gt;          // [B][Gmax][4] x1,x2,y1,y2
    const uint8_t *gt_is_bg;   // [B][Gmax]
    const int32_t *gt_count;   // [B]
    int B, Gmax, H, W, A, n_ratios;
    AnchorTable anchors;       // pixels
    double stride;
    const double *img_wh;      // [B][2]
    double max_overlap;
    double *y_cls;             // layout 0: [B][2A][H][W]   layout 1: [B][H][W][2A]
    double *y_regr;            // layout 0: [B][8A][H][W]   layout 1: [B][H][W][8A], regr half * regr_scale
    int32_t *best_anchor;      // [B][Gmax][4]
    int32_t *n_hits;           // [B][Gmax]
    int layout;
    double regr_scale;
    // workspace (all zero between launches)
    int32_t *panel_done;       // [B] double2 items of the panel filled so far
    int n_fill_ctas;           // block indices below this fill only
    int role;                  // 0 both roles in one launch, 1 fill only, 2 compute only (fill already done)
    // shared-memory layout of a compute CTA (byte offsets)
    int sm_off_tables, sm_off_items, sm_off_hits, sm_off_hash, sm_off_win, sm_off_zero, hit_cap, hash_slots, n_items_max;
    int group;                 // panels per fill round (= number of compute CTAs)
    int zero_bytes;            // fill CTAs: size of the zeroed shared-memory buffer the bulk copies read (0 = plain stores)
    long long *stamps;         // profiling build only
};

struct TargetHit {
    double iou;
    int key;       // a*H*W + cell
    int g;
};



__device__ __forceinline__ void bar_team(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifdef RADNET_TGT_PROFILE
#define TGT_STAMP(i)                                                                                  \
    do {                                                                                              \
        if (p.stamps && threadIdx.x == 0)                                                             \
            p.stamps[(size_t)blockIdx.x * 16 + (i)] = global_ns();                                    \
    } while (0)
#else
#define TGT_STAMP(i) do { } while (0)
#endif

// element offsets of anchor (a, cell) in the two output tensors, for both layouts
struct TgtAddr {
    int layout, A, HW;
    __device__ __forceinline__ size_t cls(int ch, int cell) const {           // ch in [0, 2A)
        return layout ? (size_t)cell * (2 * A) + ch : (size_t)ch * HW + cell;
    }
    __device__ __forceinline__ size_t regr(int ch, int cell) const {          // ch in [0, 8A)
        return layout ? (size_t)cell * (8 * A) + ch : (size_t)ch * HW + cell;
    }
};

// positive anchor (a, cell) matched to figure g: the four regression targets as they are stored
// (utils.py:669-687, 736; `forced` = forced positive, float32-rounded: utils.py:605, 766), times regr_scale
__device__ __forceinline__ void positive_values(const RpnTargetParams &p, int a, int cell, const double *gt4,
                                                bool forced, double v[4]) {
    const int jy = cell / p.W, ix = cell - jy * p.W;
    const AnchorPx an = anchor_px(p.stride, ix, jy, p.anchors.wh[a][0], p.anchors.wh[a][1]);
    double t[4];
    regr_targets(an, gt4[0], gt4[1], gt4[2], gt4[3], t);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __dmul_rn(forced ? (double)(float)t[k] : t[k], p.regr_scale);
}

// overlap label, np.repeat(overlap, 4) and the regression targets of a positive anchor (utils.py:728-738,
// 815-816); a forced positive also sets the valid label (utils.py:758-759)
__device__ __forceinline__ void store_positive(const RpnTargetParams &p, double *cls_b, double *regr_b, int a,
                                               int cell, const double v[4], bool forced) {
    const TgtAddr ad{p.layout, p.A, p.H * p.W};
    if (forced) cls_b[ad.cls(a, cell)] = 1.0;
    cls_b[ad.cls(p.A + a, cell)] = 1.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        regr_b[ad.regr(4 * a + k, cell)] = 1.0;
        regr_b[ad.regr(4 * p.A + 4 * a + k, cell)] = v[k];
    }
}

// ---- fill role.  Both tensors of a panel are seen as ONE array of 5*A*H*W double2 items (label tensor
//      first, then the regression tensor).  Panels are filled in groups of `group` consecutive panels (one
//      round of the compute CTAs); inside a group fill CTA k streams the k-th contiguous share of the group's
//      items.  Per panel segment: stores, a CTA barrier, then thread 0 alone fences and adds the segment's
//      item count to the panel's counter (the pattern of a grid barrier: the fence is cumulative over the
//      writes ordered before it by the barrier) while the other warps already store the next segment. ------
// shared memory of this CTA -> global memory, tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk async-groups of this thread complete (writes performed), then visible to the generic proxy
__device__ __forceinline__ void bulk_wait_group_all() {
    asm volatile("cp.async.bulk.wait_group 0;\nfence.proxy.async;" ::: "memory");
}

__device__ void fill_segment(const RpnTargetParams &p, int b, int lo, int hi, uint8_t *s_inx, uint8_t *s_iny) {
    const int HW = p.H * p.W, AHW = p.A * HW;
    double2 *cls2 = reinterpret_cast<double2 *>(p.y_cls + (size_t)b * 2 * AHW);
    double2 *regr2 = reinterpret_cast<double2 *>(p.y_regr + (size_t)b * 8 * AHW);
    {   // regression tensor: zero wherever no anchor is positive (plain-store mode; the bulk mode never gets here)
        const double2 z = make_double2(0.0, 0.0);
        const int r_lo = max(lo, AHW) - AHW, r_hi = hi - AHW;
#pragma unroll 4
        for (int i = r_lo + threadIdx.x; i < r_hi; i += kTgtThreads) regr2[i] = z;
    }
    if (lo < AHW) {
        // label tensor: [valid | overlap]; valid = anchor inside the image on both axes (utils.py:629, 638), and
        // labels are only ever written inside the GT loop: no GT, no labels (utils.py:722-738)
        const int G = min(max(p.gt_count[b], 0), p.Gmax);
        const double img_w = p.img_wh[2 * b], img_h = p.img_wh[2 * b + 1];
        for (int i = threadIdx.x; i < p.A * (p.W + p.H); i += kTgtThreads) {
            const int c = i / (p.W + p.H), r = i - c * (p.W + p.H);
            const bool isx = r < p.W;
            const int k = isx ? r : r - p.W;
            const double side = p.anchors.wh[c][isx ? 0 : 1], lim_px = isx ? img_w : img_h;
            const double ctr = __dmul_rn(p.stride, (double)k + 0.5);
            const double v1 = __dsub_rn(ctr, __dmul_rn(side, 0.5)), v2 = __dadd_rn(ctr, __dmul_rn(side, 0.5));
            const bool ok = !(v1 < 0.0 || v2 > lim_px) && v1 < v2 && G > 0;
            (isx ? s_inx + c * p.W : s_iny + c * p.H)[k] = ok ? 1 : 0;
        }
        __syncthreads();
        const int twoA = 2 * p.A, top = min(hi, AHW);
        for (int i = lo + threadIdx.x; i < top; i += kTgtThreads) {
            double v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = 2 * i + h;
                int c, cell;
                if (p.layout) { cell = e / twoA; c = e - cell * twoA; }
                else { c = e / HW; cell = e - c * HW; }
                double val = 0.0;
                if (c < p.A) {
                    const int jy = cell / p.W, ix = cell - jy * p.W;
                    val = (s_inx[c * p.W + ix] & s_iny[c * p.H + jy]) ? 1.0 : 0.0;
                }
                v[h] = val;
            }
            cls2[i] = make_double2(v[0], v[1]);
        }
    }
    __syncthreads();                                              // all stores of the segment issued; tables free
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&p.panel_done[b], hi - lo);
    }
}

// ---- bulk mode: the zeros of the regression tensors of panels [b0, b0 + nb) are ONE array of nb * 4*A*H*W double2
//      items that ALL CTAs of the launch share out (CTA k of n takes the k-th contiguous part) and stream with TMA
//      bulk copies from a zeroed shared-memory buffer: one instruction per `zero_bytes`, nothing in the load/store
//      pipe, so compute CTAs take part at no cost to their panel.  Called by thread 0 only: first with
//      publish = false (issue + commit), later - after bulk_wait_group_all() and a fence - with publish = true. ------
__device__ void bulk_zero_share(const RpnTargetParams &p, int b0, int nb, int k, int n, const unsigned char *s_zero,
                                bool publish) {
    const long long per_panel = 4LL * p.A * p.H * p.W;
    const long long total = per_panel * nb;
    const long long share = ((total + n - 1) / n + 63) & ~63LL;
    long long lo = share * k, hi = lo + share;
    if (hi > total) hi = total;
    while (lo < hi) {
        const int b = (int)(lo / per_panel);
        const long long base = (long long)b * per_panel;
        const long long seg_hi = (hi < base + per_panel) ? hi : base + per_panel;
        if (publish) {
            atomicAdd(&p.panel_done[b0 + b], (int)(seg_hi - lo));
        } else {
            unsigned char *g = reinterpret_cast<unsigned char *>(reinterpret_cast<double2 *>(p.y_regr + (size_t)(b0 + b) * 2 * per_panel) + (lo - base));
            size_t bytes = (size_t)(seg_hi - lo) * sizeof(double2);
            while (bytes) {
                const uint32_t c = (uint32_t)(bytes < (size_t)p.zero_bytes ? bytes : (size_t)p.zero_bytes);
                bulk_s2g(g, s_zero, c);
                g += c;
                bytes -= c;
            }
        }
        lo = seg_hi;
    }
    if (!publish) bulk_commit_group();
}

__device__ void fill_role(const RpnTargetParams &p, int k, int n_fill, int group, uint8_t *s_inx, uint8_t *s_iny,
                          const unsigned char *s_zero) {
    const long long AHW = (long long)p.A * p.H * p.W;
    // plain-store mode: label + regression items of a panel as one array; bulk mode: the label items only
    const long long per_panel = s_zero ? AHW : 5LL * AHW;
#pragma unroll 1
    for (int b0 = 0; b0 < p.B; b0 += group) {
        const int nb = min(group, p.B - b0);
        if (s_zero && threadIdx.x == 0) bulk_zero_share(p, b0, nb, (int)blockIdx.x, (int)gridDim.x, s_zero, false);
        const long long total = per_panel * nb;
        // shares are multiples of 64 items (1 KB) so that every CTA writes whole, aligned lines
        const long long share = ((total + n_fill - 1) / n_fill + 63) & ~63LL;
        long long lo = share * k, hi = lo + share;
        if (hi > total) hi = total;
        while (lo < hi) {
            const int b = (int)(lo / per_panel);
            const long long base = (long long)b * per_panel;
            const long long seg_hi = (hi < base + per_panel) ? hi : base + per_panel;
            fill_segment(p, b0 + b, (int)(lo - base), (int)(seg_hi - base), s_inx, s_iny);
            lo = seg_hi;
        }
        if (s_zero && threadIdx.x == 0) {
            bulk_wait_group_all();                                // the bulk copies of this round have been written
            __threadfence();
            bulk_zero_share(p, b0, nb, (int)blockIdx.x, (int)gridDim.x, s_zero, true);
        }
    }
}

__device__ __forceinline__ int ld_acquire(const int *ptr) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t hash_key(uint32_t k) { return (k * 2654435761u) >> 7; }

__global__ void __launch_bounds__(kTgtThreads, 1) rpn_targets_kernel(RpnTargetParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_ctl[8];              // 1 ready flag, 2 hits, 3 regular winners, 4 forced positives
    const int HW = p.H * p.W, AHW = p.A * HW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    // ---- shared-memory map of a compute CTA (a fill CTA only uses the in-image tables) ---------------
    double *s_gt = reinterpret_cast<double *>(smem);                                   // [G][4] x1,x2,y1,y2
    float4 *s_gt32 = reinterpret_cast<float4 *>(s_gt + 4 * p.Gmax);                      // [G] x1,y1,x2,y2 rounded
    unsigned long long *s_best = reinterpret_cast<unsigned long long *>(s_gt32 + p.Gmax);
    float *s_area32 = reinterpret_cast<float *>(s_best + p.Gmax);
    unsigned *s_floor = reinterpret_cast<unsigned *>(s_area32 + p.Gmax);                // [G] lower bound of the best IoU (f32 bits)
    int *s_hits = reinterpret_cast<int *>(s_floor + p.Gmax);
    unsigned *s_order = reinterpret_cast<unsigned *>(s_hits + p.Gmax);                  // [G] forced anchor or ~0
    uint8_t *s_skip = reinterpret_cast<uint8_t *>(s_order + p.Gmax);                    // bit0: bg/degenerate, bit1: no filter
    // per anchor shape: coordinates per column / row
    double2 *s_ax = reinterpret_cast<double2 *>(smem + p.sm_off_tables);                // [A][W] anchor x1,x2 of column ix
    double2 *s_ay = s_ax + p.A * p.W;                                                    // [A][H]
    float4 *s_axf = reinterpret_cast<float4 *>(s_ay + p.A * p.H);                        // [A][W] x1,x2,width (f32), in-image flag
    float4 *s_ayf = s_axf + p.A * p.W;                                                   // [A][H]
    int *s_use = reinterpret_cast<int *>(s_ayf + p.A * p.H);                             // [A][4] ix_lo, ix_hi, jy_lo, jy_hi in-image
    uint8_t *s_inx = reinterpret_cast<uint8_t *>(s_use + 4 * p.A);                       // [A][W] fill: column inside the image
    uint8_t *s_iny = s_inx + p.A * p.W;                                                  // [A][H]
    // work items (anchor shape, figure): window and first 32-cell chunk
    int4 *s_range = reinterpret_cast<int4 *>(smem + p.sm_off_items);                     // [A*G]
    int *s_cstart = reinterpret_cast<int *>(s_range + p.n_items_max);                    // [A*G + 1]
    TargetHit *s_hit = reinterpret_cast<TargetHit *>(smem + p.sm_off_hits);              // [hit_cap]
    unsigned long long *s_tmax = reinterpret_cast<unsigned long long *>(smem + p.sm_off_hash);   // [slots] best IoU bits
    uint32_t *s_tkey = reinterpret_cast<uint32_t *>(s_tmax + p.hash_slots);              // [slots] a*HW + cell or ~0
    int *s_tg = reinterpret_cast<int *>(s_tkey + p.hash_slots);                          // [slots] winning figure
    const int hit_cap = p.hit_cap;
    const uint32_t hmask = (uint32_t)p.hash_slots - 1;
    // positives ready to be stored once the panel is filled: regular winners, then forced ones
    double *s_wv = reinterpret_cast<double *>(smem + p.sm_off_win);                      // [hit_cap + Gmax][4]
    int *s_wkey = reinterpret_cast<int *>(s_wv + 4 * (size_t)(hit_cap + p.Gmax));        // [hit_cap + Gmax] a*HW + cell

    // bulk mode: every CTA owns a zeroed buffer the TMA unit reads (compute-only launches of the two-launch form: none)
    const unsigned char *s_zero = (p.zero_bytes > 0 && p.role != 2) ? smem + p.sm_off_zero : nullptr;
    if (s_zero) {
        for (int i = threadIdx.x; i < p.zero_bytes / 16; i += kTgtThreads)
            reinterpret_cast<double2 *>(smem + p.sm_off_zero)[i] = make_double2(0.0, 0.0);
        fence_proxy_async_smem();
        __syncthreads();
    }
    const bool fill_only = p.role == 1 || (p.role == 0 && (int)blockIdx.x < p.n_fill_ctas);
    if (fill_only) {
        TGT_STAMP(0);
        if (p.role == 1) fill_role(p, (int)blockIdx.x, (int)gridDim.x, p.B, s_inx, s_iny, s_zero);
        else fill_role(p, (int)blockIdx.x, p.n_fill_ctas, p.group, s_inx, s_iny, s_zero);
        TGT_STAMP(9);
    } else {
        const int n_comp = p.role == 2 ? (int)gridDim.x : (int)gridDim.x - p.n_fill_ctas;
        const int comp_id = p.role == 2 ? (int)blockIdx.x : (int)blockIdx.x - p.n_fill_ctas;
#pragma unroll 1
        for (int b0 = 0; b0 < p.B; b0 += n_comp) {
            const int b = b0 + comp_id;
            // bulk mode: this CTA's part of the round's regression zeros goes out first, asynchronously
            if (s_zero && threadIdx.x == 0)
                bulk_zero_share(p, b0, min(n_comp, p.B - b0), (int)blockIdx.x, (int)gridDim.x, s_zero, false);
            if (b >= p.B) {                                       // no panel left for this CTA in the last round
                if (s_zero && threadIdx.x == 0) {
                    bulk_wait_group_all();
                    __threadfence();
                    bulk_zero_share(p, b0, min(n_comp, p.B - b0), (int)blockIdx.x, (int)gridDim.x, s_zero, true);
                }
                continue;
            }
            TGT_STAMP(0);
            double *cls_b = p.y_cls + (size_t)b * 2 * AHW;
            double *regr_b = p.y_regr + (size_t)b * 8 * AHW;
            const int G = min(max(p.gt_count[b], 0), p.Gmax);
            const double img_w = p.img_wh[2 * b], img_h = p.img_wh[2 * b + 1];
            __syncthreads();                                      // previous panel fully consumed
            for (int i = threadIdx.x; i < p.Gmax; i += kTgtThreads) {
                const double2 *q = reinterpret_cast<const double2 *>(p.gt + ((size_t)b * p.Gmax + i) * 4);
                const double2 qx = q[0], qy = q[1];
                const uint8_t isbg = p.gt_is_bg[(size_t)b * p.Gmax + i];
                const double x1 = qx.x, x2 = qx.y, y1 = qy.x, y2 = qy.y;
                s_gt[4 * i + 0] = x1; s_gt[4 * i + 1] = x2; s_gt[4 * i + 2] = y1; s_gt[4 * i + 3] = y2;
                s_gt32[i] = make_float4((float)x1, (float)y1, (float)x2, (float)y2);
                s_area32[i] = (float)((x2 - x1) * (y2 - y1));
                s_best[i] = 0ull;
                s_hits[i] = 0;
                s_floor[i] = 0u;
                // 'bg' figures never produce labels (utils.py:690); degenerate ones have IoU 0 (utils.py:103)
                uint8_t f = ((isbg != 0) || (x1 >= x2) || (y1 >= y2)) ? 1 : 0;
                // the float32 estimate is only trusted for pixel-scale coordinates
                if (!(img_w <= 8192.0 && img_h <= 8192.0) ||
                    !(fabs(x1) <= 8192.0 && fabs(x2) <= 8192.0 && fabs(y1) <= 8192.0 && fabs(y2) <= 8192.0)) f |= 2;
                s_skip[i] = f;
            }
            for (int i = threadIdx.x; i < 4 * p.A; i += kTgtThreads) s_use[i] = (i & 1) ? -1 : ((i & 2) ? p.H : p.W);
            for (int i = threadIdx.x; i < p.hash_slots; i += kTgtThreads) {
                s_tkey[i] = 0xFFFFFFFFu;
                s_tmax[i] = 0ull;
                s_tg[i] = 0x7fffffff;
            }
            if (threadIdx.x == 0) { s_ctl[2] = 0; s_ctl[1] = 0; s_ctl[3] = 0; s_ctl[4] = 0; }
            __syncthreads();
            // A LOWER bound of every figure's best float32 IoU, from the exact IoU with the A anchors of
            // the cell under the figure's centre.  Pairs whose float32 estimate is below it by more than
            // the margin cannot be (or tie with) the best anchor.
            for (int i = threadIdx.x; i < G * p.A; i += kTgtThreads) {
                const int g = i / p.A, a2 = i - g * p.A;
                if (s_skip[g] & 1) continue;
                const double gx1 = s_gt[4 * g + 0], gx2 = s_gt[4 * g + 1], gy1 = s_gt[4 * g + 2], gy2 = s_gt[4 * g + 3];
                int cx = (int)floor((gx1 + gx2) * 0.5 / p.stride), cy = (int)floor((gy1 + gy2) * 0.5 / p.stride);
                cx = min(max(cx, 0), p.W - 1);
                cy = min(max(cy, 0), p.H - 1);
                const AnchorPx c = anchor_px(p.stride, cx, cy, p.anchors.wh[a2][0], p.anchors.wh[a2][1]);
                const bool ok = !(c.x1 < 0.0 || c.x2 > img_w) && !(c.y1 < 0.0 || c.y2 > img_h);
                if (ok) {
                    const float v = (float)ref_iou(gx1, gy1, gx2, gy2, c.x1, c.y1, c.x2, c.y2);
                    if (v > 0.f) atomicMax(&s_floor[g], __float_as_uint(v));
                }
            }
            // anchor coordinates per shape and column / row (utils.py:625-626, 635-636) and the per-axis in-image
            // tests (utils.py:629, 638); an anchor is used when both its column and its row pass
            for (int i = threadIdx.x; i < p.A * (p.W + p.H); i += kTgtThreads) {
                const int a = i / (p.W + p.H), r = i - a * (p.W + p.H);
                const bool isx = r < p.W;
                const int k = isx ? r : r - p.W;
                const double side = p.anchors.wh[a][isx ? 0 : 1], lim_px = isx ? img_w : img_h;
                const double c = __dmul_rn(p.stride, (double)k + 0.5);
                const double v1 = __dsub_rn(c, __dmul_rn(side, 0.5)), v2 = __dadd_rn(c, __dmul_rn(side, 0.5));
                const bool ok = !(v1 < 0.0 || v2 > lim_px) && v1 < v2;
                (isx ? s_ax + a * p.W : s_ay + a * p.H)[k] = make_double2(v1, v2);
                (isx ? s_axf + a * p.W : s_ayf + a * p.H)[k] = make_float4((float)v1, (float)v2, (float)(v2 - v1), ok ? 1.f : 0.f);
                if (ok) {
                    atomicMin(&s_use[4 * a + (isx ? 0 : 2)], k);
                    atomicMax(&s_use[4 * a + (isx ? 1 : 3)], k);
                }
            }
            __syncthreads();
            TGT_STAMP(1);

            const float thr32 = (float)p.max_overlap;
            const int n_items = p.A * G;
            // Cell window of anchor shape a that can matter for figure g.  IoU >= L needs, on each axis, an
            // overlap of at least L*max(figure side, anchor side) (because union >= the larger area and the
            // other overlap <= the smaller side); with L = min(floor, thr) - margin this is a handful of cells
            // around the figure.  The window is empty when the two shapes cannot reach L at all
            // (IoU <= smaller-overlap-box / union), and it is clipped to the in-image rectangle of the shape.
            // One cell of padding absorbs the rounding of this float64 arithmetic.
            for (int it = threadIdx.x; it < n_items; it += kTgtThreads) {
                const int a = it / G, g = it - a * G;             // item order = the reference's loop order over shapes
                const double aw = p.anchors.wh[a][0], ah = p.anchors.wh[a][1];
                const int *use = s_use + 4 * a;
                int4 r = make_int4(0, -1, 0, -1);                                    // empty
                const uint8_t f = s_skip[g];
                if (!(f & 1) && use[0] <= use[1] && use[2] <= use[3]) {
                    const double gx1 = s_gt[4 * g + 0], gx2 = s_gt[4 * g + 1], gy1 = s_gt[4 * g + 2], gy2 = s_gt[4 * g + 3];
                    double L = (double)fminf(__uint_as_float(s_floor[g]), thr32) - 2.0 * (double)kIouMargin;
                    if ((f & 2) || !(aw <= 8192.0 && ah <= 8192.0) || !(L > 0.0)) L = 0.0;
                    const double wg = gx2 - gx1, hg = gy2 - gy1;
                    const double imax = fmin(wg, aw) * fmin(hg, ah);                  // largest possible intersection
                    const bool feasible = imax >= (L - 1e-9) * (wg * hg + aw * ah - imax);
                    const double mx = L * fmax(wg, aw), my = L * fmax(hg, ah);
                    // centre c = stride*(i+0.5) must satisfy  g1 + m - side/2 <= c <= g2 - m + side/2
                    const double xl = (gx1 + mx - aw * 0.5) / p.stride - 0.5, xh = (gx2 - mx + aw * 0.5) / p.stride - 0.5;
                    const double yl = (gy1 + my - ah * 0.5) / p.stride - 0.5, yh = (gy2 - my + ah * 0.5) / p.stride - 0.5;
                    if (feasible && xl <= xh + 2.0 && yl <= yh + 2.0) {
                        r.x = max((int)fmax(floor(xl) - 1.0, 0.0), use[0]);
                        r.y = min((int)fmin(ceil(xh) + 1.0, (double)(p.W - 1)), use[1]);
                        r.z = max((int)fmax(floor(yl) - 1.0, 0.0), use[2]);
                        r.w = min((int)fmin(ceil(yh) + 1.0, (double)(p.H - 1)), use[3]);
                    }
                }
                s_range[it] = r;
                const int n = (r.x > r.y || r.z > r.w) ? 0 : (r.y - r.x + 1) * (r.w - r.z + 1);
                s_cstart[it + 1] = (n + 31) >> 5;                                    // chunks of this item, prefix below
            }
            __syncthreads();
            if (w == 0) {                  // inclusive prefix over the items' chunk counts, 32 at a time
                int carry = 0;
                for (int i0 = 0; i0 < n_items; i0 += 32) {
                    const int i = i0 + lane;
                    const int v = i < n_items ? s_cstart[i + 1] : 0;
                    int inc = v;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int n = __shfl_up_sync(0xffffffffu, inc, d);
                        if (lane >= d) inc += n;
                    }
                    if (i < n_items) s_cstart[i + 1] = carry + inc;
                    carry += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (lane == 0) s_cstart[0] = 0;
            }
            __syncthreads();
            TGT_STAMP(2);

            // Phase 1 - the windows of all (shape, figure) items are cut into chunks of 32 cells, one chunk per
            // warp and step.  Cells whose IoU exceeds rpn_max_overlap go to a hit list; which figure wins a
            // cell is settled in phase 2, so the figure order of the reference ("first figure wins ties",
            // utils.py:710-713) does not serialise the warps.
            const int n_chunks = n_items ? s_cstart[n_items] : 0;
#pragma unroll 1
            for (int c = w; c < n_chunks; c += kTgtWarps) {
                int it = 0;                                                           // last item with s_cstart[it] <= c
                for (int hi = n_items - 1; it < hi;) {
                    const int mid = (it + hi + 1) >> 1;
                    if (s_cstart[mid] <= c) it = mid; else hi = mid - 1;
                }
                const int a = it / G, g = it - a * G;
                const int4 rg = s_range[it];
                const int ww = rg.y - rg.x + 1, n = ww * (rg.w - rg.z + 1);
                const int t = (c - s_cstart[it]) * 32 + lane;
                const bool act = t < n;
                const int dy = act ? t / ww : 0;
                const int ix = rg.x + (act ? t - dy * ww : 0), jy = rg.z + dy;
                const uint8_t gflag = s_skip[g];
                const double gx1 = s_gt[4 * g + 0], gx2 = s_gt[4 * g + 1], gy1 = s_gt[4 * g + 2], gy2 = s_gt[4 * g + 3];
                const float4 gf = s_gt32[g];
                const float lim = fminf(__uint_as_float(s_floor[g]), thr32);
                const double2 X = s_ax[a * p.W + ix], Y = s_ay[a * p.H + jy];
                const float4 XF = s_axf[a * p.W + ix], YF = s_ayf[a * p.H + jy];
                // anchors crossing the image are skipped entirely (utils.py:629,638); a degenerate anchor has IoU 0
                const bool usable = act && XF.w != 0.f && YF.w != 0.f;
                // IoU > 0  <=>  the open intervals meet on both axes (exact, float64 compares only)
                const bool isect = usable && gx2 > X.x && X.y > gx1 && gy2 > Y.x && Y.y > gy1;
                // float32 estimate of the IoU: decides whether the exact float64 value can matter at all
                bool need = false;
                if (isect) {
                    const float wi = fminf(gf.z, XF.y) - fmaxf(gf.x, XF.x);
                    const float hi = fminf(gf.w, YF.y) - fmaxf(gf.y, YF.x);
                    const float itf = fmaxf(wi, 0.f) * fmaxf(hi, 0.f);
                    const float q = __fdividef(itf, s_area32[g] + XF.z * YF.z - itf);
                    need = (q + kIouMargin >= lim) ||           // could be the best anchor, or exceed rpn_max_overlap
                           (gflag & 2) || !(XF.z <= 8192.f && YF.z <= 8192.f);   // estimate not trusted: always exact
                }
                if (!__any_sync(0xffffffffu, need)) continue;                         // warp-uniform
                unsigned bits = 0;
                bool hit = false;
                if (need) {
                    const double iou = ref_iou(gx1, gy1, gx2, gy2, X.x, Y.x, X.y, Y.y);
                    const float iou32 = (float)iou;                                   // float32 accumulator (utils.py:603)
                    if (iou32 > 0.f) bits = __float_as_uint(iou32);
                    hit = iou > p.max_overlap;                                        // utils.py:704
                    if (hit) {
                        const int pos = atomicAdd(&s_ctl[2], 1);
                        if (pos < hit_cap) s_hit[pos] = TargetHit{iou, a * HW + jy * p.W + ix, g};
                    }
                }
                // best anchor of this figure: max float32 IoU, then first in loop order size->ratio->ix->jy
                const unsigned order = (unsigned)((a * p.W + ix) * p.H + jy);
                const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
                const unsigned hm = __ballot_sync(0xffffffffu, hit);
                if (wmax) {
                    const unsigned omin = __reduce_min_sync(0xffffffffu, bits == wmax ? order : 0xFFFFFFFFu);
                    if (lane == 0) atomicMax(&s_best[g], ((unsigned long long)wmax << 32) | (0xFFFFFFFFu - omin));
                }
                if (hm && lane == 0) atomicAdd(&s_hits[g], __popc(hm));
            }
            __syncthreads();
            TGT_STAMP(3);

            // Phase 2 - settle every hit anchor: highest IoU wins, equal IoU -> the earlier figure (strict '>'
            // in figure order, utils.py:710-713).  Open-addressing table keyed by the anchor; three passes.  The
            // winners and their regression targets are parked in shared memory: once the panel is filled only
            // stores are left.
            const int n_hit = s_ctl[2];
            const bool replay = n_hit > hit_cap;
            if (!replay) {
                for (int e = threadIdx.x; e < n_hit; e += kTgtThreads) {
                    const TargetHit h = s_hit[e];
                    uint32_t slot = hash_key((uint32_t)h.key) & hmask;
                    while (true) {
                        const uint32_t prev = atomicCAS(&s_tkey[slot], 0xFFFFFFFFu, (uint32_t)h.key);
                        if (prev == 0xFFFFFFFFu || prev == (uint32_t)h.key) break;
                        slot = (slot + 1) & hmask;
                    }
                    atomicMax(&s_tmax[slot], (unsigned long long)__double_as_longlong(h.iou));   // positive doubles order like their bits
                }
                __syncthreads();
                for (int e = threadIdx.x; e < n_hit; e += kTgtThreads) {
                    const TargetHit h = s_hit[e];
                    uint32_t slot = hash_key((uint32_t)h.key) & hmask;
                    while (s_tkey[slot] != (uint32_t)h.key) slot = (slot + 1) & hmask;
                    if ((unsigned long long)__double_as_longlong(h.iou) == s_tmax[slot]) atomicMin(&s_tg[slot], h.g);
                }
                __syncthreads();
                for (int e = threadIdx.x; e < n_hit; e += kTgtThreads) {
                    const TargetHit h = s_hit[e];
                    uint32_t slot = hash_key((uint32_t)h.key) & hmask;
                    while (s_tkey[slot] != (uint32_t)h.key) slot = (slot + 1) & hmask;
                    if ((unsigned long long)__double_as_longlong(h.iou) == s_tmax[slot] && s_tg[slot] == h.g) {
                        const int a2 = h.key / HW, pos = atomicAdd(&s_ctl[3], 1);
                        s_wkey[pos] = h.key;
                        positive_values(p, a2, h.key - a2 * HW, s_gt + 4 * h.g, false, s_wv + 4 * pos);
                    }
                }
            }
            // forced positives + best_anchor table (utils.py:741-766): decode the best anchor of every figure
            for (int g = threadIdx.x; g < p.Gmax; g += kTgtThreads) {
                const unsigned long long key = g < G ? s_best[g] : 0ull;
                const int nh = g < G ? s_hits[g] : 0;
                unsigned order = 0xFFFFFFFFu;
                int4 out = make_int4(-1, -1, -1, -1);
                if (key) {
                    const unsigned o = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
                    const int jy = (int)(o % (unsigned)p.H);
                    const unsigned rest = o / (unsigned)p.H;
                    const int ix = (int)(rest % (unsigned)p.W);
                    const int a2 = (int)(rest / (unsigned)p.W);
                    out = make_int4(jy, ix, a2 % p.n_ratios, a2 / p.n_ratios);            // utils.py:697
                    if (nh == 0) order = o;
                }
                *reinterpret_cast<int4 *>(p.best_anchor + ((size_t)b * p.Gmax + g) * 4) = out;
                p.n_hits[(size_t)b * p.Gmax + g] = nh;
                s_order[g] = order;
            }
            __syncthreads();
            // The reference applies the forced positives in GT order, so when several GT share the same best
            // anchor the LAST one wins: a figure is only kept if no later forced figure targets its anchor.
            const int n_win = s_ctl[3];
            for (int g = threadIdx.x; g < G; g += kTgtThreads) {
                const unsigned o = s_order[g];
                if (o == 0xFFFFFFFFu) continue;
                bool last = true;
                for (int g2 = g + 1; g2 < G; ++g2) last = last && (s_order[g2] != o);
                if (!last) continue;
                const int jy = (int)(o % (unsigned)p.H);
                const unsigned rest = o / (unsigned)p.H;
                const int ix = (int)(rest % (unsigned)p.W);
                const int a2 = (int)(rest / (unsigned)p.W);
                const int pos = hit_cap + atomicAdd(&s_ctl[4], 1);
                s_wkey[pos] = a2 * HW + jy * p.W + ix;
                positive_values(p, a2, jy * p.W + ix, s_gt + 4 * g, true, s_wv + 4 * pos);
            }
            TGT_STAMP(4);

            // ---- wait until every item of this panel has been filled -----------------------------------
            if (s_zero && threadIdx.x == 0) {                     // first publish this CTA's own part of the zeros
                bulk_wait_group_all();
                __threadfence();
                bulk_zero_share(p, b0, min(n_comp, p.B - b0), (int)blockIdx.x, (int)gridDim.x, s_zero, true);
            }
            if (p.role == 0) {
                const int want = 5 * AHW;
                const long long t_start = global_ns();
                while (true) {
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        const int done = ld_acquire(&p.panel_done[b]);
                        s_ctl[1] = done >= want;
                        if (!s_ctl[1] && global_ns() - t_start > 4000000000LL) s_ctl[1] = 2;
                    }
                    __syncthreads();
                    if (s_ctl[1]) break;
                }
                if (s_ctl[1] == 2)         // the fill never completed (4 s): results invalid, reported through n_hits
                    for (int g = threadIdx.x; g < p.Gmax; g += kTgtThreads) p.n_hits[(size_t)b * p.Gmax + g] = -1;
            } else {
                __syncthreads();
            }
            TGT_STAMP(5);
            if (threadIdx.x == 0) p.panel_done[b] = 0;            // leave the workspace clean

            // regular positives (utils.py:728-738)
            if (!replay) {
                for (int e = threadIdx.x; e < n_win; e += kTgtThreads) {
                    const int key = s_wkey[e], a2 = key / HW;
                    store_positive(p, cls_b, regr_b, a2, key - a2 * HW, s_wv + 4 * e, false);
                }
            } else {
                // more positives than the list holds (never seen in practice): shape by shape, figure by figure
                // with in-place per-cell state (the hit list and the table are not needed any more)
                double *s_lb = reinterpret_cast<double *>(smem + p.sm_off_hits);      // [HW]
                int *s_lg = reinterpret_cast<int *>(s_lb + HW);                       // [HW]
#pragma unroll 1
                for (int a = 0; a < p.A; ++a) {
                    __syncthreads();
                    for (int cell = threadIdx.x; cell < HW; cell += kTgtThreads) { s_lb[cell] = 0.0; s_lg[cell] = -1; }
                    __syncthreads();
#pragma unroll 1
                    for (int g = 0; g < G; ++g) {
                        const int4 rg = s_range[a * G + g];
                        if (rg.x > rg.y || rg.z > rg.w) continue;                     // block-uniform
                        const int ww = rg.y - rg.x + 1, n = ww * (rg.w - rg.z + 1);
                        const double gx1 = s_gt[4 * g + 0], gx2 = s_gt[4 * g + 1], gy1 = s_gt[4 * g + 2], gy2 = s_gt[4 * g + 3];
                        for (int t = threadIdx.x; t < n; t += kTgtThreads) {
                            const int dy = t / ww, ix = rg.x + t - dy * ww, jy = rg.z + dy;
                            if (s_axf[a * p.W + ix].w == 0.f || s_ayf[a * p.H + jy].w == 0.f) continue;
                            const double2 X = s_ax[a * p.W + ix], Y = s_ay[a * p.H + jy];
                            const double iou = ref_iou(gx1, gy1, gx2, gy2, X.x, Y.x, X.y, Y.y);
                            const int cell = jy * p.W + ix;
                            if (iou > p.max_overlap && iou > s_lb[cell]) { s_lb[cell] = iou; s_lg[cell] = g; }
                        }
                        __syncthreads();
                    }
                    for (int cell = threadIdx.x; cell < HW; cell += kTgtThreads) {
                        const int lg = s_lg[cell];
                        if (lg >= 0) {
                            double v[4];
                            positive_values(p, a, cell, s_gt + 4 * lg, false, v);
                            store_positive(p, cls_b, regr_b, a, cell, v, false);
                        }
                    }
                }
            }
            __syncthreads();                                                          // regular before forced writes
            const int n_forced = s_ctl[4];
            for (int e = threadIdx.x; e < n_forced; e += kTgtThreads) {
                const int key = s_wkey[hit_cap + e], a2 = key / HW;
                store_positive(p, cls_b, regr_b, a2, key - a2 * HW, s_wv + 4 * (hit_cap + e), true);
            }
#ifdef RADNET_TGT_PROFILE
            if (p.stamps && threadIdx.x == 0) {
                p.stamps[(size_t)blockIdx.x * 16 + 10] = n_chunks;
                p.stamps[(size_t)blockIdx.x * 16 + 11] = n_hit;
                p.stamps[(size_t)blockIdx.x * 16 + 12] = n_win;
                p.stamps[(size_t)blockIdx.x * 16 + 13] = n_forced;
            }
#endif
            TGT_STAMP(6);
        }
    }
}

}  // namespace radnet

using namespace radnet;

namespace {
struct TgtSmemLayout {
    size_t off_tables, off_items, off_hits, off_hash, off_win, total;
    int hit_cap, hash_slots, n_items_max;
};
TgtSmemLayout tgt_smem_layout(int Gmax, int H, int W, int A, int hit_cap) {
    TgtSmemLayout l;
    const size_t gm = Gmax > 0 ? Gmax : 1, HW = (size_t)H * W;
    size_t off = align_up(gm * (32 + 16 + 8 + 4 + 4 + 4 + 4 + 1) + 16, 16);
    l.off_tables = off;
    off += align_up((size_t)A * (W + H) * 32 + (size_t)A * 16 + (size_t)A * (W + H), 16);
    l.off_items = off;
    l.n_items_max = (int)(A * gm);
    off += align_up((size_t)l.n_items_max * 16 + ((size_t)l.n_items_max + 1) * 4, 16);
    l.off_hits = off;
    l.hit_cap = hit_cap;
    int slots = 8;
    while (slots < 2 * hit_cap) slots <<= 1;
    l.hash_slots = slots;
    size_t hits_bytes = align_up((size_t)hit_cap * sizeof(TargetHit), 16);
    const size_t hash_bytes = (size_t)slots * 16;
    // the replay path reuses both regions as {double iou[HW]; int figure[HW]}
    if (hits_bytes + hash_bytes < 12 * HW + 16) hits_bytes = align_up(12 * HW + 16 - hash_bytes, 16);
    l.off_hash = off + hits_bytes;
    l.off_win = l.off_hash + hash_bytes;
    l.total = l.off_win + ((size_t)hit_cap + gm) * (4 * sizeof(double) + sizeof(int)) + 16;
    return l;
}
size_t tgt_ws_bytes(int B) { return align_up((size_t)B * sizeof(int32_t) + 16, 256); }
}  // namespace

extern "C" size_t radnet_rpn_targets_workspace_bytes(int B, int Gmax, int H, int W, int A) {
    if (B < 1 || Gmax < 0 || H < 1 || W < 1 || A < 1) return 0;
    return tgt_ws_bytes(B);
}

extern "C" int radnet_rpn_targets_workspace_init(void *ws, size_t ws_bytes, int B, int Gmax, void *stream) {
    RADNET_CHECK_ARG(ws && B >= 1 && Gmax >= 0, "rpn_targets_workspace_init: bad arguments");
    const size_t need = tgt_ws_bytes(B);
    if (ws_bytes < need) {
        set_error("rpn_targets_workspace_init: workspace %zu < %zu", ws_bytes, need);
        return RADNET_E_WORKSPACE;
    }
    RADNET_CUDA(cudaMemsetAsync(ws, 0, need, (cudaStream_t)stream));
    return RADNET_OK;
}

#ifdef RADNET_TGT_PROFILE
static long long *g_tgt_stamps = nullptr;
extern "C" int radnet_debug_set_tgt_stamps(long long *dev_ptr) {
    g_tgt_stamps = dev_ptr;
    return 0;
}
#endif

extern "C" int radnet_rpn_targets(const double *gt, const uint8_t *gt_is_bg, const int32_t *gt_count, int B,
                                  int Gmax, int H, int W, int A, int n_ratios, const double *h_anchor_px,
                                  double rpn_stride, const double *img_wh, double max_overlap, int layout,
                                  double regr_scale, double *y_rpn_cls, double *y_rpn_regr, int32_t *best_anchor,
                                  int32_t *n_hits, void *ws, size_t ws_bytes, void *stream) {
    RADNET_CHECK_ARG(gt_count && h_anchor_px && img_wh && y_rpn_cls && y_rpn_regr && ws, "rpn_targets: null pointer");
    RADNET_CHECK_ARG(Gmax == 0 || (gt && gt_is_bg && best_anchor && n_hits), "rpn_targets: null GT buffers");
    RADNET_CHECK_ARG(B >= 1 && B <= (1 << 24) && H >= 1 && W >= 1 && A >= 1 && A <= kMaxAnchors && n_ratios >= 1 && Gmax >= 0,
                     "rpn_targets: bad sizes B=%d H=%d W=%d A=%d Gmax=%d", B, H, W, A, Gmax);
    RADNET_CHECK_ARG((long long)A * H * W < (1LL << 27), "rpn_targets: anchor count overflows the loop-order key");
    RADNET_CHECK_ARG(layout == RADNET_TARGETS_CHANNEL_FIRST || layout == RADNET_TARGETS_NHWC, "rpn_targets: bad layout %d",
                     layout);
    RADNET_CHECK_ARG(((uintptr_t)y_rpn_cls & 15) == 0 && ((uintptr_t)y_rpn_regr & 15) == 0 &&
                         ((uintptr_t)gt & 15) == 0 && ((uintptr_t)best_anchor & 15) == 0,
                     "rpn_targets: gt, best_anchor and the output tensors must be 16-byte aligned");
    if (ws_bytes < tgt_ws_bytes(B)) {
        set_error("rpn_targets: workspace %zu < %zu", ws_bytes, tgt_ws_bytes(B));
        return RADNET_E_WORKSPACE;
    }
    int hit_cap = 1024;
    {   // tests shrink the list to exercise the replay path
        const long long v = get_option(kOptTargetsHitCap);
        if (v >= 1 && v < hit_cap) hit_cap = (int)v;
    }
    const TgtSmemLayout sl = tgt_smem_layout(Gmax, H, W, A, hit_cap);
    int dev = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    const int smem_limit = device_smem_optin(dev), n_sm = device_sm_count(dev);
    if (smem_limit < 0 || n_sm < 1) return RADNET_E_CUDA;
    if (sl.total > (size_t)smem_limit) {
        set_error("rpn_targets: %d figures on a %dx%dx%d map need %zu B of shared memory (limit %d)", Gmax, H, W, A,
                  sl.total, smem_limit);
        return RADNET_E_UNSUPPORTED;
    }
    RpnTargetParams p{};
    p.gt = gt; p.gt_is_bg = gt_is_bg; p.gt_count = gt_count;
    p.B = B; p.Gmax = Gmax; p.H = H; p.W = W; p.A = A; p.n_ratios = n_ratios;
    for (int a = 0; a < A; ++a) {
        p.anchors.wh[a][0] = h_anchor_px[2 * a];
        p.anchors.wh[a][1] = h_anchor_px[2 * a + 1];
    }
    p.stride = rpn_stride; p.img_wh = img_wh; p.max_overlap = max_overlap;
    p.y_cls = y_rpn_cls; p.y_regr = y_rpn_regr; p.best_anchor = best_anchor; p.n_hits = n_hits;
    p.layout = layout; p.regr_scale = regr_scale;
    p.panel_done = reinterpret_cast<int32_t *>(ws);
    p.sm_off_tables = (int)sl.off_tables; p.sm_off_items = (int)sl.off_items; p.sm_off_hits = (int)sl.off_hits;
    p.sm_off_hash = (int)sl.off_hash; p.sm_off_win = (int)sl.off_win; p.hit_cap = sl.hit_cap; p.hash_slots = sl.hash_slots; p.n_items_max = sl.n_items_max;
#ifdef RADNET_TGT_PROFILE
    p.stamps = g_tgt_stamps;
#endif
    size_t smem_total = sl.total;
    {   // zero source of the bulk copies: a buffer behind the compute layout, as large as fits (at most 32 KB)
        // measured (64 panels, B200): 27.2 us per launch with the bulk copies against 25.4 us with plain stores - the
        // zeros land earlier (fill shares written at 14.3 us instead of 16.9 us) but every CTA starts ~2 us later, so
        // the bulk mode is off unless asked for
        const long long want = get_option(kOptTargetsFillBulk);        // 0 plain stores, 1 bulk copies (largest buffer), else bytes
        const size_t off = align_up(sl.total, 128);
        size_t z = (size_t)smem_limit - 256 > off ? (((size_t)smem_limit - 256 - off) & ~(size_t)1023) : 0;
        if (z > 32768) z = 32768;
        if (want > 1 && (size_t)want < z) z = (size_t)want & ~(size_t)1023;
        p.zero_bytes = (want <= 0 || z < 4096) ? 0 : (int)z;
        p.sm_off_zero = (int)off;
        if (p.zero_bytes) smem_total = off + z;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(rpn_targets_kernel), dev, smem_total)) return rc;
    // SM roles: about 43 % of the SMs compute (one panel at a time each), the rest stream the fill.  With one
    // CTA per SM the whole grid is resident, and fill CTAs - the lower block indices - never wait on anyone.
    long long n_comp = get_option(kOptTargetsComputeCtas);
    if (n_comp < 1) n_comp = (n_sm * 43 + 99) / 100;
    if (n_comp > B) n_comp = B;
    if (n_comp > n_sm - 1) n_comp = n_sm > 1 ? n_sm - 1 : 1;
    long long n_fill = n_sm - n_comp;
    if (n_fill < 1) n_fill = 1;
    if (get_option(kOptTargetsTwoLaunches) == 1) {
        // no co-residency assumed: the fill as its own launch, then the panels (stream order replaces the wait)
        p.role = 1; p.n_fill_ctas = 0; p.group = B;
        rpn_targets_kernel<<<(unsigned)n_sm, kTgtThreads, smem_total, st>>>(p);
        if (int rc = check_launch("rpn_targets_kernel (fill)")) return rc;
        p.role = 2;
        const long long g2 = B < n_sm ? B : n_sm;
        rpn_targets_kernel<<<(unsigned)g2, kTgtThreads, smem_total, st>>>(p);
        return check_launch("rpn_targets_kernel (panels)");
    }
    p.role = 0;
    p.n_fill_ctas = (int)n_fill;
    p.group = (int)n_comp;
    rpn_targets_kernel<<<(unsigned)(n_fill + n_comp), kTgtThreads, smem_total, st>>>(p);
    return check_launch("rpn_targets_kernel");
}


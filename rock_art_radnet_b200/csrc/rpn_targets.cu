// rpn_targets.cu - K3: RPN anchor target assignment (reference faster_rcnn/utils.py:554-775, 815-816;
// upstream name calc_rpn), batched over panels, ONE launch.
//
// The launch is persistent - two CTAs of 512 threads per SM - and there are two kinds of work, both handed
// out through counters in the workspace:
//   * fill units.  The regression tensor is zero and the label tensor holds only the "anchor lies inside
//     the image" flags everywhere except at the few positive anchors, so 10*A*H*W*8 bytes per panel are
//     known at kernel start.  A unit is a quarter of one panel's output; it is streamed with 16-byte stores
//     and then added to the panel's counter of filled items (release);
//   * panels.  All anchor shapes against the figures of the panel, in shared memory: exact float64 IoU only
//     where it can matter (see below); the positives and their regression targets are parked in shared memory.
//     When the panel's counter says it is completely filled (acquire) they are stored on top, a positive is
//     forced for figures without one, and best_anchor / n_hits are written.
// SMs are given a preference by %smid: most SMs fill first and take panels only when no fill unit is left;
// a few SMs take panels first (no store stream next to them: global stores that wait for L2 credits stall the
// SM's whole load/store pipe, shared-memory traffic included - measured 2-3x on every compute phase) and join
// the fill afterwards.  A CTA that waits for its panel pulls fill units meanwhile, so the launch completes
// under any CTA placement.  Measured on B200 (tools/exp/fill_bw.cu): a bare fill of 66.6 MB needs >= 110 SMs
// storing to reach the 4.06 TB/s such a short burst gets (6.0 TB/s for 533 MB), which is why compute cannot
// have many SMs to itself.
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

constexpr int kUnitsPerPanel = 16;
constexpr int kGroupPanels = 64;          // fill completion is published and awaited per group of panels

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
    int32_t *group_done;       // [n_groups] double2 items filled so far in each group of kGroupPanels panels
    int32_t *ctl;              // {next fill unit, next panel, CTAs finished}
    int n_compute_sm, n_sm;    // this many SMs, spread evenly over %smid, take panels first
    int role;                  // 0 both kinds of work in one launch, 1 fill only, 2 panels only (fill already done)
    // shared-memory layout (byte offsets)
    int sm_off_tables, sm_off_items, sm_off_hits, sm_off_hash, sm_off_win, hit_cap, hash_slots, n_items_max;
    long long *stamps;         // profiling build only
};

struct TargetHit {
    double iou;
    double v[4];   // regression targets as stored (times regr_scale)
    int key;       // a*H*W + cell
    int g;         // figure; set to -1 - g when the entry loses its anchor to another figure
};

constexpr int kNeedCap = 4096;            // candidate pairs queued for the exact float64 pass

__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int sm_id() {
    int v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}
__device__ __forceinline__ int ld_acquire(const int *ptr) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t hash_key(uint32_t k) { return (k * 2654435761u) >> 7; }

#ifdef RADNET_TGT_PROFILE
#define TGT_STAMP(i)                                                                                  \
    do {                                                                                              \
        if (p.stamps && threadIdx.x == 0) p.stamps[(size_t)blockIdx.x * 16 + (i)] = global_ns();      \
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

// shared-memory views of one CTA
struct TgtShared {
    double *gt;                 // [G][4] x1,x2,y1,y2
    float4 *gt32;               // [G] x1,y1,x2,y2 rounded
    unsigned long long *best;   // [G] (f32 IoU bits << 32) | (~loop order)
    float *area32;              // [G]
    unsigned *floor;            // [G] lower bound of the best IoU (f32 bits)
    int *hits;                  // [G]
    unsigned *order;            // [G] forced anchor (loop order) or ~0
    uint8_t *skip;              // [G] bit0: bg/degenerate, bit1: float32 estimate not trusted
    double2 *ax, *ay;           // [A][W], [A][H] anchor x1,x2 per column / y1,y2 per row
    float4 *axf, *ayf;          // [A][W], [A][H] x1,x2,width (f32), in-image flag
    int *use;                   // [A][4] ix_lo, ix_hi, jy_lo, jy_hi inside the image
    uint8_t *inx, *iny;         // [A][W], [A][H] fill: column / row inside the image (and G > 0)
    int4 *range;                // [A*G] cell window of (shape, figure)
    int *pstart;                // [A*G + 1] first candidate pair of the item
    TargetHit *hit;             // [hit_cap]
    int *need;                  // [kNeedCap] candidate pairs whose exact IoU is needed
    unsigned long long *tmax;   // [slots] best IoU bits of an anchor
    uint32_t *tkey;             // [slots] a*HW + cell or ~0
    int *tg;                    // [slots] winning figure
    double *wv;                 // [G][4] values of the forced positives
    int *wkey;                  // [G] their anchors
    int *ctl;                   // static: 0 pulled index, 1 ready flag, 2 hits, 3 winners, 4 forced, 5 scan carry,
                                //         6 prefetched fill unit, 7 group of the unpublished items, 8 their count,
                                //         9 queued pairs
    int *warp;                  // static [33] scan scratch
};

__device__ __forceinline__ TgtShared carve(const RpnTargetParams &p, unsigned char *smem, int *s_ctl, int *s_warp) {
    TgtShared s;
    s.gt = reinterpret_cast<double *>(smem);
    s.gt32 = reinterpret_cast<float4 *>(s.gt + 4 * p.Gmax);
    s.best = reinterpret_cast<unsigned long long *>(s.gt32 + p.Gmax);
    s.area32 = reinterpret_cast<float *>(s.best + p.Gmax);
    s.floor = reinterpret_cast<unsigned *>(s.area32 + p.Gmax);
    s.hits = reinterpret_cast<int *>(s.floor + p.Gmax);
    s.order = reinterpret_cast<unsigned *>(s.hits + p.Gmax);
    s.skip = reinterpret_cast<uint8_t *>(s.order + p.Gmax);
    s.ax = reinterpret_cast<double2 *>(smem + p.sm_off_tables);
    s.ay = s.ax + p.A * p.W;
    s.axf = reinterpret_cast<float4 *>(s.ay + p.A * p.H);
    s.ayf = s.axf + p.A * p.W;
    s.use = reinterpret_cast<int *>(s.ayf + p.A * p.H);
    s.inx = reinterpret_cast<uint8_t *>(s.use + 4 * p.A);
    s.iny = s.inx + p.A * p.W;
    s.range = reinterpret_cast<int4 *>(smem + p.sm_off_items);
    s.pstart = reinterpret_cast<int *>(s.range + p.n_items_max);
    s.hit = reinterpret_cast<TargetHit *>(smem + p.sm_off_hits);
    s.tmax = reinterpret_cast<unsigned long long *>(smem + p.sm_off_hash);
    s.tkey = reinterpret_cast<uint32_t *>(s.tmax + p.hash_slots);
    s.tg = reinterpret_cast<int *>(s.tkey + p.hash_slots);
    s.wv = reinterpret_cast<double *>(smem + p.sm_off_win);
    s.wkey = reinterpret_cast<int *>(s.wv + 4 * (size_t)p.Gmax);
    s.need = s.wkey + p.Gmax;
    s.ctl = s_ctl;
    s.warp = s_warp;
    return s;
}

// thread 0: make the items this CTA filled since its last publication visible and count them for their group.
// The fence is cumulative over the stores of all threads ordered before it by a CTA barrier.
__device__ __forceinline__ void publish_fill(const RpnTargetParams &p, const TgtShared &s) {
    if (s.ctl[8] > 0) {
        __threadfence();
        atomicAdd(&p.group_done[s.ctl[7]], s.ctl[8]);
        s.ctl[8] = 0;
    }
}

// ---- fill.  Both tensors of a panel are seen as ONE array of 5*A*H*W double2 items (label tensor first, then the
//      regression tensor); unit u = sixteenth (u % 16) of panel u / 16. ---------------------------------------
template <int NT>
__device__ void fill_unit(const RpnTargetParams &p, const TgtShared &s, int unit) {
    const int b = unit / kUnitsPerPanel, q = unit - b * kUnitsPerPanel;
    const int HW = p.H * p.W, AHW = p.A * HW;
    const int n_items = 5 * AHW;
    const int per = ((n_items + kUnitsPerPanel - 1) / kUnitsPerPanel + 63) & ~63;
    const int lo = min(q * per, n_items), hi = min(lo + per, n_items);
    double2 *cls2 = reinterpret_cast<double2 *>(p.y_cls + (size_t)b * 2 * AHW);
    double2 *regr2 = reinterpret_cast<double2 *>(p.y_regr + (size_t)b * 8 * AHW);
    {   // regression tensor: zero wherever no anchor is positive
        const double2 z = make_double2(0.0, 0.0);
        const int r_lo = max(lo, AHW) - AHW, r_hi = hi - AHW;
#pragma unroll 4
        for (int i = r_lo + threadIdx.x; i < r_hi; i += NT) regr2[i] = z;
    }
    if (lo < AHW) {
        // label tensor: [valid | overlap]; valid = anchor inside the image on both axes (utils.py:629, 638), and
        // labels are only ever written inside the GT loop: no GT, no labels (utils.py:722-738)
        const int G = min(max(p.gt_count[b], 0), p.Gmax);
        const double img_w = p.img_wh[2 * b], img_h = p.img_wh[2 * b + 1];
        for (int i = threadIdx.x; i < p.A * (p.W + p.H); i += NT) {
            const int c = i / (p.W + p.H), r = i - c * (p.W + p.H);
            const bool isx = r < p.W;
            const int k = isx ? r : r - p.W;
            const double side = p.anchors.wh[c][isx ? 0 : 1], lim_px = isx ? img_w : img_h;
            const double ctr = __dmul_rn(p.stride, (double)k + 0.5);
            const double v1 = __dsub_rn(ctr, __dmul_rn(side, 0.5)), v2 = __dadd_rn(ctr, __dmul_rn(side, 0.5));
            const bool ok = !(v1 < 0.0 || v2 > lim_px) && v1 < v2 && G > 0;
            (isx ? s.inx + c * p.W : s.iny + c * p.H)[k] = ok ? 1 : 0;
        }
        __syncthreads();
        const int twoA = 2 * p.A, top = min(hi, AHW);
        for (int i = lo + threadIdx.x; i < top; i += NT) {
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
                    val = (s.inx[c * p.W + ix] & s.iny[c * p.H + jy]) ? 1.0 : 0.0;
                }
                v[h] = val;
            }
            cls2[i] = make_double2(v[0], v[1]);
        }
    }
    __syncthreads();                                              // all stores of the unit issued; tables free
    if (threadIdx.x == 0) {
        // completion is published per group of panels: one fence when this CTA leaves a group (or on demand),
        // not one per unit - a fence per unit stalls the CTA at its next barrier for the whole drain time
        const int gr = b / kGroupPanels;
        if (s.ctl[7] != gr) publish_fill(p, s);
        s.ctl[7] = gr;
        s.ctl[8] += hi - lo;
    }
}

// next fill unit from the shared counter; false when none is left.  The index of the unit after this one is
// requested before the stores of this one are issued (s.ctl[6] holds a prefetched index, -1 = none), so the
// round trip of the atomic is hidden behind the store stream.
template <int NT>
__device__ bool pull_fill(const RpnTargetParams &p, const TgtShared &s) {
    __syncthreads();
    if (threadIdx.x == 0) {
        s.ctl[0] = s.ctl[6] >= 0 ? s.ctl[6] : atomicAdd(&p.ctl[0], 1);
        s.ctl[6] = -1;
    }
    __syncthreads();
    const int u = s.ctl[0];
    if (u >= p.B * kUnitsPerPanel) {
        if (threadIdx.x == 0) publish_fill(p, s);                 // nothing left to pull: hand in what is pending
        return false;
    }
    int next = -1;
    if (threadIdx.x == 0) next = atomicAdd(&p.ctl[0], 1);        // consumed after the stores below
    fill_unit<NT>(p, s, u);
    if (threadIdx.x == 0) s.ctl[6] = next;
    return true;
}

// CTA-wide exclusive scan of the item sizes in s.pstart[1..n] (in place: pstart[i] = first pair of item i)
template <int NT>
__device__ void scan_items(const TgtShared &s, int n) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s.ctl[5] = 0; s.pstart[0] = 0; }
    for (int i0 = 0; i0 < n; i0 += NT) {                 // block-uniform
        const int i = i0 + threadIdx.x;
        const int v = i < n ? s.pstart[i + 1] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        __syncthreads();
        if (lane == 31) s.warp[w] = inc;
        __syncthreads();
        if (w == 0) {
            const int x = lane < NT / 32 ? s.warp[lane] : 0;       // NT / 32 <= 32 warps
            int xi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, xi, d);
                if (lane >= d) xi += t;
            }
            if (lane < NT / 32) s.warp[lane] = xi - x;
            if (lane == 31) s.warp[32] = xi;
        }
        __syncthreads();
        const int carry = s.ctl[5];
        if (i < n) s.pstart[i + 1] = carry + s.warp[w] + inc;
        __syncthreads();
        if (threadIdx.x == 0) s.ctl[5] = carry + s.warp[32];
    }
    __syncthreads();
}

// ---- one panel --------------------------------------------------------------------------------------------------
template <int NT>
__device__ void do_panel(const RpnTargetParams &p, const TgtShared &s, int b) {
    const int HW = p.H * p.W, AHW = p.A * HW;
    const int hit_cap = p.hit_cap;
    const uint32_t hmask = (uint32_t)p.hash_slots - 1;
    TGT_STAMP(0);
    double *cls_b = p.y_cls + (size_t)b * 2 * AHW;
    double *regr_b = p.y_regr + (size_t)b * 8 * AHW;
    const int G = min(max(p.gt_count[b], 0), p.Gmax);
    const double img_w = p.img_wh[2 * b], img_h = p.img_wh[2 * b + 1];
    __syncthreads();                                      // previous work of this CTA fully consumed
    for (int i = threadIdx.x; i < p.Gmax; i += NT) {
        const double2 *q = reinterpret_cast<const double2 *>(p.gt + ((size_t)b * p.Gmax + i) * 4);
        const double2 qx = q[0], qy = q[1];
        const uint8_t isbg = p.gt_is_bg[(size_t)b * p.Gmax + i];
        const double x1 = qx.x, x2 = qx.y, y1 = qy.x, y2 = qy.y;
        s.gt[4 * i + 0] = x1; s.gt[4 * i + 1] = x2; s.gt[4 * i + 2] = y1; s.gt[4 * i + 3] = y2;
        s.gt32[i] = make_float4((float)x1, (float)y1, (float)x2, (float)y2);
        s.area32[i] = (float)((x2 - x1) * (y2 - y1));
        s.best[i] = 0ull;
        s.hits[i] = 0;
        s.floor[i] = 0u;
        // 'bg' figures never produce labels (utils.py:690); degenerate ones have IoU 0 (utils.py:103)
        uint8_t f = ((isbg != 0) || (x1 >= x2) || (y1 >= y2)) ? 1 : 0;
        // the float32 estimate is only trusted for pixel-scale coordinates
        if (!(img_w <= 8192.0 && img_h <= 8192.0) ||
            !(fabs(x1) <= 8192.0 && fabs(x2) <= 8192.0 && fabs(y1) <= 8192.0 && fabs(y2) <= 8192.0)) f |= 2;
        s.skip[i] = f;
    }
    for (int i = threadIdx.x; i < 4 * p.A; i += NT) s.use[i] = (i & 1) ? -1 : ((i & 2) ? p.H : p.W);
    for (int i = threadIdx.x; i < p.hash_slots; i += NT) {
        s.tkey[i] = 0xFFFFFFFFu;
        s.tmax[i] = 0ull;
        s.tg[i] = 0x7fffffff;
    }
    if (threadIdx.x == 0) { s.ctl[1] = 0; s.ctl[2] = 0; s.ctl[3] = 0; s.ctl[4] = 0; s.ctl[9] = 0; }
    __syncthreads();
    // A LOWER bound of every figure's best float32 IoU, from the exact IoU with the A anchors of the cell under
    // the figure's centre.  Pairs whose float32 estimate is below it by more than the margin cannot be (or tie
    // with) the best anchor.
    for (int i = threadIdx.x; i < G * p.A; i += NT) {
        const int g = i / p.A, a2 = i - g * p.A;
        if (s.skip[g] & 1) continue;
        const double gx1 = s.gt[4 * g + 0], gx2 = s.gt[4 * g + 1], gy1 = s.gt[4 * g + 2], gy2 = s.gt[4 * g + 3];
        int cx = (int)floor((gx1 + gx2) * 0.5 / p.stride), cy = (int)floor((gy1 + gy2) * 0.5 / p.stride);
        cx = min(max(cx, 0), p.W - 1);
        cy = min(max(cy, 0), p.H - 1);
        const AnchorPx c = anchor_px(p.stride, cx, cy, p.anchors.wh[a2][0], p.anchors.wh[a2][1]);
        const bool ok = !(c.x1 < 0.0 || c.x2 > img_w) && !(c.y1 < 0.0 || c.y2 > img_h);
        if (ok) {
            const float v = (float)ref_iou(gx1, gy1, gx2, gy2, c.x1, c.y1, c.x2, c.y2);
            if (v > 0.f) atomicMax(&s.floor[g], __float_as_uint(v));
        }
    }
    // anchor coordinates per shape and column / row (utils.py:625-626, 635-636) and the per-axis in-image tests
    // (utils.py:629, 638); an anchor is used when both its column and its row pass
    for (int i = threadIdx.x; i < p.A * (p.W + p.H); i += NT) {
        const int a = i / (p.W + p.H), r = i - a * (p.W + p.H);
        const bool isx = r < p.W;
        const int k = isx ? r : r - p.W;
        const double side = p.anchors.wh[a][isx ? 0 : 1], lim_px = isx ? img_w : img_h;
        const double c = __dmul_rn(p.stride, (double)k + 0.5);
        const double v1 = __dsub_rn(c, __dmul_rn(side, 0.5)), v2 = __dadd_rn(c, __dmul_rn(side, 0.5));
        const bool ok = !(v1 < 0.0 || v2 > lim_px) && v1 < v2;
        (isx ? s.ax + a * p.W : s.ay + a * p.H)[k] = make_double2(v1, v2);
        (isx ? s.axf + a * p.W : s.ayf + a * p.H)[k] = make_float4((float)v1, (float)v2, (float)(v2 - v1), ok ? 1.f : 0.f);
        if (ok) {
            atomicMin(&s.use[4 * a + (isx ? 0 : 2)], k);
            atomicMax(&s.use[4 * a + (isx ? 1 : 3)], k);
        }
    }
    __syncthreads();
    TGT_STAMP(1);

    const float thr32 = (float)p.max_overlap;
    const double inv_stride = 1.0 / p.stride;
    const int n_items = p.A * G;
    // Cell window of anchor shape a that can matter for figure g.  IoU >= L needs, on each axis, an overlap of at
    // least L*max(figure side, anchor side) (because union >= the larger area and the other overlap <= the smaller
    // side); with L = min(floor, thr) - margin this is a handful of cells around the figure.  The window is empty
    // when the two shapes cannot reach L at all (IoU <= smaller-overlap-box / union), and it is clipped to the
    // in-image rectangle of the shape.  One cell of padding absorbs the rounding of this float64 arithmetic.
    for (int it = threadIdx.x; it < n_items; it += NT) {
        const int a = it / G, g = it - a * G;
        const double aw = p.anchors.wh[a][0], ah = p.anchors.wh[a][1];
        const int *use = s.use + 4 * a;
        int4 r = make_int4(0, -1, 0, -1);                                    // empty
        const uint8_t f = s.skip[g];
        if (!(f & 1) && use[0] <= use[1] && use[2] <= use[3]) {
            const double gx1 = s.gt[4 * g + 0], gx2 = s.gt[4 * g + 1], gy1 = s.gt[4 * g + 2], gy2 = s.gt[4 * g + 3];
            double L = (double)fminf(__uint_as_float(s.floor[g]), thr32) - 2.0 * (double)kIouMargin;
            if ((f & 2) || !(aw <= 8192.0 && ah <= 8192.0) || !(L > 0.0)) L = 0.0;
            const double wg = gx2 - gx1, hg = gy2 - gy1;
            const double imax = fmin(wg, aw) * fmin(hg, ah);                  // largest possible intersection
            const bool feasible = imax >= (L - 1e-9) * (wg * hg + aw * ah - imax);
            const double mx = L * fmax(wg, aw), my = L * fmax(hg, ah);
            // centre c = stride*(i+0.5) must satisfy  g1 + m - side/2 <= c <= g2 - m + side/2
            const double xl = (gx1 + mx - aw * 0.5) * inv_stride - 0.5, xh = (gx2 - mx + aw * 0.5) * inv_stride - 0.5;
            const double yl = (gy1 + my - ah * 0.5) * inv_stride - 0.5, yh = (gy2 - my + ah * 0.5) * inv_stride - 0.5;
            if (feasible && xl <= xh + 2.0 && yl <= yh + 2.0) {
                r.x = max((int)fmax(floor(xl) - 1.0, 0.0), use[0]);
                r.y = min((int)fmin(ceil(xh) + 1.0, (double)(p.W - 1)), use[1]);
                r.z = max((int)fmax(floor(yl) - 1.0, 0.0), use[2]);
                r.w = min((int)fmin(ceil(yh) + 1.0, (double)(p.H - 1)), use[3]);
            }
        }
        s.range[it] = r;
        s.pstart[it + 1] = (r.x > r.y || r.z > r.w) ? 0 : (r.y - r.x + 1) * (r.w - r.z + 1);
    }
    __syncthreads();
    scan_items<NT>(s, n_items);
    TGT_STAMP(2);

    // Phase 1 - two passes over the candidate pairs (anchor of a window, figure).  Pass A, every pair, float32 only:
    // a cheap estimate of the IoU decides whether the exact value can matter at all; the pairs where it can are
    // queued.  Pass B, dense over the queue: exact float64 IoU, the figure's best anchor, and for IoU above
    // rpn_max_overlap a hit with its regression targets.  (One divergent pass costs every warp the float64 path
    // on every step.)  Which figure wins an anchor is settled in phase 2, so the figure order of the reference
    // ("first figure wins ties", utils.py:710-713) does not serialise anything.
    const int n_pairs = n_items ? s.pstart[n_items] : 0;
    auto exact_pair = [&](int it, int t) {
        const int a = it / G, g = it - a * G;
        const int4 rg = s.range[it];
        const int ww = rg.y - rg.x + 1;
        const int dy = t / ww;
        const int ix = rg.x + t - dy * ww, jy = rg.z + dy;
        const double gx1 = s.gt[4 * g + 0], gx2 = s.gt[4 * g + 1], gy1 = s.gt[4 * g + 2], gy2 = s.gt[4 * g + 3];
        const double2 X = s.ax[a * p.W + ix], Y = s.ay[a * p.H + jy];
        const double iou = ref_iou(gx1, gy1, gx2, gy2, X.x, Y.x, X.y, Y.y);
        const float iou32 = (float)iou;                                       // float32 accumulator (utils.py:603)
        if (iou32 > 0.f) {
            // best anchor of this figure: max float32 IoU, then first in loop order size->ratio->ix->jy.  The
            // 64-bit shared-memory max is a compare-and-swap loop: only candidates that beat the value seen go in.
            const unsigned order = (unsigned)((a * p.W + ix) * p.H + jy);
            const unsigned long long key = ((unsigned long long)__float_as_uint(iou32) << 32) | (0xFFFFFFFFu - order);
            if (key > *reinterpret_cast<volatile unsigned long long *>(&s.best[g])) atomicMax(&s.best[g], key);
        }
        if (iou > p.max_overlap) {                                            // utils.py:704
            atomicAdd(&s.hits[g], 1);
            const int pos = atomicAdd(&s.ctl[2], 1);
            if (pos < hit_cap) {
                TargetHit h;
                h.iou = iou; h.key = a * HW + jy * p.W + ix; h.g = g;
                positive_values(p, a, jy * p.W + ix, s.gt + 4 * g, false, h.v);
                s.hit[pos] = h;
            }
        }
    };
    {
        const int per_thread = (n_pairs + NT - 1) / NT;
        int q = min((int)threadIdx.x * per_thread, n_pairs);
        const int q_end = min(q + per_thread, n_pairs);
        int it = 0;                                                           // last item with pstart[it] <= q
        if (q < q_end)
            for (int hi = n_items - 1; it < hi;) {
                const int mid = (it + hi + 1) >> 1;
                if (s.pstart[mid] <= q) it = mid; else hi = mid - 1;
            }
        int it_end = q < q_end ? s.pstart[it + 1] : 0;
#pragma unroll 1
        for (; q < q_end; ++q) {
            while (q >= it_end) { ++it; it_end = s.pstart[it + 1]; }          // next non-empty item
            const int a = it / G, g = it - a * G;
            const int4 rg = s.range[it];
            const int ww = rg.y - rg.x + 1;
            const int t = q - s.pstart[it];
            const int dy = t / ww;
            const int ix = rg.x + t - dy * ww, jy = rg.z + dy;
            const float4 XF = s.axf[a * p.W + ix], YF = s.ayf[a * p.H + jy];
            // anchors crossing the image are skipped entirely (utils.py:629,638); a degenerate anchor has IoU 0
            if (XF.w == 0.f || YF.w == 0.f) continue;
            const float4 gf = s.gt32[g];
            const bool trusted = !(s.skip[g] & 2) && XF.z <= 8192.f && YF.z <= 8192.f;
            if (trusted) {
                // float32 estimate of the IoU (error < 3e-4 at pixel scale, margin 2e-3): pairs that cannot be the
                // figure's best anchor nor exceed rpn_max_overlap are dropped, including the disjoint ones (est = 0)
                const float wi = fminf(gf.z, XF.y) - fmaxf(gf.x, XF.x);
                const float hi32 = fminf(gf.w, YF.y) - fmaxf(gf.y, YF.x);
                const float itf = fmaxf(wi, 0.f) * fmaxf(hi32, 0.f);
                const float est = __fdividef(itf, s.area32[g] + XF.z * YF.z - itf);
                const float lim = fminf(__uint_as_float(s.floor[g]), thr32);
                if (!(est + kIouMargin >= lim)) continue;
            }
            const int slot = atomicAdd(&s.ctl[9], 1);
            if (slot < kNeedCap) s.need[slot] = q;
            else exact_pair(it, t);                                           // queue full: resolve in place
        }
    }
    __syncthreads();
    {
        const int n_need = min(s.ctl[9], kNeedCap);
#pragma unroll 1
        for (int e = threadIdx.x; e < n_need; e += NT) {
            const int q = s.need[e];
            int it = 0;
            for (int hi = n_items - 1; it < hi;) {
                const int mid = (it + hi + 1) >> 1;
                if (s.pstart[mid] <= q) it = mid; else hi = mid - 1;
            }
            exact_pair(it, q - s.pstart[it]);
        }
    }
    __syncthreads();
    TGT_STAMP(3);

    // Phase 2 - settle every hit anchor: highest IoU wins, equal IoU -> the earlier figure (strict '>' in figure
    // order, utils.py:710-713).  Open-addressing table keyed by the anchor; three passes; losers are flagged in
    // place.  The hits carry their regression targets already: once the panel is filled only stores are left.
    const int n_hit = s.ctl[2];
    const bool replay = n_hit > hit_cap;
    if (!replay) {
        for (int e = threadIdx.x; e < n_hit; e += NT) {
            const int key = s.hit[e].key;
            uint32_t slot = hash_key((uint32_t)key) & hmask;
            while (true) {
                const uint32_t prev = atomicCAS(&s.tkey[slot], 0xFFFFFFFFu, (uint32_t)key);
                if (prev == 0xFFFFFFFFu || prev == (uint32_t)key) break;
                slot = (slot + 1) & hmask;
            }
            atomicMax(&s.tmax[slot], (unsigned long long)__double_as_longlong(s.hit[e].iou));   // positive doubles order like their bits
        }
    }
    // forced positives + best_anchor table (utils.py:741-766): decode the best anchor of every figure
    for (int g = NT - 1 - (int)threadIdx.x; g < p.Gmax; g += NT) {
        const unsigned long long key = g < G ? s.best[g] : 0ull;
        const int nh = g < G ? s.hits[g] : 0;
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
        s.order[g] = order;
    }
    __syncthreads();
    if (!replay) {
        for (int e = threadIdx.x; e < n_hit; e += NT) {
            const int key = s.hit[e].key;
            uint32_t slot = hash_key((uint32_t)key) & hmask;
            while (s.tkey[slot] != (uint32_t)key) slot = (slot + 1) & hmask;
            if ((unsigned long long)__double_as_longlong(s.hit[e].iou) == s.tmax[slot]) atomicMin(&s.tg[slot], s.hit[e].g);
        }
    }
    // The reference applies the forced positives in GT order, so when several GT share the same best anchor the
    // LAST one wins: a figure is only kept if no later forced figure targets its anchor.  (Threads are taken from
    // the top so that this float64 work runs next to the table passes, not after them.)
    for (int g = NT - 1 - (int)threadIdx.x; g < G; g += NT) {
        const unsigned o = s.order[g];
        if (o == 0xFFFFFFFFu) continue;
        bool last = true;
        for (int g2 = g + 1; g2 < G; ++g2) last = last && (s.order[g2] != o);
        if (!last) continue;
        const int jy = (int)(o % (unsigned)p.H);
        const unsigned rest = o / (unsigned)p.H;
        const int ix = (int)(rest % (unsigned)p.W);
        const int a2 = (int)(rest / (unsigned)p.W);
        const int pos = atomicAdd(&s.ctl[4], 1);
        s.wkey[pos] = a2 * HW + jy * p.W + ix;
        positive_values(p, a2, jy * p.W + ix, s.gt + 4 * g, true, s.wv + 4 * pos);
    }
    __syncthreads();
    if (!replay) {
        for (int e = threadIdx.x; e < n_hit; e += NT) {
            const int key = s.hit[e].key, g = s.hit[e].g;
            uint32_t slot = hash_key((uint32_t)key) & hmask;
            while (s.tkey[slot] != (uint32_t)key) slot = (slot + 1) & hmask;
            const bool win = (unsigned long long)__double_as_longlong(s.hit[e].iou) == s.tmax[slot] && s.tg[slot] == g;
            if (!win) s.hit[e].g = -1 - g;
        }
    }
    TGT_STAMP(4);

    // ---- wait until every item of this panel has been filled; fill meanwhile if there is anything left ---------
    if (p.role == 0) {
        const int gr = b / kGroupPanels;
        const int want = 5 * AHW * min(kGroupPanels, p.B - gr * kGroupPanels);
        const long long t_start = global_ns();
        bool more = true;
        while (true) {
            __syncthreads();
            if (threadIdx.x == 0) {
                publish_fill(p, s);                               // never wait on items this CTA still holds back
                const int done = ld_acquire(&p.group_done[gr]);
                s.ctl[1] = done >= want;
                if (!s.ctl[1] && global_ns() - t_start > 4000000000LL) s.ctl[1] = 2;
            }
            __syncthreads();
            if (s.ctl[1]) break;
            // the parked positives live in wv / wkey / ctl[3..4]; a fill unit only touches the in-image tables
            if (more) more = pull_fill<NT>(p, s);
        }
        if (s.ctl[1] == 2)         // the fill never completed (4 s): results invalid, reported through n_hits
            for (int g = threadIdx.x; g < p.Gmax; g += NT) p.n_hits[(size_t)b * p.Gmax + g] = -1;
    } else {
        __syncthreads();
    }
    TGT_STAMP(5);

    // regular positives (utils.py:728-738)
    int n_win = 0;
    if (!replay) {
        for (int e = threadIdx.x; e < n_hit; e += NT) {
            const TargetHit &h = s.hit[e];
            if (h.g < 0) continue;
            const int a2 = h.key / HW;
            store_positive(p, cls_b, regr_b, a2, h.key - a2 * HW, h.v, false);
            ++n_win;
        }
    } else {
        // more positives than the list holds (never seen in practice): shape by shape, figure by figure with
        // in-place per-cell state (the hit list and the table are not needed any more)
        double *s_lb = reinterpret_cast<double *>(s.hit);                     // [HW]
        int *s_lg = reinterpret_cast<int *>(s_lb + HW);                       // [HW]
#pragma unroll 1
        for (int a = 0; a < p.A; ++a) {
            __syncthreads();
            for (int cell = threadIdx.x; cell < HW; cell += NT) { s_lb[cell] = 0.0; s_lg[cell] = -1; }
            __syncthreads();
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                const int4 rg = s.range[a * G + g];
                if (rg.x > rg.y || rg.z > rg.w) continue;                     // block-uniform
                const int ww = rg.y - rg.x + 1, n = ww * (rg.w - rg.z + 1);
                const double gx1 = s.gt[4 * g + 0], gx2 = s.gt[4 * g + 1], gy1 = s.gt[4 * g + 2], gy2 = s.gt[4 * g + 3];
                for (int t = threadIdx.x; t < n; t += NT) {
                    const int dy = t / ww, ix = rg.x + t - dy * ww, jy = rg.z + dy;
                    if (s.axf[a * p.W + ix].w == 0.f || s.ayf[a * p.H + jy].w == 0.f) continue;
                    const double2 X = s.ax[a * p.W + ix], Y = s.ay[a * p.H + jy];
                    const double iou = ref_iou(gx1, gy1, gx2, gy2, X.x, Y.x, X.y, Y.y);
                    const int cell = jy * p.W + ix;
                    if (iou > p.max_overlap && iou > s_lb[cell]) { s_lb[cell] = iou; s_lg[cell] = g; }
                }
                __syncthreads();
            }
            for (int cell = threadIdx.x; cell < HW; cell += NT) {
                const int lg = s_lg[cell];
                if (lg >= 0) {
                    double v[4];
                    positive_values(p, a, cell, s.gt + 4 * lg, false, v);
                    store_positive(p, cls_b, regr_b, a, cell, v, false);
                }
            }
        }
    }
    __syncthreads();                                                          // regular before forced writes
    const int n_forced = s.ctl[4];
    for (int e = threadIdx.x; e < n_forced; e += NT) {
        const int key = s.wkey[e], a2 = key / HW;
        store_positive(p, cls_b, regr_b, a2, key - a2 * HW, s.wv + 4 * e, true);
    }
#ifdef RADNET_TGT_PROFILE
    if (p.stamps && threadIdx.x == 0) {
        p.stamps[(size_t)blockIdx.x * 16 + 10] = n_pairs;
        p.stamps[(size_t)blockIdx.x * 16 + 11] = n_hit;
        p.stamps[(size_t)blockIdx.x * 16 + 12] = n_win;
        p.stamps[(size_t)blockIdx.x * 16 + 13] = n_forced;
    }
#endif
    TGT_STAMP(6);
}

// next panel from the shared counter; false when none is left
template <int NT>
__device__ bool pull_panel(const RpnTargetParams &p, const TgtShared &s) {
    __syncthreads();
    if (threadIdx.x == 0) s.ctl[0] = atomicAdd(&p.ctl[1], 1);
    __syncthreads();
    const int b = s.ctl[0];
    if (b >= p.B) return false;
    do_panel<NT>(p, s, b);
    return true;
}

template <int NT>
__global__ void __launch_bounds__(NT, 2048 / NT / 2) rpn_targets_kernel(RpnTargetParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_ctl[12];
    __shared__ int s_warp[34];
    const TgtShared s = carve(p, smem, s_ctl, s_warp);
    if (threadIdx.x == 0) { s_ctl[6] = -1; s_ctl[7] = -1; s_ctl[8] = 0; }
#ifdef RADNET_TGT_PROFILE
    if (p.stamps && threadIdx.x == 0) p.stamps[(size_t)blockIdx.x * 16 + 8] = global_ns();
#endif
    if (p.role == 1) {
        while (pull_fill<NT>(p, s)) {}
    } else if (p.role == 2) {
        while (pull_panel<NT>(p, s)) {}
    } else if (((long long)sm_id() * p.n_compute_sm) % p.n_sm < p.n_compute_sm) {     // n_compute_sm SMs, evenly spread
        while (pull_panel<NT>(p, s)) {}
        while (pull_fill<NT>(p, s)) {}
    } else {
        while (pull_fill<NT>(p, s)) {}
#ifdef RADNET_TGT_PROFILE
        if (p.stamps && threadIdx.x == 0) p.stamps[(size_t)blockIdx.x * 16 + 9] = global_ns();
#endif
        while (pull_panel<NT>(p, s)) {}
    }
    // ---- the last CTA out resets the launch-wide counters ----------------------------------------------------
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int prev = atomicAdd(&p.ctl[2], 1);
        if (prev == (int)gridDim.x - 1) {
            p.ctl[0] = 0;
            p.ctl[1] = 0;
            p.ctl[2] = 0;
            if (p.role != 1)
                for (int g = 0; g < (p.B + kGroupPanels - 1) / kGroupPanels; ++g) p.group_done[g] = 0;
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
    l.total = l.off_win + gm * (4 * sizeof(double) + sizeof(int)) + (size_t)kNeedCap * sizeof(int) + 16;
    return l;
}
int tgt_groups(int B) { return (B + kGroupPanels - 1) / kGroupPanels; }
size_t tgt_ws_bytes(int B) { return align_up(((size_t)tgt_groups(B) + 4) * sizeof(int32_t), 256); }
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
    int hit_cap = 512;
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
    p.group_done = reinterpret_cast<int32_t *>(ws);
    p.ctl = p.group_done + tgt_groups(B);
    p.sm_off_tables = (int)sl.off_tables; p.sm_off_items = (int)sl.off_items; p.sm_off_hits = (int)sl.off_hits;
    p.sm_off_hash = (int)sl.off_hash; p.sm_off_win = (int)sl.off_win; p.hit_cap = sl.hit_cap;
    p.hash_slots = sl.hash_slots; p.n_items_max = sl.n_items_max;
#ifdef RADNET_TGT_PROFILE
    p.stamps = g_tgt_stamps;
#endif
    cudaStream_t st = (cudaStream_t)stream;
    // Two launch shapes.  Few panels (one round of at most ~43 % of the SMs): one CTA of 1024 threads per SM, so
    // that a panel has a whole SM's issue slots and the round is short.  Many panels: two CTAs of 512 threads per
    // SM - two panels per computing SM hide each other's latencies and more SMs are left for the fill.
    const bool two_per_sm = 2 * (sl.total + 1024) <= (size_t)smem_limit + 1024;
    long long n_comp_sm = get_option(kOptTargetsComputeCtas);
    const bool wide = !two_per_sm || (n_comp_sm < 1 && B <= (n_sm * 43 + 99) / 100);
    const int per_sm = wide ? 1 : 2;
    if (n_comp_sm < 1) n_comp_sm = wide ? (n_sm * 43 + 99) / 100 : (n_sm * 22 + 99) / 100;
    if (n_comp_sm > (B + per_sm - 1) / per_sm) n_comp_sm = (B + per_sm - 1) / per_sm;
    if (n_comp_sm > n_sm - 1) n_comp_sm = n_sm > 1 ? n_sm - 1 : 1;
    p.n_compute_sm = (int)n_comp_sm;
    p.n_sm = n_sm;
    const long long n_units = (long long)B * kUnitsPerPanel;
    long long grid = (long long)n_sm * per_sm;
    if (grid > n_units + B) grid = n_units + B;
    auto launch = [&](unsigned g) -> int {
        if (wide) {
            if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(rpn_targets_kernel<1024>), dev, sl.total)) return rc;
            rpn_targets_kernel<1024><<<g, 1024, sl.total, st>>>(p);
        } else {
            if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(rpn_targets_kernel<512>), dev, sl.total)) return rc;
            rpn_targets_kernel<512><<<g, 512, sl.total, st>>>(p);
        }
        return check_launch("rpn_targets_kernel");
    };
    if (get_option(kOptTargetsTwoLaunches) == 1) {
        // no co-residency assumed: the fill as its own launch, then the panels (stream order replaces the wait)
        p.role = 1;
        if (int rc = launch((unsigned)(grid < n_units ? grid : n_units))) return rc;
        p.role = 2;
        return launch((unsigned)(grid < B ? grid : B));
    }
    p.role = 0;
    return launch((unsigned)grid);
}

// targets.cu - K3: RPN anchor target assignment (reference faster_rcnn/utils.py:554-775,
// 815-816; upstream name calc_rpn) and a4: RoI target assignment (reference
// faster_rcnn/rpn.py:209-282).
//
// K3 layout: one thread per anchor (a, jy, ix) with ix fastest, so each of the 10 float64
// output planes an anchor touches is written with coalesced 8-byte stores.  The GT boxes
// of the panel sit in shared memory.  The kernel is bound by the float64 output stream
// (10*A*H*W*8 B per panel); IoU work is skipped for non-intersecting pairs (IoU == 0.0
// exactly) so the float64 divide only runs where boxes meet.
//
// Order-dependent reference semantics and how they are kept without a serial loop:
//   * best anchor per GT = first anchor, in the reference's loop order
//     size -> ratio -> ix -> jy (utils.py:616-632), whose float32-rounded IoU is the
//     maximum (float32 accumulator utils.py:603; NumPy>=2 compares in float32, see
//     SURVEY.md row a3').  Realised as a 64-bit atomicMax on
//     (float32 bits of IoU) << 32 | (0xFFFFFFFF - loop_order).
//   * per-anchor best GT: strict '>' from 0.0, first GT wins ties (utils.py:710-713) - a
//     serial loop over the GT inside the thread.
//   * forced positives (utils.py:741-766) are applied by a second tiny kernel in GT order,
//     with the float32-rounded targets (utils.py:605,766).
#include <stdlib.h>

#include "common.cuh"

namespace radnet {

constexpr int kTgtThreads = 256;

struct RpnTargetParams {
    const double *gt;          // [B][Gmax][4] x1,x2,y1,y2
    const uint8_t *gt_is_bg;   // [B][Gmax]
    const int32_t *gt_count;   // [B]
    int Gmax, H, W, A, n_ratios;
    AnchorTable anchors;       // pixels
    double stride;
    const double *img_wh;      // [B][2]
    double max_overlap;
    double *y_cls;             // [B][2A][H][W]
    double *y_regr;            // [B][8A][H][W]
    int32_t *best_anchor;      // [B][Gmax][4]
    int32_t *n_hits;           // [B][Gmax]
    unsigned long long *best_key;  // [B][Gmax] workspace
    int sm_off_cells;          // shared-memory offset of the per-cell state
    int hit_cap;               // capacity of the positive-cell list
};

struct TargetHit {
    double iou;
    int cell;
    int g;
};

// reference utils.py:77-109 with a = GT (x1,y1,x2,y2), b = anchor
__device__ __forceinline__ double ref_iou(double ax1, double ay1, double ax2, double ay2, double bx1,
                                          double by1, double bx2, double by2) {
    if (ax1 >= ax2 || ay1 >= ay2 || bx1 >= bx2 || by1 >= by2) return 0.0;
    double x = fmax(ax1, bx1), y = fmax(ay1, by1);
    double w = __dsub_rn(fmin(ax2, bx2), x), h = __dsub_rn(fmin(ay2, by2), y);
    if (w < 0.0 || h < 0.0) return 0.0;
    double inter = __dmul_rn(w, h);
    if (inter == 0.0) return 0.0;
    double area_a = __dmul_rn(__dsub_rn(ax2, ax1), __dsub_rn(ay2, ay1));
    double area_b = __dmul_rn(__dsub_rn(bx2, bx1), __dsub_rn(by2, by1));
    double uni = __dsub_rn(__dadd_rn(area_a, area_b), inter);
    return __ddiv_rn(inter, __dadd_rn(uni, 1e-6));
}

struct AnchorPx { double x1, x2, y1, y2; };

__device__ __forceinline__ AnchorPx anchor_px(double stride, int ix, int jy, double aw, double ah) {
    AnchorPx a;
    double cx = __dmul_rn(stride, (double)ix + 0.5), cy = __dmul_rn(stride, (double)jy + 0.5);
    a.x1 = __dsub_rn(cx, __ddiv_rn(aw, 2.0));      // utils.py:625
    a.x2 = __dadd_rn(cx, __ddiv_rn(aw, 2.0));      // utils.py:626
    a.y1 = __dsub_rn(cy, __ddiv_rn(ah, 2.0));      // utils.py:635
    a.y2 = __dadd_rn(cy, __ddiv_rn(ah, 2.0));      // utils.py:636
    return a;
}

// (tx,ty,tw,th) of utils.py:669-687
__device__ __forceinline__ void regr_targets(const AnchorPx &a, double gx1, double gx2, double gy1,
                                             double gy2, double t[4]) {
    double cx = __ddiv_rn(__dadd_rn(gx1, gx2), 2.0), cy = __ddiv_rn(__dadd_rn(gy1, gy2), 2.0);
    double cxa = __ddiv_rn(__dadd_rn(a.x1, a.x2), 2.0), cya = __ddiv_rn(__dadd_rn(a.y1, a.y2), 2.0);
    double wa = __dsub_rn(a.x2, a.x1), ha = __dsub_rn(a.y2, a.y1);
    t[0] = __ddiv_rn(__dsub_rn(cx, cxa), wa);
    t[1] = __ddiv_rn(__dsub_rn(cy, cya), ha);
    t[2] = log(__ddiv_rn(__dsub_rn(gx2, gx1), wa));
    t[3] = log(__ddiv_rn(__dsub_rn(gy2, gy1), ha));
}

// Margin of the float32 IoU estimate used to skip work that cannot matter (see below).  The
// estimate is off by < 3e-4 absolute for boxes up to a few thousand pixels; 2e-3 is generous.
constexpr float kIouMargin = 2e-3f;

// One CTA per (panel, anchor shape): the figures and their filters are set up once and the CTA
// then walks the H*W cells of its anchor plane in chunks of kTgtThreads.
__global__ void __launch_bounds__(kTgtThreads, 4) rpn_targets_kernel(RpnTargetParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *s_gt = reinterpret_cast<double *>(smem);                                   // [G][4] x1,x2,y1,y2
    float4 *s_gt32 = reinterpret_cast<float4 *>(s_gt + 4 * p.Gmax);                      // [G] x1,y1,x2,y2 rounded
    int4 *s_range = reinterpret_cast<int4 *>(s_gt32 + p.Gmax);                          // [G] ix_lo, ix_hi, jy_lo, jy_hi
    unsigned long long *s_best = reinterpret_cast<unsigned long long *>(s_range + p.Gmax);
    float *s_area32 = reinterpret_cast<float *>(s_best + p.Gmax);
    unsigned *s_floor = reinterpret_cast<unsigned *>(s_area32 + p.Gmax);                // [G] lower bound of the best IoU (f32 bits)
    int *s_hits = reinterpret_cast<int *>(s_floor + p.Gmax);
    uint8_t *s_skip = reinterpret_cast<uint8_t *>(s_hits + p.Gmax);                     // bit0: bg/degenerate, bit1: no filter

    const int b = blockIdx.y, a = blockIdx.x;
    const int HW = p.H * p.W;
    const int G = p.gt_count[b];
    const double aw = p.anchors.wh[a][0], ah = p.anchors.wh[a][1];
    const double img_w = p.img_wh[2 * b], img_h = p.img_wh[2 * b + 1];
    double *cls_b = p.y_cls + (size_t)b * 2 * p.A * HW;
    double *regr_b = p.y_regr + (size_t)b * 8 * p.A * HW;
    const int lane = threadIdx.x & 31;

    const bool coords_small = img_w <= 8192.0 && img_h <= 8192.0 && aw <= 8192.0 && ah <= 8192.0;
    for (int i = threadIdx.x; i < G; i += kTgtThreads) {
        const double *q = p.gt + ((size_t)b * p.Gmax + i) * 4;
        const double x1 = q[0], x2 = q[1], y1 = q[2], y2 = q[3];
        s_gt[4 * i + 0] = x1; s_gt[4 * i + 1] = x2; s_gt[4 * i + 2] = y1; s_gt[4 * i + 3] = y2;
        s_gt32[i] = make_float4((float)x1, (float)y1, (float)x2, (float)y2);
        s_area32[i] = (float)((x2 - x1) * (y2 - y1));
        s_best[i] = 0ull;
        s_hits[i] = 0;
        s_floor[i] = 0u;
        // 'bg' figures never produce labels (utils.py:690); degenerate ones have IoU 0 (utils.py:103)
        uint8_t f = ((p.gt_is_bg[(size_t)b * p.Gmax + i] != 0) || (x1 >= x2) || (y1 >= y2)) ? 1 : 0;
        // the float32 estimate is only trusted for pixel-scale coordinates
        if (!coords_small || !(fabs(x1) <= 8192.0 && fabs(x2) <= 8192.0 && fabs(y1) <= 8192.0 && fabs(y2) <= 8192.0)) f |= 2;
        s_skip[i] = f;
    }
    __syncthreads();
    // A LOWER bound of every figure's best float32 IoU, from the exact IoU with the A anchors of
    // the cell under the figure's centre (all anchor shapes, not only this CTA's).  Pairs whose
    // float32 estimate is below it by more than the margin cannot be (or tie with) the best anchor.
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
    __syncthreads();

    const float thr32 = (float)p.max_overlap;
    // per-cell state of this anchor plane: best IoU above rpn_max_overlap and the figure it came from
    double *s_lb = reinterpret_cast<double *>(smem + p.sm_off_cells);                    // [HW]
    double2 *s_ax = reinterpret_cast<double2 *>(s_lb + ((HW + 1) & ~1));                 // [W] anchor x1,x2 of column ix
    double2 *s_ay = s_ax + p.W;                                                          // [H] anchor y1,y2 of row jy
    float4 *s_axf = reinterpret_cast<float4 *>(s_ay + p.H);                              // [W] x1,x2,width (f32), in-image flag
    float4 *s_ayf = s_axf + p.W;                                                         // [H]
    TargetHit *s_hit = reinterpret_cast<TargetHit *>(s_ayf + p.H);                       // [hit_cap]
    const int hit_cap = p.hit_cap;
    int *s_nhit = reinterpret_cast<int *>(s_hit + hit_cap);                              // 4 ints
    short *s_lg = reinterpret_cast<short *>(s_nhit + 8);                                 // [HW]
    int *s_use = s_nhit + 4;                                                             // ix_lo, ix_hi, jy_lo, jy_hi in-image
    if (threadIdx.x == 0) {
        *s_nhit = 0;
        s_use[0] = p.W; s_use[1] = -1; s_use[2] = p.H; s_use[3] = -1;
    }
    __syncthreads();
    for (int cell = threadIdx.x; cell < HW; cell += kTgtThreads) {
        s_lb[cell] = 0.0;
        s_lg[cell] = -1;
    }
    // anchor coordinates per column / row (utils.py:625-626, 635-636) and the per-axis in-image tests
    // (utils.py:629, 638); an anchor is used when both its column and its row pass
    for (int i = threadIdx.x; i < p.W + p.H; i += kTgtThreads) {
        const bool isx = i < p.W;
        const int k = isx ? i : i - p.W;
        const double side = isx ? aw : ah, lim_px = isx ? img_w : img_h;
        const double c = __dmul_rn(p.stride, (double)k + 0.5);
        const double v1 = __dsub_rn(c, __dmul_rn(side, 0.5)), v2 = __dadd_rn(c, __dmul_rn(side, 0.5));
        const bool ok = !(v1 < 0.0 || v2 > lim_px) && v1 < v2;
        (isx ? s_ax : s_ay)[k] = make_double2(v1, v2);
        (isx ? s_axf : s_ayf)[k] = make_float4((float)v1, (float)v2, (float)(v2 - v1), ok ? 1.f : 0.f);
        if (ok) {
            atomicMin(&s_use[isx ? 0 : 2], k);
            atomicMax(&s_use[isx ? 1 : 3], k);
        }
    }
    __syncthreads();

    // Cell window of this anchor shape that can matter for each figure.  IoU >= L needs, on each
    // axis, an overlap of at least L*max(figure side, anchor side) (because union >= the larger
    // area and the other overlap <= the smaller side); with L = min(floor, thr) - margin this is
    // a handful of cells around the figure.  The window is empty when the two shapes cannot reach
    // L at all (IoU <= smaller-overlap-box / union), and it is clipped to the in-image rectangle
    // of this anchor shape.  One cell of padding absorbs the rounding of this float64 arithmetic.
    for (int g = threadIdx.x; g < G; g += kTgtThreads) {
        int4 r = make_int4(0, -1, 0, -1);                                    // empty
        const uint8_t f = s_skip[g];
        if (!(f & 1) && s_use[0] <= s_use[1] && s_use[2] <= s_use[3]) {
            const double gx1 = s_gt[4 * g + 0], gx2 = s_gt[4 * g + 1], gy1 = s_gt[4 * g + 2], gy2 = s_gt[4 * g + 3];
            double L = (double)fminf(__uint_as_float(s_floor[g]), thr32) - 2.0 * (double)kIouMargin;
            if ((f & 2) || !(L > 0.0)) L = 0.0;
            const double wg = gx2 - gx1, hg = gy2 - gy1;
            const double imax = fmin(wg, aw) * fmin(hg, ah);                  // largest possible intersection
            const bool feasible = imax >= (L - 1e-9) * (wg * hg + aw * ah - imax);
            const double mx = L * fmax(wg, aw), my = L * fmax(hg, ah);
            // centre c = stride*(i+0.5) must satisfy  g1 + m - side/2 <= c <= g2 - m + side/2
            const double xl = (gx1 + mx - aw * 0.5) / p.stride - 0.5, xh = (gx2 - mx + aw * 0.5) / p.stride - 0.5;
            const double yl = (gy1 + my - ah * 0.5) / p.stride - 0.5, yh = (gy2 - my + ah * 0.5) / p.stride - 0.5;
            if (feasible && xl <= xh + 2.0 && yl <= yh + 2.0) {
                r.x = max((int)fmax(floor(xl) - 1.0, 0.0), s_use[0]);
                r.y = min((int)fmin(ceil(xh) + 1.0, (double)(p.W - 1)), s_use[1]);
                r.z = max((int)fmax(floor(yl) - 1.0, 0.0), s_use[2]);
                r.w = min((int)fmin(ceil(yh) + 1.0, (double)(p.H - 1)), s_use[3]);
            }
        }
        s_range[g] = r;
    }
    __syncthreads();

    // Phase 1 - one WARP per figure (8 figures in flight per CTA; a CTA-wide pass per figure was
    // latency-bound on the float64 divide and the block barrier).  The warp enumerates the cells
    // of the figure's window 32 at a time.  Cells whose IoU exceeds rpn_max_overlap go to a hit
    // list; which figure wins a cell is settled in phase 2, so the figure order of the reference
    // ("first figure wins ties", utils.py:710-713) does not serialise the warps.
    constexpr int kWarps = kTgtThreads / 32;
    const int w = threadIdx.x >> 5;
#pragma unroll 1
    for (int g = w; g < G; g += kWarps) {
        const int4 rg = s_range[g];
        if (rg.x > rg.y || rg.z > rg.w) continue;                             // warp-uniform (also bg / degenerate)
        const int ww = rg.y - rg.x + 1, n = ww * (rg.w - rg.z + 1);
        const uint8_t gflag = s_skip[g];
        const double gx1 = s_gt[4 * g + 0], gx2 = s_gt[4 * g + 1], gy1 = s_gt[4 * g + 2], gy2 = s_gt[4 * g + 3];
        const float4 gf = s_gt32[g];
        const float lim = fminf(__uint_as_float(s_floor[g]), thr32);
        unsigned long long best = 0ull;
        int nhit = 0;
#pragma unroll 1
        for (int t0 = 0; t0 < n; t0 += 32) {
            const int t = t0 + lane;
            const bool act = t < n;
            const int dy = act ? t / ww : 0;
            const int ix = rg.x + (act ? t - dy * ww : 0), jy = rg.z + dy;
            const double2 X = s_ax[ix], Y = s_ay[jy];
            const float4 XF = s_axf[ix], YF = s_ayf[jy];
            // anchors crossing the image are skipped entirely (utils.py:629,638); a degenerate anchor has IoU 0
            const bool usable = act && XF.w != 0.f && YF.w != 0.f;
            // IoU > 0  <=>  the open intervals meet on both axes (exact, float64 compares only)
            const bool isect = usable && gx2 > X.x && X.y > gx1 && gy2 > Y.x && Y.y > gy1;
            // float32 estimate of the IoU: decides whether the exact float64 value can matter at all
            bool need = false;
            if (isect) {
                const float wi = fminf(gf.z, XF.y) - fmaxf(gf.x, XF.x);
                const float hi = fminf(gf.w, YF.y) - fmaxf(gf.y, YF.x);
                const float it = fmaxf(wi, 0.f) * fmaxf(hi, 0.f);
                const float q = __fdividef(it, s_area32[g] + XF.z * YF.z - it);
                need = (q + kIouMargin >= lim) ||           // could be the best anchor, or exceed rpn_max_overlap
                       (gflag & 2);                         // estimate not trusted: always exact
            }
            if (!__any_sync(0xffffffffu, need)) continue;                     // warp-uniform
            unsigned bits = 0;
            bool hit = false;
            double iou = 0.0;
            if (need) {
                iou = ref_iou(gx1, gy1, gx2, gy2, X.x, Y.x, X.y, Y.y);
                const float iou32 = (float)iou;                               // float32 accumulator (utils.py:603)
                if (iou32 > 0.f) bits = __float_as_uint(iou32);
                hit = iou > p.max_overlap;                                    // utils.py:704
            }
            // best anchor of this figure: max float32 IoU, then first in loop order.  Two REDUX ops.
            const unsigned order = (unsigned)((a * p.W + ix) * p.H + jy);     // size->ratio->ix->jy
            const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
            if (wmax) {
                const unsigned omin = __reduce_min_sync(0xffffffffu, bits == wmax ? order : 0xFFFFFFFFu);
                const unsigned long long key = ((unsigned long long)wmax << 32) | (0xFFFFFFFFu - omin);
                best = key > best ? key : best;
            }
            const unsigned hm = __ballot_sync(0xffffffffu, hit);
            if (hm) {
                int base = 0;
                if (lane == 0) base = atomicAdd(s_nhit, __popc(hm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) {
                    const int pos = base + __popc(hm & lanemask_lt());
                    if (pos < hit_cap) s_hit[pos] = TargetHit{iou, jy * p.W + ix, g};
                }
                nhit += __popc(hm);
            }
        }
        if (lane == 0) {              // this warp is the only writer of figure g in this CTA
            s_best[g] = best;
            s_hits[g] = nhit;
        }
    }
    __syncthreads();

    // Phase 2 - settle every hit cell: highest IoU wins, equal IoU -> the earlier figure (strict '>'
    // in figure order, utils.py:710-713).
    const int n_hit = *s_nhit;
    if (n_hit <= hit_cap) {
        for (int e = threadIdx.x; e < n_hit; e += kTgtThreads) {
            const TargetHit h = s_hit[e];
            bool win = true;
            for (int j = 0; j < n_hit; ++j) {
                const TargetHit o = s_hit[j];
                if (o.cell == h.cell && (o.iou > h.iou || (o.iou == h.iou && o.g < h.g))) win = false;
            }
            if (win) s_lg[h.cell] = (short)h.g;
        }
    } else {
        // more positives than the list holds (never seen in practice): replay figure by figure with
        // in-place per-cell state, the whole CTA on one figure at a time
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            const int4 rg = s_range[g];
            if (rg.x > rg.y || rg.z > rg.w) continue;                         // block-uniform
            const int ww = rg.y - rg.x + 1, n = ww * (rg.w - rg.z + 1);
            const double gx1 = s_gt[4 * g + 0], gx2 = s_gt[4 * g + 1], gy1 = s_gt[4 * g + 2], gy2 = s_gt[4 * g + 3];
            for (int t = threadIdx.x; t < n; t += kTgtThreads) {
                const int dy = t / ww, ix = rg.x + t - dy * ww, jy = rg.z + dy;
                if (s_axf[ix].w == 0.f || s_ayf[jy].w == 0.f) continue;
                const double2 X = s_ax[ix], Y = s_ay[jy];
                const double iou = ref_iou(gx1, gy1, gx2, gy2, X.x, Y.x, X.y, Y.y);
                const int cell = jy * p.W + ix;
                if (iou > p.max_overlap && iou > s_lb[cell]) { s_lb[cell] = iou; s_lg[cell] = (short)g; }
            }
            __syncthreads();
        }
    }
    __syncthreads();

    // write-out: every cell of the anchor plane, 10 float64 planes, coalesced along ix
#pragma unroll 1
    for (int cell = threadIdx.x; cell < HW; cell += kTgtThreads) {
        const int jy = cell / p.W, ix = cell - jy * p.W;
        // `inside` is the per-axis test only; a degenerate anchor (side <= 0) never occurs in-image
        const bool inside = s_axf[ix].w != 0.f && s_ayf[jy].w != 0.f;
        // labels are written inside the GT loop of the reference: no GT, no labels (utils.py:722-738)
        const double valid = (inside && G > 0) ? 1.0 : 0.0;
        const int lg = s_lg[cell];
        const double ov = lg >= 0 ? 1.0 : 0.0;
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        if (lg >= 0) {
            AnchorPx an;
            const double2 X = s_ax[ix], Y = s_ay[jy];
            an.x1 = X.x; an.x2 = X.y; an.y1 = Y.x; an.y2 = Y.y;
            regr_targets(an, s_gt[4 * lg + 0], s_gt[4 * lg + 1], s_gt[4 * lg + 2], s_gt[4 * lg + 3], t);
        }
        cls_b[(size_t)a * HW + cell] = valid;
        cls_b[(size_t)(p.A + a) * HW + cell] = ov;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            regr_b[(size_t)(4 * a + k) * HW + cell] = ov;                       // np.repeat(overlap,4)
            regr_b[(size_t)(4 * p.A + 4 * a + k) * HW + cell] = t[k];
        }
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += kTgtThreads) {
        if (s_best[g]) atomicMax(&p.best_key[(size_t)b * p.Gmax + g], s_best[g]);
        if (s_hits[g]) atomicAdd(&p.n_hits[(size_t)b * p.Gmax + g], s_hits[g]);
    }
}

// forced positives + best_anchor table (utils.py:741-766).  One CTA per panel, one thread per GT.
// The reference applies the forced positives in GT order, so when several GT share the same
// best anchor the LAST one wins: a thread only writes if no later forced GT targets its anchor.
__global__ void __launch_bounds__(128) rpn_targets_finalize_kernel(RpnTargetParams p) {
    extern __shared__ unsigned s_order[];          // [Gmax] loop-order id of the forced anchor, or ~0
    const int b = blockIdx.x;
    const int HW = p.H * p.W;
    const int G = p.gt_count[b];
    double *cls_b = p.y_cls + (size_t)b * 2 * p.A * HW;
    double *regr_b = p.y_regr + (size_t)b * 8 * p.A * HW;
    for (int g = threadIdx.x; g < p.Gmax; g += blockDim.x) {
        int32_t *ba = p.best_anchor + ((size_t)b * p.Gmax + g) * 4;
        const unsigned long long key = (g < G) ? p.best_key[(size_t)b * p.Gmax + g] : 0ull;
        unsigned order = 0xFFFFFFFFu;
        if (!key) {
            ba[0] = ba[1] = ba[2] = ba[3] = -1;
        } else {
            const unsigned o = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
            const int jy = (int)(o % (unsigned)p.H);
            const unsigned rest = o / (unsigned)p.H;
            const int ix = (int)(rest % (unsigned)p.W);
            const int a = (int)(rest / (unsigned)p.W);
            ba[0] = jy; ba[1] = ix; ba[2] = a % p.n_ratios; ba[3] = a / p.n_ratios;     // utils.py:697
            if (p.n_hits[(size_t)b * p.Gmax + g] == 0) order = o;
        }
        s_order[g] = order;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const unsigned o = s_order[g];
        if (o == 0xFFFFFFFFu) continue;
        bool last = true;
        for (int g2 = g + 1; g2 < G; ++g2) last = last && (s_order[g2] != o);
        if (!last) continue;
        const int jy = (int)(o % (unsigned)p.H);
        const unsigned rest = o / (unsigned)p.H;
        const int ix = (int)(rest % (unsigned)p.W);
        const int a = (int)(rest / (unsigned)p.W);
        const double *gt = p.gt + ((size_t)b * p.Gmax + g) * 4;
        AnchorPx an = anchor_px(p.stride, ix, jy, p.anchors.wh[a][0], p.anchors.wh[a][1]);
        double t[4];
        regr_targets(an, gt[0], gt[1], gt[2], gt[3], t);
        const int cell = jy * p.W + ix;
        cls_b[(size_t)a * HW + cell] = 1.0;
        cls_b[(size_t)(p.A + a) * HW + cell] = 1.0;
        for (int k = 0; k < 4; ++k) {
            regr_b[(size_t)(4 * a + k) * HW + cell] = 1.0;
            regr_b[(size_t)(4 * p.A + 4 * a + k) * HW + cell] = (double)(float)t[k];   // float32 store utils.py:605
        }
    }
}

// ----------------------------------------------------------------------------------
// a4: calc_iou per-RoI loop.  Single CTA, order-preserving compaction by block scan.
// ----------------------------------------------------------------------------------
struct RoiTargetParams {
    const int32_t *rois; int R;
    const double *gt; const int32_t *gt_class; int G;
    int n_cls, bg_class;
    double min_overlap, max_overlap;
    double std4[4];
    int32_t *x_roi; int32_t *y_class; double *y_regr; double *ious; int32_t *count;
};

__global__ void __launch_bounds__(1024) roi_targets_kernel(RoiTargetParams p) {
    __shared__ int s_warp[33];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int n_regr = 4 * (p.n_cls - 1);
    for (int r0 = 0; r0 < p.R; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        bool keep = false;
        double best = 0.0;
        int best_g = -1;
        int x1 = 0, y1 = 0, x2 = 0, y2 = 0;
        if (r < p.R) {
            int4 bx = reinterpret_cast<const int4 *>(p.rois)[r];
            x1 = bx.x; y1 = bx.y; x2 = bx.z; y2 = bx.w;
            for (int g = 0; g < p.G; ++g) {                                        // rpn.py:220-226
                const double *q = p.gt + 4 * g;
                double cur = ref_iou(q[0], q[2], q[1], q[3], (double)x1, (double)y1, (double)x2, (double)y2);
                if (cur > best) { best = cur; best_g = g; }
            }
            keep = !(best < p.min_overlap);                                        // rpn.py:228
        }
        // order-preserving slot
        unsigned km = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[w] = __popc(km);
        __syncthreads();
        if (w == 0) {
            int v = s_warp[lane], inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int n = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += n;
            }
            s_warp[lane] = inc - v;
            if (lane == 31) s_warp[32] = inc;
        }
        __syncthreads();
        const int base = s_base;
        if (keep) {
            const int slot = base + s_warp[w] + __popc(km & lanemask_lt());
            const int wd = x2 - x1, ht = y2 - y1;
            reinterpret_cast<int4 *>(p.x_roi)[slot] = make_int4(x1, y1, wd, ht);   // rpn.py:234-236
            p.ious[slot] = best;
            int cls = p.bg_class;
            double t[4] = {0, 0, 0, 0};
            if (best >= p.max_overlap) {                                           // rpn.py:244-256
                cls = p.gt_class[best_g];
                const double *q = p.gt + 4 * best_g;
                double cxg = __ddiv_rn(__dadd_rn(q[0], q[1]), 2.0), cyg = __ddiv_rn(__dadd_rn(q[2], q[3]), 2.0);
                double cx = __dadd_rn((double)x1, __ddiv_rn((double)wd, 2.0));
                double cy = __dadd_rn((double)y1, __ddiv_rn((double)ht, 2.0));
                t[0] = __ddiv_rn(__dsub_rn(cxg, cx), (double)wd);
                t[1] = __ddiv_rn(__dsub_rn(cyg, cy), (double)ht);
                t[2] = log(__ddiv_rn(__dsub_rn(q[1], q[0]), (double)wd));
                t[3] = log(__ddiv_rn(__dsub_rn(q[3], q[2]), (double)ht));
            }
            int32_t *yc = p.y_class + (size_t)slot * p.n_cls;
            for (int c = 0; c < p.n_cls; ++c) yc[c] = (c == cls) ? 1 : 0;          // rpn.py:263-266
            double *yr = p.y_regr + (size_t)slot * 2 * n_regr;
            for (int c = 0; c < 2 * n_regr; ++c) yr[c] = 0.0;
            if (cls != p.bg_class) {                                               // rpn.py:270-277
                for (int k = 0; k < 4; ++k) {
                    yr[4 * cls + k] = 1.0;
                    yr[n_regr + 4 * cls + k] = __dmul_rn(p.std4[k], t[k]);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + s_warp[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *p.count = s_base;
}

}  // namespace radnet

using namespace radnet;

extern "C" size_t radnet_rpn_targets_workspace_bytes(int B, int Gmax) {
    if (B < 1 || Gmax < 0) return 0;
    return align_up((size_t)B * (Gmax > 0 ? Gmax : 1) * sizeof(unsigned long long), 256);
}

extern "C" int radnet_rpn_targets(const double *gt, const uint8_t *gt_is_bg, const int32_t *gt_count, int B,
                                  int Gmax, int H, int W, int A, int n_ratios, const double *h_anchor_px,
                                  double rpn_stride, const double *img_wh, double max_overlap,
                                  double *y_rpn_cls, double *y_rpn_regr, int32_t *best_anchor, int32_t *n_hits,
                                  void *ws, size_t ws_bytes, void *stream) {
    RADNET_CHECK_ARG(gt_count && h_anchor_px && img_wh && y_rpn_cls && y_rpn_regr && ws, "rpn_targets: null pointer");
    RADNET_CHECK_ARG(Gmax == 0 || (gt && gt_is_bg && best_anchor && n_hits), "rpn_targets: null GT buffers");
    RADNET_CHECK_ARG(B >= 1 && B <= 65535 && H >= 1 && W >= 1 && A >= 1 && A <= kMaxAnchors && n_ratios >= 1 && Gmax >= 0,
                     "rpn_targets: bad sizes B=%d H=%d W=%d A=%d Gmax=%d", B, H, W, A, Gmax);
    RADNET_CHECK_ARG((long long)A * H * W < 0x7fffffffLL, "rpn_targets: anchor count overflows the loop-order key");
    size_t need = radnet_rpn_targets_workspace_bytes(B, Gmax);
    if (ws_bytes < need) {
        set_error("rpn_targets: workspace %zu < %zu", ws_bytes, need);
        return RADNET_E_WORKSPACE;
    }
    size_t gt_bytes = align_up((size_t)Gmax * (4 * 8 + 16 + 16 + 8 + 4 + 4 + 4 + 1) + 16, 16);
    int hit_cap = H * W < 4096 ? H * W : 4096;
    if (const char *e = getenv("RADNET_TARGETS_HIT_CAP")) {      // tests shrink the list to exercise the replay path
        int v = atoi(e);
        if (v >= 1 && v < hit_cap) hit_cap = v;
    }
    size_t smem = gt_bytes + (size_t)H * W * (sizeof(double) + sizeof(short)) + (size_t)(H + W) * 32 +
                  (size_t)hit_cap * sizeof(TargetHit) + 64;
    int dev = 0, smem_limit = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    RADNET_CUDA(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)smem_limit) {
        set_error("rpn_targets: %d figures on a %dx%d map need %zu B of shared memory (limit %d)", Gmax, H, W, smem,
                  smem_limit);
        return RADNET_E_UNSUPPORTED;
    }
    RADNET_CHECK_ARG(Gmax < 32768, "rpn_targets: Gmax=%d too large", Gmax);
    RpnTargetParams p{};
    p.gt = gt; p.gt_is_bg = gt_is_bg; p.gt_count = gt_count;
    p.Gmax = Gmax; p.H = H; p.W = W; p.A = A; p.n_ratios = n_ratios;
    for (int a = 0; a < A; ++a) {
        p.anchors.wh[a][0] = h_anchor_px[2 * a];
        p.anchors.wh[a][1] = h_anchor_px[2 * a + 1];
    }
    p.stride = rpn_stride; p.img_wh = img_wh; p.max_overlap = max_overlap;
    p.y_cls = y_rpn_cls; p.y_regr = y_rpn_regr; p.best_anchor = best_anchor; p.n_hits = n_hits;
    p.best_key = reinterpret_cast<unsigned long long *>(ws);
    p.sm_off_cells = (int)gt_bytes;
    p.hit_cap = hit_cap;
    cudaStream_t st = (cudaStream_t)stream;
    RADNET_CUDA(cudaMemsetAsync(ws, 0, need, st));
    if (Gmax > 0) RADNET_CUDA(cudaMemsetAsync(n_hits, 0, sizeof(int32_t) * (size_t)B * Gmax, st));
    if (smem > 48 * 1024)
        RADNET_CUDA(cudaFuncSetAttribute(rpn_targets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(A, B);
    rpn_targets_kernel<<<grid, kTgtThreads, smem, st>>>(p);
    int rc = check_launch("rpn_targets_kernel");
    if (rc) return rc;
    if (Gmax > 0) {
        rpn_targets_finalize_kernel<<<B, 128, (size_t)Gmax * sizeof(unsigned), st>>>(p);
        rc = check_launch("rpn_targets_finalize_kernel");
    }
    return rc;
}

extern "C" int radnet_roi_targets(const int32_t *rois, int R, const double *gt, const int32_t *gt_class, int G,
                                  int n_cls, int bg_class, double min_overlap, double max_overlap,
                                  const double *h_regr_std4, int32_t *x_roi, int32_t *y_class, double *y_regr,
                                  double *ious, int32_t *count, void *stream) {
    RADNET_CHECK_ARG(rois && h_regr_std4 && x_roi && y_class && y_regr && ious && count, "roi_targets: null pointer");
    RADNET_CHECK_ARG(R >= 1 && G >= 0 && n_cls >= 2 && bg_class >= 0 && bg_class < n_cls, "roi_targets: bad sizes");
    RADNET_CHECK_ARG(G == 0 || (gt && gt_class), "roi_targets: null GT buffers");
    RoiTargetParams p{};
    p.rois = rois; p.R = R; p.gt = gt; p.gt_class = gt_class; p.G = G;
    p.n_cls = n_cls; p.bg_class = bg_class; p.min_overlap = min_overlap; p.max_overlap = max_overlap;
    for (int k = 0; k < 4; ++k) p.std4[k] = h_regr_std4[k];
    p.x_roi = x_roi; p.y_class = y_class; p.y_regr = y_regr; p.ious = ious; p.count = count;
    roi_targets_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p);
    return check_launch("roi_targets_kernel");
}

// utils.iou(a, b) for n box pairs (reference utils.py:77-109): API parity, not on the batched hot path
namespace radnet {
__global__ void iou_pairs_kernel(const double *__restrict__ a, const double *__restrict__ b, long long n,
                                 double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *p = a + 4 * i, *q = b + 4 * i;
    out[i] = ref_iou(p[0], p[1], p[2], p[3], q[0], q[1], q[2], q[3]);
}
}  // namespace radnet

extern "C" int radnet_iou_pairs(const double *a, const double *b, long long n, double *out, void *stream) {
    RADNET_CHECK_ARG(a && b && out && n >= 0, "iou_pairs: bad arguments");
    if (n == 0) return RADNET_OK;
    const long long blocks = (n + 255) / 256;
    RADNET_CHECK_ARG(blocks <= 0x7fffffffLL, "iou_pairs: n too large");
    radnet::iou_pairs_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
    return radnet::check_launch("iou_pairs_kernel");
}


// iou.cuh - the reference's scalar IoU and regression-target formulas (faster_rcnn/utils.py:77-109, 669-687),
// shared by K3 (rpn_targets.cu) and a4 (targets.cu).  Every operation rounds once, in the reference's order.
#pragma once
#include "common.cuh"

namespace radnet {

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

}  // namespace radnet

// targets.cu - a4: RoI target assignment (reference faster_rcnn/rpn.py:209-282), single image and batched,
// and utils.iou for box pairs.  K3 (RPN anchor targets) lives in rpn_targets.cu.
#include "iou.cuh"

namespace radnet {

// ----------------------------------------------------------------------------------
// a4: calc_iou per-RoI loop, one CTA per panel, order-preserving compaction by block scan.
// RoIs come either from the detection records of K2 (det != NULL) or from a dense array.
// ----------------------------------------------------------------------------------
struct RoiTargetParams {
    const int32_t *rois;       // [B][R][4] x1,y1,x2,y2 (det == NULL)
    const uint8_t *det;        // [B] detection records (or NULL)
    size_t det_stride;
    const int32_t *roi_count;  // [B] or NULL (= R)
    int R;
    const double *gt;          // [B][Gmax][4] x1,x2,y1,y2 feature cells
    const int32_t *gt_class;   // [B][Gmax]
    const int32_t *gt_count;   // [B] or NULL (= Gmax)
    int Gmax;
    int n_cls, bg_class;
    double min_overlap, max_overlap;
    double std4[4];
    int32_t *x_roi;            // [B][R][4]
    int32_t *y_class;          // [B][R][n_cls]
    double *y_regr;            // [B][R][8(n_cls-1)]
    double *ious;              // [B][R]
    int32_t *best_gt;          // [B][R] or NULL
    int32_t *count;            // [B]
};

__global__ void __launch_bounds__(1024) roi_targets_kernel(RoiTargetParams p) {
    __shared__ int s_warp[33];
    __shared__ int s_base;
    extern __shared__ __align__(16) unsigned char smem[];
    double *s_gt = reinterpret_cast<double *>(smem);                                   // [G][4]
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_regr = 4 * (p.n_cls - 1);
    int R = p.R;
    const int32_t *rois = nullptr;
    if (p.det) {
        const int32_t *rec = reinterpret_cast<const int32_t *>(p.det + (size_t)b * p.det_stride);
        R = min(max(rec[0], 0), p.R);
        rois = rec + 4;
    } else {
        rois = p.rois + (size_t)b * p.R * 4;
        if (p.roi_count) R = min(max(p.roi_count[b], 0), p.R);
    }
    const int G = p.gt_count ? min(max(p.gt_count[b], 0), p.Gmax) : p.Gmax;
    const double *gt = p.gt + (size_t)b * p.Gmax * 4;
    const int32_t *gcls = p.gt_class + (size_t)b * p.Gmax;
    int32_t *x_roi = p.x_roi + (size_t)b * p.R * 4;
    int32_t *y_class = p.y_class + (size_t)b * p.R * p.n_cls;
    double *y_regr = p.y_regr + (size_t)b * p.R * 2 * n_regr;
    double *ious = p.ious + (size_t)b * p.R;
    int32_t *best_out = p.best_gt ? p.best_gt + (size_t)b * p.R : nullptr;
    for (int i = threadIdx.x; i < 4 * G; i += blockDim.x) s_gt[i] = gt[i];
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int r0 = 0; r0 < R; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        bool keep = false;
        double best = 0.0;
        int best_g = -1;
        int x1 = 0, y1 = 0, x2 = 0, y2 = 0;
        if (r < R) {
            int4 bx = reinterpret_cast<const int4 *>(rois)[r];
            x1 = bx.x; y1 = bx.y; x2 = bx.z; y2 = bx.w;
            for (int g = 0; g < G; ++g) {                                          // rpn.py:220-226
                const double *q = s_gt + 4 * g;
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
            reinterpret_cast<int4 *>(x_roi)[slot] = make_int4(x1, y1, wd, ht);     // rpn.py:234-236
            ious[slot] = best;
            int cls = p.bg_class;
            int bg_out = -1;
            double t[4] = {0, 0, 0, 0};
            // a positive RoI has w,h >= 1: a zero-extent RoI has IoU 0.0 (utils.py:99-100) and never gets here
            if (best >= p.max_overlap) {                                           // rpn.py:244-256
                cls = gcls[best_g];
                bg_out = best_g;
                const double *q = s_gt + 4 * best_g;
                double cxg = __ddiv_rn(__dadd_rn(q[0], q[1]), 2.0), cyg = __ddiv_rn(__dadd_rn(q[2], q[3]), 2.0);
                double cx = __dadd_rn((double)x1, __ddiv_rn((double)wd, 2.0));
                double cy = __dadd_rn((double)y1, __ddiv_rn((double)ht, 2.0));
                t[0] = __ddiv_rn(__dsub_rn(cxg, cx), (double)wd);
                t[1] = __ddiv_rn(__dsub_rn(cyg, cy), (double)ht);
                t[2] = log(__ddiv_rn(__dsub_rn(q[1], q[0]), (double)wd));
                t[3] = log(__ddiv_rn(__dsub_rn(q[3], q[2]), (double)ht));
            }
            if (best_out) best_out[slot] = bg_out;
            int32_t *yc = y_class + (size_t)slot * p.n_cls;
            for (int c = 0; c < p.n_cls; ++c) yc[c] = (c == cls) ? 1 : 0;          // rpn.py:263-266
            double *yr = y_regr + (size_t)slot * 2 * n_regr;
            for (int c = 0; c < 2 * n_regr; ++c) yr[c] = 0.0;
            // cls < 0: the figure's class is not in class_mapping - the host raises KeyError (rpn.py:263)
            if (cls != p.bg_class && cls >= 0 && cls < p.n_cls - 1) {              // rpn.py:270-277
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
    if (threadIdx.x == 0) p.count[b] = s_base;
}

}  // namespace radnet

using namespace radnet;

static int roi_targets_launch(RoiTargetParams &p, int B, const double *h_regr_std4, void *stream) {
    for (int k = 0; k < 4; ++k) p.std4[k] = h_regr_std4[k];
    const size_t smem = (size_t)(p.Gmax > 0 ? p.Gmax : 1) * 4 * sizeof(double);
    RADNET_CHECK_ARG(smem <= 40 * 1024, "roi_targets: too many figures (%d)", p.Gmax);
    roi_targets_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>(p);
    return check_launch("roi_targets_kernel");
}

extern "C" int radnet_roi_targets(const int32_t *rois, int R, const double *gt, const int32_t *gt_class, int G,
                                  int n_cls, int bg_class, double min_overlap, double max_overlap,
                                  const double *h_regr_std4, int32_t *x_roi, int32_t *y_class, double *y_regr,
                                  double *ious, int32_t *best_gt, int32_t *count, void *stream) {
    RADNET_CHECK_ARG(rois && h_regr_std4 && x_roi && y_class && y_regr && ious && count, "roi_targets: null pointer");
    RADNET_CHECK_ARG(R >= 1 && G >= 0 && n_cls >= 2 && bg_class == n_cls - 1, "roi_targets: bad sizes");
    RADNET_CHECK_ARG(G == 0 || (gt && gt_class), "roi_targets: null GT buffers");
    RoiTargetParams p{};
    p.rois = rois; p.R = R; p.gt = gt; p.gt_class = gt_class; p.Gmax = G;
    p.n_cls = n_cls; p.bg_class = bg_class; p.min_overlap = min_overlap; p.max_overlap = max_overlap;
    p.x_roi = x_roi; p.y_class = y_class; p.y_regr = y_regr; p.ious = ious; p.best_gt = best_gt; p.count = count;
    return roi_targets_launch(p, 1, h_regr_std4, stream);
}

extern "C" int radnet_roi_targets_batch(const void *det, int det_max_boxes, const int32_t *rois,
                                        const int32_t *roi_count, int B, int R, const double *gt,
                                        const int32_t *gt_class, const int32_t *gt_count, int Gmax, int n_cls,
                                        int bg_class, double min_overlap, double max_overlap,
                                        const double *h_regr_std4, int32_t *x_roi, int32_t *y_class, double *y_regr,
                                        double *ious, int32_t *best_gt, int32_t *count, void *stream) {
    RADNET_CHECK_ARG((det || rois) && h_regr_std4 && x_roi && y_class && y_regr && ious && count,
                     "roi_targets_batch: null pointer");
    RADNET_CHECK_ARG(B >= 1 && R >= 1 && Gmax >= 0 && n_cls >= 2 && bg_class == n_cls - 1, "roi_targets_batch: bad sizes");
    RADNET_CHECK_ARG(Gmax == 0 || (gt && gt_class), "roi_targets_batch: null GT buffers");
    RADNET_CHECK_ARG(!det || (det_max_boxes >= 1 && R <= det_max_boxes), "roi_targets_batch: R exceeds the record capacity");
    RoiTargetParams p{};
    p.det = reinterpret_cast<const uint8_t *>(det);
    p.det_stride = det ? radnet_det_record_bytes(det_max_boxes) : 0;
    p.rois = rois; p.roi_count = roi_count; p.R = R;
    p.gt = gt; p.gt_class = gt_class; p.gt_count = gt_count; p.Gmax = Gmax;
    p.n_cls = n_cls; p.bg_class = bg_class; p.min_overlap = min_overlap; p.max_overlap = max_overlap;
    p.x_roi = x_roi; p.y_class = y_class; p.y_regr = y_regr; p.ious = ious; p.best_gt = best_gt; p.count = count;
    return roi_targets_launch(p, B, h_regr_std4, stream);
}

// utils.iou(a, b) for n box pairs (reference utils.py:77-109)
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

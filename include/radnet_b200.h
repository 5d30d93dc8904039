/*
 * radnet_b200.h - C ABI of libradnet_b200.so (hand-written sm_100a CUDA kernels).
 *
 * The reference (Swedish-Rock-Art-Research-Archives/rock-art-radnet) has no FFI:
 * its proposal / RoI hot path is plain Python + NumPy (+ one Keras layer).  This
 * header is therefore the boundary the *replacement* binds: the Python package
 * `rock_art_radnet_b200` keeps the reference's function names and array layouts
 * and calls these entry points through ctypes.  Each entry
 * point names the reference interface (file:line under the reference tree) whose
 * arithmetic it takes over.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no C++ or torch types.
 *   - every pointer is a DEVICE pointer unless the parameter name starts with
 *     `h_` (host).  Buffers are caller-allocated; sizes of scratch areas come
 *     from the matching `*_workspace_bytes` function.
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous with
 *     respect to the host.  The library keeps no device state of its own: every buffer,
 *     including scratch, belongs to the caller.  Host-side it keeps only per-device
 *     caches of device attributes (atomics) and the tuning options below.
 *   - return value: 0 = RADNET_OK, negative = RADNET_E_*.  Nothing throws.
 *     `radnet_last_error_string()` returns a thread-local description.
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with RADNET_E_CUDA.
 */
#ifndef RADNET_B200_H
#define RADNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RADNET_ABI_VERSION 2

enum {
    RADNET_OK = 0,
    RADNET_E_INVALID = -1,   /* bad argument (null pointer, size out of range)   */
    RADNET_E_CUDA = -2,      /* CUDA runtime error, see radnet_last_error_string */
    RADNET_E_WORKSPACE = -3, /* workspace too small                              */
    RADNET_E_UNSUPPORTED = -4
};

/* Per-panel detection record written by radnet_sort_nms_i32 and exchanged between
 * ranks (one NCCL gather per batch).  A record is
 *   int32 header[4] = {count (-1 = internal fault), n_candidates, n_score_ties among the
 *                      n_sorted top-scoring candidates the kernel had to sort, n_sorted}
 *   int32 boxes [max_boxes][4]   x1,y1,x2,y2 in feature cells, score-descending
 *   float scores[max_boxes]
 *   int32 index [max_boxes]      flat anchor index a*H*W + r*W + c of each kept box
 * padded to a multiple of 16 bytes; radnet_det_record_bytes gives the stride.
 * Rows >= count are zero. */
size_t radnet_det_record_bytes(int max_boxes);

int radnet_version(void);
const char *radnet_last_error_string(void);
const char *radnet_error_name(int code);
/* Device facts used by the host to size launches: out[0]=SM count, out[1]=max opt-in
 * shared memory per block, out[2]=compute capability major*10+minor. */
int radnet_device_info(int *h_out3);

/* Process-wide tuning options, read (atomically) by every launch.  Each is seeded once, when the
 * library is loaded, from the environment variable RADNET_<NAME IN CAPITALS>.  Every value gives the
 * same results; the options choose between equivalent code paths (tests use them to exercise all of
 * them):  nms_cluster (-1 auto, 0 one CTA per panel), nms_cluster_size (0 auto | 2 | 4 | 8 | 16),
 * nms_cluster_ranks, nms_sel_target, nms_lookahead (0 = default), roipool_force_direct (0 | 1),
 * roipool_form (0 auto | 1 whole-map slices | 2 row bands), roipool_bands / roipool_lanes (band form: bands per map,
 * float4 lanes per pixel, 8 | 16 | 32; whole-map form: 4 | 2 | 1 caps the slice width; 0 = auto), roipool_cluster (-1 = lockstep tuned by timing at the first large call per shape |
 * 0 none | 2 | 4 | 8 CTAs per cluster) with roipool_sync_every (barrier every n column rounds), roipool_ctas (0 = one
 * CTA per work item, else that many persistent CTAs), targets_hit_cap (0 = default),
 * targets_compute_ctas (0 = 43 % of the SMs), targets_two_launches (0 | 1: fill and panels as two
 * launches, no co-residency assumed), targets_fill_bulk (0 plain stores | 1 regression zeros by TMA bulk copies
 * shared out over all CTAs | n = size of the zero buffer), sampler_force_exact (0 | 1). */
int radnet_set_option(const char *h_name, long long value);
int radnet_get_option(const char *h_name, long long *h_value);

/* ---------------------------------------------------------------- K1: decode + clip
 * Replaces the anchor loop of rpn_to_roi (reference faster_rcnn/rpn.py:91-166) and
 * apply_regr_np (rpn.py:299-344) for a batch of B panels.
 *   cls   [B][H][W][A]    float32 objectness           (rpn.py:75-77)
 *   regr  [B][H][W][4A]   float32 (tx,ty,tw,th) per anchor (rpn.py:78-80)
 *   h_anchor_wh [A][2]    float64 anchor (w,h) in feature cells, i.e.
 *                         scale*ratio/rpn_stride evaluated by the caller (rpn.py:112-113)
 *   std_scaling           divisor applied in float32 (rpn.py:91)
 *   use_regr              0 = anchors only (rpn.py:133)
 * Outputs, anchor-major flat order i = a*H*W + r*W + c (rpn.py:154-155):
 *   boxes_i32 [B][N][4]   x1,y1,x2,y2 after round/min-size/clip, as int32 (use_regr=1:
 *                         always integer valued; use_regr=0: only valid when every
 *                         anchor half-size is an integer, else use the f64 variant)
 *   keys      [B][N]      order-preserving uint32 image of the score; 0 marks a box the
 *                         reference deletes as degenerate (rpn.py:163-166)
 *   stats     [B][4]      int32 {n_valid, n_nonfinite, n_round_near_ties, n_noninteger}
 */
int radnet_decode_clip_i32(const float *cls, const float *regr, int B, int H, int W, int A,
                           const double *h_anchor_wh, float std_scaling, int use_regr,
                           int32_t *boxes_i32, uint32_t *keys, int32_t *stats, void *stream);

/* Same arithmetic, float64 boxes + float32 scores + validity bytes (general path and the
 * parity surface for decoded boxes).  boxes_f64 [B][N][4], scores [B][N], valid [B][N]. */
int radnet_decode_clip_f64(const float *cls, const float *regr, int B, int H, int W, int A,
                           const double *h_anchor_wh, float std_scaling, int use_regr,
                           double *boxes_f64, float *scores, uint8_t *valid, int32_t *stats,
                           void *stream);

/* apply_regr_np(X, T) (rpn.py:299-344): X [4][n] float64 (x,y,w,h planes), T [4][n]
 * float64 (tx,ty,tw,th planes; the float32 regression map widened exactly by the caller)
 * -> out [4][n] float64 (x1,y1,w1,h1 rounded half-even). */
int radnet_apply_regr(const double *X, const double *T, long long n, double *out, void *stream);

/* ------------------------------------------------- K2: segmented score sort + greedy NMS
 * Replaces non_max_suppression_fast on the rpn_to_roi path (rpn.py:380-456, called at
 * rpn.py:170) for B independent panels (segments) of N candidates each.  One CTA per
 * panel: radix sort of the keys (ties: higher flat index first), then exact greedy
 * suppression `inter/(union+1e-6) > thr` (rpn.py:443-447) with a stop at max_boxes
 * (rpn.py:449-450).  Integer boxes make inter/union exact; the float64 predicate is
 * tabulated per union value at kernel start, so the decision is bit-exact.
 *   boxes_i32 [B][N][4], keys [B][N] as produced by radnet_decode_clip_i32
 *   map_h,map_w  feature-map size (bounds the union table: 2*(H-1)*(W-1) entries)
 *   det        B records (layout above), record stride radnet_det_record_bytes(max_boxes)
 *   ws         scratch of radnet_sort_nms_i32_workspace_bytes(B,N,map_h,map_w,max_boxes)
 */
size_t radnet_sort_nms_i32_workspace_bytes(int B, int N, int map_h, int map_w, int max_boxes);
int radnet_sort_nms_i32(const int32_t *boxes_i32, const uint32_t *keys, int B, int N,
                        int map_h, int map_w, double thr, int max_boxes, void *det,
                        void *ws, size_t ws_bytes, void *stream);

/* General non_max_suppression_fast(boxes, probs, overlap_thresh, max_boxes)
 * (rpn.py:380-456; other call sites RADNet.py:574,639,698): float64 boxes of any
 * magnitude, float64 scores, float64 IoU arithmetic in the reference's association
 * order.  One segment of M boxes.
 *   boxes [M][4] float64, probs [M] float64
 *   valid [M] uint8 or NULL: rows with valid==0 are skipped (the rows rpn_to_roi deletes
 *         as degenerate, rpn.py:163-166); NULL = all rows are candidates
 *   pick  [min(max_boxes,M)] int32 picked row indices, score-descending
 *   count [3] int32 {n_picked (-1 = internal fault), n_score_ties among the n_sorted
 *         top-scoring candidates the kernel had to sort, n_sorted}
 */
size_t radnet_nms_f64_workspace_bytes(int M, int max_boxes);
int radnet_nms_f64(const double *boxes, const double *probs, const uint8_t *valid, int M, double thr,
                   int max_boxes, int32_t *pick, int32_t *count, void *ws, size_t ws_bytes, void *stream);

/* ----------------------------------------------------------------- K4: RoI pooling
 * Replaces RoiPoolingConv.call (reference faster_rcnn/RoiPoolingConv.py:48-88): crop
 * + TF-1 legacy bilinear resize (tf.image.resize_images, align_corners=False) of every
 * RoI to pool x pool, float32, no fused multiply-add.
 *   feat   [B][H][W][C]  float32 NHWC feature maps
 *   rois   xywh int32, feature cells.  Two addressing modes:
 *          det != NULL : RoIs are the kept boxes of the B detection records (xyxy ->
 *                        x,y,w=x2-x1,h=y2-y1, RADNet.py:564-565), rois_per_panel slots each;
 *          det == NULL : rois [B][rois_per_panel][4] int32 (x,y,w,h), roi_count [B] (may be NULL
 *                        = all slots used).
 *   out    [B][rois_per_panel][pool][pool][C] float32; slots >= count are zero-filled.
 * RoIs must lie inside the map after TF's slice clamping (x,y >= 0); the host shim checks.
 */
int radnet_roi_pool(const float *feat, int B, int H, int W, int C, const void *det,
                    int det_max_boxes, const int32_t *rois, const int32_t *roi_count,
                    int rois_per_panel, int pool, float *out, void *stream);
/* The launch form radnet_roi_pool settled on for this shape on the current device (a large call times a few
 * equivalent forms once - see option roipool_cluster): h_out3 = {float4 lanes per pixel of a slice, CTAs per
 * cluster, barrier every n column rounds}, or {-1, -1, -1} when no call of this shape has been tuned yet.  The form
 * is per (device, H, W, C, pool, rois_per_panel); B is accepted for symmetry with radnet_roi_pool and ignored. */
int radnet_roi_pool_form(int B, int H, int W, int C, int pool, int rois_per_panel, int *h_out3);

/* --------------------------------------------------- K3: RPN anchor target assignment
 * Replaces the deterministic part of calc_region_props (reference faster_rcnn/utils.py:
 * 585-775 and 815-816; upstream name calc_rpn) for B panels, in ONE launch.  The RNG-driven
 * 256-region subsampling (utils.py:777-813) is radnet_rpn_subsample below.
 *   gt        [B][Gmax][4] float64 x1,x2,y1,y2 in resized-image pixels (utils.py:608-613)
 *   gt_is_bg  [B][Gmax]    uint8, 1 = class 'bg' (utils.py:690)
 *   gt_count  [B]          int32
 *   h_anchor_px [A][2]     float64 anchor (w,h) in pixels, a = ratio_idx + n_ratios*size_idx
 *   img_wh    [B][2]       float64 resized width,height (utils.py:629,638)
 *   n_ratios               len(anchor_box_ratios), to split a into (ratio_idx,size_idx)
 *   layout                 RADNET_TARGETS_CHANNEL_FIRST: as returned by the reference
 *                            y_rpn_cls  [B][2A][H][W] float64 = [valid | overlap]          (utils.py:815)
 *                            y_rpn_regr [B][8A][H][W] float64 = [repeat(overlap,4) | regr] (utils.py:816)
 *                          RADNET_TARGETS_NHWC: as the training loop consumes them (utils.py:477-478)
 *                            y_rpn_cls  [B][H][W][2A],  y_rpn_regr [B][H][W][8A]
 *   regr_scale             factor applied to the regr half (1.0, or C.std_scaling: utils.py:475)
 *   best_anchor [B][Gmax][4] int32 {jy, ix, ratio_idx, size_idx} or -1 (utils.py:697)
 *   n_hits      [B][Gmax]    int32 positives per GT before forcing    (utils.py:707)
 *               A panel whose fill never completed (co-residency lost for 4 s) has n_hits = -1.
 *   ws          radnet_rpn_targets_workspace_bytes(B,Gmax,H,W,A) bytes of counters that every successful
 *               launch leaves zeroed: call radnet_rpn_targets_workspace_init once after allocating it
 *               (and after a failed launch).  One workspace serves one launch at a time.
 * gt, best_anchor and both output tensors must be 16-byte aligned. */
enum { RADNET_TARGETS_CHANNEL_FIRST = 0, RADNET_TARGETS_NHWC = 1 };
size_t radnet_rpn_targets_workspace_bytes(int B, int Gmax, int H, int W, int A);
int radnet_rpn_targets_workspace_init(void *ws, size_t ws_bytes, int B, int Gmax, void *stream);
int radnet_rpn_targets(const double *gt, const uint8_t *gt_is_bg, const int32_t *gt_count,
                       int B, int Gmax, int H, int W, int A, int n_ratios,
                       const double *h_anchor_px, double rpn_stride, const double *img_wh,
                       double max_overlap, int layout, double regr_scale, double *y_rpn_cls,
                       double *y_rpn_regr, int32_t *best_anchor, int32_t *n_hits, void *ws,
                       size_t ws_bytes, void *stream);

/* ----------------------------------------------------------- a4: RoI target assignment
 * Replaces the per-RoI loop of calc_iou (reference faster_rcnn/rpn.py:209-282).
 *   rois   [R][4] int32 x1,y1,x2,y2;  gt [G][4] float64 x1,x2,y1,y2 in feature cells
 *   (already rounded by the host, rpn.py:197-200);  gt_class [G] int32 class index, -1 = the
 *   figure's class is not in class_mapping (only an error if it is some RoI's best match).
 *   bg_class  index of 'bg' = n_cls-1;  n_cls = len(class_mapping);  regr_std [4].
 * Outputs compacted in RoI order: x_roi [R][4] int32 (x,y,w,h), y_class [R][n_cls] int32,
 * y_regr [R][8*(n_cls-1)] float64 = [labels | coords], ious [R] float64,
 * best_gt [R] int32 (index of the matched figure of a positive row, -1 for 'bg' rows; may be NULL),
 * count [1] int32 = number of rows written.
 */
int radnet_roi_targets(const int32_t *rois, int R, const double *gt, const int32_t *gt_class,
                       int G, int n_cls, int bg_class, double min_overlap, double max_overlap,
                       const double *h_regr_std4, int32_t *x_roi, int32_t *y_class,
                       double *y_regr, double *ious, int32_t *best_gt, int32_t *count, void *stream);

/* The same for B panels in one launch (one CTA per panel) - the training path right after K2.
 * RoIs: det != NULL: the kept boxes of the B detection records (xyxy); else rois [B][R][4] int32
 * xyxy with roi_count [B] (NULL = R).  gt [B][Gmax][4], gt_class [B][Gmax], gt_count [B] (NULL = Gmax).
 * Outputs are [B][R][...] with rows >= count[b] left untouched. */
int radnet_roi_targets_batch(const void *det, int det_max_boxes, const int32_t *rois,
                             const int32_t *roi_count, int B, int R, const double *gt,
                             const int32_t *gt_class, const int32_t *gt_count, int Gmax, int n_cls,
                             int bg_class, double min_overlap, double max_overlap,
                             const double *h_regr_std4, int32_t *x_roi, int32_t *y_class,
                             double *y_regr, double *ious, int32_t *best_gt, int32_t *count,
                             void *stream);

/* ------------------------------------------------ f3: training-side sampling
 * The reference samples with NumPy's legacy global generator (np.random.seed, train.py:134-135).  A
 * generator state here is uint32[625] = key[624] + pos, the layout of np.random.get_state()[1:3]; every
 * entry point reads the state of panel b at rng_states + 625*b, consumes exactly the words NumPy would
 * and writes the advanced state back, so a host can pass np.random's own state in and out.
 *
 * radnet_mt19937_seed: np.random.seed(seeds[b]) for 32-bit integer seeds (init_genrand). */
int radnet_mt19937_seed(const uint32_t *seeds, int B, uint32_t *rng_states, void *stream);

/* The 256-region balancing at the end of calc_region_props (reference faster_rcnn/utils.py:777-813), in
 * place on the label tensor written by radnet_rpn_targets (same layout argument): if more than
 * max_regions/2 anchors are positive, np.random.choice(n_pos, n_pos - max_regions/2, replace=False, p)
 * of them are switched off (valid = 0); then, if positives + negatives exceed max_regions,
 * np.random.choice(n_neg, n_neg - n_pos, replace=False, p) negatives.  p is the reference's per-channel
 * weight, looked up in a table keyed by the channels of the NEGATIVES in both branches (utils.py:789,804).
 *   out [B][8] int32 {n_pos returned by calc_region_props, positives found, negatives found,
 *                     status (0 ok, 1 = the reference raises KeyError: a positive's channel has no
 *                     negative; nothing is drawn or changed), rounds redone with the serial cumsum,
 *                     draws, 0, 0}
 *   ws  radnet_rpn_subsample_workspace_bytes(B,H,W,A) bytes. */
size_t radnet_rpn_subsample_workspace_bytes(int B, int H, int W, int A);
int radnet_rpn_subsample(double *y_rpn_cls, int B, int H, int W, int A, int layout, int max_regions,
                         uint32_t *rng_states, int32_t *out, void *ws, size_t ws_bytes, void *stream);

/* get_selected_samples (reference train.py:93-129) for B panels: from the one-hot rows of calc_iou
 * (y_class [B][R][n_cls] int32, count [B] rows used or NULL = R; 'bg' is the last class) pick n_rois rows,
 * positives first: all positives if fewer than n_rois/2, else np.random.choice(pos, n_rois/2,
 * replace=False); then np.random.choice(neg, rest, replace=False) (replace=True when there are too few);
 * without any negative the reference's second branch (train.py:122-127).
 *   sel [B][n_rois] int32 selected row indices;  out [B][4] int32 {rows selected, n_pos, n_neg, status
 *   (0 ok, 2 = the reference raises ValueError: nothing to sample from)}.  R <= 2048. */
int radnet_select_samples(const int32_t *y_class, const int32_t *count, int B, int R, int n_cls,
                          int n_rois, uint32_t *rng_states, int32_t *sel, int32_t *out, void *stream);

/* utils.iou(a, b) (reference faster_rcnn/utils.py:77-109) for n box pairs: a, b [n][4] float64
 * (x1,y1,x2,y2) -> out [n] float64; 0.0 for degenerate boxes, else inter/(union+1e-6). */
int radnet_iou_pairs(const double *a, const double *b, long long n, double *out, void *stream);

/* ------------------------------------------------ f1/f2: detection post-processing
 * Labelled detection record, the unit exchanged between these entry points and between ranks
 * (stride = radnet_cls_record_bytes(max_det)):
 *   int32 header[8] = {n_det (-1 = an input record carried a fault, -2 = capacity exceeded: more
 *                      entries than the kernel holds or than the output record has slots),
 *                      n_in (RoIs examined / entries concatenated), n_degenerate (boxes with
 *                      x1>=x2 or y1>=y2: the reference's NMS asserts, rpn.py:400-401),
 *                      n_score_ties, n_classes_present,
 *                      n_regr_fallback (classify_*) or n_empty_clusters (final_nms),
 *                      n_round_near_ties, n_out_of_range (|coordinate| > 2^25)}
 *   int32 class_order[32]  class ids in first-appearance order (the insertion order of the
 *                          reference's per-class dicts), -1 padded
 *   int32 class_count[32]  entries per class id
 *   entry[max_det]         {int32 cls; float prob; int32 x1,y1,x2,y2; int32 src; int32 aux}
 * Entries of radnet_classify_decode are in RoI order; those of the NMS entry points are grouped by
 * class (class_order) and, inside a class, in pick order (descending score). */
size_t radnet_cls_record_bytes(int max_det);

/* Per-RoI class decision and box decode of RADNet.apply_spatial_pyramid_pooling (reference
 * faster_rcnn/RADNet.py:123-150) with the scalar apply_regr (rpn.py:346-378), B tiles at once.
 *   p_cls  [B][R][n_cls] float32, p_regr [B][R][4*(n_cls-1)] float32: classifier-head outputs
 *   RoIs (x,y,w,h in feature cells), two addressing modes as in radnet_roi_pool:
 *     det != NULL : the kept boxes of the B detection records (x2-x1, y2-y1; RADNet.py:564-565)
 *     det == NULL : rois [B][R][4] int32, roi_count [B] or NULL (= R); the caller has already
 *                   padded the last chunk as RADNet.py:106-118 does
 *   bbox_threshold   compared in float32 (RADNet.py:126, NumPy >= 2 scalar rule)
 *   h_regr_std4      classifier_regr_std; the division is float32 (RADNet.py:140-143)
 *   rpn_stride       integer stride (RADNet.py:149); R <= 1024, n_cls <= 32 ('bg' is the last class)
 * Output: B records, entries in RoI order, src = RoI index. */
int radnet_classify_decode(const float *p_cls, const float *p_regr, int B, int R, int n_cls,
                           const void *det, int det_max_boxes, const int32_t *rois,
                           const int32_t *roi_count, double bbox_threshold,
                           const double *h_regr_std4, int rpn_stride, void *rec_out,
                           int rec_max_det, void *stream);

/* The same decode fused with the per-class NMS of RADNet.predict (rpn.non_max_suppression_fast at
 * RADNet.py:574, threshold nms_thr, stop at max_boxes per class), get_real_coordinates
 * (RADNet.py:44-51: Python floor division by ratio[b]) and the tile offset origin[b] = {x0,y0}
 * (RADNet.py:582-600).  ratio / origin may be NULL (no scaling / no offset). */
int radnet_classify_nms(const float *p_cls, const float *p_regr, int B, int R, int n_cls,
                        const void *det, int det_max_boxes, const int32_t *rois,
                        const int32_t *roi_count, double bbox_threshold,
                        const double *h_regr_std4, int rpn_stride, double nms_thr, int max_boxes,
                        const double *ratio, const int32_t *origin, void *rec_out,
                        int rec_max_det, void *stream);

/* Per-class greedy NMS over the concatenation of n_in records per segment (the cross-image NMS of
 * RADNet.py:695-716, or the NMS half of radnet_classify_nms on records).  rec_in is
 * [S][n_in] records of capacity in_max_det; in_count [S] (NULL = n_in) limits the records used.
 * Optional ratio [S] / origin [S][2] as above. */
size_t radnet_class_nms_workspace_bytes(int S, int n_in, int in_max_det, int n_cls);
int radnet_class_nms(const void *rec_in, int in_max_det, int S, int n_in, const int32_t *in_count,
                     int n_cls, double thr, int max_boxes, const double *ratio,
                     const int32_t *origin, void *rec_out, int out_max_det, void *ws,
                     size_t ws_bytes, void *stream);

/* RADNet.final_nms (RADNet.py:156-240): cluster-and-average merge, per class, of the n_in tile
 * records of each of S images.  A cluster = the best remaining box and every remaining box with
 * inter/(union+1e-6) > avg_thr; it is represented by the members scoring above conf_thr (float32
 * comparison) or, if even its best score is below conf_thr, by its n_obj_avg best members:
 * box = rint(mean), prob = float32 mean in NumPy's pairwise order.  Entry.src = input index of the
 * cluster's best box, entry.aux = number of members averaged.  At most 4096 boxes per (image,
 * class); beyond that header[0] = -2. */
size_t radnet_final_nms_workspace_bytes(int S, int n_in, int in_max_det, int n_cls);
int radnet_final_nms(const void *rec_in, int in_max_det, int S, int n_in, const int32_t *in_count,
                     int n_cls, double avg_thr, double conf_thr, int n_obj_avg, void *rec_out,
                     int out_max_det, void *ws, size_t ws_bytes, void *stream);

/* RADNet.get_real_coordinates (RADNet.py:44-51) for n coordinates: out[i] = int(round(v[i] // ratio)),
 * Python float floor division.  (radnet_classify_nms / radnet_class_nms apply the same rule fused.) */
int radnet_real_coordinates(const int32_t *v, long long n, double ratio, int32_t *out, void *stream);


/* ------------------------------------------------ f4: losses and evaluation
 * The four training losses of the reference (faster_rcnn/losses.py:16-95) as fused masked reductions over the
 * NHWC target tensors radnet_rpn_targets writes (layout RADNET_TARGETS_NHWC, regr half scaled) and the rows
 * radnet_roi_targets_batch writes.  One value per panel (the reference trains with batch size 1).  Elements are
 * evaluated in float32 exactly as the Keras-2.2 / TF-1 backend calls of losses.py define them (including
 * K.binary_crossentropy(y_pred, y_true[..., A:]) with the prediction in the `target` slot, losses.py:65); sums are
 * float64 in a fixed order (deterministic).  TF's reduction order is unspecified: parity bar 1e-5 relative.
 *   loss [B][2] float32: radnet_rpn_losses {rpn_loss_cls, rpn_loss_regr}; radnet_class_losses {class_loss_cls,
 *   class_loss_regr}.  ws of radnet_rpn_losses: radnet_rpn_losses_workspace_bytes(B), zeroed once with
 *   radnet_rpn_losses_workspace_init and left zeroed by every launch. */
size_t radnet_rpn_losses_workspace_bytes(int B);
int radnet_rpn_losses_workspace_init(void *ws, size_t ws_bytes, int B, void *stream);
int radnet_rpn_losses(const double *y_rpn_cls, const double *y_rpn_regr, const float *p_cls,
                      const float *p_regr, int B, int H, int W, int A, float *loss, void *ws,
                      size_t ws_bytes, void *stream);
/* y_class [B][R][n_cls] int32, y_regr [B][R][8(n_cls-1)] float64 (radnet_roi_targets_batch); sel [B][n_sel] row
 * indices (radnet_select_samples) or NULL (rows 0..n_sel-1); n_sel_per_panel [B] or NULL (= n_sel);
 * p_cls [B][n_sel][n_cls], p_regr [B][n_sel][4(n_cls-1)] float32 classifier-head outputs for those rows. */
int radnet_class_losses(const int32_t *y_class, const double *y_regr, const int32_t *sel,
                        const int32_t *n_sel_per_panel, int B, int R, int n_cls, int n_sel,
                        const float *p_cls, const float *p_regr, float *loss, void *stream);

/* get_objects (reference test.py:48-113): greedy matching of n_det detections, visited in descending score order
 * (ties: higher index first), each to the first not yet matched figure of its class with utils.iou >= thr - one
 * pool of figures, as the reference concatenates the whole test set (test.py:221-227).
 *   det_box [n_det][4] float64 x1,y1,x2,y2, det_cls [n_det] int32, det_prob [n_det] float64; gt_box / gt_cls alike.
 *   visit [n_det] int32: detection visited at rank r; match [n_det] int32: figure it matched or -1. */
size_t radnet_match_detections_workspace_bytes(int n_det, int n_gt);
int radnet_match_detections(const double *det_box, const int32_t *det_cls, const double *det_prob, int n_det,
                            const double *gt_box, const int32_t *gt_cls, int n_gt, double thr,
                            int32_t *visit, int32_t *match, void *ws, size_t ws_bytes, void *stream);

/* calc_class_ap (reference test.py:117-173) for one class: y_true [n] int32, y_pred [n] float64 ->
 * precision / recall / interpolated precision / interpolated recall [n] float64 in visiting order (descending
 * score) and ap [1] float64 (the reference adds its terms left to right; here the sum is a tree: 1e-12 relative). */
size_t radnet_class_ap_workspace_bytes(int n);
int radnet_class_ap(const int32_t *y_true, const double *y_pred, int n, double *precision, double *recall,
                    double *interp_precision, double *interp_recall, double *ap, void *ws, size_t ws_bytes,
                    void *stream);

/* ------------------------------------------------ bench / test utility: synthetic panels on the device
 * BASELINE configs[4] (the 10,000 panels of the archive sweep are generated on the
 * device from their seed).  Counter-based: every value is a hash of (seed, panel id, tensor, element), so panel
 * first_panel + b*panel_stride is the same bytes wherever and in whatever batch it is generated.
 *   cls [B][H][W][A] uniform scores in (0,1); regr [B][H][W][4A] 0.5*N(0,1); feat [B][H][W][C] N(0,1). */
int radnet_synth_panels(unsigned long long seed, long long first_panel, long long panel_stride, int B,
                        int H, int W, int A, int C, float *cls, float *regr, float *feat, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RADNET_B200_H */

// losses.cu - f4: the four training losses of the reference (faster_rcnn/losses.py:16-95) as fused masked
// reductions over the target layouts K3 / a4 write, one value per panel (the reference trains
// with batch size 1: each value is what Keras would report for that image).
//
// Every element is evaluated in float32 with one rounding per Keras / TF backend operation (no FMA contraction:
// the library is built with -fmad=false); the sums are accumulated in float64 in a FIXED order (per-thread strided
// partial sums, tree inside the CTA, partials of a panel's CTAs added in CTA order by the last one to finish), so
// results are deterministic run to run.  TF's own reduction order is unspecified: parity bar 1e-5 relative.
//
//   rpn_loss_regr  (losses.py:16-44)   smooth-L1 over y_rpn_regr [H][W][8A] = [repeat(overlap,4) | regr*std] (float64
//                                      as K3 writes it, cast to float32 like a Keras feed) vs p_regr [H][W][4A]
//   rpn_loss_cls   (losses.py:47-67)   K.binary_crossentropy(y_pred, y_true[..., A:]) masked by valid.  NOTE the
//                                      argument order: the prediction sits in the `target` slot and the 0/1 label in
//                                      the `output` slot, which Keras clips to [1e-7, 1-1e-7] and turns into a logit
//   class_loss_regr (losses.py:70-88)  smooth-L1 over Y2 rows [8(n_cls-1)] = [labels | coords] vs p_regr rows
//   class_loss_cls  (losses.py:93-95)  mean over rows of keras.objectives.categorical_crossentropy
#include "common.cuh"

namespace radnet {

constexpr int kLossThreads = 256;
constexpr int kLossCtasPerPanel = 16;
constexpr float kLossEps = 1e-4f;     // losses.py:14
constexpr float kKerasEps = 1e-7f;    // keras.backend.epsilon()

__device__ __forceinline__ float smooth_l1(float mask, float t, float pred) {
    const float x = __fsub_rn(t, pred);
    const float xa = fabsf(x);
    const float xb = xa <= 1.0f ? 1.0f : 0.0f;
    // mask * (x_bool * (0.5*x*x) + (1 - x_bool) * (x_abs - 0.5))
    const float a = __fmul_rn(xb, __fmul_rn(__fmul_rn(0.5f, x), x));
    const float b = __fmul_rn(__fsub_rn(1.0f, xb), __fsub_rn(xa, 0.5f));
    return __fmul_rn(mask, __fadd_rn(a, b));
}

// K.binary_crossentropy(target, output) of Keras 2.2 on the TF backend, float32
__device__ __forceinline__ float keras_bce(float target, float output) {
    const float o = fminf(fmaxf(output, kKerasEps), __fsub_rn(1.0f, kKerasEps));
    const float x = logf(__fdiv_rn(o, __fsub_rn(1.0f, o)));                        // logit
    // tf.nn.sigmoid_cross_entropy_with_logits: max(x, 0) - x*z + log(1 + exp(-|x|))
    return __fadd_rn(__fsub_rn(fmaxf(x, 0.0f), __fmul_rn(x, target)), log1pf(expf(-fabsf(x))));
}

// block sum of a double, result valid in thread 0 (fixed tree)
__device__ double block_sum(double v, double *s_red) {
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
    return t;
}

struct RpnLossParams {
    const double *y_cls;      // [B][HW][2A]
    const double *y_regr;     // [B][HW][8A]
    const float *p_cls;       // [B][HW][A]
    const float *p_regr;      // [B][HW][4A]
    int HW, A;
    float *loss;              // [B][2] {rpn_loss_cls, rpn_loss_regr}
    double *partial;          // [B][kLossCtasPerPanel][4]
    int32_t *done;            // [B], zero between launches
};

__global__ void __launch_bounds__(kLossThreads) rpn_losses_kernel(RpnLossParams p) {
    __shared__ double s_red[kLossThreads / 32];
    __shared__ int s_last;
    const int b = blockIdx.y, part = blockIdx.x;
    const int A = p.A;
    const long long n_cell = p.HW;
    const double *yc = p.y_cls + (size_t)b * n_cell * 2 * A;
    const double *yr = p.y_regr + (size_t)b * n_cell * 8 * A;
    const float *pc = p.p_cls + (size_t)b * n_cell * A;
    const float *pr = p.p_regr + (size_t)b * n_cell * 4 * A;
    double cls_num = 0.0, cls_den = 0.0, reg_num = 0.0, reg_den = 0.0;
    // classification: one thread per (cell, anchor)
    const long long n1 = n_cell * A;
    for (long long i = (long long)part * kLossThreads + threadIdx.x; i < n1; i += (long long)kLossCtasPerPanel * kLossThreads) {
        const long long cell = i / A;
        const int a = (int)(i - cell * A);
        const float valid = (float)yc[cell * 2 * A + a], label = (float)yc[cell * 2 * A + A + a];
        cls_num += (double)__fmul_rn(valid, keras_bce(pc[i], label));
        cls_den += (double)__fadd_rn(kLossEps, valid);
    }
    // regression: one thread per (cell, 4A channel)
    const long long n4 = n_cell * 4 * A;
    for (long long i = (long long)part * kLossThreads + threadIdx.x; i < n4; i += (long long)kLossCtasPerPanel * kLossThreads) {
        const long long cell = i / (4 * A);
        const int c = (int)(i - cell * 4 * A);
        const float mask = (float)yr[cell * 8 * A + c], t = (float)yr[cell * 8 * A + 4 * A + c];
        reg_num += (double)smooth_l1(mask, t, pr[i]);
        reg_den += (double)__fadd_rn(kLossEps, mask);
    }
    double *mine = p.partial + ((size_t)b * kLossCtasPerPanel + part) * 4;
    const double s0 = block_sum(cls_num, s_red), s1 = block_sum(cls_den, s_red);
    const double s2 = block_sum(reg_num, s_red), s3 = block_sum(reg_den, s_red);
    if (threadIdx.x == 0) {
        mine[0] = s0; mine[1] = s1; mine[2] = s2; mine[3] = s3;
        __threadfence();
        s_last = atomicAdd(&p.done[b], 1) == kLossCtasPerPanel - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        double t[4] = {0, 0, 0, 0};
        for (int k = 0; k < kLossCtasPerPanel; ++k)
            for (int j = 0; j < 4; ++j) t[j] += __ldcg(p.partial + ((size_t)b * kLossCtasPerPanel + k) * 4 + j);
        // K.sum gives float32 sums; lambda_rpn_* = 1.0
        p.loss[2 * b + 0] = __fdiv_rn((float)t[0], (float)t[1]);
        p.loss[2 * b + 1] = __fdiv_rn((float)t[2], (float)t[3]);
        p.done[b] = 0;
    }
}

struct ClassLossParams {
    const int32_t *y_class;   // [B][R][n_cls] one-hot rows of calc_iou
    const double *y_regr;     // [B][R][8(n_cls-1)]
    const int32_t *sel;       // [B][n_sel] row indices or NULL (rows 0..n_sel-1)
    const int32_t *n_sel_b;   // [B] rows used per panel or NULL (= n_sel)
    int R, n_cls, n_sel;
    const float *p_cls;       // [B][n_sel][n_cls]
    const float *p_regr;      // [B][n_sel][4(n_cls-1)]
    float *loss;              // [B][2] {class_loss_cls, class_loss_regr}
};

__global__ void __launch_bounds__(kLossThreads) class_losses_kernel(ClassLossParams p) {
    __shared__ double s_red[kLossThreads / 32];
    const int b = blockIdx.x;
    const int nc = p.n_cls, n4 = 4 * (nc - 1);
    const int rows = p.n_sel_b ? min(max(p.n_sel_b[b], 0), p.n_sel) : p.n_sel;
    const int32_t *yc = p.y_class + (size_t)b * p.R * nc;
    const double *yr = p.y_regr + (size_t)b * p.R * 2 * n4;
    const int32_t *sel = p.sel ? p.sel + (size_t)b * p.n_sel : nullptr;
    const float *pc = p.p_cls + (size_t)b * p.n_sel * nc;
    const float *pr = p.p_regr + (size_t)b * p.n_sel * n4;
    // categorical cross-entropy, one thread per row: output /= sum(output); clip; -sum(target * log(output))
    double ce = 0.0;
    for (int r = threadIdx.x; r < rows; r += kLossThreads) {
        const int src = sel ? sel[r] : r;
        float s = 0.0f;
        for (int c = 0; c < nc; ++c) s = __fadd_rn(s, pc[r * nc + c]);
        float acc = 0.0f;
        for (int c = 0; c < nc; ++c) {
            const float o = fminf(fmaxf(__fdiv_rn(pc[r * nc + c], s), kKerasEps), __fsub_rn(1.0f, kKerasEps));
            acc = __fadd_rn(acc, __fmul_rn((float)yc[(size_t)src * nc + c], logf(o)));
        }
        ce += (double)(-acc);
    }
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < rows * n4; i += kLossThreads) {
        const int r = i / n4, c = i - r * n4;
        const int src = sel ? sel[r] : r;
        const float mask = (float)yr[(size_t)src * 2 * n4 + c], t = (float)yr[(size_t)src * 2 * n4 + n4 + c];
        num += (double)smooth_l1(mask, t, pr[i]);
        den += (double)__fadd_rn(kLossEps, mask);
    }
    const double s0 = block_sum(ce, s_red), s1 = block_sum(num, s_red), s2 = block_sum(den, s_red);
    if (threadIdx.x == 0) {
        p.loss[2 * b + 0] = rows > 0 ? (float)(s0 / rows) : __int_as_float(0x7fc00000);      // K.mean of nothing = NaN
        p.loss[2 * b + 1] = __fdiv_rn((float)s1, (float)s2);
    }
}

}  // namespace radnet

using namespace radnet;

extern "C" size_t radnet_rpn_losses_workspace_bytes(int B) {
    if (B < 1) return 0;
    return align_up((size_t)B * sizeof(int32_t), 256) + (size_t)B * kLossCtasPerPanel * 4 * sizeof(double);
}

extern "C" int radnet_rpn_losses_workspace_init(void *ws, size_t ws_bytes, int B, void *stream) {
    RADNET_CHECK_ARG(ws && B >= 1 && ws_bytes >= radnet_rpn_losses_workspace_bytes(B), "rpn_losses_workspace_init: bad arguments");
    RADNET_CUDA(cudaMemsetAsync(ws, 0, align_up((size_t)B * sizeof(int32_t), 256), (cudaStream_t)stream));
    return RADNET_OK;
}

extern "C" int radnet_rpn_losses(const double *y_rpn_cls, const double *y_rpn_regr, const float *p_cls,
                                 const float *p_regr, int B, int H, int W, int A, float *loss, void *ws,
                                 size_t ws_bytes, void *stream) {
    RADNET_CHECK_ARG(y_rpn_cls && y_rpn_regr && p_cls && p_regr && loss && ws, "rpn_losses: null pointer");
    RADNET_CHECK_ARG(B >= 1 && B <= 65535 && H >= 1 && W >= 1 && A >= 1, "rpn_losses: bad sizes");
    if (ws_bytes < radnet_rpn_losses_workspace_bytes(B)) {
        set_error("rpn_losses: workspace %zu < %zu", ws_bytes, radnet_rpn_losses_workspace_bytes(B));
        return RADNET_E_WORKSPACE;
    }
    RpnLossParams p{};
    p.y_cls = y_rpn_cls; p.y_regr = y_rpn_regr; p.p_cls = p_cls; p.p_regr = p_regr;
    p.HW = H * W; p.A = A; p.loss = loss;
    p.done = reinterpret_cast<int32_t *>(ws);
    p.partial = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(ws) + align_up((size_t)B * sizeof(int32_t), 256));
    rpn_losses_kernel<<<dim3(kLossCtasPerPanel, B), kLossThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch("rpn_losses_kernel");
}

extern "C" int radnet_class_losses(const int32_t *y_class, const double *y_regr, const int32_t *sel,
                                   const int32_t *n_sel_per_panel, int B, int R, int n_cls, int n_sel,
                                   const float *p_cls, const float *p_regr, float *loss, void *stream) {
    RADNET_CHECK_ARG(y_class && y_regr && p_cls && p_regr && loss, "class_losses: null pointer");
    RADNET_CHECK_ARG(B >= 1 && R >= 1 && n_cls >= 2 && n_sel >= 0, "class_losses: bad sizes");
    ClassLossParams p{};
    p.y_class = y_class; p.y_regr = y_regr; p.sel = sel; p.n_sel_b = n_sel_per_panel;
    p.R = R; p.n_cls = n_cls; p.n_sel = n_sel; p.p_cls = p_cls; p.p_regr = p_regr; p.loss = loss;
    class_losses_kernel<<<B, kLossThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch("class_losses_kernel");
}

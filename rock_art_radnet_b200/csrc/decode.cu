// decode.cu - K1: anchor decode + clip (reference faster_rcnn/rpn.py:91-166, 299-344).
//
// One thread per (cell, anchor) pair in the INPUT order, so the float4 regression
// loads and the score loads are fully coalesced; results are transposed through
// shared memory and leave in the reference's anchor-major flat order
// i = a*H*W + r*W + c with 16-byte coalesced stores per anchor row.
//
// Arithmetic contract (bit-exact against the NumPy reference, except exp() ulps):
//   t = regr / std_scaling           float32 IEEE divide            (rpn.py:91)
//   all box math in float64, one rounding per NumPy operation, no FMA contraction
//   (the library is compiled with -fmad=false and the products below use __dmul_rn)
//   np.round == rint() (half to even)                                (rpn.py:335-338)
//   x/2. is evaluated as x*0.5: both are exact in binary floating point, the multiply is
//   one DMUL instead of a ~25-instruction DDIV sequence
//
// Float32 shortcut: the four values that get rounded (x1, y1, w1, h1) are first evaluated in float32.
// With every magnitude below 128 the float32 result is within 7 * 2^-24 * 128 = 5.4e-5 of the float64 one
// (three roundings on the centre, six on the size incl. expf's 2 ulp, one on the difference), so whenever
// it lies more than 2e-4 away from a half-integer, rintf() of it IS np.round() of the reference value;
// the float64 path (two exp() in double) then runs only for the ~0.2 % of anchors near a rounding
// boundary, for large maps / sizes, and for non-finite inputs.  Outputs stay bit-exact.
#include <math.h>

#include "common.cuh"

namespace radnet {

constexpr int kDecodeCells = 128;     // cells per CTA (4.5 anchors per thread at A = 9: loads of several anchors in flight)
static_assert(kDecodeCells == 128, "the write-out loop shifts by 7");
constexpr int kDecodeThreads = 256;

struct DecodedBox {
    double x1, y1, x2, y2;
    int flags;   // bit0 valid, bit1 nonfinite, bit2 round near-tie, bit3 non-integer coordinate
};

__device__ __forceinline__ int near_half(double v) {
    // |frac(v) - 0.5| tiny: an exp() ulp difference against NumPy could flip np.round
    double f = v - floor(v);
    return fabs(f - 0.5) < 1e-9 ? 4 : 0;
}

__device__ __forceinline__ DecodedBox decode_one(int c, int r, double aw, double ah, float4 t,
                                                 int use_regr, int rows, int cols) {
    DecodedBox o;
    double x = __dsub_rn((double)c, __dmul_rn(aw, 0.5));     // X - anchor_x/2   (rpn.py:127)
    double y = __dsub_rn((double)r, __dmul_rn(ah, 0.5));     // Y - anchor_y/2   (rpn.py:128)
    double w = aw, h = ah;
    int flags = 0;
    if (use_regr) {                                          // apply_regr_np     (rpn.py:325-338)
        double cx = __dadd_rn(x, __dmul_rn(w, 0.5));
        double cy = __dadd_rn(y, __dmul_rn(h, 0.5));
        double cx1 = __dadd_rn(__dmul_rn((double)t.x, w), cx);
        double cy1 = __dadd_rn(__dmul_rn((double)t.y, h), cy);
        double w1 = __dmul_rn(exp((double)t.z), w);
        double h1 = __dmul_rn(exp((double)t.w), h);
        double x1 = __dsub_rn(cx1, __dmul_rn(w1, 0.5));
        double y1 = __dsub_rn(cy1, __dmul_rn(h1, 0.5));
        flags |= near_half(x1) | near_half(y1) | near_half(w1) | near_half(h1);
        x = rint(x1);
        y = rint(y1);
        w = rint(w1);
        h = rint(h1);
    }
    // np.maximum / np.minimum propagate NaN; fmax/fmin do not, so test explicitly
    w = (w != w) ? w : fmax(1.0, w);                         // rpn.py:137
    h = (h != h) ? h : fmax(1.0, h);                         // rpn.py:138
    double x2 = __dadd_rn(w, x);                             // rpn.py:143
    double y2 = __dadd_rn(h, y);                             // rpn.py:144
    x = (x != x) ? x : fmax(0.0, x);                         // rpn.py:147
    y = (y != y) ? y : fmax(0.0, y);                         // rpn.py:148
    x2 = (x2 != x2) ? x2 : fmin((double)(cols - 1), x2);     // rpn.py:149
    y2 = (y2 != y2) ? y2 : fmin((double)(rows - 1), y2);     // rpn.py:150
    bool nan_any = (x != x) || (y != y) || (x2 != x2) || (y2 != y2);
    // rows the reference deletes: (x1 - x2 >= 0) | (y1 - y2 >= 0)         (rpn.py:163)
    bool drop = (__dsub_rn(x, x2) >= 0.0) || (__dsub_rn(y, y2) >= 0.0);
    if (!drop) flags |= 1;
    if (nan_any && !drop) flags |= 2;   // survives the delete and trips the NMS assert (rpn.py:400-401)
    if (!drop && !nan_any &&
        (x != rint(x) || y != rint(y) || x2 != rint(x2) || y2 != rint(y2)))
        flags |= 8;
    o.x1 = x; o.y1 = y; o.x2 = x2; o.y2 = y2; o.flags = flags;
    return o;
}

// float32 evaluation of one anchor; returns false when the float64 path has to decide
__device__ __forceinline__ bool decode_fast(int c, int r, float wf, float hf, float4 t, int rows, int cols,
                                            int4 &box, int &valid) {
    const float txw = t.x * wf, tyh = t.y * hf;
    const float cx1 = txw + (float)c, cy1 = tyh + (float)r;            // anchor centre is the cell corner (rpn.py:127-128, 325-326)
    const float w1 = expf(t.z) * wf, h1 = expf(t.w) * hf;
    const float x1 = cx1 - 0.5f * w1, y1 = cy1 - 0.5f * h1;
    const float big = fmaxf(fmaxf(fmaxf(fabsf(txw), fabsf(tyh)), fmaxf(fabsf(cx1), fabsf(cy1))),
                            fmaxf(fmaxf(w1, h1), fmaxf(fabsf(x1), fabsf(y1))));
    // distance of each rounded value to the nearest half-integer
    const float dx = fabsf(x1 - floorf(x1) - 0.5f), dy = fabsf(y1 - floorf(y1) - 0.5f);
    const float dw = fabsf(w1 - floorf(w1) - 0.5f), dh = fabsf(h1 - floorf(h1) - 0.5f);
    const float near = fminf(fminf(dx, dy), fminf(dw, dh));
    // fmaxf / fminf drop NaN operands, so NaN is tested explicitly; inf fails the magnitude test
    const bool no_nan = x1 == x1 && y1 == y1 && w1 == w1 && h1 == h1;
    if (!no_nan || !(big < 128.f) || !(near > 2e-4f)) return false;
    int x = __float2int_rn(x1), y = __float2int_rn(y1);
    int w = max(1, __float2int_rn(w1)), h = max(1, __float2int_rn(h1));   // rpn.py:137-138
    int x2 = w + x, y2 = h + y;                                        // rpn.py:143-144
    x = max(0, x); y = max(0, y);                                      // rpn.py:147-148
    x2 = min(cols - 1, x2); y2 = min(rows - 1, y2);                    // rpn.py:149-150
    valid = !((x - x2 >= 0) || (y - y2 >= 0));                         // rpn.py:163
    box = make_int4(x, y, x2, y2);
    return true;
}

template <bool kF64>
__global__ void __launch_bounds__(kDecodeThreads, kF64 ? 2 : 4)
decode_clip_kernel(const float *__restrict__ cls, const float *__restrict__ regr, int H, int W,
                   int A, unsigned magic_a, unsigned magic_w, AnchorTable anchors, float std_scaling, float inv_scale, int use_regr,
                   int32_t *__restrict__ boxes_i32, uint32_t *__restrict__ keys,
                   double *__restrict__ boxes_f64, float *__restrict__ scores,
                   uint8_t *__restrict__ valid, int32_t *__restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // staging: per anchor row kDecodeCells(+1 pad) entries
    constexpr int kPad = kDecodeCells + 1;
    int4 *s_box = reinterpret_cast<int4 *>(smem_raw);                       // [A][kPad] (i32 path)
    double4 *s_boxd = reinterpret_cast<double4 *>(smem_raw);                // [A][kPad] (f64 path)
    uint32_t *s_key = reinterpret_cast<uint32_t *>(
        smem_raw + (size_t)A * kPad * (kF64 ? sizeof(double4) : sizeof(int4)));  // [A][kPad]
    __shared__ int s_stats[4];

    const int HW = H * W;
    const int b = blockIdx.y;
    const int cell0 = blockIdx.x * kDecodeCells;
    const int ncell = min(kDecodeCells, HW - cell0);
    const size_t N = (size_t)HW * A;
    if (threadIdx.x < 4) s_stats[threadIdx.x] = 0;
    __syncthreads();

    const float *cls_b = cls + (size_t)b * N + (size_t)cell0 * A;
    const float4 *regr_b = reinterpret_cast<const float4 *>(regr + ((size_t)b * N + (size_t)cell0 * A) * 4);
    int n_valid = 0, n_nonfinite = 0, n_tie = 0, n_nonint = 0;
    // the loads of the next anchor of this thread are issued before the current one is decoded (the decode is a
    // ~120-instruction dependent chain; without the prefetch every anchor exposes one trip to L2 / HBM)
    float4 t_next = make_float4(0.f, 0.f, 0.f, 0.f);
    float s_next = 0.f;
    if ((int)threadIdx.x < ncell * A) {
        t_next = __ldg(regr_b + threadIdx.x);
        s_next = __ldg(cls_b + threadIdx.x);
    }
    for (int p = threadIdx.x; p < ncell * A; p += kDecodeThreads) {
        // p / A and cell / W by multiply-high with host-made reciprocals (exact in the checked ranges)
        const int cl = magic_a ? (int)__umulhi((unsigned)p, magic_a) : p / A;
        const int a = p - cl * A;
        const int cell = cell0 + cl;
        const int r = magic_w ? (int)__umulhi((unsigned)cell, magic_w) : cell / W;
        const int c = cell - r * W;
        float4 t = t_next;
        float s = s_next;
        if (p + kDecodeThreads < ncell * A) {
            t_next = __ldg(regr_b + p + kDecodeThreads);
            s_next = __ldg(cls_b + p + kDecodeThreads);
        }
        if (inv_scale != 0.f) {          // std_scaling is a power of two: the float32 quotient is the exact product
            t.x = __fmul_rn(t.x, inv_scale);
            t.y = __fmul_rn(t.y, inv_scale);
            t.z = __fmul_rn(t.z, inv_scale);
            t.w = __fmul_rn(t.w, inv_scale);
        } else {
            t.x = __fdiv_rn(t.x, std_scaling);
            t.y = __fdiv_rn(t.y, std_scaling);
            t.z = __fdiv_rn(t.z, std_scaling);
            t.w = __fdiv_rn(t.w, std_scaling);
        }
        int4 fbox;
        int fvalid = 0;
        const double aw = anchors.wh[a][0], ah = anchors.wh[a][1];
        if (use_regr && decode_fast(c, r, (float)aw, (float)ah, t, H, W, fbox, fvalid)) {
            n_valid += fvalid;
            if (kF64) {
                s_boxd[a * kPad + cl] = make_double4((double)fbox.x, (double)fbox.y, (double)fbox.z, (double)fbox.w);
                s_key[a * kPad + cl] = __float_as_uint(s);
                reinterpret_cast<uint8_t *>(s_key + (size_t)A * kPad)[a * kPad + cl] = (uint8_t)fvalid;
            } else {
                s_box[a * kPad + cl] = fbox;
                s_key[a * kPad + cl] = fvalid ? score_to_key(s) : 0u;
            }
            continue;
        }
        DecodedBox d = decode_one(c, r, aw, ah, t, use_regr, H, W);
        n_valid += d.flags & 1;
        n_nonfinite += (d.flags >> 1) & 1;
        n_tie += (d.flags >> 2) & 1;
        n_nonint += (d.flags >> 3) & 1;
        if (kF64) {
            s_boxd[a * kPad + cl] = make_double4(d.x1, d.y1, d.x2, d.y2);
            // key slot carries the raw score bits and validity in the f64 path
            s_key[a * kPad + cl] = __float_as_uint(s);
            reinterpret_cast<uint8_t *>(s_key + (size_t)A * kPad)[a * kPad + cl] = (uint8_t)(d.flags & 1);
        } else {
            // saturating conversions; deleted / non-finite boxes never reach the NMS
            int4 bi = make_int4(__double2int_rn(d.x1), __double2int_rn(d.y1), __double2int_rn(d.x2),
                                __double2int_rn(d.y2));
            s_box[a * kPad + cl] = bi;
            s_key[a * kPad + cl] = ((d.flags & 3) == 1) ? score_to_key(s) : 0u;
        }
    }
    // block-level stats
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    n_nonfinite = __reduce_add_sync(0xffffffffu, n_nonfinite);
    n_tie = __reduce_add_sync(0xffffffffu, n_tie);
    n_nonint = __reduce_add_sync(0xffffffffu, n_nonint);
    if ((threadIdx.x & 31) == 0) {
        if (n_valid) atomicAdd(&s_stats[0], n_valid);
        if (n_nonfinite) atomicAdd(&s_stats[1], n_nonfinite);
        if (n_tie) atomicAdd(&s_stats[2], n_tie);
        if (n_nonint) atomicAdd(&s_stats[3], n_nonint);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_stats[threadIdx.x]) atomicAdd(&stats[b * 4 + threadIdx.x], s_stats[threadIdx.x]);

    // transposed, coalesced write-out: anchor-major rows of ncell entries
    for (int p = threadIdx.x; p < ncell * A; p += kDecodeThreads) {
        const int a = (ncell == kDecodeCells) ? (p >> 7) : p / ncell;
        const int cl = p - a * ncell;
        size_t o = (size_t)b * N + (size_t)a * HW + cell0 + cl;
        if (kF64) {
            double4 v = s_boxd[a * kPad + cl];
            reinterpret_cast<double4 *>(boxes_f64)[o] = v;
            scores[o] = __uint_as_float(s_key[a * kPad + cl]);
            valid[o] = reinterpret_cast<uint8_t *>(s_key + (size_t)A * kPad)[a * kPad + cl];
        } else {
            reinterpret_cast<int4 *>(boxes_i32)[o] = s_box[a * kPad + cl];
            keys[o] = s_key[a * kPad + cl];
        }
    }
}

// apply_regr_np as a flat elementwise kernel (API parity; not on the batched hot path)
__global__ void apply_regr_kernel(const double *__restrict__ X, const double *__restrict__ T,
                                  long long n, double *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = X[i], y = X[n + i], w = X[2 * n + i], h = X[3 * n + i];
    double tx = T[i], ty = T[n + i], tw = T[2 * n + i], th = T[3 * n + i];
    double cx = __dadd_rn(x, __dmul_rn(w, 0.5));
    double cy = __dadd_rn(y, __dmul_rn(h, 0.5));
    double cx1 = __dadd_rn(__dmul_rn(tx, w), cx);
    double cy1 = __dadd_rn(__dmul_rn(ty, h), cy);
    double w1 = __dmul_rn(exp(tw), w);
    double h1 = __dmul_rn(exp(th), h);
    out[i] = rint(__dsub_rn(cx1, __dmul_rn(w1, 0.5)));
    out[n + i] = rint(__dsub_rn(cy1, __dmul_rn(h1, 0.5)));
    out[2 * n + i] = rint(w1);
    out[3 * n + i] = rint(h1);
}

static int decode_common(bool f64, const float *cls, const float *regr, int B, int H, int W, int A,
                         const double *h_anchor_wh, float std_scaling, int use_regr,
                         int32_t *boxes_i32, uint32_t *keys, double *boxes_f64, float *scores,
                         uint8_t *valid, int32_t *stats, void *stream) {
    RADNET_CHECK_ARG(cls && regr && h_anchor_wh && stats, "decode_clip: null pointer");
    RADNET_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && A >= 1 && A <= kMaxAnchors,
                     "decode_clip: bad shape B=%d H=%d W=%d A=%d (A<=%d)", B, H, W, A, kMaxAnchors);
    RADNET_CHECK_ARG(B <= 65535, "decode_clip: B=%d exceeds grid.y", B);
    RADNET_CHECK_ARG(std_scaling != 0.f, "decode_clip: std_scaling == 0");
    if (f64) RADNET_CHECK_ARG(boxes_f64 && scores && valid, "decode_clip_f64: null output");
    else RADNET_CHECK_ARG(boxes_i32 && keys, "decode_clip_i32: null output");
    AnchorTable tab;
    for (int a = 0; a < A; ++a) {
        tab.wh[a][0] = h_anchor_wh[2 * a];
        tab.wh[a][1] = h_anchor_wh[2 * a + 1];
    }
    cudaStream_t st = (cudaStream_t)stream;
    RADNET_CUDA(cudaMemsetAsync(stats, 0, sizeof(int32_t) * 4 * B, st));
    const int HW = H * W;
    // floor(n / d) == umulhi(n, ceil(2^32 / d)) whenever n * d < 2^32
    const unsigned magic_a = A > 1 ? (unsigned)((0x100000000ULL + (unsigned)A - 1) / (unsigned)A) : 0u;   // n < 64 * 64
    const unsigned magic_w = ((unsigned long long)HW * (unsigned)W < 0x100000000ULL && W > 1)
                                 ? (unsigned)((0x100000000ULL + (unsigned)W - 1) / (unsigned)W) : 0u;
    // x / s == x * (1 / s) bit for bit when s is a power of two whose reciprocal is a normal float32 (the reference's
    // std_scaling is 4.0); denormal quotients round identically because the product is exact before rounding
    float inv_scale = 0.f;
    {
        int e = 0;
        const float m = frexpf(fabsf(std_scaling), &e);
        if (m == 0.5f && e > -100 && e < 100) inv_scale = 1.0f / std_scaling;
    }
    dim3 grid((HW + kDecodeCells - 1) / kDecodeCells, B);
    size_t per = f64 ? (sizeof(double4) + sizeof(uint32_t) + 1) : (sizeof(int4) + sizeof(uint32_t));
    size_t smem = (size_t)A * (kDecodeCells + 1) * per + 16;
    int dev = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    if (f64) {
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(decode_clip_kernel<true>), dev, smem)) return rc;
        decode_clip_kernel<true><<<grid, kDecodeThreads, smem, st>>>(
            cls, regr, H, W, A, magic_a, magic_w, tab, std_scaling, inv_scale, use_regr, nullptr, nullptr, boxes_f64, scores, valid, stats);
    } else {
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(decode_clip_kernel<false>), dev, smem)) return rc;
        decode_clip_kernel<false><<<grid, kDecodeThreads, smem, st>>>(
            cls, regr, H, W, A, magic_a, magic_w, tab, std_scaling, inv_scale, use_regr, boxes_i32, keys, nullptr, nullptr, nullptr, stats);
    }
    return check_launch("decode_clip_kernel");
}

}  // namespace radnet

extern "C" int radnet_decode_clip_i32(const float *cls, const float *regr, int B, int H, int W, int A,
                                      const double *h_anchor_wh, float std_scaling, int use_regr,
                                      int32_t *boxes_i32, uint32_t *keys, int32_t *stats, void *stream) {
    return radnet::decode_common(false, cls, regr, B, H, W, A, h_anchor_wh, std_scaling, use_regr,
                                 boxes_i32, keys, nullptr, nullptr, nullptr, stats, stream);
}

extern "C" int radnet_decode_clip_f64(const float *cls, const float *regr, int B, int H, int W, int A,
                                      const double *h_anchor_wh, float std_scaling, int use_regr,
                                      double *boxes_f64, float *scores, uint8_t *valid, int32_t *stats,
                                      void *stream) {
    return radnet::decode_common(true, cls, regr, B, H, W, A, h_anchor_wh, std_scaling, use_regr,
                                 nullptr, nullptr, boxes_f64, scores, valid, stats, stream);
}

extern "C" int radnet_apply_regr(const double *X, const double *T, long long n, double *out, void *stream) {
    RADNET_CHECK_ARG(X && T && out && n >= 0, "apply_regr: bad arguments");
    if (n == 0) return RADNET_OK;
    long long blocks = (n + 255) / 256;
    RADNET_CHECK_ARG(blocks <= 0x7fffffffLL, "apply_regr: n too large");
    radnet::apply_regr_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(X, T, n, out);
    return radnet::check_launch("apply_regr_kernel");
}

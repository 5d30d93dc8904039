// roipool.cu - K4: RoI crop + TF-1 legacy bilinear resize (reference
// faster_rcnn/RoiPoolingConv.py:48-88, i.e. tf.image.resize_images(crop,(pool,pool)),
// BILINEAR, align_corners=False, no half-pixel centres).
//
// The kernel is bound by the HBM WRITE stream: a 600-px panel reads one 5.9 MB map and
// writes 300 x 14 x 14 x 1024 floats = 240.8 MB.  Each output pixel samples 4 source
// pixels, so reading the map through L2 would cost up to 4x the write traffic.  Design:
//
//   * channel-sliced, map-resident CTAs: a CTA owns (panel, slice of 32 channels) and
//     copies that slice of the WHOLE map (38*38*128 B = 185 KB) into shared memory once
//     (cp.async 16 B, coalesced 128 B per pixel).  The map is then read from HBM exactly
//     once per panel and every bilinear tap is a conflict-free LDS.128.
//   * per chunk of 32 RoIs the CTA tabulates, per axis and output index, the two source
//     offsets and the float32 lerp weight exactly as TF's compute_interpolation_weights
//     does (scale = in/out float divide; in = i*scale; floor/ceil; lerp = in - floor(in)).
//   * 8 lanes x float4 cover the 32 channels of one output pixel; a warp writes 4 full
//     128-byte lines per store instruction with a streaming (evict-first) hint.
//   * value = top + (bottom-top)*ylerp, top = tl + (tr-tl)*xlerp, separate IEEE float32
//     multiply and add (never fused) - bit-identical to the CPU kernel of TF-1.
//
// A direct (global-load) kernel covers maps whose slice does not fit in shared memory
// and channel counts that are not a multiple of 4.
#include "common.cuh"

namespace radnet {

constexpr int kPoolThreads = 512;
constexpr int kRoiChunk = 32;
constexpr int kMaxPool = 32;

struct RoiPoolParams {
    const float *feat;       // [B][H][W][C]
    int H, W, C;
    const unsigned char *det;
    size_t det_stride;
    int det_max_boxes;
    const int32_t *rois;     // [B][R][4] xywh (det == null)
    const int32_t *roi_count;
    int R;                   // slots per panel
    int pool;
    float *out;              // [B][R][pool][pool][C]
    int n_slices;
};

struct AxisEntry {           // 16 bytes, one LDS.128
    int off0, off1;          // byte offsets into the shared slice
    float lerp;
    int pad;
};

// RoI k of panel b as (x,y,w,h), already int32-truncated by the caller (RoiPoolingConv.py:69-72);
// returns false for empty slots / RoIs that crop to nothing
__device__ __forceinline__ bool fetch_roi(const RoiPoolParams &p, int b, int k, int &x, int &y, int &cw, int &ch) {
    int count, w, h;
    if (p.det) {
        const unsigned char *rec = p.det + (size_t)b * p.det_stride;
        count = *reinterpret_cast<const int32_t *>(rec);
        int4 bx = reinterpret_cast<const int4 *>(rec + 16)[k < p.det_max_boxes ? k : 0];
        x = bx.x; y = bx.y; w = bx.z - bx.x; h = bx.w - bx.y;        // RADNet.py:564-565
    } else {
        count = p.roi_count ? p.roi_count[b] : p.R;
        int4 bx = reinterpret_cast<const int4 *>(p.rois)[(size_t)b * p.R + k];
        x = bx.x; y = bx.y; w = bx.z; h = bx.w;
    }
    if (k >= count || x < 0 || y < 0) return false;
    // TF strided-slice clamps the end of img[:, y:y+h, x:x+w, :] to the map
    ch = min(y + h, p.H) - min(y, p.H);
    cw = min(x + w, p.W) - min(x, p.W);
    return ch > 0 && cw > 0;
}

// TF-1 legacy interpolation weights for output index i of an axis with in_size source cells
__device__ __forceinline__ void legacy_axis(int i, int in_size, int pool, int &lo, int &hi, float &lerp) {
    float scale = __fdiv_rn((float)in_size, (float)pool);
    float src = __fmul_rn((float)i, scale);
    float fl = floorf(src);
    lo = max((int)fl, 0);
    hi = min((int)ceilf(src), in_size - 1);
    lerp = __fsub_rn(src, fl);
}

__device__ __forceinline__ float lerp1(float a, float b, float t) {
    return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}
__device__ __forceinline__ float4 lerp4(const float4 &a, const float4 &b, float t) {
    return make_float4(lerp1(a.x, b.x, t), lerp1(a.y, b.y, t), lerp1(a.z, b.z, t), lerp1(a.w, b.w, t));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// LANES = float4 lanes per pixel in the slice (8 -> 32 channels, 4 -> 16, 2 -> 8, 1 -> 4)
template <int LANES>
__global__ void __launch_bounds__(kPoolThreads, 1) roi_pool_slice_kernel(RoiPoolParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kPixBytes = LANES * 16;
    const int HW = p.H * p.W;
    float4 *s_map = reinterpret_cast<float4 *>(smem);                              // [HW+1][LANES]
    AxisEntry *s_tab = reinterpret_cast<AxisEntry *>(smem + (size_t)(HW + 1) * kPixBytes);  // [chunk][2][pool]

    const int b = blockIdx.x / p.n_slices;
    const int s = blockIdx.x - b * p.n_slices;
    const int C4 = p.C >> 2;
    const int q = threadIdx.x % LANES;
    const int g = threadIdx.x / LANES;
    constexpr int G = kPoolThreads / LANES;          // pixel groups per CTA iteration

    // ---- stage the channel slice of the whole map ---------------------------------
    {
        const float4 *src = reinterpret_cast<const float4 *>(p.feat) + (size_t)b * HW * C4 + (size_t)s * LANES + q;
        for (int pix = g; pix < HW; pix += G) cp_async16(&s_map[pix * LANES + q], src + (size_t)pix * C4);
        if (g == 0) s_map[HW * LANES + q] = make_float4(0.f, 0.f, 0.f, 0.f);      // the "zero pixel"
        cp_async_wait_all();
    }
    const int pool = p.pool, PP = pool * pool;
    // per-iteration stepping of (roi, py, px) by G pixels
    const int d_roi = G / PP, d_rem = G % PP, d_py = d_rem / pool, d_px = d_rem % pool;
    const unsigned char *mapb = reinterpret_cast<const unsigned char *>(s_map) + q * 16;

    for (int r0 = 0; r0 < p.R; r0 += kRoiChunk) {
        const int nr = min(kRoiChunk, p.R - r0);
        __syncthreads();     // previous chunk done with the tables (and the map has landed)
        for (int e = threadIdx.x; e < nr * 2 * pool; e += kPoolThreads) {
            int rl = e / (2 * pool);
            int rem = e - rl * 2 * pool;
            int axis = rem / pool;               // 0 = x, 1 = y
            int i = rem - axis * pool;
            int x, y, cw, ch;
            AxisEntry en;
            if (fetch_roi(p, b, r0 + rl, x, y, cw, ch)) {
                int lo, hi;
                legacy_axis(i, axis ? ch : cw, pool, lo, hi, en.lerp);
                int stride = axis ? p.W * kPixBytes : kPixBytes;
                int org = axis ? y : x;
                en.off0 = (org + lo) * stride;
                en.off1 = (org + hi) * stride;
            } else {
                // x entries point at the zero pixel, y entries add nothing: output is exactly 0
                en.off0 = en.off1 = axis ? 0 : HW * kPixBytes;
                en.lerp = 0.f;
            }
            en.pad = 0;
            s_tab[e] = en;
        }
        __syncthreads();

        // walk the chunk's output pixels: item = rl*PP + py*pool + px, G items per iteration
        int rl = g / PP, rem = g - rl * PP;
        int py = rem / pool, px = rem - py * pool;
        float4 *dst = reinterpret_cast<float4 *>(p.out) +
                      (((size_t)b * p.R + r0) * PP + g) * C4 + (size_t)s * LANES + q;
        const size_t dst_step = (size_t)G * C4;
        while (rl < nr) {
            const AxisEntry ex = s_tab[(rl * 2 + 0) * pool + px];
            const AxisEntry ey = s_tab[(rl * 2 + 1) * pool + py];
            const float4 tl = *reinterpret_cast<const float4 *>(mapb + ey.off0 + ex.off0);
            const float4 tr = *reinterpret_cast<const float4 *>(mapb + ey.off0 + ex.off1);
            const float4 bl = *reinterpret_cast<const float4 *>(mapb + ey.off1 + ex.off0);
            const float4 br = *reinterpret_cast<const float4 *>(mapb + ey.off1 + ex.off1);
            const float4 top = lerp4(tl, tr, ex.lerp);
            const float4 bot = lerp4(bl, br, ex.lerp);
            st_stream_f4(dst, lerp4(top, bot, ey.lerp));
            dst += dst_step;
            rl += d_roi; py += d_py; px += d_px;
            if (px >= pool) { px -= pool; ++py; }
            if (py >= pool) { py -= pool; ++rl; }
        }
    }
}

// Direct kernel: one CTA per (roi slot, output row); threads stride over px and channels.
// VEC = 4 (C % 4 == 0, float4 path) or 1.
template <int VEC>
__global__ void __launch_bounds__(256) roi_pool_direct_kernel(RoiPoolParams p) {
    const int pool = p.pool;
    const int slot = blockIdx.x / pool;          // b*R + k
    const int py = blockIdx.x - slot * pool;
    const int b = slot / p.R, k = slot - b * p.R;
    int x, y, cw, ch;
    const bool ok = fetch_roi(p, b, k, x, y, cw, ch);
    const int CV = p.C / VEC;
    float *orow = p.out + ((size_t)slot * pool + py) * pool * p.C;
    if (!ok) {
        for (int i = threadIdx.x; i < pool * p.C; i += blockDim.x) orow[i] = 0.f;
        return;
    }
    int y0, y1;
    float ly;
    legacy_axis(py, ch, pool, y0, y1, ly);
    const float *r0 = p.feat + (((size_t)b * p.H + y + y0) * p.W + x) * p.C;
    const float *r1 = p.feat + (((size_t)b * p.H + y + y1) * p.W + x) * p.C;
    for (int i = threadIdx.x; i < pool * CV; i += blockDim.x) {
        int px = i / CV, c = i - px * CV;
        int x0, x1;
        float lx;
        legacy_axis(px, cw, pool, x0, x1, lx);
        if (VEC == 4) {
            const float4 tl = __ldg(reinterpret_cast<const float4 *>(r0 + (size_t)x0 * p.C) + c);
            const float4 tr = __ldg(reinterpret_cast<const float4 *>(r0 + (size_t)x1 * p.C) + c);
            const float4 bl = __ldg(reinterpret_cast<const float4 *>(r1 + (size_t)x0 * p.C) + c);
            const float4 br = __ldg(reinterpret_cast<const float4 *>(r1 + (size_t)x1 * p.C) + c);
            st_stream_f4(reinterpret_cast<float4 *>(orow + (size_t)px * p.C) + c,
                         lerp4(lerp4(tl, tr, lx), lerp4(bl, br, lx), ly));
        } else {
            float tl = __ldg(r0 + (size_t)x0 * p.C + c), tr = __ldg(r0 + (size_t)x1 * p.C + c);
            float bl = __ldg(r1 + (size_t)x0 * p.C + c), br = __ldg(r1 + (size_t)x1 * p.C + c);
            orow[(size_t)px * p.C + c] = lerp1(lerp1(tl, tr, lx), lerp1(bl, br, lx), ly);
        }
    }
}

template <int LANES>
static int launch_slice(const RoiPoolParams &p, int B, size_t smem, cudaStream_t st) {
    RADNET_CUDA(cudaFuncSetAttribute(roi_pool_slice_kernel<LANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_pool_slice_kernel<LANES><<<B * p.n_slices, kPoolThreads, smem, st>>>(p);
    return check_launch("roi_pool_slice_kernel");
}

}  // namespace radnet

using namespace radnet;

// RADNET_ROIPOOL_FORCE_DIRECT=1 forces the direct kernel (parity tests exercise both)
static bool force_direct() {
    const char *e = getenv("RADNET_ROIPOOL_FORCE_DIRECT");
    return e && e[0] == '1';
}

extern "C" int radnet_roi_pool(const float *feat, int B, int H, int W, int C, const void *det, int det_max_boxes,
                               const int32_t *rois, const int32_t *roi_count, int rois_per_panel, int pool,
                               float *out, void *stream) {
    RADNET_CHECK_ARG(feat && out && (det || rois), "roi_pool: null pointer");
    RADNET_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 1 && rois_per_panel >= 1,
                     "roi_pool: bad sizes B=%d H=%d W=%d C=%d R=%d", B, H, W, C, rois_per_panel);
    RADNET_CHECK_ARG(pool >= 1 && pool <= kMaxPool, "roi_pool: pool=%d out of range [1,%d]", pool, kMaxPool);
    RADNET_CHECK_ARG(!det || rois_per_panel <= det_max_boxes, "roi_pool: rois_per_panel > det_max_boxes");
    RADNET_CHECK_ARG((long long)B * rois_per_panel * pool < 0x7fffffffLL, "roi_pool: grid too large");
    RoiPoolParams p{};
    p.feat = feat; p.H = H; p.W = W; p.C = C;
    p.det = reinterpret_cast<const unsigned char *>(det);
    p.det_stride = radnet_det_record_bytes(det_max_boxes);
    p.det_max_boxes = det_max_boxes;
    p.rois = rois; p.roi_count = roi_count; p.R = rois_per_panel; p.pool = pool; p.out = out;
    cudaStream_t st = (cudaStream_t)stream;

    int dev = 0, smem_limit = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    RADNET_CUDA(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t tab_bytes = (size_t)kRoiChunk * 2 * pool * sizeof(AxisEntry);
    const size_t HW = (size_t)H * W;
    if (C % 4 == 0 && !force_direct() && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const int C4 = C / 4;
        const int lanes_opts[4] = {8, 4, 2, 1};
        for (int li = 0; li < 4; ++li) {
            int L = lanes_opts[li];
            size_t smem = (HW + 1) * L * 16 + tab_bytes;
            if (C4 % L != 0 || smem > (size_t)smem_limit) continue;
            if (HW * L * 16 >= 0x7fffffffULL) continue;
            p.n_slices = C4 / L;
            if ((long long)B * p.n_slices >= 0x7fffffffLL) continue;
            switch (L) {
                case 8: return launch_slice<8>(p, B, smem, st);
                case 4: return launch_slice<4>(p, B, smem, st);
                case 2: return launch_slice<2>(p, B, smem, st);
                default: return launch_slice<1>(p, B, smem, st);
            }
        }
    }
    unsigned grid = (unsigned)((long long)B * rois_per_panel * pool);
    if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        roi_pool_direct_kernel<4><<<grid, 256, 0, st>>>(p);
    else
        roi_pool_direct_kernel<1><<<grid, 256, 0, st>>>(p);
    return check_launch("roi_pool_direct_kernel");
}

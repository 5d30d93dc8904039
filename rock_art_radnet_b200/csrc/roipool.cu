// roipool.cu - K4: RoI crop + TF-1 legacy bilinear resize (reference
// faster_rcnn/RoiPoolingConv.py:48-88, i.e. tf.image.resize_images(crop,(pool,pool)),
// BILINEAR, align_corners=False, no half-pixel centres).
//
// The kernel is bound by the HBM WRITE stream: a 600-px panel reads one 5.9 MB map and
// writes 300 x 14 x 14 x 1024 floats = 240.8 MB.  Each output pixel samples 4 source
// pixels, so reading the map through L2 would cost up to 4x the write traffic.  Design:
//
//   * channel-sliced, map-resident CTAs: a CTA owns (panel, slice of 32 channels) and
//     brings that slice of the WHOLE map (38*38*128 B = 185 KB) into shared memory with ONE
//     TMA tensor copy (cp.async.bulk.tensor.4d over the map seen as (C, W, H, B); SASS
//     UTMALDG) that completes on an mbarrier while the first y tables are being built.  The
//     map is then read from HBM exactly once per panel and every bilinear tap is a
//     conflict-free LDS.128.  (cp.async 16 B is kept for maps wider than a TMA box.)
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
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace radnet {

constexpr int kPoolThreads = 512;
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
    int roi_chunk;           // RoIs whose y tables fit in shared memory at once
    int map_rows_pad;        // map rows held in shared memory (>= H: whole TMA boxes)
    int tma_rows;            // map rows per TMA box (0 = stage with cp.async)
};

// global -> shared TMA tile copy of a rank-4 tensor, completion (bytes) on an mbarrier
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

// Per (RoI, output row) entry of the y table, 8 bytes = one LDS.64.
//   bits  0..15  pixel index of the first cell of source row y+y0  ((y+y0)*W)
//   bit   16     1 when y1 = y0+1 (else y1 = y0)
//   bits 28..30  what the row cache has to do relative to the previous output row (kAct*)
struct YEntry {
    unsigned code;
    float lerp;
};
enum : unsigned {
    kActKeep = 0,      // same two source rows as the previous output row
    kActShift = 1,     // y0 = previous y1, new y1: h0 <- h1, sample h1
    kActShiftDup = 2,  // y0 = y1 = previous y1: h0 <- h1
    kActBoth = 3,      // sample both rows
    kActOneDup = 4,    // y0 = y1, new row: sample h0, h1 <- h0
    kActExtend = 5     // y0 kept, previous rows were equal, new y1: sample h1
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

// a + (b-a)*t on two lanes at once: packed subtract and multiply (FADD2/FMUL2), SCALAR final
// adds.  ptxas 12.9 contracts a packed multiply feeding a packed add into FFMA2 even for
// explicitly rounded PTX (and for the __fmul2_rn/__fadd2_rn intrinsics), which would break
// bit-parity with TF's unfused arithmetic; a scalar add.rn.f32 is never contracted.
// tests/test_build_sass.py checks the SASS of these kernels for FFMA.
__device__ __forceinline__ void lerp2(float a0, float a1, float b0, float b1, unsigned long long tt,
                                      float &r0, float &r1) {
    unsigned long long a, b, d, m;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(b), "l"(a));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(d), "l"(tt));
    float m0, m1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(m0), "=f"(m1) : "l"(m));
    r0 = __fadd_rn(a0, m0);
    r1 = __fadd_rn(a1, m1);
}
__device__ __forceinline__ float4 lerp4(const float4 &a, const float4 &b, float t) {
    unsigned long long tt;
    asm("mov.b64 %0, {%1,%1};" : "=l"(tt) : "f"(t));
    float4 r;
    lerp2(a.x, a.y, b.x, b.y, tt, r.x, r.y);
    lerp2(a.z, a.w, b.z, b.w, tt, r.z, r.w);
    return r;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// LANES = float4 lanes per pixel in the slice (8 -> 32 channels, 4 -> 16, 2 -> 8, 1 -> 4).
//
// Work decomposition: a group of LANES threads owns one output COLUMN (roi, px) and walks
// py = 0..pool-1.  The horizontal interpolation of a source row, hrow(r) = tl + (tr-tl)*xlerp,
// does not depend on py, so the two rows an output needs are cached in registers and reused
// while y0/y1 repeat (always, when the RoI is upsampled: h < pool); only rows that change are
// re-sampled (2 LDS.128 each).  top/bottom of TF's formula are exactly hrow(y0)/hrow(y1), so
// the result is bit-identical to evaluating all four taps per output.
template <int LANES>
__global__ void __launch_bounds__(kPoolThreads, 1) roi_pool_slice_kernel(RoiPoolParams p,
                                                                          const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    constexpr int kPixBytes = LANES * 16;
    constexpr int G = kPoolThreads / LANES;          // columns in flight per CTA
    const int HW = p.H * p.W;
    const int HWp = p.map_rows_pad * p.W;            // pixels held in shared memory (whole TMA boxes)
    const int pool = p.pool, PP = pool * pool;
    float4 *s_map = reinterpret_cast<float4 *>(smem);                                   // [HWp+1][LANES]
    YEntry *s_ytab = reinterpret_cast<YEntry *>(smem + (size_t)(HWp + 1) * kPixBytes);  // [chunk][pool]
    int2 *s_roi = reinterpret_cast<int2 *>(s_ytab + (size_t)p.roi_chunk * pool);        // [chunk] {x, cw}
    float *s_scale = reinterpret_cast<float *>(s_roi + p.roi_chunk);                    // [W+1] cw / pool (float32 divide)

    const int b = blockIdx.x / p.n_slices;
    const int s = blockIdx.x - b * p.n_slices;
    const int C4 = p.C >> 2;
    const int q = threadIdx.x % LANES;
    const int g = threadIdx.x / LANES;

    // ---- stage the channel slice of the whole map (read from HBM exactly once) --------
    if (p.tma_rows > 0) {
        // one elected thread: a TMA tile copy per box of `tma_rows` map rows (one box for maps up to 256 rows),
        // all completing on the same mbarrier; everybody else goes straight on to the first y tables
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
            const int n_box = p.map_rows_pad / p.tma_rows;
            mbar_expect_tx(&s_bar, (uint32_t)((size_t)HWp * kPixBytes));
            for (int k = 0; k < n_box; ++k)
                tma_load_4d(reinterpret_cast<unsigned char *>(s_map) + (size_t)k * p.tma_rows * p.W * kPixBytes, &tmap,
                            s * LANES * 4, 0, k * p.tma_rows, b, &s_bar);
        }
        if (g == 0) s_map[HWp * LANES + q] = make_float4(0.f, 0.f, 0.f, 0.f);     // the "zero pixel"
        for (int i = threadIdx.x; i <= p.W; i += kPoolThreads) s_scale[i] = __fdiv_rn((float)i, (float)p.pool);
    } else {
        const float4 *src = reinterpret_cast<const float4 *>(p.feat) + (size_t)b * HW * C4 + (size_t)s * LANES + q;
        for (int pix = g; pix < HW; pix += G) cp_async16(&s_map[pix * LANES + q], src + (size_t)pix * C4);
        if (g == 0) s_map[HWp * LANES + q] = make_float4(0.f, 0.f, 0.f, 0.f);     // the "zero pixel"
        for (int i = threadIdx.x; i <= p.W; i += kPoolThreads) s_scale[i] = __fdiv_rn((float)i, (float)p.pool);
        cp_async_wait_all();
    }
    bool map_ready = p.tma_rows == 0;
    const unsigned char *mapb = reinterpret_cast<const unsigned char *>(s_map) + q * 16;
    const size_t py_step = (size_t)pool * C4;

    for (int r0 = 0; r0 < p.R; r0 += p.roi_chunk) {
        const int nr = min(p.roi_chunk, p.R - r0);
        __syncthreads();     // previous chunk done with the tables (and the map has landed)
        for (int e = threadIdx.x; e < nr * pool; e += kPoolThreads) {
            int rl = e / pool, i = e - rl * pool;
            int x, y, cw, ch;
            YEntry en;
            if (fetch_roi(p, b, r0 + rl, x, y, cw, ch)) {
                int lo, hi, plo = -1, phi = -1;
                float pl;
                legacy_axis(i, ch, pool, lo, hi, en.lerp);
                if (i > 0) legacy_axis(i - 1, ch, pool, plo, phi, pl);
                unsigned act;
                if (i == 0) act = (hi == lo) ? kActOneDup : kActBoth;
                else if (lo == plo && hi == phi) act = kActKeep;
                else if (lo == phi && hi != lo && phi != plo) act = kActShift;
                else if (lo == phi && hi == lo && phi != plo) act = kActShiftDup;
                else if (lo == plo && phi == plo && hi != lo) act = kActExtend;
                else act = (hi == lo) ? kActOneDup : kActBoth;
                en.code = (unsigned)((y + lo) * p.W) | ((hi != lo) ? 0x10000u : 0u) | (act << 28);
                if (i == 0) s_roi[rl] = make_int2(x, cw);
            } else {
                // rows add nothing and the column points at the zero pixel: output is exactly 0
                en.code = (i == 0) ? (kActOneDup << 28) : (kActKeep << 28);
                en.lerp = 0.f;
                if (i == 0) s_roi[rl] = make_int2(0, 0);
            }
            s_ytab[e] = en;
        }
        __syncthreads();
        if (!map_ready) {            // the barrier above also ordered the mbarrier's initialisation before this wait
            mbar_wait(&s_bar, 0);
            map_ready = true;
        }

        const int ncol = nr * pool;
        // (rl, px) advance incrementally by G columns: no integer division in the column loop
        int rl = g / pool, px = g - rl * pool;
        const int d_rl = G / pool, d_px = G - d_rl * pool;
        for (int col = g; col < ncol; col += G, rl += d_rl, px += d_px) {
            if (px >= pool) { px -= pool; ++rl; }
            const int2 rx = s_roi[rl];
            int xo0, xo1;
            float lx;
            if (rx.y > 0) {
                // TF-1 legacy weights with the float32 scale cw/pool taken from a table
                const float src = __fmul_rn((float)px, s_scale[rx.y]);
                const float fl = floorf(src);
                const int lo = max((int)fl, 0), hi = min((int)ceilf(src), rx.y - 1);
                lx = __fsub_rn(src, fl);
                xo0 = (rx.x + lo) * kPixBytes;
                xo1 = (rx.x + hi) * kPixBytes;
            } else {
                xo0 = xo1 = HWp * kPixBytes;
                lx = 0.f;
            }
            const YEntry *yt = s_ytab + rl * pool;
            float4 *dst = reinterpret_cast<float4 *>(p.out) +
                          (((size_t)b * p.R + r0 + rl) * PP + px) * C4 + (size_t)s * LANES + q;
            const int row_step = p.W * kPixBytes;
            float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
#pragma unroll 2
            for (int py = 0; py < pool; ++py) {
                const YEntry e = yt[py];
                const unsigned act = e.code >> 28;
                if (act != kActKeep) {
                    const unsigned char *row = mapb + (e.code & 0xFFFFu) * kPixBytes;
                    if (act == kActShift || act == kActShiftDup) h0 = h1;
                    if (act == kActBoth || act == kActOneDup)
                        h0 = lerp4(*reinterpret_cast<const float4 *>(row + xo0),
                                   *reinterpret_cast<const float4 *>(row + xo1), lx);
                    if (act == kActOneDup) h1 = h0;
                    if (act == kActShift || act == kActBoth || act == kActExtend)
                        h1 = lerp4(*reinterpret_cast<const float4 *>(row + row_step + xo0),
                                   *reinterpret_cast<const float4 *>(row + row_step + xo1), lx);
                }
                st_stream_f4(dst, lerp4(h0, h1, e.lerp));
                dst += py_step;
            }
        }
    }
}

// ---- pair form: a cluster of two CTAs holds one channel slice, half of the map rows each --------------------
// For maps whose 32-channel slice does not fit one CTA's shared memory (600x800 px: 38*50*128 B = 243 KB) the
// whole-map form had to fall back to 16-channel slices: 64-byte store segments, twice the CTAs (0.83 of the HBM
// peak instead of 0.95).  Here CTA `rank` of a cluster of two stages rows [rank*Hh, rank*Hh + Hh) of the slice
// (one TMA tile copy each) and the pair splits the output columns; a bilinear tap on a row of the other half is
// a distributed-shared-memory load (ld.shared::cluster on the mapa-translated address).  Store lines stay 128
// bytes, and when a half is small enough (38x38: 92 KB) two CTAs share an SM, so the staging of one cluster
// overlaps the store stream of another.
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int LANES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPoolThreads, 1)
    roi_pool_pair_kernel(RoiPoolParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    constexpr int kPixBytes = LANES * 16;
    constexpr int G = kPoolThreads / LANES;          // columns in flight per CTA
    const int Hh = p.tma_rows;                       // map rows per half
    const int HWh = Hh * p.W;                        // pixels of a half (whole TMA box)
    const int pool = p.pool, PP = pool * pool;
    float4 *s_map = reinterpret_cast<float4 *>(smem);                                   // [HWh+1][LANES]
    YEntry *s_ytab = reinterpret_cast<YEntry *>(smem + (size_t)(HWh + 1) * kPixBytes);  // [chunk][pool]
    int2 *s_roi = reinterpret_cast<int2 *>(s_ytab + (size_t)p.roi_chunk * pool);        // [chunk] {x, cw}
    float *s_scale = reinterpret_cast<float *>(s_roi + p.roi_chunk);                    // [W+1] cw / pool (float32 divide)

    const int rank = (int)cluster_ctarank();
    const int pairi = blockIdx.x >> 1;
    const int b = pairi / p.n_slices;
    const int s = pairi - b * p.n_slices;
    const int C4 = p.C >> 2;
    const int q = threadIdx.x % LANES;
    const int g = threadIdx.x / LANES;

    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, (uint32_t)((size_t)HWh * kPixBytes));
        // rows beyond the map (second half of an odd H) are zero-filled by the TMA unit
        tma_load_4d(s_map, &tmap, s * LANES * 4, 0, rank * Hh, b, &s_bar);
    }
    if (g == 0) s_map[HWh * LANES + q] = make_float4(0.f, 0.f, 0.f, 0.f);         // the "zero pixel"
    for (int i = threadIdx.x; i <= p.W; i += kPoolThreads) s_scale[i] = __fdiv_rn((float)i, (float)p.pool);
    const unsigned char *mapb = reinterpret_cast<const unsigned char *>(s_map) + q * 16;
    const uint32_t peer = cluster_map_shared(mapb, (uint32_t)(rank ^ 1));
    const size_t py_step = (size_t)pool * C4;
    bool map_ready = false;

    // a tap: 16 bytes at `off` inside the half `owner` of the slice
    auto tap = [&](int owner, unsigned off) -> float4 {
        return owner == rank ? *reinterpret_cast<const float4 *>(mapb + off) : ld_dsmem_f4(peer + off);
    };

    for (int r0 = 0; r0 < p.R; r0 += p.roi_chunk) {
        const int nr = min(p.roi_chunk, p.R - r0);
        __syncthreads();     // previous chunk done with the tables
        for (int e = threadIdx.x; e < nr * pool; e += kPoolThreads) {
            int rl = e / pool, i = e - rl * pool;
            int x, y, cw, ch;
            YEntry en;
            if (fetch_roi(p, b, r0 + rl, x, y, cw, ch)) {
                int lo, hi, plo = -1, phi = -1;
                float pl;
                legacy_axis(i, ch, pool, lo, hi, en.lerp);
                if (i > 0) legacy_axis(i - 1, ch, pool, plo, phi, pl);
                unsigned act;
                if (i == 0) act = (hi == lo) ? kActOneDup : kActBoth;
                else if (lo == plo && hi == phi) act = kActKeep;
                else if (lo == phi && hi != lo && phi != plo) act = kActShift;
                else if (lo == phi && hi == lo && phi != plo) act = kActShiftDup;
                else if (lo == plo && phi == plo && hi != lo) act = kActExtend;
                else act = (hi == lo) ? kActOneDup : kActBoth;
                // absolute source row y + lo; the kernel splits it into (half, row inside the half)
                en.code = (unsigned)(y + lo) | ((hi != lo) ? 0x10000u : 0u) | (act << 28);
                if (i == 0) s_roi[rl] = make_int2(x, cw);
            } else {
                en.code = (i == 0) ? (kActOneDup << 28) : (kActKeep << 28);
                en.lerp = 0.f;
                if (i == 0) s_roi[rl] = make_int2(0, 0);
            }
            s_ytab[e] = en;
        }
        __syncthreads();
        if (!map_ready) {
            mbar_wait(&s_bar, 0);        // my half has landed ...
            cluster_sync_all();          // ... and so has the partner's
            map_ready = true;
        }

        const int ncol = nr * pool;
        // the pair walks the columns 2G at a time: CTA `rank` takes the rank-th group of G
        const int g2 = g + rank * G;
        int rl = g2 / pool, px = g2 - rl * pool;
        const int d_rl = (2 * G) / pool, d_px = 2 * G - d_rl * pool;
        for (int col = g2; col < ncol; col += 2 * G, rl += d_rl, px += d_px) {
            while (px >= pool) { px -= pool; ++rl; }
            const int2 rx = s_roi[rl];
            const bool live = rx.y > 0;
            unsigned xo0 = 0, xo1 = 0;
            float lx = 0.f;
            if (live) {
                const float src = __fmul_rn((float)px, s_scale[rx.y]);
                const float fl = floorf(src);
                const int lo = max((int)fl, 0), hi = min((int)ceilf(src), rx.y - 1);
                lx = __fsub_rn(src, fl);
                xo0 = (unsigned)(rx.x + lo) * kPixBytes;
                xo1 = (unsigned)(rx.x + hi) * kPixBytes;
            }
            const YEntry *yt = s_ytab + rl * pool;
            float4 *dst = reinterpret_cast<float4 *>(p.out) +
                          (((size_t)b * p.R + r0 + rl) * PP + px) * C4 + (size_t)s * LANES + q;
            const unsigned row_bytes = (unsigned)p.W * kPixBytes;
            float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
#pragma unroll 2
            for (int py = 0; py < pool; ++py) {
                const YEntry e = yt[py];
                const unsigned act = e.code >> 28;
                if (act != kActKeep && live) {
                    const int ya = (int)(e.code & 0xFFFFu), yb = ya + 1;           // absolute rows y0 and y0 + 1
                    const int oa = ya >= Hh, ob = yb >= Hh;
                    const unsigned ra = (unsigned)(ya - oa * Hh) * row_bytes, rb = (unsigned)(yb - ob * Hh) * row_bytes;
                    if (act == kActShift || act == kActShiftDup) h0 = h1;
                    if (act == kActBoth || act == kActOneDup) h0 = lerp4(tap(oa, ra + xo0), tap(oa, ra + xo1), lx);
                    if (act == kActOneDup) h1 = h0;
                    if (act == kActShift || act == kActBoth || act == kActExtend)
                        h1 = lerp4(tap(ob, rb + xo0), tap(ob, rb + xo1), lx);
                }
                st_stream_f4(dst, lerp4(h0, h1, e.lerp));
                dst += py_step;
            }
        }
    }
    if (!map_ready) {                    // no chunk at all (R == 0 cannot happen, but never leave the partner waiting)
        mbar_wait(&s_bar, 0);
        cluster_sync_all();
    }
    cluster_sync_all();                  // the partner may still be reading this half
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
static int launch_pair(const RoiPoolParams &p, const CUtensorMap &tmap, int B, size_t smem, cudaStream_t st) {
    int dev = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(roi_pool_pair_kernel<LANES>), dev, smem)) return rc;
    roi_pool_pair_kernel<LANES><<<2 * B * p.n_slices, kPoolThreads, smem, st>>>(p, tmap);
    return check_launch("roi_pool_pair_kernel");
}

template <int LANES>
static int launch_slice(const RoiPoolParams &p, const CUtensorMap &tmap, int B, size_t smem, cudaStream_t st) {
    int dev = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(roi_pool_slice_kernel<LANES>), dev, smem)) return rc;
    roi_pool_slice_kernel<LANES><<<B * p.n_slices, kPoolThreads, smem, st>>>(p, tmap);
    return check_launch("roi_pool_slice_kernel");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// the feature maps as a rank-4 tensor (C, W, H, B), box = (lanes*4 channels, W, rows, 1); false when the shape
// does not fit a TMA box (the kernel then stages with cp.async)
static bool make_map_tensor(CUtensorMap *tmap, const float *feat, int B, int H, int W, int C, int lanes, int rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || W > 256 || rows > 256 || rows < 1 || (C * 4) % 16 != 0) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {(cuuint32_t)lanes * 4, (cuuint32_t)W, (cuuint32_t)rows, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(feat), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// rows per TMA box: whole map when it has at most 256 rows, else the split with the least padding
static int tma_box_rows(int H) {
    if (H <= 256) return H;
    int best = 1, best_pad = 1 << 30;
    for (int k = 256; k >= 64; --k) {
        const int pad = (H + k - 1) / k * k - H;
        if (pad < best_pad) { best_pad = pad; best = k; }
    }
    return best;
}

}  // namespace radnet

using namespace radnet;

// option roipool_force_direct = 1 forces the direct kernel (parity tests exercise both)
static bool force_direct() { return get_option(kOptRoipoolForceDirect) == 1; }

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

    int dev = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    const int smem_limit = device_smem_optin(dev);
    if (smem_limit < 0) return RADNET_E_CUDA;
    const size_t HW = (size_t)H * W;
    const size_t per_roi = (size_t)pool * sizeof(YEntry) + sizeof(int2);
    const long long form = get_option(kOptRoipoolForm);          // 0 auto, 1 whole-map slices, 2 cluster pairs
    if (C % 4 == 0 && !force_direct() && HW < 65536 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const int C4 = C / 4;
        // pair form (32-channel slices, half the rows per CTA of a cluster of two): when the whole-map form cannot
        // keep 32 channels in one CTA, or on request
        const int Hh = (H + 1) / 2;
        const size_t half_bytes = ((size_t)Hh * W + 1) * 8 * 16;
        const size_t whole8 = (HW + 1) * 8 * 16 + ((size_t)W + 1) * sizeof(float) + 16 + 8 * per_roi;
        const bool pair_fits = C4 % 8 == 0 && W <= 256 && Hh <= 256 && H >= 2 && encode_tiled_fn() &&
                               half_bytes + ((size_t)W + 1) * sizeof(float) + 16 + 8 * per_roi <= (size_t)smem_limit &&
                               2LL * B * (C4 / 8) < 0x7fffffffLL;
        if (pair_fits && (form == 2 || (form == 0 && whole8 > (size_t)smem_limit))) {
            const size_t scale_bytes = ((size_t)W + 1) * sizeof(float) + 16;
            // two CTAs per SM when a half is small enough: cap the shared memory at half an SM
            size_t budget = (size_t)smem_limit;
            if (2 * (half_bytes + scale_bytes + 32 * per_roi + 1024) <= (size_t)smem_limit + 1024)
                budget = ((size_t)smem_limit + 1024) / 2 - 1024;
            size_t chunk = (budget - half_bytes - scale_bytes) / per_roi;
            if (chunk > (size_t)rois_per_panel) chunk = rois_per_panel;
            alignas(64) CUtensorMap tmap;
            memset(&tmap, 0, sizeof(tmap));
            if (chunk >= 1 && make_map_tensor(&tmap, feat, B, H, W, C, 8, Hh)) {
                p.n_slices = C4 / 8;
                p.roi_chunk = (int)chunk;
                p.map_rows_pad = Hh;
                p.tma_rows = Hh;
                return launch_pair<8>(p, tmap, B, half_bytes + chunk * per_roi + scale_bytes, st);
            }
        }
        const int lanes_opts[4] = {8, 4, 2, 1};
        const int box_rows = (W <= 256 && encode_tiled_fn()) ? tma_box_rows(H) : 0;
        const int rows_pad = box_rows ? (H + box_rows - 1) / box_rows * box_rows : H;
        for (int li = 0; li < 4; ++li) {
            int L = lanes_opts[li];
            size_t map_bytes = ((size_t)rows_pad * W + 1) * L * 16;
            const size_t scale_bytes = ((size_t)W + 1) * sizeof(float) + 16;
            if (C4 % L != 0 || map_bytes + scale_bytes + 8 * per_roi > (size_t)smem_limit) continue;
            size_t chunk = ((size_t)smem_limit - map_bytes - scale_bytes) / per_roi;
            if (chunk > (size_t)rois_per_panel) chunk = rois_per_panel;
            size_t smem = map_bytes + chunk * per_roi + scale_bytes;
            p.n_slices = C4 / L;
            p.roi_chunk = (int)chunk;
            if ((long long)B * p.n_slices >= 0x7fffffffLL) continue;
            alignas(64) CUtensorMap tmap;
            memset(&tmap, 0, sizeof(tmap));
            p.map_rows_pad = rows_pad;
            p.tma_rows = (box_rows && make_map_tensor(&tmap, feat, B, H, W, C, L, box_rows)) ? box_rows : 0;
            if (!p.tma_rows && rows_pad != H) continue;            // sized for TMA boxes but no descriptor: next option
            switch (L) {
                case 8: return launch_slice<8>(p, tmap, B, smem, st);
                case 4: return launch_slice<4>(p, tmap, B, smem, st);
                case 2: return launch_slice<2>(p, tmap, B, smem, st);
                default: return launch_slice<1>(p, tmap, B, smem, st);
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

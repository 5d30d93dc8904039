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
//     multiply and add (never fused) - bit-identical to the CPU kernel of TF-1.  The multiply is
//     issued as a packed FMA with a -0.0 addend so that ptxas cannot contract it (see lerp2).
//   * the kernel is co-limited by instruction issue (one SM sustains ~46 GB/s of 16-byte stores)
//     and by how HBM takes the write stream; the inner loop costs ~30 instructions per 16-byte
//     store (flag-coded y table, cached rows), and the launcher tunes how tightly the warps of a
//     CTA / the CTAs of neighbouring slices are kept in lockstep (see "automatic lockstep").
//
// A direct (global-load) kernel covers maps whose slice does not fit in shared memory
// and channel counts that are not a multiple of 4.
#include <cuda.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace radnet {

#ifndef RADNET_POOL_THREADS
#define RADNET_POOL_THREADS 512
#endif
constexpr int kSliceThreads = RADNET_POOL_THREADS;   // whole-map form (one CTA per SM)
constexpr int kPoolThreads = 512;                    // band form (two CTAs per SM)
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
    int band_rows, n_bands;  // band form: source rows owned by a band, bands per map
    int grid;                // whole-map form: CTAs launched
    int n_work;              // whole-map form: (panel, slice) work items; the grid may be smaller (persistent CTAs)
    int cluster;             // whole-map form: CTAs per cluster (neighbouring slices of a panel), 1 = no cluster
    int sync_every;          // whole-map form: CTA (or cluster) barrier every this many column rounds (0 = never)
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
//   bits  0..23  byte offset, inside the staged slice, of the first cell of source row y+y0
//   bits 27..31  what the two cached source rows have to do before this output row (kRow*)
// A column keeps h0 = hrow(y0) and h1 = hrow(y1) in registers (hrow(r) = tl + (tr-tl)*xlerp of source row r).
struct YEntry {
    unsigned code;
    float lerp;
};
enum : unsigned {
    kRowAny = 1u << 31,      // something changes (sign bit: one compare)
    kRowShift = 1u << 30,    // h0 <- h1           (y0 = previous y1)
    kRowS0 = 1u << 29,       // h0 <- hrow(y0)     sampled
    kRowDup = 1u << 28,      // h1 <- h0           (y1 = y0)
    kRowS1 = 1u << 27,       // h1 <- hrow(y0 + 1) sampled
    kRowOffMask = (1u << 24) - 1u
};

// the update the row cache needs going from source rows (plo, phi) to (lo, hi); first = no previous row
__device__ __forceinline__ unsigned row_action(bool first, int lo, int hi, int plo, int phi) {
    if (first) return kRowAny | kRowS0 | (hi == lo ? kRowDup : kRowS1);
    if (lo == plo && hi == phi) return 0u;                                   // same two rows
    if (lo == phi && phi != plo) return kRowAny | kRowShift | (hi != lo ? kRowS1 : 0u);
    if (lo == plo && phi == plo && hi != lo) return kRowAny | kRowS1;        // rows were equal, new y1
    return kRowAny | kRowS0 | (hi == lo ? kRowDup : kRowS1);
}

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

// a + (b-a)*t on two lanes at once with three packed instructions and NO contraction of the multiply into the
// add.  ptxas 12.9 contracts a packed multiply feeding a packed add into FFMA2 even for explicitly rounded PTX,
// which would break bit-parity with TF's unfused arithmetic.  So the product is itself issued as a packed FMA
// with an addend of -0.0 (d*t + (-0.0) rounds exactly like d*t, signed zeros included; the -0.0 comes from a
// kernel parameter so that ptxas cannot fold it away), and an FMA result feeding an add cannot be fused again:
// FADD2 (b-a), FFMA2 (d*t - 0), FADD2 (a + m).  tests/test_build_sass.py checks that the kernels contain no
// FMUL2 and exactly two FADD2 per FFMA2 (a contraction would trade an FADD2 for an FFMA2); the parity tests check the bits.
__device__ __forceinline__ void lerp2(float a0, float a1, float b0, float b1, unsigned long long tt,
                                      unsigned long long nz, float &r0, float &r1) {
    unsigned long long a, b, d, m, r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(b), "l"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(m) : "l"(d), "l"(tt), "l"(nz));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(m));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}
__device__ __forceinline__ float4 lerp4(const float4 &a, const float4 &b, float t, unsigned long long nz) {
    unsigned long long tt;
    asm("mov.b64 %0, {%1,%1};" : "=l"(tt) : "f"(t));
    float4 r;
    lerp2(a.x, a.y, b.x, b.y, tt, nz, r.x, r.y);
    lerp2(a.z, a.w, b.z, b.w, tt, nz, r.z, r.w);
    return r;
}
// {-0.0f, -0.0f} that the compiler cannot constant-fold (pool is a kernel parameter, always >= 1)
__device__ __forceinline__ unsigned long long opaque_neg_zero2(int pool) {
    const unsigned w = 0x80000000u | ((unsigned)pool >> 31);
    unsigned long long nz;
    asm("mov.b64 %0, {%1,%1};" : "=l"(nz) : "r"(w));
    return nz;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// One output column (roi, px), rows [py0, py1): a group of LANES threads (16 bytes each) walks py.  The
// horizontal interpolation of a source row, hrow(r) = tl + (tr-tl)*xlerp, does not depend on py, so the two rows
// an output needs are cached in registers and reused while y0/y1 repeat (always, when the RoI is upsampled:
// h < pool); only rows that change are re-sampled (2 LDS.128 each).  top/bottom of TF's formula are exactly
// hrow(y0)/hrow(y1), so the result is bit-identical to evaluating all four taps per output.
// mapb = this lane's 16 bytes of pixel 0 of the staged slice; xo0/xo1 = byte offsets of the two x taps.
#ifndef RADNET_POOL_STORE
#define RADNET_POOL_STORE 0
#endif
__device__ __forceinline__ void st_pool_f4(float4 *p, const float4 &v) {
#if RADNET_POOL_STORE == 0
    st_stream_f4(p, v);
#elif RADNET_POOL_STORE == 1
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#elif RADNET_POOL_STORE == 2
    asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#else
    asm volatile("st.global.L1::no_allocate.L2::evict_last.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#endif
}

template <int POOL>
__device__ __forceinline__ void pool_column(const unsigned char *mapb, const YEntry *yt, int py0, int py1,
                                            unsigned xo0, unsigned xo1, float lx, unsigned row_step, float4 *dst,
                                            size_t py_step, unsigned long long nz) {
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
    const unsigned char *t0 = mapb + xo0, *t1 = mapb + xo1;         // the two x taps in source row 0
    auto body = [&](int py) {
        const YEntry e = yt[py];
        if ((int)e.code < 0) {
            const unsigned off = e.code & kRowOffMask;
            if (e.code & kRowS0)
                h0 = lerp4(*reinterpret_cast<const float4 *>(t0 + off), *reinterpret_cast<const float4 *>(t1 + off),
                           lx, nz);
            else if (e.code & kRowShift)
                h0 = h1;
            if (e.code & kRowS1)
                h1 = lerp4(*reinterpret_cast<const float4 *>(t0 + off + row_step),
                           *reinterpret_cast<const float4 *>(t1 + off + row_step), lx, nz);
            else if (e.code & kRowDup)
                h1 = h0;
        }
        st_pool_f4(dst, lerp4(h0, h1, e.lerp, nz));
        dst += py_step;
    };
    if (POOL > 0) {
#pragma unroll
        for (int py = 0; py < POOL; ++py) body(py);
    } else {
#pragma unroll 2
        for (int py = py0; py < py1; ++py) body(py);
    }
}

// Two neighbouring output columns (roi, px) and (roi, px + 1) at once: they share the RoI's y table, so the table
// entry, its decoding and the branches on it are paid once per pair of 16-byte stores, and the two columns' taps
// and lerps are independent work the scheduler can overlap.  has_b = false: the pair has only its first column
// (odd pool sizes); the caller makes column B shadow column A and its stores are skipped.
template <int POOL>
__device__ __forceinline__ void pool_column_pair(const unsigned char *mapb, const YEntry *yt, int npy, unsigned xa0,
                                                 unsigned xa1, float lxa, unsigned xb0, unsigned xb1, float lxb,
                                                 bool has_b, unsigned row_step, float4 *dst, size_t px_step,
                                                 size_t py_step, unsigned long long nz) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
    const unsigned char *ta0 = mapb + xa0, *ta1 = mapb + xa1, *tb0 = mapb + xb0, *tb1 = mapb + xb1;
    auto ld = [](const unsigned char *p) { return *reinterpret_cast<const float4 *>(p); };
    auto body = [&](int py) {
        const YEntry e = yt[py];
        if ((int)e.code < 0) {
            const unsigned off = e.code & kRowOffMask;
            if (e.code & kRowS0) {
                a0 = lerp4(ld(ta0 + off), ld(ta1 + off), lxa, nz);
                b0 = lerp4(ld(tb0 + off), ld(tb1 + off), lxb, nz);
            } else if (e.code & kRowShift) {
                a0 = a1;
                b0 = b1;
            }
            if (e.code & kRowS1) {
                a1 = lerp4(ld(ta0 + off + row_step), ld(ta1 + off + row_step), lxa, nz);
                b1 = lerp4(ld(tb0 + off + row_step), ld(tb1 + off + row_step), lxb, nz);
            } else if (e.code & kRowDup) {
                a1 = a0;
                b1 = b0;
            }
        }
        st_pool_f4(dst, lerp4(a0, a1, e.lerp, nz));
        if (has_b) st_pool_f4(dst + px_step, lerp4(b0, b1, e.lerp, nz));
        dst += py_step;
    };
    if (POOL > 0) {
#pragma unroll
        for (int py = 0; py < POOL; ++py) body(py);
    } else {
#pragma unroll 2
        for (int py = 0; py < npy; ++py) body(py);
    }
}

// x taps of output column px of a RoI that is cw cells wide and starts at cell x (TF-1 legacy weights with the
// float32 scale cw/pool taken from a table); cw == 0 marks an empty slot: both taps on the zero pixel
__device__ __forceinline__ void x_taps(int px, int x, int cw, const float *s_scale, unsigned pix_bytes,
                                       unsigned zero_off, unsigned &xo0, unsigned &xo1, float &lx) {
    if (cw > 0) {
        const float src = __fmul_rn((float)px, s_scale[cw]);
        const float fl = floorf(src);
        const int lo = max((int)fl, 0), hi = min((int)ceilf(src), cw - 1);
        lx = __fsub_rn(src, fl);
        xo0 = (unsigned)(x + lo) * pix_bytes;
        xo1 = (unsigned)(x + hi) * pix_bytes;
    } else {
        xo0 = xo1 = zero_off;
        lx = 0.f;
    }
}

// LANES = float4 lanes per pixel in the slice (8 -> 32 channels, 4 -> 16, 2 -> 8, 1 -> 4); POOL = compile-time
// pool size (fully unrolled rows) or 0 for any pool size.
template <int LANES, int POOL>
__global__ void __launch_bounds__(kSliceThreads, 1) roi_pool_slice_kernel(RoiPoolParams p,
                                                                          const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    constexpr int kPixBytes = LANES * 16;
    constexpr int G = kSliceThreads / LANES;          // columns in flight per CTA
    const int HW = p.H * p.W;
    const int HWp = p.map_rows_pad * p.W;            // pixels held in shared memory (whole TMA boxes)
    const int pool = POOL > 0 ? POOL : p.pool, PP = pool * pool;
    float4 *s_map = reinterpret_cast<float4 *>(smem);                                   // [HWp+1][LANES]
    YEntry *s_ytab = reinterpret_cast<YEntry *>(smem + (size_t)(HWp + 1) * kPixBytes);  // [chunk][pool]
    int2 *s_roi = reinterpret_cast<int2 *>(s_ytab + (size_t)p.roi_chunk * pool);        // [chunk] {x, cw}
    float *s_scale = reinterpret_cast<float *>(s_roi + p.roi_chunk);                    // [W+1] cw / pool (float32 divide)

    const int C4 = p.C >> 2;
    const int q = threadIdx.x % LANES;
    const int g = threadIdx.x / LANES;
    const unsigned long long nz = opaque_neg_zero2(p.pool);
    if (threadIdx.x == 0 && p.tma_rows > 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i <= p.W; i += kSliceThreads) s_scale[i] = __fdiv_rn((float)i, (float)p.pool);
    if (g == 0) s_map[HWp * LANES + q] = make_float4(0.f, 0.f, 0.f, 0.f);         // the "zero pixel"
    const unsigned char *mapb = reinterpret_cast<const unsigned char *>(s_map) + q * 16;
    const size_t py_step = (size_t)pool * C4;
    const unsigned row_step = (unsigned)p.W * kPixBytes;

    // a CTA takes (panel, slice) work items round robin: with as many CTAs as items this is one item each; with
    // fewer (a multiple of the slices per panel) every CTA keeps its slice and the slices stay evenly loaded
    uint32_t phase = 0;
    for (int work = blockIdx.x; work < p.n_work; work += gridDim.x, phase ^= 1u) {
    const int b = work / p.n_slices;
    const int s = work - b * p.n_slices;
    if (work != (int)blockIdx.x) __syncthreads();          // everybody is done with the previous map

    // ---- stage the channel slice of the whole map (read from HBM exactly once) --------
    if (p.tma_rows > 0) {
        // one elected thread: a TMA tile copy per box of `tma_rows` map rows (one box for maps up to 256 rows),
        // all completing on the same mbarrier; everybody else goes straight on to the first y tables
        if (threadIdx.x == 0) {
            const int n_box = p.map_rows_pad / p.tma_rows;
            mbar_expect_tx(&s_bar, (uint32_t)((size_t)HWp * kPixBytes));
            for (int k = 0; k < n_box; ++k)
                tma_load_4d(reinterpret_cast<unsigned char *>(s_map) + (size_t)k * p.tma_rows * p.W * kPixBytes, &tmap,
                            s * LANES * 4, 0, k * p.tma_rows, b, &s_bar);
        }
    } else {
        const float4 *src = reinterpret_cast<const float4 *>(p.feat) + (size_t)b * HW * C4 + (size_t)s * LANES + q;
        for (int pix = g; pix < HW; pix += G) cp_async16(&s_map[pix * LANES + q], src + (size_t)pix * C4);
        cp_async_wait_all();
    }
    bool map_ready = p.tma_rows == 0;

    for (int r0 = 0; r0 < p.R; r0 += p.roi_chunk) {
        const int nr = min(p.roi_chunk, p.R - r0);
        __syncthreads();     // previous chunk done with the tables (and the map has landed)
        for (int e = threadIdx.x; e < nr * pool; e += kSliceThreads) {
            int rl = e / pool, i = e - rl * pool;
            int x, y, cw, ch;
            YEntry en;
            if (fetch_roi(p, b, r0 + rl, x, y, cw, ch)) {
                int lo, hi, plo = -1, phi = -1;
                float pl;
                legacy_axis(i, ch, pool, lo, hi, en.lerp);
                if (i > 0) legacy_axis(i - 1, ch, pool, plo, phi, pl);
                en.code = (unsigned)((y + lo) * p.W) * kPixBytes | row_action(i == 0, lo, hi, plo, phi);
                if (i == 0) s_roi[rl] = make_int2(x, cw);
            } else {
                // rows add nothing and the column points at the zero pixel: output is exactly 0
                en.code = (i == 0) ? (kRowAny | kRowS0 | kRowDup) : 0u;
                en.lerp = 0.f;
                if (i == 0) s_roi[rl] = make_int2(0, 0);
            }
            s_ytab[e] = en;
        }
        __syncthreads();
        if (!map_ready) {            // the barrier above also ordered the mbarrier's initialisation before this wait
            mbar_wait(&s_bar, phase);
            map_ready = true;
        }

        if (LANES == 8) {
        // 32-channel slices: a group takes PAIRS of neighbouring columns (px, px + 1) - measured 0.87 -> 0.92 of the HBM
        // peak on 7x7x512, 0.956 -> 0.963 sustained on 14x14x1024; on 16-channel slices (38x50 maps) pairs measured
        // 0.89 against 0.95 for single columns, so narrower slices keep one column per group
        // ppair pairs per RoI, the last one single when pool is odd; (rl, k) advance incrementally by G pairs: no
        // integer division in the loop
        const int ppair = (pool + 1) >> 1;
        const int npair = nr * ppair;
        int rl = g / ppair, k = g - rl * ppair;
        const int d_rl = G / ppair, d_k = G - d_rl * ppair;
        int it = 0;
        for (int base = 0; base < npair; base += G, rl += d_rl, k += d_k, ++it) {
            // lockstep: the warps of the CTA (and, in a cluster, the CTAs of neighbouring channel slices of the same
            // panel) are kept within `sync_every` rounds of each other, so that what the SM writes at any moment stays
            // within a few RoIs - see the launcher for the measured effect
            if (p.sync_every > 0 && it % p.sync_every == 0) {
                if (p.cluster > 1) cluster_sync_relaxed();
                else __syncthreads();
            }
            if (k >= ppair) { k -= ppair; ++rl; }
            if (base + g >= npair) continue;
            const int2 rx = s_roi[rl];
            const int px = 2 * k;
            const bool has_b = px + 1 < pool;
            unsigned xa0, xa1, xb0, xb1;
            float lxa, lxb;
            x_taps(px, rx.x, rx.y, s_scale, kPixBytes, (unsigned)HWp * kPixBytes, xa0, xa1, lxa);
            if (has_b) {
                x_taps(px + 1, rx.x, rx.y, s_scale, kPixBytes, (unsigned)HWp * kPixBytes, xb0, xb1, lxb);
            } else {               // single column: B shadows A (valid addresses for every row offset, nothing stored)
                xb0 = xa0; xb1 = xa1; lxb = lxa;
            }
            float4 *dst = reinterpret_cast<float4 *>(p.out) +
                          (((size_t)b * p.R + r0 + rl) * PP + px) * C4 + (size_t)s * LANES + q;
            pool_column_pair<POOL>(mapb, s_ytab + rl * pool, pool, xa0, xa1, lxa, xb0, xb1, lxb, has_b, row_step, dst,
                                   (size_t)C4, py_step, nz);
        }
        } else {
        const int ncol = nr * pool;
        int rl = g / pool, px = g - rl * pool;
        const int d_rl = G / pool, d_px = G - d_rl * pool;
        int it = 0;
        for (int base = 0; base < ncol; base += G, rl += d_rl, px += d_px, ++it) {
            if (p.sync_every > 0 && it % p.sync_every == 0) {
                if (p.cluster > 1) cluster_sync_relaxed();
                else __syncthreads();
            }
            if (px >= pool) { px -= pool; ++rl; }
            if (base + g >= ncol) continue;
            const int2 rx = s_roi[rl];
            unsigned xo0, xo1;
            float lx;
            x_taps(px, rx.x, rx.y, s_scale, kPixBytes, (unsigned)HWp * kPixBytes, xo0, xo1, lx);
            float4 *dst = reinterpret_cast<float4 *>(p.out) +
                          (((size_t)b * p.R + r0 + rl) * PP + px) * C4 + (size_t)s * LANES + q;
            pool_column<POOL>(mapb, s_ytab + rl * pool, 0, pool, xo0, xo1, lx, row_step, dst, py_step, nz);
        }
        }
    }
    }   // work items
}

// ---- band form: a CTA holds a band of map rows and emits the output rows that sample it -----------------------
// For maps whose 32-channel slice does not fit one CTA's shared memory (600x800 px: 38*50*128 B = 243 KB) the
// whole-map form has to fall back to 16-channel slices: 64-byte store segments, twice the CTAs (0.84 of the HBM
// peak instead of 0.96).  (A cluster of two CTAs holding half of the rows each, with the taps on the other half
// going through distributed shared memory, was measured at 0.53-0.55: remote LDS.128 taps are too slow.)
// Here the map rows are cut into `n_bands` bands of `band_rows` rows.  Output row py of a RoI samples source rows
// ya = y + lo(py) and ya or ya + 1, so it belongs to exactly one band, ya / band_rows, and the CTA of that band
// stages band_rows + 1 rows (one TMA tile copy; rows below the map are zero-filled by the TMA unit and never
// sampled).  Every CTA of a (panel, slice) walks all RoIs but only the py range that falls into its band (ya is
// non-decreasing in py, so the range is contiguous; RoIs without a row in the band are dropped from the column
// list).  No data crosses CTAs, store lines stay 128 bytes, the map is read (band_rows+1)/band_rows times, and
// two CTAs share an SM, so the staging of one overlaps the store stream of the other.
template <int LANES>
__global__ void __launch_bounds__(kPoolThreads, 2) roi_pool_band_kernel(RoiPoolParams p,
                                                                         const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_nlive;
    constexpr int kPixBytes = LANES * 16;
    constexpr int G = kPoolThreads / LANES;          // columns in flight per CTA
    const int Hb = p.band_rows;
    const int HWb = p.tma_rows * p.W;                // pixels staged (band_rows + 1 rows = one TMA box)
    const int pool = p.pool, PP = pool * pool;
    float4 *s_map = reinterpret_cast<float4 *>(smem);                                   // [HWb+1][LANES]
    int4 *s_roi = reinterpret_cast<int4 *>(smem + (size_t)(HWb + 1) * kPixBytes);       // [chunk] {x, cw, py_lo, py_hi}
    YEntry *s_ytab = reinterpret_cast<YEntry *>(s_roi + p.roi_chunk);                   // [chunk][pool]
    int *s_live = reinterpret_cast<int *>(s_ytab + (size_t)p.roi_chunk * pool);         // [chunk] RoIs with rows here
    float *s_scale = reinterpret_cast<float *>(s_live + p.roi_chunk);                   // [W+1] cw / pool (float32 divide)

    const int band = blockIdx.x % p.n_bands;
    const int bs = blockIdx.x / p.n_bands;
    const int b = bs / p.n_slices;
    const int s = bs - b * p.n_slices;
    const int C4 = p.C >> 2;
    const int q = threadIdx.x % LANES;
    const int g = threadIdx.x / LANES;
    const int row_base = band * Hb;
    const unsigned long long nz = opaque_neg_zero2(p.pool);

    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, (uint32_t)((size_t)HWb * kPixBytes));
        tma_load_4d(s_map, &tmap, s * LANES * 4, 0, row_base, b, &s_bar);
    }
    if (g == 0) s_map[HWb * LANES + q] = make_float4(0.f, 0.f, 0.f, 0.f);         // the "zero pixel"
    for (int i = threadIdx.x; i <= p.W; i += kPoolThreads) s_scale[i] = __fdiv_rn((float)i, (float)p.pool);
    const unsigned char *mapb = reinterpret_cast<const unsigned char *>(s_map) + q * 16;
    const size_t py_step = (size_t)pool * C4;
    const unsigned row_step = (unsigned)p.W * kPixBytes;
    bool map_ready = false;

    for (int r0 = 0; r0 < p.R; r0 += p.roi_chunk) {
        const int nr = min(p.roi_chunk, p.R - r0);
        __syncthreads();     // previous chunk done with the tables
        for (int rl = threadIdx.x; rl < nr; rl += kPoolThreads) s_roi[rl] = make_int4(0, 0, 0, 0);
        __syncthreads();
        for (int e = threadIdx.x; e < nr * pool; e += kPoolThreads) {
            int rl = e / pool, i = e - rl * pool;
            int x, y, cw, ch;
            YEntry en;
            en.code = 0u;
            en.lerp = 0.f;
            if (fetch_roi(p, b, r0 + rl, x, y, cw, ch)) {
                int lo, hi, plo = -1, phi = -1, nlo = -1, nhi;
                float pl;
                legacy_axis(i, ch, pool, lo, hi, en.lerp);
                if (i > 0) legacy_axis(i - 1, ch, pool, plo, phi, pl);
                if (i + 1 < pool) legacy_axis(i + 1, ch, pool, nlo, nhi, pl);
                const int ya = y + lo;
                if (ya / Hb == band) {
                    const bool first = i == 0 || (y + plo) / Hb != band;
                    const bool last = i + 1 == pool || (y + nlo) / Hb != band;
                    en.code = (unsigned)((ya - row_base) * p.W) * kPixBytes | row_action(first, lo, hi, plo, phi);
                    if (first) { s_roi[rl].x = x; s_roi[rl].y = cw; s_roi[rl].z = i; }
                    if (last) s_roi[rl].w = i + 1;
                }
            } else if (band == 0) {
                // empty slot: band 0 writes its zeros (rows add nothing and the column points at the zero pixel)
                if (i == 0) { en.code = kRowAny | kRowS0 | kRowDup; s_roi[rl].z = 0; }
                if (i + 1 == pool) s_roi[rl].w = pool;
            }
            s_ytab[e] = en;
        }
        __syncthreads();
        if (threadIdx.x < 32) {          // ordered list of the RoIs of this chunk that have rows in this band
            int n = 0;
            for (int base = 0; base < nr; base += 32) {
                const int rl = base + (int)threadIdx.x;
                const bool live = rl < nr && s_roi[rl].w > s_roi[rl].z;
                const unsigned m = __ballot_sync(0xffffffffu, live);
                if (live) s_live[n + __popc(m & ((1u << threadIdx.x) - 1u))] = rl;
                n += __popc(m);
            }
            if (threadIdx.x == 0) s_nlive = n;
        }
        __syncthreads();
        if (!map_ready) {
            mbar_wait(&s_bar, 0);
            map_ready = true;
        }

        const int ncol = s_nlive * pool;
        int li = g / pool, px = g - li * pool;
        const int d_li = G / pool, d_px = G - d_li * pool;
        for (int col = g; col < ncol; col += G, li += d_li, px += d_px) {
            if (px >= pool) { px -= pool; ++li; }
            const int rl = s_live[li];
            const int4 rx = s_roi[rl];
            unsigned xo0, xo1;
            float lx;
            x_taps(px, rx.x, rx.y, s_scale, kPixBytes, (unsigned)HWb * kPixBytes, xo0, xo1, lx);
            float4 *dst = reinterpret_cast<float4 *>(p.out) +
                          (((size_t)b * p.R + r0 + rl) * PP + (size_t)rx.z * pool + px) * C4 + (size_t)s * LANES + q;
            pool_column<0>(mapb, s_ytab + rl * pool, rx.z, rx.w, xo0, xo1, lx, row_step, dst, py_step, nz);
        }
    }
    if (!map_ready) mbar_wait(&s_bar, 0);      // never exit with the tile copy in flight
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
    const unsigned long long nz = opaque_neg_zero2(p.pool);
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
                         lerp4(lerp4(tl, tr, lx, nz), lerp4(bl, br, lx, nz), ly, nz));
        } else {
            float tl = __ldg(r0 + (size_t)x0 * p.C + c), tr = __ldg(r0 + (size_t)x1 * p.C + c);
            float bl = __ldg(r1 + (size_t)x0 * p.C + c), br = __ldg(r1 + (size_t)x1 * p.C + c);
            orow[(size_t)px * p.C + c] = lerp1(lerp1(tl, tr, lx), lerp1(bl, br, lx), ly);
        }
    }
}

template <int LANES>
static int launch_band(const RoiPoolParams &p, const CUtensorMap &tmap, int B, size_t smem, cudaStream_t st) {
    int dev = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(roi_pool_band_kernel<LANES>), dev, smem)) return rc;
    roi_pool_band_kernel<LANES><<<B * p.n_slices * p.n_bands, kPoolThreads, smem, st>>>(p, tmap);
    return check_launch("roi_pool_band_kernel");
}

template <int LANES, int POOL>
static int launch_slice_pool(const RoiPoolParams &p, const CUtensorMap &tmap, int B, size_t smem, cudaStream_t st) {
    int dev = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(roi_pool_slice_kernel<LANES, POOL>), dev, smem))
        return rc;
    if (p.cluster > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)p.grid);
        cfg.blockDim = dim3(kSliceThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, roi_pool_slice_kernel<LANES, POOL>, p, tmap) == cudaSuccess)
            return check_launch("roi_pool_slice_kernel (cluster)");
        cudaGetLastError();            // a refused cluster launch (MPS, partitioned GPU): same barriers inside each CTA only
    }
    RoiPoolParams q = p;
    q.cluster = 1;
    roi_pool_slice_kernel<LANES, POOL><<<q.grid, kSliceThreads, smem, st>>>(q, tmap);
    return check_launch("roi_pool_slice_kernel");
}
// the reference's two pool sizes (14: ResNet-50, 7: VGG-16) have their rows fully unrolled
template <int LANES>
static int launch_slice(const RoiPoolParams &p, const CUtensorMap &tmap, int B, size_t smem, cudaStream_t st) {
#ifndef RADNET_POOL_NO_UNROLL
    if (p.pool == 14) return launch_slice_pool<LANES, 14>(p, tmap, B, smem, st);
    if (p.pool == 7) return launch_slice_pool<LANES, 7>(p, tmap, B, smem, st);
#endif
    return launch_slice_pool<LANES, 0>(p, tmap, B, smem, st);
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


// ---- automatic lockstep of the whole-map form -------------------------------------------------------------------
// K4 is a pure HBM write stream of 128-byte (or 64-byte) pieces, and how well HBM takes it depends on WHAT the SMs
// write at the same time, not only on how much.  Measured on B200 (64 panels, 38x38x1024, pool 14; fraction of the
// 6.56 TB/s copy peak): warps free-running 0.86-0.90; a CTA barrier after every round of 64 output columns 0.97; two
// neighbouring slices in a cluster with a cluster barrier every second round 1.01.  The same barriers COST 10-20 %
// on 7x7x512 (VGG-16) and on 16-channel slices (38x50 maps), where the free-running form reaches 0.87 / 0.96.
// (Related: a plain fill kernel writes 6.05 TB/s from 120 SMs but 5.78 TB/s from 148, cudaMemsetAsync 6.8-7.3 TB/s -
// profiles/r02_fill_bw.log, r02_tma_store_bw.log.)  So the choice is measured, not guessed: the first eager call for
// a (device, shape) times the candidates below on the caller's own buffers and the result is cached.
static const int kLockstep[][2] = {{1, 0}, {1, 1}, {1, 2}, {2, 1}, {2, 2}, {2, 4}};     // {CTAs per cluster, sync_every}
constexpr int kLockstepN = sizeof(kLockstep) / sizeof(kLockstep[0]);

// the batch size is not part of the key: the form is a property of what one CTA writes (map shape, channels, pool,
// RoIs per panel, slice width), and a ragged last batch must not be tuned afresh in the middle of a sweep
struct TuneKey {
    int dev, H, W, C, pool, R, lanes;
    bool operator==(const TuneKey &o) const {
        return dev == o.dev && H == o.H && W == o.W && C == o.C && pool == o.pool && R == o.R && lanes == o.lanes;
    }
};
struct TuneEntry {
    TuneKey key;
    int choice;
};
static std::mutex g_tune_mutex;
static TuneEntry g_tune[128];
static int g_tune_n = 0;

static int cached_choice(const TuneKey &k) {
    std::lock_guard<std::mutex> lock(g_tune_mutex);
    for (int i = 0; i < g_tune_n; ++i)
        if (g_tune[i].key == k) return g_tune[i].choice;
    return -1;
}
// first tuned entry of this shape, whatever its slice width
static bool find_shape(int dev, int H, int W, int C, int pool, int R, TuneEntry *out) {
    std::lock_guard<std::mutex> lock(g_tune_mutex);
    for (int i = 0; i < g_tune_n; ++i) {
        const TuneKey &k = g_tune[i].key;
        if (k.dev == dev && k.H == H && k.W == W && k.C == C && k.pool == pool && k.R == R) {
            *out = g_tune[i];
            return true;
        }
    }
    return false;
}
static void store_choice(const TuneKey &k, int choice) {
    std::lock_guard<std::mutex> lock(g_tune_mutex);
    for (int i = 0; i < g_tune_n; ++i)
        if (g_tune[i].key == k) { g_tune[i].choice = choice; return; }
    if (g_tune_n < 128) {
        g_tune[g_tune_n++] = TuneEntry{k, choice};
    } else {                                        // table full: recycle the slots round robin rather than re-tune forever
        static int next = 0;
        g_tune[next] = TuneEntry{k, choice};
        next = (next + 1) % 128;
    }
}

template <typename Launch>
static int tune_lockstep(Launch &&go, cudaStream_t st, int *best_choice) {
    cudaEvent_t e0, e1;
    RADNET_CUDA(cudaEventCreate(&e0));
    RADNET_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    int rc = RADNET_OK;
    *best_choice = 0;
    for (int c = 0; c < kLockstepN; ++c) {
        if ((rc = go(kLockstep[c][0], kLockstep[c][1])) != RADNET_OK) break;     // warm-up
        cudaEventRecord(e0, st);
        for (int k = 0; k < 2 && rc == RADNET_OK; ++k) rc = go(kLockstep[c][0], kLockstep[c][1]);
        cudaEventRecord(e1, st);
        if (rc != RADNET_OK) break;
        if (cudaEventSynchronize(e1) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "lockstep tuning"); break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best * 0.985f) { best = ms; *best_choice = c; }       // a later candidate has to win clearly
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

}  // namespace radnet

using namespace radnet;

extern "C" int radnet_roi_pool_form(int B, int H, int W, int C, int pool, int rois_per_panel, int *h_out3) {
    RADNET_CHECK_ARG(h_out3, "roi_pool_form: null pointer");
    int dev = 0;
    RADNET_CUDA(cudaGetDevice(&dev));
    TuneEntry e;
    (void)B;                        // the form does not depend on the batch size (see TuneKey)
    if (find_shape(dev, H, W, C, pool, rois_per_panel, &e)) {
        h_out3[0] = e.key.lanes;
        h_out3[1] = kLockstep[e.choice][0];
        h_out3[2] = kLockstep[e.choice][1];
    } else {
        h_out3[0] = h_out3[1] = h_out3[2] = -1;
    }
    return RADNET_OK;
}

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
    const int smem_optin = device_smem_optin(dev);
    if (smem_optin < 0) return RADNET_E_CUDA;
    const int smem_limit = smem_optin - 256;        // dynamic bytes a CTA may ask for: the kernels' static variables come on top
    const size_t HW = (size_t)H * W;
    const size_t per_roi = (size_t)pool * sizeof(YEntry) + sizeof(int2);
    const long long form = get_option(kOptRoipoolForm);          // 0 auto, 1 whole-map slices, 2 row bands
    if (C % 4 == 0 && !force_direct() && HW < 65536 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const int C4 = C / 4;
        // band form (a band of map rows per CTA, 32 / 64 / 128-channel slices): only on request (roipool_form = 2,
        // roipool_bands = number of bands, roipool_lanes = float4 lanes per pixel).  It keeps 128-byte store lines for
        // maps whose 32-channel slice does not fit one CTA (38x50), but once the inner loop was no longer
        // issue-bound the whole-map form with 16-channel slices measured 0.95-0.96 of the HBM peak there against
        // 0.73-0.80 for two or three bands (every band CTA rebuilds the y tables and has a different amount of work)
        const int bl_opt = (int)get_option(kOptRoipoolLanes);
        int bl = bl_opt;
        if (bl != 8 && bl != 16 && bl != 32) bl = 8;
        if (C4 % bl == 0 && W <= 256 && H >= 2 && encode_tiled_fn() &&
            form == 2) {
            const size_t per_roi_b = (size_t)pool * sizeof(YEntry) + sizeof(int4) + sizeof(int);
            const size_t scale_bytes = ((size_t)W + 1) * sizeof(float) + 16;
            const size_t budget2 = ((size_t)smem_optin + 1024) / 2 - 1024 - 256; // two CTAs per SM (1 KB reserved + static, each)
            auto band_bytes = [&](int nb) { return ((size_t)((H + nb - 1) / nb + 1) * W + 1) * bl * 16; };
            int nb = (int)get_option(kOptRoipoolBands);
            if (nb < 2 || nb > H) {
                // fewest bands that let two CTAs share an SM; else the fewest that fit at all
                nb = 0;
                for (int k = 2; k <= H && k <= 16 && !nb; ++k)
                    if (band_bytes(k) + scale_bytes + 64 * per_roi_b <= budget2) nb = k;
                for (int k = 2; k <= H && !nb; ++k)
                    if (band_bytes(k) + scale_bytes + 8 * per_roi_b <= (size_t)smem_limit) nb = k;
            }
            const int Hb = nb ? (H + nb - 1) / nb : 0;
            nb = nb ? (H + Hb - 1) / Hb : 0;                                   // bands that own at least one row
            if (nb >= 1 && Hb + 1 <= 256 && band_bytes(nb) + scale_bytes + 8 * per_roi_b <= (size_t)smem_limit &&
                (long long)B * (C4 / bl) * nb < 0x7fffffffLL) {
                const size_t map_bytes = ((size_t)(Hb + 1) * W + 1) * bl * 16;
                size_t budget = map_bytes + scale_bytes + 8 * per_roi_b <= budget2 ? budget2 : (size_t)smem_limit;
                size_t chunk = (budget - map_bytes - scale_bytes) / per_roi_b;
                if (chunk > (size_t)rois_per_panel) chunk = rois_per_panel;
                alignas(64) CUtensorMap tmap;
                memset(&tmap, 0, sizeof(tmap));
                if (make_map_tensor(&tmap, feat, B, H, W, C, bl, Hb + 1)) {
                    p.n_slices = C4 / bl;
                    p.roi_chunk = (int)chunk;
                    p.map_rows_pad = Hb + 1;
                    p.tma_rows = Hb + 1;
                    p.band_rows = Hb;
                    p.n_bands = nb;
                    const size_t smem = map_bytes + chunk * per_roi_b + scale_bytes;
                    if (bl == 8) return launch_band<8>(p, tmap, B, smem, st);
                    if (bl == 16) return launch_band<16>(p, tmap, B, smem, st);
                    return launch_band<32>(p, tmap, B, smem, st);
                }
            }
        }
        const int lanes_opts[4] = {8, 4, 2, 1};
        const int box_rows = (W <= 256 && encode_tiled_fn()) ? tma_box_rows(H) : 0;
        const int rows_pad = box_rows ? (H + box_rows - 1) / box_rows * box_rows : H;
        const int max_lanes = (bl_opt == 4 || bl_opt == 2 || bl_opt == 1) ? bl_opt : 8;   // roipool_lanes caps the slice width
        for (int li = 0; li < 4; ++li) {
            int L = lanes_opts[li];
            if (L > max_lanes) continue;
            size_t map_bytes = ((size_t)rows_pad * W + 1) * L * 16;
            const size_t scale_bytes = ((size_t)W + 1) * sizeof(float) + 16;
            if (C4 % L != 0 || map_bytes + scale_bytes + 8 * per_roi > (size_t)smem_limit) continue;
            size_t chunk = ((size_t)smem_limit - map_bytes - scale_bytes) / per_roi;
            if (chunk > (size_t)rois_per_panel) chunk = rois_per_panel;
            size_t smem = map_bytes + chunk * per_roi + scale_bytes;
            p.n_slices = C4 / L;
            p.roi_chunk = (int)chunk;
            if ((long long)B * p.n_slices >= 0x7fffffffLL) continue;
            p.n_work = B * p.n_slices;
            const long long ctas = get_option(kOptRoipoolCtas);
            p.grid = (ctas > 0 && ctas < p.n_work) ? (int)ctas : p.n_work;
            alignas(64) CUtensorMap tmap;
            memset(&tmap, 0, sizeof(tmap));
            p.map_rows_pad = rows_pad;
            p.tma_rows = (box_rows && make_map_tensor(&tmap, feat, B, H, W, C, L, box_rows)) ? box_rows : 0;
            if (!p.tma_rows && rows_pad != H) continue;            // sized for TMA boxes but no descriptor: next option
            auto go = [&](int cluster, int sync_every) {
                p.cluster = (cluster == 2 || cluster == 4 || cluster == 8) && p.n_slices % cluster == 0 &&
                                    p.grid == p.n_work ? cluster : 1;
                p.sync_every = sync_every;
                switch (L) {
                    case 8: return launch_slice<8>(p, tmap, B, smem, st);
                    case 4: return launch_slice<4>(p, tmap, B, smem, st);
                    case 2: return launch_slice<2>(p, tmap, B, smem, st);
                    default: return launch_slice<1>(p, tmap, B, smem, st);
                }
            };
            const int cs_opt = (int)get_option(kOptRoipoolCluster), every_opt = (int)get_option(kOptRoipoolSyncEvery);
            if (cs_opt >= 0) return go(cs_opt, every_opt > 0 ? every_opt : (cs_opt > 1 ? 2 : 0));
            // automatic lockstep: measured once per (device, shape) on the caller's own buffers (the launches are
            // idempotent), never while the stream is being captured into a graph
            if (p.n_work < 4 * device_sm_count(dev)) return go(1, 0);     // too small to load HBM: free-running
            const TuneKey key{dev, H, W, C, pool, rois_per_panel, L};
            int choice = cached_choice(key);
            if (choice < 0) {
                cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
                if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) cap = cudaStreamCaptureStatusActive;
                if (cap != cudaStreamCaptureStatusNone) return go(1, 0);       // not cached: tuned on the first eager call
                if (int rc = tune_lockstep(go, st, &choice)) return rc;
                store_choice(key, choice);
            }
            return go(kLockstep[choice][0], kLockstep[choice][1]);
        }
    }
    unsigned grid = (unsigned)((long long)B * rois_per_panel * pool);
    if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        roi_pool_direct_kernel<4><<<grid, 256, 0, st>>>(p);
    else
        roi_pool_direct_kernel<1><<<grid, 256, 0, st>>>(p);
    return check_launch("roi_pool_direct_kernel");
}

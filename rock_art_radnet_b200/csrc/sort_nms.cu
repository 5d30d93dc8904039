// sort_nms.cu - K2: segmented score sort + exact greedy NMS, one CTA per panel.
//
// Replaces non_max_suppression_fast (reference faster_rcnn/rpn.py:380-456).
//
// The reference sorts every candidate and then runs up to max_boxes greedy iterations
// (rpn.py:415-450).  The greedy loop stops at max_boxes keeps, which normally happens within
// the first few hundred candidates, so this kernel only orders what it needs:
//
// Stage 0  keys (order-preserving uint images of the scores, 0 = deleted) are staged into
//          shared memory with one TMA bulk copy (cp.async.bulk + mbarrier) on the hot path.
// Stage 1  SELECT: a block-wide bisection on the key value (one count-compare pass per step,
//          ~12 steps) finds a threshold with about `sel_target` (>= 2048) keys at or above it;
//          those keys are compacted in flat-index order.
// Stage 2  SORT: stable LSD radix sort (8-bit digits) of the selected slice only.  Each warp
//          owns a contiguous run; the rank of an element inside a 32-wide row comes from warp
//          ballots (no atomics), per-warp digit counters live in shared memory.  Stable +
//          ascending => reading from the end visits equal scores "higher flat index first",
//          the documented tie rule.
// Stage 3  NMS over tiles of 1024 sorted candidates, one candidate per thread.  (a) test
//          against the boxes kept by earlier tiles; (b) each warp builds the 32x32 "lower lane
//          overlaps me" bit matrix of its row from a shared-memory tile; (c) rows retire in
//          rank order: a warp sleeps on the mbarrier of each earlier row in turn, tests its
//          candidates against the boxes kept since it last looked, resolves its own row with
//          a ballot fixed point and appends its keeps.
// If the selected slice is exhausted before max_boxes boxes are kept (rare), a second round
// sorts everything and the NMS continues at the first unvisited rank.
//
// Exactness: the i32 path evaluates `inter/(union+1e-6) > thr` (rpn.py:443-447) through a
// per-union table of the smallest suppressing intersection, built at kernel start with the
// float64 division itself (boxes are integers, so inter and union are exact); coordinates
// are packed 2 x int16 and compared with VIMNMX.S16x2 / VIADDMNMX.S16x2.RELU.  The f64 path
// performs the float64 arithmetic in the reference's association order.  Both decide
// bit-exactly like NumPy.
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace radnet {

constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kCntStride = 257;                 // padded digit row: conflict-free scan
constexpr int kTile = kNmsThreads;              // candidates per NMS tile
constexpr int kMaxTableEntries = 65535;
constexpr int kKeptSmemMax = 1024;              // kept boxes held in shared memory
constexpr int kLookAhead = 6;                   // rows that may run ahead of the retiring row
constexpr float kLeaderShare = 1.0f;            // cluster form: the leader's share of matrix work relative to a helper's
constexpr int kChainGroups = 4;                 // cluster form: groups of 32 ranks resolved by one warp of the chain
constexpr int kSmemHeader = 3072;               // barriers, misc words, scan scratch, row counts, 256-bin histogram

// ----------------------------------------------------------------------------------
// candidate representations
// ----------------------------------------------------------------------------------
struct BoxI32 {
    // packed candidate: x = (-y1 << 16) | (-x1 & 0xffff), y = (y2 << 16) | x2, z = area, w = flat index
    using Cand = int4;
    static constexpr size_t kBoxBytes = sizeof(int4);
    struct Ctx { const uint16_t *tab; };
    static __device__ __forceinline__ Cand empty() { return make_int4(0, 0, 0, 0); }
    static __device__ __forceinline__ Cand load(const void *boxes, size_t i, int flat) {
        const int4 b = __ldg(reinterpret_cast<const int4 *>(boxes) + i);     // x1,y1,x2,y2
        Cand c;
        c.x = (int)((((unsigned)(-b.y)) << 16) | (((unsigned)(-b.x)) & 0xFFFFu));
        c.y = (int)((((unsigned)b.w) << 16) | (((unsigned)b.z) & 0xFFFFu));
        c.z = (b.z - b.x) * (b.w - b.y);                                     // area, no +1 (rpn.py:412)
        c.w = flat;
        return c;
    }
    static __device__ __forceinline__ int flat(const Cand &c) { return c.w; }
    // one LDS.128 (the compiler otherwise narrows the read to three 32-bit loads)
    static __device__ __forceinline__ Cand load_shared(const Cand *p) {
        Cand c;
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w)
                     : "r"(smem_u32(p)));
        return c;
    }
    static __device__ __forceinline__ int4 unpack(const Cand &c) {
        const int nx1 = (int)(short)(c.x & 0xFFFF), ny1 = c.x >> 16;
        return make_int4(-nx1, -ny1, (int)(short)(c.y & 0xFFFF), c.y >> 16);
    }
    // true when kept box `k` suppresses candidate `c`
    static __device__ __forceinline__ bool suppress(const Cand &k, const Cand &c, const Ctx &ctx) {
        const unsigned nlo = __vmins2((unsigned)k.x, (unsigned)c.x);         // -(max x1), -(max y1)
        const unsigned hi = __vmins2((unsigned)k.y, (unsigned)c.y);          //   min x2 ,   min y2
        const unsigned wh = __viaddmax_s16x2_relu(hi, nlo, 0u);              // max(0, w), max(0, h)
        const int inter = (int)(wh & 0xFFFFu) * (int)(wh >> 16);
        const int uni = k.z + c.z - inter;
        return inter >= (int)ctx.tab[uni];
    }
};

struct CandF64 {
    double x1, y1, x2, y2, area;
    int flat, pad;
};

struct BoxF64 {
    using Cand = CandF64;
    static constexpr size_t kBoxBytes = sizeof(double4);
    struct Ctx { double thr; };
    static __device__ __forceinline__ Cand empty() { return Cand{0, 0, 0, 0, 0, 0, 0}; }
    static __device__ __forceinline__ Cand load(const void *boxes, size_t i, int flat) {
        const double4 b = reinterpret_cast<const double4 *>(boxes)[i];
        Cand c;
        c.x1 = b.x; c.y1 = b.y; c.x2 = b.z; c.y2 = b.w;
        c.area = __dmul_rn(__dsub_rn(b.z, b.x), __dsub_rn(b.w, b.y));                // rpn.py:412
        c.flat = flat; c.pad = 0;
        return c;
    }
    static __device__ __forceinline__ int flat(const Cand &c) { return c.flat; }
    static __device__ __forceinline__ Cand load_shared(const Cand *p) { return *p; }
    static __device__ __forceinline__ bool suppress(const Cand &k, const Cand &c, const Ctx &ctx) {
        double xx1 = fmax(k.x1, c.x1), yy1 = fmax(k.y1, c.y1);                       // rpn.py:429-430
        double xx2 = fmin(k.x2, c.x2), yy2 = fmin(k.y2, c.y2);                       // rpn.py:431-432
        double ww = fmax(0.0, __dsub_rn(xx2, xx1));                                  // rpn.py:434
        double hh = fmax(0.0, __dsub_rn(yy2, yy1));                                  // rpn.py:435
        double inter = __dmul_rn(ww, hh);                                            // rpn.py:437
        double uni = __dsub_rn(__dadd_rn(k.area, c.area), inter);                    // rpn.py:440
        return __ddiv_rn(inter, __dadd_rn(uni, 1e-6)) > ctx.thr;                     // rpn.py:443,447
    }
};

template <typename KeyT> struct KeyInfo;
template <> struct KeyInfo<uint32_t> { static constexpr int passes = 4; static constexpr uint32_t max = 0xFFFFFFFFu; };
template <> struct KeyInfo<uint64_t> { static constexpr int passes = 8; static constexpr uint64_t max = ~0ull; };

struct SortNmsParams {
    const void *boxes;        // [B][N] boxes (int4 or double4)
    const void *keys;         // [B][N] KeyT, 0 = deleted
    int N;
    int max_boxes;            // effective K = min(max_boxes, valid candidates)
    double thr;
    int table_entries;        // i32 only: umax + 1
    int sel_target;           // candidates the first (selective) sort round aims for
    int look_ahead;           // rows that may run ahead of the retiring row in the NMS hand-off
    int cluster_ranks;        // cluster form: top ranks resolved through the cluster-wide overlap matrix
    // outputs
    unsigned char *det;       // i32: detection records
    size_t det_stride;
    int det_max_boxes;
    int32_t *pick;            // f64: picked indices
    int32_t *count;           // f64: {n, ties, n_sorted}
    // global scratch per panel (generic-path sort buffers, big kept lists)
    unsigned char *ws;
    size_t ws_stride;
    size_t ws_off_kA, ws_off_kB, ws_off_iA, ws_off_iB, ws_off_kept;
    // dynamic shared memory carve-up (bytes from the base)
    int sm_off_cnt, sm_off_sort, sm_off_table, sm_off_kept;
    int sort_cap;             // elements per smem sort buffer
    // hot path (uint32 keys staged in shared memory): staged keys, sorted (key, position) words, cluster matrix
    int sm_off_stage, sm_off_sk, sm_off_L;
    int sk_cap;               // sorted words that fit
    int l_bytes;              // bytes available for the cluster overlap matrix
};

// block-wide exclusive scan of one int per thread (1024 threads); returns exclusive prefix,
// total via *total (same value in every thread)
__device__ __forceinline__ int block_exscan(int v, int *s_warp, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = s_warp[lane];
        int ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += n;
        }
        s_warp[lane] = ti - t;            // exclusive warp base
        if (lane == 31) s_warp[32] = ti;  // grand total
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[w] + inc - v;
}

// Lanes of the warp holding the same 8-bit digit as me (valid for active lanes).  Eight
// ballots instead of __match_any_sync: MATCH.ANY serialises over the distinct values of the
// warp (~30 for a random byte) and was the top stall of the first version of this kernel.
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, bool act) {
    uint32_t peers = __ballot_sync(0xffffffffu, act);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool bit = (d >> k) & 1u;
        const uint32_t b = __ballot_sync(0xffffffffu, act && bit);
        peers &= bit ? b : ~b;
    }
    return peers;
}

// One stable radix pass over n elements (all of them take part).  Returns false when the pass
// was skipped because every element shares the digit (block-uniform).
template <typename KeyT, typename IdxT>
__device__ bool radix_pass(const KeyT *in_k, const IdxT *in_i, KeyT *out_k, IdxT *out_i, int n, int shift,
                           uint32_t *s_cnt, int *s_scan, int *s_flag) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kNmsWarps * kCntStride; i += kNmsThreads) s_cnt[i] = 0;
    if (threadIdx.x == 0) *s_flag = 0;
    __syncthreads();
    const int chunk = ((n + kNmsWarps - 1) / kNmsWarps + 31) & ~31;
    const int wbeg = min(w * chunk, n), wend = min(wbeg + chunk, n);
    uint32_t *cnt = s_cnt + w * kCntStride;
    for (int base = wbeg; base < wend; base += 32) {
        const int i = base + lane;
        const bool act = i < wend;
        const KeyT key = act ? in_k[i] : (KeyT)0;
        const uint32_t d = (uint32_t)(key >> shift) & 0xFFu;
        const uint32_t peers = digit_peers(d, act);
        if (act && (peers & lanemask_lt()) == 0) cnt[d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // scan in (digit major, warp minor) order; thread t owns digit t>>2, warps (t&3)*8..+7
    {
        const int d = threadIdx.x >> 2, wq = (threadIdx.x & 3) * 8;
        uint32_t c[8];
        int sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c[j] = s_cnt[(wq + j) * kCntStride + d];
            sum += (int)c[j];
        }
        int dsum = sum;
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
        int total;
        int ex = block_exscan(sum, s_scan, &total);
        if (dsum == total && total > 0 && (threadIdx.x & 3) == 0) *s_flag = 1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s_cnt[(wq + j) * kCntStride + d] = (uint32_t)ex;
            ex += (int)c[j];
        }
        __syncthreads();
        if (*s_flag) return false;
    }
    for (int base = wbeg; base < wend; base += 32) {
        const int i = base + lane;
        const bool act = i < wend;
        const KeyT key = act ? in_k[i] : (KeyT)0;
        const uint32_t d = (uint32_t)(key >> shift) & 0xFFu;
        const uint32_t peers = digit_peers(d, act);
        uint32_t off = 0;
        if (act) {
            off = cnt[d];
            const uint32_t pos = off + __popc(peers & lanemask_lt());
            out_k[pos] = key;
            out_i[pos] = in_i[i];
        }
        __syncwarp();
        if (act && (peers & lanemask_lt()) == 0) cnt[d] = off + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    return true;
}

__device__ __forceinline__ void atomic_min_key(uint32_t *a, uint32_t v) { atomicMin(a, v); }
__device__ __forceinline__ void atomic_max_key(uint32_t *a, uint32_t v) { atomicMax(a, v); }
__device__ __forceinline__ void atomic_min_key(uint64_t *a, uint64_t v) {
    atomicMin(reinterpret_cast<unsigned long long *>(a), (unsigned long long)v);
}
__device__ __forceinline__ void atomic_max_key(uint64_t *a, uint64_t v) {
    atomicMax(reinterpret_cast<unsigned long long *>(a), (unsigned long long)v);
}

// number of valid (non-zero) keys and their min / max -> s_sel[0], s_minmax[0..1]
template <typename KeyT>
__device__ void key_stats(const KeyT *keys, int N, int *s_sel, KeyT *s_minmax) {
    const int lane = threadIdx.x & 31;
    int c = 0;
    KeyT mn = KeyInfo<KeyT>::max, mx = 0;
    for (int i = threadIdx.x; i < N; i += kNmsThreads) {
        const KeyT k = keys[i];
        if (k) {
            ++c;
            mn = k < mn ? k : mn;
            mx = k > mx ? k : mx;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const KeyT on = __shfl_xor_sync(0xffffffffu, mn, d), ox = __shfl_xor_sync(0xffffffffu, mx, d);
        mn = on < mn ? on : mn;
        mx = ox > mx ? ox : mx;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (threadIdx.x == 0) {
        s_sel[0] = 0;
        s_minmax[0] = KeyInfo<KeyT>::max;
        s_minmax[1] = 0;
    }
    __syncthreads();
    if (lane == 0 && c) {
        atomicAdd(&s_sel[0], c);
        atomic_min_key(&s_minmax[0], mn);
        atomic_max_key(&s_minmax[1], mx);
    }
    __syncthreads();
}

// SELECT: a threshold t (>= 1) with count(key >= t) >= target and only a small bucket of extra
// keys.  256-ary search on the key VALUE range: per-warp 256-bin histograms of (key - lo) >> shift
// (shared-memory RED, no returns), column sums, a suffix scan by warp 0 from the top bin; the
// bucket that straddles the target rank becomes the new [lo, hi].  One or two rounds in practice
// (13k keys over 256 bins leave ~50 keys in the boundary bucket).
// M = number of valid (non-zero) keys, S = count(key >= t).
template <typename KeyT>
__device__ KeyT select_threshold(const KeyT *keys, int N, int target, uint32_t *s_cnt, uint32_t *s_hist,
                                 int *s_sel, KeyT *s_minmax, int &M, int &S) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    key_stats<KeyT>(keys, N, s_sel, s_minmax);
    M = s_sel[0];
    KeyT lo = s_minmax[0], hi = s_minmax[1];
    S = M;
    if (M <= target) return (KeyT)1;
    int above = 0, need = target, inb = M;
    const int slack = max(target >> 3, 1);
#pragma unroll 1
    for (int it = 0; it < (int)sizeof(KeyT) + 1; ++it) {
        const KeyT range = hi - lo;
        if (range == 0) break;                                   // a single key value: all ties
        const int bits = (sizeof(KeyT) == 8) ? 64 - __clzll((long long)range) : 32 - __clz((int)range);
        const int shift = bits > 8 ? bits - 8 : 0;
        __syncthreads();                                         // previous round done with s_cnt / s_sel
        for (int i = threadIdx.x; i < kNmsWarps * kCntStride; i += kNmsThreads) s_cnt[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += kNmsThreads) {
            const KeyT k = keys[i];
            if (k >= lo && k <= hi) atomicAdd(&s_cnt[w * kCntStride + (int)((k - lo) >> shift)], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t t = 0;
            for (int ww = 0; ww < kNmsWarps; ++ww) t += s_cnt[ww * kCntStride + threadIdx.x];
            s_hist[threadIdx.x] = t;
        }
        __syncthreads();
        if (w == 0) {
            // lane l owns bins 255-8l .. 248-8l, so lane order = descending key order
            uint32_t loc[8];
            int sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                loc[j] = s_hist[255 - 8 * lane - j];
                sum += (int)loc[j];
            }
            int inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int n = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += n;
            }
            int abv = inc - sum;                                 // keys in the bins of lower lanes (larger keys)
            if (abv < need && need <= inc) {                     // exactly one lane (need <= keys in range)
                int b = 255 - 8 * lane, cnt = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (abv + (int)loc[j] >= need) { b = 255 - 8 * lane - j; cnt = (int)loc[j]; break; }
                    abv += (int)loc[j];
                }
                s_sel[1] = b; s_sel[2] = abv; s_sel[3] = cnt;
            }
        }
        __syncthreads();
        const int b = s_sel[1];
        above += s_sel[2];
        need -= s_sel[2];
        inb = s_sel[3];
        const KeyT nlo = lo + ((KeyT)b << shift);
        const KeyT span = (((KeyT)1) << shift) - 1;
        hi = (hi - nlo > span) ? nlo + span : hi;
        lo = nlo;
        if (shift == 0 || inb <= slack) break;
    }
    S = above + inb;
    return lo;
}

constexpr int kWideBins = 2048;

// SELECT by sampling (hot path): the threshold only has to cut off "about sel_target" top keys - the
// NMS normally stops long before the slice is used up, and if it does not, a second round sorts the
// rest - so it is estimated from every 8th key: statistics and a 2048-bin histogram of ~N/8 keys instead
// of two passes over all of them.  The cut is placed 15 % (+8 keys) below the target rank of the sample,
// which makes a slice smaller than sel_target unlikely (and harmless).  Returns the threshold key; the
// exact size of the slice and the number of valid keys come out of the bucket sort that follows.
constexpr int kSampleStride = 8;
__device__ uint32_t select_threshold_sampled(const uint32_t *keys, int N, int target, uint32_t *s_hist2k,
                                             int *s_scan, int *s_sel, uint32_t *s_minmax) {
    const int lane = threadIdx.x & 31;
    const int ns = (N + kSampleStride - 1) / kSampleStride;
    {
        int c = 0;
        uint32_t mn = 0xFFFFFFFFu, mx = 0;
        for (int j = threadIdx.x; j < ns; j += kNmsThreads) {
            const uint32_t k = keys[j * kSampleStride];
            if (k) { ++c; mn = min(mn, k); mx = max(mx, k); }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        c = __reduce_add_sync(0xffffffffu, c);
        if (threadIdx.x == 0) { s_sel[0] = 0; s_minmax[0] = 0xFFFFFFFFu; s_minmax[1] = 0; }
        __syncthreads();
        if (lane == 0 && c) {
            atomicAdd(&s_sel[0], c);
            atomicMin(&s_minmax[0], mn);
            atomicMax(&s_minmax[1], mx);
        }
        __syncthreads();
    }
    const int Ms = s_sel[0];
    const uint32_t lo = s_minmax[0], hi = s_minmax[1];
    const int need = (target + kSampleStride - 1) / kSampleStride * 23 / 20 + 8;
    if (Ms <= need || hi <= lo) return 1u;              // few valid keys (or one value): take them all
    const uint32_t range = hi - lo;
    const int bits = 32 - __clz((int)range);
    const int shift = bits > 11 ? bits - 11 : 0;
    for (int i = threadIdx.x; i < kWideBins; i += kNmsThreads) s_hist2k[i] = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < ns; j += kNmsThreads) {
        const uint32_t k = keys[j * kSampleStride];
        if (k) atomicAdd(&s_hist2k[(k - lo) >> shift], 1u);
    }
    __syncthreads();
    const int b0 = kWideBins - 1 - 2 * (int)threadIdx.x;           // thread order = descending key order
    const int c0 = (int)s_hist2k[b0], c1 = (int)s_hist2k[b0 - 1];
    int total;
    const int abv = block_exscan(c0 + c1, s_scan, &total);
    if (abv < need && need <= abv + c0 + c1) s_sel[1] = (abv + c0 < need) ? b0 - 1 : b0;   // exactly one thread
    __syncthreads();
    const uint32_t t = lo + ((uint32_t)s_sel[1] << shift);
    return t ? t : 1u;
}

// SORT, hot-path form: bucket sort of the selected keys by value.  A 2048-bin histogram over [t, max]
// holds about one key per bin, a block scan from the top bin turns it into descending start offsets,
// every selected key is scattered into its bucket as a (key << 32 | position) word and each thread
// orders its two buckets (descending, i.e. higher position first among equal keys) by insertion.
// Returns the number of sorted words, or -1 when a bucket is too crowded for this scheme (many nearly
// equal scores) - the caller then takes the general radix path.
constexpr int kBucketMax = 24;
__device__ int bucket_sort_desc(const uint32_t *keys, int N, uint32_t t, uint32_t hi, unsigned long long *out,
                                int cap, uint32_t *s_hist2k, int *s_start, int *s_scan, int *s_flag, int *s_valid,
                                int *s_ties, long long *dbg = nullptr) {
#define BS_STAMP(i) do { if (dbg && threadIdx.x == 0) dbg[i] = clock64(); } while (0)
    BS_STAMP(0);
    const uint32_t range = hi > t ? hi - t : 0u;
    const int bits = range ? 32 - __clz((int)range) : 0;
    const int shift = bits > 11 ? bits - 11 : 0;
    for (int i = threadIdx.x; i < kWideBins; i += kNmsThreads) s_hist2k[i] = 0;
    if (threadIdx.x == 0) { *s_flag = 0; *s_valid = 0; }
    __syncthreads();
    BS_STAMP(1);
    // pass 1 over all keys: histogram of the selected ones (keys above `hi`, which may come from a sample,
    // share the top bin), count of the valid ones, and a per-thread bit mask of which of my keys are selected
    uint32_t mine = 0;
    int nvalid = 0;
    {
        int q = 0;
        for (int i = threadIdx.x; i < N; i += kNmsThreads, ++q) {
            const uint32_t k = keys[i];
            nvalid += k ? 1 : 0;
            if (k >= t) {
                atomicAdd(&s_hist2k[min((k - t) >> shift, (uint32_t)(kWideBins - 1))], 1u);
                if (q < 32) mine |= 1u << q;
            }
        }
    }
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    if ((threadIdx.x & 31) == 0 && nvalid) atomicAdd(s_valid, nvalid);
    __syncthreads();
    BS_STAMP(2);
    const int b0 = kWideBins - 1 - 2 * (int)threadIdx.x;           // thread order = descending key order
    const int c0 = (int)s_hist2k[b0], c1 = (int)s_hist2k[b0 - 1];
    int total;
    const int ex = block_exscan(c0 + c1, s_scan, &total);
    s_start[b0] = ex;
    s_start[b0 - 1] = ex + c0;
    if (c0 > kBucketMax || c1 > kBucketMax || total > cap || N > 32 * kNmsThreads) *s_flag = 1;
    __syncthreads();
    BS_STAMP(3);
    if (*s_flag) return -1;
    // pass 2 touches only the selected keys
    while (mine) {
        const int q = __ffs(mine) - 1;
        mine &= mine - 1;
        const int i = threadIdx.x + q * kNmsThreads;
        const uint32_t k = keys[i];
        const uint32_t b = min((k - t) >> shift, (uint32_t)(kWideBins - 1));
        const int slot = (int)atomicSub(&s_hist2k[b], 1u) - 1;
        out[s_start[b] + slot] = ((unsigned long long)k << 32) | (unsigned)i;
    }
    __syncthreads();
    BS_STAMP(4);
    int ties = 0;              // equal keys share a bucket: count the tied entries here, no extra pass
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        unsigned long long *a = out + (q ? ex + c0 : ex);
        const int n = q ? c1 : c0;
        for (int i = 1; i < n; ++i) {
            const unsigned long long v = a[i];
            int j = i - 1;
            while (j >= 0 && a[j] < v) { a[j + 1] = a[j]; --j; }
            a[j + 1] = v;
        }
        for (int i = 0; i < n; ++i) {
            const uint32_t k = (uint32_t)(a[i] >> 32);
            ties += ((i > 0 && (uint32_t)(a[i - 1] >> 32) == k) || (i + 1 < n && (uint32_t)(a[i + 1] >> 32) == k)) ? 1 : 0;
        }
    }
    if (ties) atomicAdd(s_ties, ties);
    __syncthreads();
    BS_STAMP(5);
#undef BS_STAMP
    return total;
}

// order-preserving compaction of the keys >= t with their positions; returns the count
template <typename KeyT, typename IdxT>
__device__ int compact_ge(const KeyT *keys, int N, KeyT t, KeyT *out_k, IdxT *out_i, int *s_scan) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int chunk = ((N + kNmsWarps - 1) / kNmsWarps + 31) & ~31;
    const int wbeg = min(w * chunk, N), wend = min(wbeg + chunk, N);
    int cnt = 0;
    for (int base = wbeg; base < wend; base += 32) {
        const int i = base + lane;
        cnt += __popc(__ballot_sync(0xffffffffu, i < wend && keys[i] >= t));
    }
    if (lane == 0) s_scan[w] = cnt;
    __syncthreads();
    if (w == 0) {
        const int v = s_scan[lane];
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += n;
        }
        s_scan[lane] = inc - v;
        if (lane == 31) s_scan[32] = inc;
    }
    __syncthreads();
    int pos = s_scan[w];
    const int total = s_scan[32];
    for (int base = wbeg; base < wend; base += 32) {
        const int i = base + lane;
        const KeyT k = (i < wend) ? keys[i] : (KeyT)0;
        const bool act = (i < wend) && k >= t;
        const uint32_t m = __ballot_sync(0xffffffffu, act);
        if (act) {
            const int o = pos + __popc(m & lanemask_lt());
            out_k[o] = k;
            out_i[o] = (IdxT)i;
        }
        pos += __popc(m);
    }
    __syncthreads();
    return total;
}

// Is candidate `cand` suppressed by any of kept[from, to)?  The kept boxes are fetched in
// batches of kBatch before any of them is tested, so the dependent LDS -> ALU -> LDS(table)
// chains of a batch overlap; the last batch re-reads the final entry instead of running a
// one-by-one remainder loop (testing the same box twice is harmless).
template <typename Traits, bool kKeptSmem, int kBatch>
__device__ __forceinline__ bool suppressed_by(const typename Traits::Cand *kept, int from, int to,
                                              const typename Traits::Cand &cand,
                                              const typename Traits::Ctx &ctx) {
    bool hit = false;
    for (int j = from; j < to; j += kBatch) {
        typename Traits::Cand kb[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int jj = min(j + u, to - 1);
            kb[u] = kKeptSmem ? Traits::load_shared(kept + jj) : kept[jj];
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) hit |= Traits::suppress(kb[u], cand, ctx);
    }
    return hit;
}

template <typename Traits, typename KeyT, typename IdxT, bool kSmemSort, bool kKeptSmem, bool kCluster = false,
          bool kStagedOnly = false>
__global__ void __launch_bounds__(kNmsThreads, 1) sort_nms_kernel(SortNmsParams p) {
    using Cand = typename Traits::Cand;
    constexpr bool kI32 = sizeof(Cand) == sizeof(int4);
    // kSmemSort: the general sort's ping-pong buffers live in shared memory (and the hot path borrows them);
    // kStagedOnly: only the hot path's buffers do (larger panels), the general sort runs in global scratch
    constexpr bool kStaged = kSmemSort || kStagedOnly;
    constexpr bool kFast = kStaged && sizeof(KeyT) == 4;
    constexpr int kTestBatch = kI32 ? 8 : 2;      // kept boxes fetched per batch (register budget)

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem);                 // TMA barrier
    uint64_t *s_lbar = reinterpret_cast<uint64_t *>(smem + 8);            // cluster form: matrix blocks landed (leader)
    int *s_misc = reinterpret_cast<int *>(smem + 16);                     // 16 ints
    int *s_scan = reinterpret_cast<int *>(smem + 96);                     // 34 ints
    uint64_t *s_turn = reinterpret_cast<uint64_t *>(smem + 256);          // [32] hand-off barriers
    KeyT *s_minmax = reinterpret_cast<KeyT *>(smem + 512);                // [2]
    volatile int *s_kafter = reinterpret_cast<volatile int *>(smem + 1536);   // [32] kept count after row w
    // cluster form: per group of 32 ranks {kept count after the group + 1, lanes kept}; 0 = not yet published
    volatile unsigned long long *s_pub = reinterpret_cast<volatile unsigned long long *>(smem + 1664);   // [32]
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem + p.sm_off_cnt);  // [32][257]
    volatile int *s_kcount = s_misc + 0;
    int *s_flag = s_misc + 2;
    int *s_ties = s_misc + 3;
    int *s_sel = s_misc + 4;                                              // 4 ints
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + 2048);         // [256]
    volatile int *s_fault = s_misc + 8;

    // cluster form: the CTAs of one cluster share a panel; rank 0 leads, the others help with the
    // overlap matrix of the first ranks and then leave
    const uint32_t crank = kCluster ? cluster_ctarank() : 0u;
    const uint32_t csize = kCluster ? cluster_nctarank() : 1u;
    const int seg = kCluster ? (int)(blockIdx.x / csize) : (int)blockIdx.x;
    const int N = p.N;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const KeyT *g_keys = reinterpret_cast<const KeyT *>(p.keys) + (size_t)seg * N;
    const unsigned char *g_boxes = reinterpret_cast<const unsigned char *>(p.boxes) + (size_t)seg * N * Traits::kBoxBytes;
    unsigned char *ws = p.ws + (size_t)seg * p.ws_stride;

    KeyT *kA, *kB;
    IdxT *iA, *iB;
    if (kSmemSort) {
        unsigned char *sb = smem + p.sm_off_sort;
        kA = reinterpret_cast<KeyT *>(sb);
        kB = kA + p.sort_cap;
        iA = reinterpret_cast<IdxT *>(kB + p.sort_cap);
        iB = iA + p.sort_cap;
    } else {
        kA = reinterpret_cast<KeyT *>(ws + p.ws_off_kA);
        kB = reinterpret_cast<KeyT *>(ws + p.ws_off_kB);
        iA = reinterpret_cast<IdxT *>(ws + p.ws_off_iA);
        iB = reinterpret_cast<IdxT *>(ws + p.ws_off_iB);
    }

#ifdef RADNET_NMS_PROFILE
    long long *prof = reinterpret_cast<long long *>(ws + p.ws_off_kept + 32768);
    int prof_n = 0;
#define NMS_STAMP() do { __syncthreads(); if (threadIdx.x == 0 && prof_n < 16 && crank == 0) prof[prof_n] = clock64(); ++prof_n; } while (0)
#define NMS_ROW_STAMP(k) do { if (lane == 0 && tile_no == 0) prof[16 + w * 8 + (k)] = clock64(); } while (0)
#else
#define NMS_STAMP() do {} while (0)
#define NMS_ROW_STAMP(k) do {} while (0)
#endif
    NMS_STAMP();
    if (kCluster && threadIdx.x < 32) s_pub[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) {
        *s_kcount = 0;
        *s_ties = 0;
        *s_fault = 0;
        mbar_init(s_bar, 1);
        for (int i = 0; i < kNmsWarps; ++i) mbar_init(&s_turn[i], 1);
        if (kCluster) mbar_init(s_lbar, csize);       // one arrival per CTA of the cluster (matrix blocks)
        mbar_fence_init();
    }
    __syncthreads();
    // cluster form, split barrier phase A: "my barriers exist" - waited for right before the first
    // access to another CTA's shared memory, by which time everybody has long arrived
    if constexpr (kCluster) cluster_arrive();

    // ---- stage 0: raw keys -> shared memory (one TMA bulk copy on the hot path) ----------
    const KeyT *raw_k = g_keys;
    bool tma_pending = false;
    if (kStaged) {
        KeyT *stage = kSmemSort ? kA : reinterpret_cast<KeyT *>(smem + p.sm_off_stage);
        const size_t bytes = (size_t)N * sizeof(KeyT);
        const uint32_t bulk = ((reinterpret_cast<uintptr_t>(g_keys) & 15) == 0) ? (uint32_t)(bytes & ~(size_t)15) : 0u;
        if (threadIdx.x == 0 && bulk) {
            mbar_expect_tx(s_bar, bulk);
            tma_bulk_g2s(stage, g_keys, bulk, s_bar);
        }
        // tail (and the whole array when the source is not 16-byte aligned)
        for (int i = (int)(bulk / sizeof(KeyT)) + threadIdx.x; i < N; i += kNmsThreads) stage[i] = g_keys[i];
        raw_k = stage;
        tma_pending = bulk != 0;
    }

    // ---- suppression table, built while the keys land (i32 path) ------------------------
    uint16_t *s_tab = reinterpret_cast<uint16_t *>(smem + p.sm_off_table);
    if (kI32) {
        const double thr = p.thr;
        for (int u = threadIdx.x; u < p.table_entries; u += kNmsThreads) {
            // tab[u] = smallest inter with inter/(u+1e-6) > thr.  In exact arithmetic that is the first
            // integer above pm = thr*(u+1e-6); the rounded divide can only disagree when pm is within a few
            // ulps of an integer, so the divide itself is consulted only then (the table stays bit-exact and
            // costs two float64 multiplies per entry instead of three divides).
            const double d = __dadd_rn((double)u, 1e-6);
            const double pm = __dmul_rn(thr, d);
            const double fl = floor(pm);
            const double eps = __dmul_rn(1e-12, fabs(pm) + 1.0);
            int c;
            bool ok;
            if (pm - fl > eps && (fl + 1.0) - pm > eps) {
                const double first = fmax(fl + 1.0, 0.0);
                ok = first <= (double)u && first < 65535.0;
                c = ok ? (int)first : 0;
            } else {
                const double g = fl - 1.0;
                c = (g > 0.0) ? ((g < 65534.0) ? (int)g : 65535) : 0;
                const int lim = c + 4;
                while (c < lim && c <= u && !(__ddiv_rn((double)c, d) > thr)) ++c;
                ok = (c <= u) && (c < lim) && (__ddiv_rn((double)c, d) > thr);
            }
            s_tab[u] = ok ? (uint16_t)c : (uint16_t)0xFFFF;
        }
    }
    if (tma_pending) mbar_wait(s_bar, 0);
    __syncthreads();
    NMS_STAMP();     // 1: keys staged + table built

    // kept list
    Cand *kept;
    {
        unsigned char *kb = kKeptSmem ? (smem + p.sm_off_kept) : (ws + p.ws_off_kept);
        kept = reinterpret_cast<Cand *>(kb);
    }
    Cand *s_tile = reinterpret_cast<Cand *>(s_cnt);   // counters are dead during the NMS
    typename Traits::Ctx ctx;
    if constexpr (kI32) ctx.tab = s_tab; else ctx.thr = p.thr;

    int M = 0;            // valid candidates
    int S = 0;            // candidates sorted so far (ranks [0,S) are final)
    int done = 0;         // ranks already visited by the NMS
    int k0 = 0;
    int K = 0;
    int tile_no = 0;
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        // ---- stage 1 + 2: select, compact, sort --------------------------------------------
        const KeyT *in_k = nullptr;
        const IdxT *in_i = nullptr;
        bool sorted_fast = false;
        const unsigned long long *sk_sorted = nullptr;      // hot path: descending (key, position) words
        if constexpr (kFast) {
            if (round == 0) {
                // hot path: sampled select, then a value-bucket sort of the selected keys as (key, position)
                // words in descending order.  Neither touches the area of the cluster overlap matrix (in the
                // full shared-memory layout: sorted words in the second index buffer, matrix in the second
                // key buffer + first index buffer).
                unsigned long long *sk = reinterpret_cast<unsigned long long *>(smem + p.sm_off_sk);
                const int sk_cap = p.sk_cap;
                uint32_t *hist = reinterpret_cast<uint32_t *>(s_cnt);
                const KeyT thr_key = select_threshold_sampled(raw_k, N, p.sel_target, hist, s_scan, s_sel, s_minmax);
                NMS_STAMP();     // 2: threshold selected
                const int got = bucket_sort_desc(raw_k, N, thr_key, s_minmax[1], sk, sk_cap, hist,
                                                 reinterpret_cast<int *>(hist + kWideBins), s_scan, s_flag, &s_sel[2], s_ties
#ifdef RADNET_NMS_PROFILE
                                                 , (crank == 0) ? prof + 300 : nullptr
#endif
                                                 );
                NMS_STAMP();     // 3: (compacted)
                if (got >= 0) {
                    S = got;
                    M = s_sel[2];
                    K = min(p.max_boxes, M);
                    sk_sorted = sk;
                    sorted_fast = true;
                }
                // else: crowded buckets (many nearly equal scores) - the general path below selects and sorts
            }
        }
        if (!sorted_fast) {
        KeyT thr_key = (KeyT)1;
        if (round == 0) {
            thr_key = select_threshold<KeyT>(raw_k, N, p.sel_target, s_cnt, s_hist, s_sel, s_minmax, M, S);
            K = min(p.max_boxes, M);
            if constexpr (!kFast) NMS_STAMP();     // 2: threshold selected
        } else if (kSmemSort) {
            // the ping-pong buffers overwrote the staged keys: fetch them again
            for (int i = threadIdx.x; i < N; i += kNmsThreads) kA[i] = g_keys[i];
            __syncthreads();
        }
        // smem path: raw keys live in kA, so the compacted slice goes to kB
        KeyT *ck = kSmemSort ? kB : kA;
        IdxT *ci = kSmemSort ? iB : iA;
        S = compact_ge<KeyT, IdxT>(raw_k, N, thr_key, ck, ci, s_scan);
        if (round == 1) M = S;
        NMS_STAMP();         // 3: compacted

        // ---- stable LSD radix sort of the slice ------------------------------------------
        in_k = ck;
        in_i = ci;
        KeyT *out_k = (ck == kA) ? kB : kA;
        IdxT *out_i = (ci == iA) ? iB : iA;
        if (S > 1) {
#pragma unroll 1
            for (int pass = 0; pass < KeyInfo<KeyT>::passes; ++pass) {
                const bool moved = radix_pass<KeyT, IdxT>(in_k, in_i, out_k, out_i, S, pass * 8, s_cnt, s_scan, s_flag);
#ifdef RADNET_NMS_PROFILE
                if (threadIdx.x == 0 && round == 0) prof[8 + pass] = clock64() * 2 + (moved ? 1 : 0);
#endif
                if (moved) {
                    const KeyT *nk = out_k;
                    const IdxT *ni = out_i;
                    out_k = const_cast<KeyT *>(in_k);
                    out_i = const_cast<IdxT *>(in_i);
                    in_k = nk;
                    in_i = ni;
                }
            }
        }
        }
        // in_k / in_i: S entries ascending by (score, flat index)
        NMS_STAMP();         // 4: sorted

        // ---- score ties among the sorted candidates (reported) -----------
        // (the hot path counted them inside its buckets)
        if (!sorted_fast) {
            if (threadIdx.x == 0) *s_ties = 0;
            __syncthreads();
            int t = 0;
            for (int i = threadIdx.x; i < S; i += kNmsThreads) {
                const KeyT k = in_k[i];
                const bool tie = (i > 0 && in_k[i - 1] == k) || (i + 1 < S && in_k[i + 1] == k);
                t += tie ? 1 : 0;
            }
            t = __reduce_add_sync(0xffffffffu, t);
            if (lane == 0 && t) atomicAdd(s_ties, t);
        }
        __syncthreads();

        // ---- stage 3a (cluster form): the first C1 ranks through a cluster-wide overlap matrix ----
        // Every CTA of the cluster has just computed the same sorted order.  Row b of the lower-
        // triangular matrix L holds, one bit per earlier rank a < b, "a overlaps b beyond the
        // threshold"; rows are dealt round-robin over the CTAs (warp per row, one ballot per 32 pairs)
        // and written straight into the leader's shared memory (DSMEM).  After one cluster barrier the
        // helpers leave and ONE warp of the leader resolves the greedy chain with bit operations only:
        // for each group of 32 ranks, dead = OR_g' (L[b][g'] & keepmask[g']), then the ballot fixed
        // point inside the group.  No box test remains on the sequential path.
        int first = done;
        if constexpr (kCluster) {
            if (round == 0) {
                cluster_wait();          // phase A: every CTA of the cluster runs and has its barriers
                int C1 = 0, stride = 1;
                const size_t l_bytes = (size_t)p.l_bytes;
                if (sorted_fast) {
                    C1 = min(S, p.cluster_ranks);
                    while (C1 > 0 && (size_t)((C1 + 31) & ~31) * (size_t)(((C1 + 31) >> 5) | 1) * 4 > l_bytes) C1 -= 32;
                    if (C1 < 64) C1 = 0;
                    stride = ((C1 + 31) >> 5) | 1;               // odd: lanes of a group hit distinct banks
                }
                uint32_t *Lm = reinterpret_cast<uint32_t *>(smem + p.sm_off_L);
                if (C1 > 0) {
                    const int r = threadIdx.x;
                    Cand cand = Traits::empty();
                    if (r < C1) {
                        const int flat = (int)(uint32_t)sk_sorted[r];
                        cand = Traits::load(g_boxes, (size_t)flat, flat);
                    }
                    s_tile[r] = cand;
                    __syncthreads();
                    NMS_STAMP();         // (cluster) candidates gathered
                    // my block of rows: row b costs ~b pair tests, so equal work means boundaries ~ sqrt
                    const int C1r = (C1 + 31) & ~31;
                    // (the leader takes a smaller share: its rows need no copy, the helpers' blocks still
                    // have to cross the cluster after they are computed)
                    auto bound = [&](uint32_t c) {
                        if (c >= csize) return C1r;
                        if (c == 0) return 0;
                        const float lead = kLeaderShare / (float)csize;
                        const float area = (csize > 1) ? lead + (float)(c - 1) * (1.f - lead) / (float)(csize - 1) : 1.f;
                        const int v = (int)sqrtf(area * (float)C1r * (float)C1r);
                        return min(C1r, v & ~3);       // 4 rows of an odd number of words = a multiple of 16 bytes
                    };
                    const int b0 = bound(crank), b1 = bound(crank + 1);
                    for (int b = b0 + w; b < b1; b += kNmsWarps) {
                        const Cand cb = Traits::load_shared(s_tile + b);       // b < C1r <= kTile; rows >= C1 hold empty boxes
                        const int last = b >> 5;
                        uint32_t *row = Lm + b * stride;
                        // every pair is tested unconditionally (no divergent guard); the diagonal word is
                        // masked to the earlier ranks afterwards
#pragma unroll 4
                        for (int wd = 0; wd < last; ++wd) {
                            const bool hit = Traits::suppress(Traits::load_shared(s_tile + (wd << 5) + lane), cb, ctx);
                            const uint32_t bits = __ballot_sync(0xffffffffu, hit);
                            if (lane == 0) row[wd] = bits;
                        }
                        const bool hit = Traits::suppress(Traits::load_shared(s_tile + (last << 5) + lane), cb, ctx);
                        const uint32_t bits = __ballot_sync(0xffffffffu, hit) & ((1u << (b & 31)) - 1u);
                        if (lane == 0) row[last] = (b < C1) ? bits : 0u;
                    }
                    if (crank != 0) fence_proxy_async_smem();     // my rows -> visible to the bulk-copy engine
                    __syncthreads();
                    NMS_STAMP();         // (cluster) my rows of the matrix written
                    if (crank != 0) {
                        // ship my block into the leader's matrix with one bulk copy that completes on ITS mbarrier
                        if (threadIdx.x == 0) {
                            const uint32_t bytes = (uint32_t)(b1 - b0) * (uint32_t)stride * 4u;
                            const uint32_t rbar = cluster_map_shared(s_lbar, 0);
                            mbar_remote_arrive_expect_tx(rbar, bytes);
                            if (bytes) bulk_s2s_cluster(cluster_map_shared(Lm + b0 * stride, 0), Lm + b0 * stride, bytes, rbar);
                        }
                        // phase B: stay until the leader has seen every block (my shared memory is the source)
                        cluster_arrive();
                        cluster_wait();
                        return;
                    }
                    if (threadIdx.x == 0) mbar_arrive(s_lbar);
                    mbar_wait(s_lbar, 0);
                    cluster_arrive();    // phase B: the helpers may leave
                    cluster_wait();
                    NMS_STAMP();         // (cluster) matrix complete
                    // greedy chain.  A warp owns kChainGroups consecutive groups of 32 ranks (one candidate of
                    // each group per lane).  While the earlier warps work it folds their published keep masks
                    // into the dead bits of all its candidates; at its turn it resolves its groups one after the
                    // other in registers (ballot + vote per group when no two live candidates of the group
                    // overlap, the ballot fixed point otherwise) and publishes ONE 64-bit word per group
                    // {kept count after the group + 1, keep mask}.  The hand-off between warps is a single
                    // shared-memory store and a polled load; no box test is left on the sequential path.
                    {
                        const int groups = (C1 + 31) >> 5;
                        const int g0 = w * kChainGroups;
                        const uint32_t pub_addr = smem_u32(const_cast<unsigned long long *>(s_pub));
                        if (g0 < groups) {
                            const uint32_t *row[kChainGroups];
                            uint32_t dead[kChainGroups];
                            bool valid[kChainGroups];
#pragma unroll
                            for (int q = 0; q < kChainGroups; ++q) {
                                const int b = ((g0 + q) << 5) + lane;
                                valid[q] = b < C1;
                                row[q] = Lm + (size_t)min(b, C1r - 1) * stride;        // rows up to C1r exist
                                dead[q] = 0;
                            }
                            int kc = 0;
                            for (int gp = 0; gp < g0; ++gp) {
                                uint32_t cur[kChainGroups];
#pragma unroll
                                for (int q = 0; q < kChainGroups; ++q) cur[q] = row[q][gp];
                                unsigned long long pub;
                                while ((pub = lds_volatile_u64(pub_addr + 8u * (uint32_t)gp)) == 0ull) {}
#pragma unroll
                                for (int q = 0; q < kChainGroups; ++q) dead[q] |= cur[q] & (uint32_t)pub;
                                kc = (int)(pub >> 32) - 1;
                            }
                            uint32_t mine[kChainGroups];               // keep masks of my own groups
                            int at[kChainGroups];                      // kept count before each of them
                            // (the matrix words my groups need from each other were fetched while waiting)
                            uint32_t lower_of[kChainGroups], cross[kChainGroups][kChainGroups];
#pragma unroll
                            for (int q = 0; q < kChainGroups; ++q) {
                                lower_of[q] = (g0 + q < groups) ? row[q][g0 + q] : 0u;
#pragma unroll
                                for (int e = 0; e < q; ++e) cross[q][e] = (g0 + q < groups) ? row[q][g0 + e] : 0u;
                            }
#pragma unroll
                            for (int q = 0; q < kChainGroups; ++q) {
                                const int g = g0 + q;
                                mine[q] = 0;
                                at[q] = kc;
                                if (g >= groups) continue;
                                uint32_t keep = 0;
                                int nk = 0;
                                if (kc < K) {
                                    uint32_t d = dead[q];
#pragma unroll
                                    for (int e = 0; e < q; ++e) d |= cross[q][e] & mine[e];
                                    const uint32_t lower = lower_of[q];
                                    const bool me0 = valid[q] && !d;
                                    uint32_t und = __ballot_sync(0xffffffffu, me0);
                                    if (!__any_sync(0xffffffffu, me0 && (lower & und))) {
                                        keep = und;                    // no two live candidates of the group overlap
                                    } else {
                                        while (und) {
                                            const bool me = me0 && ((und >> lane) & 1u);
                                            const uint32_t know = __ballot_sync(0xffffffffu, me && !(lower & keep) && !(lower & und));
                                            keep |= know;
                                            const uint32_t dnow = __ballot_sync(0xffffffffu, me && (lower & keep));
                                            und &= ~(know | dnow);
                                        }
                                    }
                                    const int room = K - kc;
                                    nk = __popc(keep);
                                    if (nk > room) {
                                        keep &= (1u << __fns(keep, 0, room + 1)) - 1u;
                                        nk = room;
                                    }
                                }
                                // publish first: the successors only need the mask and the count
                                if (lane == 0)
                                    sts_volatile_u64(pub_addr + 8u * (uint32_t)g, ((unsigned long long)(unsigned)(kc + nk + 1) << 32) | keep);
                                mine[q] = keep;
                                kc += nk;
#ifdef RADNET_NMS_PROFILE
                                if (lane == 0) prof[16 + g * 8 + 7] = clock64();      // after the hand-off
#endif
                                if (g == groups - 1 && lane == 0) *s_kcount = kc;
                            }
                            // the kept boxes themselves are only read after the block-wide barrier below
#pragma unroll
                            for (int q = 0; q < kChainGroups; ++q)
                                if ((mine[q] >> lane) & 1u)
                                    kept[at[q] + __popc(mine[q] & lanemask_lt())] = Traits::load_shared(s_tile + ((g0 + q) << 5) + lane);
                        }
                    }
                    __syncthreads();
                    k0 = *s_kcount;
                    first = C1;
                    __syncthreads();
                } else if (crank != 0) {
                    return;              // nothing to share: the leader goes on alone
                }
            }
        }

        // ---- stage 3: greedy suppression over ranks [first, S) ---------------------------
#pragma unroll 1
        for (int base = first; base < S && k0 < K; base += kTile, ++tile_no) {
            const uint32_t parity = tile_no & 1;
            const int r = base + threadIdx.x;
            const bool active = r < S;
            Cand cand = Traits::empty();
            if (active) {
                const int flat = sorted_fast ? (int)(uint32_t)sk_sorted[r] : (int)in_i[S - 1 - r];
                cand = Traits::load(g_boxes, (size_t)flat, flat);
            }
            s_tile[threadIdx.x] = cand;
            __syncwarp();

            NMS_ROW_STAMP(0);    // candidate gathered
            // (a) against boxes kept by earlier tiles
            bool alive = active;
            if (suppressed_by<Traits, kKeptSmem, kTestBatch>(kept, 0, k0, cand, ctx)) alive = false;
            // (b) intra-row matrix: which lower lanes of my row overlap me.  Each unordered pair is
            //     evaluated once: in step t lane l tests lane (l-t) mod 32 and the ballot hands the
            //     result to whichever of the two has the higher rank index.
            uint32_t lower = 0;
            {
                const uint32_t row_alive = __ballot_sync(0xffffffffu, alive);
#pragma unroll 4
                for (int t = 1; t <= 16; ++t) {
                    const int a = (lane - t) & 31, b = (lane + t) & 31;
                    const Cand ob = Traits::load_shared(s_tile + (w << 5) + a);
                    const bool hit = alive && ((row_alive >> a) & 1u) && Traits::suppress(ob, cand, ctx);
                    const uint32_t hits = __ballot_sync(0xffffffffu, hit);
                    if (a < lane && hit) lower |= 1u << a;
                    if (b < lane && ((hits >> b) & 1u)) lower |= 1u << b;
                }
            }
            // (c) rows retire in rank order.  A warp first sleeps until the row kLookAhead before it
            //     has retired (so rows far beyond the cut-off never start), then waits on the mbarrier
            //     of each remaining earlier row in turn (hardware wait, no issue slots burnt) and tests
            //     its candidates against the boxes kept since it last looked; once max_boxes are kept
            //     it stops looking.
            int seen = k0;
            int kc = k0;
            NMS_ROW_STAMP(1);    // intra-row matrix done
            for (int pw = max(0, w - p.look_ahead); pw < w; ++pw) {
                if (pw == w - 1) NMS_ROW_STAMP(2);    // about to wait for the immediate predecessor
                if (!mbar_wait_bounded(&s_turn[pw], parity)) { *s_fault = 1; break; }
                kc = s_kafter[pw];            // published by exactly the row just acquired
                if (*s_kcount >= K) kc = K;   // a later row already reached max_boxes: leave at once
                if (kc >= K) break;
                if (suppressed_by<Traits, kKeptSmem, kTestBatch>(kept, seen, kc, cand, ctx)) alive = false;
                seen = kc;
            }
            NMS_ROW_STAMP(3);    // my turn: all earlier keeps tested
            if (kc < K) {
                uint32_t und = __ballot_sync(0xffffffffu, alive);
                uint32_t keep = 0;
                const bool me0 = alive;
                while (und) {
                    const bool me = me0 && ((und >> lane) & 1u);
                    const uint32_t know = __ballot_sync(0xffffffffu, me && !(lower & keep) && !(lower & und));
                    keep |= know;
                    const uint32_t dnow = __ballot_sync(0xffffffffu, me && (lower & keep));
                    und &= ~(know | dnow);
                }
                const int room = K - kc;
                int nk = __popc(keep);
                if (nk > room) {                       // keep only the first `room` of this row
                    keep &= (1u << __fns(keep, 0, room + 1)) - 1u;
                    nk = room;
                }
                if ((keep >> lane) & 1u) kept[kc + __popc(keep & lanemask_lt())] = cand;
                kc += nk;
            }
            __syncwarp();
            if (lane == 0) {
                s_kafter[w] = kc;
                if (kc > *s_kcount) *s_kcount = kc;       // rows retire in order, so this only grows
                mbar_arrive(&s_turn[w]);                  // release: keeps + counts visible to waiters
            }
            NMS_ROW_STAMP(4);    // retired
            __syncthreads();
            k0 = *s_kcount;
            __syncthreads();
        }
        NMS_STAMP();         // 5: NMS tiles done
        done = S;
        if (k0 >= K || S >= M) break;
    }
    const int kept_n = *s_kcount;

    // ---- epilogue -------------------------------------------------------------------
    if constexpr (kI32) {
        unsigned char *rec = p.det + (size_t)seg * p.det_stride;
        int32_t *hdr = reinterpret_cast<int32_t *>(rec);
        int4 *rb = reinterpret_cast<int4 *>(rec + 16);
        float *rs = reinterpret_cast<float *>(rec + 16 + (size_t)p.det_max_boxes * 16);
        int32_t *ri = reinterpret_cast<int32_t *>(rec + 16 + (size_t)p.det_max_boxes * 20);
        if (threadIdx.x == 0) {
            hdr[0] = *s_fault ? -1 : kept_n;        // -1: a hand-off did not complete within 2 s, results invalid
            hdr[1] = M;
            hdr[2] = *s_ties;
            hdr[3] = S;
        }
        for (int j = threadIdx.x; j < p.det_max_boxes; j += kNmsThreads) {
            if (j < kept_n) {
                const Cand c = kept[j];
                const int flat = Traits::flat(c);
                rb[j] = BoxI32::unpack(reinterpret_cast<const int4 &>(c));
                rs[j] = key_to_score((uint32_t)g_keys[flat]);
                ri[j] = flat;
            } else {
                rb[j] = make_int4(0, 0, 0, 0);
                rs[j] = 0.f;
                ri[j] = 0;
            }
        }
        NMS_STAMP();         // 6: record written
    } else {
        if (threadIdx.x == 0) {
            p.count[0] = *s_fault ? -1 : kept_n;
            p.count[1] = *s_ties;
            p.count[2] = S;
        }
        for (int j = threadIdx.x; j < kept_n; j += kNmsThreads) p.pick[j] = Traits::flat(kept[j]);
    }
}

// ----------------------------------------------------------------------------------
// host side: shared-memory plan + launch
// ----------------------------------------------------------------------------------
struct NmsPlan {
    bool smem_sort;
    bool staged_only;         // hot-path buffers in shared memory, general sort in global scratch
    bool kept_smem;
    size_t smem_bytes;
    size_t ws_stride;
    SortNmsParams p;
};

static int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    return dev;
}

static int max_optin_smem() {
    const int v = device_smem_optin(current_device());
    return v > 0 ? v : 232448;
}

template <typename Cand, typename KeyT>
static NmsPlan make_plan(int N, int max_boxes, int table_entries, bool allow_smem_sort, int smem_limit) {
    NmsPlan pl{};
    SortNmsParams &p = pl.p;
    const int K = max_boxes < N ? max_boxes : N;
    size_t off = kSmemHeader;
    p.sm_off_cnt = (int)off;
    size_t cnt_bytes = (size_t)kNmsWarps * kCntStride * 4;
    size_t tile_bytes = (size_t)kTile * sizeof(Cand);
    off += align_up(cnt_bytes > tile_bytes ? cnt_bytes : tile_bytes, 128);
    p.sm_off_table = (int)off;
    off += align_up((size_t)table_entries * 2, 128);
    pl.kept_smem = K <= kKeptSmemMax;
    p.sm_off_kept = (int)off;
    if (pl.kept_smem) off += align_up((size_t)K * sizeof(Cand) + 64, 128);
    // first sort round: about this many top-scoring candidates (see the kernel)
    p.sel_target = (4 * K > 2048) ? 4 * K : 2048;
    p.look_ahead = kLookAhead;
    {   // tuning knobs (any value is correct)
        const long long la = get_option(kOptNmsLookahead), stg = get_option(kOptNmsSelTarget);
        if (la >= 1 && la <= 32) p.look_ahead = (int)la;
        if (stg >= 32 && stg <= (1 << 20)) p.sel_target = (int)stg;
    }
    // cluster form: ranks resolved through the overlap matrix (about 2.5 K candidates yield K keeps at 0.7)
    p.cluster_ranks = (int)align_up((size_t)(5 * K / 2), 32);
    if (p.cluster_ranks > kTile) p.cluster_ranks = kTile;
    {
        const long long v = get_option(kOptNmsClusterRanks);
        if (v >= 64 && v <= kTile) p.cluster_ranks = ((int)v + 31) & ~31;
    }
    // shared-memory sort: keys ping-pong + uint16 index ping-pong
    size_t cap = align_up((size_t)N, 64);
    size_t sort_bytes = 2 * cap * sizeof(KeyT) + 2 * cap * sizeof(uint16_t);
    pl.smem_sort = allow_smem_sort && N <= 65535 && off + sort_bytes <= (size_t)smem_limit;
    p.sm_off_sort = (int)off;
    p.sort_cap = (int)cap;
    pl.staged_only = false;
    if (pl.smem_sort) {
        // the hot path borrows the ping-pong buffers: keys staged in kA, matrix in kB + iA, sorted words in iB
        p.sm_off_stage = (int)off;
        p.sm_off_L = (int)(off + cap * sizeof(KeyT));
        p.l_bytes = (int)(cap * (sizeof(KeyT) + sizeof(uint16_t)));
        p.sm_off_sk = (int)(off + cap * (2 * sizeof(KeyT) + sizeof(uint16_t)));
        p.sk_cap = (int)(cap * sizeof(uint16_t) / 8);
        off += sort_bytes;
    } else if (allow_smem_sort && sizeof(KeyT) == 4 && pl.kept_smem && N <= 32 * kNmsThreads) {
        // larger panels (12 anchors, 600x800 px): only the hot path's own buffers fit
        const size_t stage_bytes = align_up(cap * sizeof(KeyT), 128);
        const int sk_cap = (p.sel_target * 3 / 2 + 255) & ~255;
        const size_t sk_bytes = (size_t)sk_cap * 8;
        const int rows = p.cluster_ranks;
        const size_t l_bytes = align_up((size_t)rows * (size_t)((rows >> 5) | 1) * 4, 128);
        if (off + stage_bytes + sk_bytes + l_bytes <= (size_t)smem_limit) {
            pl.staged_only = true;
            p.sm_off_stage = (int)off;
            p.sm_off_sk = (int)(off + stage_bytes);
            p.sk_cap = sk_cap;
            p.sm_off_L = (int)(off + stage_bytes + sk_bytes);
            p.l_bytes = (int)l_bytes;
            off += stage_bytes + sk_bytes + l_bytes;
        }
    }
    pl.smem_bytes = off;
    // global scratch (always sized so that either variant can run)
    size_t w = 0;
    p.ws_off_kA = w; w += align_up(cap * sizeof(KeyT), 256);
    p.ws_off_kB = w; w += align_up(cap * sizeof(KeyT), 256);
    p.ws_off_iA = w; w += align_up(cap * 4, 256);
    p.ws_off_iB = w; w += align_up(cap * 4, 256);
    p.ws_off_kept = w; w += align_up((size_t)K * sizeof(Cand) + 64, 256) + 32768 + 256;   // + optional profile stamps
    pl.ws_stride = w;
    p.ws_stride = w;
    return pl;
}

// CTAs per panel in the cluster form: 16 (non-portable opt-in; one GPC of a B200 holds 18-20 SMs) for the
// very few panels that many clusters fit, 8 (the portable maximum) for a few more, otherwise one CTA per
// panel.  RADNET_NMS_CLUSTER_SIZE forces 2, 4, 8 or 16; RADNET_NMS_CLUSTER=0 turns the form off.
static int forced_cluster_size() {
    const long long q = get_option(kOptNmsClusterSize);
    return (q == 2 || q == 4 || q == 8 || q == 16) ? (int)q : 0;
}

// function attributes (dynamic shared memory, non-portable cluster size) are set once per kernel and device
template <typename K>
static int ensure_attrs(K kernel, size_t smem_bytes, bool nonportable) {
    const int dev = current_device();
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), dev, smem_bytes)) return rc;
    if (nonportable) return ensure_nonportable_clusters(reinterpret_cast<const void *>(kernel), dev);
    return RADNET_OK;
}

template <typename K>
static int launch_cluster(K kernel, const NmsPlan &pl, int B, int cs, cudaStream_t st) {
    if (int rc = ensure_attrs(kernel, pl.smem_bytes, cs > 8)) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cs));
    cfg.blockDim = dim3(kNmsThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RADNET_CUDA(cudaLaunchKernelEx(&cfg, kernel, pl.p));
    return check_launch("sort_nms_kernel (cluster)");
}


template <typename K>
static int launch(K kernel, const NmsPlan &pl, int B, cudaStream_t st) {
    if (int rc = ensure_attrs(kernel, pl.smem_bytes, false)) return rc;
    kernel<<<B, kNmsThreads, pl.smem_bytes, st>>>(pl.p);
    return check_launch("sort_nms_kernel");
}

__global__ void make_keys64_kernel(const double *probs, const uint8_t *valid, int M, uint64_t *keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) keys[i] = (valid && !valid[i]) ? 0ull : score_to_key64(probs[i]);
}

// clusters of `cs` CTAs of the hot-path kernel that the device can hold at once (a cluster needs its SMs
// inside one GPC, so this is fewer than SMs / cs); cached per plan size
static int max_active_clusters(const NmsPlan &pl, int cs) {
    static std::mutex mtx;
    static size_t cached_smem[64][2][17] = {{{0}}};
    static int cached_n[64][2][17] = {{{0}}};
    const int dev = current_device() & 63;
    std::lock_guard<std::mutex> lock(mtx);
    size_t *cached_smem_row = cached_smem[dev][pl.staged_only ? 1 : 0];
    int *cached = cached_n[dev][pl.staged_only ? 1 : 0];
    if (cached_smem_row[cs] == pl.smem_bytes) return cached[cs];
    const void *kernel = pl.staged_only
        ? (const void *)sort_nms_kernel<BoxI32, uint32_t, uint32_t, false, true, true, true>
        : (const void *)sort_nms_kernel<BoxI32, uint32_t, uint16_t, true, true, true>;
    int n = 0;
    // through the shared table, so that the granted shared-memory size of a kernel never shrinks
    bool ok = ensure_dynamic_smem(kernel, current_device(), pl.smem_bytes) == RADNET_OK;
    if (ok && cs > 8) ok = ensure_nonportable_clusters(kernel, current_device()) == RADNET_OK;
    if (ok) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)cs);
        cfg.blockDim = dim3(kNmsThreads);
        cfg.dynamicSmemBytes = pl.smem_bytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    cached[cs] = n;
    cached_smem_row[cs] = pl.smem_bytes;
    return n;
}

// cluster size for a launch of B panels; 0 = one CTA per panel
static int choose_cluster(const NmsPlan &pl, int B) {
    if (!(pl.smem_sort || pl.staged_only) || !pl.kept_smem) return 0;
    if (get_option(kOptNmsCluster) == 0) return 0;
    if (const int f = forced_cluster_size()) return B <= max_active_clusters(pl, f) ? f : 0;
    if (B <= max_active_clusters(pl, 16)) return 16;
    if (B <= max_active_clusters(pl, 8)) return 8;
    return 0;
}

}  // namespace radnet

using namespace radnet;

extern "C" size_t radnet_det_record_bytes(int max_boxes) {
    if (max_boxes < 0) return 0;
    return align_up(16 + (size_t)max_boxes * 24, 16);
}

extern "C" size_t radnet_sort_nms_i32_workspace_bytes(int B, int N, int map_h, int map_w, int max_boxes) {
    if (B < 1 || N < 1 || max_boxes < 1) return 0;
    (void)map_h; (void)map_w;
    NmsPlan pl = make_plan<BoxI32::Cand, uint32_t>(N, max_boxes, 1, false, 0);
    return pl.ws_stride * (size_t)B;
}

extern "C" int radnet_sort_nms_i32(const int32_t *boxes_i32, const uint32_t *keys, int B, int N, int map_h,
                                   int map_w, double thr, int max_boxes, void *det, void *ws,
                                   size_t ws_bytes, void *stream) {
    RADNET_CHECK_ARG(boxes_i32 && keys && det && ws, "sort_nms_i32: null pointer");
    RADNET_CHECK_ARG(B >= 1 && N >= 1 && max_boxes >= 1 && map_h >= 1 && map_w >= 1,
                     "sort_nms_i32: bad sizes B=%d N=%d max_boxes=%d", B, N, max_boxes);
    long long umax = 2LL * (map_h - 1) * (map_w - 1);
    if (umax + 1 > kMaxTableEntries || map_h > 16384 || map_w > 16384) {
        set_error("sort_nms_i32: map %dx%d exceeds the exact-table range; use the f64 path", map_h, map_w);
        return RADNET_E_UNSUPPORTED;
    }
    NmsPlan pl = make_plan<BoxI32::Cand, uint32_t>(N, max_boxes, (int)umax + 1, true, max_optin_smem());
    if (pl.smem_bytes > (size_t)max_optin_smem()) {
        set_error("sort_nms_i32: shared memory plan %zu B exceeds the device limit", pl.smem_bytes);
        return RADNET_E_UNSUPPORTED;
    }
    if (ws_bytes < pl.ws_stride * (size_t)B) {
        set_error("sort_nms_i32: workspace %zu < %zu", ws_bytes, pl.ws_stride * (size_t)B);
        return RADNET_E_WORKSPACE;
    }
    SortNmsParams &p = pl.p;
    p.boxes = boxes_i32; p.keys = keys; p.N = N; p.max_boxes = max_boxes; p.thr = thr;
    p.table_entries = (int)umax + 1;
    p.det = reinterpret_cast<unsigned char *>(det);
    p.det_stride = radnet_det_record_bytes(max_boxes);
    p.det_max_boxes = max_boxes;
    p.pick = nullptr; p.count = nullptr;
    p.ws = reinterpret_cast<unsigned char *>(ws);
    cudaStream_t st = (cudaStream_t)stream;
    if (pl.smem_sort) {
        // few panels: a cluster of 8-16 CTAs per panel shares the overlap tests (latency path);
        // many panels: one CTA per panel keeps every SM on its own panel (throughput path)
        if (const int cs = choose_cluster(pl, B)) {
            // a refused cluster launch (nothing has run yet) is not an error: the one-CTA form does the same job
            if (launch_cluster(sort_nms_kernel<BoxI32, uint32_t, uint16_t, true, true, true>, pl, B, cs, st) == RADNET_OK)
                return RADNET_OK;
            cudaGetLastError();
        }
        if (pl.kept_smem) return launch(sort_nms_kernel<BoxI32, uint32_t, uint16_t, true, true>, pl, B, st);
        return launch(sort_nms_kernel<BoxI32, uint32_t, uint16_t, true, false>, pl, B, st);
    }
    if (pl.staged_only) {
        if (const int cs = choose_cluster(pl, B)) {
            if (launch_cluster(sort_nms_kernel<BoxI32, uint32_t, uint32_t, false, true, true, true>, pl, B, cs, st) == RADNET_OK)
                return RADNET_OK;
            cudaGetLastError();
        }
        return launch(sort_nms_kernel<BoxI32, uint32_t, uint32_t, false, true, false, true>, pl, B, st);
    }
    if (pl.kept_smem) return launch(sort_nms_kernel<BoxI32, uint32_t, uint32_t, false, true>, pl, B, st);
    return launch(sort_nms_kernel<BoxI32, uint32_t, uint32_t, false, false>, pl, B, st);
}

extern "C" size_t radnet_nms_f64_workspace_bytes(int M, int max_boxes) {
    if (M < 1 || max_boxes < 1) return 0;
    // +M*8 for the uint64 key image built by the entry point
    NmsPlan pl = make_plan<BoxF64::Cand, uint64_t>(M, max_boxes, 1, false, 0);
    return pl.ws_stride + align_up((size_t)M * 8, 256);
}

extern "C" int radnet_nms_f64(const double *boxes, const double *probs, const uint8_t *valid, int M, double thr,
                              int max_boxes, int32_t *pick, int32_t *count, void *ws, size_t ws_bytes,
                              void *stream) {
    RADNET_CHECK_ARG(boxes && probs && pick && count && ws, "nms_f64: null pointer");
    RADNET_CHECK_ARG(M >= 1 && max_boxes >= 1, "nms_f64: bad sizes M=%d max_boxes=%d", M, max_boxes);
    NmsPlan pl = make_plan<BoxF64::Cand, uint64_t>(M, max_boxes, 1, false, max_optin_smem());
    size_t key_bytes = align_up((size_t)M * 8, 256);
    if (ws_bytes < pl.ws_stride + key_bytes) {
        set_error("nms_f64: workspace %zu < %zu", ws_bytes, pl.ws_stride + key_bytes);
        return RADNET_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t *keys = reinterpret_cast<uint64_t *>(ws);
    make_keys64_kernel<<<(M + 255) / 256, 256, 0, st>>>(probs, valid, M, keys);
    int rc = check_launch("make_keys64_kernel");
    if (rc) return rc;
    SortNmsParams &p = pl.p;
    p.boxes = boxes; p.keys = keys; p.N = M; p.max_boxes = max_boxes; p.thr = thr;
    p.table_entries = 0;
    p.det = nullptr; p.det_stride = 0; p.det_max_boxes = 0;
    p.pick = pick; p.count = count;
    p.ws = reinterpret_cast<unsigned char *>(ws) + key_bytes;
    if (pl.kept_smem) return launch(sort_nms_kernel<BoxF64, uint64_t, uint32_t, false, true>, pl, 1, st);
    return launch(sort_nms_kernel<BoxF64, uint64_t, uint32_t, false, false>, pl, 1, st);
}

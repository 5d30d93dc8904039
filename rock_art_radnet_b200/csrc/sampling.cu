// sampling.cu - f3: the RNG-driven sampling of the training data path, on the device and draw for draw
// identical to the reference's calls into NumPy's legacy global generator:
//
//   radnet_rpn_subsample    the 256-region balancing at the end of calc_region_props
//                           (reference faster_rcnn/utils.py:777-813):
//                           np.random.choice(n, k, replace=False, p=probs), twice
//   radnet_select_samples   get_selected_samples (reference train.py:93-129):
//                           np.random.choice(arr, k, replace=False) / (..., replace=True)
//   radnet_mt19937_seed     np.random.seed(int) -> generator state, per panel
//
// NumPy's RandomState is frozen (NEP 19), so its algorithms are a contract: MT19937, random_sample =
// 53-bit doubles from two 32-bit words, choice(replace=False, p) = rounds of
// cdf.searchsorted(rand(k), 'right') + first-occurrence unique, choice(replace=False) =
// permutation(n)[:k] (Fisher-Yates from the top, masked-rejection random_interval),
// choice(replace=True) = randint(0, n, k) (masked rejection).  A generator state is
// uint32 key[624] + uint32 pos, the layout of np.random.get_state(); the kernels read it, advance it
// exactly as NumPy would and write it back, so host code can hand np.random's state in and out.
//
// The one step that is inherently sequential in NumPy - cdf = np.cumsum(p), a serial float64 chain -
// is replaced by a parallel scan whose values differ from the serial ones by at most eps = 4*n*2^-53;
// a draw x is resolved with it unless x lies within 4*eps of a cdf value next to it, in which case the
// round is redone with the exact serial chain (proof sketch at `resolve_round`).  Option
// sampler_force_exact = 1 always takes the serial chain (tests compare both).
#include "common.cuh"

namespace radnet {

constexpr int kSampThreads = 1024;
constexpr int kMtN = 624, kMtM = 397;

// ------------------------------------------------------------------ MT19937 in shared memory
struct MtShared {
    uint32_t key[kMtN];
    int pos;
};

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7FFFFFFFu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}

// regenerate all 624 words (block-cooperative; three dependency phases, each read / barrier / write)
__device__ void mt_regenerate(MtShared &mt) {
    const int t = threadIdx.x;
    const int lo[3] = {0, kMtN - kMtM, 2 * (kMtN - kMtM)}, hi[3] = {kMtN - kMtM, 2 * (kMtN - kMtM), kMtN};
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
        const int i = lo[ph] + t;
        uint32_t v = 0;
        const bool act = i < hi[ph];
        if (act) v = mt_twist(mt.key[i], mt.key[(i + 1) % kMtN], mt.key[(i + kMtM) % kMtN]);
        __syncthreads();
        if (act) mt.key[i] = v;
        __syncthreads();
    }
    if (t == 0) mt.pos = 0;
    __syncthreads();
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return y;
}

// the next `count` 32-bit outputs of the stream into out[] (block-cooperative)
__device__ void mt_fill(MtShared &mt, uint32_t *out, int count) {
    int done = 0;
    while (done < count) {                       // block-uniform
        if (mt.pos >= kMtN) mt_regenerate(mt);
        const int pos = mt.pos;
        const int take = min(kMtN - pos, count - done);
        for (int i = threadIdx.x; i < take; i += blockDim.x) out[done + i] = mt_temper(mt.key[pos + i]);
        __syncthreads();
        if (threadIdx.x == 0) mt.pos = pos + take;
        __syncthreads();
        done += take;
    }
}

// random_sample: 53-bit double from two consecutive words (a >> 5, b >> 6); every step is exact
__device__ __forceinline__ double mt_double(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

__device__ void mt_load(MtShared &mt, const uint32_t *state) {
    for (int i = threadIdx.x; i < kMtN; i += blockDim.x) mt.key[i] = state[i];
    if (threadIdx.x == 0) mt.pos = (int)min(state[kMtN], (uint32_t)kMtN);
    __syncthreads();
}
__device__ void mt_store(const MtShared &mt, uint32_t *state) {
    __syncthreads();
    for (int i = threadIdx.x; i < kMtN; i += blockDim.x) state[i] = mt.key[i];
    if (threadIdx.x == 0) state[kMtN] = (uint32_t)mt.pos;
}

// ------------------------------------------------------------------ block helpers
// exclusive scan of one int per thread; returns the exclusive prefix, *total = block sum.  Two barriers.
__device__ int block_excl_scan(int v, int *s_warp /*[33]*/, int *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    __syncthreads();                      // s_warp free
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int x = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
        int xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += n;
        }
        s_warp[lane] = xi - x;
        if (lane == 31) s_warp[32] = xi;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[w] + inc - v;
}

// inclusive float64 scan p[0..n) -> cdf[0..n), any association (thread chunks + block scan of the chunk sums)
__device__ void block_scan_f64(const double *p, double *cdf, int n, double *s_part /*[kSampThreads]*/) {
    const int chunk = (n + blockDim.x - 1) / blockDim.x;
    const int lo = min((int)threadIdx.x * chunk, n), hi = min(lo + chunk, n);
    double acc = 0.0;
    for (int i = lo; i < hi; ++i) acc += p[i];
    __syncthreads();
    s_part[threadIdx.x] = acc;
    __syncthreads();
    // Hillis-Steele over the 1024 partial sums in shared memory (10 rounds)
    for (int d = 1; d < (int)blockDim.x; d <<= 1) {
        const double add = threadIdx.x >= (unsigned)d ? s_part[threadIdx.x - d] : 0.0;
        __syncthreads();
        s_part[threadIdx.x] += add;
        __syncthreads();
    }
    double run = threadIdx.x ? s_part[threadIdx.x - 1] : 0.0;
    for (int i = lo; i < hi; ++i) {
        run += p[i];
        cdf[i] = run;
    }
    __syncthreads();
}

// np.cumsum exactly: one serial float64 chain, staged through shared memory tile by tile
__device__ void block_cumsum_serial(const double *p, double *cdf, int n, double *s_tile /*[kSampThreads]*/,
                                    double *s_carry) {
    if (threadIdx.x == 0) *s_carry = 0.0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int len = min((int)blockDim.x, n - base);
        __syncthreads();
        if ((int)threadIdx.x < len) s_tile[threadIdx.x] = p[base + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            double acc = *s_carry;
            for (int i = 0; i < len; ++i) {
                acc = __dadd_rn(acc, s_tile[i]);
                s_tile[i] = acc;
            }
            *s_carry = acc;
        }
        __syncthreads();
        if ((int)threadIdx.x < len) cdf[base + threadIdx.x] = s_tile[threadIdx.x];
    }
    __syncthreads();
}

struct ChoiceScratch {
    double *p;          // [n] probabilities, zeroed as items are found
    double *cdf;        // [n]
    uint32_t *words;    // [2*size] stream words of a round, then the drawn indices
    uint8_t *taken;     // [n]
    double *s_f64;      // shared [kSampThreads]
    int *s_int;         // shared [8]: 0 n_uniq, 1 round count, 2 uncertain flag
    double *s_carry;    // shared [1]
    int force_exact;
};

// first i with cdf[i] > x (np.searchsorted(cdf, x, side='right')); the invariant holds for any array
__device__ __forceinline__ int upper_bound(const double *cdf, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] > x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// One round of RandomState.choice(replace=False, p): m draws against the current p.
//   exact == false: cdf' from the parallel scan, |cdf' - cdf| <= eps = 4*n*2^-53 (both sums carry at most
//   n*2^-53 relative error, as do both totals).  upper_bound returns an adjacent pair cdf'[i-1] <= x < cdf'[i];
//   if x is farther than delta = 4*eps from both, then every cdf'[k] is farther than eps from x (near-monotone:
//   cdf'[k] <= cdf'[i-1] + 2*eps for k < i-1, cdf'[k] >= cdf'[i] - 2*eps for k > i), so cdf[k] > x <=> cdf'[k] > x
//   for all k and the index equals NumPy's.  Otherwise the round is flagged and redone exactly.
__device__ void resolve_round(const ChoiceScratch &c, int n, int m, bool exact) {
    if (exact) block_cumsum_serial(c.p, c.cdf, n, c.s_f64, c.s_carry);
    else block_scan_f64(c.p, c.cdf, n, c.s_f64);
    const double total = c.cdf[n - 1];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) c.cdf[i] = __ddiv_rn(c.cdf[i], total);     // cdf /= cdf[-1]
    __syncthreads();
    const double delta = (double)n * 1.7763568394002505e-15;                                      // n * 2^-49
    bool unsure = false;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const double x = mt_double(c.words[2 * j], c.words[2 * j + 1]);
        const int idx = upper_bound(c.cdf, n, x);
        if (!exact) {
            const double below = idx > 0 ? c.cdf[idx - 1] : 0.0;
            const double above = idx < n ? c.cdf[idx] : 2.0;
            if (x - below <= delta || above - x <= delta || idx >= n) unsure = true;
        }
        c.words[2 * m + j] = (uint32_t)idx;
    }
    if (unsure) atomicOr(&c.s_int[2], 1);
    __syncthreads();
}

// RandomState.choice(n, size, replace=False, p): marks the chosen items in c.taken (block-cooperative).
// c.p holds the probabilities on entry (destroyed); c.words needs 3*size entries.
__device__ void block_choice_p(MtShared &mt, const ChoiceScratch &c, int n, int size) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) c.taken[i] = 0;
    if (threadIdx.x == 0) c.s_int[0] = 0;
    __syncthreads();
    while (true) {
        const int n_uniq = c.s_int[0];
        if (n_uniq >= size) break;                                   // block-uniform
        const int m = size - n_uniq;
        mt_fill(mt, c.words, 2 * m);                                 // x = self.rand(size - n_uniq)
        if (threadIdx.x == 0) c.s_int[2] = 0;
        __syncthreads();
        if (c.force_exact) {
            resolve_round(c, n, m, true);
        } else {
            resolve_round(c, n, m, false);
            if (c.s_int[2]) {                                        // block-uniform (read after the barrier)
                __syncthreads();
                resolve_round(c, n, m, true);
                if (threadIdx.x == 0) c.s_int[1] += 1;               // exact re-runs, reported
            }
        }
        // np.unique: first occurrences of this round join `found`; found items get p = 0 for the next round.
        // Only the SET matters to the caller (the items are switched off), so the draw order is not kept.
        int mine = 0;
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            const uint32_t idx = c.words[2 * m + j];
            uint32_t *word = reinterpret_cast<uint32_t *>(c.taken) + (idx >> 2);
            const uint32_t bit = 1u << (8 * (idx & 3));
            if (!(atomicOr(word, bit) & bit)) {
                c.p[idx] = 0.0;
                ++mine;
            }
        }
        if (mine) atomicAdd(&c.s_int[0], mine);
        __syncthreads();
    }
}

// ------------------------------------------------------------------ K-sub: the 256-region balancing
struct SubsampleParams {
    double *y_cls;        // layout 0: [B][2A][H][W]   layout 1: [B][H][W][2A]
    int B, H, W, A, layout, max_regions;
    uint32_t *states;     // [B][625]
    int32_t *out;         // [B][8] {n_pos returned, n_pos found, n_neg found, status, rounds redone exactly, draws, -, -}
    unsigned char *ws;
    size_t ws_stride;
    int force_exact;
};

enum { kSubOk = 0, kSubKeyError = 1 };

__device__ __forceinline__ size_t cls_offset(const SubsampleParams &p, int ch, int cell) {
    return p.layout ? (size_t)cell * (2 * p.A) + ch : (size_t)ch * p.H * p.W + cell;
}

__global__ void __launch_bounds__(kSampThreads, 1) rpn_subsample_kernel(SubsampleParams p) {
    __shared__ MtShared mt;
    __shared__ double s_f64[kSampThreads];
    __shared__ double s_carry;
    __shared__ int s_int[8];
    __shared__ int s_warp[2][33];
    __shared__ int s_cnt_neg[kMaxAnchors], s_cnt_pos_nokey;
    __shared__ int s_base[2];

    const int b = blockIdx.x;
    const int HW = p.H * p.W, N = p.A * HW;
    double *cls_b = p.y_cls + (size_t)b * 2 * p.A * HW;
    unsigned char *ws = p.ws + (size_t)b * p.ws_stride;
    int32_t *pos_list = reinterpret_cast<int32_t *>(ws);
    int32_t *neg_list = pos_list + N;
    double *pr = reinterpret_cast<double *>(neg_list + N);           // N is padded to even by the launcher
    double *cdf = pr + N;
    uint32_t *words = reinterpret_cast<uint32_t *>(cdf + N);          // [3N]
    uint8_t *taken = reinterpret_cast<uint8_t *>(words + 3 * (size_t)N);
    int32_t *out = p.out + 8 * b;

    mt_load(mt, p.states + (size_t)b * (kMtN + 1));
    for (int i = threadIdx.x; i < kMaxAnchors; i += blockDim.x) s_cnt_neg[i] = 0;
    if (threadIdx.x == 0) { s_base[0] = 0; s_base[1] = 0; s_cnt_pos_nokey = 0; s_int[1] = 0; s_int[3] = 0; }
    __syncthreads();

    // np.where order: channel-first (a, jy, ix) - ordered compaction of the positives and the negatives
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < N; base += kSampThreads) {
        const int f = base + threadIdx.x;
        bool isp = false, isn = false;
        int a = 0;
        if (f < N) {
            a = f / HW;
            const int cell = f - a * HW;
            const double valid = cls_b[cls_offset(p, a, cell)], ov = cls_b[cls_offset(p, p.A + a, cell)];
            isp = (ov == 1.0) && (valid == 1.0);                     // utils.py:777
            isn = (ov == 0.0) && (valid == 1.0);                     // utils.py:778
        }
        const unsigned mp = __ballot_sync(0xffffffffu, isp), mn = __ballot_sync(0xffffffffu, isn);
        if (lane == 0) { s_warp[0][w] = __popc(mp); s_warp[1][w] = __popc(mn); }
        __syncthreads();
        if (w < 2) {
            const int v = s_warp[w][lane];
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += n;
            }
            s_warp[w][lane] = inc - v;
            if (lane == 31) s_warp[w][32] = inc;
        }
        __syncthreads();
        if (isp) pos_list[s_base[0] + s_warp[0][w] + __popc(mp & lanemask_lt())] = f;
        if (isn) {
            neg_list[s_base[1] + s_warp[1][w] + __popc(mn & lanemask_lt())] = f;
            atomicAdd(&s_cnt_neg[a], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_base[0] += s_warp[0][32]; s_base[1] += s_warp[1][32]; }
        __syncthreads();
    }
    int n_pos = s_base[0];
    const int n_neg = s_base[1];
    const int half = p.max_regions / 2;                              // int(max_n_regions / 2)
    ChoiceScratch c{pr, cdf, words, taken, s_f64, s_int, &s_carry, p.force_exact};
    int draws = 0;
    int status = kSubOk;

    if (n_pos > half) {                                              // utils.py:785 (n_pos > max_n_regions/2)
        // probs = [pos_probs[l] / pos_counts[l] for l in pos_locs[0]] with BOTH dicts keyed by the channels of
        // the NEGATIVES (utils.py:787-793): a positive whose channel has no negative raises KeyError
        for (int i = threadIdx.x; i < n_pos; i += blockDim.x) {
            const int cnt = s_cnt_neg[pos_list[i] / HW];
            if (cnt == 0) atomicAdd(&s_cnt_pos_nokey, 1);
            else pr[i] = __ddiv_rn(__ddiv_rn((double)cnt, (double)n_pos), (double)cnt);
        }
        __syncthreads();
        if (s_cnt_pos_nokey) {
            status = kSubKeyError;
        } else {
            block_choice_p(mt, c, n_pos, n_pos - half);              // utils.py:797
            for (int i = threadIdx.x; i < n_pos; i += blockDim.x)
                if (taken[i]) {
                    const int f = pos_list[i], a = f / HW;
                    cls_b[cls_offset(p, a, f - a * HW)] = 0.0;       // utils.py:798
                }
            draws += n_pos - half;
            n_pos = half;
        }
    }
    if (status == kSubOk && n_neg + n_pos > p.max_regions) {         // utils.py:802
        for (int i = threadIdx.x; i < n_neg; i += blockDim.x) {
            const int cnt = s_cnt_neg[neg_list[i] / HW];
            pr[i] = __ddiv_rn(__ddiv_rn((double)cnt, (double)n_neg), (double)cnt);
        }
        __syncthreads();
        block_choice_p(mt, c, n_neg, n_neg - n_pos);                 // utils.py:812
        for (int i = threadIdx.x; i < n_neg; i += blockDim.x)
            if (taken[i]) {
                const int f = neg_list[i], a = f / HW;
                cls_b[cls_offset(p, a, f - a * HW)] = 0.0;           // utils.py:813
            }
        draws += n_neg - n_pos;
    }
    mt_store(mt, p.states + (size_t)b * (kMtN + 1));
    if (threadIdx.x == 0) {
        out[0] = n_pos; out[1] = s_base[0]; out[2] = n_neg; out[3] = status;
        out[4] = s_int[1]; out[5] = draws; out[6] = 0; out[7] = 0;
    }
}

// ------------------------------------------------------------------ get_selected_samples (train.py:93-129)
struct SelectParams {
    const int32_t *y_class;   // [B][R][n_cls] one-hot rows of calc_iou
    const int32_t *count;     // [B] rows used (NULL = R)
    int B, R, n_cls, n_rois;
    uint32_t *states;         // [B][625]
    int32_t *sel;             // [B][n_rois] selected row indices, positives first
    int32_t *out;             // [B][4] {n_selected, n_pos, n_neg, status (0 ok, 2 = the reference raises ValueError)}
};

constexpr int kSelMaxRows = 2048;
enum { kSelAll = 0, kSelPerm = 1, kSelPermDiscard = 2, kSelRandint = 3 };
enum { kSelOk = 0, kSelValueError = 2 };

// Thread 0 replays the reference's draws as a small program of stages over a buffer of tempered words;
// whenever the buffer runs dry the whole block regenerates the generator (block-uniform loop).  Draw kinds:
// masked-rejection random_interval(i) for the Fisher-Yates steps of permutation(n) (choice(replace=False) =
// pop[permutation(n)[:k]]) and for randint(0, n) (choice(replace=True)).
__global__ void __launch_bounds__(256, 1) select_samples_kernel(SelectParams p) {
    __shared__ MtShared mt;
    __shared__ uint32_t s_words[kMtN];
    __shared__ int s_have, s_used, s_done;
    __shared__ int s_pos[kSelMaxRows], s_neg[kSelMaxRows], s_perm[kSelMaxRows];
    __shared__ int s_npos, s_nneg;
    __shared__ int s_prog[4][4];          // stage: {population (0 pos, 1 neg), n, k, mode}
    __shared__ int s_nstage, s_stage, s_i, s_written, s_status;

    const int b = blockIdx.x;
    const int R = p.count ? min(max(p.count[b], 0), p.R) : p.R;
    const int32_t *yc = p.y_class + (size_t)b * p.R * p.n_cls;
    int32_t *sel = p.sel + (size_t)b * p.n_rois;
    const int n_rois = p.n_rois;
    mt_load(mt, p.states + (size_t)b * (kMtN + 1));
    if (threadIdx.x == 0) {
        // np.where(Y1[0,:,-1] == 1) / == 0, ascending row order (train.py:96-97); rows are few: serial
        int np_ = 0, nn = 0;
        for (int r = 0; r < R; ++r) {
            const int last = yc[(size_t)r * p.n_cls + p.n_cls - 1];
            if (last == 1) s_neg[nn++] = r;
            else if (last == 0) s_pos[np_++] = r;
        }
        s_npos = np_; s_nneg = nn;
        s_have = 0; s_used = 0; s_done = 0; s_stage = 0; s_i = -1; s_written = 0; s_status = kSelOk;
        // the program (train.py:107-127)
        int ns = 0;
        const int half = n_rois / 2;
        const int first = np_ < half ? np_ : half;                      // len(selected_pos_samples)
        if (nn > 0) {
            s_prog[ns][0] = 0; s_prog[ns][1] = np_; s_prog[ns][2] = first; s_prog[ns][3] = np_ < half ? kSelAll : kSelPerm; ++ns;
            const int k = n_rois - first;
            // replace=False raises before any draw when the population is too small; the reference then
            // repeats the call with replace=True (train.py:115-118)
            s_prog[ns][0] = 1; s_prog[ns][1] = nn; s_prog[ns][2] = k; s_prog[ns][3] = k > nn ? kSelRandint : kSelPerm; ++ns;
        } else {
            // the first selection was already drawn (and is thrown away) before the branch (train.py:108-111)
            if (np_ >= half) { s_prog[ns][0] = 0; s_prog[ns][1] = np_; s_prog[ns][2] = half; s_prog[ns][3] = kSelPermDiscard; ++ns; }
            s_prog[ns][0] = 0; s_prog[ns][1] = np_; s_prog[ns][2] = np_; s_prog[ns][3] = kSelPerm; ++ns;      // train.py:124
            const int k = n_rois - np_;
            if (k < 0 || (np_ == 0 && k > 0)) s_status = kSelValueError;     // negative size / empty population
            else { s_prog[ns][0] = 0; s_prog[ns][1] = np_; s_prog[ns][2] = k; s_prog[ns][3] = kSelRandint; ++ns; }   // train.py:125
        }
        s_nstage = ns;
    }
    const int n_stage_max = 4;
    (void)n_stage_max;
    while (true) {
        __syncthreads();
        const bool done = s_done != 0, dry = s_used >= s_have;
        __syncthreads();
        if (done) break;
        if (dry) {                                                    // block-uniform: refill the word buffer
            if (mt.pos >= kMtN) mt_regenerate(mt);
            for (int i = threadIdx.x; i < kMtN; i += blockDim.x) s_words[i] = mt_temper(mt.key[i]);
            __syncthreads();
            if (threadIdx.x == 0) { s_used = mt.pos; s_have = kMtN; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            int used = s_used, stage = s_stage, i = s_i, written = s_written;
            const int have = s_have, n_stage = s_nstage;
            bool starved = false;
            while (!starved && stage < n_stage) {
                const int *pop = s_prog[stage][0] ? s_neg : s_pos;
                const int n = s_prog[stage][1], k = s_prog[stage][2], mode = s_prog[stage][3];
                if (mode == kSelAll) {                                // every positive, no draw (train.py:108)
                    for (int j = 0; j < n && written < n_rois; ++j) sel[written++] = pop[j];
                } else if (mode == kSelRandint) {
                    // randint(0, n, k): masked rejection per draw; a population of one consumes no word
                    if (i == -1) i = 0;
                    const uint32_t mx = (uint32_t)(n - 1);
                    uint32_t mask = mx;
                    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
                    while (i < k) {
                        uint32_t v = 0;
                        bool got = mx == 0;
                        while (!got && used < have) {
                            v = s_words[used++] & mask;
                            if (v <= mx) got = true;
                        }
                        if (!got) { starved = true; break; }
                        if (written < n_rois) sel[written++] = pop[v];
                        ++i;
                    }
                    if (starved) break;
                } else if (k > 0 || mode == kSelPermDiscard) {
                    // permutation(n): Fisher-Yates from the top over arange(n), every step one random_interval(i)
                    if (i == -1) {
                        for (int j = 0; j < n; ++j) s_perm[j] = j;
                        i = n - 1;
                    }
                    while (i >= 1) {
                        uint32_t mask = (uint32_t)i;
                        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
                        bool got = false;
                        uint32_t v = 0;
                        while (used < have) {
                            v = s_words[used++] & mask;
                            if (v <= (uint32_t)i) { got = true; break; }
                        }
                        if (!got) { starved = true; break; }
                        const int t = s_perm[i]; s_perm[i] = s_perm[v]; s_perm[v] = t;
                        --i;
                    }
                    if (starved) break;
                    if (mode == kSelPerm)
                        for (int j = 0; j < k && written < n_rois; ++j) sel[written++] = pop[s_perm[j]];
                }
                ++stage;
                i = -1;
            }
            s_used = used; s_stage = stage; s_i = i; s_written = written;
            mt.pos = used;
            if (stage >= n_stage) s_done = 1;
        }
    }
    __syncthreads();
    mt_store(mt, p.states + (size_t)b * (kMtN + 1));
    if (threadIdx.x == 0) {
        p.out[4 * b] = s_written; p.out[4 * b + 1] = s_npos; p.out[4 * b + 2] = s_nneg; p.out[4 * b + 3] = s_status;
    }
}

__global__ void mt19937_seed_kernel(const uint32_t *seeds, int B, uint32_t *states) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    uint32_t *k = states + (size_t)b * (kMtN + 1);
    uint32_t prev = seeds[b];
    k[0] = prev;
    for (int i = 1; i < kMtN; ++i) {                                   // init_genrand
        prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i;
        k[i] = prev;
    }
    k[kMtN] = kMtN;
}

}  // namespace radnet

using namespace radnet;

static size_t subsample_ws_stride(int N) {
    const size_t n = ((size_t)N + 1) & ~(size_t)1;
    return align_up(n * (4 + 4 + 8 + 8 + 12) + align_up(n, 16) + 64, 256);
}

extern "C" size_t radnet_rpn_subsample_workspace_bytes(int B, int H, int W, int A) {
    if (B < 1 || H < 1 || W < 1 || A < 1) return 0;
    return (size_t)B * subsample_ws_stride(A * H * W);
}

extern "C" int radnet_rpn_subsample(double *y_rpn_cls, int B, int H, int W, int A, int layout, int max_regions,
                                    uint32_t *rng_states, int32_t *out, void *ws, size_t ws_bytes, void *stream) {
    RADNET_CHECK_ARG(y_rpn_cls && rng_states && out && ws, "rpn_subsample: null pointer");
    RADNET_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && A >= 1 && A <= kMaxAnchors && max_regions >= 2,
                     "rpn_subsample: bad sizes B=%d H=%d W=%d A=%d max_regions=%d", B, H, W, A, max_regions);
    RADNET_CHECK_ARG((long long)A * H * W < (1LL << 28), "rpn_subsample: too many anchors");
    RADNET_CHECK_ARG(layout == RADNET_TARGETS_CHANNEL_FIRST || layout == RADNET_TARGETS_NHWC, "rpn_subsample: bad layout");
    const size_t need = radnet_rpn_subsample_workspace_bytes(B, H, W, A);
    if (ws_bytes < need) {
        set_error("rpn_subsample: workspace %zu < %zu", ws_bytes, need);
        return RADNET_E_WORKSPACE;
    }
    SubsampleParams p{};
    p.y_cls = y_rpn_cls; p.B = B; p.H = H; p.W = W; p.A = A; p.layout = layout; p.max_regions = max_regions;
    p.states = rng_states; p.out = out;
    p.ws = reinterpret_cast<unsigned char *>(ws);
    p.ws_stride = subsample_ws_stride(A * H * W);
    p.force_exact = get_option(kOptSamplerForceExact) == 1;
    rpn_subsample_kernel<<<B, kSampThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch("rpn_subsample_kernel");
}

extern "C" int radnet_select_samples(const int32_t *y_class, const int32_t *count, int B, int R, int n_cls, int n_rois,
                                     uint32_t *rng_states, int32_t *sel, int32_t *out, void *stream) {
    RADNET_CHECK_ARG(y_class && rng_states && sel && out, "select_samples: null pointer");
    RADNET_CHECK_ARG(B >= 1 && R >= 1 && R <= kSelMaxRows && n_cls >= 2 && n_rois >= 1,
                     "select_samples: bad sizes B=%d R=%d n_cls=%d n_rois=%d (R <= %d)", B, R, n_cls, n_rois, kSelMaxRows);
    SelectParams p{};
    p.y_class = y_class; p.count = count; p.B = B; p.R = R; p.n_cls = n_cls; p.n_rois = n_rois;
    p.states = rng_states; p.sel = sel; p.out = out;
    select_samples_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("select_samples_kernel");
}

extern "C" int radnet_mt19937_seed(const uint32_t *seeds, int B, uint32_t *rng_states, void *stream) {
    RADNET_CHECK_ARG(seeds && rng_states && B >= 1, "mt19937_seed: bad arguments");
    mt19937_seed_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(seeds, B, rng_states);
    return check_launch("mt19937_seed_kernel");
}

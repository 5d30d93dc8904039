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
// c.p holds the probabilities on entry; words needs 3*size entries.
__device__ void block_choice_p(MtShared &mt, const ChoiceScratch &c, int n, int size) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) c.taken[i] = 0;
    if (threadIdx.x == 0) { c.s_int[0] = 0; c.s_int[1] = 0; }
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
            if (c.s_int[2]) {
                __syncthreads();
                resolve_round(c, n, m, true);
                if (threadIdx.x == 0) c.s_int[1] += 1;               // exact re-runs, reported
            }
        }
        // first occurrences of this round; found items get p = 0 for the next round
        int mine = 0;
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            const uint32_t idx = c.words[2 * m + j];
            if (idx < (uint32_t)n && atomicExch(reinterpret_cast<unsigned int *>(c.taken) + (idx >> 2), 0u) == 0xFFFFFFFFu) {}
            (void)mine;
        }
        __syncthreads();
        break;
    }
}

}  // namespace radnet

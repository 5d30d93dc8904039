"""TEST INFRASTRUCTURE ONLY - restatement of the legacy NumPy RNG calls on the training side of the
RADNet hot path, so that the CUDA sampler can be checked draw for draw.

The reference samples with the GLOBAL legacy generator (`np.random.seed(SEED)`, train.py:134):
  * calc_region_props: `np.random.choice(n, k, replace=False, p=probs)`       (utils.py:797, 812)
  * get_selected_samples: `np.random.choice(arr, k, replace=False)` and
    `np.random.choice(arr, k, replace=True)`                                  (train.py:111-127)
NumPy's `RandomState` is frozen by NEP 19 (stream compatibility), so its algorithms are a published
contract: MT19937 (Matsumoto & Nishimura 1998), `random_sample` = 53-bit doubles from two 32-bit words,
`choice(replace=False, p)` = rounds of `cdf.searchsorted(rand(k))` + first-occurrence `unique`,
`choice(replace=False)` = `permutation(n)[:k]` (Fisher-Yates from the top with masked-rejection
`random_interval`), `choice(replace=True)` = `randint(0, n, k)` (masked rejection on 32-bit words).

Parity pinning: PINNED against NumPy itself - `tests/test_oracle_rng.py` runs every function here
against `np.random.RandomState` with the same state (values AND the state left behind), and the f3
goldens under tests/golden/ come from the unmodified reference functions run with `np.random.seed`.
"""
import numpy as np

N, M = 624, 397
UPPER, LOWER, MATRIX_A = 0x80000000, 0x7FFFFFFF, 0x9908B0DF


class MT19937:
    """Explicit MT19937 with NumPy's state layout: key[624] uint32 + pos (pos == 624: regenerate first)."""

    def __init__(self, key, pos):
        self.key = [int(v) for v in key]
        self.pos = int(pos)

    @classmethod
    def from_numpy_state(cls, state):
        name, key, pos = state[0], state[1], state[2]
        assert name == "MT19937"
        return cls(key, pos)

    @classmethod
    def from_seed(cls, seed):
        """init_genrand(seed) as `np.random.seed(int)` does for a 32-bit integer seed."""
        key = [0] * N
        key[0] = seed & 0xFFFFFFFF
        for i in range(1, N):
            key[i] = (1812433253 * (key[i - 1] ^ (key[i - 1] >> 30)) + i) & 0xFFFFFFFF
        return cls(key, N)

    def numpy_state(self):
        return ("MT19937", np.array(self.key, dtype=np.uint32), self.pos, 0, 0.0)

    def _regenerate(self):
        k = self.key
        for i in range(N):
            y = (k[i] & UPPER) | (k[(i + 1) % N] & LOWER)
            k[i] = k[(i + M) % N] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
        self.pos = 0

    def next_u32(self):
        if self.pos >= N:
            self._regenerate()
        y = self.key[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def next_double(self):
        a, b = self.next_u32() >> 5, self.next_u32() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0

    def random_sample(self, n):
        return np.array([self.next_double() for _ in range(n)], dtype=np.float64)

    def interval(self, mx):
        """random_interval: uniform integer in [0, mx], masked rejection on 32-bit words (mx < 2**32)."""
        if mx == 0:
            return 0
        mask = mx
        for s in (1, 2, 4, 8, 16, 32):
            mask |= mask >> s
        assert mx <= 0xFFFFFFFF
        while True:
            v = self.next_u32() & mask
            if v <= mx:
                return v


def choice_noreplace_p(rng, n, size, p):
    """RandomState.choice(n, size, replace=False, p=p) -> int64 indices in draw order."""
    p = np.array(p, dtype=np.float64)
    assert p.shape == (n,)
    if np.count_nonzero(p > 0) < size:
        raise ValueError("Fewer non-zero entries in p than size")
    found = np.zeros(size, dtype=np.int64)
    n_uniq = 0
    while n_uniq < size:
        x = rng.random_sample(size - n_uniq)
        if n_uniq > 0:
            p[found[0:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        new = cdf.searchsorted(x, side='right')
        _, first = np.unique(new, return_index=True)
        first.sort()
        new = new.take(first)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return found


def permutation(rng, n):
    arr = np.arange(n)
    for i in range(n - 1, 0, -1):
        j = rng.interval(i)
        arr[i], arr[j] = arr[j], arr[i]
    return arr


def choice_noreplace(rng, pop, size):
    """RandomState.choice(pop, size, replace=False) for a 1-D array `pop` (or int)."""
    pop = np.arange(pop) if np.isscalar(pop) else np.asarray(pop)
    if size > len(pop):
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    return pop[permutation(rng, len(pop))[:size]]


def choice_replace(rng, pop, size):
    """RandomState.choice(pop, size, replace=True) = pop[randint(0, len(pop), size)]."""
    pop = np.arange(pop) if np.isscalar(pop) else np.asarray(pop)
    if size < 0:
        raise ValueError("negative dimensions are not allowed")
    if len(pop) == 0 and size > 0:
        raise ValueError("'a' cannot be empty unless no samples are taken")
    idx = np.array([rng.interval(len(pop) - 1) for _ in range(size)], dtype=np.int64)
    return pop[idx]


# ------------------------------------------------------------------ the two reference call sites
def subsample_regions(rng, valid_cf, overlap_cf, max_n_regions=256):
    """utils.py:777-813 on channel-first (A,H,W) arrays, in place, with an explicit generator.
    Returns n_pos.  Raises KeyError like the reference when a positive's anchor channel has no negative."""
    pos = np.where(np.logical_and(overlap_cf == 1, valid_cf == 1))
    neg = np.where(np.logical_and(overlap_cf == 0, valid_cf == 1))
    n_pos, n_neg = len(pos[0]), len(neg[0])
    half = int(max_n_regions / 2)
    ids, counts = np.unique(neg[0], return_counts=True)
    if n_pos > max_n_regions / 2:
        share = dict(zip(ids, counts / n_pos))
        size = dict(zip(ids, counts))
        probs = [share[c] / size[c] for c in pos[0]]
        drop = choice_noreplace_p(rng, n_pos, n_pos - half, probs)
        valid_cf[pos[0][drop], pos[1][drop], pos[2][drop]] = 0
        n_pos = half
    if n_neg + n_pos > max_n_regions:
        share = dict(zip(ids, counts / n_neg))
        size = dict(zip(ids, counts))
        probs = [share[c] / size[c] for c in neg[0]]
        drop = choice_noreplace_p(rng, n_neg, n_neg - n_pos, probs)
        valid_cf[neg[0][drop], neg[1][drop], neg[2][drop]] = 0
    return n_pos


def get_selected_samples(rng, Y1, n_rois):
    """train.py:93-129 with an explicit generator: (selected row indices, n_pos)."""
    neg = np.where(Y1[0, :, -1] == 1)[0]
    pos = np.where(Y1[0, :, -1] == 0)[0]
    if len(pos) < n_rois // 2:
        sel_pos = pos.tolist()
    else:
        sel_pos = choice_noreplace(rng, pos, n_rois // 2).tolist()
    if len(neg) > 0:
        k = n_rois - len(sel_pos)
        if k <= len(neg):
            sel_neg = choice_noreplace(rng, neg, k).tolist()
        else:       # the reference catches the ValueError of replace=False, which is raised before any draw
            sel_neg = choice_replace(rng, neg, k).tolist()
        return sel_pos + sel_neg, len(pos)
    sel_pos = choice_noreplace(rng, pos, len(pos)).tolist()
    sel_pos += choice_replace(rng, pos, n_rois - len(sel_pos)).tolist()
    return sel_pos, len(pos)

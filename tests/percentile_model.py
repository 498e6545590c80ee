"""numpy model of what op_percentile_apply (percentile.c:392-751) leaves in the chromosome vectors --
TEST INFRASTRUCTURE (pinned to the unmodified reference by tests/test_oracle_percentile_state.py; the GPU
tests hold gdsp_percentile_collect and the host executor to it).

collect (percentile.c:547-580), closed form: with the genome concatenated in
chromsSorted order, q_0 < q_1 < ... the qualifying positions (every `stride`-th chromosome coordinate with
min <= v <= max) and n their count:  A'[j] = A[q_j] for j < n;  for q_j >= n: A'[q_j] = g(j), g(j) = A[j] if
position j does not qualify, else g(rank(j));  every other position keeps its value.
"""
import numpy as np


def qualifying_mask(lengths, cat, stride, mn, mx):
    q = np.zeros(cat.size, bool)
    pos = 0
    for n in lengths:
        idx = np.arange(0, n, stride) + pos
        v = cat[idx]
        q[idx] = ~(v < mn) & ~(v > mx)
        pos += n
    return q


def collect(lengths, cat, stride=1, mn=-np.inf, mx=np.inf):
    """-> (state after the collect pass, number of qualifying samples)"""
    q = qualifying_mask(lengths, cat, stride, mn, mx)
    qpos = np.flatnonzero(q)
    n = qpos.size
    rank = np.cumsum(q) - q                       # rank[p] = number of qualifying positions before p
    out = cat.copy()
    out[:n] = cat[qpos]
    tail = qpos[qpos >= n]
    j = rank[tail]
    while True:                                   # chase g(j) while position j itself qualifies
        again = q[j]
        if not again.any():
            break
        j = np.where(again, rank[j], j)
    out[tail] = cat[j]
    return out, n


def key_sort(a):
    """ascending by value; the reference's qsort comparator leaves the order of equal values (+0.0/-0.0)
    unspecified, so the model is only used on signals without signed zeros"""
    return np.sort(a)


def post_state(lengths, cat, last_rank, stride=1, mn=-np.inf, mx=np.inf):
    """the vectors after op_percentile_apply when the last reported rank is `last_rank`"""
    state, n = collect(lengths, cat, stride, mn, mx)
    if n == 0:
        return state, 0
    starts = np.concatenate([[0], np.cumsum(lengths)])
    last = int(np.searchsorted(starts, n, "left")) - 1          # chromosome holding the n-th sample (percentile.c:582-583)
    front = [(int(starts[c]), int(min(starts[c + 1], n))) for c in range(last + 1)]
    for a, b in front:
        state[a:b] = key_sort(state[a:b])
    K = last
    acc = 0
    for c, (a, b) in enumerate(front):
        acc += b - a
        if last_rank < acc:
            K = c
            break
    for c in range(K + 1):                        # the bubble passes, percentile.c:623-651
        a, b = front[c]
        for d in range(c + 1, last + 1):
            e, f = front[d]
            both = key_sort(np.concatenate([state[a:b], state[e:f]]))
            state[a:b] = both[:b - a]
            state[e:f] = both[b - a:]
    return state, n

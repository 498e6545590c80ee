"""CPU model of the accumulator schedule of k_smooth_sym (genodsp_b200/csrc/gdsp_smooth_sym.cu) against the direct
ascending-tap fold of the reference (sum.c:651-664): every product w[k]*in[j] of a symmetric tap pair is computed
once and added to a "young" output (at tap k) and an "old" one (at tap W-1-k); outputs move one tap per step, take
the centre tap between the two sides, and leave finished.  T = 1, K = (W-1)/2 is the kernel as shipped (one thread
per strip); T > 1 deals the pairs out to T lanes that hand the accumulators on (the shuffle design that was measured
and dropped, profiles/r2_smooth_sym.md) and pads with T*K - (W-1)/2 zero taps.  Bit-compared, not within a tolerance."""
import numpy as np, sys

def direct(v, w):
    W = len(w); h = (W - 1) // 2; n = len(v)
    out = np.zeros(n)
    for x in range(n):
        s = 0.0
        for k in range(W):
            j = x - h + k
            if 0 <= j < n:
                s = s + w[k] * v[j]
        out[x] = s
    return out

def model(v, w, T, K, x0, x1):
    W = len(w); P = (W - 1) // 2
    hp = T * K; D = hp - P
    assert D >= 0
    wp = np.zeros(hp + 1)
    wp[D:hp] = w[:P]; wp[hp] = w[P]
    n_in = len(v)
    a = np.zeros((T, K)); b = np.zeros((T, K))
    exit_low = np.zeros(T); exit_high = np.zeros(T); c_prev = np.zeros(T)
    out = {}
    total = (x1 - x0) + 2 * hp
    nblocks = (total + K - 1) // K
    for n0 in range(0, nblocks * K, K):
        for phi in range(K):
            n = n0 + phi
            j = x0 - hp + n
            vv = v[j] if 0 <= j < n_in else 0.0
            inc_low = np.zeros(T); inc_high = np.zeros(T); cnew = np.zeros(T)
            for s in range(T):
                inc_low[s] = exit_low[s - 1] if s > 0 else 0.0
                cnew[s] = exit_low[s] + wp[hp] * vv
                inc_high[s] = exit_high[s + 1] if s < T - 1 else c_prev[s]
            c_prev = cnew
            for s in range(T):
                a[s, phi] = inc_low[s]; b[s, phi] = inc_high[s]
                for i in range(K):
                    p = wp[s * K + i] * vv
                    a[s, (phi - i) % K] = a[s, (phi - i) % K] + p
                    b[s, (phi + 1 + i) % K] = b[s, (phi + 1 + i) % K] + p
                exit_low[s] = a[s, (phi + 1) % K]; exit_high[s] = b[s, (phi + 1) % K]
            if n >= 2 * hp and n < total:
                out[x0 - 2 * hp + n] = exit_high[0]
    return out

def hann(W):
    P = (W - 1) // 2
    w = np.zeros(W)
    for k in range(P + 1):
        x = (k + 1) / (W + 1)
        w[k] = w[W - 1 - k] = (1 - np.cos(2 * np.pi * x)) / 2
    return w / w.sum()


def check(W, T, K, n=300, seed=1):
    rng = np.random.default_rng(seed)
    w = hann(W)
    v = rng.normal(size=n) * 10.0 ** rng.integers(-3, 4, n)
    want = direct(v, w)
    for (x0, x1) in [(0, n), (0, 117), (117, n), (50, 51)]:
        got = model(v, w, T, K, x0, x1)
        assert sorted(got) == list(range(x0, x1)), (W, T, K, x0, x1)
        bad = [x for x in range(x0, x1) if got[x].tobytes() != want[x].tobytes()]
        assert not bad, (W, T, K, x0, x1, bad[:5])


if __name__ == "__main__":
    for (W, T, K) in [(3, 1, 1), (5, 1, 2), (11, 1, 5), (31, 1, 15), (101, 1, 50), (7, 2, 2), (11, 2, 3), (31, 4, 4), (101, 2, 25), (21, 3, 4)]:
        check(W, T, K)
        print("ok", W, T, K)

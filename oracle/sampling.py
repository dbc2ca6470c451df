"""Philox4x32-10 joint-space sampler (TEST INFRASTRUCTURE): numpy restatement of
mmmp_b200/csrc/sample.cu, which replaces np.random.uniform + np.round(.,2) of
planners/prm/pybullet/decoupled/decoupled_prm.py:310-318.  The reference uses
numpy's global Mersenne Twister stream, which a counter-based batched sampler
cannot follow; what is pinned here is (a) the arithmetic lo + (hi-lo)*u and the
np.round(x, 2) = rint(x*100)/100 rounding, (b) bit-exact agreement of the CUDA
kernel with this file."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def sample_uniform(lo, hi, n, seed, first_index=0, round_dp=2):
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    D = lo.shape[0]
    ctr = np.uint64(first_index) + np.arange(n, dtype=np.uint64)
    out = np.empty((n, D))
    for jp in range((D + 1) // 2):
        r = philox4x32_10((ctr & MASK).astype(np.uint32), (ctr >> np.uint64(32)).astype(np.uint32),
                          np.full(n, jp, np.uint32), np.zeros(n, np.uint32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for h in range(2):
            j = 2 * jp + h
            if j >= D:
                break
            u = ((r[2 * h] >> np.uint32(5)).astype(np.float64) * 67108864.0 +
                 (r[2 * h + 1] >> np.uint32(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)
            out[:, j] = lo[j] + (hi[j] - lo[j]) * u
    if round_dp >= 0:
        s = 10.0 ** round_dp
        out = np.rint(out * s) / s
    return out

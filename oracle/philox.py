"""numpy restatement of Philox4x32-10 and of the chain's base sample (TEST INFRASTRUCTURE ONLY).
Matches memflow_b200/csrc/philox.cuh bit for bit."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n2 = hi0 ^ c3 ^ k1
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u01_from_bits(bits):
    """(k + 0.5) / 2^23 from the top 23 bits: exact in float32 and strictly inside (0, 1) (philox.cuh)."""
    return ((np.asarray(bits, dtype=np.uint32) >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)


def base_sample(n_events, F, bound, seed, event_offset=0):
    """z[e, f] in (-bound, bound), float32: counter = (global event, f >> 2), key = seed."""
    gev = np.arange(n_events, dtype=np.uint64) + np.uint64(event_offset)
    lo = (gev & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (gev >> np.uint64(32)).astype(np.uint32)
    z = np.empty((n_events, F), dtype=np.float32)
    for blk in range((F + 3) // 4):
        r = philox4x32_10(lo, hi, np.full(n_events, blk, np.uint32), np.zeros(n_events, np.uint32),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for j in range(4):
            f = 4 * blk + j
            if f < F:
                z[:, f] = np.float32(bound) * (np.float32(2.0) * u01_from_bits(r[j]) - np.float32(1.0))
    return z

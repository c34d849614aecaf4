"""ORACLE (test infrastructure only): Philox4x32-10 in numpy, the counter-based
generator of csrc/philox.cuh (Salmon et al. 2011, "Parallel random numbers: as
easy as 1, 2, 3").  Known-answer vectors from the Random123 distribution are
checked in tests/test_oracle.py."""

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c = [np.asarray(v, dtype=np.uint32).copy() for v in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c


def rows_counter(seed, rows, block, stream):
    """Counter/key layout of csrc/philox.cuh: ctr = (row_lo, row_hi, block, stream)."""
    rows = np.asarray(rows, dtype=np.uint64)
    return philox4x32_10(
        (rows & np.uint64(0xFFFFFFFF)).astype(np.uint32),
        (rows >> np.uint64(32)).astype(np.uint32),
        np.uint32(block), np.uint32(stream),
        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF,
    )


def accept_uniform(seed, rows):
    """u of the rejection step: stream 1, block 0, first word; (r + 0.5) * 2^-32."""
    r = rows_counter(seed, rows, 0, 1)[0]
    return (r.astype(np.float64) + 0.5) * 2.0**-32


def latent_normals(seed, rows, D):
    """Box-Muller normals of the latent draw (float64 evaluation of the fp32 kernel)."""
    rows = np.asarray(rows, dtype=np.uint64)
    out = np.empty((len(rows), D))
    for d0 in range(0, D, 4):
        r = rows_counter(seed, rows, d0 // 4, 0)
        u = [np.minimum(((w.astype(np.float32) + np.float32(0.5)) * np.float32(2.0**-32)), np.float32(1.0)).astype(np.float64) for w in r]
        rad0, rad1 = np.sqrt(-2 * np.log(u[0])), np.sqrt(-2 * np.log(u[2]))
        th0, th1 = 2 * np.pi * u[1] - np.pi, 2 * np.pi * u[3] - np.pi
        v = [rad0 * np.cos(th0), rad0 * np.sin(th0), rad1 * np.cos(th1), rad1 * np.sin(th1)]
        for j in range(4):
            if d0 + j < D:
                out[:, d0 + j] = v[j]
    return out

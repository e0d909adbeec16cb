"""Reference Philox4x32-10 in numpy (TEST INFRASTRUCTURE): the published counter-based
generator of Salmon et al. (SC'11), with the stream convention of include/ctgan_sm100.h:
element e of stream `seed` is lane e&3 of philox(counter=e>>2, key=seed);
u = (bits >> 8) * 2^-24; normal i = sqrt(-2 ln u[2i]) * cos(2 pi u[2i+1])."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def blocks(seed, ctr):
    """ctr: uint64 array of counters -> uint32 array [len, 4]"""
    ctr = np.asarray(ctr, dtype=np.uint64)
    c0 = (ctr & MASK).astype(np.uint32)
    c1 = (ctr >> np.uint64(32)).astype(np.uint32)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return np.stack([c0, c1, c2, c3], axis=1)


def bits(seed, offset, n):
    e = np.arange(n, dtype=np.uint64) + np.uint64(offset)
    first, last = int(e[0] >> np.uint64(2)), int(e[-1] >> np.uint64(2))
    blk = blocks(seed, np.arange(first, last + 1, dtype=np.uint64))
    idx = (e >> np.uint64(2)).astype(np.int64) - first
    return blk[idx, (e & np.uint64(3)).astype(np.int64)]


def uniform(seed, offset, n):
    if n == 0:
        return np.zeros(0, dtype=np.float32)
    return ((bits(seed, offset, n) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def normal(seed, offset, n):
    u = uniform(seed, offset, 2 * n)
    u1 = np.maximum(u[0::2], np.float32(5.9604645e-8))
    u2 = u[1::2]
    return (np.sqrt(np.float32(-2.0) * np.log(u1)) * np.cos(np.float32(2.0 * np.pi) * u2)).astype(np.float32)

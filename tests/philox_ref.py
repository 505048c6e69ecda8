"""Host restatement (numpy, vectorised) of the in-kernel dropout stream of libmhimk (include/mhimk.h, mil_dropout_t mode 2):
Philox4x32-10 (Salmon et al., SC'11; the same round function as curand / torch), counter = (row, column/8, offset_lo, offset_hi),
key = (seed_lo, seed_hi); each call yields eight 16-bit uniforms (low half-word first); keep iff u < round((1-p)*65536).
Test infrastructure only."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) for v in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & mask
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & mask
        c0, c1, c2, c3 = n0, p1 & mask, n2, p0 & mask
        k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return c0, c1, c2, c3


def keep_mask(rows, ncols, p, seed, offset):
    """bool [rows, ncols]: True = kept."""
    thresh = int(round((1.0 - p) * 65536.0))
    r = np.repeat(np.arange(rows, dtype=np.uint64), ncols // 8)
    g = np.tile(np.arange(ncols // 8, dtype=np.uint64), rows)
    o0 = np.full_like(r, offset & 0xFFFFFFFF)
    o1 = np.full_like(r, (offset >> 32) & 0xFFFFFFFF)
    out = philox4x32_10(r, g, o0, o1, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = np.empty((rows * (ncols // 8), 8), dtype=np.uint64)
    for j, w in enumerate(out):
        u[:, 2 * j] = w & np.uint64(0xFFFF)
        u[:, 2 * j + 1] = w >> np.uint64(16)
    return (u < thresh).reshape(rows, ncols)

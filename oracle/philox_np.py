"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based noise layout of iwvi_normal_fill
(include/iwvi_b200.h): Philox4x32-10 (Salmon et al., SC'11; the generator behind tf.random_normal at reference
layers.py:86 and temp_workaround.py:89, whose stream TF does not let us reproduce) keyed by (seed_lo, seed_hi), counter =
flat element index >> 1, Box-Muller pair lane = index & 1.  Flat index = (first_point + p) * C + c, i.e. the
reference's [n, k, c] row-major order."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)]
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniforms(flat_pairs, seed):
    """Returns (u1, u2) in (0,1) for each pair counter; exact 53-bit values (bit-exact with the CUDA kernel)."""
    q = np.asarray(flat_pairs, dtype=np.uint64)
    r0, r1, r2, r3 = philox4x32_10(q & MASK, q >> np.uint64(32), np.zeros_like(q), np.zeros_like(q),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    a = (r1 << np.uint64(32)) | r0
    b = (r3 << np.uint64(32)) | r2
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0
    u2 = ((b >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0
    return u1, u2


def normal(n_points, C, first_point, seed):
    f0 = first_point * C
    idx = f0 + np.arange(n_points * C, dtype=np.int64)
    u1, u2 = uniforms(idx >> 1, seed)
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.where(idx & 1, rad * np.sin(2 * np.pi * u2), rad * np.cos(2 * np.pi * u2))
    return z.reshape(n_points, C)

"""NumPy restatement of the Philox4x32-10 stream of qg_init_state (csrc/qg_api.cu:
philox_uniform): counter = (node lo, node hi, field, 0), key = seed, 53-bit uniform in [0, 1).
Test infrastructure: pins the device-side initial condition bit for bit."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox_words(node, field, seed):
    """The four 32-bit output words (as uint64 arrays) for counter (node lo, node hi, field, 0)."""
    node = np.asarray(node, dtype=np.uint64)
    c0, c1 = node & MASK, node >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(field))
    c3 = np.zeros_like(c0)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2                      # 32x32 -> 64 bit products
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_uniform(node, field, seed):
    c0, c1, _, _ = philox_words(node, field, seed)
    bits = ((c0 << np.uint64(32)) | c1) >> np.uint64(11)
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)


def rand_fields(M, P, seed, member=0):
    """The two (M+2, P+2) draws initialise_model consumes (src/model.jl:41-42), as the device
    generates them: interior node (i, j) -> counter j*M + i; ghosts are overwritten afterwards."""
    out = []
    for layer in range(2):
        j, i = np.meshgrid(np.arange(P, dtype=np.uint64), np.arange(M, dtype=np.uint64), indexing="ij")
        u = philox_uniform(j * np.uint64(M) + i, member * 2 + layer, seed)     # [j, i]
        full = np.zeros((M + 2, P + 2), order="F")
        full[1:-1, 1:-1] = u.T
        out.append(full)
    return out

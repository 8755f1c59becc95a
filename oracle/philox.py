"""ORACLE -- test infrastructure only (imported by tests/; never by the product path).

numpy restatement of the counter-based generator libb2v.so uses for the DDPM ancestral noise when the caller passes
none (csrc/ew_kernels.cu: philox4x32_10 / philox_normal4).  The reference itself draws this noise with
torch.randn_like (models/diffusion.py:333); the Python mirror reproduces that stream by passing torch's draws to the C
ABI, so this generator is an extension for C callers without torch, not a restatement of reference arithmetic.
Algorithm: Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
Random123 library's philox4x32_R(10)), pinned against Random123's published known-answer vectors in
tests/test_oracle_golden.py.  Normal draws: u = (word + 0.5) / 2^32 (fp32), Box-Muller on word pairs.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32, key: (2,) uint32 -> (..., 4) uint32"""
    c = np.array(counter, dtype=np.uint32, copy=True)
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            c1 = p1.astype(np.uint32)
            c3 = p0.astype(np.uint32)
            c0, c2 = n0, n2
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return np.stack([c0, c1, c2, c3], axis=-1)


def normal(n, seed, step):
    """the n N(0,1) draws of loop step `step`: element i comes from counter (i // 4, 0, step, 0), word pair (i % 4) // 2"""
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    ctr = np.stack([(q & np.uint64(0xFFFFFFFF)).astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32),
                    np.full(nq, step, dtype=np.uint32), np.zeros(nq, dtype=np.uint32)], axis=-1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = (r.astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -32)  # fp32 like the kernel
    u1 = np.minimum(u[:, 0::2], np.float32(0.99999994)).astype(np.float64)
    u2 = u[:, 1::2].astype(np.float64)
    rad = np.sqrt(-2.0 * np.log(u1))
    out = np.empty((nq, 4))
    out[:, 0::2] = rad * np.cos(2 * np.pi * u2)
    out[:, 1::2] = rad * np.sin(2 * np.pi * u2)
    return out.reshape(-1)[:n]

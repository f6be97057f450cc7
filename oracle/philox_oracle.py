"""numpy statement of the engine's in-kernel noise stream (TEST INFRASTRUCTURE ONLY).

The reference draws `torch.randn(K,T,nu)` (mppi.py:203): mt19937 on CPU, Philox on CUDA; neither
stream can be consumed in-kernel sample-by-sample, so the engine defines its own counter-based
stream and this file is its specification (csrc/mppi_math.cuh `philox4x32_10`, `Normals<>`):

  Philox4x32-10 (Salmon et al., SC'11; the same generator curand/ATen use), with
      key     = (seed_lo32, seed_hi32)
      counter = (c_lo32, c_hi32, k_lo32, k_hi32),   c = offset + chunk, k = GLOBAL sample index
  float32: each call yields 4 normals: (x,y)->(n0,n1), (z,w)->(n2,n3) with
      u = v*2^-32 + 2^-33 (computed in float32), r = sqrt(-2 ln u1), n0 = r sin(2 pi u2), n1 = r cos(2 pi u2)
  float64: each call yields 2 normals from two 53-bit uniforms u = (v>>11)*2^-53 + 2^-54,
      a = y<<32|x, b = w<<32|z.
  Sample k's flat noise vector (length R=T*nu, KMPPI S*nu) is the concatenation of its chunks.

Known-answer check: the Random123 reference vector philox4x32_10(ctr=0, key=0) =
6627e8d5 e169c58d bc57ac4c 9b00dbd8 (tests/test_philox_oracle.py).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (...,4) uint32 array-like, key: (2,) ints -> (...,4) uint32"""
    c = np.asarray(ctr, dtype=np.uint64).copy()
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c[..., 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[..., 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=-1)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def normals(seed, offset, k_offset, K, R, dtype):
    """(K,R) standard normals for samples k_offset..k_offset+K-1 of one command."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    key = (seed & 0xFFFFFFFF, seed >> 32)
    per = 4 if dtype == np.float32 else 2
    chunks = (R + per - 1) // per
    kk = (np.arange(K, dtype=np.uint64) + np.uint64(k_offset))[:, None].repeat(chunks, axis=1)
    cc = (np.uint64(offset) + np.arange(chunks, dtype=np.uint64))[None, :].repeat(K, axis=0)
    ctr = np.stack([cc & MASK, cc >> np.uint64(32), kk & MASK, kk >> np.uint64(32)], axis=-1)
    v = philox4x32_10(ctr, key)                       # (K,chunks,4)
    if dtype == np.float32:
        u = v.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)
        u = u.astype(np.float64)
        out = np.empty((K, chunks, 4), dtype=np.float64)
        for a, b, o in ((0, 1, 0), (2, 3, 2)):
            r = np.sqrt(-2.0 * np.log(u[..., a]))
            ang = 2.0 * np.pi * u[..., b]
            out[..., o] = r * np.sin(ang)
            out[..., o + 1] = r * np.cos(ang)
        return out.reshape(K, chunks * 4)[:, :R].astype(np.float32)
    a = (v[..., 1].astype(np.uint64) << np.uint64(32)) | v[..., 0].astype(np.uint64)
    b = (v[..., 3].astype(np.uint64) << np.uint64(32)) | v[..., 2].astype(np.uint64)
    u1 = (a >> np.uint64(11)).astype(np.float64) * 2.0 ** -53 + 2.0 ** -54
    u2 = (b >> np.uint64(11)).astype(np.float64) * 2.0 ** -53 + 2.0 ** -54
    r = np.sqrt(-2.0 * np.log(u1))
    out = np.stack([r * np.sin(2.0 * np.pi * u2), r * np.cos(2.0 * np.pi * u2)], axis=-1)
    return out.reshape(K, chunks * 2)[:, :R]

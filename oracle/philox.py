"""TEST INFRASTRUCTURE ONLY -- Philox4x32-10 noise generator, CPU (numpy) restatement.

The reference samples the reparameterisation noise with ``tf.random.normal`` (models.py:108,
nb-radial cell 5 ``tf.shape`` variant) and never seeds it, so there is no reference noise
stream to reproduce.  The noise contract is therefore OURS: the CUDA kernels and this oracle
implement the same counter-based generator, so "identical seeds" is meaningful and the result
does not depend on how the batch is sharded over GPUs.

Contract (must match csrc/dib_common.cuh :: dib_philox_normal4):
    counter = (global_sample_lo, global_sample_hi ^ (feature << 8), dim_quad, step)
    key     = (seed_lo, seed_hi)
    -> 4 uint32 r0..r3 after 10 Philox rounds
    u_k = ((r_k >> 8) + 0.5) * 2^-24      (in (0,1); exactly representable in fp32, so the
                                           device and this oracle see the SAME uniforms)
    Box-Muller: n0 = sqrt(-2 ln u0) cos(2 pi u1), n1 = sqrt(-2 ln u0) sin(2 pi u1),
                n2 = sqrt(-2 ln u2) cos(2 pi u3), n3 = sqrt(-2 ln u2) sin(2 pi u3)
    eps[sample, feature, 4*dim_quad + k] = n_k

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def normal_noise(seed, step, sample_ids, num_features, embed_dim, dtype=np.float32):
    """eps[len(sample_ids), num_features, embed_dim] for the given GLOBAL sample ids.

    embed_dim is rounded up to a multiple of 4 internally and truncated.
    """
    sample_ids = np.asarray(sample_ids, dtype=np.uint64)
    n = sample_ids.shape[0]
    q = (embed_dim + 3) // 4
    s_lo = (sample_ids & _MASK)[:, None, None]
    s_hi = (sample_ids >> np.uint64(32))[:, None, None]
    feat = np.arange(num_features, dtype=np.uint64)[None, :, None]
    quad = np.arange(q, dtype=np.uint64)[None, None, :]
    c1 = (s_hi ^ (feat << np.uint64(8))) & _MASK
    r0, r1, r2, r3 = philox4x32_10(s_lo, c1, quad, np.uint64(int(step) & 0xFFFFFFFF),
                                   int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    scale = 2.0 ** -24
    u0 = ((r0 >> np.uint32(8)).astype(np.float64) + 0.5) * scale
    u1 = ((r1 >> np.uint32(8)).astype(np.float64) + 0.5) * scale
    u2 = ((r2 >> np.uint32(8)).astype(np.float64) + 0.5) * scale
    u3 = ((r3 >> np.uint32(8)).astype(np.float64) + 0.5) * scale
    ra = np.sqrt(-2.0 * np.log(u0))
    rb = np.sqrt(-2.0 * np.log(u2))
    out = np.empty((n, num_features, q, 4), dtype=np.float64)
    out[..., 0] = ra * np.cos(2.0 * np.pi * u1)
    out[..., 1] = ra * np.sin(2.0 * np.pi * u1)
    out[..., 2] = rb * np.cos(2.0 * np.pi * u3)
    out[..., 3] = rb * np.sin(2.0 * np.pi * u3)
    return out.reshape(n, num_features, 4 * q)[:, :, :embed_dim].astype(dtype)


def dropout_keep(seed, step, sample_ids, feature, layer, width, rate):
    """Keras Dropout keep-mask of one encoder activation [len(sample_ids), width] (nb-radial cell 5: Dropout(rate) after
    every hidden Dense of the feature encoders), from the same counter-based generator as the noise so that the CUDA
    kernels and this oracle agree and the result does not depend on the sharding:
        counter = (sample_lo, sample_hi ^ (feature << 8), 0x80000000 | (layer << 24) | (column // 4), step)
        u_k as above (k = column % 4);  keep = u_k >= rate   (probability 1 - rate)
    Must match csrc/dib_elementwise.cu :: dib_dropout_kernel."""
    sample_ids = np.asarray(sample_ids, dtype=np.uint64)
    q = (width + 3) // 4
    s_lo = (sample_ids & _MASK)[:, None]
    s_hi = (sample_ids >> np.uint64(32))[:, None]
    quad = np.arange(q, dtype=np.uint64)[None, :]
    c1 = (s_hi ^ np.uint64(int(feature) << 8)) & _MASK
    c2 = (np.uint64(0x80000000) | np.uint64(int(layer) << 24) | quad) & _MASK
    r = philox4x32_10(s_lo, c1, c2, np.uint64(int(step) & 0xFFFFFFFF), int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    u = np.stack([((rk >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24) for rk in r], -1)
    return (u.reshape(len(sample_ids), 4 * q)[:, :width] >= np.float32(rate))

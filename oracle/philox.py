"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11, Random123), numpy.  TEST INFRASTRUCTURE.

The reference draws its uniforms inside vegasflow with TensorFlow's stateful generator
(`tf.random.uniform`, seed 4 in utilities.py:89); that stream is not reproducible outside TF, so
the B200 path defines its own stream and this module restates it for the parity tests:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (event & 0xffffffff, event >> 32, iteration, j)      j = 0 .. ceil(ndim/2)-1
    block j yields dimensions 2j and 2j+1:   u = ((w_a << 32 | w_b) >> 11) * 2^-53  in [0,1)

`event` is the GLOBAL event index inside the iteration, so the sample set does not depend on how
events are split over GPUs or chunks.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 uint32 arrays, key: 2 uint32 scalars/arrays -> 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in ctr]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _to_unit(a, b):
    v = (a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)
    return (v >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniforms(seed, iteration, first_event, nevents, ndim):
    """(nevents, ndim) float64 in [0,1) for global events first_event .. first_event+nevents-1."""
    ev = np.arange(first_event, first_event + nevents, dtype=np.uint64)
    lo, hi = (ev & _MASK).astype(np.uint32), (ev >> np.uint64(32)).astype(np.uint32)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((nevents, ndim))
    it = np.full(nevents, iteration, dtype=np.uint32)
    for j in range((ndim + 1) // 2):
        r = philox4x32_10((lo, hi, it, np.full(nevents, j, dtype=np.uint32)), key)
        out[:, 2 * j] = _to_unit(r[0], r[1])
        if 2 * j + 1 < ndim:
            out[:, 2 * j + 1] = _to_unit(r[2], r[3])
    return out

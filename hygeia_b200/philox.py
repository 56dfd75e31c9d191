"""Counter-based Philox4x32-10 uniforms, indexed by (seed, chain, site).

The reference draws its resampling uniforms from a sequential stream
(arma::randu() -> R's RNG, misc/resample.h:126); because resampling decisions
are discontinuous in the weights, a sequential stream cannot be replayed
across implementations.  This framework therefore indexes
the draw BY SITE: u[t] = philox(key = seed, counter = (t, chain)).  The device
kernel (csrc/hyg_philox.h) evaluates the same function; this numpy copy lets the
host and the CPU oracle see the identical numbers.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
TAG = 0x48594745  # "HYGE"


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; inputs broadcastable uint32 arrays, returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def uniforms_by_site(seed: int, chain: int, T: int, t0: int = 0) -> np.ndarray:
    """u[t] in [0,1) with 53 random bits for sites t0..t0+T-1 of chain `chain` under `seed`."""
    t = np.arange(t0, t0 + T, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10(t & _MASK, t >> np.uint64(32), np.uint64(chain & 0xFFFFFFFF), np.uint64(TAG),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    hi = (x0.astype(np.uint64) >> np.uint64(5))   # 27 bits
    lo = (x1.astype(np.uint64) >> np.uint64(6))   # 26 bits
    return (hi * np.float64(67108864.0) + lo) * np.float64(1.0 / 9007199254740992.0)

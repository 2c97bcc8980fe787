"""Counter-based complex-normal generator shared by the oracle, the tests and the benchmarks.

The reference draws `randn` through `ITensors.random_itensor` under Julia's Xoshiro
(`/root/reference/src/TensorNetworks/tensornetworkstate.jl:82-92`), which cannot be regenerated
outside Julia.  This generator is keyed by (seed, stream, element) so any
language can reproduce the same ComplexF64 inputs: SplitMix64 hash -> two uniforms -> Box-Muller.
The CUDA library has the same formula in `tnqs_fill_random` (device libm, last-ulp differences).
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_G = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)


def _mix(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * _C1
    z = (z ^ (z >> np.uint64(27))) * _C2
    return z ^ (z >> np.uint64(31))


def complex_normal(seed: int, stream: int, n: int, offset: int = 0) -> np.ndarray:
    """n samples of N(0,1) + i N(0,1), elements offset..offset+n-1 of stream `stream`."""
    with np.errstate(over="ignore"):
        base = _mix(np.uint64(seed) * _G + np.uint64(stream) + np.uint64(1))
        idx = np.arange(offset, offset + n, dtype=np.uint64)
        z1 = _mix(base + (np.uint64(2) * idx + np.uint64(1)) * _G)
        z2 = _mix(base + (np.uint64(2) * idx + np.uint64(2)) * _G)
    u1 = ((z1 >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)   # (0, 1]
    u2 = (z2 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)           # [0, 1)
    r = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    return r * np.cos(ang) + 1j * (r * np.sin(ang))

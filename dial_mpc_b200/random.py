"""PRNG keys for the planner: Threefry-2x32 keys like ``jax.random`` (uint32[2]).

``PRNGKey(seed) = (0, seed)`` and ``split`` follow JAX's legacy (non-partitionable)
counter layout; the Threefry arithmetic runs in the C library
(host function ``dial_key_split``) so that host and device share one implementation."""
from __future__ import annotations

import ctypes as C

import numpy as np

from dial_mpc_b200 import _capi


def PRNGKey(seed: int) -> np.ndarray:
    return np.array([0, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key):
    """-> (new_rng, subkey)  ==  tuple(jax.random.split(key))."""
    key = np.ascontiguousarray(key, dtype=np.uint32)
    a = (C.c_uint32 * 2)()
    b = (C.c_uint32 * 2)()
    _capi.lib().dial_key_split(key.ctypes.data_as(C.POINTER(C.c_uint32)), a, b)
    return np.array(a[:], dtype=np.uint32), np.array(b[:], dtype=np.uint32)


# ---- the pieces of jax.random that `sample_command` needs (host side, a few calls per 500 steps) --
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds (Random123 / jax._src.prng.threefry2x32), vectorised over counters."""
    k0, k1 = (np.uint32(key[0]), np.uint32(key[1]))
    ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
    x0 = np.asarray(x0, dtype=np.uint32).copy()
    x1 = np.asarray(x1, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for r in range(5):
            for rot in _ROT[r % 2]:
                x0 += x1
                x1 = (x1 << np.uint32(rot)) | (x1 >> np.uint32(32 - rot))
                x1 ^= x0
            x0 += ks[(r + 1) % 3]
            x1 += ks[(r + 2) % 3] + np.uint32(r + 1)
    return x0, x1


def split_n(key, num: int) -> np.ndarray:
    """jax.random.split(key, num) in the legacy counter layout: counters 0..2num-1, first half
    paired with second half, outputs concatenated and read back as (num, 2)."""
    c = np.arange(2 * num, dtype=np.uint32)
    y0, y1 = threefry2x32(key, c[:num], c[num:])
    return np.concatenate([y0, y1]).reshape(num, 2)


def uniform1(key, minval: float, maxval: float) -> np.float32:
    """jax.random.uniform(key, (1,), minval=, maxval=)[0]: 32 random bits (counter 0, odd length
    padded with 0) -> mantissa of a float in [1, 2) -> scaled, fp32 like JAX."""
    bits = threefry2x32(key, np.zeros(1, np.uint32), np.zeros(1, np.uint32))[0]
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32)[0] - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo)


def uniform(key, n: int, minval: float, maxval: float) -> np.ndarray:
    """jax.random.uniform(key, (n,), minval=, maxval=) in the legacy counter layout: counters
    0..n-1 (an odd count padded with one zero), first half paired with second half."""
    m = n + (n & 1)
    c = np.arange(m, dtype=np.uint32)
    if n & 1:
        c[-1] = 0
    y0, y1 = threefry2x32(key, c[:m // 2], c[m // 2:])
    bits = np.concatenate([y0, y1])[:n]
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo)

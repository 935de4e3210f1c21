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

"""Custom-reward builds of the sampling core.

The reference lets users write their own env in JAX (README.md:223-312, ``--custom-env`` at
dial_mpc/core/dial_core.py:202-204).  Here ``step``'s physics is the fused rollout kernel, so a
custom env supplies (1) an MJCF model, compiled by ``dial_mpc_b200.modelc``, and (2) its reward
as ONE CUDA device function (contract: ``include/dial_custom_reward.h``).  ``build_library``
compiles a dedicated build of ``libdial_b200`` with that function fused into the rollout
kernel — sm_100a, same flags as the stock library, only the solver instantiation the model
needs — and caches it in-tree under ``dial_mpc_b200/_custom/`` keyed by the content hash of
every source that goes into it.  There is no interpreter / CPU fallback for the reward.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from typing import Optional

from dial_mpc_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
CACHE_DIR = os.path.join(_HERE, "_custom")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-use_fast_math", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def _sources(reward_path: str):
    return [os.path.join(CSRC, f) for f in ("dial_kernels.cu", "dial_rollout_variant.cu", "dial_device.cuh", "dial_host.h")] + \
        [os.path.join(_ROOT, "include", "dial_b200.h"), os.path.join(_ROOT, "include", "dial_custom_reward.h"),
         reward_path]


def _tag(variant: int, dense_nv: Optional[int], defines=()) -> str:
    """Build key of a solver instantiation: the variant, plus the dof count for the dense solver
    (variant 3 is compiled per nv: -DDIAL_DENSE_NV; the stock library carries nv = 22), plus a mark
    for extra compile-time options (``defines``, e.g. DIAL_ROBUST_LS)."""
    return (f"v{variant}" + (f"n{dense_nv}" if variant == 3 and dense_nv not in (None, 22) else "")
            + ("x" if defines else ""))


def reward_id(reward_path: str, variant: int, dense_nv: Optional[int] = None, defines=()) -> str:
    h = hashlib.sha256()
    for s in _sources(reward_path):
        h.update(open(s, "rb").read())
    h.update((" ".join(NVCC_FLAGS) + f" variant={_tag(variant, dense_nv)[1:]}"
              + "".join(f" -D{d}" for d in sorted(defines))).encode())
    return h.hexdigest()[:16]


def library_path(reward_path: str, variant: int, dense_nv: Optional[int] = None, defines=()) -> str:
    stem = os.path.splitext(os.path.basename(reward_path))[0]
    return os.path.join(CACHE_DIR, f"libdial_b200_{stem}_{_tag(variant, dense_nv, defines)}_"
                                   f"{reward_id(reward_path, variant, dense_nv, defines)}.so")


def solver_variant(model) -> int:
    """Solver instantiation a compiled model maps to (host logic of the stock library).  Models on the
    dense (elliptic-cone) path map to variant 3 whatever their dof count: a custom build instantiates
    the dense solver for the model's own nv (``dense_nv``)."""
    md = _capi.fill_model_desc(model)
    if dense_nv(model) is not None:
        return 3
    v = _capi.lib().dial_solver_variant(md)
    if v < 0:
        raise RuntimeError(f"model not supported: {_capi.lib().dial_last_error().decode()}")
    return v


def dense_nv(model) -> Optional[int]:
    """nv of a model that needs the dense solver path (elliptic friction cones), else None."""
    md = _capi.fill_model_desc(model)
    if int(md.cone) != 1:
        return None
    if not 1 <= int(md.nv) <= 32:
        raise RuntimeError("the dense solver keeps one matrix row per lane: nv <= 32")
    return int(md.nv)


def build_library(reward_path: str, model=None, variant: Optional[int] = None, force: bool = False,
                  verbose: bool = False, defines=()) -> str:
    """Compile (or reuse) the library with ``reward_path`` fused in; returns the ``.so`` path.
    ``defines``: extra compile-time options of the kernels, e.g. ``("DIAL_ROBUST_LS",)`` — the line search
    with a bracket that only narrows (a documented deviation from the reference's rule)."""
    defines = tuple(defines)
    reward_path = os.path.abspath(reward_path)
    if not os.path.exists(reward_path):
        raise FileNotFoundError(reward_path)
    if variant is None:
        if model is None:
            raise ValueError("pass the compiled model (or the solver variant) the library is for")
        variant = solver_variant(model)
    nvd = dense_nv(model) if (model is not None and variant == 3) else None
    out = library_path(reward_path, variant, nvd, defines)
    if os.path.exists(out) and not force:
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: custom rewards are compiled CUDA, there is no fallback")
    os.makedirs(CACHE_DIR, exist_ok=True)
    tmp = out + f".tmp{os.getpid()}"
    cmd = [nvcc] + NVCC_FLAGS + [f'-DDIAL_CUSTOM_REWARD_FILE="{reward_path}"',
                                 f"-DDIAL_CUSTOM_REWARD_ID={reward_id(reward_path, variant, nvd, defines)}",
                                 f"-DDIAL_ONLY_VARIANT={variant}"] + ([f"-DDIAL_DENSE_NV={nvd}"] if nvd else []) + [f"-D{d}" for d in defines] + [
                                 "-o", tmp, os.path.join(CSRC, "dial_kernels.cu")]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError(f"nvcc failed for {reward_path}:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, out)  # atomic: concurrent ranks may build the same library
    # superseded builds of the same reward (older kernel or reward sources) are dropped
    import glob
    stem = os.path.splitext(os.path.basename(reward_path))[0]
    for old in glob.glob(os.path.join(CACHE_DIR, f"libdial_b200_{stem}_{_tag(variant, nvd, defines)}_*.so")):
        if old != out:
            try:
                os.remove(old)
            except OSError:
                pass
    return out

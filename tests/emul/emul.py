"""TEST-ONLY python driver of the CPU warp emulator (tests/emul/libdial_emul.so)."""
import ctypes as C
import os
import subprocess

import numpy as np

from dial_mpc_b200 import _capi

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libdial_emul.so")


def build(force=False):
    srcs = [os.path.join(_DIR, "emul_main.cpp"), os.path.join(_DIR, "warp_emul.h"),
            os.path.join(_DIR, "..", "..", "dial_mpc_b200", "csrc", "dial_device.cuh"),
            os.path.join(_DIR, "..", "..", "dial_mpc_b200", "csrc", "dial_host.h"),
            os.path.join(_DIR, "..", "..", "include", "dial_b200.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", _DIR, "-shared", "-fPIC", "-o", _SO,
                               os.path.join(_DIR, "emul_main.cpp")])
    return C.CDLL(_SO)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def rollout(env, plan_desc, qpos, qvel, warm, step0=0, stage0=0, us=None, eps=None, Ybar=None,
            noise=None, key=(0, 0), mode=0, nrows=None, H=None, want_traj=True):
    lib = build()
    md = _capi.fill_model_desc(env.sys.model)
    nq, nv, nu, nb = md.nq, md.nv, md.nu, md.nbody
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    qpos, qvel, warm, us, eps, Ybar, noise = map(f32, (qpos, qvel, warm, us, eps, Ybar, noise))
    if mode == 0:
        nrows, H = us.shape[0], us.shape[1]
    out = dict(rewss=np.zeros((nrows, H), np.float32), rews=np.zeros(nrows, np.float32),
               q=np.zeros((nrows, H, nq), np.float32), qd=np.zeros((nrows, H, nv), np.float32),
               xpos=np.zeros((nrows, H, nb - 1, 3), np.float32), qpos_out=np.zeros(nq, np.float32),
               qvel_out=np.zeros(nv, np.float32), warm_out=np.zeros(nv, np.float32),
               ctrl_out=np.zeros(nu, np.float32), slab=np.zeros(8192, np.float32))
    rc = lib.emul_rollout(C.byref(md), C.byref(plan_desc), mode, nrows, H, step0, stage0, _p(qpos), _p(qvel),
                          _p(warm), _p(us), _p(eps), _p(Ybar), _p(noise), C.c_uint32(key[0]), C.c_uint32(key[1]),
                          _p(out["rewss"]), _p(out["rews"]), _p(out["q"]), _p(out["qd"]), _p(out["xpos"]),
                          _p(out["qpos_out"]), _p(out["qvel_out"]), _p(out["warm_out"]), _p(out["ctrl_out"]),
                          _p(out["slab"]))
    assert rc == 0
    return out

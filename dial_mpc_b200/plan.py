"""Thin Python wrapper over the C-ABI plan handle (include/dial_b200.h).

PyTorch is used only for device memory and streams; every compute call goes to the
hand-written kernels in ``csrc/libdial_b200.so``."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from dial_mpc_b200 import _capi


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "need contiguous fp32 CUDA tensor"
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _key(key):
    if key is None:
        return None
    k = np.ascontiguousarray(key, dtype=np.uint32)
    return (C.c_uint32 * 2)(int(k[0]), int(k[1]))


class Plan:
    def __init__(self, env, desc: "_capi.dial_plan_desc", device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("dial_mpc_b200 needs a CUDA device (B200); there is no CPU fallback")
        # custom-reward envs carry their own build of the library (dial_mpc_b200.custom)
        self.lib = _capi.lib(getattr(env, "library_path", None))
        self.env = env
        self.desc = desc
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.mdesc = _capi.fill_model_desc(env.sys.model)
        with torch.cuda.device(self.device):
            self.handle = self.lib.dial_plan_create(C.byref(self.mdesc), C.byref(desc))
        if not self.handle:
            raise RuntimeError(f"dial_plan_create failed: {self.lib.dial_last_error().decode()}")
        m = self.mdesc
        self.nq, self.nv, self.nu, self.nbody = m.nq, m.nv, m.nu, m.nbody
        self.N, self.Ntotal = desc.Nsample, desc.Ntotal
        self.Hs, self.Hn = desc.Hsample, desc.Hnode
        self.exchange_on = False
        self._cmd = None          # last command override uploaded (randomize_tasks)
        self._stages = None       # identity of the last stage tables uploaded (seq-jump randomize_tasks)

    def set_command(self, override) -> None:
        """``override`` = (step, vel[3], ang[3]) or None: the one-step random command of
        ``randomize_tasks`` (envs' ``command_override``); uploaded only when it changes."""
        key = None if override is None else (int(override[0]), tuple(np.float32(override[1]).tolist()),
                                             tuple(np.float32(override[2]).tolist()))
        if key == self._cmd:
            return
        if key is None:
            self._check(self.lib.dial_plan_set_command(self.handle, -1, None, None, _stream()))
        else:
            v, a = (C.c_float * 3)(*key[1]), (C.c_float * 3)(*key[2])
            self._check(self.lib.dial_plan_set_command(self.handle, key[0], v, a, _stream()))
        self._cmd = key

    def set_stages(self, tables) -> None:
        """``tables`` = (pose [n,3], yaw [n], contact_targets [n,4,3], contact_radius [n,4]): the jump
        sequence of the state a launch starts from (seq-jump ``randomize_tasks``: drawn at reset);
        uploaded only when it differs from what the plan holds."""
        pose, yaw, tgt, rad = (np.ascontiguousarray(t, dtype=np.float32) for t in tables)
        key = (pose.tobytes(), yaw.tobytes(), tgt.tobytes(), rad.tobytes())
        if key == self._stages:
            return
        n = int(pose.shape[0])
        assert yaw.shape == (n,) and tgt.shape == (n, 4, 3) and rad.shape == (n, 4)
        F = C.POINTER(C.c_float)
        self._check(self.lib.dial_plan_set_stages(self.handle, n, pose.ctypes.data_as(F), yaw.ctypes.data_as(F),
                                                  tgt.ctypes.data_as(F), rad.ctypes.data_as(F), _stream()))
        self._stages = key

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(f"dial_b200: {self.lib.dial_last_error().decode()} (rc={rc})")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.dial_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------------------
    def f32(self, x, shape=None) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.device, dtype=torch.float32).contiguous()
        else:
            t = torch.as_tensor(np.asarray(x, dtype=np.float32), device=self.device).contiguous()
        if shape is not None:
            assert tuple(t.shape) == tuple(shape), f"expected shape {shape}, got {tuple(t.shape)}"
        return t

    def empty(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def _state(self, state, horizon: int = 1) -> Tuple["_capi.dial_state", tuple]:
        ps = state.pipeline_state
        qpos, qvel, warm = self.f32(ps.qpos, (self.nq,)), self.f32(ps.qvel, (self.nv,)), self.f32(ps.qacc_warmstart, (self.nv,))
        s = _capi.dial_state()
        s.qpos, s.qvel, s.qacc_warmstart = qpos.data_ptr(), qvel.data_ptr(), warm.data_ptr()
        s.step = int(state.info.get("step", 0))
        s.stage = int(state.info.get("contact_stage", 0))
        if state.info.get("randomize_target", False):
            # every launch that starts from `state` sees the random command its horizon may reach
            self.set_command(self.env.command_override(state.info, horizon))
            if hasattr(self.env, "stage_tables"):
                self.set_stages(self.env.stage_tables(state.info))
        return s, (qpos, qvel, warm)

    @property
    def launches(self) -> int:
        return int(self.lib.dial_launch_count(self.handle))

    # -- API ------------------------------------------------------------------------------------
    def pipeline_init(self, qpos):
        from dial_mpc_b200.envs.base_env import PipelineState
        q = self.f32(qpos, (self.nq,))
        qv = torch.zeros(self.nv, dtype=torch.float32, device=self.device)
        qo, wo = self.empty(self.nq), self.empty(self.nv)
        self._check(self.lib.dial_pipeline_init(self.handle, _ptr(q), _ptr(qv), _ptr(qo), _ptr(wo), _stream()))
        return PipelineState(qo, qv, wo, torch.zeros(self.nu, dtype=torch.float32, device=self.device))

    def env_step(self, state, action):
        from dial_mpc_b200.envs.base_env import PipelineState
        s, keep = self._state(state, 1)
        a = self.f32(action, (self.nu,))
        qo, vo, wo, r, c = self.empty(self.nq), self.empty(self.nv), self.empty(self.nv), self.empty(1), self.empty(self.nu)
        kin = self.empty(13)
        self._check(self.lib.dial_env_step_kin(self.handle, C.byref(s), _ptr(a), _ptr(qo), _ptr(vo), _ptr(wo), _ptr(r),
                                               _ptr(c), _ptr(kin), _stream()))
        ps = PipelineState(qo, vo, wo, c)
        ps.kin = kin      # torso x.pos, x.rot, body-frame velocities (what _get_obs reads of x / xd)
        return ps, r[0]

    def rollout(self, state, us, want_traj=True):
        s, keep = self._state(state, int(np.shape(us)[1]))
        us = self.f32(us)
        B, H, nu = us.shape
        assert nu == self.nu
        rewss = self.empty(B, H)
        q = self.empty(B, H, self.nq) if want_traj else None
        qd = self.empty(B, H, self.nv) if want_traj else None
        x = self.empty(B, H, self.nbody - 1, 3) if want_traj else None
        self._check(self.lib.dial_rollout(self.handle, C.byref(s), _ptr(us), B, H, _ptr(rewss), _ptr(q), _ptr(qd),
                                          _ptr(x), _stream()))
        return rewss, q, qd, x

    def reverse_rollout(self, state, eps, key, Ybar, noise_scale, rews_local):
        s, keep = self._state(state, self.Hs + 1)
        self._check(self.lib.dial_reverse_rollout(self.handle, C.byref(s), _ptr(eps), _key(key), _ptr(Ybar),
                                                  _ptr(noise_scale), _ptr(rews_local), _stream()))

    def reverse_update(self, eps, key, Ybar, noise_scale, rews_all, Ybar_out, weights=None, rews_gathered=None):
        """``rews_all=None``: the rewards come from this rank's exchange mailbox (sharded plans with a
        connected exchange); ``rews_gathered`` then receives a compact copy of all Ntotal+1 rewards."""
        self._check(self.lib.dial_reverse_update_x(self.handle, _ptr(eps), _key(key), _ptr(Ybar), _ptr(noise_scale),
                                                   _ptr(rews_all), _ptr(Ybar_out), _ptr(weights), _ptr(rews_gathered),
                                                   _stream()))

    # -- multi-GPU exchange over NVLink peer memory (include/dial_b200.h: dial_exchange_*) -------------
    def exchange_setup(self, rank: int, world: int, group=None) -> None:
        """Create this rank's mailbox, all-gather the CUDA IPC handles over ``torch.distributed`` and
        map the peers' mailboxes.  Raises if peer memory cannot be mapped (the caller then keeps NCCL)."""
        import torch.distributed as dist
        nb = _capi.DEFINES["DIAL_IPC_HANDLE_BYTES"]
        h = (C.c_ubyte * nb)()
        with torch.cuda.device(self.device):
            self._check(self.lib.dial_exchange_create(self.handle, int(rank), int(world), h))
            mine = torch.tensor(list(bytes(h)), dtype=torch.uint8, device=self.device)
            allh = torch.empty(world * nb, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            buf = (C.c_ubyte * (world * nb)).from_buffer_copy(bytes(allh.cpu().numpy().tobytes()))
            self._check(self.lib.dial_exchange_connect(self.handle, buf))
            torch.cuda.synchronize()
            dist.barrier(group=group)      # every rank has mapped every mailbox before anyone writes
        self.exchange_on = True

    def exchange_status(self) -> dict:
        out = (C.c_uint32 * 6)()
        self._check(self.lib.dial_exchange_status(self.handle, out))
        return dict(seq=int(out[0]), done=int(out[1]), error=int(out[2]), bars_seq=int(out[3]),
                    update_wait_ns=int(out[4]), bars_wait_ns=int(out[5]))

    def reverse_trajbar(self, weights, rank, qbar, qdbar, xbar):
        self._check(self.lib.dial_reverse_trajbar(self.handle, _ptr(weights), int(rank), _ptr(qbar), _ptr(qdbar),
                                                  _ptr(xbar), _stream()))

    def reverse_trajectories(self):
        """q, qd, x.pos [Nsample+1,Hs+1,*] of the last reverse_rollout (copies)."""
        H = self.Hs + 1
        q, qd, x = self.empty(self.N + 1, H, self.nq), self.empty(self.N + 1, H, self.nv), self.empty(self.N + 1, H, self.nbody - 1, 3)
        self._check(self.lib.dial_reverse_trajectories(self.handle, _ptr(q), _ptr(qd), _ptr(x), _stream()))
        return q, qd, x

    # -- device-resident MPC loop (one CUDA graph per control step) ---------------------------------
    def mpc_bind(self, bufs: dict, M_shift) -> None:
        """``bufs``: name -> CUDA tensor for every field of ``dial_mpc_buffers`` (qbar/qdbar/xbar may be
        None).  The tensors must stay alive and in place while bound (kept on ``self``)."""
        b = _capi.dial_mpc_buffers()
        want = {"counters": torch.int32, "rng": torch.int32}
        for name, _ in _capi.dial_mpc_buffers._fields_:
            t = bufs.get(name)
            if t is None:
                setattr(b, name, None)
                continue
            assert t.is_cuda and t.is_contiguous() and t.dtype == want.get(name, torch.float32), name
            setattr(b, name, t.data_ptr())
        M = np.ascontiguousarray(M_shift, dtype=np.float32)
        assert M.shape == (self.Hn + 1, self.Hn + 1)
        self._mpc_keep = (dict(bufs), b, M)
        self._check(self.lib.dial_mpc_bind(self.handle, C.byref(b), M.ctypes.data_as(C.c_void_p)))

    def mpc_step(self, n_diffuse: int, env_step=True) -> None:
        """env_step: True / 1 env step + shift, False / 0 plan only, 2 shift + plan (state untouched)."""
        self._check(self.lib.dial_mpc_step(self.handle, int(n_diffuse), int(env_step), _stream()))

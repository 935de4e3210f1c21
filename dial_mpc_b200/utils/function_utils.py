"""Host-side (numpy) versions of the reference's small helpers
(dial_mpc/utils/function_utils.py:7-43).  The CUDA rollout kernel carries its own fp32
copies (csrc/dial_device.cuh: foot_step, base_kin); these exist for API parity and tests."""
import numpy as np


def _rotate(v, q):
    s, u = q[0], np.asarray(q[1:])
    v = np.asarray(v, dtype=np.float64)
    return 2 * np.dot(u, v) * u + (s * s - np.dot(u, u)) * v + 2 * s * np.cross(u, v)


def global_to_body_velocity(v, q):
    """Rotate ``v`` by the inverse of quaternion ``q`` (wxyz)."""
    q = np.asarray(q, dtype=np.float64)
    return _rotate(v, q * np.array([1.0, -1.0, -1.0, -1.0]))


def body_to_global_velocity(v, q):
    return _rotate(v, np.asarray(q, dtype=np.float64))


def get_foot_step(duty_ratio, cadence, amplitude, phases, time):
    """Foot-height profile of the gait generator (function_utils.py:18-43)."""
    phases = np.asarray(phases, dtype=np.float64)
    t = time * 2 * np.pi * cadence + np.pi
    angle = np.mod(t + np.pi - 2 * np.pi * phases, 2 * np.pi) - np.pi
    if duty_ratio < 1:
        angle = angle * 0.5 / (1 - duty_ratio)
    clipped = np.clip(angle, -np.pi / 2, np.pi / 2)
    value = np.cos(clipped) if duty_ratio < 1 else np.zeros_like(clipped)
    final = np.where(np.abs(value) >= 1e-6, np.abs(value), 0.0)
    return amplitude * final

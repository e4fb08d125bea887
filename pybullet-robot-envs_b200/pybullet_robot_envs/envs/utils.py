"""Math helpers of the task envs (reference envs/utils.py:11-14, :78-107), batch-aware."""
import numpy as np


def goal_distance(a, b):
    """Euclidean distance along the last axis (reference utils.py:11-14)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        raise AssertionError("goal_distance(): shape of points mismatch")
    return np.linalg.norm(a - b, axis=-1)


def _check(space, data):
    data = np.asarray(data)
    # single env: identical to the reference's `assert data.shape == data_space.shape`
    # (utils.py:88); batched: every row must have the space's shape.
    assert data.shape == space.shape or (data.ndim == len(space.shape) + 1 and data.shape[1:] == space.shape)
    return data


def scale_gym_data(data_space, data):
    """[low, high] -> [-1, 1], no clipping (reference utils.py:78-91)."""
    data = _check(data_space, data)
    lo, hi = data_space.low, data_space.high
    return 2.0 * ((data - lo) / (hi - lo)) - 1.0


def unscale_gym_data(data_space, scaled_data):
    """[-1, 1] -> [low, high] (reference utils.py:94-107)."""
    scaled_data = _check(data_space, scaled_data)
    lo, hi = data_space.low, data_space.high
    return lo + (0.5 * (scaled_data + 1.0) * (hi - lo))


def euler_from_quaternion(q):
    """``p.getEulerFromQuaternion`` (btQuaternion::getEulerZYX) for one quaternion or a batch [..., 4] (x, y, z, w) ->
    [..., 3] (roll, pitch, yaw).  Same branches as the oracle's quat_to_euler (oracle/b2oracle.c)."""
    q = np.asarray(q, np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    sarg = -2.0 * (x * z - w * y)
    roll = np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z)
    pitch = np.arcsin(np.clip(sarg, -1.0, 1.0))
    yaw = np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)
    lo, hi = sarg <= -0.99999, sarg >= 0.99999
    roll = np.where(lo | hi, 0.0, roll)
    pitch = np.where(lo, -0.5 * np.pi, np.where(hi, 0.5 * np.pi, pitch))
    yaw = np.where(lo, 2 * np.arctan2(x, -y), np.where(hi, 2 * np.arctan2(-x, y), yaw))
    return np.stack([roll, pitch, yaw], axis=-1)


def quaternion_from_euler(e):
    """``p.getQuaternionFromEuler`` (btQuaternion::setEulerZYX): [..., 3] (roll, pitch, yaw) -> [..., 4] (x, y, z, w)."""
    e = np.asarray(e, np.float64)
    hr, hp, hy = 0.5 * e[..., 0], 0.5 * e[..., 1], 0.5 * e[..., 2]
    cr, sr, cp, sp, cy, sy = np.cos(hr), np.sin(hr), np.cos(hp), np.sin(hp), np.cos(hy), np.sin(hy)
    return np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy], axis=-1)

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

"""Shared helpers for the parity tests (seeded synthetic inputs for both backends)."""
import numpy as np

from pybullet_robot_envs.b2env.model import TASK_GRASP, TASK_PUSH, TASK_REACH, panda_task_setup  # noqa: F401


def sample_object_poses(B, seed=0, rnd=0.05, z=0.695):
    """Object start poses as WorldEnv._sample_pose draws them (world_env.py:145-176): uniform
    half-width `rnd` around (0.45, 0), yaw uniform in +-pi/4."""
    rng = np.random.RandomState(seed)
    pose = np.zeros((B, 7), np.float32)
    pose[:, 0] = np.clip(0.45 + rng.uniform(-rnd, rnd, B), 0.35, 0.55)
    pose[:, 1] = np.clip(rng.uniform(-rnd, rnd, B), -0.25, 0.25)
    pose[:, 2] = z
    yaw = rng.uniform(-np.pi / 4, np.pi / 4, B)
    pose[:, 5] = np.sin(yaw / 2)
    pose[:, 6] = np.cos(yaw / 2)
    return pose


def targets_for(pose, z=None):
    tg = pose[:, :3].copy()
    tg[:, 0] += 0.05
    tg[:, 1] += 0.05
    if z is not None:
        tg[:, 2] = z
    return tg.astype(np.float32)


STATE_FIELDS = ["q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "cache_key", "cache_lam", "status"]


def copy_state_to_gpu(oracle, sim, fields=("q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters",
                                            "cache_key", "cache_lam", "hand_pose")):
    for f in fields:
        sim.set(f, oracle.state[f])

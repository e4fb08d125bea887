#!/usr/bin/env python
"""Generates the golden trajectories under tests/golden/ from the fp64 build of the CPU oracle
(oracle/b2oracle.c, the restatement pinned on the SURVEY Appendix C known answers).

The reference ships no golden vectors and PyBullet cannot be imported in this image, so these
fixtures do NOT pin parity with PyBullet — they freeze the statement both implementations follow at
double precision: a regression pin for the fp32 oracle (`-m "not gpu"`) and a second, frozen target for the CUDA
path (`-m gpu`) that does not depend on the oracle being rebuilt on the GPU box.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz (small: ~20 KB each)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import b2oracle  # noqa: E402
from pybullet_robot_envs.b2env.model import TASK_PUSH, TASK_REACH, icub_task_setup, panda_task_setup  # noqa: E402

import common  # noqa: E402
import icub_cases  # noqa: E402

STATE = ("q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "hand_pose", "shaping")


def rollout(model, params, pose, target, n_settle, actions, double=True):
    B = pose.shape[0]
    orc = b2oracle.Oracle(model, params, B, double=double, nthreads=4)
    orc.reset(pose, target)
    orc.step(None, n_settle, 1, want_obs=False)
    start = {k: orc.state[k].copy() for k in STATE + ("cache_key", "cache_lam")}
    obs, rew, done = [], [], []
    for a in actions:
        o, r, d = orc.step(a, 1, 0)
        obs.append(o); rew.append(r); done.append(d)
    end = {k: orc.state[k].copy() for k in STATE}
    return start, np.stack(obs), np.stack(rew), np.stack(done), end


def save(name, start, actions, obs, rew, done, end):
    out = {"actions": actions, "obs": obs, "reward": rew, "done": done}
    out.update({"start_" + k: v for k, v in start.items()})
    out.update({"end_" + k: v for k, v in end.items()})
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "obs", obs.shape, "bytes", os.path.getsize(os.path.join(HERE, name)))


def main():
    rng = np.random.RandomState(2024)
    # pandaPush-v0 / pandaReach-v0, joint mode: 8 envs, 60 settle steps, 24 random-policy steps
    for task, name in ((TASK_PUSH, "panda_push.npz"), (TASK_REACH, "panda_reach.npz")):
        m, p = panda_task_setup(task)
        pose = common.sample_object_poses(8, seed=11)
        acts = rng.uniform(-1, 1, (24, 8, p.n_act)).astype(np.float32)
        save(name, *(lambda s, o, r, d, e: (s, acts, o, r, d, e))(*rollout(m, p, pose, common.targets_for(pose), 60, acts)))
    # iCubPush-v0 as registered (Cartesian control through the DLS IK) and iCubReach joint mode: 4 envs, 10 steps
    for task, use_ik, name in ((TASK_PUSH, 1, "icub_push_ik.npz"), (TASK_REACH, 0, "icub_reach_joint.npz")):
        m, p = icub_task_setup(task, control_arm='l', use_ik=use_ik, control_orientation=0, reward_type=0, goal_env=0)
        pose = icub_cases.object_poses(4, seed=5)
        tg = (pose[:, :3] + np.array([0.05, 0.05, -0.045], np.float32)).astype(np.float32)
        acts = rng.uniform(-1, 1, (10, 4, p.n_act)).astype(np.float32)
        save(name, *(lambda s, o, r, d, e: (s, acts, o, r, d, e))(*rollout(m, p, pose, tg, 3, acts)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Free-running comparison of the sweep-capped ("jammed") population: the same random rollout on the GPU (scheduled full-batch
launches, tail launch included) and on the CPU oracle; prints, every 200 steps, the number of sweep-capped environments, big
systems and the contact families present on both sides.  usage: python tools/jam_compare.py [B=4096] [steps=1000]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pybullet-robot-envs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from common import FAMILIES, TASK_PUSH, panda_task_setup, sample_object_poses, targets_for  # noqa: E402
from oracle import b2oracle  # noqa: E402
from pybullet_robot_envs.b2env.binding import B2Sim  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
m, p = panda_task_setup(TASK_PUSH)
orc = b2oracle.Oracle(m, p, B, nthreads=os.cpu_count() or 8)
sim = B2Sim(m, p, B, 0)
pose = sample_object_poses(B, 0)
tg = targets_for(pose, z=0.65)
orc.reset(pose, tg)
sim.reset_host(pose, tg)
orc.step(None, 200, 1, want_obs=False)
sim.step_host(None, 200, 1, want_obs=False)
rng = np.random.RandomState(0)


def summary(st, keys):
    cap = st[:, 1] >= 150
    fam = {f: int(((keys >= lo) & (keys < hi)).any(axis=1).sum()) for f, (lo, hi) in FAMILIES.items()}
    return "capped %4d  rows>25 %4d  overflow %4d  nan %d  mean iters %.1f  %s" % (
        int(cap.sum()), int((st[:, 3] > 25).sum()), int(((st[:, 0] & 2) > 0).sum()), int((st[:, 0] & 1).sum()), st[:, 1].mean(),
        {k: v for k, v in fam.items() if v})


for i in range(steps + 1):
    a = rng.uniform(-1, 1, (B, 7)).astype(np.float32)
    orc.step(a, 1, 0)
    sim.step_host(a, 1, 0)
    if i % 200 == 0:
        print("step %4d  gpu: %s" % (i, summary(sim.get("status"), sim.get("cache_key"))))
        print("           cpu: %s" % summary(orc.state["status"], orc.state["cache_key"]))
        dq = np.abs(sim.get("q") - orc.state["q"]).max(axis=1)
        print("           envs with |dq| > 1e-3: %d, > 0.1: %d" % (int((dq > 1e-3).sum()), int((dq > 0.1).sum())))

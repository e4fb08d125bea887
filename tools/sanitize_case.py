"""Small contact-rich workload for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py`
(70 envs, joint and Cartesian modes, overflow slots and 3-row-set systems in play)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "pybullet-robot-envs_b200"); sys.path.insert(0, "tests")
import numpy as np
from oracle import b2oracle
from common import *
import test_gpu_parity as T
from pybullet_robot_envs.b2env.binding import B2Sim
# joint mode, contact-rich states (slots, 3 row sets, coupled islands), IK mode, grasp: a few steps each
for kw, A in ((dict(), 7), (dict(use_ik=1), 6)):
    m, p = panda_task_setup(TASK_PUSH, **kw)
    B = 70   # not a multiple of the 16 envs per block: exercises the padding groups
    sim = B2Sim(m, p, B, 0)
    qs, poses = T._contact_rich_states(b2oracle, m, p, B, 3)
    tg = targets_for(poses) + np.array([0.3, 0, 0], np.float32)
    sim.reset_host(poses, tg)
    sim.set("q", qs); sim.set("mtarget", qs)
    rng = np.random.RandomState(0)
    if kw:
        sim.step_host(None, 1, 3, want_obs=False)
    for i in range(6):
        o, r, d = sim.step_host(rng.uniform(-1, 1, (B, A)).astype(np.float32), 1, 0)
    sim.step_host(None, 3, 1, want_obs=False)
    st = sim.get("status")
    print(kw, "rows max", st[:, 3].max(), "nan", (st[:, 0] & 1).sum(), "finite", np.isfinite(o).all())
    sim.close()
# every contact family of the collision stage (box-box at the rim / legs, finger pads, capsule through GJK / EPA, self pairs),
# stepped through the scheduled path as well (tail launch: B2ENV_SCHED_MIN lowered)
import os
m, p = panda_task_setup(TASK_PUSH)
qs, poses, fam = family_states(b2oracle, m, p, list(FAMILIES), per_family=6, seed=5)
for sched_min in ("1", "100000"):
    os.environ["B2ENV_SCHED_MIN"] = sched_min
    B = len(fam)
    sim = B2Sim(m, p, B, 0)
    sim.reset_host(poses, targets_for(poses))
    sim.set("q", qs); sim.set("mtarget", qs)
    rng = np.random.RandomState(1)
    for i in range(6):
        o, r, d = sim.step_host(rng.uniform(-1, 1, (B, 7)).astype(np.float32), 1, 0)
    st = sim.get("status")
    keys = sim.get("cache_key")
    seen = sorted(f for f, (lo, hi) in FAMILIES.items() if ((keys >= lo) & (keys < hi)).any())
    print("families, sched_min", sched_min, "rows max", st[:, 3].max(), "nan", (st[:, 0] & 1).sum(), "finite", np.isfinite(o).all(), "seen", len(seen))
    sim.close()
del os.environ["B2ENV_SCHED_MIN"]
# iCub tree kernel (one warp per env): joint mode, Cartesian mode, hand pressed onto the cube (coupled islands)
import icub_cases
from pybullet_robot_envs.b2env.model import icub_task_setup
for kw, A in ((dict(use_ik=0), 10), (dict(use_ik=1), 3)):
    m, p = icub_task_setup(TASK_PUSH, **kw)
    B = 9    # not a multiple of the 4 envs per block
    sim = B2Sim(m, p, B, 0)
    pose = icub_cases.object_poses(B, 0)
    pose[:4, 0] = [0.30, 0.33, 0.36, 0.30]; pose[:4, 1] = [0.26, 0.26, 0.27, 0.20]; pose[:, 2] = 0.651
    sim.reset_host(pose, pose[:, :3] + np.array([0.05, 0.05, 0], np.float32))
    sim.set("shaping", np.ones((B, 2), np.float32))
    rng = np.random.RandomState(0)
    if kw["use_ik"]:
        sim.step_host(None, 1, 3, want_obs=False)
        hp = sim.get("hand_pose"); hp[:, 2] = 0.70; sim.set("hand_pose", hp)
        sim.step_host(None, 30, 3, want_obs=False)
    for i in range(4):
        o, r, d = sim.step_host(rng.uniform(-1, 1, (B, A)).astype(np.float32), 1, 0)
    st = sim.get("status")
    print("icub", kw, "rows max", st[:, 3].max(), "contacts max", st[:, 2].max(), "nan", (st[:, 0] & 1).sum(), "finite", np.isfinite(o).all())
    sim.close()

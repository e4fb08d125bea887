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

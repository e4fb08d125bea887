#!/usr/bin/env python
"""Fuzz of the EMULATED iCub tree kernel (general collision path: cube near the rim of the table top, near a leg, on the floor,
random orientations and velocities) against the oracle: every step restarts from the oracle state; contact counts, keys (order
included) and overflow flags must be exact, converged environments within 2e-4.  Round 2: 864 env-steps, no mismatch
(518 with cube-vs-static-box manifolds, 333 with ground-plane vertices).  Second part: random arm postures (hands and forearms
anywhere between home and the joint limits: over the table, at its rim, pressed into it): 768 env-steps, no mismatch (55 with
sphere-vs-slab-side contacts, 70 sphere-table, 18 sphere-cube, the contact-list overflow once).
    python tools/fuzz_emu_icub_static.py"""
import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [ROOT, os.path.join(ROOT, 'pybullet-robot-envs_b200'), os.path.join(ROOT, 'tests')]
import numpy as np
from oracle import b2oracle
import icub_cases
from pybullet_robot_envs.b2env import binding
from pybullet_robot_envs.b2env.binding import B2Sim
from pybullet_robot_envs.b2env.model import TASK_PUSH, icub_task_setup
lib = binding.load_library(os.path.join(ROOT, 'tools', 'emu', 'libb2env_emu.so'))
m, p = icub_task_setup(TASK_PUSH, use_ik=0)
B = 24
sim = B2Sim(m, p, B, 0, lib=lib)
orc = b2oracle.Oracle(m, p, B, nthreads=4)
tot = 0; fams = {}
for seed in range(6):
    rng = np.random.RandomState(100 + seed)
    pose = icub_cases.object_poses(B, seed)
    for e in range(B):
        kind = e % 4
        if kind == 0:   pose[e, :3] = [0.1 + rng.uniform(-0.03, 0.04), rng.uniform(-0.45, 0.45), 0.651 + rng.uniform(0, 0.01)]       # near rim x
        elif kind == 1: pose[e, :3] = [rng.uniform(0.15, 0.4), 0.5 + rng.uniform(-0.04, 0.03), 0.651 + rng.uniform(0, 0.01)]        # near rim y
        elif kind == 2: pose[e, :3] = [0.2 + rng.uniform(-0.09, 0.09), -0.4 + rng.uniform(-0.09, 0.09), 0.026 + rng.uniform(0, 0.02)] # floor near leg
        else:           pose[e, :3] = [rng.uniform(-0.2, 0.05), rng.uniform(-0.3, 0.3), 0.03 + rng.uniform(0, 0.05)]                # floor
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = rng.uniform(0, 0.6) if kind < 2 else rng.uniform(0, np.pi)
        if kind < 2: ax = np.array([0, 0, 1.0]) if rng.rand() < 0.5 else ax
        pose[e, 3:6] = ax * np.sin(ang / 2); pose[e, 6] = np.cos(ang / 2)
    orc.reset(pose, pose[:, :3].copy())
    orc.state["obj_vel"][:, :3] = rng.uniform(-0.3, 0.3, (B, 3)).astype(np.float32)
    for i in range(6):
        icub_cases.sync(orc, sim)
        orc.step(None, 1, 1, want_obs=False); sim.step_host(None, 1, 1, want_obs=False)
        g, o = sim.get("status"), orc.state["status"]
        assert np.array_equal(g[:, 2:], o[:, 2:]), (seed, i, g[:, 2:], o[:, 2:])
        assert np.array_equal(g[:, 0] & 6, o[:, 0] & 6), (seed, i)
        assert np.array_equal(sim.get("cache_key"), orc.state["cache_key"]), (seed, i)
        conv = o[:, 1] < 150
        err = np.abs(sim.get("obj_pose") - orc.state["obj_pose"])
        assert err[conv].max() < 2e-4 and np.isfinite(sim.get("obj_pose")).all(), (seed, i, err.max())
        k = orc.state["cache_key"]
        for name, lo, hi in (("cube-table", 0, 8), ("cube-plane", 8, 16), ("cube-sbox", 4096, 12288)):
            fams[name] = fams.get(name, 0) + int(((k >= lo) & (k < hi)).any(axis=1).sum())
        tot += B
        orc.step(None, 3, 1, want_obs=False)
print("iCub static-world fuzz:", tot, "env-steps, keys / counts / flags exact;", fams, "overflowed", int(((orc.state['status'][:,0]&2)>0).sum()))
sim.close()

# ---- part 2: random arm postures (sphere proxies against the static world) ----
m, p = icub_task_setup(TASK_PUSH, use_ik=0)
B = 32
nd = m.n_dof
lo = np.array([m.lower[d] for d in range(nd)], np.float32); hi = np.array([m.upper[d] for d in range(nd)], np.float32)
home = np.array([m.home[d] for d in range(nd)], np.float32)
sim = B2Sim(m, p, B, 0, lib=lib)
orc = b2oracle.Oracle(m, p, B, nthreads=4)
tot = 0; fams = {}
for seed in range(8):
    rng = np.random.RandomState(200 + seed)
    pose = icub_cases.object_poses(B, seed); pose[:, 2] = 0.651
    orc.reset(pose, pose[:, :3].copy()); sim.reset_host(pose, pose[:, :3].copy())
    w = rng.uniform(0, 1, (B, 1)).astype(np.float32)
    q = (home + w * ((lo + (hi - lo) * rng.uniform(0.05, 0.95, (B, nd))) - home)).astype(np.float32)
    orc.state["q"][:] = q; orc.state["mtarget"][:] = q
    for i in range(3):
        icub_cases.sync(orc, sim)
        orc.step(None, 1, 1, want_obs=False); sim.step_host(None, 1, 1, want_obs=False)
        g, o = sim.get("status"), orc.state["status"]
        assert np.array_equal(g[:, 2:], o[:, 2:]), (seed, i, np.where((g[:, 2:] != o[:, 2:]).any(axis=1)))
        assert np.array_equal(g[:, 0] & 6, o[:, 0] & 6), (seed, i)
        assert np.array_equal(sim.get("cache_key"), orc.state["cache_key"]), (seed, i)
        assert np.isfinite(sim.get("q")).all()
        conv = o[:, 1] < 150
        if conv.any():
            err = np.abs(sim.get("q") - orc.state["q"])[conv].max()
            assert err < 5e-4, (seed, i, err)
        k = orc.state["cache_key"]
        for name, a, b in (("sphere-cube", 16, 32), ("sphere-table", 32, 64), ("sphere-sbox", 128, 256), ("sphere-plane", 256, 512)):
            fams[name] = fams.get(name, 0) + int(((k >= a) & (k < b)).any(axis=1).sum())
        tot += B
print("iCub random-posture fuzz:", tot, "env-steps exact;", fams, "overflow env-steps", int(((orc.state['status'][:,0]&2)>0).sum()))

#!/usr/bin/env python
"""Tail study for the Panda kernel: find a jammed environment (solver runs all sweeps over a big coupled system) deep in
a random rollout, then build a small batch with ONE copy of it per block (the other 15 environments of the block are
copies of an ordinary one) and time / profile single launches of that batch: the launch then lasts exactly as long as
the jammed solve.  Under `ncu --profile-from-start off` only the marked launch is captured.
usage: python tools/jam_profile.py [depth=900] [blocks=148]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
from pybullet_robot_envs.b2env import binding  # noqa: E402
from pybullet_robot_envs.envs import pandaPushGymEnv  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 900
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 148
B = 16384
dev = torch.device("cuda", 0)
env = pandaPushGymEnv(num_envs=B, device=0, renders=False, obj_pose_rnd_std=0.05, tg_pose_rnd_std=0, max_steps=100000)
env.seed(0)
env.reset()
sim = env._sim
gen = torch.Generator(device=dev)
gen.manual_seed(1234)
obs_t = torch.empty((B, sim.params.n_obs), device=dev)
rew_t = torch.empty(B, device=dev)
done_t = torch.empty(B, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
for i in range(depth):
    a = torch.rand((B, 7), generator=gen, device=dev) * 2 - 1
    sim.step(a, obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
torch.cuda.synchronize()
st = sim.get("status")
iters, nc, R = st[:, 1], st[:, 2], st[:, 3]
cost = iters.astype(np.int64) * R
print("depth %d: capped envs %d, rows max %d, mean iters %.1f, envs with cost>=2500: %d" % (
    depth, int((iters >= 150).sum()), int(R.max()), iters.mean(), int((cost >= 2500).sum())))
jam = int(np.argmax(cost))
norm = int(np.argmin(cost))
print("jammed env %d: iters %d rows %d contacts %d | ordinary env %d: iters %d rows %d" % (
    jam, iters[jam], R[jam], nc[jam], norm, iters[norm], R[norm]))
B2 = nblk * 16
idx = np.full(B2, norm)
idx[::16] = jam
s2 = binding.B2Sim(sim.model, sim.params, B2, 0)
saved = {}
for f in ("q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "cache_key", "cache_lam", "hand_pose"):
    saved[f] = sim.get(f)[idx]
    s2.set(f, saved[f])
o2 = torch.empty((B2, sim.params.n_obs), device=dev)
r2 = torch.empty(B2, device=dev)
d2 = torch.empty(B2, device=dev)
a_all = torch.rand((B, 7), generator=gen, device=dev) * 2 - 1
a2 = a_all[torch.as_tensor(idx, device=dev)].contiguous()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(4):
    for f, v in saved.items():
        s2.set(f, v)
    torch.cuda.synchronize()
    if rep == 2:
        torch.cuda.profiler.start()
    ev0.record()
    s2.step(a2, o2, r2, d2, 1, binding.MODE_ACTION, stream)
    ev1.record()
    torch.cuda.synchronize()
    if rep == 2:
        torch.cuda.profiler.stop()
    st2 = s2.get("status")
    print("launch %d: %.3f ms; jammed copies: iters %d status[2] %d status[3] %d (rows %d; %.0f cycles per row update at 1.965 GHz)" % (
        rep, ev0.elapsed_time(ev1), st2[0, 1], st2[0, 2], st2[0, 3], R[jam],
        ev0.elapsed_time(ev1) * 1e-3 * 1.965e9 / max(1, st2[0, 1] * R[jam])))

#!/usr/bin/env python
"""Active lanes of the warp at the stage boundaries of the iCub tree kernel (probe build -DPROFILE_STAGES -DPROFILE_WARM selected
with B2ENV_LIB): rolls iCubPush-v0 (registered kwargs) to the given depths and prints, per probe, how many environments saw how
many active lanes.  32 everywhere = the warp is converged; anything else = a straggler, every full-mask shuffle after it takes
its divergent path (DESIGN.md 4c').  usage: B2ENV_LIB=variants/libb2env_lanes.so python tools/lane_probe.py [depths=20,200]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
from pybullet_robot_envs.b2env import binding  # noqa: E402
import bench  # noqa: E402

depths = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,200").split(",")]
B = 16384
dev = torch.device("cuda", 0)
env = bench.make_env("icubpush", B, 0)
env.seed(0)
env.reset()
sim = env._sim
gen = torch.Generator(device=dev)
gen.manual_seed(1234)
na = sim.params.n_act
obs_t = torch.empty((B, sim.params.n_obs), device=dev)
rew_t = torch.empty(B, device=dev)
done_t = torch.empty(B, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
names = ["-", "action / IK", "targets stored", "dynamics", "collision", "rows + block + warm start", "affine arm island", "sweeps",
         "post-solve barrier", "integrate", "contact cache", "state store", "obs + reward"]
d = 0
for D in depths:
    while d < D:
        a = torch.rand((B, na), generator=gen, device=dev) * 2 - 1
        sim.step(a, obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
        d += 1
    torch.cuda.synchronize()
    am = sim.get("contacts").reshape(B, -1)[:, 32:48].astype(np.int64)
    st = sim.get("status")
    print("depth %d: mean sweeps %.1f, contacts max %d" % (d, st[:, 1].mean(), st[:, 2].max()))
    for k in range(1, 13):
        u, cnt = np.unique(am[:, k], return_counts=True)
        print("   active lanes after %-26s %s" % (names[k], ", ".join("%d: %d envs" % (x, y) for x, y in zip(u, cnt))))
import ctypes as C
try:   # loop sites (probe build only): visits / visits with a split warp
    out = (C.c_ulonglong * 32)()
    sim.lib.b2e_debug_lane_sites.argtypes = [C.c_void_p, C.c_int]
    if sim.lib.b2e_debug_lane_sites(out, 1) == 0:
        sites = ["Panda sweep", "Panda affine arm iteration", "Panda IK pass", "-", "iCub sweep", "iCub affine arm iteration", "iCub IK pass"]
        for k, n in enumerate(sites):
            if out[2 * k]:
                print("   loop site %-28s %d warp visits, %d with fewer than 32 active lanes" % (n, out[2 * k], out[2 * k + 1]))
except AttributeError:
    pass

#!/usr/bin/env python
"""Stage timing of the Panda kernel (instrumentation build -DPROFILE_STAGES, selected with B2ENV_LIB): rolls a
16384-env batch to the given depths and prints, per depth, the per-environment cycle stamps of one launch:
dynamics / collision / solve / wait at the post-solve barrier / rest, as mean, p50, p99 and max over the environments.
usage: B2ENV_LIB=variants/libb2env_stages.so python tools/stage_profile.py [depths=50,300,600,1000]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
from pybullet_robot_envs.b2env import binding  # noqa: E402
from pybullet_robot_envs.envs import pandaPushGymEnv  # noqa: E402

depths = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "50,300,600,1000").split(",")]
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
d = 0
names = ["fk+term", "dynamics", "collision", "barrier1", "solve", "barrier2(wait)", "rest"]
for D in depths:
    while d < D:
        a = torch.rand((B, 7), generator=gen, device=dev) * 2 - 1
        sim.step(a, obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
        d += 1
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a = torch.rand((B, 7), generator=gen, device=dev) * 2 - 1
    ev0.record()
    sim.step(a, obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
    ev1.record()
    torch.cuda.synchronize()
    d += 1
    c = sim.get("contacts").reshape(B, -1)[:, :8].astype(np.float64)
    st = sim.get("status")
    t = np.concatenate([c[:, :1], np.diff(c[:, :7], axis=1)], axis=1)   # per-stage cycles
    print("depth %d: launch %.3f ms (%d cycles at 1.965 GHz), mean iters %.1f, capped %d" % (
        d, ev0.elapsed_time(ev1), int(ev0.elapsed_time(ev1) * 1.965e6), st[:, 1].mean(), int((st[:, 1] >= 150).sum())))
    for k, n in enumerate(names):
        v = t[:, k]
        print("   %-16s mean %8.0f  p50 %8.0f  p99 %8.0f  max %8.0f" % (n, v.mean(), np.percentile(v, 50), np.percentile(v, 99), v.max()))
    am = sim.get("contacts").reshape(B, -1)[:, 32:46].astype(np.int64)
    if am.any():   # -DPROFILE_WARM build: active lanes of the warp at every stamp, over all environments (32 = converged)
        stamps = ["start", "after FK", "dynamics", "collision", "pre-solve barrier", "solve", "post-solve barrier", "integrate+cache",
                  "final FK", "filing", "state store", "obs+reward"]
        for k, n in enumerate(stamps):
            u, cnt = np.unique(am[:, k], return_counts=True)
            print("   active lanes at %-20s %s" % (n, ", ".join("%d: %d envs" % (a, b) for a, b in zip(u, cnt))))
        packs = sim.get("contacts").reshape(B, -1)[:, 17:20].astype(np.int64)
        labels = [("build entry", "rows + W", "Delassus block"), ("cache look-up", "warm-start apply", "sweeps"),
                  ("collision stage", "before slot claim", "after slot claim")]
        for j in range(3):
            for q in range(3):
                u, cnt = np.unique((packs[:, j] // (100 ** q)) % 100, return_counts=True)
                print("   active lanes after %-18s %s" % (labels[j][q], ", ".join("%d: %d envs" % (a, b) for a, b in zip(u, cnt))))
        continue
    cf = sim.get("contacts").reshape(B, -1)[:, 13:20].astype(np.float64)
    if cf.any():   # finer stamps of this build: parts of `rest` and of `solve`
        fin = {"rest: integrate+cache": cf[:, 0] - c[:, 5], "rest: final FK+termination": cf[:, 1] - cf[:, 0],
               "rest: cost-class filing": cf[:, 2] - cf[:, 1], "rest: state store": cf[:, 3] - cf[:, 2],
               "rest: observation+reward": c[:, 6] - cf[:, 3], "solve: rows + W": cf[:, 4], "solve: A + warm start": cf[:, 5] - cf[:, 4],
               "solve: affine arm": cf[:, 6] - cf[:, 5], "solve: sweeps + store": t[:, 4] - cf[:, 6]}
        for n, v in fin.items():
            print("   %-28s mean %8.0f  p50 %8.0f  p99 %8.0f" % (n, v.mean(), np.percentile(v, 50), np.percentile(v, 99)))
    tot = c[:, 6]
    print("   %-16s mean %8.0f  p50 %8.0f  p99 %8.0f  max %8.0f" % ("block lifetime", tot.mean(), np.percentile(tot, 50), np.percentile(tot, 99), tot.max()))
    top = np.argsort(-t[:, 4])[:6]
    envs, n_tail = sim.debug_sched_lists()
    print("   tail list: %d envs; slowest solves: %s" % (n_tail, ", ".join("env %d: %d cycles, %d sweeps x %d rows (%.0f cycles/row), block %d" % (
        e, t[e, 4], st[e, 1], st[e, 3], t[e, 4] / max(1, st[e, 1] * st[e, 3]), c[e, 7]) for e in top)))
    cs = sim.get("contacts").reshape(B, -1)[:, 8:13].astype(np.float64)
    if cs.any():   # -DPROFILE_SWEEP build: cycles per part of the sweeps
        for e in top[:4]:
            n = max(1, st[e, 1])
            cfe = sim.get("contacts").reshape(B, -1)[e, 13:20].astype(np.float64)
            print("   env %d: stage cycles fk %.0f dyn %.0f coll %.0f | solve %.0f = rows+W %.0f, A+warm start %.0f, affine %.0f, sweeps+store %.0f | "
                  "integrate+cache %.0f final FK %.0f filing %.0f store %.0f obs %.0f" % (
                      e, t[e, 0], t[e, 1], t[e, 2], t[e, 4], cfe[4], cfe[5] - cfe[4], cfe[6] - cfe[5], t[e, 4] - cfe[6],
                      cfe[0] - c[e, 5], cfe[1] - cfe[0], cfe[2] - cfe[1], cfe[3] - cfe[2], c[e, 6] - cfe[3]))
            ex = sim.get("contacts").reshape(B, -1)[e, 22:25]
            print("   env %d: warp mate env %d (rows %d, sweeps %d), warp runs the %s instantiation for %d generic rows, slots of the launch %d" % (
                e, int(ex[0]), st[int(ex[0]), 3], st[int(ex[0]), 1], "global-scratch" if ex[1] >= 1000 else "slot", int(ex[1]) % 1000, int(ex[2])))
            code = sim.get("contacts").reshape(B, -1)[e, 21]
            print("   env %d: launch role %d (1 main, 2 tail), system storage %d (>= 0 overflow slot, -1 global scratch, -3 own block), block %d" % (
                e, int(round(code / 100.0)), int(code - 100 * round(code / 100.0)), c[e, 7]))
            pre = sim.get("contacts").reshape(B, -1)[e, 20]
            print("   env %d: of motor+masks, before the serial motor rows (copies, block-form attempt): %.0f cycles per sweep" % (e, pre / n))
            print("   env %d sweep parts, cycles per sweep: motor+masks %.0f | non-friction rows %.0f | friction bounds %.0f | friction rows %.0f | residual %.0f  (rows %d, nc %d)" % (
                e, cs[e, 0] / n, cs[e, 1] / n, cs[e, 2] / n, cs[e, 3] / n, cs[e, 4] / n, st[e, 3], st[e, 2]))
        simple_e = np.nonzero(st[:, 3] == 21)[0][:3]
        for e in simple_e:
            print("   resting env %d: totals motor+masks %.0f | nf rows %.0f | fric bounds %.0f | fric rows %.0f | residual %.0f (status iters %d)" % (
                e, cs[e, 0], cs[e, 1], cs[e, 2], cs[e, 3], cs[e, 4], st[e, 1]))
    # solve cycles vs (iters, rows)
    it, R = st[:, 1], st[:, 3]
    simple = (R == 21)
    if simple.any():
        cyc = t[simple, 4]; its = it[simple]
        A = np.stack([its, np.ones_like(its)], axis=1).astype(np.float64)
        coef = np.linalg.lstsq(A, cyc, rcond=None)[0]
        print("   resting-cube envs (21 rows): solve cycles ~ %.0f + %.1f * sweeps  (=> %.1f cycles per row update over 12 cube rows)" % (coef[1], coef[0], coef[0] / 12))
try:   # -DPROFILE_WARM build: loop sites, visits / visits with a split warp
    import ctypes as C
    out = (C.c_ulonglong * 32)()
    sim.lib.b2e_debug_lane_sites.argtypes = [C.c_void_p, C.c_int]
    if sim.lib.b2e_debug_lane_sites(out, 1) == 0:
        for k, n in enumerate(["Panda sweep", "Panda affine arm iteration", "Panda IK pass"]):
            if out[2 * k]:
                print("   loop site %-28s %d warp visits, %d with fewer than 32 active lanes" % (n, out[2 * k], out[2 * k + 1]))
except AttributeError:
    pass

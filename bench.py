#!/usr/bin/env python
"""Headline benchmark: env-steps/s of random-policy pandaPush rollouts (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port, host cores)

One "step" = one env.step() of a whole batch (16384 envs per GPU): action -> motor targets ->
physics (dt = 1/240, <=150 PGS iterations) -> observation, reward, done, fused in one launch.
BASELINE.md protocol: 50 warm-up + 1000 timed steps after reset, done ignored.  The cost of a step grows with
the rollout depth (random-policy arms drift into the table / the cube; a few jammed contacts run all 150
sweeps), so the depth matters: with the default --warmup 400 --steps 8000 and 8 replicas of the batch stepped
round-robin (working set > L2) every replica does exactly 50 + 1000 steps.
`value`  : device-timed, K launches between two CUDA events on the launch stream, no host sync inside.
`e2e`    : same metric through the public Gym-style API with HOST numpy buffers (H2D of the
           actions and D2H of obs/reward/done inside the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))

B_ALG_PUSH = 968  # algorithmic bytes per pandaPush env-step
# dram__bytes_read.sum + dram__bytes_write.sum of one step_kernel launch at 16384 envs, from the ncu --set full
# capture summarised in profiles/r1_ncu_step_kernel_final_shallow.csv (8.29 MB + 0.29 MB); algorithmic: 15.86 MB
TRAFFIC_BYTES_PER_LAUNCH_16384 = 8.58e6
METRIC = "env-steps/sec PandaPush-v0 batch=16384"
WORKLOAD = "pandaPush-v0 joint mode, random policy U(-1,1)^7, post-reset state, done ignored"
# --workload icubpush: BASELINE.json config 4 (iCubPush-v0 as registered: Cartesian control through the DLS IK,
# 32-dof tree), an extra line next to the headline; 1492 algorithmic bytes per env-step
WORKLOADS = {
    "pandapush": dict(metric=METRIC, workload=WORKLOAD, n_act=7, n_obs=33, b_alg=B_ALG_PUSH, kernel="step_kernel",
                      traffic=TRAFFIC_BYTES_PER_LAUNCH_16384,
                      note="latency-bound by construction: ~40 sequential PGS sweeps per step, and a launch lasts as long as its "
                           "slowest environment (a jammed contact: 150 sweeps x ~45 rows x ~100 cycles, DESIGN.md 4c); DRAM "
                           "traffic per launch (ncu, profiles/) stays below the 15.9 MB algorithmic figure: state is "
                           "written back lazily from L2"),
    "icubpush": dict(metric="env-steps/sec iCubPush-v0", n_act=3, n_obs=34, b_alg=1492, kernel="tree_step_kernel<IK>",
                     traffic=12.67e6,   # profiles/r1_ncu_icub_tree_kernel_phase_locked.csv: 12.67 MB read + 0.5 KB write
                     note="instruction-issue / latency-bound: ~33 k warp instructions per env-step (32x32 inverse, DLS IK, "
                          "affine Gauss-Seidel sweeps), IPC 1.9 of 4 after phase-locking the blocks (DESIGN.md 4b); "
                          "DRAM traffic per launch 12.7 MB vs 24.4 MB algorithmic",
                     workload="iCubPush-v0 as registered (left arm, Cartesian xyz actions -> DLS IK -> 32 position motors), "
                              "random policy U(-1,1)^3, post-reset state, done ignored"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_cpu_arm_icub(B, seed0, nthreads):
    """CPU restatement set up like iCubPushGymEnv.reset(): IK of the home hand pose, settle, cube on the table."""
    from oracle import b2oracle
    from pybullet_robot_envs.b2env.model import TASK_PUSH, icub_task_setup
    m, p = icub_task_setup(TASK_PUSH, use_ik=1)
    orc = b2oracle.Oracle(m, p, B, nthreads=nthreads)
    rng = np.random.RandomState(seed0)
    pose = np.zeros((B, 7), np.float32)
    pose[:, 0] = 0.25 + rng.uniform(-0.05, 0.05, B)
    pose[:, 1] = rng.uniform(-0.05, 0.05, B)
    pose[:, 2] = 0.695
    yaw = rng.uniform(-np.pi / 4, np.pi / 4, B)
    pose[:, 5], pose[:, 6] = np.sin(yaw / 2), np.cos(yaw / 2)
    orc.reset(pose, pose[:, :3].copy())
    orc.state["shaping"][:] = 1
    orc.step(None, 1, 3, want_obs=False)
    orc.step(None, 101, 1, want_obs=False)
    tg = orc.state["obj_pose"][:, :3].copy()
    tg[:, 0] = np.clip(tg[:, 0] + 0.05, 0.17, 0.38)
    tg[:, 1] = np.clip(tg[:, 1] + 0.05, -0.3, 0.3)
    orc.state["target"][:] = tg
    return orc


def make_cpu_arm(B, seed0, nthreads, workload="pandapush"):
    """The CPU restatement (oracle port) set up on the same workload: reset + settle, like env.reset()."""
    if workload == "icubpush":
        return make_cpu_arm_icub(B, seed0, nthreads)
    from oracle import b2oracle
    from pybullet_robot_envs.b2env.model import TASK_PUSH, panda_task_setup
    from pybullet_robot_envs.gym_compat import seeding
    m, p = panda_task_setup(TASK_PUSH)
    orc = b2oracle.Oracle(m, p, B, nthreads=nthreads)
    pose = np.zeros((B, 7), np.float32)
    for i in range(B):
        rng, _ = seeding.np_random(seed0 + i)
        pose[i, 0] = np.clip(0.45 + rng.uniform(-0.05, 0.05), 0.35, 0.55)
        pose[i, 1] = np.clip(rng.uniform(-0.05, 0.05), -0.25, 0.25)
        pose[i, 2] = 0.695
        yaw = rng.uniform(-np.pi / 4, np.pi / 4)
        pose[i, 5], pose[i, 6] = np.sin(yaw / 2), np.cos(yaw / 2)
    tg = pose[:, :3].copy()
    orc.reset(pose, tg)
    orc.step(None, 101, 1, want_obs=False)
    tg = orc.state["obj_pose"][:, :3].copy()
    tg[:, 0] = np.clip(tg[:, 0] + 0.05, 0.37, 0.58)
    tg[:, 1] = np.clip(tg[:, 1] + 0.05, -0.3, 0.3)
    orc.state["target"][:] = tg
    return orc


def time_cpu(orcs, steps, warmup, seed=1234):
    """Round-robin over the replicas in `orcs` (same depth profile as the GPU arm)."""
    rng = np.random.RandomState(seed)
    B = orcs[0].B
    n = len(orcs)
    na = orcs[0].params.n_act
    for i in range(warmup):
        orcs[i % n].step(rng.uniform(-1, 1, (B, na)).astype(np.float32), 1, 0)
    t0 = time.perf_counter()
    for i in range(steps):   # i.i.d. actions drawn per step (the draw is ~2 % of a step)
        orcs[(warmup + i) % n].step(rng.uniform(-1, 1, (B, na)).astype(np.float32), 1, 0)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt


def clone_cpu_arm(orc, n):
    from oracle import b2oracle
    out = [orc]
    for _ in range(n - 1):
        o2 = b2oracle.Oracle(orc.model, orc.params, orc.B, nthreads=orc.nthreads)
        for k, v in orc.state.items():
            o2.state[k][...] = v
        out.append(o2)
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  PyBullet is not installable
    here, so this is the oracle port on all host cores, a bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = args.cpu_batch
    WL = WORKLOADS[args.workload]
    orcs = clone_cpu_arm(make_cpu_arm(B, 0, cores, args.workload), args.replicas)
    rate, dt = time_cpu(orcs, args.steps, max(args.warmup, 3))
    line = {
        "impl": "reference", "metric": WL["metric"], "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WL["workload"], "sample": "%d envs per step-call on %d host threads, %d replicas round-robin (rollout depth per replica %d steps)" % (B, cores, args.replicas, (args.steps + max(args.warmup, 3)) // args.replicas),
                   "note": "CPU restatement (oracle port), not PyBullet: pybullet is absent from this image; "
                           "as-shipped reference is additionally capped at 240 steps/s/process by time.sleep"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d envs x %d steps" % (B, args.steps)},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8000)
    ap.add_argument("--warmup", type=int, default=400)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="environments per GPU")
    ap.add_argument("--cpu-batch", type=int, default=2048)
    ap.add_argument("--e2e-steps", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", type=int, default=8, help="independent batches stepped round-robin (working set > L2)")
    ap.add_argument("--workload", default="pandapush", choices=sorted(WORKLOADS), help="pandapush = the BASELINE.json metric")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pybullet_robot_envs.envs import iCubPushGymEnv, pandaPushGymEnv

    WL = WORKLOADS[args.workload]
    NA, NO, B_ALG = WL["n_act"], WL["n_obs"], WL["b_alg"]
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    if args.workload == "icubpush":   # registered kwargs of iCubPush-v0 (reference __init__.py:19-31)
        env = iCubPushGymEnv(num_envs=B, device=local, renders=False, use_IK=1, control_arm='l', control_orientation=0,
                             obj_pose_rnd_std=0.05, tg_pose_rnd_std=0, max_steps=1000, reward_type=0)
    else:
        env = pandaPushGymEnv(num_envs=B, device=local, renders=False, obj_pose_rnd_std=0.05, tg_pose_rnd_std=0,
                              max_steps=1000)
    env.seed(rank * B)   # env i of rank r is the reference env seeded r*B + i
    env.reset()
    sim = env._sim
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    NACT = W + K               # one i.i.d. action batch per launch (no recycling: a periodic sequence is a drift, not a random walk)
    actions = torch.rand((NACT, B, NA), generator=gen, device=dev, dtype=torch.float32)
    actions.mul_(2.0).sub_(1.0)
    returns = torch.zeros(B, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # Working set larger than L2: NREP independent replicas of the batch (each ~19 MB of state) are stepped
    # round-robin, so a replica's state has been evicted from the 126 MB L2 by the time its turn comes again.
    # No flush kernels, no host synchronisation inside the timed region: K launches between two CUDA events.
    from pybullet_robot_envs.b2env import binding
    NREP = args.replicas
    sims = [sim]
    for r in range(1, NREP):
        s2 = binding.B2Sim(sim.model, sim.params, B, local)
        for f in ("q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "cache_key", "cache_lam", "hand_pose", "shaping"):
            s2.set(f, sim.get(f))
        sims.append(s2)
    obs_t = torch.empty((B, sim.params.n_obs), device=dev)
    rew_t = torch.empty(B, device=dev)
    done_t = torch.empty(B, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)      # nvidia-smi needs a moment before the first sample; the timed region is seconds long
    for i in range(W):
        sims[i % NREP].step(actions[i % NACT], obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = sum(s_.launch_count() for s_ in sims)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(K):
        sims[(W + i) % NREP].step(actions[(W + i) % NACT], obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)   # ONE launch of step_kernel
        returns += rew_t
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sum(s_.launch_count() for s_ in sims) - l0
    dev_ms = ev0.elapsed_time(ev1)
    # kernel-only duration for the roofline: events directly around a few launches (no accumulation kernel)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(K, 32))]
    extra_actions = torch.rand((len(kev), B, NA), generator=gen, device=dev, dtype=torch.float32) * 2 - 1
    for i, (a_, b_) in enumerate(kev):
        a_.record()
        sims[i % NREP].step(extra_actions[i], obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
        b_.record()
    barrier()
    kernel_ms = float(np.median([a_.elapsed_time(b_) for a_, b_ in kev]))
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    status = sims[0].get("status")
    mean_iters = float(status[:, 1].mean())
    nan_flags = int(sum((s_.get("status")[:, 0] & 1).sum() for s_ in sims))
    for s_ in sims[1:]:
        s_.close()

    # ---- end-to-end through the public API with host buffers: same protocol (fresh reset, 50 warm-up steps,
    #      then the timed steps; every step copies the actions H2D from page-locked memory and reads obs /
    #      reward / done back D2H) ----
    depth = (W + K) // NREP
    n_warm_e = min(50, depth // 2)
    Ke = max(8, min(args.e2e_steps, depth - n_warm_e))
    NH = n_warm_e + Ke
    host_actions = sim.pinned_array((NH, B, NA))          # the policy's outputs live in page-locked host memory
    rs = np.random.RandomState(99 + rank)
    for i in range(NH):                                   # i.i.d. per step
        host_actions[i] = rs.uniform(-1, 1, (B, NA)).astype(np.float32)
    env.reset()
    for i in range(n_warm_e):
        env.step(host_actions[i])
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        o, r, d, _ = env.step(host_actions[n_warm_e + i])        # H2D actions, launch, D2H obs/reward/done
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_rate = B * world * Ke / float(te.item())

    # ---- episode returns gathered over NCCL (the only collective of the path) ----
    if world > 1:
        allret = [torch.empty_like(returns) for _ in range(world)]
        dist.all_gather(allret, returns)
        mean_return = float(torch.stack(allret).mean().item())
    else:
        mean_return = float(returns.mean().item())

    if rank == 0:
        peak, which = peaks()
        value = B * world * K / (dev_ms_max * 1e-3)
        per_gpu_rate = B * K / (dev_ms_max * 1e-3)
        # dominant kernel: algorithmic bytes per launch / its AVERAGE launch duration over the timed region (the K
        # launches run back to back between the two events; the 3 us return-accumulation add is included)
        achieved = B * B_ALG / (dev_ms_max / K * 1e-3) / 1e9
        line = {
            "metric": WL["metric"], "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WL["workload"], "envs_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                       "dt": 1.0 / 240, "solver_iters_max": 150, "residual_tol": 1e-7,
                       "mean_pgs_iters_last_step": mean_iters, "nan_flags": nan_flags,
                       "l2": "no flush: %d replicas of the batch stepped round-robin, working set %.0f MB > 126 MB L2" % (NREP, NREP * B * 1.2e-3),
                       "rollout_depth_per_replica": (W + K) // NREP, "protocol": "BASELINE.md: 50 warm-up + 1000 timed steps after reset per batch, done ignored",
                       "wall_ms_per_step": 1e3 * t_wall / K, "kernel_ms_at_final_depth": kernel_ms, "mean_episode_return": mean_return},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (WL["traffic"] if B == 16384 else None), "traffic_unit": "bytes per launch (ncu)", "peak_source": which, "alg_bytes_per_env_step": B_ALG,
                         "kernel": WL["kernel"], "kernel_ms_avg": dev_ms_max / K, "kernel_ms_at_final_depth": kernel_ms,
                         "note": WL["note"]},
            "e2e": {"value": e2e_rate, "unit": "env-steps/s", "h2d_bytes_per_step": B * NA * 4,
                    "d2h_bytes_per_step": B * (NO + 2) * 4, "steps": Ke, "warmup": n_warm_e,
                    "note": "fresh reset, then warm-up + timed steps through env.step() with host arrays: H2D copy of the actions, results stored by the kernel straight into page-locked host arrays (zero-copy D2H; B2ENV_ZEROCOPY=0 for explicit copies)"},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            orc = make_cpu_arm(args.cpu_batch, 0, cores, args.workload)
            depth = (W + K) // NREP
            n_warm = min(50, depth // 2)
            n_timed = max(10, depth - n_warm)
            rate, dt = time_cpu([orc], n_timed, n_warm)
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": "%d envs x (%d warm-up + %d timed) steps from reset (%.1f s): same rollout depth as one GPU replica"
                                              % (args.cpu_batch, n_warm, n_timed, dt)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

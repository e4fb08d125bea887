#!/usr/bin/env python
"""Headline benchmark: env-steps/s of random-policy pandaPush rollouts (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port, host cores)

One "step" = one env.step() of a whole batch (16384 envs per GPU): action -> motor targets ->
physics (dt = 1/240, <=150 PGS iterations) -> observation, reward, done, fused in one launch.

BASELINE.md protocol: 50 warm-up + 1000 timed steps after reset, done ignored.  The cost of a step GROWS with
the rollout depth (random-policy arms drift into the table / the cube; a few jammed contacts run all 150
sweeps), so a number is only meaningful with its depth.  Every invocation therefore samples the protocol
window, whatever K is: NREP replicas of the batch are stepped round-robin (working set > L2) and, BEFORE the
timed region, replica r is rolled untimed to the depth at which its share of the K timed steps starts, the
starts being spread evenly over 50..1050.  With the default --warmup 400 --steps 8000 every replica does
exactly 50 warm-up + 1000 timed steps from reset (no pre-roll: `config.protocol` says so); with the driver's
--steps 20 --warmup 5 replica r is timed at depth 50 + 142 r (+3 steps).  The CPU arm gets the same profile.

`value`  : device-timed, K launches between two CUDA events on the launch stream, no host sync inside.
`e2e`    : same metric through the public Gym-style API with HOST numpy buffers (H2D of the
           actions and D2H of obs/reward/done inside the timed region), sampled over the same depth window.
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

PROTO_WARM, PROTO_STEPS = 50, 1000
# Algorithmic bytes per env-step: 4 (A + O + 2) + 8 S
WORKLOADS = {
    "pandapush": dict(metric="env-steps/sec PandaPush-v0 batch=16384", batch=16384, n_act=7, n_obs=33, b_alg=968,
                      kernel="step_kernel<false>",
                      workload="pandaPush-v0 joint mode, random policy U(-1,1)^7, post-reset state, done ignored",
                      note="latency-bound by construction: ~40 sequential PGS sweeps per step, and a launch lasts as long as its "
                           "slowest environment (a jammed contact runs all 150 sweeps over ~45 rows, DESIGN.md 4c)"),
    "pandareach": dict(metric="env-steps/sec PandaReach-v0 batch=4096", batch=4096, n_act=7, n_obs=30, b_alg=932,
                       kernel="step_kernel<false>",
                       workload="pandaReach-v0 as registered (joint mode, cube_small on the table), random policy U(-1,1)^7, "
                                "post-reset state, done ignored",
                       note="BASELINE.json config 2; 4096 envs = 256 blocks of 16 on 148 SMs (< 1 wave at 2 blocks / SM): "
                            "latency-bound, the GPU is not full"),
    "pandagrasp": dict(metric="env-steps/sec PandaGrasp-v0", batch=16384, n_act=8, n_obs=33, b_alg=972,
                       kernel="step_kernel<false>",
                       workload="PandaGrasp-v0 (new task, no reference env): 7 joint increments + gripper command, finger "
                                "force 10 / maxVelocity 1, random policy U(-1,1)^8, post-reset state, done ignored",
                       note="BASELINE.json config 5; throughput only (the reference has no grasp env)"),
    "icubpush": dict(metric="env-steps/sec iCubPush-v0", batch=16384, n_act=3, n_obs=34, b_alg=1492,
                     kernel="tree_step_kernel<true>",
                     workload="iCubPush-v0 as registered (left arm, Cartesian xyz actions -> DLS IK -> 32 position motors), "
                              "random policy U(-1,1)^3, post-reset state, done ignored",
                     note="instruction-issue / latency-bound: ~33 k warp instructions per env-step (32x32 inverse, DLS IK, "
                          "affine Gauss-Seidel sweeps); BASELINE.json config 4"),
}
STATE_FIELDS = ("q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "cache_key", "cache_lam",
                "hand_pose", "shaping")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def ncu_traffic(workload, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from an `ncu --set full`
    capture of this benchmark (replicas stepped round-robin, i.e. the regime of the timed loop); summaries and their
    provenance live in profiles/traffic.json.  None when no capture exists for this workload / batch."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f).get(workload, {}).get(str(batch))
        return (float(e["bytes_per_launch"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def depth_plan(K, W, nrep):
    """Per-replica (pre-roll, first timed depth, timed steps): the timed windows of the replicas are spread evenly
    over the protocol window [50, 1050).  Replica r takes warm-up steps r, r+nrep, ... and timed steps likewise."""
    wr = [len(range(r, W, nrep)) for r in range(nrep)]                       # warm-up steps of replica r
    kr = [len(range((r - W) % nrep, K, nrep)) for r in range(nrep)]          # timed steps of replica r
    kmax = max(kr)
    plan = []
    for r in range(nrep):
        if kmax >= PROTO_STEPS:
            start = max(PROTO_WARM, wr[r])
        elif nrep == 1:
            start = PROTO_WARM + (PROTO_STEPS - kmax) // 2
        else:
            start = PROTO_WARM + int(round(r * (PROTO_STEPS - kmax) / (nrep - 1.0)))
        start = max(start, wr[r])
        plan.append({"preroll": start - wr[r], "start": start, "timed": kr[r]})
    return plan


def plan_is_protocol(plan):
    return all(p["start"] == PROTO_WARM and p["timed"] == PROTO_STEPS for p in plan)


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


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/b2oracle.c), set up on the same workload as the GPU arm

def _cube_poses(B, seed0, x0, gym_seeding):
    pose = np.zeros((B, 7), np.float32)
    if gym_seeding:   # env i is the reference env seeded seed0 + i (world_env.py:145-176 draw order x, y, yaw)
        from pybullet_robot_envs.gym_compat import seeding
        for i in range(B):
            rng, _ = seeding.np_random(seed0 + i)
            pose[i, 0] = x0 + rng.uniform(-0.05, 0.05)
            pose[i, 1] = rng.uniform(-0.05, 0.05)
            yaw = rng.uniform(-np.pi / 4, np.pi / 4)
            pose[i, 5], pose[i, 6] = np.sin(yaw / 2), np.cos(yaw / 2)
    else:
        rng = np.random.RandomState(seed0)
        pose[:, 0] = x0 + rng.uniform(-0.05, 0.05, B)
        pose[:, 1] = rng.uniform(-0.05, 0.05, B)
        yaw = rng.uniform(-np.pi / 4, np.pi / 4, B)
        pose[:, 5], pose[:, 6] = np.sin(yaw / 2), np.cos(yaw / 2)
    pose[:, 2] = 0.695
    return pose


def make_cpu_arm(B, seed0, nthreads, workload="pandapush"):
    """The CPU restatement set up like env.reset() of the workload: settle, cube on the table, target."""
    from oracle import b2oracle
    from pybullet_robot_envs.b2env.model import TASK_GRASP, TASK_PUSH, TASK_REACH, icub_task_setup, panda_task_setup
    if workload == "icubpush":
        m, p = icub_task_setup(TASK_PUSH, use_ik=1)
        orc = b2oracle.Oracle(m, p, B, nthreads=nthreads)
        pose = _cube_poses(B, seed0, 0.25, False)
        orc.reset(pose, pose[:, :3].copy())
        orc.state["shaping"][:] = 1
        orc.step(None, 1, 3, want_obs=False)
        orc.step(None, 101, 1, want_obs=False)
        tg = orc.state["obj_pose"][:, :3].copy()
        tg[:, 0] = np.clip(tg[:, 0] + 0.05, 0.17, 0.38)
        tg[:, 1] = np.clip(tg[:, 1] + 0.05, -0.3, 0.3)
        orc.state["target"][:] = tg
        return orc
    task = {"pandapush": TASK_PUSH, "pandareach": TASK_REACH, "pandagrasp": TASK_GRASP}[workload]
    m, p = panda_task_setup(task)
    if task == TASK_GRASP:   # as pandaGraspGymEnv: 7 joint increments + gripper command, finger force 10 / maxVelocity 1
        p.n_act = 8
        for d in (7, 8):
            m.max_force[d] = 10.0
            m.max_vel[d] = 1.0
    orc = b2oracle.Oracle(m, p, B, nthreads=nthreads)
    pose = _cube_poses(B, seed0, 0.45, B <= 4096)
    orc.reset(pose, pose[:, :3].copy())
    orc.step(None, 101, 1, want_obs=False)
    tg = orc.state["obj_pose"][:, :3].copy()
    if task == TASK_GRASP:
        tg[:, 2] += 0.1
    else:
        tg[:, 0] = np.clip(tg[:, 0] + 0.05, 0.37, 0.58)
        tg[:, 1] = np.clip(tg[:, 1] + 0.05, -0.3, 0.3)
    orc.state["target"][:] = tg
    return orc


def make_cpu_arm_staggered(B, nthreads, workload, starts, seed=4321):
    """One CPU batch of B envs whose slices sit at the rollout depths `starts` (the GPU arm's replica depths):
    slice r is built from reset and rolled `starts[r]` random steps, then the slices are merged, so every timed
    step over the merged batch samples the whole protocol window.  Returns (oracle, pre-roll seconds)."""
    from oracle import b2oracle
    n = len(starts)
    per = max(1, B // n)
    t0 = time.perf_counter()
    parts = []
    for r, d in enumerate(starts):
        o = make_cpu_arm(per, 1000 * r, nthreads, workload)
        rng = np.random.RandomState(seed + r)
        na = o.params.n_act
        for _ in range(d):
            o.step(rng.uniform(-1, 1, (per, na)).astype(np.float32), 1, 0, want_obs=False)
        parts.append(o)
    if n == 1:
        return parts[0], time.perf_counter() - t0
    big = b2oracle.Oracle(parts[0].model, parts[0].params, per * n, nthreads=nthreads)
    for r, o in enumerate(parts):
        for k, v in o.state.items():
            big.state[k][r * per:(r + 1) * per] = v
    return big, time.perf_counter() - t0


def time_cpu(orc, steps, warmup, seed=1234):
    rng = np.random.RandomState(seed)
    B, na = orc.B, orc.params.n_act
    for _ in range(warmup):
        orc.step(rng.uniform(-1, 1, (B, na)).astype(np.float32), 1, 0)
    t0 = time.perf_counter()
    for _ in range(steps):   # i.i.d. actions drawn per step (the draw is ~2 % of a step)
        orc.step(rng.uniform(-1, 1, (B, na)).astype(np.float32), 1, 0)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt


def run_cpu_arm(workload, K, W, nrep, cpu_batch, cores, budget_env_steps=12e6):
    """The CPU arm on a BOUNDED sample with the GPU arm's depth profile.  Returns (rate, seconds, steps run, sample).
    Sampled mode (fewer than 1000 timed steps per replica): one batch whose `nrep` slices are pre-rolled (untimed) to
    depths spread over 50..1050, then W warm-up + K timed step-calls over the whole batch.  Full-protocol mode: every
    env does 50 warm-up + 1000 timed steps from reset — the rollout of ONE of the GPU arm's replicas (the others repeat
    the same depth profile).  The batch is bounded by a work budget (~20 s on 16 threads); the per-env cost does not
    depend on the batch (environments never interact)."""
    plan = depth_plan(K, W, nrep)
    kmax = max(p["timed"] for p in plan)
    if kmax >= PROTO_STEPS:
        B = int(min(cpu_batch, max(256, budget_env_steps / (PROTO_WARM + PROTO_STEPS) // 256 * 256)))
        orc = make_cpu_arm(B, 0, cores, workload)
        rate, dt = time_cpu(orc, PROTO_STEPS, PROTO_WARM)
        return rate, dt, PROTO_STEPS, ("%d envs per step-call on %d host threads, %d warm-up + %d timed steps from reset (%.1f s): the "
                                       "rollout of one GPU replica" % (B, cores, PROTO_WARM, PROTO_STEPS, dt))
    if nrep == 1:
        starts = [PROTO_WARM + (PROTO_STEPS - K) // 2]
    else:
        starts = [PROTO_WARM + int(round(r * (PROTO_STEPS - K) / (nrep - 1.0))) for r in range(nrep)]
    cost_per_env = sum(starts) / float(len(starts)) + K + W     # steps per env incl. pre-roll
    B = int(min(cpu_batch, max(256, budget_env_steps / cost_per_env // 256 * 256)))
    orc, pre_s = make_cpu_arm_staggered(B, cores, workload, starts)
    rate, dt = time_cpu(orc, K, W)
    return rate, dt, K, ("%d envs per step-call on %d host threads; %d slices pre-rolled (untimed, %.1f s) to rollout depths %s, "
                         "then %d warm-up + %d timed steps (%.1f s)" % (orc.B, cores, len(starts), pre_s, starts, W, K, dt))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  PyBullet is not installable
    here, so this is the oracle port on all host cores, a bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    WL = WORKLOADS[args.workload]
    batch = args.batch or WL["batch"]
    K, W = args.steps, max(args.warmup, 3)
    rate, dt, k_run, sample = run_cpu_arm(args.workload, K, W, max(1, args.replicas), args.cpu_batch or batch, cores)
    line = {
        "impl": "reference", "metric": WL["metric"], "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / k_run,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WL["workload"], "envs_per_gpu": batch, "global_batch": batch * (world if args.scaling == "weak" else 1),
                   "dt": 1.0 / 240, "solver_iters_max": 150, "residual_tol": 1e-7,
                   "sample": sample,
                   "note": "CPU restatement (oracle port), not PyBullet: pybullet is absent from this image (SURVEY 8c); per-env cost is "
                           "independent of the batch (envs never interact), so a bounded batch measures the same env-steps/s; the "
                           "as-shipped reference is additionally capped at 240 steps/s/process by time.sleep(1/240)"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def make_env(workload, B, device):
    from pybullet_robot_envs.envs import iCubPushGymEnv, pandaPushGymEnv, pandaReachGymEnv
    from pybullet_robot_envs.envs.panda_envs.panda_grasp_gym_env import pandaGraspGymEnv
    if workload == "icubpush":   # registered kwargs of iCubPush-v0 (reference __init__.py:19-31)
        return iCubPushGymEnv(num_envs=B, device=device, renders=False, use_IK=1, control_arm='l', control_orientation=0,
                              obj_pose_rnd_std=0.05, tg_pose_rnd_std=0, max_steps=1000, reward_type=0)
    if workload == "pandareach":   # reference __init__.py:47-56
        return pandaReachGymEnv(num_envs=B, device=device, renders=False, use_IK=0, obj_pose_rnd_std=0.05, max_steps=1000)
    if workload == "pandagrasp":
        return pandaGraspGymEnv(num_envs=B, device=device, renders=False, obj_pose_rnd_std=0.05, max_steps=1000)
    return pandaPushGymEnv(num_envs=B, device=device, renders=False, obj_pose_rnd_std=0.05, tg_pose_rnd_std=0,
                           max_steps=1000)   # reference __init__.py:58-68


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8000)
    ap.add_argument("--warmup", type=int, default=400)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="environments per GPU (weak) / in total (strong); default: the workload's")
    ap.add_argument("--cpu-batch", type=int, default=0, help="CPU arm: envs per step-call (default: the batch; bounded by a work budget)")
    ap.add_argument("--e2e-steps", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", type=int, default=8, help="independent batches stepped round-robin (working set > L2)")
    ap.add_argument("--workload", default="pandapush", choices=sorted(WORKLOADS), help="pandapush = the BASELINE.json metric")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch envs per GPU; strong: --batch envs in total, sharded contiguously over the ranks")
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
    from pybullet_robot_envs.b2env import binding
    from pybullet_robot_envs.b2env.shard import shard_range

    WL = WORKLOADS[args.workload]
    NA, NO, B_ALG = WL["n_act"], WL["n_obs"], WL["b_alg"]
    batch = args.batch or WL["batch"]
    if args.scaling == "strong":
        lo, hi = shard_range(batch, rank, world)
        B, seed0, global_batch = hi - lo, lo, batch
    else:
        B, seed0, global_batch = batch, rank * batch, batch * world
    K, W = args.steps, max(args.warmup, 3)
    NREP = max(1, args.replicas)
    env = make_env(args.workload, B, local)
    env.seed(seed0)   # env i of this rank is the reference env seeded seed0 + i
    env.reset()
    sim = env._sim
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    returns = torch.zeros(B, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # Working set larger than L2: NREP independent replicas of the batch (each ~1.2 KB of state per env) are stepped
    # round-robin, so a replica's state has been evicted from the 126 MB L2 by the time its turn comes again.
    # No flush kernels, no host synchronisation inside the timed region: K launches between two CUDA events.
    sims = [sim]
    for r in range(1, NREP):
        s2 = binding.B2Sim(sim.model, sim.params, B, local)
        for f in STATE_FIELDS:
            s2.set(f, sim.get(f))
        sims.append(s2)
    obs_t = torch.empty((B, sim.params.n_obs), device=dev)
    rew_t = torch.empty(B, device=dev)
    done_t = torch.empty(B, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def roll(s_, n, chunk=64):
        """n untimed random-policy steps of one replica (actions drawn on the device in chunks)."""
        done_ = 0
        while done_ < n:
            m_ = min(chunk, n - done_)
            a_ = torch.rand((m_, B, NA), generator=gen, device=dev, dtype=torch.float32) * 2 - 1
            for j in range(m_):
                s_.step(a_[j], obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
            done_ += m_
        torch.cuda.synchronize(dev)

    plan = depth_plan(K, W, NREP)
    t_pre0 = time.perf_counter()
    for r in range(NREP):
        roll(sims[r], plan[r]["preroll"])
    preroll_s = time.perf_counter() - t_pre0
    NACT = W + K               # one i.i.d. action batch per launch (no recycling: a periodic sequence is a drift, not a random walk)
    actions = torch.rand((NACT, B, NA), generator=gen, device=dev, dtype=torch.float32)
    actions.mul_(2.0).sub_(1.0)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)      # nvidia-smi needs a moment before the first sample
    for i in range(W):
        sims[i % NREP].step(actions[i], obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = sum(s_.launch_count() for s_ in sims)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(K):
        sims[(W + i) % NREP].step(actions[W + i], obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)   # ONE launch of the step kernel
        returns += rew_t
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sum(s_.launch_count() for s_ in sims) - l0
    dev_ms = ev0.elapsed_time(ev1)
    # kernel-only durations (events directly around single launches, no accumulation kernel), one per replica = one per depth
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2 * NREP)]
    extra_actions = torch.rand((len(kev), B, NA), generator=gen, device=dev, dtype=torch.float32) * 2 - 1
    for i, (a_, b_) in enumerate(kev):
        a_.record()
        sims[(W + K + i) % NREP].step(extra_actions[i], obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
        b_.record()
    barrier()
    kms = [a_.elapsed_time(b_) for a_, b_ in kev]
    kernel_ms_by_replica = [round(float(np.mean([kms[i] for i in range(len(kev)) if (W + K + i) % NREP == r])), 4) for r in range(NREP)]
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    mean_iters = float(np.mean([s_.get("status")[:, 1].mean() for s_ in sims]))
    capped = int(sum((s_.get("status")[:, 1] >= 150).sum() for s_ in sims))
    nan_flags = int(sum((s_.get("status")[:, 0] & 1).sum() for s_ in sims))
    for s_ in sims[1:]:
        s_.close()

    # ---- end-to-end through the public API with host buffers.  Every timed step copies the actions H2D from
    #      page-locked memory and reads obs / reward / done back; the timed steps are taken in windows spread over the
    #      same protocol depths (one env object: between the windows it is rolled on untimed with device actions) ----
    Ke = max(8, min(args.e2e_steps, K))
    n_win = 1 if Ke >= PROTO_STEPS else min(8, Ke)
    ke = Ke // n_win
    Ke = ke * n_win
    if n_win == 1:
        e_starts = [PROTO_WARM]
    else:
        e_starts = [PROTO_WARM + int(round(j * (PROTO_STEPS - ke) / (n_win - 1.0))) for j in range(n_win)]
    host_actions = sim.pinned_array((Ke, B, NA))          # the policy's outputs live in page-locked host memory
    rs = np.random.RandomState(99 + rank)
    for i in range(Ke):                                   # i.i.d. per step
        host_actions[i] = rs.uniform(-1, 1, (B, NA)).astype(np.float32)
    env.zero_copy_results = True     # opt-in: step() returns views of the page-locked result arrays (valid until the next step)
    env.reset()
    depth_e, e2e_s, i_act = 0, 0.0, 0
    for j in range(n_win):
        roll(sim, max(0, e_starts[j] - depth_e - 1))
        env.step(host_actions[i_act])                     # one untimed step through the host path (first-touch, caches)
        depth_e = e_starts[j]
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            o, r_, d_, _info = env.step(host_actions[i_act])   # H2D actions, launch, D2H obs/reward/done
            i_act += 1
        torch.cuda.synchronize(dev)
        e2e_s += time.perf_counter() - t0
        depth_e += ke
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_total = B * Ke
    if world > 1:
        tb = torch.tensor([e2e_total], device=dev, dtype=torch.float64)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        e2e_total = float(tb.item())
    e2e_rate = e2e_total / float(te.item())

    # ---- episode returns gathered over NCCL (the only collective of the path) ----
    if world > 1:
        from pybullet_robot_envs.b2env.shard import gather_returns
        mean_return = float(gather_returns(returns).mean().item())
    else:
        mean_return = float(returns.mean().item())

    if rank == 0:
        peak, which = peaks()
        value = global_batch * K / (dev_ms_max * 1e-3)
        # dominant kernel: algorithmic bytes per launch / its AVERAGE launch duration over the timed region (the K
        # launches run back to back between the two events; the 3 us return-accumulation add is included)
        achieved = B * B_ALG / (dev_ms_max / K * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(args.workload, B)
        cfg = {"workload": WL["workload"], "envs_per_gpu": B, "global_batch": global_batch, "parallelism": "dp%d" % world,
               "dt": 1.0 / 240, "solver_iters_max": 150, "residual_tol": 1e-7,
               "mean_pgs_iters_last_step": mean_iters, "sweep_capped_envs_last_step": capped, "nan_flags": nan_flags,
               "l2": "no flush: %d replicas of the batch stepped round-robin, working set %.0f MB > 126 MB L2" % (NREP, NREP * B * 1.2e-3),
               "rollout_depth_per_replica": [[p["start"], p["start"] + p["timed"]] for p in plan],
               "preroll": "untimed, %.1f s: replica r rolled to its first timed depth so that the timed steps sample the "
                          "BASELINE.md window 50..1050 whatever --steps is" % preroll_s,
               "wall_ms_per_step": 1e3 * t_wall / K, "kernel_ms_by_replica": kernel_ms_by_replica, "mean_episode_return": mean_return}
        if plan_is_protocol(plan):
            cfg["protocol"] = "BASELINE.md: 50 warm-up + 1000 timed steps after reset per batch, done ignored"
        else:
            cfg["protocol"] = "sampled: %d timed steps per replica at the depths above (BASELINE.md window 50..1050), done ignored" % plan[0]["timed"]
        line = {
            "metric": WL["metric"], "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)", "traffic_source": traffic_src,
                         "peak_source": which, "alg_bytes_per_env_step": B_ALG, "alg_bytes_per_launch": B * B_ALG,
                         "kernel": WL["kernel"], "kernel_ms_avg": dev_ms_max / K, "note": WL["note"]},
            "e2e": {"value": e2e_rate, "unit": "env-steps/s", "h2d_bytes_per_step": B * NA * 4,
                    "d2h_bytes_per_step": B * (NO + 2) * 4, "steps": Ke, "windows": [[s_, s_ + ke] for s_ in e_starts],
                    "note": "env.step() with host arrays, timed in windows over the same rollout depths: H2D copy of the actions from "
                            "page-locked memory, results stored by the kernel straight into page-locked host arrays "
                            "(env.zero_copy_results = True, the opt-in view path; the default returns copies)"},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            kc = K if plan[0]["timed"] >= PROTO_STEPS else 60     # sampled mode: 60 step-calls over the staggered batch
            rate, dt, k_run, sample = run_cpu_arm(args.workload, kc, min(W, 5), NREP, args.cpu_batch or B, cores)
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Roll a 16384-env pandaPush batch `depth` random-policy steps from reset, then mark ONE more launch for the profiler
(`ncu --profile-from-start off ... python tools/deep_capture.py 1000`): the deep-state capture of the step kernel without
paying ncu's per-launch interception for the whole rollout."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
from pybullet_robot_envs.b2env import binding  # noqa: E402
from pybullet_robot_envs.envs import pandaPushGymEnv  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
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
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(depth + 3):
    a = torch.rand((B, 7), generator=gen, device=dev) * 2 - 1
    if i == depth:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ev0.record()
    sim.step(a, obs_t, rew_t, done_t, 1, binding.MODE_ACTION, stream)
    ev1.record()
    if i >= depth:
        torch.cuda.synchronize()
        print("launch at depth %d: %.3f ms" % (i, ev0.elapsed_time(ev1)))
    if i == depth:
        torch.cuda.profiler.stop()
st = sim.get("status")
print("capped envs %d, rows max %d, mean sweeps %.1f" % (int((st[:, 1] >= 150).sum()), int(st[:, 3].max()), st[:, 1].mean()))

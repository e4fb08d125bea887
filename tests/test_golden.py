"""Golden trajectories (tests/golden/*.npz, written by tests/golden/make_golden.py from the fp64 oracle build).
They do not pin parity with PyBullet (nothing can here: SURVEY.md §8c) — they freeze the statement of DESIGN.md §2 at
double precision.  `-m "not gpu"`: the fp32 oracle replays them; `-m gpu`: the CUDA path replays them through the C-ABI.

Tolerances (free-running fp32 vs fp64 over 24 / 10 steps, written here): scaled observation 5e-3 (the standardised
end-effector velocity entries amplify the solver's 3e-4 velocity residual by 1/0.03), reward 2e-3, final joint
angles 2e-4, done flags and step counters exact."""
import os

import numpy as np
import pytest

from pybullet_robot_envs.b2env.model import TASK_PUSH, TASK_REACH, icub_task_setup, panda_task_setup

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "panda_push.npz": lambda: panda_task_setup(TASK_PUSH),
    "panda_reach.npz": lambda: panda_task_setup(TASK_REACH),
    "icub_push_ik.npz": lambda: icub_task_setup(TASK_PUSH, control_arm='l', use_ik=1, control_orientation=0, reward_type=0, goal_env=0),
    "icub_reach_joint.npz": lambda: icub_task_setup(TASK_REACH, control_arm='l', use_ik=0, control_orientation=0, reward_type=0, goal_env=0),
}
START = ("q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "cache_key", "cache_lam", "hand_pose", "shaping")


def replay(g, set_field, step, get_field):
    for k in START:
        set_field(k, g["start_" + k])
    worst = {"obs": 0.0, "reward": 0.0}
    for t in range(g["actions"].shape[0]):
        obs, rew, done = step(g["actions"][t])
        worst["obs"] = max(worst["obs"], float(np.abs(obs - g["obs"][t]).max()))
        worst["reward"] = max(worst["reward"], float(np.abs(rew - g["reward"][t]).max()))
        np.testing.assert_array_equal(done, g["done"][t], err_msg="done flags at step %d" % t)
    worst["q"] = float(np.abs(get_field("q") - g["end_q"]).max())
    worst["obj"] = float(np.abs(get_field("obj_pose") - g["end_obj_pose"]).max())
    np.testing.assert_array_equal(get_field("counters"), g["end_counters"])
    return worst


def check(worst, name):
    assert worst["obs"] <= 5e-3, (name, worst)
    assert worst["reward"] <= 2e-3, (name, worst)
    assert worst["q"] <= 2e-4, (name, worst)
    assert worst["obj"] <= 2e-4, (name, worst)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_fp32_replays_golden(name, oracle_lib):
    g = np.load(os.path.join(HERE, name))
    m, p = CASES[name]()
    orc = oracle_lib.Oracle(m, p, g["actions"].shape[1], nthreads=2)

    def set_field(k, v):
        orc.state[k][...] = v
    check(replay(g, set_field, lambda a: orc.step(a, 1, 0), lambda k: orc.state[k]), name)


@pytest.mark.parametrize("name", sorted(CASES))
def test_emulated_kernels_replay_golden(name):
    """The CUDA kernel source compiled for the host (tools/emu, test infrastructure) replays the golden trajectories."""
    import shutil
    import subprocess
    from pybullet_robot_envs.b2env import binding
    from pybullet_robot_envs.b2env.binding import B2Sim
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "tools", "emu", "libb2env_emu.so")
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    srcs = [os.path.join(root, "pybullet-robot-envs_b200", "csrc", f) for f in ("b2env.cu", "b2env_tree.cuh")]
    srcs += [os.path.join(root, "tools", "emu", "cuda_emu.h"), os.path.join(root, "include", "b2env.h")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        subprocess.check_call(["bash", os.path.join(root, "tools", "emu", "build.sh")])
    lib = binding.load_library(so)
    g = np.load(os.path.join(HERE, name))
    m, p = CASES[name]()
    sim = B2Sim(m, p, g["actions"].shape[1], 0, lib=lib)
    try:
        check(replay(g, sim.set, lambda a: sim.step_host(a, 1, 0), sim.get), name)
    finally:
        sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_replays_golden(name):
    from pybullet_robot_envs.b2env.binding import B2Sim
    g = np.load(os.path.join(HERE, name))
    m, p = CASES[name]()
    sim = B2Sim(m, p, g["actions"].shape[1], 0)
    try:
        check(replay(g, sim.set, lambda a: sim.step_host(a, 1, 0), sim.get), name)
    finally:
        sim.close()

"""iCub on the B200: the warp-per-environment tree kernel against the oracle through the C-ABI (libb2env.so), and the
iCub gym surface.  Cases and tolerances: tests/icub_cases.py (the same cases run on the CPU emulation of the kernel
source in tests/test_emu_kernels.py)."""
import numpy as np
import pytest

import icub_cases
from pybullet_robot_envs.b2env.model import TASK_PUSH, TASK_REACH, icub_task_setup

pytestmark = pytest.mark.gpu


@pytest.fixture()
def make_sim():
    from pybullet_robot_envs.b2env.binding import B2Sim
    sims = []

    def mk(m, p, B):
        s = B2Sim(m, p, B, 0)
        assert b"EMULATION" not in s.lib.b2e_version()
        sims.append(s)
        return s
    yield mk
    for s in sims:
        s.close()


def test_joint_mode(make_sim, oracle_lib):
    icub_cases.single_step_parity(make_sim, oracle_lib, B=130, use_ik=0, n_hold=3, n_act=12)


def test_ik_mode_registered_push(make_sim, oracle_lib):
    icub_cases.single_step_parity(make_sim, oracle_lib, B=130, use_ik=1, n_hold=3, n_act=12)


def test_ik_orientation_right_arm(make_sim, oracle_lib):
    icub_cases.single_step_parity(make_sim, oracle_lib, B=64, use_ik=1, control_orientation=1, arm='r', n_hold=2, n_act=8,
                                  reward_type=1, goal_env=1)


def test_reach(make_sim, oracle_lib):
    icub_cases.single_step_parity(make_sim, oracle_lib, B=64, use_ik=1, task=TASK_REACH, n_hold=2, n_act=8)


def test_hand_contacts(make_sim, oracle_lib):
    icub_cases.hand_contact_parity(make_sim, oracle_lib, n_check=10)


@pytest.mark.xfail(strict=False, reason="general collision path of the tree kernel: written after round 2's GPU time was spent, "
                   "verified on the host emulation only (tests/test_emu_kernels.py); its first GPU execution is the driver's "
                   "round-end run.  XPASS = verified on the B200; a difference here must not hide the other GPU tests")
def test_static_world_general_path(make_sim, oracle_lib):
    """Rim / legs / floor for the iCub path (tests/icub_cases.py: static_world_parity), CUDA vs oracle."""
    icub_cases.static_world_parity(make_sim, oracle_lib)


def test_free_running_rollout(make_sim, oracle_lib):
    """120 random Cartesian steps without re-synchronising: the hand stays away from the cube (targets move by
    <= 0.6 m in 120 steps only in the worst case; envs whose hand touched something are excluded), so the
    trajectories stay together: |dq| < 2e-3, reward < 1e-2."""
    B = 96
    m, p = icub_task_setup(TASK_PUSH, use_ik=1)
    sim = make_sim(m, p, B)
    orc = oracle_lib.Oracle(m, p, B, nthreads=8)
    pose = icub_cases.object_poses(B, 3)
    tg = (pose[:, :3] + np.array([0.05, 0.05, -0.045], np.float32)).astype(np.float32)
    orc.reset(pose, tg)
    sim.reset_host(pose, tg)
    for s in (sim, ):
        s.set("shaping", np.ones((B, 2), np.float32))
    orc.state["shaping"][:] = 1
    orc.step(None, 1, 3, want_obs=False); sim.step_host(None, 1, 3, want_obs=False)
    orc.step(None, 100, 1, want_obs=False); sim.step_host(None, 100, 1, want_obs=False)
    rng = np.random.RandomState(5)
    touched = np.zeros(B, bool)
    for i in range(120):
        a = rng.uniform(-1, 1, (B, 3)).astype(np.float32)
        a[:, 2] = np.abs(a[:, 2]) * 0.3           # drift upwards: keep the hand off the table
        o_obs, o_rew, o_done = orc.step(a, 1, 0)
        g_obs, g_rew, g_done = sim.step_host(a, 1, 0)
        touched |= (orc.state["cache_key"] >= 16).any(axis=1) | (sim.get("cache_key") >= 16).any(axis=1)
    ok = ~touched
    assert ok.sum() >= B // 2
    assert np.abs(sim.get("q") - orc.state["q"])[ok].max() < 2e-3
    assert np.abs(g_rew - o_rew)[ok].max() < 1e-2
    np.testing.assert_array_equal(sim.get("counters")[ok], orc.state["counters"][ok])
    assert (sim.get("status")[:, 0] & 1).sum() == 0   # no NaN flag


def test_gym_surface_batched():
    import pybullet_robot_envs  # noqa: F401
    from pybullet_robot_envs import gym_compat as gym
    B = 64
    env = gym.make('iCubPush-v0', num_envs=B)
    u = env.unwrapped if hasattr(env, "unwrapped") else env
    env.seed(0)
    obs = env.reset()
    assert obs.shape == (B, 34)
    raw = u._physics_client_id.observe()[3]
    np.testing.assert_allclose(raw[:, :3], np.tile([0.3, 0.26, 0.8], (B, 1)), atol=3e-3)
    np.testing.assert_allclose(raw[:, 21], 0.65, atol=2e-3)
    rng = np.random.RandomState(0)
    for i in range(20):
        o, r, d, info = env.step(rng.uniform(-1, 1, (B, 3)).astype(np.float32))
    assert o.shape == (B, 34) and r.shape == (B,) and d.shape == (B,)
    assert np.isfinite(o).all() and np.isfinite(r).all()
    # per-env reset of a subset (row f1) on the tree kernel
    ids = np.array([1, 5, 17], np.int32)
    o2 = u.reset(ids)
    assert o2.shape == (3, 34)
    raw = u._physics_client_id.observe()[3]
    np.testing.assert_allclose(raw[ids, :3], np.tile([0.3, 0.26, 0.8], (3, 1)), atol=3e-3)
    env.close()
    for name, na, no in (('iCubReach-v0', 3, 31), ('iCubPushGoal-v0', 6, 34)):
        e = gym.make(name, num_envs=8)
        o = e.reset()
        a = np.zeros((8, na), np.float32)
        o, r, d, info = e.step(a)
        if isinstance(o, dict):
            assert o['observation'].shape == (8, no) and o['achieved_goal'].shape == (8, 3) and 'is_success' in info
        else:
            assert o.shape == (8, no)
        e.close()

"""GPU parity of the Cartesian (IK) control mode and the GoalEnv rules, through the C-ABI."""
import numpy as np
import pytest

from common import TASK_PUSH, copy_state_to_gpu, panda_task_setup, sample_object_poses, targets_for

pytestmark = pytest.mark.gpu


def _mk(oracle_lib, B, **kw):
    from pybullet_robot_envs.b2env.binding import B2Sim
    m, p = panda_task_setup(TASK_PUSH, **kw)
    orc = oracle_lib.Oracle(m, p, B, nthreads=8)
    sim = B2Sim(m, p, B, 0)
    pose = sample_object_poses(B, 4)
    tg = targets_for(pose, z=0.65)
    tg[:, 0] += 0.25
    orc.reset(pose, tg)
    sim.reset_host(pose, tg)
    return m, p, orc, sim


@pytest.mark.parametrize("orient", [1, 0])
def test_ik_mode_single_step_parity(oracle_lib, orient):
    """IK solution (motor targets), commanded hand pose and the physics step agree with the oracle when
    both start every step from the oracle's state.  Tolerances: targets 2e-4 rad (the IK loop stops on a
    1e-3 m residual; one fp32 ulp can flip the last iteration), joints 2e-5."""
    B = 128
    m, p, orc, sim = _mk(oracle_lib, B, use_ik=1, ik_orientation=orient)
    A = 6 if orient else 3
    assert p.n_act == A
    orc.step(None, 1, 3, want_obs=False)       # robot.reset(): IK of the home hand pose
    copy_state_to_gpu(orc, sim)
    orc.step(None, 150, 1, want_obs=False)
    sim.step_host(None, 150, 1, want_obs=False)
    rng = np.random.RandomState(0)
    flips = 0
    for i in range(60):
        copy_state_to_gpu(orc, sim)
        a = rng.uniform(-1, 1, (B, A)).astype(np.float32)
        o_obs, o_rew, o_done = orc.step(a, 1, 0)
        g_obs, g_rew, g_done = sim.step_host(a, 1, 0)
        np.testing.assert_allclose(sim.get("hand_pose"), orc.state["hand_pose"], atol=1e-6)
        dt = np.abs(sim.get("mtarget") - orc.state["mtarget"]).max(axis=1)
        flips += int((dt > 2e-4).sum())
        ok = dt <= 2e-4
        assert np.abs(sim.get("q") - orc.state["q"])[ok].max() < 2e-5
        assert np.abs(g_rew - o_rew)[ok].max() < 1e-3
        np.testing.assert_array_equal(g_done, o_done)
    assert flips <= 0.01 * 60 * B, flips        # an extra/missing last IK iteration must be rare


def test_cartesian_env_api():
    from pybullet_robot_envs.envs import pandaPushGymEnv
    env = pandaPushGymEnv(use_IK=1, num_envs=16, obj_pose_rnd_std=0.05)
    assert env.action_space.shape == (6,) and env.observation_space.shape == (33,)
    env.seed(1)
    obs = env.reset()
    assert obs.shape == (16, 33)
    raw, _ = env.get_extended_observation()
    # after the reset the hand sits near the home hand pose (0.2, 0, 0.8) (panda_env.py:85-88); not exactly:
    # the IK ignores joint limits and joint 4 of its solution lies beyond its lower limit
    assert np.all(np.abs(raw[:, 0] - 0.2) < 0.12) and np.all(np.abs(raw[:, 2] - 0.8) < 0.12)
    a = np.zeros((16, 6), np.float32)
    a[:, 2] = -1.0
    for _ in range(50):
        obs, rew, done, info = env.step(a)
    hp = env._physics_client_id.get("hand_pose")
    np.testing.assert_allclose(hp[:, 2], 0.8 - 50 * 0.005, atol=1e-5)
    raw2, _ = env.get_extended_observation()
    assert np.all(raw2[:, 2] < raw[:, 2] - 0.1)                 # the hand went down
    with pytest.raises(AssertionError):
        env.step(np.zeros((16, 7), np.float32))
    env.close()


def test_goal_env():
    import pybullet_robot_envs  # noqa: F401
    from pybullet_robot_envs import gym
    env = gym.make("pandaPushGoal-v0", num_envs=8)
    env.seed(0)
    obs = env.reset()
    assert set(obs) == {"observation", "achieved_goal", "desired_goal"}
    assert obs["observation"].shape == (8, 33) and obs["achieved_goal"].shape == (8, 3)
    o, r, d, info = env.step(np.zeros((8, 7), np.float32))
    # registered tg_pose_rnd_std = 0: the goal is 0.0707 m from the cube, inside the 0.1 m radius
    assert np.all(info["is_success"]) and np.all(r == 0) and np.all(d)
    np.testing.assert_array_equal(env.compute_reward(o["achieved_goal"], o["desired_goal"], info), r)
    far = o["desired_goal"] + np.array([0.3, 0, 0])
    np.testing.assert_array_equal(env.compute_reward(o["achieved_goal"], far, info), -np.ones(8, np.float32))
    env.close()


def test_goal_rules_parity(oracle_lib):
    B = 64
    m, p, orc, sim = _mk(oracle_lib, B, goal_env=1, max_steps=3)
    orc.step(None, 101, 1, want_obs=False)
    sim.step_host(None, 101, 1, want_obs=False)
    tg = orc.state["obj_pose"][:, :3].copy()
    tg[: B // 2, 0] += 0.05          # half the batch succeeds at once
    tg[B // 2:, 0] += 0.3
    orc.state["target"][:] = tg
    copy_state_to_gpu(orc, sim)
    rng = np.random.RandomState(1)
    for i in range(6):
        a = rng.uniform(-1, 1, (B, 7)).astype(np.float32)
        o_obs, o_rew, o_done = orc.step(a, 1, 0)
        g_obs, g_rew, g_done = sim.step_host(a, 1, 0)
        np.testing.assert_array_equal(g_rew, o_rew)
        np.testing.assert_array_equal(g_done, o_done)
        np.testing.assert_array_equal(sim.get("counters"), orc.state["counters"])


def test_grasp_task_parity_and_env(oracle_lib):
    """PandaGrasp: finger rows carry force 10 / maxVelocity 1, so the motor-island shortcut regularly hits an
    active bound and must fall back to the serial sweep — both must match the oracle."""
    from common import TASK_GRASP
    from pybullet_robot_envs.b2env.binding import B2Sim
    B = 128
    m, p = panda_task_setup(TASK_GRASP)
    orc = oracle_lib.Oracle(m, p, B, nthreads=8)
    sim = B2Sim(m, p, B, 0)
    pose = sample_object_poses(B, 9)
    tg = targets_for(pose, z=0.75)
    orc.reset(pose, tg)
    sim.reset_host(pose, tg)
    orc.step(None, 101, 1, want_obs=False)
    sim.step_host(None, 101, 1, want_obs=False)
    rng = np.random.RandomState(2)
    for i in range(80):
        copy_state_to_gpu(orc, sim)
        a = rng.uniform(-1, 1, (B, 8)).astype(np.float32)
        if i % 20 < 10:
            a[:, 7] = np.sign(a[:, 7])                # saturating open/close commands
        o_obs, o_rew, o_done = orc.step(a, 1, 0)
        g_obs, g_rew, g_done = sim.step_host(a, 1, 0)
        assert np.abs(sim.get("q") - orc.state["q"]).max() < 2e-5, i
        assert np.abs(sim.get("qd") - orc.state["qd"]).max() < 5e-3, i
        np.testing.assert_allclose(sim.get("mtarget"), orc.state["mtarget"], atol=1e-6)
        np.testing.assert_allclose(g_rew, o_rew, atol=1e-3)
        np.testing.assert_array_equal(g_done, o_done)
    sim.close()
    import pybullet_robot_envs  # noqa: F401
    from pybullet_robot_envs import gym
    env = gym.make("PandaGrasp-v0", num_envs=32)
    assert env.action_space.shape == (8,) and env.observation_space.shape == (33,)
    env.seed(0)
    obs = env.reset()
    raw, _ = env.get_extended_observation()
    np.testing.assert_allclose(raw[:, 32], raw[:, 20] + 0.1, atol=1e-5)     # target = object + 0.1 m lift
    o, r, d, _ = env.step(env.action_space.sample()[None, :].repeat(32, 0))
    assert o.shape == (32, 33) and np.all(d == 0)
    env.close()

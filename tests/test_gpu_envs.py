"""The reference-facing Gym surface on the CUDA backend: ids, shapes, dtypes, reset/step semantics
(reference panda_push_gym_env.py:105-115, :244-255)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _unscale(space, x):
    return space.low + 0.5 * (x + 1.0) * (space.high - space.low)


def test_single_env_matches_reference_shapes_and_kats():
    import pybullet_robot_envs  # noqa: F401
    from pybullet_robot_envs import gym
    env = gym.make("pandaPush-v0")
    assert env.observation_space.shape == (33,) and env.action_space.shape == (7,)
    assert env.seed(0)[0] == 0
    obs = env.reset()
    assert obs.shape == (33,) and obs.dtype == np.float64
    raw = _unscale(env.observation_space, obs)
    # C.1: motors hold the home pose through the settle -> EE position and joint angles
    np.testing.assert_allclose(raw[0:3], (0.365860, -0.036740, 0.986768), atol=2e-4)
    np.testing.assert_allclose(raw[9:18], (0, -0.54, 0, -2.6, -0.30, 2.0, 1.0, 0.02, 0.02), atol=2e-4)
    # C.4: seed 0 object pose (x, y, yaw); cube settled on the table
    np.testing.assert_allclose(raw[18:20], (0.4054360056, 0.0465390937), atol=2e-3)
    assert abs(raw[20] - 0.65) < 1e-3
    assert abs(raw[23] - 0.2084304499) < 5e-3
    # target = object + (0.05, 0.05, 0) (panda_push_gym_env.py:342-344)
    np.testing.assert_allclose(raw[30:33], raw[18:21] + np.array([0.05, 0.05, 0]), atol=1e-5)
    o, r, d, info = env.step(env.action_space.sample())
    assert o.shape == (33,) and np.shape(r) == () and np.shape(d) == () and info == {}
    # registered config succeeds immediately
    assert float(d) == 1.0 and 1000 < float(r) < 1100
    with pytest.raises(AssertionError):
        env.step(np.zeros(6))
    env.close()


def test_hooks_agree_with_fused_step():
    from pybullet_robot_envs.envs import pandaPushGymEnv
    env = pandaPushGymEnv(num_envs=32, obj_pose_rnd_std=0.05, tg_pose_rnd_std=0.2)
    env.seed(5)
    np.random.seed(0)
    env.reset()
    a = np.random.RandomState(1).uniform(-1, 1, (32, 7)).astype(np.float32)
    obs, rew, done, _ = env.step(a)
    raw, lim = env.get_extended_observation()
    assert raw.shape == (32, 33) and len(lim) == 33
    np.testing.assert_allclose(2 * (raw - env.observation_space.low) / (env.observation_space.high - env.observation_space.low) - 1,
                               obs, atol=1e-5)
    np.testing.assert_allclose(env._compute_reward(), rew, atol=1e-6)
    np.testing.assert_array_equal(env._termination(), done)
    r_obs, r_lim = env._robot.get_observation()
    w_obs, w_lim = env._world.get_observation()
    assert r_obs.shape == (32, 18) and w_obs.shape == (32, 6) and len(r_lim) == 18 and len(w_lim) == 6
    np.testing.assert_allclose(raw[:, :18], r_obs)
    np.testing.assert_allclose(raw[:, 18:24], w_obs)
    env.close()


def test_overridden_hook_takes_the_reference_call_sequence():
    from pybullet_robot_envs.envs import pandaReachGymEnv

    class Shaped(pandaReachGymEnv):
        def _compute_reward(self):
            return super()._compute_reward() * 2.0

    base = pandaReachGymEnv(num_envs=8, obj_pose_rnd_std=0.05)
    sub = Shaped(num_envs=8, obj_pose_rnd_std=0.05)
    for e in (base, sub):
        e.seed(3)
        e.reset()
    a = np.random.RandomState(2).uniform(-1, 1, (8, 7)).astype(np.float32)
    o1, r1, d1, _ = base.step(a)
    o2, r2, d2, _ = sub.step(a)
    assert not sub._fused and base._fused
    np.testing.assert_allclose(o2, o1, atol=1e-5, rtol=1e-5)   # host float64 scaling vs device fp32 scaling
    np.testing.assert_allclose(r2, 2.0 * r1, rtol=1e-6)
    np.testing.assert_array_equal(d1, d2)
    base.close(); sub.close()


def test_batched_env_equals_independent_single_envs():
    """Env i of a batch seeded s is the reference env seeded s + i."""
    from pybullet_robot_envs.envs import pandaReachGymEnv
    B = 4
    batch = pandaReachGymEnv(num_envs=B, obj_pose_rnd_std=0.05)
    batch.seed(10)
    ob = batch.reset()
    a = np.random.RandomState(4).uniform(-1, 1, (B, 7)).astype(np.float32)
    for _ in range(5):
        ob, rb, db, _ = batch.step(a)
    for i in range(B):
        single = pandaReachGymEnv(num_envs=1, obj_pose_rnd_std=0.05)
        single.seed(10 + i)
        single.reset()
        for _ in range(5):
            o, r, d, _ = single.step(a[i])
        np.testing.assert_allclose(o, ob[i], atol=1e-6)
        np.testing.assert_allclose(r, rb[i], atol=1e-6)
        single.close()
    batch.close()


def test_device_tensor_step_path():
    import torch
    from pybullet_robot_envs.envs import pandaPushGymEnv
    env = pandaPushGymEnv(num_envs=256, obj_pose_rnd_std=0.05)
    env.seed(0)
    env.reset()
    a = torch.rand((256, 7), device="cuda") * 2 - 1
    obs, rew, done, _ = env.step(a)
    assert obs.is_cuda and obs.shape == (256, 33) and rew.shape == (256,) and done.shape == (256,)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    env.close()


def test_subset_reset_and_auto_reset():
    """reset(env_ids) re-runs the reference reset sequence for some envs of the batch only."""
    from pybullet_robot_envs.envs import pandaPushGymEnv
    B = 16
    env = pandaPushGymEnv(num_envs=B, obj_pose_rnd_std=0.05, tg_pose_rnd_std=0.2)
    env.seed(7)
    np.random.seed(1)
    env.reset()
    rng = np.random.RandomState(0)
    for _ in range(30):
        env.step(rng.uniform(-1, 1, (B, 7)).astype(np.float32))
    c = env._physics_client_id
    before = {k: c.get(k).copy() for k in ("q", "qd", "obj_pose", "obj_vel", "target", "counters", "mtarget")}
    ids = np.array([1, 5, 6], np.int32)
    obs = env.reset(ids)
    assert obs.shape == (3, 33)
    after = {k: c.get(k) for k in before}
    keep = np.setdiff1d(np.arange(B), ids)
    for k in before:
        np.testing.assert_array_equal(after[k][keep], before[k][keep], err_msg=k)     # untouched envs: bit-identical
    home = np.array([0, -0.54, 0, -2.6, -0.30, 2.0, 1.0, 0.02, 0.02], np.float32)
    np.testing.assert_allclose(after["q"][ids], np.tile(home, (3, 1)), atol=2e-5)
    assert np.all(after["counters"][ids] == 0)
    assert np.all(np.abs(after["obj_pose"][ids, 2] - 0.65) < 1e-3)                   # new cube settled on the table
    assert np.all(np.abs(after["obj_vel"][ids]) < 5e-3)
    assert np.all(np.abs(after["obj_pose"][ids, 0] - 0.45) <= 0.0501)
    # auto-reset: finished envs restart inside step()
    env.auto_reset = True
    tg = c.get("target")
    tg[:4] = c.get("obj_pose")[:4, :3] + np.array([0.02, 0, 0], np.float32)           # envs 0-3 succeed next step
    c.set("target", tg)
    o, r, d, _ = env.step(np.zeros((B, 7), np.float32))
    assert np.all(d[:4] == 1) and np.all(r[:4] > 1000)
    cnt = c.get("counters")
    assert np.all(cnt[:4] == 0)                                                       # fresh episodes
    assert np.all(np.isfinite(o))
    env.close()

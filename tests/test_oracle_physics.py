"""Closed-form and invariant checks of the oracle's physics step."""
import numpy as np
import pytest

from common import TASK_PUSH, TASK_REACH, panda_task_setup, sample_object_poses, targets_for


def _settled(oracle_lib, task, B, seed=0, double=False):
    m, p = panda_task_setup(task)
    o = oracle_lib.Oracle(m, p, B, double=double, nthreads=4)
    pose = sample_object_poses(B, seed)
    o.reset(pose, targets_for(pose, z=0.65))
    o.step(None, 101, 1, want_obs=False)
    return m, p, o


def test_position_motor_closed_form(oracle_lib):
    """Contact-free: one step moves each controlled joint by kp * 0.05 * a = 0.025 a (App. D.3),
    independent of gravity, up to the solver's velocity residual."""
    m, p, o = _settled(oracle_lib, TASK_PUSH, 8)
    q0 = o.state["q"].copy()
    a = np.random.RandomState(0).uniform(-1, 1, (8, 7)).astype(np.float32)
    o.step(a, 1, 0)
    dq = o.state["q"] - q0
    np.testing.assert_allclose(dq[:, :7], 0.025 * a, atol=3e-6)       # 3e-4 m/s residual * dt
    np.testing.assert_allclose(dq[:, 7:], 0, atol=3e-6)                # fingers hold 0.02
    np.testing.assert_allclose(o.state["qd"][:, :7], 0.025 * a * 240, atol=8e-4)


def test_cube_rests_on_table_and_arm_holds_home(oracle_lib):
    m, p, o = _settled(oracle_lib, TASK_PUSH, 16)
    z = o.state["obj_pose"][:, 2]
    assert np.all(np.abs(z - 0.65) < 2e-4)                             # table top 0.625 + half extent
    assert np.all(np.abs(o.state["obj_vel"]) < 1e-3)
    home = np.array([m.home[i] for i in range(9)], np.float32)
    np.testing.assert_allclose(o.state["q"], np.tile(home, (16, 1)), atol=1e-5)
    # four contact points (the bottom face corners), keys are cube vertices 0..3 vs table
    assert np.all(o.state["status"][:, 2] == 4)
    np.testing.assert_array_equal(np.sort(o.state["cache_key"][:, :4], axis=1), np.tile(np.arange(4), (16, 1)))
    # normal impulses carry the weight: sum(lambda_n) = m g dt
    lam = o.state["cache_lam"].reshape(16, 16, 3)[:, :4, 0].sum(axis=1)
    np.testing.assert_allclose(lam, 0.1 * 9.8 / 240, rtol=2e-2)


def test_free_fall_before_contact(oracle_lib):
    """Dropping cube: semi-implicit Euler with the (small) free-body damping of the statement."""
    m, p = panda_task_setup(TASK_PUSH)
    o = oracle_lib.Oracle(m, p, 1, double=True)
    pose = sample_object_poses(1, 0)
    o.reset(pose, targets_for(pose))
    v, z = 0.0, 0.695
    for _ in range(10):
        o.step(None, 1, 1, want_obs=False)
        v = v + p.dt * (-9.8 - v * (p.damp_lin_k1 + p.damp_lin_k2 * abs(v)))
        z = z + p.dt * v
    assert abs(o.state["obj_vel"][0, 2] - v) < 1e-6
    assert abs(o.state["obj_pose"][0, 2] - z) < 1e-6
    assert o.state["status"][0, 2] == 0                                # no contact yet


def test_fp32_vs_fp64_oracle_trajectories(oracle_lib):
    """Bounds the fp32 rounding error of the statement itself over 240 contact-free steps."""
    B = 8
    _, _, o32 = _settled(oracle_lib, TASK_REACH, B, seed=2)
    _, _, o64 = _settled(oracle_lib, TASK_REACH, B, seed=2, double=True)
    rng = np.random.RandomState(3)
    for _ in range(240):
        a = rng.uniform(-1, 1, (B, 7)).astype(np.float32)
        r32 = o32.step(a, 1, 0)
        r64 = o64.step(a, 1, 0)
    free = (o32.state["status"][:, 2] <= 4) & (o64.state["status"][:, 2] <= 4)
    assert free.any()
    assert np.abs(o32.state["q"] - o64.state["q"])[free].max() < 5e-4
    assert np.abs(r32[1] - r64[1])[free].max() < 5e-3


def test_joint_limits_hold(oracle_lib):
    """Drive joint 4 (range [-3.1416, 0]) into its upper limit: target is clamped (panda_env.py:303)
    and the limit row keeps q <= 0."""
    m, p, o = _settled(oracle_lib, TASK_REACH, 2)
    a = np.zeros((2, 7), np.float32)
    a[:, 3] = 1.0
    for _ in range(200):
        o.step(a, 1, 0)
    assert np.all(o.state["q"][:, 3] <= 1e-4)
    assert np.all(o.state["q"][:, 3] > -0.2)
    assert np.all(o.state["mtarget"][:, 3] <= 0.0)


def test_termination_and_reward_semantics(oracle_lib):
    """pandaPush as registered succeeds at once: target = object + (0.05, 0.05, 0) is 0.0707 m away and
    the success radius is 0.1: done = 1, reward = 1000 + (100 - 80 d2); the step counter
    does not advance on terminating steps (panda_push_gym_env.py:239-242)."""
    m, p, o = _settled(oracle_lib, TASK_PUSH, 4)
    tg = o.state["obj_pose"][:, :3].copy()
    tg[:, :2] += 0.05
    o.state["target"][:] = tg
    obs, rew, done = o.step(np.zeros((4, 7), np.float32), 1, 0)
    assert np.all(done == 1)
    d2 = np.linalg.norm(o.state["obj_pose"][:, :3] - tg, axis=1)
    np.testing.assert_allclose(rew, 1000 + (100 - 80 * d2), rtol=1e-6)
    np.testing.assert_array_equal(o.state["counters"], [[0, 1]] * 4)
    # far target: running reward -d1 - d2, counter advances, done only after max_steps
    o.state["target"][:, 0] += 0.3
    o.state["counters"][:] = 0
    obs, rew, done = o.step(np.zeros((4, 7), np.float32), 1, 0)
    assert np.all(done == 0) and np.all(rew < 0)
    np.testing.assert_array_equal(o.state["counters"], [[1, 0]] * 4)
    raw = o.state["raw_obs"]
    d1 = np.linalg.norm(raw[:, :3] - raw[:, 18:21], axis=1)
    d2 = np.linalg.norm(raw[:, 18:21] - raw[:, 30:33], axis=1)
    np.testing.assert_allclose(rew, -d1 - d2, atol=1e-6)
    # observation scaling = scale_gym_data (utils.py:91)
    lo = np.array(list(p.obs_low)[:33], np.float32)
    hi = np.array(list(p.obs_high)[:33], np.float32)
    np.testing.assert_allclose(obs, 2 * (raw - lo) / (hi - lo) - 1, atol=2e-6)
    o.state["counters"][:, 0] = 1001
    _, _, done = o.step(np.zeros((4, 7), np.float32), 1, 0)
    assert np.all(done == 1)                                            # counter > max_steps (:313)


def _christoffel_bias(o, q, qd, h=2e-3):
    """c_i = sum_jk 1/2 (dM_ij/dq_k + dM_ik/dq_j - dM_jk/dq_i) qd_j qd_k from central differences of the oracle's own
    mass matrix M(q) = (M^-1(q))^-1 — the velocity-dependent forces every rigid-body system must show, computed
    without the recursive algorithm under test."""
    nd = len(q)
    M = lambda qq: np.linalg.inv(o.minv(qq.astype(np.float32)).astype(np.float64))
    dM = np.zeros((nd, nd, nd))
    for k in range(nd):
        e = np.zeros(nd); e[k] = h
        qp, qm = (q + e).astype(np.float32).astype(np.float64), (q - e).astype(np.float32).astype(np.float64)
        dM[:, :, k] = (M(qp) - M(qm)) / (qp[k] - qm[k])
    c = np.zeros(nd)
    for i in range(nd):
        G = 0.5 * (dM[i, :, :] + dM[i, :, :].T - dM[:, :, i])
        c[i] = qd @ G @ qd
    return c, M(q)


@pytest.mark.parametrize("robot", ["panda", "icub"])
def test_bias_forces_are_the_christoffel_symbols_of_the_mass_matrix(oracle_lib, robot):
    """Pins the Coriolis / centrifugal part of the articulated-body algorithm (chain and branching tree) on first principles:
    with tau = 0, M (qdd(q, qd) - qdd(q, 0)) must equal -c(q, qd), c built from derivatives of the mass matrix (which
    tests/test_oracle_kat.py pins on SURVEY App. C.3).  Independent of the recursion: a wrong velocity-product term, a wrong
    frame for the spatial cross products or a missing branch contribution shows here."""
    from pybullet_robot_envs.b2env.model import default_params, load_icub, TASK_REACH
    if robot == "panda":
        m, p = panda_task_setup(TASK_PUSH)
    else:
        m, _ = load_icub()
        p = default_params(TASK_REACH, [0] * 31, [1] * 31)
    nd = m.n_dof
    o = oracle_lib.Oracle(m, p, 1, double=True)
    rng = np.random.RandomState(3)
    lo = np.array([m.lower[d] for d in range(nd)], np.float64)
    hi = np.array([m.upper[d] for d in range(nd)], np.float64)
    z = np.zeros(nd, np.float32)
    worst = 0.0
    for trial in range(2 if robot == "icub" else 4):
        q = (lo + (hi - lo) * rng.uniform(0.2, 0.8, nd)).astype(np.float32).astype(np.float64)
        qd = rng.uniform(-1.5, 1.5, nd)
        c, M = _christoffel_bias(o, q, qd)
        a_v = o.forward_dynamics(q.astype(np.float32), qd.astype(np.float32), z).astype(np.float64)
        a_0 = o.forward_dynamics(q.astype(np.float32), z, z).astype(np.float64)
        lhs = M @ (a_v - a_0)
        scale = max(1.0, np.abs(c).max())
        assert np.abs(c).max() > 0.05                       # the configuration does exercise the velocity products
        np.testing.assert_allclose(lhs, -c, atol=4e-3 * scale)
        worst = max(worst, np.abs(lhs + c).max() / scale)
    assert worst < 4e-3

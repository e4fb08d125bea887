"""Oracle checks for the Cartesian (IK) control mode and the GoalEnv rules (CPU only)."""
import ctypes as C
import math

import numpy as np

from common import TASK_PUSH, panda_task_setup, sample_object_poses, targets_for


def _e2q(r, p, y):
    cr, sr, cp, sp, cy, sy = (f(x / 2) for x in (r, p, y) for f in (math.cos, math.sin))
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy], np.float32)


def _ik(o, m, p, q0, tpos, tquat):
    out = np.zeros(9, np.float32)
    q0 = np.ascontiguousarray(q0, np.float32)
    tpos = np.ascontiguousarray(tpos, np.float32)
    tquat = np.ascontiguousarray(tquat, np.float32)
    o.lib.b2o_ik.restype = C.c_int
    it = o.lib.b2o_ik(C.byref(m), C.byref(p), q0.ctypes.data_as(C.c_void_p), tpos.ctypes.data_as(C.c_void_p),
                      tquat.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return it, out


def test_dls_ik_reaches_the_home_hand_pose(oracle_lib):
    """calculateInverseKinematics(maxNumIterations=100, residualThreshold=1e-3) from the home joint state
    to the home hand pose (0.2, 0, 0.8, euler (pi,0,0)) of panda_env.py:85-88."""
    m, p = panda_task_setup(TASK_PUSH, use_ik=1)
    o = oracle_lib.Oracle(m, p, 1)
    home = np.array([m.home[i] for i in range(9)], np.float32)
    it, q = _ik(o, m, p, home, (0.2, 0.0, 0.8), _e2q(math.pi, 0, 0))
    assert 0 < it < 100
    pos, rot = o.fk(q)
    assert np.linalg.norm(pos[11] - (0.2, 0.0, 0.8)) <= 1e-3 * 1.01          # residual threshold
    J, _, quat = o.ee_jacobian(q)
    assert abs(abs(np.dot(quat, _e2q(math.pi, 0, 0))) - 1) < 1e-4            # orientation reached
    np.testing.assert_array_equal(q[7:], home[7:])                            # finger dofs are not on the EE path
    # already at the target: zero iterations, joints untouched
    it2, q2 = _ik(o, m, p, q, pos[11], quat)
    assert it2 == 0
    np.testing.assert_array_equal(q2, q)


def test_cartesian_mode_rollout_tracks_the_commanded_pose(oracle_lib):
    B = 4
    m, p = panda_task_setup(TASK_PUSH, use_ik=1)
    assert p.n_act == 6
    o = oracle_lib.Oracle(m, p, B, nthreads=4)
    pose = sample_object_poses(B, 1)
    o.reset(pose, targets_for(pose, z=0.65) + np.array([0.3, 0, 0], np.float32))
    o.step(None, 1, 3, want_obs=False)            # robot.reset(): IK of the home hand pose + one step
    o.step(None, 200, 1, want_obs=False)          # settle with those targets
    raw0 = None
    a = np.zeros((B, 6), np.float32)
    a[:, 0] = 1.0                                  # +x at 0.005 per step (panda_push_gym_env.py:203)
    for i in range(60):
        obs, rew, done = o.step(a, 1, 0)
    hp = o.state["hand_pose"]
    # commanded x: 0.2 (home) -> first command clamps into the workspace [0.3, 0.65] -> +0.005 per step
    np.testing.assert_allclose(hp[:, 0], 0.3 + 59 * 0.005, atol=1e-5)
    for i in range(40):
        obs, rew, done = o.step(a, 1, 0)
    hp = o.state["hand_pose"]
    np.testing.assert_allclose(hp[:, 0], 0.65, atol=1e-6)       # workspace clamp (panda_push_gym_env.py:215-219)
    np.testing.assert_allclose(hp[:, 3:], np.tile([math.pi, 0, 0], (B, 1)), atol=1e-6)
    ee = o.state["raw_obs"][:, :3]
    assert np.all(np.abs(ee[:, 0] - hp[:, 0]) < 0.06)           # the hand follows (kp 0.2 lag)
    assert np.all(np.abs(ee[:, 2] - hp[:, 2]) < 0.06)
    assert (o.state["status"][:, 0] & 1).sum() == 0


def test_goal_env_rules(oracle_lib):
    """panda_push_gym_goal_env.py:96-122: reward -(d > 0.1), done = counter > max_steps or success,
    the counter keeps advancing after success."""
    B = 4
    m, p = panda_task_setup(TASK_PUSH, goal_env=1, max_steps=5)
    o = oracle_lib.Oracle(m, p, B)
    pose = sample_object_poses(B, 2)
    tg = targets_for(pose, z=0.65)
    tg[2:, 0] += 0.3                                # envs 2,3: far target
    o.reset(pose, tg)
    o.step(None, 101, 1, want_obs=False)
    z = np.zeros((B, 7), np.float32)
    obs, rew, done = o.step(z, 1, 0)
    np.testing.assert_array_equal(rew, [0, 0, -1, -1])
    np.testing.assert_array_equal(done, [1, 1, 0, 0])
    np.testing.assert_array_equal(o.state["counters"], [[1, 0]] * 4)
    for _ in range(5):
        obs, rew, done = o.step(z, 1, 0)
    np.testing.assert_array_equal(o.state["counters"][:, 0], [6] * 4)
    np.testing.assert_array_equal(done, [1, 1, 1, 1])            # 6 > max_steps


def test_grasp_task_rules(oracle_lib):
    """PandaGrasp (new task): gripper command drives both finger targets with the grasp gains/force;
    lifting the object 0.1 m above its rest height succeeds."""
    from common import TASK_GRASP
    B = 4
    m, p = panda_task_setup(TASK_GRASP)
    assert p.n_act == 8 and m.max_force[7] == 10.0 and m.max_vel[8] == 1.0
    o = oracle_lib.Oracle(m, p, B)
    pose = sample_object_poses(B, 3)
    o.reset(pose, targets_for(pose, z=0.75))
    o.step(None, 101, 1, want_obs=False)
    a = np.zeros((B, 8), np.float32)
    a[:, 7] = 1.0                                    # open
    for _ in range(80):
        obs, rew, done = o.step(a, 1, 0)
    np.testing.assert_allclose(o.state["mtarget"][:, 7:], 0.04, atol=1e-7)
    np.testing.assert_allclose(o.state["q"][:, 7:], 0.04, atol=2e-3)
    assert np.all(done == 0) and np.all(rew < 0)
    a[:, 7] = -1.0                                   # close: maxVelocity 1 m/s caps the finger speed
    o.step(a, 1, 0)
    assert np.all(np.abs(o.state["qd"][:, 7:]) <= 1.0 + 1e-3)
    # teleport the object upwards: success
    o.state["obj_pose"][:, 2] = 0.65 + 0.11
    obs, rew, done = o.step(a, 1, 0)
    assert np.all(done == 1) and np.all(rew == 1000)


def test_cartesian_velocity_cap(oracle_lib):
    """robot.apply_action(action, max_vel) with max_vel != -1 (panda_env.py:285-291): the 7 arm motors run with PyBullet's default
    position gain and maxVelocity = max_vel, so a far Cartesian command moves no arm joint faster than the cap; without the
    cap the same command does."""
    B = 2
    peak = {}
    for cap in (-1.0, 0.3):
        m, p = panda_task_setup(TASK_PUSH, use_ik=1)
        p.ik_max_vel = cap
        o = oracle_lib.Oracle(m, p, B, nthreads=2)
        pose = sample_object_poses(B, 1)
        o.reset(pose, targets_for(pose, z=0.65))
        o.step(None, 1, 3, want_obs=False)
        o.step(None, 100, 3, want_obs=False)                       # hold the home hand pose through the IK mode
        o.state["hand_pose"][:, :3] = [0.55, 0.25, 0.75]           # a far command
        worst = 0.0
        for i in range(30):
            o.step(None, 1, 3, want_obs=False)
            worst = max(worst, float(np.abs(o.state["qd"][:, :7]).max()))
        peak[cap] = worst
    assert peak[0.3] <= 0.3 * 1.02, peak
    assert peak[-1.0] > 0.6, peak


def test_quaternion_helpers_round_trip():
    from pybullet_robot_envs.envs.utils import euler_from_quaternion, quaternion_from_euler
    rng = np.random.RandomState(0)
    e = rng.uniform(-1.5, 1.5, (200, 3))
    q = quaternion_from_euler(e)
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(euler_from_quaternion(q), e, atol=1e-9)
    np.testing.assert_allclose(quaternion_from_euler([math.pi, 0, 0]), [1, 0, 0, 0], atol=1e-12)   # the home hand orientation
    for k in range(5):   # same numbers as the oracle's conversions
        np.testing.assert_allclose(quaternion_from_euler(e[k]), _e2q(*e[k]), atol=1e-6)

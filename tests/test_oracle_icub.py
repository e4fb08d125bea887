"""iCub restatement (oracle) pinned on what can be pinned without PyBullet: SURVEY App. B.2 / C.5 known answers
through the 32-body merged model, closed forms (position-motor response, App. D.3), and the Python-level semantics
of the iCub task envs that ARE in the reference (icub_env.py, icub_push_gym_env.py, icub_reach_gym_env.py)."""
import numpy as np

from pybullet_robot_envs.b2env.model import (REWARD_ICUB_PUSH0, REWARD_ICUB_PUSH1, REWARD_ICUB_REACH, TASK_PUSH, TASK_REACH,
                                             icub_task_setup, load_icub_arm)

PARK = np.array([[5, 5, 0.025, 0, 0, 0, 1]], np.float32)


def test_merged_model_equals_38_link_model(oracle_lib):
    """Folding the six welded F/T-sensor links into their parents changes neither kinematics nor dynamics."""
    m38, info = load_icub_arm('l', merged=False)
    m32, info2 = load_icub_arm('l', merged=True)
    assert (m38.n_links, m38.n_dof, m32.n_links, m32.n_dof) == (38, 32, 32, 32)
    assert m32.ee_link == 21 and info2["movable"][21] == "l_wrist_yaw"
    assert abs(sum(m32.mass[i] for i in range(32)) - sum(m38.mass[i] for i in range(38))) < 1e-5
    _, p = icub_task_setup(TASK_PUSH)
    o38, o32 = oracle_lib.Oracle(m38, p, 1, double=True), oracle_lib.Oracle(m32, p, 1, double=True)
    rng = np.random.RandomState(0)
    for _ in range(4):
        q = np.array([m38.lower[i] + rng.rand() * (m38.upper[i] - m38.lower[i]) for i in range(32)], np.float32)
        qd, tau = rng.randn(32).astype(np.float32), rng.randn(32).astype(np.float32)
        a, b = o38.forward_dynamics(q, qd, tau), o32.forward_dynamics(q, qd, tau)
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-4)
        p38, _ = o38.fk(q)
        p32, _ = o32.fk(q)
        np.testing.assert_allclose(p38[26], p32[21], atol=1e-6)     # l_hand: PyBullet joint 26 = body 21


def test_pinned_base_and_home_hand(oracle_lib):
    """SURVEY App. C.5 (+ the pin: z x 1.2, App. E.14): l_hand link origin z = 0.737243 + 0.126; x, y move by the
    0.0064 rad yaw flip of the pin (3.14 -> -3.14)."""
    m, info = load_icub_arm('l', merged=True)
    assert abs(m.base_pos[2] - 0.756) < 1e-6
    _, p = icub_task_setup(TASK_PUSH)
    o = oracle_lib.Oracle(m, p, 1, double=True)
    q = np.array([m.home[i] for i in range(32)], np.float32)
    pos, rot = o.fk(q)
    assert abs(pos[21][2] - 0.863243) < 2e-5
    np.testing.assert_allclose(pos[21][:2], [0.259158, 0.192303], atol=2.5e-3)


def test_joint_mode_motor_closed_form(oracle_lib):
    """App. D.3: from rest, target q + 0.05 a with kp 0.5 moves a controlled joint by 0.025 a in one step
    (icub_push_gym_env.py:256-257, icub_env.py:347-361); blocked joints hold their rest pose."""
    m, p = icub_task_setup(TASK_PUSH, use_ik=0)
    assert p.n_act == 10 and [p.ctrl_dof[k] for k in range(10)] == [12, 13, 14, 15, 16, 17, 18, 19, 20, 21]
    o = oracle_lib.Oracle(m, p, 1)
    o.reset(PARK, np.zeros((1, 3), np.float32))
    o.step(None, 30, 1, want_obs=False)
    q0 = o.state["q"].copy()
    a = np.array([[0.5, -0.5, 1.0, 0.3, -0.2, 0.7, -1.0, 0.4, -0.6, 0.1]], np.float32)
    o.step(a, 1, 0)
    dq = (o.state["q"] - q0)[0]
    np.testing.assert_allclose(dq[12:22], 0.025 * a[0], atol=2e-5)
    assert np.abs(np.delete(dq, np.arange(12, 22))).max() < 1e-5


def test_ik_reaches_home_hand_pose_and_blocks_joints(oracle_lib):
    """robot.reset() in IK mode (icub_env.py:148-151): the IK of the home hand pose; blocked joints keep the rest
    pose (:314-317); after the settle the hand COM sits at (0.3, 0.26, 0.8) within the 1e-3 IK residual."""
    m, p = icub_task_setup(TASK_PUSH, use_ik=1)
    assert p.n_act == 3 and p.n_obs == 34
    o = oracle_lib.Oracle(m, p, 1)
    o.reset(PARK, np.zeros((1, 3), np.float32))
    o.step(None, 1, 3, want_obs=False)
    mt = o.state["mtarget"][0]
    home = np.array([m.home[i] for i in range(32)], np.float32)
    blocked = np.ones(32, bool)
    blocked[12:22] = False
    np.testing.assert_array_equal(mt[blocked], home[blocked])
    assert np.abs(mt[~blocked] - home[~blocked]).max() > 1e-2
    o.step(None, 200, 1, want_obs=False)
    o.step(None, 0, 4)
    np.testing.assert_allclose(o.state["raw_obs"][0, :3], [0.3, 0.26, 0.8], atol=2e-3)
    assert np.abs(o.state["raw_obs"][0, 6:9]).max() < 1e-3      # raw (unstandardised) hand velocity, at rest


def test_cartesian_action_increments_and_clamps(oracle_lib):
    """hand_pose += 0.005 a, clamped to the workspace whose floor is the table height (icub_push_gym_env.py:79-81,
    :229-250)."""
    m, p = icub_task_setup(TASK_PUSH, use_ik=1)
    o = oracle_lib.Oracle(m, p, 1)
    o.reset(PARK, np.zeros((1, 3), np.float32))
    o.step(np.array([[1, -1, 0.5]], np.float32), 1, 0)
    np.testing.assert_allclose(o.state["hand_pose"][0, :3], [0.305, 0.255, 0.8025], atol=1e-6)
    o.state["hand_pose"][0, :3] = [0.449, 0.299, 0.626]
    o.step(np.array([[1, 1, -1]], np.float32), 1, 0)
    np.testing.assert_allclose(o.state["hand_pose"][0, :3], [0.45, 0.3, 0.625], atol=1e-6)


def _reward_case(oracle_lib, task, reward_type, cube, target, shaping=(0.4, 0.2)):
    m, p = icub_task_setup(task, use_ik=1, reward_type=reward_type)
    o = oracle_lib.Oracle(m, p, 1)
    pose = np.array([[cube[0], cube[1], cube[2], 0, 0, 0, 1]], np.float32)
    o.reset(pose, np.array([target], np.float32))
    o.state["shaping"][:] = shaping
    obs, rew, done = o.step(None, 0, 4)
    hand = o.state["raw_obs"][0, :3]
    return p, hand, float(rew[0]), float(done[0])


def test_reward_semantics(oracle_lib):
    cube, far, near = (0.25, 0.0, 0.65), (0.35, 0.1, 0.65), (0.26, 0.01, 0.65)
    # push type 0: -d1 - d2, +1000 when d2 <= 0.03 (icub_push_gym_env.py:353-356)
    p, hand, r, d = _reward_case(oracle_lib, TASK_PUSH, 0, cube, far)
    assert p.reward_kind == REWARD_ICUB_PUSH0
    d1, d2 = np.linalg.norm(hand - cube), np.linalg.norm(np.subtract(cube, far))
    assert abs(r - (-d1 - d2)) < 1e-5 and d == 0.0
    p, hand, r, d = _reward_case(oracle_lib, TASK_PUSH, 0, cube, near)
    d2 = np.linalg.norm(np.subtract(cube, near))
    assert abs(r - (-d1 - d2 + 1000.0)) < 1e-3 and d == 1.0
    # push type 1: 0.125 (1 - d1/d0) [+ 0.25 (1 - d2/dmax) once the hand is within 0.1] (:358-371)
    p, hand, r, d = _reward_case(oracle_lib, TASK_PUSH, 1, cube, far, shaping=(0.4, 0.2))
    assert p.reward_kind == REWARD_ICUB_PUSH1
    d2 = np.linalg.norm(np.subtract(cube, far))
    assert d1 > 0.1 and abs(r - 0.125 * (1 - d1 / 0.4)) < 1e-5
    # reach: -d, bonus ADDED (icub_reach_gym_env.py:326-328), success radius 0.03
    p, hand, r, d = _reward_case(oracle_lib, TASK_REACH, 0, cube, (0, 0, 0))
    assert p.reward_kind == REWARD_ICUB_REACH and p.n_obs == 31
    assert abs(r + d1) < 1e-5 and d == 0.0
    m, p = icub_task_setup(TASK_REACH, use_ik=1)
    o = oracle_lib.Oracle(m, p, 1)
    o.reset(PARK, np.zeros((1, 3), np.float32))
    o.step(None, 0, 4)
    hand = o.state["raw_obs"][0, :3].copy()
    o.state["obj_pose"][0, :3] = hand + np.array([0.01, 0, 0], np.float32)
    obs, rew, done = o.step(None, 0, 4)
    assert abs(float(rew[0]) - (-0.01 + 1000.0 + (100 - 0.8))) < 1e-2 and done[0] == 1.0

"""Pins the CPU oracle on the first-principles known-answer vectors of SURVEY.md Appendix C
(FK, Jacobian, mass matrix, gravity vector, free acceleration, seeded object poses).  These are
NOT PyBullet outputs (pybullet is absent; parity with PyBullet itself is unpinned) — they were
derived independently from the URDF and pin the articulated-body restatement."""
import numpy as np
import pytest

from pybullet_robot_envs.b2env.model import TASK_PUSH, load_panda, panda_task_setup, parse_urdf, PANDA_JSON
from pybullet_robot_envs.gym_compat import seeding


@pytest.fixture(scope="module")
def orc(oracle_lib):
    m, p = panda_task_setup(TASK_PUSH)
    return oracle_lib.Oracle(m, p, 1), oracle_lib.Oracle(m, p, 1, double=True), m, p


def home_q(m):
    return np.array([m.home[i] for i in range(9)], np.float32)


def test_model_descriptor():
    m, d = load_panda()
    assert m.n_links == 12 and m.n_dof == 9 and m.ee_link == 11
    assert [m.parent[i] for i in range(12)] == [-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 8, 8]
    assert [m.jtype[i] for i in range(12)] == [0, 0, 0, 0, 0, 0, 0, 4, 4, 1, 1, 4]   # PyBullet joint types
    total = sum(m.mass[i] for i in range(12)) + d["base"]["mass"]
    assert abs(total - 17.961) < 1e-5                                                  # App. B.1
    lim = [(m.lower[i], m.upper[i]) for i in range(9)]
    np.testing.assert_allclose(lim[3], (-3.1416, 0.0), atol=1e-6)
    np.testing.assert_allclose(lim[7], (0.0, 0.04), atol=1e-7)


def test_fk_home_and_zero(orc):
    o32, o64, m, p = orc
    for o, tol in ((o32, 2e-6), (o64, 1e-6)):
        pos, rot = o.fk(home_q(m))
        np.testing.assert_allclose(pos[11], (0.365860, -0.036740, 0.986768), atol=tol)     # C.1 EE
        np.testing.assert_allclose(pos[8], (0.368721, -0.017930, 1.054133), atol=tol)      # hand
        np.testing.assert_allclose(pos[6], (0.373095, 0.010822, 1.157104), atol=tol)       # link7
        np.testing.assert_allclose(pos[9], (0.364833, -0.052816, 1.003354), atol=tol)      # left finger
        np.testing.assert_allclose(pos[10], (0.367834, -0.014431, 0.992509), atol=tol)     # right finger
        pos0, _ = o.fk(np.zeros(9, np.float32))
        np.testing.assert_allclose(pos0[11], (0.088, 0.0, 1.481), atol=tol)
        np.testing.assert_allclose(pos0[8], (0.088, 0.0, 1.551), atol=tol)


def test_ee_quaternion_and_jacobian(orc):
    o32, o64, m, p = orc
    J, pos, quat = o64.ee_jacobian(home_q(m))
    np.testing.assert_allclose(quat, (0.989736, -0.039990, -0.015142, 0.136363), atol=2e-6)
    lin = np.array([[0.03674, 0.028768, 0.031513, 0.284684, -0.017265, 0.178733, 0],
                    [0.36586, 0, 0.328592, 0, 0.118772, 0.045414, 0],
                    [0, -0.36586, 0.01889, 0.457566, -0.032431, 0.071169, 0]])
    ang = np.array([[0, 0, -0.514136, 0, 0.882707, 0.138872, -0.040879],
                    [0, 1, 0, -1, 0, -0.955336, -0.268716],
                    [1, 0, 0.857709, 0, -0.469923, 0.260858, -0.962352]])
    np.testing.assert_allclose(J[:3, :7], lin, atol=2e-6)   # C.3
    np.testing.assert_allclose(J[3:, :7], ang, atol=2e-6)
    np.testing.assert_allclose(J[:, 7:], 0, atol=1e-9)      # finger columns: EE hangs off the hand


def test_free_acceleration_gravity_and_mass_matrix(orc):
    o32, o64, m, p = orc
    q = home_q(m)
    z = np.zeros(9, np.float32)
    qdd_ref = (-1.215651, -3.186940, -0.197523, -23.636983, -2.775132, 20.344422, 5.321340, -0.612245, 0.612245)
    np.testing.assert_allclose(o64.forward_dynamics(q, z, z), qdd_ref, atol=5e-6)   # C.2
    np.testing.assert_allclose(o32.forward_dynamics(q, z, z), qdd_ref, atol=2e-4)
    M = np.linalg.inv(o64.minv(q).astype(np.float64))
    np.testing.assert_allclose(np.diag(M), (1.653442, 2.117163, 1.609751, 1.702818, 0.816721, 0.737155, 0.600180,
                                            0.1, 0.1), atol=3e-6)                   # C.3
    np.testing.assert_allclose(M[0], (1.653442, -0.025176, 1.310472, 0.000305, -0.339717, 0.210539, -0.577584,
                                      -0.035368, 0.035368), atol=3e-6)
    np.testing.assert_allclose(M, M.T, atol=1e-5)
    assert abs(np.linalg.eigvalsh(M).min() - 0.094927) < 2e-6
    G = -M @ o64.forward_dynamics(q, z, z).astype(np.float64)
    np.testing.assert_allclose(G, (0, -9.48347, -0.335738, 19.418489, 0.759811, 1.366487, 0, 0.265715, -0.265715),
                               atol=2e-5)                                            # C.2 gravity vector


def test_seeding_and_object_pose_vectors():
    # C.4: gym-0.12.5 np_random seed words and the first (x, y, yaw) draws of WorldEnv._sample_pose
    assert seeding.seed_words(0) == [547404849, 309914516]
    assert seeding.seed_words(1) == [2739863373, 598274112]
    assert seeding.seed_words(42) == [3917269561, 1772078828]
    want = {0: (0.4054360056, 0.0465390937, 0.2084304499), 1: (0.4807390436, 0.0014500072, -0.4851904736),
            42: (0.4374143378, -0.0015661448, 0.6609452598)}
    for s, (x, y, yaw) in want.items():
        r, _ = seeding.np_random(s)
        got = (0.45 + r.uniform(-0.05, 0.05), r.uniform(-0.05, 0.05), r.uniform(-np.pi / 4, np.pi / 4))
        np.testing.assert_allclose(got, (x, y, yaw), atol=1e-9)


def test_observation_space_vectors():
    # C.6: pandaPush low | high
    m, p = panda_task_setup(TASK_PUSH)
    pi = np.pi
    low = [0.3, -0.3, 0.425, -pi, -pi, -pi, -1, -1, -1, -2.9671, -1.8326, -2.9671, -3.1416, -2.9671, -0.0873, -2.9671,
           0, 0, 0.3, -0.3, 0.625, -pi, -pi, -pi, -0.5, -0.5, -0.5, 0, 0, 0, 0.3, -0.3, 0.625]
    high = [0.65, 0.3, 1.5, pi, pi, pi, 1, 1, 1, 2.9671, 1.8326, 2.9671, 0, 2.9671, 3.8223, 2.9671, 0.04, 0.04,
            0.65, 0.3, 0.925, pi, pi, pi, 0.5, 0.5, 0.5, 2 * pi, 2 * pi, 2 * pi, 0.65, 0.3, 0.925]
    assert p.n_obs == 33
    np.testing.assert_array_equal(np.array(list(p.obs_low)[:33], np.float32), np.array(low, np.float32))
    np.testing.assert_array_equal(np.array(list(p.obs_high)[:33], np.float32), np.array(high, np.float32))


def test_committed_model_json_matches_reference_urdf():
    """tests/golden-style pin: the committed descriptor equals a fresh parse of the reference URDF
    (only runs where /root/reference exists, i.e. in the build container)."""
    import json
    import os
    src = "/root/reference/pybullet_robot_envs/robot_data/franka_panda/panda_model.urdf"
    if not os.path.exists(src):
        pytest.skip("reference tree not present (GPU box)")
    fresh = parse_urdf(src)
    with open(PANDA_JSON) as f:
        committed = json.load(f)
    assert json.loads(json.dumps(fresh["joints"])) == committed["joints"]
    assert json.loads(json.dumps(fresh["links"])) == committed["links"]


def test_icub_model_and_tree_kinematics(oracle_lib):
    """iCub groundwork (SDF loader + the generic tree code of the oracle), pinned on SURVEY App. B.2 / C.5:
    39 links / 38 joints / 32 dofs, total mass 33.06 kg, joint indices in PyBullet order, hand positions."""
    from pybullet_robot_envs.b2env.model import default_params, load_icub, TASK_REACH
    m, d = load_icub()
    assert m.n_links == 38 and m.n_dof == 32
    assert abs(sum(m.mass[i] for i in range(38)) + d["base"]["mass"] - 33.06) < 5e-3
    names = [j["name"] for j in d["joints"]]
    assert [d["joints"][i]["type"] for i in (2, 7, 10, 15, 22, 33)] == ["fixed"] * 6          # FT sensor joints
    assert names[16:19] == ["torso_pitch", "torso_roll", "torso_yaw"]
    assert names[26] == "l_wrist_yaw" and names[37] == "r_wrist_yaw" and m.ee_link == 26        # icub_env.py:121-143
    ctrl_l = [16, 17, 18, 19, 20, 21, 23, 24, 25, 26]
    assert all(d["joints"][i]["type"] == "revolute" for i in ctrl_l)
    p = default_params(TASK_REACH, [0] * 31, [1] * 31)
    o = oracle_lib.Oracle(m, p, 1, double=True)
    home = np.array([m.home[i] for i in range(32)], np.float32)
    pos, rot = o.fk(home)
    com = lambda i: pos[i] + rot[i] @ np.array([m.com[i][k] for k in range(3)])
    np.testing.assert_allclose(pos[26], (0.259158, 0.192303, 0.737243), atol=2e-6)              # l_hand link origin
    np.testing.assert_allclose(com(26), (0.319191, 0.206436, 0.767843), atol=2e-6)              # l_hand COM
    np.testing.assert_allclose(pos[37], (0.258560, -0.224899, 0.736992), atol=2e-6)             # r_hand
    np.testing.assert_allclose(com(37), (0.318462, -0.239160, 0.767587), atol=2e-6)
    pos0, rot0 = o.fk(np.zeros(32, np.float32))
    np.testing.assert_allclose(pos0[26] + rot0[26] @ np.array([m.com[26][k] for k in range(3)]),
                               (0.028214, 0.079672, 0.435014), atol=2e-6)
    # articulated-body algorithm on a branching tree: M^-1 symmetric positive definite, gravity torques finite
    Minv = o.minv(home).astype(np.float64)
    np.testing.assert_allclose(Minv, Minv.T, atol=1e-4 * np.abs(Minv).max())
    assert np.linalg.eigvalsh(0.5 * (Minv + Minv.T)).min() > 0
    qdd = o.forward_dynamics(home, np.zeros(32, np.float32), np.zeros(32, np.float32))
    assert np.isfinite(qdd).all() and np.abs(qdd).max() > 1.0

"""iCub parity cases shared by the CPU-emulation tests (tests/test_emu_kernels.py, ``-m "not gpu"``) and the GPU
tests (tests/test_gpu_icub.py, ``-m gpu``): the same kernel source, run by two back ends, checked against the
oracle on the same seeded inputs.  ``make_sim(model, params, B)`` builds the simulation under test.

Tolerances (fp32 both sides; oracle = ABA + delta-velocity PGS on the 32-body model, kernel = world-frame CRBA +
Gauss-Jordan + Delassus-form PGS): single step from identical state q 2e-5, qd 5e-3 (the solver exits on a
velocity residual of sqrt(1e-7) = 3e-4), object pose 2e-5; counts / keys / counters exact."""
import numpy as np

from pybullet_robot_envs.b2env.model import TASK_PUSH, TASK_REACH, icub_task_setup

SYNC_FIELDS = ("q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "cache_key", "cache_lam", "hand_pose",
               "shaping")


def object_poses(B, seed=0):
    rng = np.random.RandomState(seed)
    pose = np.zeros((B, 7), np.float32)
    pose[:, 0] = 0.25 + rng.uniform(-0.05, 0.05, B)
    pose[:, 1] = rng.uniform(-0.05, 0.05, B)
    pose[:, 2] = 0.695
    yaw = rng.uniform(-np.pi / 4, np.pi / 4, B)
    pose[:, 5], pose[:, 6] = np.sin(yaw / 2), np.cos(yaw / 2)
    return pose


def sync(orc, sim):
    for f in SYNC_FIELDS:
        sim.set(f, orc.state[f])


def check_state(orc, sim, tag, tol_q=2e-5, tol_qd=5e-3, tol_obj=2e-5, tol_vel=5e-3, exact_rows=True, ik=False):
    """Returns the number of environments excluded because the IK loop stopped one iteration apart on the two sides
    (it exits on a 1e-3 m residual: one fp32 ulp can flip the last iteration, which moves the joint targets by up to
    ~1e-3 rad); everything else is compared on the remaining environments."""
    dt = np.abs(sim.get("mtarget") - orc.state["mtarget"]).max(axis=1)
    ok = dt <= (2e-4 if ik else 2e-5)
    assert ok.any(), (tag, dt.max())
    for f, tol in (("q", tol_q), ("qd", tol_qd), ("obj_pose", tol_obj), ("obj_vel", tol_vel)):
        err = np.abs(sim.get(f) - orc.state[f])[ok].max()
        assert err <= tol, (tag, f, err)
    np.testing.assert_array_equal(sim.get("counters"), orc.state["counters"], err_msg=tag)
    if exact_rows:
        g, o = sim.get("status"), orc.state["status"]
        np.testing.assert_array_equal(g[ok, 2:], o[ok, 2:], err_msg="n_contacts / n_rows " + tag)
        np.testing.assert_array_equal(np.sort(sim.get("cache_key"), axis=1)[ok], np.sort(orc.state["cache_key"], axis=1)[ok],
                                      err_msg="contact keys " + tag)
    return int((~ok).sum()), ok


def single_step_parity(make_sim, oracle_lib, B, use_ik, control_orientation=0, arm='l', task=TASK_PUSH, n_hold=3, n_act=6,
                       reward_type=0, seed=0, goal_env=0):
    """Every step restarts the kernel from the oracle's state (isolates the per-step error)."""
    m, p = icub_task_setup(task, control_arm=arm, use_ik=use_ik, control_orientation=control_orientation,
                           reward_type=reward_type, goal_env=goal_env)
    sim = make_sim(m, p, B)
    orc = oracle_lib.Oracle(m, p, B, nthreads=4)
    pose = object_poses(B, seed)
    tg = (pose[:, :3] + np.array([0.05, 0.05, -0.045], np.float32)).astype(np.float32)
    orc.reset(pose, tg)
    orc.state["shaping"][:] = np.array([0.3, 0.07], np.float32)
    rng = np.random.RandomState(seed + 1)
    flips = 0
    if use_ik:   # robot.reset(): IK of the home hand pose, one step (icub_env.py:148-151)
        sync(orc, sim)
        orc.step(None, 1, 3, want_obs=False)
        sim.step_host(None, 1, 3, want_obs=False)
        flips += check_state(orc, sim, "ik-pose", ik=True)[0]
    for i in range(n_hold):
        sync(orc, sim)
        orc.step(None, 20, 1, want_obs=False)
        sim.step_host(None, 20, 1, want_obs=False)
        check_state(orc, sim, "hold %d" % i, tol_q=5e-5)
    for i in range(n_act):
        sync(orc, sim)
        a = rng.uniform(-1, 1, (B, p.n_act)).astype(np.float32)
        o_obs, o_rew, o_done = orc.step(a, 1, 0)
        g_obs, g_rew, g_done = sim.step_host(a, 1, 0)
        nf, ok = check_state(orc, sim, "act %d" % i, ik=bool(use_ik))
        flips += nf
        np.testing.assert_allclose(sim.get("raw_obs")[ok], orc.state["raw_obs"][ok], atol=2e-3, rtol=0, err_msg="raw obs %d" % i)
        np.testing.assert_allclose(g_obs[ok], o_obs[ok], atol=2e-2, rtol=0, err_msg="obs %d" % i)
        np.testing.assert_allclose(g_rew[ok], o_rew[ok], atol=1e-3, rtol=1e-5, err_msg="reward %d" % i)
        np.testing.assert_array_equal(g_done, o_done)
        np.testing.assert_allclose(sim.get("hand_pose"), orc.state["hand_pose"], atol=1e-6)
    assert flips <= max(1, 0.02 * (n_act + 1) * B), flips   # an extra / missing last IK iteration must be rare
    return orc, sim, m, p


def hand_contact_parity(make_sim, oracle_lib, n_check=6):
    """The hand is driven down onto the cube / next to it (MODE_IK_POSE with a low hand pose): proxy-cube contacts
    couple the arm and cube islands.  Oracle runs the approach; the kernel is checked step by step from the oracle's
    state once contacts exist."""
    B = 4
    m, p = icub_task_setup(TASK_PUSH, use_ik=1)
    sim = make_sim(m, p, B)
    orc = oracle_lib.Oracle(m, p, B, nthreads=4)
    pose = np.zeros((B, 7), np.float32)
    pose[:, 0] = [0.30, 0.33, 0.36, 0.30]
    pose[:, 1] = [0.26, 0.26, 0.27, 0.20]
    pose[:, 2] = 0.651
    pose[:, 6] = 1
    tg = (pose[:, :3] + np.array([0.05, 0.05, 0], np.float32)).astype(np.float32)
    orc.reset(pose, tg)
    orc.state["shaping"][:] = 1
    orc.step(None, 1, 3, want_obs=False)
    orc.step(None, 60, 1, want_obs=False)
    orc.state["hand_pose"][:, 2] = 0.70
    orc.step(None, 14, 3, want_obs=False)
    seen = False
    for i in range(n_check):
        sync(orc, sim)
        orc.step(None, 1, 3, want_obs=False)
        sim.step_host(None, 1, 3, want_obs=False)
        seen = seen or bool((orc.state["cache_key"] >= 16).any())
        # a jammed hand-cube-table system does not meet the residual: compare with solver-level tolerances
        check_state(orc, sim, "contact %d" % i, tol_q=1e-4, tol_qd=3e-2, tol_obj=1e-4, tol_vel=3e-2, ik=True)
    assert seen, "the approach produced no hand-cube contact"
    return orc, sim


def static_world_parity(make_sim, oracle_lib, rounds=5):
    """The static world beyond the table top (the tree kernel's general collision path): cubes at the rim of the top slab, tipping
    over it, on the floor next to a table leg, and hands driven to the table edge so that sphere proxies touch the rim from the
    side — next to ordinary environments in the same launch.  The oracle runs the scenario; every few steps the kernel restarts
    from the oracle's state and one step is compared: contact counts and keys exact, states to solver tolerances (most of these
    systems are sweep-capped)."""
    B = 10
    m, p = icub_task_setup(TASK_PUSH, use_ik=1)
    sim = make_sim(m, p, B)
    orc = oracle_lib.Oracle(m, p, B, nthreads=4)
    pose = object_poses(B, 0)
    pose[:, 2] = 0.651
    orc.reset(pose, pose[:, :3].copy())
    orc.state["shaping"][:] = 1
    orc.step(None, 1, 3, want_obs=False)
    hp = orc.state["hand_pose"].copy()
    for e, t in enumerate([(0.1, 0.25, 0.63), (0.1, 0.3, 0.66), (0.12, 0.1, 0.63), (0.1, 0.0, 0.64), (0.2, 0.3, 0.63), (0.1, -0.1, 0.63)]):
        hp[e, :3] = t                                   # hands towards the near rim of the table (x = 0.1)
    orc.state["hand_pose"][:] = hp
    op = orc.state["obj_pose"].copy()
    op[6, :3] = [0.11, 0.0, 0.651]                      # cube 1 cm inside the rim: clipped rim manifold (box-box against the slab)
    op[7, :3] = [0.092, -0.2, 0.651]                    # cube overhanging the rim: tips over it
    op[8, :3] = [0.2, -0.4 + 0.08, 0.026]               # cube on the floor, 0.5 cm from a table leg's face
    op[9, :3] = [0.115, 0.2, 0.651]                     # rim again, yawed 45 degrees
    op[9, 3:] = [0.0, 0.0, np.sin(np.pi / 8), np.cos(np.pi / 8)]
    orc.state["obj_pose"][:] = op
    orc.state["obj_vel"][8, 1] = -0.3                   # sliding towards the leg
    seen = set()
    for r in range(rounds):
        orc.step(None, 12 if r else 2, 3, want_obs=False)
        sync(orc, sim)
        orc.step(None, 1, 3, want_obs=False)
        sim.step_host(None, 1, 3, want_obs=False)
        g, o = sim.get("status"), orc.state["status"]
        np.testing.assert_array_equal(g[:, 2:], o[:, 2:], err_msg="n_contacts / n_rows, round %d" % r)
        np.testing.assert_array_equal(g[:, 0] & 6, o[:, 0] & 6, err_msg="overflow flags, round %d" % r)
        np.testing.assert_array_equal(sim.get("cache_key"), orc.state["cache_key"], err_msg="contact keys (order included), round %d" % r)
        keys = orc.state["cache_key"]
        for name, lo, hi in (("cube-plane", 8, 16), ("sphere-sbox", 128, 256), ("sphere-plane", 256, 512), ("cube-sbox", 4096, 12288)):
            if ((keys >= lo) & (keys < hi)).any():
                seen.add(name)
        # converged systems whose IK loop stopped at the same iteration on both sides (see check_state)
        conv = (o[:, 1] < 150) & (np.abs(sim.get("mtarget") - orc.state["mtarget"]).max(axis=1) <= 2e-4)
        for f, tol in (("q", 2e-4), ("obj_pose", 2e-4)):
            err = np.abs(sim.get(f) - orc.state[f])
            assert err[conv].max() <= tol, (r, f, err[conv].max())
            assert np.isfinite(sim.get(f)).all() and err.max() < 5e-3, (r, f, err.max())
    assert {"cube-sbox", "cube-plane", "sphere-sbox"} <= seen, seen
    return orc, sim

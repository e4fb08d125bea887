"""GPU-vs-oracle parity of the fused step kernel, through the C-ABI (libb2env.so).

Tolerances (fp32 on both sides, different but mathematically equivalent formulations:
oracle = ABA + delta-velocity PGS, GPU = CRBA/Gauss-Jordan + Delassus-form PGS):
  * single step from identical state: joint/object state 2e-4 abs (solver exits on a velocity
    residual of sqrt(1e-7) ~ 3e-4 m/s, i.e. ~1.3e-6 m per step), contact keys exact;
  * trajectories: stated per test.
"""
import numpy as np
import pytest

from common import TASK_PUSH, TASK_REACH, copy_state_to_gpu, panda_task_setup, sample_object_poses, targets_for

pytestmark = pytest.mark.gpu


def _mk(task, B, oracle_lib, seed=0, nthreads=8):
    from pybullet_robot_envs.b2env.binding import B2Sim, OPT_RECORD_CONTACTS
    m, p = panda_task_setup(task)
    orc = oracle_lib.Oracle(m, p, B, nthreads=nthreads)
    sim = B2Sim(m, p, B, 0)
    sim.set_option(OPT_RECORD_CONTACTS, 1)
    pose = sample_object_poses(B, seed)
    tg = targets_for(pose, z=0.65)
    orc.reset(pose, tg)
    sim.reset_host(pose, tg)
    return m, p, orc, sim


def test_reset_state_matches(oracle_lib):
    m, p, orc, sim = _mk(TASK_PUSH, 64, oracle_lib)
    for f in ("q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "cache_key", "cache_lam"):
        np.testing.assert_array_equal(sim.get(f), orc.state[f], err_msg=f)


def test_single_step_from_identical_state(oracle_lib):
    """Every step restarts the GPU from the oracle's state: isolates per-step error."""
    B = 256
    m, p, orc, sim = _mk(TASK_PUSH, B, oracle_lib)
    rng = np.random.RandomState(1)
    worst = {}
    for i in range(160):
        copy_state_to_gpu(orc, sim)
        if i < 60:
            orc.step(None, 1, 1, want_obs=False)
            sim.step_host(None, 1, 1, want_obs=False)
        else:
            a = rng.uniform(-1, 1, (B, 7)).astype(np.float32)
            o_obs, o_rew, o_done = orc.step(a, 1, 0)
            g_obs, g_rew, g_done = sim.step_host(a, 1, 0)
            np.testing.assert_allclose(g_obs, o_obs, atol=2e-2, rtol=0, err_msg="obs step %d" % i)
            np.testing.assert_allclose(g_rew, o_rew, atol=1e-3, rtol=1e-5, err_msg="reward step %d" % i)
            np.testing.assert_array_equal(g_done, o_done)
        for f, tol in (("q", 2e-5), ("qd", 5e-3), ("obj_pose", 2e-5), ("obj_vel", 5e-3), ("mtarget", 1e-6)):
            err = np.abs(sim.get(f) - orc.state[f]).max()
            worst[f] = max(worst.get(f, 0), err)
            assert err <= tol, (f, i, err)
        # contact feature keys and counts are integers: exact
        g_st, o_st = sim.get("status"), orc.state["status"]
        np.testing.assert_array_equal(g_st[:, 2:], o_st[:, 2:], err_msg="n_contacts/n_rows step %d" % i)
        np.testing.assert_array_equal(sim.get("cache_key"), orc.state["cache_key"], err_msg="keys step %d" % i)
        np.testing.assert_array_equal(sim.get("counters"), orc.state["counters"])
    print("worst single-step abs errors:", worst)


def test_settle_trajectory(oracle_lib):
    """reset + 101 settle steps run free on both sides (panda_push_gym_env.py:132-148 analogue)."""
    B = 128
    m, p, orc, sim = _mk(TASK_PUSH, B, oracle_lib)
    for _ in range(101):
        orc.step(None, 1, 1, want_obs=False)
    sim.step_host(None, 101, 1, want_obs=False)  # 101 sub-steps fused in one launch
    np.testing.assert_allclose(sim.get("q"), orc.state["q"], atol=1e-5)
    np.testing.assert_allclose(sim.get("obj_pose"), orc.state["obj_pose"], atol=2e-4)
    # the cube rests on the table: centre at table top + half extent - slop-level penetration
    assert np.all(np.abs(sim.get("obj_pose")[:, 2] - 0.65) < 1e-3)
    assert np.all(np.abs(sim.get("obj_vel")) < 5e-3)
    np.testing.assert_array_equal(np.sort(sim.get("cache_key"), axis=1), np.sort(orc.state["cache_key"], axis=1))


@pytest.mark.parametrize("task", [TASK_PUSH, TASK_REACH])
def test_free_running_rollout(oracle_lib, task):
    """240 random-action steps, both sides free running.  Arm states are motor-dominated and stay
    within 1e-3 rad; observations within 5e-2 (scaled units) for envs without robot contact."""
    B = 128
    m, p, orc, sim = _mk(task, B, oracle_lib, seed=3)
    for _ in range(101):
        orc.step(None, 1, 1, want_obs=False)
    sim.step_host(None, 101, 1, want_obs=False)
    rng = np.random.RandomState(7)
    touched = np.zeros(B, bool)
    for i in range(240):
        a = rng.uniform(-1, 1, (B, 7)).astype(np.float32)
        o_obs, o_rew, o_done = orc.step(a, 1, 0)
        g_obs, g_rew, g_done = sim.step_host(a, 1, 0)
        touched |= orc.state["status"][:, 2] > 4
        touched |= sim.get("status")[:, 2] > 4
    ok = ~touched
    assert ok.sum() > B // 2
    assert np.abs(sim.get("q") - orc.state["q"])[ok].max() < 1e-3
    assert np.abs(sim.get("obj_pose") - orc.state["obj_pose"])[ok].max() < 1e-3
    assert np.abs(g_obs - o_obs)[ok][:, [0, 1, 2, 3, 4, 5] + list(range(9, 24))].max() < 5e-2
    assert np.abs(g_rew - o_rew)[ok].max() < 1e-2
    assert (sim.get("status")[:, 0] & 1).sum() == 0  # no NaN flag


def test_full_batch_properties():
    """BASELINE batch (16384): size-independent properties — no NaN flags, unit quaternions,
    joint limits respected, resting cubes stay on the table, determinism of a repeated run."""
    from pybullet_robot_envs.b2env.binding import B2Sim
    B = 16384
    m, p = panda_task_setup(TASK_PUSH)
    pose = sample_object_poses(B, 11)
    tg = targets_for(pose, z=0.65)
    outs = []
    for rep in range(2):
        sim = B2Sim(m, p, B, 0)
        sim.reset_host(pose, tg)
        sim.step_host(None, 101, 1, want_obs=False)
        rng = np.random.RandomState(5)
        for i in range(20):
            a = rng.uniform(-1, 1, (B, 7)).astype(np.float32)
            # rep 0: single launch from pageable host arrays; rep 1: page-locked, chunk-pipelined path (4 launches)
            obs, rew, done = sim.step_host(a, 1, 0) if rep == 0 else sim.step_pinned(a, 1, 0)
        st = sim.get("status")
        assert (st[:, 0] & 1).sum() == 0
        q = sim.get("q")
        lo = np.array([m.lower[i] for i in range(9)]) - 1e-3
        hi = np.array([m.upper[i] for i in range(9)]) + 1e-3
        assert np.all(q >= lo) and np.all(q <= hi)
        qt = sim.get("obj_pose")[:, 3:]
        np.testing.assert_allclose(np.linalg.norm(qt, axis=1), 1.0, atol=1e-5)
        assert np.isfinite(obs).all() and np.isfinite(rew).all()
        outs.append((q.copy(), obs.copy(), rew.copy()))
        sim.close()
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(a, b)  # bitwise deterministic, and identical through both host paths


def _contact_rich_states(oracle_lib, m, p, B, seed):
    """Joint states that put the hand / fingers at the cube and on the table: oracle IK towards points
    around the resting cube, so that sphere-cube and sphere-table contacts, coupled islands, limit rows
    and more than 16 generic rows occur."""
    import ctypes as C
    import math
    o1 = oracle_lib.Oracle(m, p, 1)
    o1.lib.b2o_ik.restype = C.c_int
    rng = np.random.RandomState(seed)
    home = np.array([m.home[i] for i in range(9)], np.float32)
    tq = np.array([1.0, 0.0, 0.0, 6.123234e-17], np.float32)        # hand pointing down (euler pi,0,0)
    qs = np.zeros((B, 9), np.float32)
    poses = sample_object_poses(B, seed)
    poses[:, 2] = 0.65
    for b in range(B):
        off = rng.uniform(-0.04, 0.04, 3).astype(np.float32)
        off[2] = rng.uniform(0.0, 0.09)
        tp = (poses[b, :3] + off).astype(np.float32)
        out = np.zeros(9, np.float32)
        o1.lib.b2o_ik(C.byref(m), C.byref(p), home.ctypes.data_as(C.c_void_p), tp.ctypes.data_as(C.c_void_p),
                      tq.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        out[7:] = rng.uniform(0.0, 0.04, 2)
        qs[b] = np.clip(out, [m.lower[i] for i in range(9)], [m.upper[i] for i in range(9)])
    return qs, poses


def test_contact_rich_single_step_parity(oracle_lib):
    """Robot-cube and robot-table contacts: coupled islands, serial motor rows, 2-3 generic row sets.
    GPU restarts from the oracle state every step.  Contact keys / counts / row counts must be exact;
    states within the single-step tolerances (contacts are stiff: velocities 2e-2, positions 1e-4)."""
    from pybullet_robot_envs.b2env.binding import B2Sim, OPT_RECORD_CONTACTS
    B = 256
    m, p = panda_task_setup(TASK_PUSH)
    orc = oracle_lib.Oracle(m, p, B, nthreads=8)
    sim = B2Sim(m, p, B, 0)
    sim.set_option(OPT_RECORD_CONTACTS, 1)
    qs, poses = _contact_rich_states(oracle_lib, m, p, B, 21)
    tg = targets_for(poses) + np.array([0.3, 0, 0], np.float32)
    orc.reset(poses, tg)
    orc.state["q"][:] = qs
    orc.state["mtarget"][:] = qs
    rng = np.random.RandomState(5)
    seen_rows, seen_coupled, seen_limit = 0, 0, 0
    n_lim, n_lim_close, n_res_close, n_state_far = 0, 0, 0, 0
    kp = np.array([p.kp_ctrl] * 7 + [p.kp_hold] * 2, np.float32)
    res_worst = 0.0
    for i in range(40):
        copy_state_to_gpu(orc, sim)
        a = rng.uniform(-1, 1, (B, 7)).astype(np.float32)
        q0 = orc.state["q"].copy()
        o_obs, o_rew, o_done = orc.step(a, 1, 0)
        g_obs, g_rew, g_done = sim.step_host(a, 1, 0)
        g_st, o_st = sim.get("status"), orc.state["status"]
        np.testing.assert_array_equal(g_st[:, 2:], o_st[:, 2:], err_msg="n_contacts / n_rows, step %d" % i)
        np.testing.assert_array_equal(g_st[:, 0] & 6, o_st[:, 0] & 6, err_msg="overflow flags, step %d" % i)
        np.testing.assert_array_equal(sim.get("cache_key"), orc.state["cache_key"], err_msg="contact keys, step %d" % i)
        assert (g_st[:, 0] & 1).sum() == 0
        ok = np.isfinite(orc.state["q"]).all(axis=1)
        assert ok.all()
        dq = np.abs(sim.get("q") - orc.state["q"]).max(axis=1)
        dv = np.abs(sim.get("qd") - orc.state["qd"]).max(axis=1)
        dc = np.abs(sim.get("obj_pose") - orc.state["obj_pose"]).max(axis=1)
        # iteration-limited envs (150 sweeps without meeting the residual) agree less tightly: both sides stop a
        # non-converged Gauss-Seidel at the same sweep, rounding differences are not damped out
        conv = o_st[:, 1] < 150
        assert dq[conv].max() < 1e-4 and dc[conv].max() < 1e-4 and dv[conv].max() < 2e-2, (i, dq[conv].max(), dv[conv].max(), dc[conv].max())
        # jammed configurations (cube squeezed between a robot sphere and the table) hit the 150-sweep cap without
        # converging: both sides stop the same Gauss-Seidel iteration at sweep 150, so the constraint residual they are
        # left with must agree — the violation of the position-motor rows |qd+ - kp (target - q) / dt| (rad/s): 90 % of them
        # within 25 % + 0.1, the rest finite — and the states within 5e-2 but for rare outliers, every one within
        # 0.2 (the truncated iterate is ill-conditioned: rounding is not damped out)
        assert np.isfinite(sim.get("obj_pose")).all() and dq.max() < 0.2 and dc.max() < 0.2, (i, dq.max(), dc.max())
        n_state_far += int(((dq >= 5e-2) | (dc >= 5e-2)).sum())
        if (~conv).any():
            cap = ~conv
            r_o = np.abs(orc.state["qd"] - kp * (orc.state["mtarget"] - q0) / p.dt).max(axis=1)[cap]
            r_g = np.abs(sim.get("qd") - kp * (sim.get("mtarget") - q0) / p.dt).max(axis=1)[cap]
            res_worst = max(res_worst, float((np.abs(r_g - r_o) / (0.25 * r_o + 0.1)).max()))
            n_res_close += int((np.abs(r_g - r_o) <= 0.25 * r_o + 0.1).sum())
            assert np.isfinite(r_g).all()   # the few that disagree more are truncated iterates of ill-conditioned systems: finite, counted
        n_lim += int((~conv).sum())
        n_lim_close += int(((~conv) & (dq < 2e-3) & (dc < 2e-3)).sum())
        seen_rows = max(seen_rows, int(o_st[:, 3].max()))
        keys = orc.state["cache_key"]
        seen_coupled += int((((keys >= 16) & (keys < 32)) | ((keys >= 12288) & (keys < 16384)) | ((keys >= 576) & (keys < 580))).any(axis=1).sum())
        seen_limit += int((o_st[:, 3] - 9 - 3 * o_st[:, 2] > 0).sum())
    assert n_res_close >= 0.9 * n_lim, (n_res_close, n_lim)   # sweep-capped: 90 % agree on the residual within 25 % + 0.1
    assert n_state_far <= 0.02 * max(n_lim, 1) + 1, (n_state_far, n_lim)   # states beyond 5e-2: rare outliers of truncated iterates
    assert seen_rows > 9 + 16 + 16, seen_rows        # three generic row sets were exercised
    assert seen_coupled > 50 and seen_limit > 0, (seen_coupled, seen_limit)
    print("contact-rich: max rows %d, coupled env-steps %d, limit-row env-steps %d; sweep-capped env-steps %d of which %d within 2e-3; "
          "worst motor-residual disagreement %.3f of its bound" % (seen_rows, seen_coupled, seen_limit, n_lim, n_lim_close, res_worst))


def test_contact_families_parity(oracle_lib):
    """N1: every contact family of the collision stage on hardware — cube at the table rim / against a leg / on the ground
    plane (box-box SAT + clipping), finger-pad boxes vs cube / table / static boxes, spheres vs cube / table / rim, robot
    self-collision, forearm capsule vs cube / static boxes (GJK / EPA) — GPU vs oracle from identical states: contact keys,
    contact and row counts exact; converged environments within the single-step tolerances."""
    from common import FAMILIES, family_states
    from pybullet_robot_envs.b2env.binding import B2Sim, OPT_RECORD_CONTACTS
    m, p = panda_task_setup(TASK_PUSH)
    qs, poses, fam = family_states(oracle_lib, m, p, list(FAMILIES), per_family=24, seed=11)
    B = len(fam)
    orc = oracle_lib.Oracle(m, p, B, nthreads=8)
    sim = B2Sim(m, p, B, 0)
    sim.set_option(OPT_RECORD_CONTACTS, 1)
    orc.reset(poses, targets_for(poses))
    orc.state["q"][:] = qs
    orc.state["mtarget"][:] = qs
    seen = {f: 0 for f in FAMILIES}
    rng = np.random.RandomState(2)
    worst = 0.0
    for i in range(12):
        copy_state_to_gpu(orc, sim)
        a = rng.uniform(-1, 1, (B, 7)).astype(np.float32) * (0.2 if i < 6 else 1.0)
        orc.step(a, 1, 0)
        sim.step_host(a, 1, 0)
        g_st, o_st = sim.get("status"), orc.state["status"]
        np.testing.assert_array_equal(g_st[:, 2:], o_st[:, 2:], err_msg="n_contacts / n_rows, step %d" % i)
        np.testing.assert_array_equal(g_st[:, 0] & 6, o_st[:, 0] & 6, err_msg="overflow flags, step %d" % i)
        np.testing.assert_array_equal(sim.get("cache_key"), orc.state["cache_key"], err_msg="contact keys, step %d" % i)
        keys = orc.state["cache_key"]
        for f, (lo, hi) in FAMILIES.items():
            seen[f] += int(((keys >= lo) & (keys < hi)).any(axis=1).sum())
        conv = o_st[:, 1] < 150
        dq = np.abs(sim.get("q") - orc.state["q"]).max(axis=1)
        dc = np.abs(sim.get("obj_pose") - orc.state["obj_pose"]).max(axis=1)
        assert np.isfinite(sim.get("q")).all() and np.isfinite(sim.get("obj_pose")).all()
        assert dq[conv].max() < 2e-4 and dc[conv].max() < 2e-4, (i, dq[conv].max(), dc[conv].max())
        assert dq.max() < 5e-2 and dc.max() < 5e-2, (i, dq.max(), dc.max())
        worst = max(worst, float(dq[conv].max()), float(dc[conv].max()))
    assert all(v >= 12 for v in seen.values()), seen
    print("contact families (env-steps with the family present):", seen, "worst converged |dq|, |dpose| %.1e" % worst)

"""Behavioural known answers for the round-2 contact families (oracle, CPU only): what a cube does at the table rim, against a
table leg and on the ground plane; what the robot's box / capsule / self-collision proxies do when the arm is driven into the
table or into itself.  These pin the PHYSICS the families produce (a stable rest, a fall, a stop), which the GPU-vs-oracle
parity tests cannot: both sides share the narrowphase source and the statement."""
import numpy as np

from common import FAMILIES, TASK_PUSH, family_states, panda_task_setup


def _oracle(oracle_lib, B, **kw):
    m, p = panda_task_setup(TASK_PUSH, **kw)
    return m, p, oracle_lib.Oracle(m, p, B, nthreads=4)


def _place(orc, poses):
    orc.reset(poses, poses[:, :3].copy())


def test_cube_at_the_rim_rests_or_falls(oracle_lib):
    """Table top: x in [0.1, 1.6].  A cube whose centre is 1 cm INSIDE the rim rests on the clipped manifold (two cube vertices +
    two rim crossings); with its centre 1 cm OUTSIDE it tips over the rim and ends on the ground plane."""
    m, p, orc = _oracle(oracle_lib, 2)
    poses = np.zeros((2, 7), np.float32)
    poses[:, 6] = 1.0
    poses[0, :3] = [0.11, 0.0, 0.65]
    poses[1, :3] = [0.09, 0.0, 0.65]
    _place(orc, poses)
    orc.step(None, 60, 1, want_obs=False)
    pose = orc.state["obj_pose"]
    assert abs(pose[0, 2] - 0.65) < 2e-3 and abs(pose[0, 0] - 0.11) < 2e-3, pose[0]           # at rest on the rim manifold
    keys = orc.state["cache_key"][0]
    assert ((keys >= 4096) & (keys < 4096 + 1024)).sum() == 4, keys                           # cube vs static box 0, general path
    orc.step(None, 400, 1, want_obs=False)
    pose = orc.state["obj_pose"]
    assert abs(pose[0, 2] - 0.65) < 2e-3                                                      # still there
    assert pose[1, 2] < 0.1 and pose[1, 0] < 0.1, pose[1]                                      # fell off, lies beside the table
    assert abs(pose[1, 2] - 0.025) < 5e-3 or pose[1, 2] < 0.06                                 # on the ground plane (any face)
    assert np.abs(orc.state["obj_vel"][1]).max() < 0.05                                        # and came to rest


def test_cube_on_the_floor_stops_at_a_leg(oracle_lib):
    """A cube sliding on the ground plane towards a table leg (0.1 x 0.1 m at (0.2, -0.4)) is stopped by the leg's face."""
    m, p, orc = _oracle(oracle_lib, 2)
    poses = np.zeros((2, 7), np.float32)
    poses[:, 6] = 1.0
    poses[0, :3] = [0.2, -0.4 + 0.105, 0.025]      # in front of the leg's +y face, 3 cm of clearance (friction stops a 1 m/s cube in 5 cm)
    poses[1, :3] = [0.45, -0.4 + 0.105, 0.025]     # same motion, no leg in the way (x = 0.45: between the legs)
    _place(orc, poses)
    orc.step(None, 5, 1, want_obs=False)
    v = orc.state["obj_vel"]
    v[:, 1] = -1.0
    orc.state["obj_vel"][:] = v
    orc.step(None, 120, 1, want_obs=False)
    pose = orc.state["obj_pose"]
    assert pose[0, 1] > -0.4 + 0.05 + 0.025 - 2e-3, pose[0]          # never beyond the leg's face (centre >= face + half size)
    assert abs(pose[0, 1] - (-0.4 + 0.05 + 0.025)) < 1e-3, pose[0]    # and it rests against that face
    assert pose[1, 1] < pose[0, 1] - 0.01, pose                       # the free cube slid further


def test_arm_pressed_onto_the_table_is_held_by_its_proxies(oracle_lib):
    """Postures in which a finger pad, a hand sphere or the forearm capsule touches the table, with motor targets 0.3 rad further
    along the joint direction that lowers the hand fastest: the contacts of the box / sphere / capsule families hold every proxy
    at the table top (penetration below 5 mm after the transient) against the position motors."""
    import ctypes as C
    m, p = panda_task_setup(TASK_PUSH)
    qs, poses, fam = family_states(oracle_lib, m, p, ["pad_table", "sphere_table", "cap_sbox"], per_family=3, seed=9)
    B = len(fam)
    poses[:, :3] = [1.2, 0.3, 0.65]                 # cube out of the way
    orc = oracle_lib.Oracle(m, p, B, nthreads=4)
    orc.reset(poses, poses[:, :3].copy())
    orc.state["q"][:] = qs
    tg = qs.copy()
    lo = np.array([m.lower[i] for i in range(9)], np.float32)
    hi = np.array([m.upper[i] for i in range(9)], np.float32)
    for b in range(B):
        z0 = orc.fk(qs[b])[0][8][2]                 # hand link height
        g = np.zeros(7)
        for j in range(7):
            dq = qs[b].copy()
            dq[j] += 1e-3
            g[j] = (orc.fk(dq)[0][8][2] - z0) / 1e-3
        tg[b, :7] = qs[b, :7] - 0.3 * g / max(np.linalg.norm(g), 1e-6)
    tg = np.clip(tg, lo, hi)
    orc.state["mtarget"][:] = tg
    out = np.zeros((12, 16), np.float32)
    orc.lib.b2o_collide.restype = C.c_int

    def deepest():
        """per env: (deepest rigid robot-vs-static contact, deepest soft finger-pad contact), 0 where there is none"""
        rigid, soft, fams = np.zeros(B), np.zeros(B), set()
        for b in range(B):
            ov = C.c_int(0)
            q = np.ascontiguousarray(orc.state["q"][b])
            pose = np.ascontiguousarray(orc.state["obj_pose"][b])
            n = orc.lib.b2o_collide(C.byref(m), C.byref(p), q.ctypes.data_as(C.c_void_p), pose.ctypes.data_as(C.c_void_p),
                                    out.ctypes.data_as(C.c_void_p), C.byref(ov))
            for i in range(n):
                k, d = int(out[i, 0]), float(out[i, 13])
                if int(out[i, 1]) != 2:               # CT_ARM_STATIC only
                    continue
                pad = (64 <= k < 128) or k >= 16384   # finger-pad vertices / pad box-box: soft contacts (URDF stiffness / damping)
                if pad:
                    soft[b] = min(soft[b], d)
                else:
                    rigid[b] = min(rigid[b], d)
                fams |= {f for f, (a, c) in FAMILIES.items() if a <= k < c}
        return rigid, soft, fams

    orc.step(None, 300, 1, want_obs=False)
    r1, s1, f1 = deepest()
    orc.step(None, 300, 1, want_obs=False)
    r2, s2, f2 = deepest()
    assert ((r2 < 0) | (s2 < 0)).sum() >= B - 2       # the arm is still pressed onto the table
    # rigid proxies (spheres, forearm capsule; erp 0.2): held at the surface
    assert r2.min() > -5e-3, r2
    # finger pads are SOFT contacts (panda_model.urdf:256-263: stiffness 30000, damping 1000 -> erp 0.11, cfm 0.21): pressed by
    # a position motor they sink until the contact spring balances it, and stay there — a static equilibrium inside the slab
    assert s2.min() > -0.05 and np.abs(s2 - s1).max() < 2e-3, (s1, s2)
    assert np.isfinite(orc.state["q"]).all() and np.abs(orc.state["qd"]).max() < 1.0
    assert len((f1 | f2) & {"pad_table", "pad_sbox", "sphere_table", "cap_sbox", "sphere_sbox"}) >= 2, f1 | f2


def test_self_collision_keeps_the_hand_off_the_upper_arm(oracle_lib):
    """URDF_USE_SELF_COLLISION (panda_env.py:53): states that start with a hand / wrist proxy inside a shoulder / upper-arm
    proxy are pushed apart; with the self pairs removed from the model the same states stay interpenetrating."""
    m, p = panda_task_setup(TASK_PUSH)
    qs, poses, fam = family_states(oracle_lib, m, p, ["self"], per_family=6, seed=4)
    gaps = {}
    for with_pairs in (True, False):
        m2, p2 = panda_task_setup(TASK_PUSH)
        if not with_pairs:
            m2.n_self_pairs = 0
        orc = oracle_lib.Oracle(m2, p2, len(fam), nthreads=4)
        orc.reset(poses, poses[:, :3].copy())
        orc.state["q"][:] = qs
        # targets deeper into the contact: every joint 0.1 rad further along the direction that closes the gap is not known in
        # general, so simply hold the start posture: the contact alone must open the overlap
        orc.state["mtarget"][:] = qs
        orc.step(None, 240, 1, want_obs=False)
        # smallest gap over the model's self pairs (of the full model), from the oracle's forward kinematics
        g = []
        for b in range(len(fam)):
            pos, rot = orc.fk(orc.state["q"][b])
            c = np.array([pos[m.sph_link[s]] + rot[m.sph_link[s]] @ np.array(m.sph_c[s][:]) for s in range(m.n_spheres)])
            g.append(min(np.linalg.norm(c[m.self_a[k]] - c[m.self_b[k]]) - m.sph_r[m.self_a[k]] - m.sph_r[m.self_b[k]]
                         for k in range(m.n_self_pairs)))
        gaps[with_pairs] = np.array(g)
    assert gaps[True].min() > -3e-3, gaps[True]                    # pushed out to (almost) touching
    assert gaps[True].min() > gaps[False].min() + 1e-3 or gaps[False].min() > -3e-3, (gaps[True], gaps[False])


def test_sliding_cube_decelerates_at_mu_g_per_friction_direction(oracle_lib):
    """Coulomb friction through the solver rows, closed form: a cube sliding on the table top loses mu g dt of speed per step
    along each of the two friction directions of the contact (Bullet's friction pyramid: the rows are bounded by mu x normal
    impulse separately, so a diagonal slide is braked sqrt(2) harder than an axis-aligned one), on top of the base damping
    v <- v (1 - dt (k1 + k2 |v|)); the normal rows carry m g dt in total, the cube neither lifts nor sinks nor turns."""
    mu, v0, n = 0.3, 0.3, 10
    m, p = panda_task_setup(TASK_PUSH)
    p.cube_mu = mu                                   # table friction 1.0: combined coefficient = mu (product rule)
    orc = oracle_lib.Oracle(m, p, 3, nthreads=1)
    poses = np.zeros((3, 7), np.float32)
    poses[:, 6] = 1.0
    poses[:, :3] = [0.8, 0.0, 0.65]                  # far from the arm and the rim
    _place(orc, poses)
    orc.step(None, 40, 1, want_obs=False)            # settle on the four-point manifold
    z0 = orc.state["obj_pose"][:, 2].copy()
    v = orc.state["obj_vel"].copy()
    v[0, 0] = v0                                     # along x (a friction direction of a z-normal contact)
    v[1, 1] = -v0                                    # along -y (the other one)
    v[2, 0] = v[2, 1] = v0 / np.sqrt(2.0)            # diagonal, same speed
    orc.state["obj_vel"][:] = v
    dt, g = p.dt, 9.81
    ex = np.array([v0, v0, v0 / np.sqrt(2.0)])       # per-axis speeds
    speed = np.array([v0, v0, v0])
    for _ in range(n):
        orc.step(None, 1, 1, want_obs=False)
        damp = 1.0 - dt * (p.damp_lin_k1 + p.damp_lin_k2 * speed)
        ex = ex * damp - mu * g * dt
        speed = np.array([ex[0], ex[1], ex[2] * np.sqrt(2.0)])
    got = orc.state["obj_vel"]
    np.testing.assert_allclose(got[0, 0], ex[0], atol=1e-3)
    np.testing.assert_allclose(got[1, 1], -ex[1], atol=1e-3)
    np.testing.assert_allclose(got[2, :2], [ex[2], ex[2]], atol=1e-3)
    assert abs(ex[0] - (v0 - n * mu * g * dt)) < 2e-3                     # (the damping terms are a small correction)
    assert np.abs(got[0, 1:3]).max() < 2e-3 and np.abs(got[1, [0, 2]]).max() < 2e-3     # no sideways or vertical drift
    assert np.abs(orc.state["obj_pose"][:, 2] - z0).max() < 2e-4
    lam = orc.state["cache_lam"].reshape(3, -1, 3)                                        # [B, slots, 3]: normal, friction 1, friction 2
    np.testing.assert_allclose(lam[:, :, 0].sum(axis=1), p.cube_mass * g * dt, rtol=2e-2)   # the normal rows carry the weight
    # every friction row of the sliding cubes sits on its bound mu x normal impulse
    np.testing.assert_allclose(np.abs(lam[0, :, 1:]).max(axis=1), mu * lam[0, :, 0], atol=1e-6)
    orc.step(None, 40, 1, want_obs=False)
    assert np.abs(orc.state["obj_vel"][:, :3]).max() < 1e-3                              # friction brings all three to rest

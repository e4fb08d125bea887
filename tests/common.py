"""Shared helpers for the parity tests (seeded synthetic inputs for both backends)."""
import numpy as np

from pybullet_robot_envs.b2env.model import TASK_GRASP, TASK_PUSH, TASK_REACH, panda_task_setup  # noqa: F401


def sample_object_poses(B, seed=0, rnd=0.05, z=0.695):
    """Object start poses as WorldEnv._sample_pose draws them (world_env.py:145-176): uniform
    half-width `rnd` around (0.45, 0), yaw uniform in +-pi/4."""
    rng = np.random.RandomState(seed)
    pose = np.zeros((B, 7), np.float32)
    pose[:, 0] = np.clip(0.45 + rng.uniform(-rnd, rnd, B), 0.35, 0.55)
    pose[:, 1] = np.clip(rng.uniform(-rnd, rnd, B), -0.25, 0.25)
    pose[:, 2] = z
    yaw = rng.uniform(-np.pi / 4, np.pi / 4, B)
    pose[:, 5] = np.sin(yaw / 2)
    pose[:, 6] = np.cos(yaw / 2)
    return pose


def targets_for(pose, z=None):
    tg = pose[:, :3].copy()
    tg[:, 0] += 0.05
    tg[:, 1] += 0.05
    if z is not None:
        tg[:, 2] = z
    return tg.astype(np.float32)


STATE_FIELDS = ["q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "cache_key", "cache_lam", "status"]


def copy_state_to_gpu(oracle, sim, fields=("q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters",
                                            "cache_key", "cache_lam", "hand_pose")):
    for f in fields:
        sim.set(f, oracle.state[f])


# contact families by key range (include/b2env.h: B2E_KEY_*)
FAMILIES = {
    "cube_rim_or_leg": (4096, 12288),     # cube vs static box, general box-box (rim of the top slab, legs)
    "cube_plane": (8, 16),
    "pad_cube": (12288, 16384),           # finger-pad box vs cube, box-box
    "pad_table": (64, 96),                # finger-pad vertices on the table top
    "pad_sbox": (16384, 1 << 30),         # finger-pad box vs static box, general box-box
    "sphere_cube": (16, 32),
    "sphere_table": (32, 48),
    "sphere_sbox": (128, 256),            # sphere vs a static box other than through the table top (rim, legs)
    "self": (512, 544),                   # robot self-collision
    "cap_cube": (576, 580),               # forearm capsule vs cube (GJK / EPA)
    "cap_sbox": (592, 624),               # forearm capsule vs static box (GJK / EPA)
}


def family_states(oracle_lib, m, p, want, per_family=8, seed=0, max_tries=60000):
    """Joint states + object poses whose contact set (oracle collide()) holds at least one contact of each family in `want`:
    rejection sampling over random joint vectors / hand-placed object poses.  Returns (q [n, 9], obj_pose [n, 7], the
    family each row was kept for)."""
    import ctypes as C
    o1 = oracle_lib.Oracle(m, p, 1)
    o1.lib.b2o_collide.restype = C.c_int
    o1.lib.b2o_fk.restype = C.c_int
    rng = np.random.RandomState(seed)
    lo = np.array([m.lower[i] for i in range(9)], np.float32)
    hi = np.array([m.upper[i] for i in range(9)], np.float32)
    home = np.array([m.home[i] for i in range(9)], np.float32)
    out = np.zeros((12, 16), np.float32)
    ov = C.c_int(0)
    lp, lr = np.zeros((m.n_links, 3), np.float32), np.zeros((m.n_links, 9), np.float32)

    def keys_of(q, pose):
        n = o1.lib.b2o_collide(C.byref(m), C.byref(p), q.ctypes.data_as(C.c_void_p), pose.ctypes.data_as(C.c_void_p),
                               out.ctypes.data_as(C.c_void_p), C.byref(ov))
        return out[:n, 0].astype(np.int64), out[:n, 13]

    def quat_z(yaw):
        return [0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2)]

    got = {f: [] for f in want}
    for t in range(max_tries):
        if all(len(v) >= per_family for v in got.values()):
            break
        f = [k for k in want if len(got[k]) < per_family][t % len([k for k in want if len(got[k]) < per_family])]
        q = home.copy()
        pose = np.array([0.45, 0.0, 0.65, 0, 0, 0, 1], np.float32)
        pose[3:] = quat_z(rng.uniform(-0.7, 0.7))
        if f == "cube_rim_or_leg":
            if rng.rand() < 0.6:   # over the rim x = 0.1 / y = +-0.5
                if rng.rand() < 0.5:
                    pose[:3] = [0.1 + rng.uniform(-0.02, 0.02), rng.uniform(-0.3, 0.3), 0.65 + rng.uniform(-0.001, 0.002)]
                else:
                    pose[:3] = [rng.uniform(0.2, 0.6), np.sign(rng.rand() - 0.5) * (0.5 + rng.uniform(-0.02, 0.02)), 0.65]
            else:                  # on the ground against a leg
                pose[:3] = [0.2 + rng.uniform(-0.02, 0.02), -0.4 + 0.075 + rng.uniform(-0.003, 0.003), 0.025 + rng.uniform(-0.001, 0.001)]
        elif f == "cube_plane":
            pose[:3] = [rng.uniform(-0.3, 0.05), rng.uniform(-0.3, 0.3), 0.025 + rng.uniform(-0.001, 0.002)]
        elif f in ("pad_cube", "pad_table", "sphere_cube", "sphere_table"):
            # hand above / around the cube: perturb a reaching posture
            q[:7] = np.array([0.0, 0.35, 0.0, -2.2, 0.0, 2.55, 0.8], np.float32) + rng.normal(0, 0.12, 7).astype(np.float32)
            q[7:] = rng.uniform(0.0, 0.04, 2)
            o1.lib.b2o_fk(C.byref(m), q.ctypes.data_as(C.c_void_p), lp.ctypes.data_as(C.c_void_p), lr.ctypes.data_as(C.c_void_p))
            tip = lp[9] + lr[9].reshape(3, 3) @ np.array([0, 0, 0.04], np.float32)
            pose[:3] = [tip[0] + rng.uniform(-0.03, 0.03), tip[1] + rng.uniform(-0.03, 0.03), 0.65]
        else:
            q = (lo + (hi - lo) * rng.uniform(0, 1, 9)).astype(np.float32)
            if f == "cap_cube":
                o1.lib.b2o_fk(C.byref(m), q.ctypes.data_as(C.c_void_p), lp.ctypes.data_as(C.c_void_p), lr.ctypes.data_as(C.c_void_p))
                mid = lp[5] + lr[5].reshape(3, 3) @ np.array([0, 0.07, -0.14], np.float32)
                d = rng.normal(size=3)
                d /= np.linalg.norm(d)
                pose[:3] = mid + d * (0.052 + 0.03 + rng.uniform(-0.01, 0.01))
                pose[3:] = rng.normal(size=4)
                pose[3:] /= np.linalg.norm(pose[3:])
        q = np.clip(q, lo, hi).astype(np.float32)
        keys, dist = keys_of(q, pose)
        a, b = FAMILIES[f]
        sel = (keys >= a) & (keys < b)
        if sel.any() and dist[sel].min() > -0.02 and len(keys) <= 12 and dist.min() > -0.03:
            got[f].append((q, pose.copy()))
    missing = [f for f, v in got.items() if len(v) < per_family]
    assert not missing, "family_states: no sample for %s" % missing
    rows = [(f, q, pose) for f, v in got.items() for q, pose in v[:per_family]]
    return (np.array([r[1] for r in rows], np.float32), np.array([r[2] for r in rows], np.float32), [r[0] for r in rows])

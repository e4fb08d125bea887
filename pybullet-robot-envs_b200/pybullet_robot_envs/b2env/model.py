"""Model descriptor: URDF -> flat arrays (the host half of what ``p.loadURDF`` does).

The reference loads ``robot_data/franka_panda/panda_model.urdf`` through
``p.loadURDF(..., useFixedBase=True, flags=URDF_USE_INERTIA_FROM_FILE | ...)``
(reference panda_env.py:53-56) and then addresses joints by PyBullet joint index
(= URDF joint order, panda_env.py:60-79).  This module parses the same file into
the ``b2e_model`` struct of include/b2env.h.  Because ``/root/reference`` does not
travel to the GPU box, ``tools/gen_model.py`` runs this parser once and commits the
result as ``robot_data/franka_panda/panda_model.json``; ``load_panda()`` reads that.

Collision geometry: the reference's collision meshes are git-LFS stubs,
so links carry sphere proxies defined in ``proxies.py`` (documented deviation).
"""
import ctypes as C
import json
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

MAX_LINKS = 40
MAX_DOF = 32
MAX_SPHERES = 16
MAX_OBS = 40
MAX_CONTACTS = 12
CACHE_SLOTS = 16
MAX_BOXES = 4
MAX_SELF_PAIRS = 32
MAX_SBOXES = 6
MAX_CAPS = 4

JOINT_REVOLUTE = 0
JOINT_PRISMATIC = 1
JOINT_FIXED = 4

TASK_REACH = 0
TASK_PUSH = 1
TASK_GRASP = 2
REWARD_PANDA, REWARD_ICUB_REACH, REWARD_ICUB_PUSH0, REWARD_ICUB_PUSH1 = 0, 1, 2, 3

_f = C.c_float
_i = C.c_int32


class B2EModel(C.Structure):
    """ctypes mirror of ``struct b2e_model`` (include/b2env.h)."""
    _fields_ = [
        ("n_links", _i), ("n_dof", _i), ("ee_link", _i), ("n_spheres", _i),
        ("parent", _i * MAX_LINKS), ("jtype", _i * MAX_LINKS), ("dof", _i * MAX_LINKS),
        ("jpos", (_f * 3) * MAX_LINKS), ("jrot", (_f * 9) * MAX_LINKS),
        ("axis", (_f * 3) * MAX_LINKS),
        ("mass", _f * MAX_LINKS), ("com", (_f * 3) * MAX_LINKS),
        ("inertia", (_f * 9) * MAX_LINKS),
        ("lower", _f * MAX_DOF), ("upper", _f * MAX_DOF), ("limit_margin", _f * MAX_DOF),
        ("max_force", _f * MAX_DOF), ("max_vel", _f * MAX_DOF),
        ("joint_damping", _f * MAX_DOF), ("home", _f * MAX_DOF),
        ("base_pos", _f * 3), ("base_rot", _f * 9),
        ("sph_link", _i * MAX_SPHERES), ("sph_c", (_f * 3) * MAX_SPHERES),
        ("sph_r", _f * MAX_SPHERES), ("sph_mu", _f * MAX_SPHERES),
        ("sph_erp", _f * MAX_SPHERES), ("sph_cfm", _f * MAX_SPHERES),
        ("n_boxes", _i), ("box_link", _i * MAX_BOXES), ("box_c", (_f * 3) * MAX_BOXES), ("box_h", (_f * 3) * MAX_BOXES),
        ("box_mu", _f * MAX_BOXES), ("box_erp", _f * MAX_BOXES), ("box_cfm", _f * MAX_BOXES),
        ("n_self_pairs", _i), ("self_a", _i * MAX_SELF_PAIRS), ("self_b", _i * MAX_SELF_PAIRS),
        ("sph_flags", _i * MAX_SPHERES),
        ("n_caps", _i), ("cap_link", _i * MAX_CAPS), ("cap_p0", (_f * 3) * MAX_CAPS), ("cap_p1", (_f * 3) * MAX_CAPS),
        ("cap_r", _f * MAX_CAPS), ("cap_mu", _f * MAX_CAPS),
    ]


class B2EParams(C.Structure):
    """ctypes mirror of ``struct b2e_params`` (include/b2env.h)."""
    _fields_ = [
        ("dt", _f), ("gravity", _f * 3), ("solver_iters", _i), ("residual_tol", _f),
        ("erp", _f), ("slop", _f), ("warmstart", _f), ("contact_margin", _f),
        ("table_min", _f * 3), ("table_max", _f * 3), ("table_mu", _f), ("plane_mu", _f),
        ("cube_half", _f), ("cube_mass", _f), ("cube_inertia", _f), ("cube_mu", _f),
        ("damp_lin_k1", _f), ("damp_lin_k2", _f), ("damp_ang_k1", _f), ("damp_ang_k2", _f),
        ("task", _i), ("n_act", _i), ("n_ctrl", _i), ("use_ik", _i), ("ik_orientation", _i),
        ("ik_iters", _i), ("ik_residual", _f), ("ik_damping", _f),
        ("act_scale", _f), ("act_scale_pos", _f), ("act_scale_rot", _f),
        ("kp_ctrl", _f), ("kp_hold", _f), ("dist_min", _f), ("max_steps", _i), ("n_obs", _i),
        ("obs_low", _f * MAX_OBS), ("obs_high", _f * MAX_OBS),
        ("vel_mean", _f * 3), ("vel_std", _f * 3),
        ("ws_lim", (_f * 2) * 3), ("eu_lim", (_f * 2) * 3), ("home_hand_pose", _f * 6),
        ("kp_grip", _f), ("grasp_lift", _f), ("grasp_rest_z", _f), ("goal_env", _i),
        ("n_obs_joints", _i), ("obs_dof", _i * 16), ("ctrl_dof", _i * 16), ("ctrl_mask", C.c_uint32),
        ("ik_link_offset", _f * 3), ("reward_kind", _i), ("max_contacts", _i),
        ("n_sboxes", _i), ("sbox_c", (_f * 3) * MAX_SBOXES), ("sbox_h", (_f * 3) * MAX_SBOXES), ("sbox_mu", _f * MAX_SBOXES),
        ("ik_max_vel", _f), ("kp_ik_max_vel", _f),
    ]


def rpy_to_matrix(r, p, y):
    """URDF fixed-axis roll-pitch-yaw -> rotation matrix R = Rz(y) Ry(p) Rx(r)."""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr]], dtype=np.float64)


def _vec(s, n=3):
    v = [float(x) for x in s.split()]
    assert len(v) == n
    return v


def parse_urdf(path):
    """Parse a URDF into a plain dict (JSON-serialisable), joints in PyBullet index order.

    PyBullet numbers joints by a depth-first walk from the root link that visits a link's
    child joints in file order; for the Panda file this equals file order.
    """
    root = ET.parse(path).getroot()
    links = {}
    for ln in root.findall("link"):
        name = ln.get("name")
        inert = ln.find("inertial")
        mass, com, rpy, I = 0.0, [0, 0, 0], [0, 0, 0], np.zeros((3, 3))
        if inert is not None:
            o = inert.find("origin")
            if o is not None:
                com = _vec(o.get("xyz", "0 0 0"))
                rpy = _vec(o.get("rpy", "0 0 0"))
            mass = float(inert.find("mass").get("value"))
            it = inert.find("inertia")
            ixx, ixy, ixz = float(it.get("ixx")), float(it.get("ixy")), float(it.get("ixz"))
            iyy, iyz, izz = float(it.get("iyy")), float(it.get("iyz")), float(it.get("izz"))
            I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
            Ri = rpy_to_matrix(*rpy)
            I = Ri @ I @ Ri.T  # express about COM in link-frame axes
        contact = {}
        ct = ln.find("contact")
        if ct is not None:
            for tag in ("lateral_friction", "stiffness", "damping", "rolling_friction", "spinning_friction"):
                e = ct.find(tag)
                if e is not None:
                    contact[tag] = float(e.get("value"))
        links[name] = dict(name=name, mass=mass, com=com, inertia=I.tolist(), contact=contact)

    joints = []
    for jn in root.findall("joint"):
        o = jn.find("origin")
        xyz = _vec(o.get("xyz", "0 0 0")) if o is not None else [0, 0, 0]
        rpy = _vec(o.get("rpy", "0 0 0")) if o is not None else [0, 0, 0]
        ax = jn.find("axis")
        axis = _vec(ax.get("xyz")) if ax is not None else [1, 0, 0]
        lim = jn.find("limit")
        joints.append(dict(
            name=jn.get("name"), type=jn.get("type"),
            parent=jn.find("parent").get("link"), child=jn.find("child").get("link"),
            xyz=xyz, rpy=rpy, axis=axis,
            lower=float(lim.get("lower", 0)) if lim is not None else 0.0,
            upper=float(lim.get("upper", 0)) if lim is not None else 0.0,
            effort=float(lim.get("effort", 0)) if lim is not None else 0.0,
            velocity=float(lim.get("velocity", 0)) if lim is not None else 0.0))

    children = {j["child"] for j in joints}
    roots = [n for n in links if n not in children]
    assert len(roots) == 1, roots
    base = roots[0]

    order = []

    def walk(link_name):
        for j in joints:
            if j["parent"] == link_name:
                order.append(j)
                walk(j["child"])
    walk(base)
    assert len(order) == len(joints)

    link_index = {base: -1}
    for idx, j in enumerate(order):
        link_index[j["child"]] = idx
    out_j = []
    for idx, j in enumerate(order):
        out_j.append(dict(j, index=idx, parent_index=link_index[j["parent"]]))
    return dict(name=root.get("name"), base=links[base], joints=out_j,
                links=[links[j["child"]] for j in order])


def descriptor_from_urdf_dict(d, base_position, home, ee_link, spheres, boxes=(), self_pairs=(), capsules=()):
    """Flatten the parsed URDF dict into a ``B2EModel``."""
    m = B2EModel()
    n = len(d["joints"])
    assert n <= MAX_LINKS
    m.n_links = n
    m.ee_link = ee_link
    ndof = 0
    for i, (j, l) in enumerate(zip(d["joints"], d["links"])):
        m.parent[i] = j["parent_index"]
        t = {"revolute": JOINT_REVOLUTE, "continuous": JOINT_REVOLUTE,
             "prismatic": JOINT_PRISMATIC, "fixed": JOINT_FIXED}[j["type"]]
        m.jtype[i] = t
        R = rpy_to_matrix(*j["rpy"])
        for k in range(3):
            m.jpos[i][k] = j["xyz"][k]
            m.com[i][k] = l["com"][k]
        for k in range(9):
            m.jrot[i][k] = R.flat[k]
            m.inertia[i][k] = np.asarray(l["inertia"]).flat[k]
        m.mass[i] = l["mass"]
        if t != JOINT_FIXED:
            a = np.asarray(j["axis"], dtype=np.float64)
            a = a / np.linalg.norm(a)
            for k in range(3):
                m.axis[i][k] = a[k]
            m.dof[i] = ndof
            m.lower[ndof] = j["lower"]
            m.upper[ndof] = j["upper"]
            # limit rows are only instantiated inside this distance of a limit; a joint would
            # have to move faster than margin/dt (24 rad/s, 2.4 m/s) for the omission to matter.
            m.limit_margin[ndof] = 0.1 if t == JOINT_REVOLUTE else 0.01
            m.max_force[ndof] = 1.0e5   # pybullet setJointMotorControl2 default force [EXT-recalled]
            m.max_vel[ndof] = -1.0
            m.joint_damping[ndof] = 0.0  # the Panda URDF has no <dynamics>
            m.home[ndof] = home[j["name"]]
            ndof += 1
        else:
            m.dof[i] = -1
    m.n_dof = ndof
    for k in range(3):
        m.base_pos[k] = base_position[k]
    for k in range(9):
        m.base_rot[k] = float(np.eye(3).flat[k])
    assert len(spheres) <= MAX_SPHERES
    m.n_spheres = len(spheres)
    names = [l["name"] for l in d["links"]]
    for s, sp in enumerate(spheres):
        li = names.index(sp["link"])
        m.sph_link[s] = li
        for k in range(3):
            m.sph_c[s][k] = sp["c"][k]
        m.sph_r[s] = sp["r"]
        ct = d["links"][li]["contact"]
        m.sph_mu[s] = ct.get("lateral_friction", 0.5)  # Bullet default link friction 0.5 [EXT-recalled]
        m.sph_erp[s] = -1.0
        m.sph_cfm[s] = 0.0
        m.sph_flags[s] = 1 if sp.get("self_only") else 0
    assert len(boxes) <= MAX_BOXES
    m.n_boxes = len(boxes)
    for b, bx in enumerate(boxes):
        li = names.index(bx["link"])
        m.box_link[b] = li
        for k in range(3):
            m.box_c[b][k] = bx["c"][k]
            m.box_h[b][k] = bx["h"][k]
        m.box_mu[b] = d["links"][li]["contact"].get("lateral_friction", 0.5)
        m.box_erp[b] = -1.0
        m.box_cfm[b] = 0.0
    assert len(self_pairs) <= MAX_SELF_PAIRS
    m.n_self_pairs = len(self_pairs)
    for k, (a, b) in enumerate(self_pairs):
        assert 0 <= a < len(spheres) and 0 <= b < len(spheres)
        m.self_a[k] = a
        m.self_b[k] = b
    assert len(capsules) <= MAX_CAPS
    m.n_caps = len(capsules)
    for c, cp in enumerate(capsules):
        li = names.index(cp["link"])
        m.cap_link[c] = li
        for k in range(3):
            m.cap_p0[c][k] = cp["p0"][k]
            m.cap_p1[c][k] = cp["p1"][k]
        m.cap_r[c] = cp["r"]
        m.cap_mu[c] = d["links"][li]["contact"].get("lateral_friction", 0.5)
    return m


PANDA_HOME = {  # reference panda_env.py:19-23
    'panda_joint1': 0.0, 'panda_joint2': -0.54, 'panda_joint3': 0.0,
    'panda_joint4': -2.6, 'panda_joint5': -0.30, 'panda_joint6': 2.0,
    'panda_joint7': 1.0, 'panda_finger_joint1': 0.02, 'panda_finger_joint2': 0.02,
}

_HERE = os.path.dirname(os.path.abspath(__file__))
PANDA_JSON = os.path.join(_HERE, "..", "robot_data", "franka_panda", "panda_model.json")


def load_panda(base_position=(0.0, 0.0, 0.625), urdf_path=None, dt=1.0 / 240.0):
    """Build the Panda ``B2EModel``.  ``urdf_path`` (e.g. the reference's own
    ``robot_data/franka_panda/panda_model.urdf``) overrides the committed JSON."""
    from .proxies import PANDA_BOXES, PANDA_CAPSULES, PANDA_SELF_PAIRS, PANDA_SPHERES
    if urdf_path is not None:
        d = parse_urdf(urdf_path)
    else:
        with open(PANDA_JSON) as f:
            d = json.load(f)
    if os.environ.get("B2ENV_CONTACT_MODEL") == "r1":   # round 1's proxies, for like-for-like measurements only
        from .proxies import PANDA_SPHERES_R1
        m = descriptor_from_urdf_dict(d, base_position, PANDA_HOME, ee_link=11, spheres=PANDA_SPHERES_R1)
    else:
        m = descriptor_from_urdf_dict(d, base_position, PANDA_HOME, ee_link=11, spheres=PANDA_SPHERES, boxes=PANDA_BOXES,
                                      self_pairs=PANDA_SELF_PAIRS, capsules=PANDA_CAPSULES)
    # soft finger contacts: <stiffness>/<damping> (URDF:256-263) -> per-contact erp/cfm the way
    # Bullet derives them: denom = dt*k + d, erp = dt*k/denom, cfm = 1/(denom*dt) [EXT-recalled]
    names = [l["name"] for l in d["links"]]
    for s in range(m.n_spheres):
        ct = d["links"][m.sph_link[s]]["contact"]
        if "stiffness" in ct:
            k = ct["stiffness"]
            dmp = ct.get("damping", 0.0) + 0.1  # + default contact damping of the other body
            denom = dt * k + dmp
            m.sph_erp[s] = dt * k / denom
            m.sph_cfm[s] = 1.0 / (denom * dt)
    for b in range(m.n_boxes):
        ct = d["links"][m.box_link[b]]["contact"]
        if "stiffness" in ct:
            k = ct["stiffness"]
            dmp = ct.get("damping", 0.0) + 0.1
            denom = dt * k + dmp
            m.box_erp[b] = dt * k / denom
            m.box_cfm[b] = 1.0 / (denom * dt)
    del names
    return m, d


def default_params(task, obs_low, obs_high, n_act=7, n_ctrl=7, use_ik=0, ik_orientation=0,
                   max_steps=1000, dist_min=None, ws_lim=None, eu_lim=None, goal_env=0):
    """World/solver/task constants.  Sources: reference panda_push_gym_env.py:39,52,122-126,
    panda_env.py:76,174-175,308; world assets from pybullet_data [EXT-recalled, SURVEY App. B.3]."""
    p = B2EParams()
    p.dt = 1.0 / 240.0
    p.gravity[0], p.gravity[1], p.gravity[2] = 0.0, 0.0, -9.8
    p.solver_iters = 150
    p.residual_tol = 1e-7   # pybullet default solverResidualThreshold [EXT-recalled]
    p.erp = 0.2             # pybullet default contact erp [EXT-recalled]
    p.slop = 1e-5
    p.warmstart = 0.85      # btContactSolverInfo default [EXT-recalled]
    p.contact_margin = 0.01
    # table/table.urdf at (0.85,0,0): top slab 1.5 x 1.0 x 0.05 centred z=0.6 (top 0.625 = world_env.py:68-69)
    p.table_min[0], p.table_min[1], p.table_min[2] = 0.85 - 0.75, -0.5, 0.575
    p.table_max[0], p.table_max[1], p.table_max[2] = 0.85 + 0.75, 0.5, 0.625
    p.table_mu = 1.0
    p.plane_mu = 1.0
    # static boxes: [0] the top slab, [1..4] the legs 0.1 x 0.1 x 0.58 at (0.85 +- 0.65, +- 0.4, 0.29) [EXT-recalled, App. B.3]
    boxes = [((0.85, 0.0, 0.6), (0.75, 0.5, 0.025))]
    boxes += [((0.85 + sx * 0.65, sy * 0.4, 0.29), (0.05, 0.05, 0.29)) for sx in (-1, 1) for sy in (-1, 1)]
    if os.environ.get("B2ENV_CONTACT_MODEL") == "r1":
        boxes = boxes[:1]           # round 1: the table top only
    p.n_sboxes = len(boxes)
    for k, (c, h) in enumerate(boxes):
        for j in range(3):
            p.sbox_c[k][j] = c[j]
            p.sbox_h[k][j] = h[j]
        p.sbox_mu[k] = 1.0
    p.ik_max_vel = -1.0
    p.kp_ik_max_vel = 0.1
    p.cube_half = 0.025
    p.cube_mass = 0.1
    p.cube_inertia = 0.1 * (0.05 ** 2) / 6.0
    p.cube_mu = 1.0
    p.damp_lin_k1, p.damp_lin_k2 = 0.1, 0.04   # btMultiBody base damping [EXT-recalled]
    p.damp_ang_k1, p.damp_ang_k2 = 0.2, 0.04
    p.task = task
    p.n_act = n_act
    p.n_ctrl = n_ctrl
    p.use_ik = use_ik
    p.goal_env = goal_env
    p.kp_grip = 0.1
    p.grasp_lift = 0.1
    p.grasp_rest_z = 0.65
    p.ik_orientation = ik_orientation
    p.ik_iters = 100
    p.ik_residual = 1e-3
    p.ik_damping = 0.1
    p.act_scale = 0.05
    p.act_scale_pos = 0.005
    p.act_scale_rot = 0.01
    p.kp_ctrl = 0.5
    p.kp_hold = 0.2
    if dist_min is None:
        dist_min = 0.1 if task == TASK_PUSH else 0.03
    p.dist_min = dist_min
    p.max_steps = max_steps
    n_obs = len(obs_low)
    assert n_obs <= MAX_OBS and len(obs_high) == n_obs
    p.n_obs = n_obs
    lo32 = np.asarray(obs_low, dtype=np.float32)
    hi32 = np.asarray(obs_high, dtype=np.float32)
    for k in range(n_obs):
        p.obs_low[k] = float(lo32[k])
        p.obs_high[k] = float(hi32[k])
    for k, (mu, sd) in enumerate(zip((0.0, 0.01, 0.0), (0.04, 0.07, 0.03))):
        p.vel_mean[k] = mu
        p.vel_std[k] = sd
    if ws_lim is None:
        ws_lim = [[0.3, 0.65], [-0.3, 0.3], [0.65, 1.5]]  # panda_env.py:37
    if eu_lim is None:
        eu_lim = [[-math.pi, math.pi]] * 3
    for k in range(3):
        p.ws_lim[k][0], p.ws_lim[k][1] = ws_lim[k]
        p.eu_lim[k][0], p.eu_lim[k][1] = eu_lim[k]
    for k, v in enumerate((0.2, 0.0, 0.8, math.pi, 0.0, 0.0)):  # panda_env.py:85-88
        p.home_hand_pose[k] = v
    return p


def panda_obs_limits(task, robot_ws, world_ws, lower, upper, eu_lim=None):
    """Per-entry (low, high) of the extended observation: robot obs (panda_env.py:141-193),
    world obs (world_env.py:109-126), relative pose and target (panda_push_gym_env.py:178-185)."""
    pi = math.pi
    if eu_lim is None:
        eu_lim = [[-pi, pi]] * 3
    lim = [list(x) for x in robot_ws] + [list(x) for x in eu_lim] + [[-1, 1]] * 3
    lim += [[lower[i], upper[i]] for i in range(len(lower))]
    lim += [list(x) for x in world_ws] + [[-pi, pi]] * 3
    lim += [[-0.5, 0.5]] * 3 + [[0, 2 * pi]] * 3
    if task in (TASK_PUSH, TASK_GRASP):
        lim += [list(x) for x in world_ws]
    low = [x[0] for x in lim]
    high = [x[1] for x in lim]
    return low, high


def panda_task_setup(task=TASK_PUSH, max_steps=1000, n_ctrl=7, use_ik=0, ik_orientation=1, goal_env=0):
    """Model + params of the registered pandaPush-v0 / pandaReach-v0 configuration
    (reference pybullet_robot_envs/__init__.py:47-68); ``use_ik=1`` gives the Cartesian control mode
    (action = hand-pose increment, 6 wide with orientation control, 3 without)."""
    m, _ = load_panda()
    h_table = 0.625
    robot_ws = [[0.3, 0.65], [-0.3, 0.3], [0.65, 1.5]]
    world_ws = [[0.3, 0.65], [-0.3, 0.3], [h_table, h_table + 0.3]]     # world_env.py:72
    robot_ws[2][0] = h_table if task == TASK_REACH else h_table - 0.2   # push :74 / reach :69
    lower = [m.lower[i] for i in range(m.n_dof)]
    upper = [m.upper[i] for i in range(m.n_dof)]
    low, high = panda_obs_limits(task, robot_ws, world_ws, lower, upper)
    n_act = n_ctrl if not use_ik else (6 if ik_orientation else 3)
    if task == TASK_GRASP:
        assert not use_ik
        n_act = n_ctrl + 1                         # + gripper open/close command
        for d in (7, 8):                           # apply_action_fingers: force=10, maxVelocity=1 (panda_env.py:218-224)
            m.max_force[d] = 10.0
            m.max_vel[d] = 1.0
    p = default_params(task, low, high, n_act=n_act, n_ctrl=n_ctrl, use_ik=use_ik, ik_orientation=ik_orientation,
                       max_steps=max_steps, ws_lim=robot_ws, goal_env=goal_env)
    return m, p


# ---------------------------------------------------------------------------------------------
# iCub (SDF).  Groundwork for the iCub envs (reference envs/icub_envs/icub_env.py): the loader and the
# model descriptor are pinned through the generic CPU oracle on SURVEY Appendix C.5; the CUDA kernel
# for the 38-link / 32-dof tree is not built yet.
def _pose6(s):
    v = [float(x) for x in s.split()]
    assert len(v) == 6
    return v


def _T(xyz, rpy):
    T = np.eye(4)
    T[:3, :3] = rpy_to_matrix(*rpy)
    T[:3, 3] = xyz
    return T


def parse_sdf(path):
    """Parse an SDF model (the reference's robot_data/iCub/icub_model.sdf) into the same dict layout as
    ``parse_urdf``.  SDF link poses are absolute in the model frame at the zero configuration, every joint
    pose is zero relative to its child (joint frame = child link frame) and axes are given in the child
    frame; PyBullet numbers joints by a depth-first walk in file order."""
    model = ET.parse(path).getroot().find("world").find("model")
    links, Tlink = {}, {}
    for ln in model.findall("link"):
        name = ln.get("name")
        pose = _pose6(ln.find("pose").text) if ln.find("pose") is not None else [0] * 6
        Tlink[name] = _T(pose[:3], pose[3:])
        inert = ln.find("inertial")
        ip = _pose6(inert.find("pose").text) if inert.find("pose") is not None else [0] * 6
        it = inert.find("inertia")
        g = lambda k: float(it.find(k).text)
        I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
        Ri = rpy_to_matrix(*ip[3:])
        links[name] = dict(name=name, mass=float(inert.find("mass").text), com=ip[:3], inertia=(Ri @ I @ Ri.T).tolist(), contact={})
    joints = []
    for jn in model.findall("joint"):
        ax = jn.find("axis")
        lim = ax.find("limit") if ax is not None else None
        dyn = ax.find("dynamics") if ax is not None else None
        joints.append(dict(
            name=jn.get("name"), type=jn.get("type"), parent=jn.find("parent").text, child=jn.find("child").text,
            axis=_vec(ax.find("xyz").text) if ax is not None else [1, 0, 0],
            lower=float(lim.find("lower").text) if lim is not None and lim.find("lower") is not None else 0.0,
            upper=float(lim.find("upper").text) if lim is not None and lim.find("upper") is not None else 0.0,
            effort=float(lim.find("effort").text) if lim is not None and lim.find("effort") is not None else 0.0,
            velocity=float(lim.find("velocity").text) if lim is not None and lim.find("velocity") is not None else 0.0,
            damping=float(dyn.find("damping").text) if dyn is not None and dyn.find("damping") is not None else 0.0))
    children = {j["child"] for j in joints}
    base = [n for n in links if n not in children]
    assert len(base) == 1, base
    base = base[0]
    order = []

    def walk(link_name):
        for j in joints:
            if j["parent"] == link_name:
                order.append(j)
                walk(j["child"])
    walk(base)
    assert len(order) == len(joints)
    index = {base: -1}
    for i, j in enumerate(order):
        index[j["child"]] = i
    out = []
    for i, j in enumerate(order):
        Trel = np.linalg.inv(Tlink[j["parent"]]) @ Tlink[j["child"]]   # joint origin in the parent link frame
        R = Trel[:3, :3]
        # roll-pitch-yaw of R (fixed axes, R = Rz Ry Rx), so that the dict stays URDF-like
        pitch = -np.arcsin(np.clip(R[2, 0], -1, 1))
        roll = np.arctan2(R[2, 1], R[2, 2])
        yaw = np.arctan2(R[1, 0], R[0, 0])
        out.append(dict(j, index=i, parent_index=index[j["parent"]], xyz=Trel[:3, 3].tolist(),
                        rpy=[float(roll), float(pitch), float(yaw)], rotation=R.tolist()))
    mp = model.find("pose")
    return dict(name=model.get("name"), base=links[base], joints=out, links=[links[j["child"]] for j in order],
                model_pose=_pose6(mp.text) if mp is not None else [0] * 6)


ICUB_HOME = {  # reference icub_env.py:19-40 (all other joints 0)
    'neck_pitch': 0.008, 'l_shoulder_pitch': -0.51, 'l_shoulder_roll': 0.7, 'l_elbow': 1.22,
    'r_shoulder_pitch': -0.51, 'r_shoulder_roll': 0.7, 'r_elbow': 1.22,
}
ICUB_JSON = os.path.join(_HERE, "..", "robot_data", "iCub", "icub_model.json")


def load_icub(sdf_path=None, pinned=False):
    """iCub ``B2EModel`` (38 joints, 32 dofs).  The base sits at the SDF model pose; ``pinned=True`` applies the
    reference's base pin quirk (anchor z x 1.2, icub_env.py:95-101): +0.126 m."""
    if sdf_path is not None:
        d = parse_sdf(sdf_path)
    else:
        with open(ICUB_JSON) as f:
            d = json.load(f)
    mp = d["model_pose"]
    base = [mp[0], mp[1], mp[2] * (1.2 if pinned else 1.0)]
    home = {j["name"]: ICUB_HOME.get(j["name"], 0.0) for j in d["joints"]}
    names = [l["name"] for l in d["links"]]
    m = descriptor_from_urdf_dict(d, base, home, ee_link=names.index("l_hand"), spheres=[])
    Rb = rpy_to_matrix(*mp[3:])
    for k in range(9):
        m.base_rot[k] = float(Rb.flat[k])
    # exact relative rotations (the rpy round trip of descriptor_from_urdf_dict is only accurate to ~1e-16)
    for i, j in enumerate(d["joints"]):
        R = np.asarray(j["rotation"])
        for k in range(9):
            m.jrot[i][k] = float(R.flat[k])
    dnum = 0
    for j in d["joints"]:
        if j["type"] != "fixed":
            m.joint_damping[dnum] = j.get("damping", 0.0)
            dnum += 1
    return m, d


# ---------------------------------------------------------------------------------------------
# iCub task setup: what icub_env.py / icub_{reach,push}_gym_env.py configure, as model + params.
ICUB_CTRL_GROUPS = {  # reference icub_env.py:42-51
    'torso': ['torso_pitch', 'torso_roll', 'torso_yaw'],
    'l_arm': ['l_shoulder_pitch', 'l_shoulder_roll', 'l_shoulder_yaw', 'l_elbow', 'l_wrist_pitch', 'l_wrist_prosup', 'l_wrist_yaw'],
    'r_arm': ['r_shoulder_pitch', 'r_shoulder_roll', 'r_shoulder_yaw', 'r_elbow', 'r_wrist_pitch', 'r_wrist_prosup', 'r_wrist_yaw'],
}
ICUB_HAND_OFFSET = {  # com_T_link_hand, reference icub_env.py:251-257
    'l': (-0.064768, -0.00563, -0.02266), 'r': (0.064668, -0.0056, -0.022681)}
ICUB_HOME_HAND_POSE = {  # reference icub_env.py:66-72
    'l': [0.3, 0.26, 0.8, 0.0, 0.0, 0.0], 'r': [0.3, -0.26, 0.8, 0.0, 0.0, math.pi]}
ICUB_EU_LIM = {
    'l': [[-math.pi / 2, math.pi / 2]] * 3,
    'r': [[-math.pi / 2, math.pi / 2], [-math.pi / 2, math.pi / 2], [math.pi / 2, 3 / 2 * math.pi]]}


def icub_spheres(d):
    """Collision proxies of the iCub arms (the collision meshes are git-LFS stubs): two spheres
    along each forearm (towards the wrist joint) and two on each hand (at the COM and towards the fingers).
    Torso, head and legs carry none — they cannot reach the table top in the pinned pose."""
    out = []
    names = [l["name"] for l in d["links"]]
    for side in ('l', 'r'):
        fore, wrist = side + "_forearm", side + "_wrist_1"
        jw = d["joints"][names.index(wrist)]          # joint forearm -> wrist_1: origin in the forearm frame
        c = np.asarray(jw["xyz"])
        out.append(dict(link=fore, c=(0.35 * c).tolist(), r=0.035))
        out.append(dict(link=fore, c=(0.80 * c).tolist(), r=0.03))
        com = np.asarray(d["links"][names.index(side + "_hand")]["com"])
        out.append(dict(link=side + "_hand", c=com.tolist(), r=0.03))
        out.append(dict(link=side + "_hand", c=(1.6 * com).tolist(), r=0.025))
    return out


def merge_fixed_links(m):
    """Fold every fixed-joint link into the body it is welded to: the result has one body per movable joint
    (``n_links == n_dof``, ``dof[i] == i``, bodies in dof order) — the layout of the warp-per-environment tree
    kernel (lane = body = dof).  Same dynamics: mass, first moment and rotational inertia of the welded links
    are summed in the host body's frame.  Returns (merged model, host body of every original link)."""
    n = m.n_links
    host = [-1] * n                     # body index carrying original link i (-1: the fixed base)
    Th = [np.eye(4) for _ in range(n)]  # pose of original link i in its host body's frame
    out = B2EModel()
    acc = []                            # per body: [mass, first moment (3), inertia about body origin (3x3)]
    nb = 0

    def add_mass(b, T, mass, com, I):
        R, t = T[:3, :3], T[:3, 3]
        c = R @ np.asarray(com, np.float64) + t
        Io = R @ np.asarray(I, np.float64).reshape(3, 3) @ R.T + mass * (np.dot(c, c) * np.eye(3) - np.outer(c, c))
        acc[b][0] += mass
        acc[b][1] += mass * c
        acc[b][2] += Io

    for i in range(n):
        p = m.parent[i]
        Tj = np.eye(4)
        Tj[:3, :3] = np.asarray(list(m.jrot[i]), np.float64).reshape(3, 3)
        Tj[:3, 3] = list(m.jpos[i])
        Tp = np.eye(4) if p < 0 else Th[p]
        hp = -1 if p < 0 else host[p]
        if m.jtype[i] == JOINT_FIXED:
            host[i] = hp
            Th[i] = Tp @ Tj
            if hp >= 0:
                add_mass(hp, Th[i], m.mass[i], list(m.com[i]), list(m.inertia[i]))
            continue
        assert m.dof[i] == nb, "dofs must be numbered in link order"
        b = nb
        nb += 1
        host[i] = b
        Th[i] = np.eye(4)
        T = Tp @ Tj                      # joint frame in the parent BODY frame
        out.parent[b] = hp
        out.jtype[b] = m.jtype[i]
        out.dof[b] = b
        for k in range(3):
            out.jpos[b][k] = T[k, 3]
            out.axis[b][k] = m.axis[i][k]
        for k in range(9):
            out.jrot[b][k] = T[:3, :3].flat[k]
        acc.append([0.0, np.zeros(3), np.zeros((3, 3))])
        add_mass(b, np.eye(4), m.mass[i], list(m.com[i]), list(m.inertia[i]))
    out.n_links = nb
    out.n_dof = m.n_dof
    assert nb == m.n_dof
    for b in range(nb):
        mass, h, Io = acc[b]
        c = h / mass if mass > 0 else np.zeros(3)
        Ic = Io - mass * (np.dot(c, c) * np.eye(3) - np.outer(c, c))
        out.mass[b] = mass
        for k in range(3):
            out.com[b][k] = c[k]
        for k in range(9):
            out.inertia[b][k] = Ic.flat[k]
    for name in ("lower", "upper", "limit_margin", "max_force", "max_vel", "joint_damping", "home"):
        for k in range(m.n_dof):
            getattr(out, name)[k] = getattr(m, name)[k]
    for k in range(3):
        out.base_pos[k] = m.base_pos[k]
    for k in range(9):
        out.base_rot[k] = m.base_rot[k]
    out.ee_link = host[m.ee_link]
    assert m.jtype[m.ee_link] != JOINT_FIXED
    out.n_spheres = m.n_spheres
    for s in range(m.n_spheres):
        li = m.sph_link[s]
        c = Th[li] @ np.array([m.sph_c[s][0], m.sph_c[s][1], m.sph_c[s][2], 1.0])
        out.sph_link[s] = host[li]
        assert host[li] >= 0
        for k in range(3):
            out.sph_c[s][k] = c[k]
        out.sph_r[s], out.sph_mu[s], out.sph_erp[s], out.sph_cfm[s] = m.sph_r[s], m.sph_mu[s], m.sph_erp[s], m.sph_cfm[s]
    return out, host


def load_icub_arm(control_arm='l', merged=True):
    """iCub model as the reference's ``iCubEnv`` sets it up (icub_env.py:85-151): base pinned by the fixed
    constraint whose anchor is the base position with z x 1.2 and whose frame carries the base orientation
    (the base ends up 0.126 m higher, at yaw -3.14 instead of +3.14) — applied here as a
    KINEMATIC pin (documented deviation: Bullet's pin is a soft 6-row constraint).  End effector = ``l_hand`` /
    ``r_hand`` (joint index 26 / 37); sphere proxies on the forearms and hands.  ``merged=True`` folds the six
    fixed F/T-sensor links into their parents (32 bodies = 32 dofs)."""
    with open(ICUB_JSON) as f:
        d = json.load(f)
    mp = d["model_pose"]
    base = [mp[0], mp[1], mp[2] * 1.2]
    home = {j["name"]: ICUB_HOME.get(j["name"], 0.0) for j in d["joints"]}
    names = [l["name"] for l in d["links"]]
    m = descriptor_from_urdf_dict(d, base, home, ee_link=names.index(control_arm + "_hand"), spheres=icub_spheres(d))
    Rb = rpy_to_matrix(mp[3], mp[4], -mp[5])
    for k in range(9):
        m.base_rot[k] = float(Rb.flat[k])
    for i, j in enumerate(d["joints"]):
        R = np.asarray(j["rotation"])
        for k in range(9):
            m.jrot[i][k] = float(R.flat[k])
    dnum = 0
    for j in d["joints"]:
        if j["type"] != "fixed":
            m.joint_damping[dnum] = j.get("damping", 0.0)
            dnum += 1
    info = dict(d=d, joint_names=[j["name"] for j in d["joints"]],
                movable=[j["name"] for j in d["joints"] if j["type"] != "fixed"])
    if merged:
        m, host = merge_fixed_links(m)
        info["host"] = host
    return m, info


def icub_ctrl_dofs(info, control_arm='l'):
    """dof indices of ``_joints_to_control`` in joint-index order (icub_env.py:121-136): torso, then the arm."""
    ctrl = set(ICUB_CTRL_GROUPS['torso'] + ICUB_CTRL_GROUPS[control_arm + '_arm'])
    return [k for k, name in enumerate(info["movable"]) if name in ctrl]


def icub_obs_limits(task, robot_ws, world_ws, eu_lim, joint_lim):
    """icub_env.py:202-249 (19 entries) + world_env.py:109-126 + icub_push_gym_env.py:166-203."""
    pi = math.pi
    lim = [list(x) for x in robot_ws] + [list(x) for x in eu_lim] + [[-1, 1]] * 3 + [list(x) for x in joint_lim]
    lim += [list(x) for x in world_ws] + [[-pi, pi]] * 3
    lim += [[-0.5, 0.5]] * 3 + [[0, 2 * pi]] * 3
    if task == TASK_PUSH:
        lim += [list(x) for x in world_ws]
    return lim


def icub_task_setup(task=TASK_PUSH, control_arm='l', use_ik=1, control_orientation=0, max_steps=1000, reward_type=0,
                    goal_env=0, merged=True):
    """Model + params of the registered iCubPush-v0 / iCubReach-v0 configurations
    (reference pybullet_robot_envs/__init__.py:7-31)."""
    m, info = load_icub_arm(control_arm, merged=merged)
    h_table = 0.625
    robot_ws = [[0.1, 0.45], [-0.3, 0.3], [h_table, 1.0]]              # icub_env.py:62, task env: z low := table height
    world_ws = [[0.1, 0.45], [-0.3, 0.3], [h_table, h_table + 0.3]]    # world_env.py:46, :72
    eu_lim = ICUB_EU_LIM[control_arm]
    ctrl = icub_ctrl_dofs(info, control_arm)
    joint_lim = [[m.lower[d], m.upper[d]] for d in ctrl]
    lim = icub_obs_limits(task, robot_ws, world_ws, eu_lim, joint_lim)
    n_act = len(ctrl) if not use_ik else (6 if control_orientation else 3)
    p = default_params(task, [x[0] for x in lim], [x[1] for x in lim], n_act=n_act, n_ctrl=len(ctrl), use_ik=use_ik,
                       ik_orientation=int(bool(control_orientation)), max_steps=max_steps, dist_min=0.03,
                       ws_lim=robot_ws, eu_lim=eu_lim, goal_env=goal_env)
    icub_params(p, task, control_arm, ctrl, control_orientation, reward_type)
    return m, p


def icub_params(p, task, control_arm, ctrl, control_orientation, reward_type):
    """The iCub-specific constants on top of ``default_params``."""
    p.n_obs_joints = len(ctrl)
    mask = 0
    for k, d in enumerate(ctrl):
        p.obs_dof[k] = d
        p.ctrl_dof[k] = d
        mask |= 1 << d
    p.ctrl_mask = mask
    for k in range(3):
        p.ik_link_offset[k] = ICUB_HAND_OFFSET[control_arm][k]
        p.vel_mean[k] = 0.0          # raw EE velocity in the observation (icub_env.py:234-237)
        p.vel_std[k] = 1.0
    for k, v in enumerate(ICUB_HOME_HAND_POSE[control_arm]):
        p.home_hand_pose[k] = v
    if control_orientation:          # icub_push_gym_env.py:234-235
        p.act_scale_pos, p.act_scale_rot = 0.01, 0.02
    else:                            # :229
        p.act_scale_pos, p.act_scale_rot = 0.005, 0.0
    p.reward_kind = REWARD_ICUB_REACH if task == TASK_REACH else (REWARD_ICUB_PUSH1 if reward_type == 1 else REWARD_ICUB_PUSH0)
    p.max_contacts = 8
    return p

"""Batched iCub robot object — same constructor, attributes and methods as the reference's ``iCubEnv``
(reference envs/icub_envs/icub_env.py:17-392), backed by the CUDA tree kernel instead of PyBullet.
Joint indices are PyBullet's (SDF joint order, 38 joints of which 32 movable); the simulation itself runs on
the 32-body model with the six welded F/T-sensor links folded into their parents (``merge_fixed_links``)."""
import math as m

import numpy as np

from pybullet_robot_envs.b2env import binding
from pybullet_robot_envs.b2env.client import B2Client, squeeze1
from pybullet_robot_envs.b2env.model import (ICUB_CTRL_GROUPS, ICUB_EU_LIM, ICUB_HAND_OFFSET, ICUB_HOME,
                                             ICUB_HOME_HAND_POSE, icub_ctrl_dofs, load_icub_arm)
from pybullet_robot_envs.envs.utils import euler_from_quaternion
from pybullet_robot_envs.gym_compat import seeding


class iCubEnv:
    joint_groups = {'l_leg': ['l_knee', 'l_ankle_pitch', 'l_hip_pitch'],      # reference icub_env.py:42-51
                    'r_leg': ['r_knee', 'r_ankle_pitch', 'r_hip_pitch'],
                    'head': ['neck_pitch', 'neck_roll', 'neck_yaw'],
                    'torso': ICUB_CTRL_GROUPS['torso'], 'l_arm': ICUB_CTRL_GROUPS['l_arm'], 'r_arm': ICUB_CTRL_GROUPS['r_arm']}

    def __init__(self, physicsClientId, use_IK=0, control_arm='l', control_orientation=1, control_eu_or_quat=0):
        if not isinstance(physicsClientId, B2Client):
            raise TypeError("physicsClientId must be a B2Client (the batched stand-in for p.connect())")
        self._physics_client_id = physicsClientId
        self._client = physicsClientId
        self._use_IK = use_IK
        self._control_orientation = control_orientation
        self._control_eu_or_quat = control_eu_or_quat
        self._control_arm = control_arm if control_arm in ('r', 'l') else 'l'
        self._workspace_lim = [[0.1, 0.45], [-0.3, 0.3], [0.5, 1.0]]
        self._eu_lim = [list(x) for x in ICUB_EU_LIM[self._control_arm]]
        self._home_hand_pose = list(ICUB_HOME_HAND_POSE[self._control_arm])
        self.robot_id = 0
        self.model, self._info = load_icub_arm(self._control_arm, merged=True)
        d = self._info["d"]
        self.initial_positions = {j["name"]: ICUB_HOME.get(j["name"], 0.0) for j in d["joints"] if j["type"] != "fixed"}
        # joint name -> PyBullet joint index, movable joints only, in joint-index order (reference :104-117)
        self._joint_name_to_ids = {j["name"]: j["index"] for j in d["joints"] if j["type"] != "fixed"}
        ctrl_names = set(self.joint_groups['torso'] + self.joint_groups[self._control_arm + '_arm'])
        self._joints_to_control = [i for n, i in self._joint_name_to_ids.items() if n in ctrl_names]
        self._joints_to_block = [i for n, i in self._joint_name_to_ids.items() if n not in ctrl_names]
        self.end_eff_idx = self._joint_name_to_ids[self._control_arm + '_wrist_yaw']
        self._ctrl_dofs = icub_ctrl_dofs(self._info, self._control_arm)   # dof index of every controlled joint
        self.ll, self.ul, self.jr, self.rs, self.jd = self.get_joint_ranges()
        self.seed()

    # ------------------------------------------------------------------ lifecycle
    def reset(self, env_ids=None):
        """Home joint state, position motors (gain 0.2) on every joint; IK mode: home hand pose -> IK -> motor
        targets; then one physics step (reference :85-151)."""
        c = self._client
        nd = self.model.n_dof
        ids = None if env_ids is None else np.asarray(env_ids, np.int32)
        n = c.num_envs if ids is None else len(ids)
        home = np.tile(np.array([self.model.home[i] for i in range(nd)], np.float32), (n, 1))
        put = (lambda k, v: c.set(k, v)) if ids is None else (lambda k, v: c.set_rows(k, ids, v))
        put("q", home)
        put("qd", np.zeros((n, nd), np.float32))
        put("mtarget", home)
        self.ll, self.ul, self.jr, self.rs, self.jd = self.get_joint_ranges()
        mode = binding.MODE_HOLD
        if self._use_IK:
            put("hand_pose", np.tile(np.array(self._home_hand_pose, np.float32), (n, 1)))
            mode = binding.MODE_IK_POSE
        if ids is None:
            c.step_simulation(1, mode)
        else:
            c.step_subset(ids, 1, mode)

    def delete_simulated_robot(self):
        pass

    # ------------------------------------------------------------------ static info
    def get_joint_ranges(self):
        lower, upper, ranges, rest, damp = [], [], [], [], []
        movable = self._info["movable"]
        for name, idx in self._joint_name_to_ids.items():
            d = movable.index(name)
            lo, hi = self.model.lower[d], self.model.upper[d]
            lower.append(lo)
            upper.append(hi)
            ranges.append(hi - lo)
            rest.append(self.initial_positions[name])
            damp.append(0.1 if idx in self._joints_to_control else 100.)
        return lower, upper, ranges, rest, damp

    def get_workspace(self):
        return [i[:] for i in self._workspace_lim]

    def set_workspace(self, ws):
        self._workspace_lim = [i[:] for i in ws]

    def get_rotation_lim(self):
        return [i[:] for i in self._eu_lim]

    def set_rotation_lim(self, eu):
        self._eu_lim = [i[:] for i in eu]

    def get_action_dim(self):
        if not self._use_IK:
            return len(self._joints_to_control)
        if self._control_orientation and self._control_eu_or_quat == 0:
            return 6
        if self._control_orientation and self._control_eu_or_quat == 1:
            return 7
        return 3

    def get_observation_dim(self):
        return 19

    def observation_limits(self):
        lim = [list(x) for x in self._workspace_lim]
        lim += [list(x) for x in self._eu_lim] if self._control_eu_or_quat == 0 else [[-1, 1]] * 4
        lim += [[-1, 1]] * 3
        lim += [[self.ll[i], self.ul[i]] for i, idx in enumerate(self._joint_name_to_ids.values())
                if idx in self._joints_to_control]
        return lim

    # ------------------------------------------------------------------ state queries
    def get_observation(self):
        """Hand COM pose, raw hand linear velocity, positions of the controlled joints (reference :202-249)
        -> (obs [B,19], limits)."""
        raw = self._client.observe()[3]
        return squeeze1(raw[:, :19].astype(np.float64), self._client.num_envs), self.observation_limits()

    def _com_to_link_hand_frame(self):
        return (ICUB_HAND_OFFSET[self._control_arm], (0., 0., 0., 1.))

    # ------------------------------------------------------------------ control
    def apply_action(self, action, max_vel=-1):
        B = self._client.num_envs
        action = np.asarray(action, np.float32)
        if action.ndim == 1:
            action = np.tile(action, (B, 1))
        if self._use_IK:
            if action.shape[1] not in (3, 6, 7):
                raise AssertionError('number of action commands must be \n- 3: (dx,dy,dz)'
                                     '\n- 6: (dx,dy,dz,droll,dpitch,dyaw)'
                                     '\n- 7: (dx,dy,dz,qx,qy,qz,w)'
                                     '\ninstead it is: ', action.shape[1])
            # the pose is stored; workspace / rotation clamps, COM -> link frame, IK, blocked joints and the 32
            # position targets (reference :274-326) happen inside the next physics launch (MODE_IK_POSE)
            hp = self._client.get("hand_pose")
            ws = self._workspace_lim
            for k in range(3):
                hp[:, k] = np.clip(action[:, k], ws[k][0], ws[k][1])
            if action.shape[1] == 6 and self._control_orientation:
                for k in range(3):
                    hp[:, 3 + k] = np.clip(action[:, 3 + k], self._eu_lim[k][0], self._eu_lim[k][1])
            elif action.shape[1] == 7 and self._control_orientation:
                # quaternion command: kept as Euler angles on the device, converted back there (same rotation up to rounding)
                hp[:, 3:6] = euler_from_quaternion(action[:, 3:7])
            else:
                hp[:, 3:6] = np.asarray(self._home_hand_pose[3:6], np.float32)
            self._client.set("hand_pose", hp)
            # max_vel != -1 (reference :331-337): every joint keeps the gain 0.2, maxVelocity = max_vel, until the next call
            prm = self._client.params
            want = float(max_vel) if max_vel != -1 else -1.0
            if prm is not None and prm.ik_max_vel != want:
                prm.ik_max_vel = want
                prm.kp_ik_max_vel = prm.kp_hold
                if self._client.sim is not None:
                    self._client.sim.set_params(prm)
            self._client.pending_mode = binding.MODE_IK_POSE
            return
        if action.shape[1] != len(self._joints_to_control):
            raise AssertionError('number of motor commands differs from number of motor to control',
                                 action.shape[1], len(self._joints_to_control))
        mt = self._client.get("mtarget")
        for k, d in enumerate(self._ctrl_dofs):
            mt[:, d] = np.clip(action[:, k], self.model.lower[d], self.model.upper[d])
        self._client.set("mtarget", mt)
        self._client.pending_mode = binding.MODE_TARGETS

    def seed(self, seed=None):
        self.np_random, seed = seeding.np_random(seed)
        return [seed]

    def debug_gui(self):
        pass

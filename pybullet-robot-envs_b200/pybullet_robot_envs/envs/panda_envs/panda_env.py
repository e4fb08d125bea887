"""Batched Panda robot object — same constructor, attributes and methods as the reference's
``pandaEnv`` (reference envs/panda_envs/panda_env.py:17-395), backed by the CUDA simulation
instead of PyBullet.  Every method acts on all ``num_envs`` environments at once; with
``num_envs == 1`` the returned shapes equal the reference's."""
import math as m

import numpy as np

from pybullet_robot_envs.b2env import binding
from pybullet_robot_envs.b2env.client import B2Client, squeeze1
from pybullet_robot_envs.b2env.model import PANDA_HOME, load_panda
from pybullet_robot_envs.b2env.proxies import KEY_BOX_CUBE, KEY_CAP_CUBE, KEY_SPHERE_CUBE
from pybullet_robot_envs.envs.utils import euler_from_quaternion, quaternion_from_euler
from pybullet_robot_envs.gym_compat import seeding


class pandaEnv:
    initial_positions = dict(PANDA_HOME)  # reference panda_env.py:19-23

    def __init__(self, physicsClientId, use_IK=0, base_position=(0.0, 0, 0.625), control_orientation=1,
                 control_eu_or_quat=0, joint_action_space=9, includeVelObs=True):
        if not isinstance(physicsClientId, B2Client):
            raise TypeError("physicsClientId must be a B2Client (the batched stand-in for p.connect())")
        self._physics_client_id = physicsClientId
        self._client = physicsClientId
        self._use_IK = use_IK
        self._control_orientation = control_orientation
        self._base_position = base_position
        self.joint_action_space = joint_action_space
        self._include_vel_obs = includeVelObs
        self._control_eu_or_quat = control_eu_or_quat
        self._workspace_lim = [[0.3, 0.65], [-0.3, 0.3], [0.65, 1.5]]
        self._eu_lim = [[-m.pi, m.pi], [-m.pi, m.pi], [-m.pi, m.pi]]
        self.end_eff_idx = 11
        self._home_hand_pose = []
        self._num_dof = 7
        self.robot_id = 0
        self.model, self._urdf = load_panda(base_position)
        # joint name -> PyBullet joint index, movable joints only, in joint-index order
        self._joint_name_to_ids = {j["name"]: j["index"] for j in self._urdf["joints"]
                                   if j["type"] in ("revolute", "prismatic")}
        self.ll, self.ul, self.jr, self.rs = self.get_joint_ranges()
        self.seed()
        if self._use_IK:
            self._home_hand_pose = [0.2, 0.0, 0.8, min(m.pi, max(-m.pi, m.pi)), 0.0, 0.0]

    # ------------------------------------------------------------------ lifecycle
    def reset(self, env_ids=None):
        """Home joint state, zero velocity, position motors targeting home (reference :51-91).
        ``env_ids``: only those environments (per-env reset of a batch)."""
        c = self._client
        nd = self.model.n_dof
        ids = None if env_ids is None else np.asarray(env_ids, np.int32)
        n = c.num_envs if ids is None else len(ids)
        home = np.tile(np.array([self.model.home[i] for i in range(nd)], np.float32), (n, 1))
        put = (lambda k, v: c.set(k, v)) if ids is None else (lambda k, v: c.set_rows(k, ids, v))
        put("q", home)
        put("qd", np.zeros((n, nd), np.float32))
        put("mtarget", home)
        self.ll, self.ul, self.jr, self.rs = self.get_joint_ranges()
        if self._use_IK:
            # home hand pose -> IK -> motor targets, then one physics step (reference :83-91)
            self._home_hand_pose = [0.2, 0.0, 0.8, min(m.pi, max(-m.pi, m.pi)), 0.0, 0.0]
            put("hand_pose", np.tile(np.array(self._home_hand_pose, np.float32), (n, 1)))
            if ids is None:
                c.step_simulation(1, binding.MODE_IK_POSE)
            else:
                c.step_subset(ids, 1, binding.MODE_IK_POSE)

    def delete_simulated_robot(self):
        pass  # bodies are fixed members of the batched simulation

    # ------------------------------------------------------------------ static info
    def get_joint_ranges(self):
        lower, upper, ranges, rest = [], [], [], []
        for name in self._joint_name_to_ids:
            d = self.model.dof[self._joint_name_to_ids[name]]
            lo, hi = self.model.lower[d], self.model.upper[d]
            lower.append(lo)
            upper.append(hi)
            ranges.append(hi - lo)
            rest.append(self.initial_positions[name])
        return lower, upper, ranges, rest

    def get_action_dim(self):
        if not self._use_IK:
            return self.joint_action_space
        if self._control_orientation and self._control_eu_or_quat == 0:
            return 6
        if self._control_orientation and self._control_eu_or_quat == 1:
            return 7
        return 3

    def get_observation_dim(self):
        return 18 + (1 if self._control_eu_or_quat == 1 else 0) - (0 if self._include_vel_obs else 3)

    def get_workspace(self):
        return [i[:] for i in self._workspace_lim]

    def set_workspace(self, ws):
        self._workspace_lim = [i[:] for i in ws]

    def get_rotation_lim(self):
        return [i[:] for i in self._eu_lim]

    def set_rotation_lim(self, eu):
        self._eu_lim = [i[:] for i in eu]

    def observation_limits(self):
        lim = [list(x) for x in self._workspace_lim]
        lim += [list(x) for x in self._eu_lim] if self._control_eu_or_quat == 0 else [[-1, 1]] * 4
        if self._include_vel_obs:
            lim += [[-1, 1]] * 3
        lim += [[self.ll[i], self.ul[i]] for i in range(len(self._joint_name_to_ids))]
        return lim

    # ------------------------------------------------------------------ state queries
    def get_observation(self):
        """EE pose (link 11 COM), standardised EE linear velocity, joint positions
        (reference :141-193) -> (obs [B,18], limits)."""
        raw = self._client.observe()[3][:, :18].astype(np.float64)
        if self._control_eu_or_quat == 1:   # hand orientation as a quaternion (reference :165-167): 19 entries
            raw = np.concatenate([raw[:, :3], quaternion_from_euler(raw[:, 3:6]), raw[:, 6:]], axis=1)
        if not self._include_vel_obs:
            k = 7 if self._control_eu_or_quat == 1 else 6
            raw = np.concatenate([raw[:, :k], raw[:, k + 3:]], axis=1)
        return squeeze1(raw, self._client.num_envs), self.observation_limits()

    # ------------------------------------------------------------------ control
    def pre_grasp(self):
        self.apply_action_fingers([0.04, 0.04])

    def grasp(self, obj_id=None):
        self.apply_action_fingers([0.0, 0.0], obj_id)

    def apply_action_fingers(self, action, obj_id=None):
        assert len(action) == 2, ('finger joints are 2! The number of actions you passed is ', len(action))
        mt = self._client.get("mtarget")
        mt[:, 7] = action[0]
        mt[:, 8] = action[1]
        self._client.set("mtarget", mt)

    def apply_action(self, action, max_vel=-1):
        """Joint mode: clamp to the joint limits and make them the position-motor targets
        (reference :293-310).  The physics step that follows must run with control gains
        (``B2Client.step_simulation(mode=MODE_TARGETS)``)."""
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
            # Cartesian command: the pose is stored; IK + motor targets happen inside the next physics
            # launch (B2Client.step_simulation picks MODE_IK_POSE), like calculateInverseKinematics +
            # setJointMotorControlArray(kp 0.2) at reference :269-282
            hp = self._client.get("hand_pose")
            hp[:, :3] = action[:, :3]
            if not self._control_orientation or action.shape[1] == 3:
                # orientation not under control: the home orientation (reference :247-248).  [A 3-wide command with
                # control_orientation=1 keeps the CURRENT orientation in the reference (:264-266); here it keeps the
                # commanded one, which the task envs never let drift]
                hp[:, 3:6] = np.asarray(self._home_hand_pose[3:6], np.float32)
            elif action.shape[1] == 6:
                hp[:, 3:6] = np.clip(action[:, 3:6], -m.pi, m.pi)                      # reference :251-257
            else:
                # quaternion command (reference :260-261): the device keeps the commanded pose as Euler angles and converts
                # back with getQuaternionFromEuler, the same rotation up to rounding
                hp[:, 3:6] = euler_from_quaternion(action[:, 3:7])
            self._client.set("hand_pose", hp)
            self._set_ik_max_vel(max_vel)
            self._client.pending_mode = binding.MODE_IK_POSE
            return
        assert action.shape[1] == self.joint_action_space, \
            ('number of motor commands differs from number of motor to control', action.shape[1])
        n = action.shape[1]
        mt = self._client.get("mtarget")
        mt[:, :n] = np.minimum(np.asarray(self.ul[:n], np.float32), np.maximum(np.asarray(self.ll[:n], np.float32), action))
        self._client.set("mtarget", mt)
        self._client.pending_mode = binding.MODE_TARGETS

    def _set_ik_max_vel(self, max_vel):
        """``max_vel != -1`` (reference :285-291): the 7 arm joints run through setJointMotorControl2(maxVelocity=max_vel)
        with PyBullet's default position gain; it stays in force until the next apply_action, like a motor setting."""
        prm = self._client.params
        want = float(max_vel) if max_vel != -1 else -1.0
        if prm is not None and abs(prm.ik_max_vel - want) > 0:
            prm.ik_max_vel = want
            prm.kp_ik_max_vel = 0.1
            if self._client.sim is not None:
                self._client.sim.set_params(prm)

    def check_collision(self, obj_id=None):
        """True where the robot touches the object somewhere else than with the finger pads (sphere and capsule proxies
        of link4 .. hand; the pads are the box proxies)."""
        keys = self._client.get("cache_key")
        body = ((keys >= KEY_SPHERE_CUBE) & (keys < KEY_SPHERE_CUBE + 16)) | ((keys >= KEY_CAP_CUBE) & (keys < KEY_CAP_CUBE + 4))
        return squeeze1(body.any(axis=1), self._client.num_envs)

    def check_contact_fingertips(self, obj_id=None):
        keys = self._client.get("cache_key")
        lam = self._client.get("cache_lam").reshape(keys.shape[0], keys.shape[1], 3)
        dt = self._client.params.dt
        left = (keys >= KEY_BOX_CUBE) & (keys < KEY_BOX_CUBE + 1024)               # pad 0 = panda_leftfinger
        right = (keys >= KEY_BOX_CUBE + 1024) & (keys < KEY_BOX_CUBE + 2048)       # pad 1 = panda_rightfinger
        n = left.any(axis=1).astype(int) + right.any(axis=1).astype(int)
        f0 = np.where(left, lam[:, :, 0] / dt, 0).sum(axis=1) / np.maximum(left.sum(axis=1), 1)
        f1 = np.where(right, lam[:, :, 0] / dt, 0).sum(axis=1) / np.maximum(right.sum(axis=1), 1)
        B = self._client.num_envs
        return squeeze1(n, B), (squeeze1(f0, B), squeeze1(f1, B))

    def seed(self, seed=None):
        self.np_random, seed = seeding.np_random(seed)
        return [seed]

    def debug_gui(self):
        pass  # no GUI in the batched backend

"""Batched world object: table + one object per environment — same surface as the reference's
``WorldEnv`` (reference envs/world_envs/world_env.py:33-176).  Only ``cube_small`` has a collision
model in the CUDA backend (mesh objects are out of scope; their assets live in
the un-vendored pybullet_data package)."""
import math as m

import numpy as np

from pybullet_robot_envs.b2env.client import B2Client, squeeze1
from pybullet_robot_envs.gym_compat import seeding


def get_objects_list():
    return ['duck_vhacd', 'cube_small', 'teddy_vhacd', 'domino/domino']


def _quat_from_yaw(yaw):
    return (0.0, 0.0, m.sin(0.5 * yaw), m.cos(0.5 * yaw))


class WorldEnv:
    _warned_proxy = False

    def __init__(self, physicsClientId, obj_name='duck_vhacd', obj_pose_rnd_std=0.05, workspace_lim=None,
                 control_eu_or_quat=0):
        if not isinstance(physicsClientId, B2Client):
            raise TypeError("physicsClientId must be a B2Client")
        if workspace_lim is None:
            workspace_lim = [[0.25, 0.52], [-0.3, 0.3], [0.5, 1.0]]
        self._physics_client_id = physicsClientId
        self._client = physicsClientId
        self._ws_lim = tuple(list(x) for x in workspace_lim)
        self._h_table = 0.6 + 0.05 / 2   # table top of table/table.urdf at z = 0 (world_env.py:68-69)
        self._obj_name = obj_name
        self._obj_pose_rnd_std = obj_pose_rnd_std
        self._obj_init_pose = []
        self._control_eu_or_quat = control_eu_or_quat
        self.obj_id = 1
        self.table_id = 2
        self._ws_lim[2][:] = [self._h_table, self._h_table + 0.3]
        self._rngs = None
        self.seed()

    # per-env generators: env i draws from np_random(seed + i) so a batch reproduces B
    # independent reference envs seeded seed, seed+1, ...
    def seed(self, seed=None):
        self.np_random, seed0 = seeding.np_random(seed)
        B = self._client.num_envs
        if B == 1:
            self._rngs = [self.np_random]
        elif seed is None:
            self._rngs = [seeding.np_random(None)[0] for _ in range(B)]
        else:
            self._rngs = [self.np_random] + [seeding.np_random(seed + i)[0] for i in range(1, B)]
        return [seed0]

    def reset(self, env_ids=None):
        """(Re)place the object (reference :61-84): sample a start pose per env, zero velocity,
        forget the contact cache.  ``env_ids``: only those environments."""
        self._ws_lim[2][:] = [self._h_table, self._h_table + 0.3]
        self.load_object(self._obj_name, env_ids)

    def load_object(self, obj_name, env_ids=None):
        if obj_name not in get_objects_list():
            raise NotImplementedError("unknown object '%s'" % obj_name)
        if obj_name != 'cube_small' and not self._warned_proxy:
            # the mesh objects (duck_vhacd is the reference default of iCubReachGymEnv) live in the un-vendored
            # pybullet_data package: they are simulated with the cube_small collision model (documented deviation)
            import warnings
            warnings.warn("object '%s' is simulated with the 'cube_small' collision proxy (mesh assets are not available)" % obj_name)
            self._warned_proxy = True   # once per WorldEnv object (every construction warns)
        self._obj_name = obj_name
        c = self._client
        B = c.num_envs
        if env_ids is None:
            poses = np.array([self._sample_pose(i) for i in range(B)], np.float32)
            self._obj_init_pose = poses
            c.set("obj_pose", poses)
            c.set("obj_vel", np.zeros((B, 6), np.float32))
            c.set("cache_key", np.full((B, 16), -1, np.int32))
            c.set("cache_lam", np.zeros((B, 48), np.float32))
            return
        ids = np.asarray(env_ids, np.int32)
        poses = np.array([self._sample_pose(int(i)) for i in ids], np.float32).reshape(len(ids), 7)
        if len(np.shape(self._obj_init_pose)) == 2:
            self._obj_init_pose[ids] = poses
        c.set_rows("obj_pose", ids, poses)
        c.set_rows("obj_vel", ids, np.zeros((len(ids), 6), np.float32))
        c.set_rows("cache_key", ids, np.full((len(ids), 16), -1, np.int32))
        c.set_rows("cache_lam", ids, np.zeros((len(ids), 48), np.float32))

    def get_object_init_pose(self):
        p = np.asarray(self._obj_init_pose)
        return squeeze1(p[:, :3], self._client.num_envs), squeeze1(p[:, 3:7], self._client.num_envs)

    def set_obj_pose(self, new_pos, new_quat):
        B = self._client.num_envs
        pose = np.concatenate([np.broadcast_to(np.asarray(new_pos, np.float32), (B, 3)),
                               np.broadcast_to(np.asarray(new_quat, np.float32), (B, 4))], axis=1)
        self._client.set("obj_pose", pose)

    def get_table_height(self):
        return self._h_table

    def get_object_shape_info(self):
        a = 2 * self._client.ensure().params.cube_half
        return [self.obj_id, -1, 3, (a, a, a), '', (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0)]

    def get_workspace(self):
        return [i[:] for i in self._ws_lim]

    def get_observation_dimension(self):
        return 6

    def observation_limits(self):
        lim = [list(x) for x in self._ws_lim]
        lim += [[-m.pi, m.pi]] * 3 if self._control_eu_or_quat == 0 else [[-1, 1]] * 4
        return lim

    def get_observation(self):
        """Object base position + Euler angles (reference :109-126) -> (obs [B,6], limits)."""
        raw = self._client.observe()[3]
        nr = getattr(self._client, 'n_robot_obs', 18)   # width of the robot observation that precedes it
        return squeeze1(raw[:, nr:nr + 6].astype(np.float64), self._client.num_envs), self.observation_limits()

    def check_contact(self, body_id=None, obj_id=None):
        keys = self._client.get("cache_key")
        hit = ((keys >= 16) & (keys < 32)).any(axis=1)
        return squeeze1(hit, self._client.num_envs)

    def debug_gui(self):
        pass

    def _sample_pose(self, env_index=0):
        """Start pose of the object (reference :145-176): centre of the reduced workspace plus a
        uniform offset of half-width ``obj_pose_rnd_std`` (the reference's "std" is a half-width),
        yaw uniform in +-pi/4, z = table + 0.07.  Draw order x, y, yaw."""
        rng = self._rngs[env_index]
        x_min, x_max = self._ws_lim[0][0] + 0.05, self._ws_lim[0][1] - 0.1
        y_min, y_max = self._ws_lim[1][0] + 0.05, self._ws_lim[1][1] - 0.05
        px = x_min + 0.5 * (x_max - x_min)
        py = y_min + 0.5 * (y_max - y_min)
        pz = self._h_table + 0.07
        quat = _quat_from_yaw(m.pi / 4)
        if self._obj_pose_rnd_std > 0:
            s = self._obj_pose_rnd_std
            px += rng.uniform(low=-s, high=s)
            py += rng.uniform(low=-s, high=s)
            quat = _quat_from_yaw(rng.uniform(low=-m.pi / 4, high=m.pi / 4))
        px = float(np.clip(px, x_min, x_max))
        py = float(np.clip(py, y_min, y_max))
        return (px, py, pz) + quat

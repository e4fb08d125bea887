"""``pandaPushGymGoalEnv`` — the HER / GoalEnv variant of the push task (reference
envs/panda_envs/panda_push_gym_goal_env.py:19-122): dict observation, sparse reward, ``is_success``."""
import numpy as np

from pybullet_robot_envs import gym_compat as gym
from pybullet_robot_envs.gym_compat import spaces
from pybullet_robot_envs.b2env import binding
from pybullet_robot_envs.b2env.client import squeeze1
from pybullet_robot_envs.envs.utils import goal_distance, scale_gym_data
from pybullet_robot_envs.envs.world_envs.world_env import get_objects_list
from pybullet_robot_envs.envs.panda_envs.panda_push_gym_env import pandaPushGymEnv


class GoalMixin:
    """GoalEnv surface shared by ``pandaPushGymGoalEnv`` and ``iCubPushGymGoalEnv`` (reference
    panda_push_gym_goal_env.py:37-122, icub_push_gym_goal_env.py:40-132 — the two files are the same code)."""
    _box_cls = None   # the task class whose Box spaces become the 'observation' entry

    def create_gym_spaces(self):
        box, action_space = self._box_cls.create_gym_spaces(self)
        observation_space = spaces.Dict(dict(
            desired_goal=spaces.Box(-10, 10, shape=(3,), dtype='float32'),
            achieved_goal=spaces.Box(-10, 10, shape=(3,), dtype='float32'),
            observation=box))
        return observation_space, action_space

    def _goal_dict(self, scaled, raw):
        B = self.num_envs
        nr = self._n_robot_obs   # robot obs | object pose (6) | relative pose (6) | target (3)
        return {'observation': squeeze1(np.asarray(scaled, np.float64), B),
                'achieved_goal': squeeze1(np.asarray(raw[:, nr:nr + 3], np.float64), B),
                'desired_goal': squeeze1(np.asarray(raw[:, nr + 12:nr + 15], np.float64), B)}

    def get_goal_observation(self):
        scaled, _, _, raw = self._physics_client_id.observe()
        d = self._goal_dict(scaled, raw)
        d['observation'] = squeeze1(raw.astype(np.float64), self.num_envs)   # unscaled, like the reference
        return d

    def reset(self):
        gym.GoalEnv.reset(self)
        self.reset_simulation()
        world_obs, _ = self._world.get_observation()
        self._target_pose = self.sample_tg_pose(np.asarray(world_obs).reshape(self.num_envs, 6)[:, :3])
        self._sync_target()
        self._after_target()
        scaled, _, _, raw = self._physics_client_id.observe()
        return self._goal_dict(scaled, raw)

    def step(self, action):
        a = np.asarray(action, np.float32)
        assert a.shape[-1:] == self.action_space.shape
        obs, rew, done = self._sim.step_host(self._as_batch(a), self._action_repeat, binding.MODE_ACTION)
        raw = self._sim.get("raw_obs")
        self._physics_client_id.invalidate()
        out = self._goal_dict(obs, raw)
        success = self._is_success(out['achieved_goal'], out['desired_goal'])
        info = {'is_success': success}
        B = self.num_envs
        return out, squeeze1(rew, B), squeeze1(done.astype(bool), B), info

    def _termination(self):
        c = self._physics_client_id.get("counters")[:, 0]
        return squeeze1((c > self._max_steps).astype(np.float32), self.num_envs)

    def _is_success(self, achieved_goal, goal):
        d = goal_distance(np.asarray(achieved_goal)[..., :3], np.asarray(goal)[..., :3])
        return d <= self._target_dist_min

    def compute_reward(self, achieved_goal, goal, info):
        """Vectorisable sparse reward (HER relabels whole batches through this, reference :118-122)."""
        d = goal_distance(np.asarray(achieved_goal)[..., :3], np.asarray(goal)[..., :3])
        return -(d > self._target_dist_min).astype(np.float32)


class pandaPushGymGoalEnv(GoalMixin, gym.GoalEnv, pandaPushGymEnv):
    _goal_env = 1   # the kernel then applies the GoalEnv termination / reward rules (reference :96-122)
    _box_cls = pandaPushGymEnv

    def __init__(self, numControlledJoints=7, use_IK=0, action_repeat=1, obj_name=get_objects_list()[1],
                 renders=False, max_steps=1000, obj_pose_rnd_std=0, tg_pose_rnd_std=0.2, includeVelObs=True,
                 num_envs=1, device=0):
        pandaPushGymEnv.__init__(self, numControlledJoints, use_IK, action_repeat, obj_name, renders, max_steps,
                                 obj_pose_rnd_std, tg_pose_rnd_std, includeVelObs, num_envs=num_envs, device=device)

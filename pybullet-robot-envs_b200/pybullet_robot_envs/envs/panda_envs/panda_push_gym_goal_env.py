"""``pandaPushGymGoalEnv`` — the HER / GoalEnv variant of the push task (reference
envs/panda_envs/panda_push_gym_goal_env.py:19-122): dict observation, sparse reward, ``is_success``."""
import numpy as np

from pybullet_robot_envs import gym_compat as gym
from pybullet_robot_envs.gym_compat import spaces
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

    def _goal_dict(self, scaled, raw, squeeze=True):
        B = self.num_envs if squeeze else 0
        nr = self._n_robot_obs   # robot obs | object pose (6) | relative pose (6) | target (3)
        scaled = np.asarray(scaled, np.float64).reshape(-1, raw.shape[-1])
        return {'observation': squeeze1(scaled, B),
                'achieved_goal': squeeze1(np.asarray(raw[:, nr:nr + 3], np.float64), B),
                'desired_goal': squeeze1(np.asarray(raw[:, nr + 12:nr + 15], np.float64), B)}

    def get_goal_observation(self):
        scaled, _, _, raw = self._physics_client_id.observe()
        d = self._goal_dict(scaled, raw)
        d['observation'] = squeeze1(raw.astype(np.float64), self.num_envs)   # unscaled, like the reference
        return d

    def reset(self, env_ids=None):
        """``reset()`` as the reference (:74-87); ``reset(env_ids)`` resets the listed environments only and returns
        their dict observation (batched extension, same contract as the Box envs)."""
        gym.GoalEnv.reset(self)
        if env_ids is not None:
            ids = np.asarray(env_ids, np.int32)
            scaled = self._box_cls.reset(self, ids)
            raw = self._physics_client_id.observe()[3][ids]
            return self._goal_dict(scaled, raw, squeeze=False)
        self._box_cls.reset(self)
        scaled, _, _, raw = self._physics_client_id.observe()
        return self._goal_dict(scaled, raw)

    def step(self, action):
        """One fused launch like the Box envs (same host / CUDA-tensor / auto_reset paths), results re-packed as the
        GoalEnv dict (reference :89-104)."""
        nr = self._n_robot_obs
        if hasattr(action, "is_cuda") and action.is_cuda:
            obs, rew, done, _ = self._step_fused(action)
            raw = self._sim.get_device("raw_obs", obs)
            out = {'observation': obs, 'achieved_goal': raw[:, nr:nr + 3], 'desired_goal': raw[:, nr + 12:nr + 15]}
            d = (out['achieved_goal'] - out['desired_goal']).norm(dim=1)
            return out, rew, done > 0, {'is_success': d <= self._target_dist_min}
        obs, rew, done, _ = self._step_fused(action)
        raw = self._sim.get("raw_obs")      # after an auto-reset: rows of the restarted envs describe the new episode, like obs
        out = self._goal_dict(obs, raw)
        success = self._is_success(out['achieved_goal'], out['desired_goal'])
        return out, rew, np.asarray(done).astype(bool), {'is_success': success}

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

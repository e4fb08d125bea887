"""Shared machinery of the Panda task envs (push / reach): the batched, fused-kernel
implementation of what reference panda_push_gym_env.py / panda_reach_gym_env.py spell out
twice.  The class layout (robot / world / task split and the hook names) is the reference's;
``step()`` runs ONE CUDA launch that does apply_action + stepSimulation + observation +
termination + reward for every environment."""
import math as m

import numpy as np

from pybullet_robot_envs import gym_compat as gym
from pybullet_robot_envs.gym_compat import spaces, seeding
from pybullet_robot_envs.b2env import binding
from pybullet_robot_envs.b2env.client import B2Client, squeeze1
from pybullet_robot_envs.b2env.model import TASK_GRASP, TASK_PUSH, TASK_REACH, default_params
from pybullet_robot_envs.envs.panda_envs.panda_env import pandaEnv
from pybullet_robot_envs.envs.world_envs.world_env import WorldEnv
from pybullet_robot_envs.envs.utils import goal_distance, scale_gym_data

_PARK_POSE = (5.0, 5.0, 0.025, 0.0, 0.0, 0.0, 1.0)  # where the object waits while "not loaded"


class PandaTaskBase(gym.Env):
    metadata = {'render.modes': ['human', 'rgb_array'], 'video.frames_per_second': 50}
    _task = TASK_PUSH
    _n_robot_obs = 18   # entries of the robot observation (reference panda_env.py:141-193; iCub: 19)
    _hooks = ("apply_action", "get_extended_observation", "_termination", "_compute_reward")

    def _setup(self, numControlledJoints, use_IK, action_repeat, obj_name, renders, max_steps, obj_pose_rnd_std,
               includeVelObs, num_envs, device, target_dist_min, z_low_offset):
        self._timeStep = 1. / 240.
        self.action_dim = []
        self._use_IK = use_IK
        self._action_repeat = action_repeat
        self._observation = []
        self._env_step_counter = 0
        self._renders = renders
        self._max_steps = max_steps
        self.terminated = 0
        self._target_dist_min = target_dist_min
        self.includeVelObs = includeVelObs
        self.num_envs = int(num_envs)
        # "connect": one batched simulation instead of p.connect(p.DIRECT)
        self._physics_client_id = B2Client(num_envs, device)
        self._physics_client_id.n_robot_obs = self._n_robot_obs
        self._robot = pandaEnv(self._physics_client_id, use_IK=self._use_IK, joint_action_space=numControlledJoints)
        self._world = WorldEnv(self._physics_client_id, obj_name=obj_name, obj_pose_rnd_std=obj_pose_rnd_std,
                               workspace_lim=self._robot.get_workspace())
        workspace = self._robot.get_workspace()
        workspace[2][0] = self._world.get_table_height() + z_low_offset
        self._robot.set_workspace(workspace)
        self._target_pose = np.zeros((self.num_envs, 3), np.float32)
        # constants -> simulation (observation limits are known from the lists above)
        lim = self._observation_limits()
        params = default_params(self._task, [x[0] for x in lim], [x[1] for x in lim],
                                n_act=self._robot.get_action_dim() + (1 if self._task == TASK_GRASP else 0),
                                n_ctrl=numControlledJoints, use_ik=use_IK,
                                ik_orientation=int(bool(self._robot._control_orientation)),
                                max_steps=max_steps, dist_min=target_dist_min, ws_lim=workspace,
                                eu_lim=self._robot.get_rotation_lim(), goal_env=int(getattr(self, '_goal_env', 0)))
        if self._task == TASK_GRASP:   # apply_action_fingers: force=10, maxVelocity=1 (reference panda_env.py:218-224)
            for d in (7, 8):
                self._robot.model.max_force[d] = 10.0
                self._robot.model.max_vel[d] = 1.0
        self._sim = self._physics_client_id.configure(self._robot.model, params)
        self._place_initial_world()
        self.observation_space, self.action_space = self.create_gym_spaces()
        self._fused = all(getattr(type(self), h) is getattr(self._base_cls(), h) for h in self._hooks)
        self._torch_out = None
        self.auto_reset = False   # set True for auto-resetting vectorised rollouts (host-array step path)
        # step() returns fresh arrays like the reference.  True: return VIEWS of the library's page-locked result
        # buffers instead (no copy; overwritten by the next step(), dangling after close()) — see INTEGRATION.md
        self.zero_copy_results = False
        self.seed()

    @classmethod
    def _base_cls(cls):
        for k in cls.__mro__:
            if k.__dict__.get("_is_task_impl"):
                return k
        return cls

    def _place_initial_world(self):
        # the reference constructs robot and world (object included) before the first reset()
        self._robot.reset()
        self._world.reset()
        self._sync_target()
        B = self.num_envs
        self._physics_client_id.set("counters", np.zeros((B, 2), np.int32))

    def _observation_limits(self):
        lim = self._robot.observation_limits() + self._world.observation_limits()
        lim += [[-0.5, 0.5]] * 3 + [[0, 2 * m.pi]] * 3
        if self._task in (TASK_PUSH, TASK_GRASP):
            lim += self._world.observation_limits()[:3]
        return lim

    def create_gym_spaces(self):
        lim = self._observation_limits()
        observation_space = spaces.Box(np.array([x[0] for x in lim]), np.array([x[1] for x in lim]), dtype='float32')
        action_dim = self._robot.get_action_dim()
        self.action_dim = action_dim
        action_high = np.array([1] * action_dim)
        action_space = spaces.Box(-action_high, action_high, dtype='float32')
        return observation_space, action_space

    # ------------------------------------------------------------------ reset
    def reset(self, env_ids=None):
        """``reset()`` resets every environment like the reference (:105-115).  ``reset(env_ids)`` resets only
        the listed environments of the batch (same 100 + 100 + 1 settle sequence, run on that subset)
        and returns their observations [len(env_ids), O]."""
        if env_ids is not None:
            return self._reset_subset(np.asarray(env_ids, np.int32))
        self.reset_simulation()
        if self._task in (TASK_PUSH, TASK_GRASP):
            world_obs, _ = self._world.get_observation()
            self._target_pose = self.sample_tg_pose(np.asarray(world_obs).reshape(self.num_envs, 6)[:, :3])
            self._sync_target()
            self._after_target()
        scaled = self._physics_client_id.observe()[0]
        return squeeze1(scaled.astype(np.float64), self.num_envs)

    def _after_target(self, ids=None):
        """Hook: per-episode constants that depend on the sampled target (iCub push: reward shaping distances)."""

    def reset_simulation(self):
        """resetSimulation + robot.reset + 100 steps + world.reset + 100 steps + 1 step
        (reference panda_push_gym_env.py:117-148)."""
        c = self._physics_client_id
        B = self.num_envs
        self.terminated = 0
        self._env_step_counter = 0
        c.set("counters", np.zeros((B, 2), np.int32))
        c.set("status", np.zeros((B, 4), np.int32))
        # p.resetSimulation removes every body: the object waits, at rest, away from the table
        c.set("obj_pose", np.tile(np.array(_PARK_POSE, np.float32), (B, 1)))
        c.set("obj_vel", np.zeros((B, 6), np.float32))
        c.set("cache_key", np.full((B, 16), -1, np.int32))
        c.set("cache_lam", np.zeros((B, 48), np.float32))
        self._robot.reset()
        c.step_simulation(100, binding.MODE_HOLD)
        self._world.reset()
        c.step_simulation(100, binding.MODE_HOLD)
        if self._use_IK:   # self._hand_pose = self._robot._home_hand_pose (reference :142-143)
            self._hand_pose = list(self._robot._home_hand_pose)
            c.set("hand_pose", np.tile(np.array(self._hand_pose, np.float32), (B, 1)))
        c.step_simulation(1, binding.MODE_HOLD)

    def _reset_subset(self, ids):
        c = self._physics_client_id
        n = len(ids)
        if n == 0:
            return np.zeros((0, self._sim.params.n_obs), np.float64)
        c.set_rows("counters", ids, np.zeros((n, 2), np.int32))
        c.set_rows("status", ids, np.zeros((n, 4), np.int32))
        c.set_rows("obj_pose", ids, np.tile(np.array(_PARK_POSE, np.float32), (n, 1)))
        c.set_rows("obj_vel", ids, np.zeros((n, 6), np.float32))
        c.set_rows("cache_key", ids, np.full((n, 16), -1, np.int32))
        c.set_rows("cache_lam", ids, np.zeros((n, 48), np.float32))
        self._robot.reset(ids)
        c.step_subset(ids, 100, binding.MODE_HOLD)
        self._world.reset(ids)
        c.step_subset(ids, 100, binding.MODE_HOLD)
        if self._use_IK:
            c.set_rows("hand_pose", ids, np.tile(np.array(self._robot._home_hand_pose, np.float32), (n, 1)))
        c.step_subset(ids, 1, binding.MODE_HOLD)
        if self._task in (TASK_PUSH, TASK_GRASP):
            obj = c.sim.get_rows("obj_pose", ids)[:, :3]
            tg = np.asarray(self.sample_tg_pose(obj), np.float32).reshape(n, 3)
            tp = np.asarray(self._target_pose, np.float32).reshape(-1, 3)
            if tp.shape[0] == self.num_envs:
                tp[ids] = tg
                self._target_pose = tp
            c.set_rows("target", ids, tg)
            self._after_target(ids)
        scaled = c.observe()[0]
        return scaled[ids].astype(np.float64)

    def _sync_target(self):
        tp = np.asarray(self._target_pose, np.float32).reshape(-1, 3)
        self._physics_client_id.set("target", np.broadcast_to(tp, (self.num_envs, 3)).copy())

    # ------------------------------------------------------------------ hooks (reference names)
    def get_extended_observation(self):
        raw = self._physics_client_id.observe()[3]
        self._observation = squeeze1(raw.astype(np.float64), self.num_envs)
        return np.array(self._observation), self._observation_limits()

    def apply_action(self, action):
        """Scaled action -> motor targets -> physics step(s) -> termination bookkeeping
        (reference panda_push_gym_env.py:189-242), one launch, no outputs."""
        a = self._as_batch(scale_gym_data(self.action_space, np.asarray(action, np.float32)))
        self._sim.step_host(a, self._action_repeat, binding.MODE_ACTION, want_obs=False)
        self._physics_client_id.invalidate()

    def _termination(self):
        done = self._physics_client_id.observe(latch=True)[2]
        self.terminated = squeeze1(self._physics_client_id.get("counters")[:, 1], self.num_envs)
        return squeeze1(done.astype(np.float32), self.num_envs)

    def _compute_reward(self):
        return squeeze1(self._physics_client_id.observe()[1], self.num_envs)

    # ------------------------------------------------------------------ step
    def _as_batch(self, a):
        a = np.asarray(a, np.float32)
        if a.ndim == 1:
            a = a[None, :]
        return np.ascontiguousarray(np.broadcast_to(a, (self.num_envs, a.shape[-1])))

    def step(self, action):
        if not self._fused:  # a subclass overrides a hook: follow the reference's call sequence
            self.apply_action(action)
            obs, _ = self.get_extended_observation()
            scaled_obs = scale_gym_data(self.observation_space, obs)
            done = self._termination()
            reward = self._compute_reward()
            return scaled_obs, np.array(reward), np.array(done), {}
        return self._step_fused(action)

    def _step_fused(self, action):
        """apply_action + stepSimulation + observation + termination + reward in ONE launch."""
        if hasattr(action, "is_cuda") and action.is_cuda:
            return self._step_device(action)
        a = np.asarray(action, np.float32)
        assert a.shape[-1:] == self.action_space.shape  # scale_gym_data's shape assert (utils.py:88)
        if self.num_envs == 1 and not self.auto_reset:
            obs, rew, done = self._sim.step_host(self._as_batch(a), self._action_repeat, binding.MODE_ACTION)
            self._physics_client_id.invalidate()
            return obs[0].astype(np.float64), np.array(rew[0]), np.array(done[0]), {}
        # batched host path: page-locked staging; the results land in page-locked arrays owned by the library
        obs, rew, done = self._sim.step_pinned(a if a.ndim == 2 else self._as_batch(a), self._action_repeat,
                                               binding.MODE_ACTION)
        self._physics_client_id.invalidate()
        if not getattr(self, "zero_copy_results", False):
            obs, rew, done = np.array(obs), np.array(rew), np.array(done)   # fresh arrays, like the reference
        if self.auto_reset and done.any():
            # vectorised-env convention: finished envs restart at once, their row of `obs` is the first
            # observation of the new episode; reward / done still describe the finished episode
            ids = np.nonzero(done)[0].astype(np.int32)
            obs = np.array(obs)
            rew, done = np.array(rew), np.array(done)
            obs[ids] = self._reset_subset(ids)
        return obs, rew, done, {}

    def _step_device(self, action):
        """Zero-copy path: ``action`` is a float32 CUDA tensor [B, A]; returns CUDA tensors."""
        import torch
        if self._torch_out is None:
            dev = action.device
            B = self.num_envs
            self._torch_out = (torch.empty((B, self._sim.params.n_obs), device=dev, dtype=torch.float32),
                               torch.empty(B, device=dev, dtype=torch.float32),
                               torch.empty(B, device=dev, dtype=torch.float32))
        obs, rew, done = self._torch_out
        assert action.dtype == torch.float32 and action.is_contiguous() and tuple(action.shape) == (self.num_envs, self.action_dim)
        self._sim.step(action, obs, rew, done, self._action_repeat, binding.MODE_ACTION,
                       stream=torch.cuda.current_stream(action.device).cuda_stream)
        self._physics_client_id.invalidate()
        return obs, rew, done, {}

    def seed(self, seed=None):
        self.np_random, seed = seeding.np_random(seed)
        self._world.seed(seed)
        self._robot.seed(seed)
        return [seed]

    def render(self, mode="rgb_array"):
        return np.array([])  # no rasteriser in this tier

    def close(self):
        self._physics_client_id.close()

    def debug_gui(self):
        pass

"""Shared machinery of the iCub task envs (reach / push / push-goal): the batched counterpart of reference
icub_envs/icub_{reach,push}_gym_env.py.  Everything that is not iCub-specific (reset sequence, hooks, the fused
``step()``) is inherited from the Panda task base — the reference spells the same code out once per robot."""
import math as m

import numpy as np

from pybullet_robot_envs.b2env.client import B2Client
from pybullet_robot_envs.b2env.model import TASK_PUSH, default_params, icub_params
from pybullet_robot_envs.envs.icub_envs.icub_env import iCubEnv
from pybullet_robot_envs.envs.panda_envs._panda_task import PandaTaskBase
from pybullet_robot_envs.envs.world_envs.world_env import WorldEnv


class ICubTaskBase(PandaTaskBase):
    _n_robot_obs = 19
    _reward_type = 0

    def _setup_icub(self, action_repeat, use_IK, control_arm, control_orientation, obj_name, obj_pose_rnd_std, renders,
                    max_steps, num_envs, device):
        self._time_step = 1. / 240.
        self._timeStep = self._time_step
        self._control_arm = control_arm
        self._use_IK = use_IK
        self._control_orientation = control_orientation
        self._action_repeat = action_repeat
        self._observation = []
        self._hand_pose = []
        self._env_step_counter = 0
        self._renders = renders
        self._max_steps = max_steps
        self._last_frame_time = 0
        self.terminated = 0
        self._target_dist_min = 0.03   # reference icub_push_gym_env.py:56, icub_reach_gym_env.py:52
        self.num_envs = int(num_envs)
        self._physics_client_id = B2Client(num_envs, device)
        self._physics_client_id.n_robot_obs = self._n_robot_obs
        self._robot = iCubEnv(self._physics_client_id, use_IK=self._use_IK, control_arm=self._control_arm,
                              control_orientation=self._control_orientation)
        self._world = WorldEnv(self._physics_client_id, obj_name=obj_name, obj_pose_rnd_std=obj_pose_rnd_std,
                               workspace_lim=self._robot.get_workspace())
        workspace = self._robot.get_workspace()
        workspace[2][0] = self._world.get_table_height()       # reference icub_push_gym_env.py:79-81
        self._robot.set_workspace(workspace)
        self._target_pose = np.zeros((self.num_envs, 3), np.float32)
        lim = self._observation_limits()
        ctrl = self._robot._ctrl_dofs
        params = default_params(self._task, [x[0] for x in lim], [x[1] for x in lim], n_act=self._robot.get_action_dim(),
                                n_ctrl=len(ctrl), use_ik=use_IK, ik_orientation=int(bool(control_orientation)),
                                max_steps=max_steps, dist_min=self._target_dist_min, ws_lim=workspace,
                                eu_lim=self._robot.get_rotation_lim(), goal_env=int(getattr(self, '_goal_env', 0)))
        icub_params(params, self._task, self._robot._control_arm, ctrl, control_orientation, self._reward_type)
        self._sim = self._physics_client_id.configure(self._robot.model, params)
        self._place_initial_world()
        self.observation_space, self.action_space = self.create_gym_spaces()
        self._fused = all(getattr(type(self), h) is getattr(self._base_cls(), h) for h in self._hooks)
        self._torch_out = None
        self.auto_reset = False
        self.zero_copy_results = False   # see PandaTaskBase._setup
        self.seed()

    @property
    def _tg_pose(self):          # the reference's iCub envs call the target `_tg_pose`
        return self._target_pose

    @_tg_pose.setter
    def _tg_pose(self, v):
        self._target_pose = v

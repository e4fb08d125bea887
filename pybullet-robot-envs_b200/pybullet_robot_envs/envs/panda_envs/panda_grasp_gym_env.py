"""``pandaGraspGymEnv`` / ``PandaGrasp-v0`` — BASELINE.json config 5.  The reference has NO grasp
env: this task is assembled from the reference's grasp primitives —
``pre_grasp / grasp / apply_action_fingers`` (reference panda_env.py:195-225: finger position control
with force 10, maxVelocity 1) and the scripted pick of examples/helloworlds/helloworld_panda.py:94-148.
Action = 7 joint increments (as pandaPush) + 1 gripper command in [-1, 1] (finger targets
0.02 + 0.02 g, i.e. 1 = open 0.04, -1 = closed).  Observation = the push layout (33) with the target
= object rest position + (0, 0, 0.1).  Reward = -|EE - obj| + 50 * lift (clamped), 1000 and done when
the object is 0.1 m above its rest height.  Throughput-only: there is no reference parity target."""
import numpy as np

from pybullet_robot_envs.b2env.model import TASK_GRASP
from pybullet_robot_envs.gym_compat import spaces
from pybullet_robot_envs.envs.world_envs.world_env import get_objects_list
from pybullet_robot_envs.envs.panda_envs._panda_task import PandaTaskBase


class pandaGraspGymEnv(PandaTaskBase):
    _task = TASK_GRASP
    _is_task_impl = True

    def __init__(self, numControlledJoints=7, use_IK=0, action_repeat=1, obj_name=get_objects_list()[1],
                 renders=False, max_steps=1000, obj_pose_rnd_std=0.05, includeVelObs=True, num_envs=1, device=0):
        if use_IK:
            raise NotImplementedError("PandaGrasp: joint control only")
        self._lift = 0.1
        self._setup(numControlledJoints, use_IK, action_repeat, obj_name, renders, max_steps, obj_pose_rnd_std,
                    includeVelObs, num_envs, device, target_dist_min=0.03, z_low_offset=-0.2)

    def create_gym_spaces(self):
        observation_space, _ = PandaTaskBase.create_gym_spaces(self)
        self.action_dim = self._robot.get_action_dim() + 1
        action_space = spaces.Box(-np.ones(self.action_dim), np.ones(self.action_dim), dtype='float32')
        return observation_space, action_space

    def sample_tg_pose(self, obj_pos):
        obj_pos = np.asarray(obj_pos, np.float64).reshape(-1, 3)
        pose = obj_pos + np.array([0.0, 0.0, self._lift])
        return tuple(pose[0]) if self.num_envs == 1 else pose

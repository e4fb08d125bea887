"""``pandaReachGymEnv`` — batched, CUDA-backed counterpart of reference
envs/panda_envs/panda_reach_gym_env.py:22-313."""
from pybullet_robot_envs.b2env.model import TASK_REACH
from pybullet_robot_envs.envs.world_envs.world_env import get_objects_list
from pybullet_robot_envs.envs.panda_envs._panda_task import PandaTaskBase


class pandaReachGymEnv(PandaTaskBase):
    _task = TASK_REACH
    _is_task_impl = True

    def __init__(self, numControlledJoints=7, use_IK=0, action_repeat=1, obj_name=get_objects_list()[1],
                 renders=False, max_steps=1000, obj_pose_rnd_std=0.0, includeVelObs=True, num_envs=1, device=0):
        # success radius 0.03 (reference :47); robot workspace floor = table height (:69)
        self._setup(numControlledJoints, use_IK, action_repeat, obj_name, renders, max_steps, obj_pose_rnd_std,
                    includeVelObs, num_envs, device, target_dist_min=0.03, z_low_offset=0.0)

"""``iCubReachGymEnv`` — batched, CUDA-backed counterpart of reference
envs/icub_envs/icub_reach_gym_env.py:23-330 (same constructor kwargs + ``num_envs``/``device``)."""
from pybullet_robot_envs.b2env.model import TASK_REACH
from pybullet_robot_envs.envs.icub_envs._icub_task import ICubTaskBase
from pybullet_robot_envs.envs.world_envs.world_env import get_objects_list


class iCubReachGymEnv(ICubTaskBase):
    _task = TASK_REACH
    _is_task_impl = True

    def __init__(self, action_repeat=1, use_IK=1, control_arm='l', control_orientation=0, obj_name=get_objects_list()[0],
                 obj_pose_rnd_std=0, renders=False, max_steps=2000, num_envs=1, device=0):
        self._setup_icub(action_repeat, use_IK, control_arm, control_orientation, obj_name, obj_pose_rnd_std, renders,
                         max_steps, num_envs, device)

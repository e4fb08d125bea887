"""``iCubPushGymGoalEnv`` — the HER / GoalEnv variant of the iCub push task (reference
envs/icub_envs/icub_push_gym_goal_env.py:20-132): dict observation, sparse reward, ``is_success``."""
from pybullet_robot_envs import gym_compat as gym
from pybullet_robot_envs.envs.icub_envs.icub_push_gym_env import iCubPushGymEnv
from pybullet_robot_envs.envs.panda_envs.panda_push_gym_goal_env import GoalMixin
from pybullet_robot_envs.envs.world_envs.world_env import get_objects_list


class iCubPushGymGoalEnv(GoalMixin, gym.GoalEnv, iCubPushGymEnv):
    _goal_env = 1
    _box_cls = iCubPushGymEnv

    def __init__(self, action_repeat=1, use_IK=1, control_arm='l', control_orientation=0, obj_name=get_objects_list()[1],
                 obj_pose_rnd_std=0, tg_pose_rnd_std=0.2, renders=False, max_steps=2000, reward_type=1, num_envs=1, device=0):
        iCubPushGymEnv.__init__(self, action_repeat, use_IK, control_arm, control_orientation, obj_name, obj_pose_rnd_std,
                                tg_pose_rnd_std, renders, max_steps, reward_type, num_envs=num_envs, device=device)

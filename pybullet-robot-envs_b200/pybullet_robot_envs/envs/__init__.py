from pybullet_robot_envs.envs.panda_envs.panda_env import pandaEnv  # noqa: F401
from pybullet_robot_envs.envs.panda_envs.panda_reach_gym_env import pandaReachGymEnv  # noqa: F401
from pybullet_robot_envs.envs.panda_envs.panda_push_gym_env import pandaPushGymEnv  # noqa: F401
from pybullet_robot_envs.envs.panda_envs.panda_push_gym_goal_env import pandaPushGymGoalEnv  # noqa: F401
from pybullet_robot_envs.envs.panda_envs.panda_grasp_gym_env import pandaGraspGymEnv  # noqa: F401
from pybullet_robot_envs.envs.world_envs.world_env import WorldEnv, get_objects_list  # noqa: F401
from pybullet_robot_envs.envs.icub_envs.icub_env import iCubEnv  # noqa: F401
from pybullet_robot_envs.envs.icub_envs.icub_reach_gym_env import iCubReachGymEnv  # noqa: F401
from pybullet_robot_envs.envs.icub_envs.icub_push_gym_env import iCubPushGymEnv  # noqa: F401
from pybullet_robot_envs.envs.icub_envs.icub_push_gym_goal_env import iCubPushGymGoalEnv  # noqa: F401

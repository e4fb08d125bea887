"""Gym registration — ids and default kwargs of reference pybullet_robot_envs/__init__.py:47-80
(Panda ids; the iCub ids need the 32-dof tree kernel).  The reference registers
``renders: True``; the CUDA backend has no GUI, so ids register with ``renders: False``.
BASELINE.json spells the ids with a capital P: both spellings are registered."""
try:  # a real gym wins if present
    import gym  # noqa: F401
    from gym.envs.registration import register
except Exception:  # this image: use the bundled shim
    from pybullet_robot_envs import gym_compat as gym  # noqa: F401
    from pybullet_robot_envs.gym_compat import register

_REACH_KW = {'numControlledJoints': 7, 'use_IK': 0, 'obj_pose_rnd_std': 0.05, 'includeVelObs': True,
             'max_steps': 1000, 'renders': False}
_PUSH_KW = {'numControlledJoints': 7, 'use_IK': 0, 'obj_pose_rnd_std': 0.05, 'tg_pose_rnd_std': 0,
            'includeVelObs': True, 'max_steps': 1000, 'renders': False}

for _name in ('pandaReach-v0', 'PandaReach-v0'):
    register(id=_name, entry_point='pybullet_robot_envs.envs:pandaReachGymEnv', max_episode_steps=1000,
             kwargs=dict(_REACH_KW))
for _name in ('pandaPush-v0', 'PandaPush-v0'):
    register(id=_name, entry_point='pybullet_robot_envs.envs:pandaPushGymEnv', max_episode_steps=1000,
             kwargs=dict(_PUSH_KW))
for _name in ('pandaPushGoal-v0', 'PandaPushGoal-v0'):   # reference __init__.py:70-80
    register(id=_name, entry_point='pybullet_robot_envs.envs:pandaPushGymGoalEnv', max_episode_steps=1000,
             kwargs=dict(_PUSH_KW))
for _name in ('pandaGrasp-v0', 'PandaGrasp-v0'):   # BASELINE.json config 5: new task, no reference counterpart
    register(id=_name, entry_point='pybullet_robot_envs.envs:pandaGraspGymEnv', max_episode_steps=1000,
             kwargs={'numControlledJoints': 7, 'use_IK': 0, 'obj_pose_rnd_std': 0.05, 'max_steps': 1000, 'renders': False})


def getList():
    return ['pandaReach-v0', 'pandaPush-v0', 'pandaPushGoal-v0', 'PandaReach-v0', 'PandaPush-v0', 'PandaPushGoal-v0',
            'pandaGrasp-v0', 'PandaGrasp-v0']

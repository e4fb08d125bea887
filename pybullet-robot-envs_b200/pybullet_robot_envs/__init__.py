"""Gym registration — ids and default kwargs of reference pybullet_robot_envs/__init__.py:47-80
(Panda ids) and :7-44 (iCub ids).  The reference registers
``renders: True``; the CUDA backend has no GUI, so ids register with ``renders: False``.
BASELINE.json spells the ids with a capital P: both spellings are registered."""
from pybullet_robot_envs import gym_compat
from pybullet_robot_envs.gym_compat import register as _register_compat

try:  # a real gym, when present, gets the same ids (the env classes subclass gym_compat.Env: duck-typed gym API)
    import gym  # noqa: F401
    from gym.envs.registration import register as _register_gym
except Exception:  # this image: only the bundled shim
    gym = gym_compat
    _register_gym = None


def register(**kw):
    """Every id is registered with the bundled shim (``gym_compat.make`` always works) and, when a real ``gym`` is
    importable, with its registry as well."""
    _register_compat(**kw)
    if _register_gym is not None:
        try:
            _register_gym(**kw)
        except Exception:   # e.g. the id is already taken in a long-lived interpreter
            pass


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


# iCub ids (reference __init__.py:7-44)
register(id='iCubReach-v0', entry_point='pybullet_robot_envs.envs:iCubReachGymEnv', max_episode_steps=1000,
         kwargs={'use_IK': 1, 'control_arm': 'l', 'control_orientation': 0, 'obj_pose_rnd_std': 0, 'max_steps': 1000,
                 'renders': False})
register(id='iCubPush-v0', entry_point='pybullet_robot_envs.envs:iCubPushGymEnv', max_episode_steps=1000,
         kwargs={'use_IK': 1, 'control_arm': 'l', 'control_orientation': 0, 'obj_pose_rnd_std': 0.05,
                 'tg_pose_rnd_std': 0, 'max_steps': 1000, 'reward_type': 0, 'renders': False})
register(id='iCubPushGoal-v0', entry_point='pybullet_robot_envs.envs:iCubPushGymGoalEnv', max_episode_steps=1000,
         kwargs={'use_IK': 1, 'control_arm': 'r', 'control_orientation': 1, 'obj_pose_rnd_std': 0.05,
                 'tg_pose_rnd_std': 0, 'max_steps': 1000, 'renders': False})


def getList():
    return ['iCubReach-v0', 'iCubPush-v0', 'iCubPushGoal-v0', 'pandaReach-v0', 'pandaPush-v0', 'pandaPushGoal-v0', 'PandaReach-v0', 'PandaPush-v0', 'PandaPushGoal-v0',
            'pandaGrasp-v0', 'PandaGrasp-v0']

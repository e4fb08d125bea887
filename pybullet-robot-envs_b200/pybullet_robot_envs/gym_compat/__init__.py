"""Minimal stand-in for the slice of OpenAI gym 0.12.5 the reference uses
(requirements.txt:1): ``Env``, ``GoalEnv``, ``spaces.Box/Dict``, ``utils.seeding.np_random``,
``envs.registration.register`` and ``make`` with the ``TimeLimit`` wrapper.  ``gym`` is not
installable in this image; if a real ``gym`` is importable the package uses it instead
(see ``pybullet_robot_envs/__init__.py``).
"""
from . import spaces, seeding  # noqa: F401
from .core import Env, GoalEnv, Wrapper, TimeLimit  # noqa: F401
from .registration import register, make, registry, spec  # noqa: F401

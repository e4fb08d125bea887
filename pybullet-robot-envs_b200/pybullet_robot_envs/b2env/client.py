"""The object that stands where the reference passes ``physicsClientId``.

In the reference, ``p.connect(p.DIRECT)`` (panda_push_gym_env.py:62) returns an int naming an
in-process Bullet server that every robot/world/task object then addresses.  Here the same
role is played by a ``B2Client``: it owns one batched CUDA simulation (``B2Sim``) of
``num_envs`` environments on one GPU and hands it to the robot, world and task objects.
"""
import numpy as np

from . import binding
from .model import TASK_REACH, panda_task_setup


class B2Client:
    def __init__(self, num_envs=1, device=0):
        self.num_envs = int(num_envs)
        self.device = int(device)
        self.sim = None
        self.model = None
        self.params = None
        self._obs_cache_valid = False
        self._last = None
        self.pending_mode = None   # set by robot.apply_action: how the next stepSimulation must treat the motors

    def configure(self, model, params):
        """(Re)create the simulation for a model + task constants."""
        if self.sim is not None:
            self.sim.close()
        self.model, self.params = model, params
        self.sim = binding.B2Sim(model, params, self.num_envs, self.device)
        self._obs_cache_valid = False
        return self.sim

    def ensure(self):
        if self.sim is None:
            m, p = panda_task_setup(TASK_REACH)
            self.configure(m, p)
        return self.sim

    # ---- p.stepSimulation ---------------------------------------------------------------
    def step_simulation(self, n=1, mode=None):
        if mode is None:
            mode = self.pending_mode if self.pending_mode is not None else binding.MODE_HOLD
        self.pending_mode = None
        self.ensure().step_host(None, n, mode, want_obs=False)
        self._obs_cache_valid = False

    # ---- observation / reward / done of the current state (no physics) --------------------
    def observe(self, latch=False):
        """Returns (scaled_obs, reward, done, raw_obs) for the current state.  ``latch=True`` also stores the
        success latch like the reference's ``_termination()`` does (panda_push_gym_env.py:305-306)."""
        if latch or not self._obs_cache_valid:
            sim = self.ensure()
            mode = binding.MODE_HOLD if latch else binding.MODE_OBSERVE
            obs, rew, done = sim.step_host(None, 0, mode, want_obs=True)
            self._last = (obs, rew, done, sim.get("raw_obs"))
            self._obs_cache_valid = True
        return self._last

    def invalidate(self):
        self._obs_cache_valid = False

    def get(self, name):
        return self.ensure().get(name)

    def set(self, name, value):
        self.ensure().set(name, value)
        self._obs_cache_valid = False

    def set_rows(self, name, ids, value):
        self.ensure().set_rows(name, ids, value)
        self._obs_cache_valid = False

    def step_subset(self, ids, n=1, mode=binding.MODE_HOLD):
        self.ensure().step_subset(ids, n, mode)
        self._obs_cache_valid = False

    def close(self):
        if self.sim is not None:
            self.sim.close()
            self.sim = None


def squeeze1(x, num_envs):
    """Single-env compatibility: drop the batch axis so shapes equal the reference's."""
    return x[0] if num_envs == 1 else x

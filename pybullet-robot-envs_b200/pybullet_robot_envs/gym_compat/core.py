"""``gym.Env`` / ``GoalEnv`` / ``Wrapper`` / ``TimeLimit`` (gym 0.12.5 semantics)."""


class Env:
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return

    @property
    def unwrapped(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


class GoalEnv(Env):
    def reset(self):
        from . import spaces
        if not isinstance(self.observation_space, spaces.Dict):
            raise RuntimeError("GoalEnv requires an observation space of type gym.spaces.Dict")
        for k in ("observation", "achieved_goal", "desired_goal"):
            if k not in self.observation_space.spaces:
                raise RuntimeError("GoalEnv requires the key %s in the observation space" % k)

    def compute_reward(self, achieved_goal, desired_goal, info):
        raise NotImplementedError


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.reward_range = env.reward_range
        self.metadata = env.metadata

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def render(self, mode="human", **kw):
        return self.env.render(mode, **kw)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)

    def compute_reward(self, achieved_goal, desired_goal, info):
        return self.env.compute_reward(achieved_goal, desired_goal, info)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class TimeLimit(Wrapper):
    """``register(max_episode_steps=1000)`` (reference __init__.py:10) wraps every env in this.  Batched envs return
    per-env ``done`` arrays: the limit then raises every entry, and with ``auto_reset`` the elapsed steps are counted
    per environment (an env that restarted inside ``step()`` starts from 0 again)."""

    def __init__(self, env, max_episode_steps=None):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0

    def step(self, action):
        import numpy as np
        observation, reward, done, info = self.env.step(action)
        batched = hasattr(done, "shape") and getattr(done, "shape", ()) != ()
        if batched and getattr(self.env.unwrapped, "auto_reset", False) and not hasattr(done, "is_cuda"):
            if np.isscalar(self._elapsed_steps):
                self._elapsed_steps = np.full(done.shape, self._elapsed_steps, np.int64)
            self._elapsed_steps += 1
            over = self._elapsed_steps >= self._max_episode_steps if self._max_episode_steps is not None else np.zeros(done.shape, bool)
            finished = np.asarray(done).astype(bool)
            done = np.where(over, np.ones_like(done), done)
            self._elapsed_steps[finished] = 0     # restarted by the env itself
            return observation, reward, done, info
        self._elapsed_steps += 1
        if self._max_episode_steps is not None and np.all(self._elapsed_steps >= self._max_episode_steps):
            if batched:
                done = done.clone().fill_(1) if hasattr(done, "is_cuda") else np.ones_like(done)
            else:
                done = True
        return observation, reward, done, info

    def reset(self, **kw):
        self._elapsed_steps = 0
        return self.env.reset(**kw)

"""Minimal stand-in for the parts of OpenAI gym the reference touches (train.py:17-25,
learner/gnn_dagger.py:150-163,242; test_model.py:17-47): ``gym.make`` returning a time-limited wrapper
whose raw environment is ``.env``; old-style API (4-tuple ``step``, ``env.seed``)."""

_REGISTRY = {}


def register(id, entry_point, max_episode_steps=200):
    _REGISTRY[id] = (entry_point, max_episode_steps)


class TimeLimit:
    """Ends an episode after ``max_episode_steps`` steps (the raw env never sets ``done``)."""

    def __init__(self, env, max_episode_steps):
        self.env = env
        self._max_episode_steps = max_episode_steps
        self._elapsed = 0

    def seed(self, seed=None):
        return self.env.seed(seed)

    def reset(self):
        self._elapsed = 0
        return self.env.reset()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self._elapsed += 1
        if self._elapsed >= self._max_episode_steps:
            done = True
        return obs, reward, done, info

    def render(self, mode="human"):
        return self.env.render(mode)

    def close(self):
        return self.env.close()

    @property
    def unwrapped(self):
        return self.env


def make(id):
    if id not in _REGISTRY:
        import gym_flock  # noqa: F401  (registers the flocking ids)
    if id not in _REGISTRY:
        raise KeyError(f"no registered env with id {id!r}")
    entry, steps = _REGISTRY[id]
    return TimeLimit(entry(), steps)

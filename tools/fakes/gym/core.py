class Env(object):
    metadata = {}
    spec = None

    def seed(self, seed=None):
        return []

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_") and name not in ("_max_episode_steps", "_elapsed_steps"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def spec(self):
        return None

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def seed(self, seed=None):
        return self.env.seed(seed)

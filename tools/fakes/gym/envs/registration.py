class _Registry:
    env_specs = {}


registry = _Registry()


def register(id, **kwargs):
    registry.env_specs[id] = kwargs

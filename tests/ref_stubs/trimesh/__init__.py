"""Import-only stand-in (test infrastructure): the entry points exercised (train.py, render.py) never call into trimesh."""


def __getattr__(name):
    raise AttributeError(f"trimesh stand-in: '{name}' is not available in this image")

"""Import-only stand-in (test infrastructure): the entry points exercised (train.py, render.py) never call into open3d."""


def __getattr__(name):
    raise AttributeError(f"open3d stand-in: '{name}' is not available in this image")

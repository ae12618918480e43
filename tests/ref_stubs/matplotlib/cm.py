def __getattr__(name):
    raise AttributeError(f"matplotlib stand-in: cm.{name} is not available in this image")

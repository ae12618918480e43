def __getattr__(name):
    raise AttributeError(f"matplotlib stand-in: pyplot.{name} is not available in this image")

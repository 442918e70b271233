def __getattr__(name):
    raise RuntimeError("matplotlib stub: plotting is not available (%s)" % name)

class PatchCollection(object):
    def __init__(self, *a, **k):
        raise RuntimeError("matplotlib stub: plotting is not available")

"""vegasflow stand-in so that madflow.utilities can be imported.  TEST INFRASTRUCTURE ONLY."""


def vegas_wrapper(*a, **k):
    raise RuntimeError("tfshim: vegasflow is not available offline")


class VegasFlow:
    def __init__(self, *a, **k):
        raise RuntimeError("tfshim: vegasflow is not available offline")

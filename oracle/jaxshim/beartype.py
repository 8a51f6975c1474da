"""beartype stand-in (the reference imports it; none of the hot-path functions is decorated)."""


def beartype(fn):
    return fn

"""jaxtyping stand-in: annotations only (reference: `Float[Array, "b t h w c"]`)."""


class _Sub:
    def __class_getitem__(cls, item):
        return cls


class Float(_Sub):
    pass


class Int(_Sub):
    pass


class Bool(_Sub):
    pass


class Array:
    pass


def jaxtyped(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn

"""`flax` stand-in (only `flax.nnx`) -- see oracle/jaxshim/README.md.  TEST INFRASTRUCTURE ONLY."""
from . import nnx  # noqa: F401

__version__ = "0+torchshim"

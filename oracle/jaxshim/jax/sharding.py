"""jax.sharding on ONE host device: meshes and shardings are descriptions that place nothing.  TEST INFRASTRUCTURE ONLY.

The reference's scripts build a 1-D ('data',) mesh over jax.devices() and put replicated / batch-sharded arrays on it
(claude_distributed/test_rl_model.py:25-27, distributed_train.py:107-109).  With a single device both placements are the
array itself, which is what jax does on a one-device mesh too."""


class PartitionSpec(tuple):
    def __new__(cls, *axes):
        return super().__new__(cls, axes)


class Mesh:
    def __init__(self, devices, axis_names):
        self.devices = list(devices)
        self.axis_names = tuple(axis_names) if not isinstance(axis_names, str) else (axis_names,)
        self.shape = {self.axis_names[0]: len(self.devices)} if self.axis_names else {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class NamedSharding:
    def __init__(self, mesh, spec):
        self.mesh, self.spec = mesh, spec

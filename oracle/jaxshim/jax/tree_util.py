"""jax.tree_util subset over nested dict / list / tuple containers."""


class DictKey:
    def __init__(self, key):
        self.key = key

    def __repr__(self):
        return f"[{self.key!r}]"


class SequenceKey:
    def __init__(self, idx):
        self.idx = idx

    def __repr__(self):
        return f"[{self.idx}]"


def tree_leaves_with_path(tree, path=()):
    if hasattr(tree, "to_pure_dict"):
        tree = tree.to_pure_dict()
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree, key=lambda v: (str(type(v)), v)):
            out += tree_leaves_with_path(tree[k], path + (DictKey(k),))
        return out
    if isinstance(tree, (list, tuple)):
        out = []
        for i, v in enumerate(tree):
            out += tree_leaves_with_path(v, path + (SequenceKey(i),))
        return out
    if tree is None:
        return []
    return [(path, tree)]


def tree_leaves(tree):
    return [v for _, v in tree_leaves_with_path(tree)]


def tree_map(fn, tree, *rest):
    if hasattr(tree, "to_pure_dict"):
        tree = tree.to_pure_dict()
        rest = [r.to_pure_dict() if hasattr(r, "to_pure_dict") else r for r in rest]
    if isinstance(tree, dict):
        return {k: tree_map(fn, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    if tree is None:
        return None
    return fn(tree, *rest)

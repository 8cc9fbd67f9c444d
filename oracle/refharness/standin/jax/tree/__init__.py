from ..tree_util import tree_map as map  # noqa

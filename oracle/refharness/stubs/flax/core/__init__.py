class FrozenDict(dict): pass

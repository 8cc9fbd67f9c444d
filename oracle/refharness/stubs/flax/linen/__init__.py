class Module: pass
def compact(f): return f

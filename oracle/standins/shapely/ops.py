"""shapely.ops stand-in: only the bbox of the union is ever read (model.py:294-306)."""


class _Union(object):
    def __init__(self, bounds):
        self.bounds = bounds


def unary_union(polys):
    b = [p.bounds for p in polys]
    return _Union((min(x[0] for x in b), min(x[1] for x in b), max(x[2] for x in b), max(x[3] for x in b)))

"""shapely.geometry stand-in.  Data plumbing only: closed ring in input order, bounds."""
import numpy as np


def _xy(p):
    a = np.asarray(p, dtype=np.float64).reshape(-1)
    return (float(a[0]), float(a[1]))


class _Ring(object):
    def __init__(self, coords):
        self.coords = coords


class Polygon(object):
    def __init__(self, shell):
        pts = [_xy(p) for p in shell]
        if pts[0] != pts[-1]:
            pts.append(pts[0])
        self.exterior = _Ring(pts)
        xs = [p[0] for p in pts]
        ys = [p[1] for p in pts]
        self.bounds = (min(xs), min(ys), max(xs), max(ys))


class Point(object):
    def __init__(self, x, y):
        self.x = float(x)
        self.y = float(y)

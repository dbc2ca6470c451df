"""Minimal stand-ins for the matplotlib.patches the reference's planar environment takes as
obstacles (planners/prm/planar/env.py:12): only the accessors its predicates use."""
import numpy as np


class _BBox:
    def __init__(self, x0, y0, x1, y1):
        self._p = np.array([[x0, y0], [x1, y1]], dtype=float)

    def get_points(self):
        return self._p

    @property
    def bounds(self):
        (x0, y0), (x1, y1) = self._p
        return (x0, y0, x1 - x0, y1 - y0)


class Rectangle:
    def __init__(self, xy, width, height, **kwargs):
        self.xy, self.width, self.height = xy, width, height

    def get_bbox(self):
        x, y = self.xy
        return _BBox(x, y, x + self.width, y + self.height)


class Circle:
    def __init__(self, xy, radius, **kwargs):
        self.center, self.radius = np.array(xy, dtype=float), radius


def is_rectangle(ob):
    return hasattr(ob, "get_bbox")


def is_circle(ob):
    return hasattr(ob, "center") and hasattr(ob, "radius")

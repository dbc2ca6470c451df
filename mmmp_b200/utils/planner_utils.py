"""Sampling helpers of the reference's utils/planner_utils.py that the planners rely on
(reference utils/planner_utils.py:17-50): the Interval tuple and the limit constants."""
from collections import namedtuple

import numpy as np

Interval = namedtuple("Interval", ["lower", "upper"])

UNIT_LIMITS = Interval(0.0, 1.0)
CIRCULAR_LIMITS = Interval(-np.pi, np.pi)
UNBOUNDED_LIMITS = Interval(-np.inf, np.inf)


def intervals_from_limits(limits):
    """[(lo, hi), ...] -> [Interval]; a joint whose upper < lower is circular
    (is_circular_joint, planner_utils.py:24-29)."""
    return [CIRCULAR_LIMITS if hi < lo else Interval(float(lo), float(hi)) for lo, hi in limits]

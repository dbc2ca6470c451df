"""Planar Prioritized PRM, learning phase (reference
planners/prm/planar/multi_arm/decoupled/prioritized_prm.py:16-105): identical roadmap
construction to the planar CBS-PRM (:46-58); only the query differs."""
from .cbs_prm import CBSPRM


class PrioritizedPRM(CBSPRM):
    def query(self):
        raise NotImplementedError("the planar prioritized search is host code outside the roadmap-construction "
                                  "path (SURVEY.md 8f)")

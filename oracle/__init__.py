"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatements (numpy float64 / plain C) of the reference's algorithms on
the roadmap-construction hot path.  Every function cites the reference
file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product
package `mmmp_b200` never does.

Pinning status:
  planar FK + predicates, Panda FK + metrics, heap k-NN / greedy connect:
      pinned against the reference itself, imported in the build container
      (tools/gen_golden.py -> tests/golden/*.npz) and against the reference's
      own fixtures (res/images/*_roadmap.csv, tests/robots/test_planar_arm.py).
  Panda / KUKA collision (hull GJK, oracle/hull_gjk.c): PARITY UNPINNED --
      the arithmetic lives in pybullet==3.2.6 which is not in the reference
      tree and not installable here; no reference test records a verdict.
"""

"""mmmp_b200 -- B200-native PRM roadmap construction behind the MMMP planner API.

Only the roadmap-construction hot path of ViktorLaurens/MMMP lives here (SURVEY.md
section 8): batched sampling, forward kinematics, configuration / edge collision
checks and brute-force neighbour search as hand-written sm_100a CUDA kernels behind
a C ABI (include/mmmp_b200.h), plus the host-side mirror of the reference's
`robots/` and `planners/prm/` interfaces.  There is no CPU fallback.
"""
__version__ = "0.1.0"

"""mom5_b200 -- B200-native (sm_100a) implementation of MOM5's tracer-advection hot path
(ocean_tracer_advect_mod: advect_tracer_sweby_all / advect_tracer_mdfl_sweby / quicker / upwind).

The compute path lives in mom5_b200/csrc (hand-written CUDA behind the C ABI of include/mom5adv.h, built
in-tree as mom5_b200/libmom5adv.so).  There is no CPU fallback: calling any operator without the built
library / without a GPU raises.
"""
__version__ = "0.1.0"

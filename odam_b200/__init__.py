"""odam_b200 -- B200-native drop-in for ONE hot path of likojack/ODAM: the multi-view superquadric optimisation.

    from odam_b200.sq_libs import SuperQuadricOptimizer, SuperQuadric      # reference src/super_quadric/sq_libs.py
    from odam_b200.run_multi_view import optim_process                      # reference src/scripts/run_multi_view.py
    from odam_b200 import api                                               # packed arrays <-> C ABI (include/odam_sq.h)

The compute lives in odam_b200/lib/libodam_sq.so (hand-written sm_100a CUDA, built by `python -m odam_b200.build`).
Nothing in this package falls back to the CPU, and nothing in it imports oracle/.
"""
__version__ = "0.1.0"

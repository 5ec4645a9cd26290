"""pcg_b200 — B200-native backend for PCG's pairwise direct-optimization (DO) hot path.

Only what the path needs: ``csrc/`` (CUDA kernels + the C ABI of ``include/pcg_do_b200.h``),
``backend`` (ctypes binding), ``pairwise`` (host mirror of the reference's operator interface).
"""
from .backend import BackendError, Context, Forest, Memo, Multi, Preorder, Tcm, load  # noqa: F401
from .pairwise import (PackedPairs, align2d_batch, foreign_pairwise_do,  # noqa: F401
                       foreign_pairwise_do_batch)

"""ctypes binding of the C-ABI CUDA library (include/pcg_do_b200.h).

There is no CPU fallback: importing works anywhere (so the symbol/ABI tests can run without a
GPU), but creating a context raises unless a CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)

# every symbol include/pcg_do_b200.h declares
EXPORTED_SYMBOLS = [
    "setUp2dCostMtx", "align2d", "align2dAffine",
    "pcgdo_create", "pcgdo_destroy", "pcgdo_last_error", "pcgdo_configure", "pcgdo_host_alloc", "pcgdo_host_free",
    "pcgdo_tcm_create", "pcgdo_tcm_destroy",
    "pcgdo_align2d_batch_device", "pcgdo_align2d_batch_host", "pcgdo_last_kernel_ms",
    "pcgdo_configure_ex", "pcgdo_int32_peak",
    "pcgdo_metric_batch_device", "pcgdo_metric_batch_host",
    "pcgdo_tcm_create_affine", "pcgdo_align2d_affine_batch_device", "pcgdo_align2d_affine_batch_host",
    "pcgdo_forest_create", "pcgdo_forest_destroy", "pcgdo_forest_set_leaves", "pcgdo_forest_score",
    "pcgdo_forest_tree_cost_ptr", "pcgdo_forest_tree_costs", "pcgdo_forest_copy_tree_costs", "pcgdo_forest_get_node", "pcgdo_forest_levels",
    "pcgdo_forest_alignments",
    "pcgdo_forest_enable_rerooting", "pcgdo_forest_score_rerooted", "pcgdo_forest_get_edges", "pcgdo_forest_edges", "pcgdo_forest_try_leaf", "pcgdo_forest_create32", "pcgdo_forest_set_leaves32",
    "pcgdo_memo_create", "pcgdo_memo_create_metric", "pcgdo_memo_structure", "pcgdo_memo_alphabet", "pcgdo_memo_destroy", "pcgdo_overlap2d_batch_device", "pcgdo_overlap2d_batch_host",
    "pcgdo_explicit_batch_device", "pcgdo_explicit_batch_host",
    "matrixInit", "matrixDestroy", "getCostAndMedian2D", "getCostAndMedian3D",
    "setUp3dCostMtx", "freeCostMtx", "pcgdo_configure_env",
    "pcgdo_overlap_sets_batch_device", "pcgdo_overlap_sets_batch_host",
    "pcgdo_forest_set_trees", "pcgdo_forest_cells",
    "pcgdo_multi_create", "pcgdo_multi_destroy", "pcgdo_multi_devices", "pcgdo_multi_ctx", "pcgdo_multi_last_error",
    "pcgdo_multi_configure", "pcgdo_multi_tcm_create", "pcgdo_multi_tcm_destroy",
    "pcgdo_mforest_create", "pcgdo_mforest_destroy", "pcgdo_mforest_set_leaves", "pcgdo_mforest_set_trees",
    "pcgdo_mforest_enable_rerooting", "pcgdo_mforest_score", "pcgdo_mforest_block", "pcgdo_multi_align2d_batch_host",
    "pcgdo_preorder_create", "pcgdo_preorder_destroy", "pcgdo_preorder_set_contexts", "pcgdo_preorder_run",
    "pcgdo_preorder_length", "pcgdo_preorder_get", "pcgdo_preorder_device_ptr",
]


class AlignIO(C.Structure):
    """alignIO_t (reference: c_alignment_interface.h:14-18)."""
    _fields_ = [("character", _u32p), ("length", C.c_size_t), ("capacity", C.c_size_t)]


class CostMatrices2D(C.Structure):
    """cost_matrices_2d_t (reference: costMatrix.h:35-93)."""
    _fields_ = [("alphSize", C.c_size_t), ("costMatrixDimension", C.c_size_t), ("gap_char", C.c_uint),
                ("cost_model_type", C.c_int), ("include_ambiguities", C.c_int), ("gap_open_cost", C.c_uint),
                ("is_metric", C.c_int), ("num_elements", C.c_size_t), ("cost", _u32p), ("median", _u32p),
                ("worst", _u32p), ("prepend_cost", _u32p), ("tail_cost", _u32p)]


class CostMatrices3D(C.Structure):
    """cost_matrices_3d_t (reference: costMatrix.h:102-130)."""
    _fields_ = [("alphSize", C.c_size_t), ("costMatrixDimension", C.c_size_t), ("gap_char", C.c_uint),
                ("cost_model_type", C.c_int), ("include_ambiguities", C.c_int), ("gap_open_cost", C.c_uint),
                ("num_elements", C.c_size_t), ("cost", _u32p), ("median", _u32p)]


class Batch(C.Structure):
    """pcgdo_batch."""
    _fields_ = [("n_pairs", C.c_size_t), ("elems", C.c_void_p), ("elems_bytes", C.c_size_t),
                ("off1", C.c_void_p), ("off2", C.c_void_p), ("len1", C.c_void_p), ("len2", C.c_void_p),
                ("out_off", C.c_void_p), ("out_bytes", C.c_size_t),
                ("out_gapped", C.c_void_p), ("out_slot1", C.c_void_p), ("out_slot2", C.c_void_p),
                ("out_len", C.c_void_p), ("cost", C.c_void_p), ("cells", C.c_void_p), ("out_off_actual", C.c_void_p)]


class Batch32(C.Structure):
    """pcgdo_batch32."""
    _fields_ = [("n_pairs", C.c_size_t), ("elems", C.c_void_p), ("elems_count", C.c_size_t),
                ("off1", C.c_void_p), ("off2", C.c_void_p), ("len1", C.c_void_p), ("len2", C.c_void_p),
                ("out_off", C.c_void_p), ("out_count", C.c_size_t),
                ("out_med", C.c_void_p), ("out_a1", C.c_void_p), ("out_a2", C.c_void_p),
                ("out_len", C.c_void_p), ("cost", C.c_void_p), ("cells", C.c_void_p), ("final_offset", C.c_void_p)]


class DcElement(C.Structure):
    """dcElement_t (reference: lib/tcm-memo/ffi/memoized-tcm/dynamicCharacterOperations.h)."""
    _fields_ = [("alphSize", C.c_size_t), ("element", _u64p)]


class BackendError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the library (building it first if needed). Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        if not build_if_missing:
            raise BackendError(f"{path} is missing: the CUDA extension has not been built")
        _build.build()
    L = C.CDLL(path)
    L.pcgdo_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.pcgdo_destroy.argtypes = [C.c_void_p]
    L.pcgdo_last_error.restype = C.c_char_p
    L.pcgdo_last_error.argtypes = [C.c_void_p]
    L.pcgdo_configure.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.pcgdo_host_alloc.argtypes = [C.c_void_p, C.c_size_t]
    L.pcgdo_host_alloc.restype = C.c_void_p
    L.pcgdo_host_free.argtypes = [C.c_void_p, C.c_void_p]
    L.pcgdo_host_free.restype = None
    L.pcgdo_configure_ex.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
    L.pcgdo_tcm_create.argtypes = [C.c_void_p, _u32p, C.c_size_t, C.c_uint, C.POINTER(C.c_void_p)]
    L.pcgdo_tcm_destroy.argtypes = [C.c_void_p, C.c_void_p]
    L.pcgdo_tcm_create_affine.argtypes = [C.c_void_p, _u32p, C.c_size_t, C.c_uint, C.c_int, C.POINTER(C.c_void_p)]
    L.pcgdo_align2d_affine_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Batch), C.c_void_p]
    L.pcgdo_align2d_affine_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Batch)]
    L.pcgdo_align2d_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Batch), C.c_void_p]
    L.pcgdo_align2d_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Batch)]
    L.pcgdo_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.pcgdo_int32_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.pcgdo_metric_batch_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Batch32), C.c_void_p]
    L.pcgdo_metric_batch_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Batch32)]
    L.pcgdo_forest_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                      C.POINTER(C.c_void_p)]
    L.pcgdo_forest_destroy.argtypes = [C.c_void_p, C.c_void_p]
    L.pcgdo_forest_create32.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.pcgdo_forest_set_leaves32.argtypes = [C.c_void_p] * 5
    L.pcgdo_forest_set_leaves.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pcgdo_forest_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pcgdo_forest_tree_cost_ptr.argtypes = [C.c_void_p]
    L.pcgdo_forest_tree_cost_ptr.restype = C.c_void_p
    L.pcgdo_forest_tree_costs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.pcgdo_forest_copy_tree_costs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pcgdo_forest_get_node.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.pcgdo_forest_levels.argtypes = [C.c_void_p]
    L.pcgdo_forest_levels.restype = C.c_uint32
    L.pcgdo_forest_alignments.argtypes = [C.c_void_p]
    L.pcgdo_forest_alignments.restype = C.c_uint64
    L.pcgdo_forest_enable_rerooting.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.pcgdo_forest_try_leaf.argtypes = [C.c_void_p] * 8
    L.pcgdo_forest_score_rerooted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pcgdo_forest_get_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pcgdo_forest_edges.argtypes = [C.c_void_p]
    L.pcgdo_forest_edges.restype = C.c_uint32
    L.pcgdo_memo_create.argtypes = [C.c_void_p, _u32p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.pcgdo_memo_create_metric.argtypes = [C.c_void_p, _u32p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
    L.pcgdo_memo_structure.argtypes = [C.c_void_p]
    L.pcgdo_memo_alphabet.argtypes = [C.c_void_p]
    L.pcgdo_memo_alphabet.restype = C.c_uint32
    L.pcgdo_memo_destroy.argtypes = [C.c_void_p, C.c_void_p]
    L.pcgdo_overlap2d_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t] + [C.c_void_p] * 5
    L.pcgdo_overlap2d_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t] + [C.c_void_p] * 4
    L.pcgdo_explicit_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Batch32), C.c_void_p]
    L.pcgdo_explicit_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Batch32)]
    L.matrixInit.argtypes = [C.c_size_t, _u32p]
    L.matrixInit.restype = C.c_void_p
    L.matrixDestroy.argtypes = [C.c_void_p]
    L.matrixDestroy.restype = None
    L.getCostAndMedian2D.argtypes = [C.POINTER(DcElement)] * 3 + [C.c_void_p]
    L.getCostAndMedian2D.restype = C.c_uint
    L.setUp2dCostMtx.argtypes = [C.POINTER(CostMatrices2D), _u32p, C.c_size_t, C.c_uint]
    L.setUp2dCostMtx.restype = None
    L.setUp3dCostMtx.argtypes = [C.POINTER(CostMatrices3D), _u32p, C.c_size_t, C.c_uint]
    L.setUp3dCostMtx.restype = None
    L.freeCostMtx.argtypes = [C.c_void_p, C.c_int]
    L.freeCostMtx.restype = None
    L.getCostAndMedian3D.argtypes = [C.POINTER(DcElement)] * 4 + [C.c_void_p]
    L.getCostAndMedian3D.restype = C.c_uint
    L.pcgdo_configure_env.argtypes = [C.c_void_p]
    L.pcgdo_forest_set_trees.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.pcgdo_forest_cells.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.pcgdo_preorder_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]
    L.pcgdo_preorder_destroy.argtypes = [C.c_void_p, C.c_void_p]
    L.pcgdo_preorder_destroy.restype = None
    L.pcgdo_preorder_set_contexts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                              C.c_void_p, C.c_void_p]
    L.pcgdo_preorder_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.pcgdo_preorder_length.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.pcgdo_preorder_length.restype = C.c_uint32
    L.pcgdo_preorder_get.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32] + [C.c_void_p] * 4 + \
                                    [C.POINTER(C.c_int)]
    L.pcgdo_preorder_device_ptr.argtypes = [C.c_void_p]
    L.pcgdo_preorder_device_ptr.restype = C.c_void_p
    L.pcgdo_multi_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]
    L.pcgdo_multi_destroy.argtypes = [C.c_void_p]
    L.pcgdo_multi_destroy.restype = None
    L.pcgdo_multi_devices.argtypes = [C.c_void_p]
    L.pcgdo_multi_ctx.argtypes = [C.c_void_p, C.c_int]
    L.pcgdo_multi_ctx.restype = C.c_void_p
    L.pcgdo_multi_last_error.argtypes = [C.c_void_p]
    L.pcgdo_multi_last_error.restype = C.c_char_p
    L.pcgdo_multi_configure.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.pcgdo_multi_tcm_create.argtypes = [C.c_void_p, _u32p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.pcgdo_multi_tcm_destroy.argtypes = [C.c_void_p, C.c_void_p]
    L.pcgdo_multi_tcm_destroy.restype = None
    L.pcgdo_mforest_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                       C.POINTER(C.c_void_p)]
    L.pcgdo_mforest_destroy.argtypes = [C.c_void_p, C.c_void_p]
    L.pcgdo_mforest_destroy.restype = None
    L.pcgdo_mforest_set_leaves.argtypes = [C.c_void_p] * 5
    L.pcgdo_mforest_set_trees.argtypes = [C.c_void_p] * 3
    L.pcgdo_mforest_enable_rerooting.argtypes = [C.c_void_p] * 2
    L.pcgdo_mforest_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_double)]
    L.pcgdo_mforest_block.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    L.pcgdo_multi_align2d_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Batch)]
    L.pcgdo_overlap_sets_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t] + [C.c_void_p] * 6
    L.pcgdo_overlap_sets_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t] + [C.c_void_p] * 5
    L.align2d.argtypes = [C.POINTER(AlignIO)] * 4 + [C.POINTER(CostMatrices2D), C.c_int, C.c_int, C.c_int]
    L.align2d.restype = C.c_int
    L.align2dAffine.argtypes = [C.POINTER(AlignIO)] * 4 + [C.POINTER(CostMatrices2D), C.c_int]
    L.align2dAffine.restype = C.c_int
    _lib = L
    return L


class Context:
    """One GPU context (pcgdo_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.pcgdo_create(C.byref(h), device)
        if rc != 0 or not h.value:
            raise BackendError(f"pcgdo_create(device={device}) failed with code {rc}: "
                               "a CUDA device is required (no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.pcgdo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, what: str):
        if rc != 0:
            raise BackendError(f"{what} failed ({rc}): {self.lib.pcgdo_last_error(self.h).decode()}")

    def configure(self, pairs_per_warp=0, warps_per_cta=0, max_len=0, max_b=0, arena_words=0):
        self.check(self.lib.pcgdo_configure(self.h, pairs_per_warp, warps_per_cta, max_len), "pcgdo_configure")
        self.check(self.lib.pcgdo_configure_ex(self.h, max_b, arena_words), "pcgdo_configure_ex")

    def reload_env(self):
        """Re-read the PCGDO_* experiment switches from the environment (they are otherwise read once, at creation)."""
        self.check(self.lib.pcgdo_configure_env(self.h), "pcgdo_configure_env")

    def pinned_empty(self, count: int, dtype=np.uint8) -> np.ndarray:
        """A numpy array over pinned, mapped host memory (pcgdo_host_alloc); freed with the array."""
        import weakref
        nbytes = int(count) * np.dtype(dtype).itemsize
        p = self.lib.pcgdo_host_alloc(self.h, nbytes)
        if not p:
            raise BackendError(f"pcgdo_host_alloc({nbytes}) failed: {self.lib.pcgdo_last_error(self.h).decode()}")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype, count=int(count))
        lib, h = self.lib, self.h
        weakref.finalize(buf, lambda: lib.pcgdo_host_free(h, p))
        return arr

    def tcm(self, sigma, gap_open: int = 0, affine_variant: int = 0) -> "Tcm":
        return Tcm(self, sigma, gap_open, affine_variant)

    def memo(self, sigma, structure="explicit", a=None) -> "Memo":
        return Memo(self, sigma, structure, a)

    def int32_peak(self, mode: int = 0) -> float:
        """Measured integer lane-ops/s of this GPU (roofline denominator for the DP fill)."""
        v = C.c_double(0)
        self.check(self.lib.pcgdo_int32_peak(self.h, mode, C.byref(v), None), "pcgdo_int32_peak")
        return float(v.value)

    def last_kernel_ms(self):
        ms, n = C.c_float(0), C.c_int(0)
        self.check(self.lib.pcgdo_last_kernel_ms(self.h, C.byref(ms), C.byref(n)), "pcgdo_last_kernel_ms")
        return float(ms.value), int(n.value)


class Tcm:
    def __init__(self, ctx: Context, sigma, gap_open: int = 0, affine_variant: int = 0):
        self.ctx = ctx
        self.gap_open = gap_open
        sigma = np.ascontiguousarray(np.asarray(sigma, dtype=np.uint32))
        self.a = int(sigma.shape[0])
        assert sigma.shape == (self.a, self.a)
        self.gap = 1 << (self.a - 1)
        h = C.c_void_p()
        if gap_open:
            ctx.check(ctx.lib.pcgdo_tcm_create_affine(ctx.h, sigma.ctypes.data_as(_u32p), self.a, gap_open,
                                                      affine_variant, C.byref(h)), "pcgdo_tcm_create_affine")
        else:
            ctx.check(ctx.lib.pcgdo_tcm_create(ctx.h, sigma.ctypes.data_as(_u32p), self.a, gap_open, C.byref(h)),
                      "pcgdo_tcm_create")
        self.h = h
        # dense host tables from the library's own host code (setUp2dCostMtx), for the wrapper's
        # "one string empty" case and for callers that want lookupPairwise
        cm = CostMatrices2D()
        ctx.lib.setUp2dCostMtx(C.byref(cm), sigma.ctypes.data_as(_u32p), self.a, gap_open)
        dim = 1 << self.a
        self.cost = np.ctypeslib.as_array(cm.cost, shape=(dim, dim)).copy()
        self.median = np.ctypeslib.as_array(cm.median, shape=(dim, dim)).copy()
        self.median_of_gap = self.median[:, self.gap].astype(np.uint8)
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        for f in ("cost", "median", "worst", "prepend_cost", "tail_cost"):
            libc.free(C.cast(getattr(cm, f), C.c_void_p))

    def __del__(self):
        try:
            if self.h.value and self.ctx.h.value:
                self.ctx.lib.pcgdo_tcm_destroy(self.ctx.h, self.h)
        except Exception:
            pass


METRIC_AUTO, METRIC_EXPLICIT, METRIC_DISCRETE, METRIC_L1 = -1, 0, 1, 2
MISSING = 0xffffffff      # PCGDO_MISSING: the length of a Missing dynamic character
_STRUCTURES = {"auto": METRIC_AUTO, "explicit": METRIC_EXPLICIT, "discrete": METRIC_DISCRETE, "nonadditive": METRIC_DISCRETE,
               "l1": METRIC_L1, "additive": METRIC_L1}


class Memo:
    """Overlap function of the a > 8 path (pcgdo_memo): the device-side replacement of the reference's memoized TCM
    (lib/tcm-memo) for an explicit matrix — sigma[s, j]: median symbol s, input symbol j — or one of the closed forms PCG
    picks after diagnosing the matrix (``structure``: "explicit", "discrete", "l1", or "auto" = diagnose sigma like
    ``dynamicMetadata``).  Alphabets of up to 512 symbols; beyond 32 an element is ``words`` = ceil(a / 32) uint32."""

    def __init__(self, ctx: Context, sigma, structure="explicit", a=None):
        self.ctx = ctx
        if sigma is not None:
            sigma = np.ascontiguousarray(np.asarray(sigma, dtype=np.uint32))
            self.a = int(sigma.shape[0])
            assert sigma.shape == (self.a, self.a)
        else:
            self.a = int(a)
        self.words = (self.a + 31) // 32
        self.gap = 1 << (self.a - 1)
        h = C.c_void_p()
        ctx.check(ctx.lib.pcgdo_memo_create_metric(ctx.h, sigma.ctypes.data_as(_u32p) if sigma is not None else None, self.a,
                                                   _STRUCTURES[structure] if isinstance(structure, str) else int(structure),
                                                   C.byref(h)), "pcgdo_memo_create_metric")
        self.h = h
        self.structure = int(ctx.lib.pcgdo_memo_structure(h))

    def overlap(self, e1, e2):
        """``getMedianAndCost2D`` over arrays of ambiguity sets: (median[uint32], cost[uint32])."""
        e1 = np.ascontiguousarray(np.asarray(e1, np.uint32))
        e2 = np.ascontiguousarray(np.asarray(e2, np.uint32))
        assert e1.shape == e2.shape
        med, cost = np.zeros(e1.size, np.uint32), np.zeros(e1.size, np.uint32)
        self.ctx.check(self.ctx.lib.pcgdo_overlap2d_batch_host(self.ctx.h, self.h, e1.size, e1.ctypes.data, e2.ctypes.data,
                                                               med.ctypes.data, cost.ctypes.data),
                       "pcgdo_overlap2d_batch_host")
        return med.reshape(e1.shape), cost.reshape(e1.shape)

    def overlap_sets(self, *sets):
        """Sankoff median / cost over two or three arrays of ambiguity sets, any alphabet up to 512 symbols: each array is
        (n, words) uint32.  Returns (median (n, words), cost (n,)) — getCostAndMedian2D / getCostAndMedian3D in bulk."""
        assert len(sets) in (2, 3)
        arrs = [np.ascontiguousarray(np.asarray(e, np.uint32).reshape(-1, self.words)) for e in sets]
        n = arrs[0].shape[0]
        assert all(a.shape == arrs[0].shape for a in arrs)
        med, cost = np.zeros((n, self.words), np.uint32), np.zeros(n, np.uint32)
        self.ctx.check(self.ctx.lib.pcgdo_overlap_sets_batch_host(
            self.ctx.h, self.h, len(arrs), n, arrs[0].ctypes.data, arrs[1].ctypes.data,
            arrs[2].ctypes.data if len(arrs) == 3 else None, med.ctypes.data, cost.ctypes.data), "pcgdo_overlap_sets_batch_host")
        return med, cost

    def __del__(self):
        try:
            if self.h.value and self.ctx.h.value:
                self.ctx.lib.pcgdo_memo_destroy(self.ctx.h, self.h)
        except Exception:
            pass


class Forest:
    """Candidate trees x dynamic characters resident on one GPU (pcgdo_forest): post-order scoring.

    ``children``: int32 array (n_trees, n_taxa-1, 2) of child node ids (leaves 0..n_taxa-1, internal node x
    has id n_taxa + x).  ``leaves``: list over taxa of lists over characters of uint8 element strings."""

    def __init__(self, ctx: Context, children, n_taxa: int, leaves, node_cap: int, metric=None, dialect: int = 2):
        """``metric``: None = 8-bit elements under a dense TCM (passed to ``score``); an int = 32-bit elements under the
        discrete metric over that many symbols; a ``Memo`` = 32-bit elements under its explicit matrix (large
        alphabets: every alignment by Haskell dialect ``dialect``, default ``unboxedUkkonenFullSpaceDO``)."""
        self.ctx = ctx
        self.metric = metric
        self.dtype = np.uint8 if metric is None else np.uint32
        # words per element: large alphabets beyond 32 symbols use (n, words) uint32 strings
        self.words = 1 if metric is None else ((int(metric) if isinstance(metric, int) else metric.a) + 31) // 32
        ch = np.ascontiguousarray(np.asarray(children, dtype=np.int32))
        assert ch.ndim == 3 and ch.shape[1] == n_taxa - 1 and ch.shape[2] == 2
        self.n_trees, self.n_taxa = int(ch.shape[0]), int(n_taxa)
        self.n_chars = len(leaves[0])
        self.node_cap = (int(node_cap) + 3) & ~3
        h = C.c_void_p()
        if metric is None:
            ctx.check(ctx.lib.pcgdo_forest_create(ctx.h, self.n_trees, self.n_taxa, self.n_chars, self.node_cap,
                                                  ch.ctypes.data, C.byref(h)), "pcgdo_forest_create")
        else:
            memo_h, alpha = (None, int(metric)) if isinstance(metric, int) else (metric.h, metric.a)
            ctx.check(ctx.lib.pcgdo_forest_create32(ctx.h, self.n_trees, self.n_taxa, self.n_chars, self.node_cap,
                                                    ch.ctypes.data, memo_h, alpha, dialect, C.byref(h)),
                      "pcgdo_forest_create32")
        self.h = h
        # a leaf string of None is a Missing character (PCGDO_MISSING)
        raw = [s for taxon in leaves for s in taxon]
        flat = [np.zeros((0, self.words), self.dtype) if s is None else np.asarray(s, self.dtype).reshape(-1, self.words)
                for s in raw]
        real = np.array([len(s) for s in flat], np.uint64)
        ln = np.array([MISSING if r is None else len(s) for r, s in zip(raw, flat)], np.uint32)
        off = np.concatenate([[0], np.cumsum(real)[:-1]]).astype(np.uint64)
        elems = np.concatenate(flat) if flat else np.zeros((0, self.words), self.dtype)
        elems = np.ascontiguousarray(elems.reshape(-1))
        fn = ctx.lib.pcgdo_forest_set_leaves if metric is None else ctx.lib.pcgdo_forest_set_leaves32
        ctx.check(fn(ctx.h, self.h, elems.ctypes.data, off.ctypes.data, ln.ctypes.data), "pcgdo_forest_set_leaves")

    def score(self, tcm: Tcm, stream: int = 0) -> np.ndarray:
        self.ctx.check(self.ctx.lib.pcgdo_forest_score(self.ctx.h, tcm.h if tcm is not None else None, self.h, C.c_void_p(stream)),
                       "pcgdo_forest_score")
        out = np.zeros(self.n_trees, np.int64)
        self.ctx.check(self.ctx.lib.pcgdo_forest_tree_costs(self.ctx.h, self.h, out.ctypes.data),
                       "pcgdo_forest_tree_costs")
        return out

    def score_async(self, tcm: Tcm, stream: int = 0) -> int:
        """Sweep only; returns the DEVICE pointer of the int64[n_trees] cost vector."""
        self.ctx.check(self.ctx.lib.pcgdo_forest_score(self.ctx.h, tcm.h if tcm is not None else None, self.h, C.c_void_p(stream)),
                       "pcgdo_forest_score")
        return int(self.ctx.lib.pcgdo_forest_tree_cost_ptr(self.h))

    def score_into(self, tcm: Tcm, dst_device_ptr: int, stream: int = 0) -> None:
        """Sweep on `stream`, then copy the int64[n_trees] costs into device memory at dst_device_ptr."""
        self.ctx.check(self.ctx.lib.pcgdo_forest_score(self.ctx.h, tcm.h if tcm is not None else None, self.h, C.c_void_p(stream)),
                       "pcgdo_forest_score")
        self.ctx.check(self.ctx.lib.pcgdo_forest_copy_tree_costs(self.ctx.h, self.h, C.c_void_p(dst_device_ptr),
                                                                 C.c_void_p(stream)), "pcgdo_forest_copy_tree_costs")

    def enable_rerooting(self, keep_edge_data: bool = False) -> None:
        self.ctx.check(self.ctx.lib.pcgdo_forest_enable_rerooting(self.ctx.h, self.h, int(keep_edge_data)),
                       "pcgdo_forest_enable_rerooting")

    def try_leaf(self, tcm: Tcm, leaf_strings, stream: int = 0) -> np.ndarray:
        """Wagner edge trial (``iterativeBuild.tryEdge``): cost increase of attaching a new taxon (one element string
        per character) to every edge of every tree -> int64 (n_trees, n_edges).  Needs a re-rooted sweep with
        ``enable_rerooting(keep_edge_data=True)``."""
        raw = list(leaf_strings)
        flat = [np.zeros((0, self.words), self.dtype) if x is None else np.asarray(x, self.dtype).reshape(-1, self.words)
                for x in raw]
        real = np.array([len(x) for x in flat], np.uint64)
        ln = np.array([MISSING if r is None else len(x) for r, x in zip(raw, flat)], np.uint32)
        off = np.concatenate([[0], np.cumsum(real)[:-1]]).astype(np.uint64)
        elems = np.ascontiguousarray(np.concatenate(flat).reshape(-1)) if flat else np.zeros(1, self.dtype)
        ne = int(self.ctx.lib.pcgdo_forest_edges(self.h))
        out = np.zeros((self.n_trees, ne), np.int64)
        self.ctx.check(self.ctx.lib.pcgdo_forest_try_leaf(self.ctx.h, tcm.h if tcm is not None else None, self.h, elems.ctypes.data, off.ctypes.data,
                                                          ln.ctypes.data, out.ctypes.data, C.c_void_p(stream)),
                       "pcgdo_forest_try_leaf")
        return out

    def score_rerooted(self, tcm: Tcm, stream: int = 0) -> np.ndarray:
        """Post-order + re-rooting: per-tree sum over characters of the cheapest root edge."""
        self.ctx.check(self.ctx.lib.pcgdo_forest_score_rerooted(self.ctx.h, tcm.h if tcm is not None else None, self.h, C.c_void_p(stream)),
                       "pcgdo_forest_score_rerooted")
        out = np.zeros(self.n_trees, np.int64)
        self.ctx.check(self.ctx.lib.pcgdo_forest_tree_costs(self.ctx.h, self.h, out.ctypes.data),
                       "pcgdo_forest_tree_costs")
        return out

    def score_rerooted_into(self, tcm: Tcm, dst_device_ptr: int, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.pcgdo_forest_score_rerooted(self.ctx.h, tcm.h if tcm is not None else None, self.h, C.c_void_p(stream)),
                       "pcgdo_forest_score_rerooted")
        self.ctx.check(self.ctx.lib.pcgdo_forest_copy_tree_costs(self.ctx.h, self.h, C.c_void_p(dst_device_ptr),
                                                                 C.c_void_p(stream)), "pcgdo_forest_copy_tree_costs")

    def edges(self, tree: int, character: int):
        """((u, v) per unrooted edge, total cost of rooting there, local cost); edge 0 = the original root edge."""
        ne = int(self.ctx.lib.pcgdo_forest_edges(self.h))
        uv, tot, loc = np.zeros((ne, 2), np.int32), np.zeros(ne, np.int64), np.zeros(ne, np.int32)
        self.ctx.check(self.ctx.lib.pcgdo_forest_get_edges(self.ctx.h, self.h, tree, character, uv.ctypes.data,
                                                           tot.ctypes.data, loc.ctypes.data), "pcgdo_forest_get_edges")
        return uv, tot, loc

    def node(self, tree: int, character: int, node: int):
        buf = np.zeros(self.node_cap * self.words, self.dtype)
        n, lc, tc = C.c_uint32(0), C.c_int32(0), C.c_int64(0)
        self.ctx.check(self.ctx.lib.pcgdo_forest_get_node(self.ctx.h, self.h, tree, character, node, buf.ctypes.data,
                                                          C.byref(n), C.byref(lc), C.byref(tc)),
                       "pcgdo_forest_get_node")
        if n.value == MISSING:
            return None, int(lc.value), int(tc.value)
        med = buf[:n.value * self.words].copy()
        return (med if self.words == 1 else med.reshape(-1, self.words)), int(lc.value), int(tc.value)

    def set_trees(self, children) -> None:
        """Replace all candidate topologies over the resident leaves (``pcgdo_forest_set_trees``): levelled on the
        device into the buffers allocated at creation."""
        ch = np.ascontiguousarray(np.asarray(children, dtype=np.int32))
        assert ch.shape == (self.n_trees, self.n_taxa - 1, 2)
        self.ctx.check(self.ctx.lib.pcgdo_forest_set_trees(self.ctx.h, self.h, ch.ctypes.data), "pcgdo_forest_set_trees")

    def cells(self) -> int:
        """DP cells of the reference's band over the alignments of the last post-order sweep."""
        v = C.c_uint64(0)
        self.ctx.check(self.ctx.lib.pcgdo_forest_cells(self.ctx.h, self.h, C.byref(v)), "pcgdo_forest_cells")
        return int(v.value)

    @property
    def levels(self) -> int:
        return int(self.ctx.lib.pcgdo_forest_levels(self.h))

    @property
    def alignments(self) -> int:
        return int(self.ctx.lib.pcgdo_forest_alignments(self.h))

    def close(self):
        if getattr(self, "h", None) and self.h.value and self.ctx.h.value:
            self.ctx.lib.pcgdo_forest_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


REDUCE_HOST, REDUCE_PEER, REDUCE_NCCL = 0, 1, 2


class Preorder:
    """Pre-order pass on the device (``pcgdo_preorder_*``, do_preorder.cu): final alignments of every node of n_trees rooted
    trees x n_chars characters from their post-order contexts.  ``contexts[t][c]``: dict / sequence node -> (n, 3) uint8
    array (median, left, right) or None (Missing); leaves included ((s, s, s))."""

    def __init__(self, ctx: Context, children, n_taxa: int, n_chars: int):
        ch = np.ascontiguousarray(np.asarray(children, dtype=np.int32))
        if ch.ndim == 2:
            ch = ch[None]
        assert ch.shape[1:] == (n_taxa - 1, 2)
        self.ctx, self.n_trees, self.n_taxa, self.n_chars = ctx, ch.shape[0], n_taxa, n_chars
        self.n_nodes = 2 * n_taxa - 1
        self.h = C.c_void_p()
        ctx.check(ctx.lib.pcgdo_preorder_create(ctx.h, self.n_trees, n_taxa, n_chars, ch.ctypes.data, C.byref(self.h)),
                  "pcgdo_preorder_create")

    def set_contexts(self, contexts) -> None:
        n = self.n_trees * self.n_chars * self.n_nodes
        off, ln = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
        parts, at, i = [], 0, 0
        for t in range(self.n_trees):
            for c in range(self.n_chars):
                col = contexts[t][c]
                for v in range(self.n_nodes):
                    x = col[v]
                    if x is None:
                        ln[i] = MISSING
                    else:
                        x = np.ascontiguousarray(np.asarray(x, np.uint8).reshape(-1, 3))
                        off[i], ln[i] = at, len(x)
                        parts.append(x)
                        at += len(x)
                    i += 1
        allc = np.concatenate(parts, axis=0) if parts else np.zeros((0, 3), np.uint8)
        m, l, r = (np.ascontiguousarray(allc[:, k]) for k in range(3))
        self.ctx.check(self.ctx.lib.pcgdo_preorder_set_contexts(self.ctx.h, self.h, m.ctypes.data, l.ctypes.data, r.ctypes.data,
                                                                at, off.ctypes.data, ln.ctypes.data),
                       "pcgdo_preorder_set_contexts")

    def run(self, alphabet: int, stream: int = 0) -> None:
        self.ctx.check(self.ctx.lib.pcgdo_preorder_run(self.ctx.h, self.h, alphabet, C.c_void_p(stream)), "pcgdo_preorder_run")

    def node(self, tree: int, character: int, node: int):
        """((n, 3) final alignment, (n,) single disambiguation), or (None, None) for a Missing character."""
        n = int(self.ctx.lib.pcgdo_preorder_length(self.h, tree, character))
        if n == MISSING:
            return None, None
        bufs = [np.zeros(n, np.uint8) for _ in range(4)]
        miss = C.c_int(0)
        self.ctx.check(self.ctx.lib.pcgdo_preorder_get(self.ctx.h, self.h, tree, character, node, *[b.ctypes.data for b in bufs],
                                                       C.byref(miss)), "pcgdo_preorder_get")
        if miss.value:
            return None, None
        return np.stack(bufs[:3], axis=1), bufs[3]

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.ctx.lib.pcgdo_preorder_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Multi:
    """Several GPUs driven from this one process (``pcgdo_multi``): forests sharded by character blocks, pair batches
    split evenly; one host thread per device inside the library."""

    def __init__(self, devices):
        self.lib = load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.pcgdo_multi_create(C.byref(h), devs, len(devices))
        if rc != 0 or not h.value:
            raise BackendError(f"pcgdo_multi_create({list(devices)}) failed with code {rc} (no CPU fallback)")
        self.h = h
        self.n = len(devices)

    def check(self, rc, what):
        if rc != 0:
            raise BackendError(f"{what} failed ({rc}): {self.lib.pcgdo_multi_last_error(self.h).decode()}")

    def configure(self, pairs_per_warp=0, warps_per_cta=0, max_len=0, max_b=0):
        self.check(self.lib.pcgdo_multi_configure(self.h, pairs_per_warp, warps_per_cta, max_len, max_b), "pcgdo_multi_configure")

    def tcm(self, sigma):
        sigma = np.ascontiguousarray(np.asarray(sigma, dtype=np.uint32))
        h = C.c_void_p()
        self.check(self.lib.pcgdo_multi_tcm_create(self.h, sigma.ctypes.data_as(_u32p), int(sigma.shape[0]), C.byref(h)),
                   "pcgdo_multi_tcm_create")
        return h

    def forest(self, children, n_taxa, leaves, node_cap):
        """``leaves``: uint8 array (n_taxa, n_chars, L) of equal-length leaf strings, or a list over taxa of lists over
        characters.  Returns an opaque handle for ``score`` / ``set_trees``."""
        ch = np.ascontiguousarray(np.asarray(children, dtype=np.int32))
        n_trees = int(ch.shape[0])
        flat = [np.asarray(s, np.uint8) for taxon in leaves for s in taxon]
        n_chars = len(leaves[0])
        ln = np.array([len(s) for s in flat], np.uint32)
        off = np.concatenate([[0], np.cumsum(ln.astype(np.uint64))[:-1]]).astype(np.uint64)
        elems = np.ascontiguousarray(np.concatenate(flat)) if flat else np.zeros(1, np.uint8)
        h = C.c_void_p()
        self.check(self.lib.pcgdo_mforest_create(self.h, n_trees, int(n_taxa), n_chars, int(node_cap), ch.ctypes.data, C.byref(h)),
                   "pcgdo_mforest_create")
        self.check(self.lib.pcgdo_mforest_set_leaves(self.h, h, elems.ctypes.data, off.ctypes.data, ln.ctypes.data),
                   "pcgdo_mforest_set_leaves")
        return h, n_trees

    def set_trees(self, forest, children):
        ch = np.ascontiguousarray(np.asarray(children, dtype=np.int32))
        self.check(self.lib.pcgdo_mforest_set_trees(self.h, forest[0], ch.ctypes.data), "pcgdo_mforest_set_trees")

    def enable_rerooting(self, forest):
        self.check(self.lib.pcgdo_mforest_enable_rerooting(self.h, forest[0]), "pcgdo_mforest_enable_rerooting")

    def score(self, tcm, forest, rerooted=False, reduce_mode=REDUCE_HOST):
        """(tree costs int64[n_trees], wall ms, slowest device's sweep ms, reduction ms)."""
        out = np.zeros(forest[1], np.int64)
        wall, sweep, red = C.c_double(0), C.c_float(0), C.c_double(0)
        self.check(self.lib.pcgdo_mforest_score(self.h, tcm, forest[0], int(rerooted), int(reduce_mode), out.ctypes.data,
                                                C.byref(wall), C.byref(sweep), C.byref(red)), "pcgdo_mforest_score")
        return out, float(wall.value), float(sweep.value), float(red.value)

    def block(self, forest, i):
        lo, hi, bt = C.c_uint32(0), C.c_uint32(0), C.c_int(0)
        self.check(self.lib.pcgdo_mforest_block(forest[0], i, C.byref(lo), C.byref(hi), C.byref(bt)), "pcgdo_mforest_block")
        return int(lo.value), int(hi.value), bool(bt.value)

    def destroy_forest(self, forest):
        self.lib.pcgdo_mforest_destroy(self.h, forest[0])

    def align2d_batch_host(self, tcm, batch):
        self.check(self.lib.pcgdo_multi_align2d_batch_host(self.h, tcm, C.byref(batch)), "pcgdo_multi_align2d_batch_host")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.pcgdo_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

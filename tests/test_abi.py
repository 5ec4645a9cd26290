"""CPU: the C-ABI library loads, exports every symbol include/pcg_do_b200.h declares, its host-side
code (dense TCM build) matches the oracle, and it fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.helpers import ROOT, tcm_matrix


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pcg_do_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", src, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import pcg_b200.backend as be
    lib = be.load()
    declared = _declared_symbols()
    assert "align2d" in declared and "pcgdo_align2d_batch_host" in declared and len(declared) >= 11
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(be.EXPORTED_SYMBOLS) == declared


@pytest.mark.parametrize("name,a", [("discrete", 5), ("l1", 5), ("weird", 5), ("discrete", 4), ("l1", 7)])
def test_setUp2dCostMtx_matches_oracle(oracle_mod, name, a):
    import pcg_b200.backend as be
    lib = be.load()
    sig = tcm_matrix(name, a)
    cm = be.CostMatrices2D()
    lib.setUp2dCostMtx(C.byref(cm), sig.ctypes.data_as(C.POINTER(C.c_uint32)), a, 0)
    dim = 1 << a
    o = oracle_mod.Oracle(sig)
    assert cm.alphSize == a and cm.costMatrixDimension == dim and cm.gap_char == 1 << (a - 1)
    assert cm.cost_model_type == 0
    assert np.array_equal(np.ctypeslib.as_array(cm.cost, shape=(dim, dim)), o.cost)
    assert np.array_equal(np.ctypeslib.as_array(cm.median, shape=(dim, dim)), o.median)
    gap = 1 << (a - 1)
    assert np.array_equal(np.ctypeslib.as_array(cm.tail_cost, shape=(dim,))[1:], o.cost[1:, gap])
    assert np.array_equal(np.ctypeslib.as_array(cm.prepend_cost, shape=(dim,))[1:], o.cost[gap, 1:])


def test_struct_layouts_match_reference_abi():
    import pcg_b200.backend as be
    # alignIO_t {elem_t*; size_t; size_t}; cost_matrices_2d_t: 13 fields (costMatrix.h:35-93)
    assert C.sizeof(be.AlignIO) == 24
    assert C.sizeof(be.CostMatrices2D) == 88
    assert be.CostMatrices2D.cost.offset == 48 and be.CostMatrices2D.median.offset == 56


def test_no_cpu_fallback_without_gpu():
    import torch
    import pcg_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pcg_b200.BackendError):
        pcg_b200.Context(0)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "pcg_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower(), os.path.join(dp, f)

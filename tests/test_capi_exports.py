"""CPU-only: the C-ABI library loads, exports every symbol include/spand_b200.h declares, and fails loudly
(no CPU fallback) when a compute entry point is called without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import spand_public_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "spand_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spand_[a-z_0-9A-Z]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = ctypes.CDLL(S.LIB_PATH)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), n
    assert sorted(S.EXPORTS) == names


def test_no_product_code_touches_the_oracle():
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "spand_public_b200")):
        if "_build" in dp:
            continue
        for fn in fns:
            if fn.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"oracle_lib|liboracle|spand_oracle|orc_[a-z]+\(", txt):
                    bad.append(os.path.join(dp, fn))
    assert bad == []


def test_generators_match_reference_fixture_format():
    A = S.neglapl(5, 2)
    assert A.shape == (25, 25) and A.nnz == 105
    assert abs(A - A.T).max() == 0 and A.diagonal().tolist() == [4.0] * 25
    B = S.aniso_convdiff(6)
    assert B.shape == (216, 216)
    assert abs(B - B.T).max() > 0  # non-symmetric
    d = B.diagonal()
    off = np.asarray(abs(B).sum(axis=1)).ravel() - abs(d)
    assert (d > 0).all() and (d >= off - 1e-12).all()  # M-matrix, weakly diagonally dominant rows
    r = S.random(5, 2019)
    assert np.allclose(r[:3], [-0.78447229, 0.0473689, 0.09126481], atol=1e-8)


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    A = S.neglapl(5, 2)
    t = S.Tree(3)
    t.partition(A)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        t.assemble(A)

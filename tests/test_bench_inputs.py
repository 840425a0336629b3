"""The CPU (reference) arm of bench.py generates its inputs with numpy / scipy only, so that its process maps nothing
of the product. They must be the product's inputs bit for bit."""
import numpy as np
import pytest

import bench
import spand_public_b200 as S


@pytest.mark.parametrize("n,d", [(6, 3), (9, 2), (5, 1)])
def test_numpy_laplacian_and_coordinates(n, d):
    A, B = S.neglapl(n, d), bench.np_neglapl(n, d)
    assert A.nnz == B.nnz and abs(A - B).max() == 0
    assert np.array_equal(S.linspace_nd(n, d), bench.np_linspace_nd(n, d))
    G, H = S.symmetric_graph(A), bench.np_symmetric_graph(B)
    assert np.array_equal(G.indptr, H.indptr) and np.array_equal(G.indices, H.indices)


def test_numpy_random_is_the_reference_stream():
    assert np.array_equal(S.random(1000, 2019), bench.np_random(1000, 2019))
    assert np.array_equal(S.random(17, 7), bench.np_random(17, 7))


def test_numpy_c5_family():
    A, B = S.aniso_convdiff(7), bench.np_aniso_convdiff(7)
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert abs(A - B).max() <= 1e-13 * abs(A).max()


def test_reference_arm_does_not_import_the_product():
    import ast
    import inspect
    src = inspect.getsource(bench.run_oracle)
    assert "spand_public_b200" not in src and "import S" not in src
    names = {n.id for n in ast.walk(ast.parse(src.strip())) if isinstance(n, ast.Name)}
    assert "S" not in names

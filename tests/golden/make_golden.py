"""Generates tests/golden/*.json. Run in the build container (needs /root/reference for the .mm fixtures).

partition_square.json  transcribes the known answer of reference tests/tests.cpp:378-412.
c1_oracle.json         is the oracle's own output on config C1 read from the reference's fixture files
                       mats/neglapl_2_32.mm + mats/32x32.mm (the reference holds no golden ranks/nnz; this file pins
                       that the generator path (neglapl, linspace_nd + 1) reproduces the fixture path exactly).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
import spand_public_b200 as S  # noqa: E402

REF = "/root/reference/mats"

sep = [(0, 0), (0, 0), (1, 0), (0, 1), (0, 1), (0, 0), (0, 0), (1, 0), (0, 1), (0, 1), (2, 0), (2, 0), (2, 0), (2, 0),
       (2, 0), (0, 2), (0, 2), (1, 1), (0, 3), (0, 3), (0, 2), (0, 2), (1, 1), (0, 3), (0, 3)]
left = [(0, 0), (0, 0), (0, 0), (0, 1), (0, 1), (0, 0), (0, 0), (0, 0), (0, 1), (0, 1), (0, 0), (0, 0), (1, 0), (0, 1),
        (0, 1), (0, 2), (0, 2), (0, 2), (0, 3), (0, 3), (0, 2), (0, 2), (0, 2), (0, 3), (0, 3)]
right = [(0, 0), (0, 0), (0, 1), (0, 1), (0, 1), (0, 0), (0, 0), (0, 1), (0, 1), (0, 1), (0, 2), (0, 2), (1, 1), (0, 3),
         (0, 3), (0, 2), (0, 2), (0, 3), (0, 3), (0, 3), (0, 2), (0, 2), (0, 3), (0, 3), (0, 3)]
json.dump({"self": [list(x) for x in sep], "l": [list(x) for x in left], "r": [list(x) for x in right]},
          open(os.path.join(HERE, "partition_square.json"), "w"))

A = S.mm_read(os.path.join(REF, "neglapl_2_32.mm"))
X = S.mm_read_dense(os.path.join(REF, "32x32.mm"))
assert abs(A - S.neglapl(32, 2)).max() == 0
assert np.array_equal(X, S.linspace_nd(32, 2)[::-1] + 1.0)  # the file stores x fastest
t = O.OracleTree(5, tol=1e-2)
t.set_coords(X)
t.partition(S.symmetric_graph(A))
t.assemble(A)
t.factorize()
lg = t.log()
it, _ = t.cg(A, S.random(1024, 2019), 100, 1e-12)
ids, size, rank = t.stats()
json.dump({"coords_first8": X[:, :8].tolist(), "dofs_left_elim": lg["dofs_left_elim"].astype(int).tolist(),
           "dofs_left_spars": lg["dofs_left_spars"].astype(int).tolist(), "nnz": int(t.nnz()), "cg": int(it),
           "size": size.tolist(), "rank": rank.tolist()}, open(os.path.join(HERE, "c1_oracle.json"), "w"))
print("golden written")

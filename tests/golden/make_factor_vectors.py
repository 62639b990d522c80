#!/usr/bin/env python
"""
Records golden factor vectors from the REFERENCE's generated headers (oracle/_ref/libref_factors.so,
built by oracle/Makefile from /root/reference): 8 random inputs per kind with res, J, H(lower), rhs.
Output: tests/golden/factor_vectors.json (committed; lets the factor parity test run without
/root/reference).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from symforce_b200 import desc as D  # noqa: E402
from tests import oracle_capi as O  # noqa: E402
from tests.test_oracle_factors import random_args  # noqa: E402

ref = O.load_ref()
assert ref is not None, "build oracle/_ref first (make -C oracle)"
out = {"source": "ref_eval_factor over /root/reference generated headers", "vectors": []}
for kind in range(len(D.KINDS)):
    rng = np.random.default_rng(777 + kind)
    for _ in range(8):
        args = random_args(kind, rng)
        r, J, H, g = O.eval_factor(ref, "ref_eval_factor", kind, args)
        out["vectors"].append(dict(kind=kind, args=[a.tolist() for a in args], res=r.tolist(),
                                   J=J.reshape(-1).tolist(), H=np.tril(H).reshape(-1).tolist(), rhs=g.tolist()))
with open(os.path.join(os.path.dirname(__file__), "factor_vectors.json"), "w") as f:
    json.dump(out, f)
print("wrote", len(out["vectors"]), "vectors")

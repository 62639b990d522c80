"""
Generates tests/golden/covariance_test_matrix.npz from the reference's own fixture of the marginal-covariance path,
/root/reference/test/test_data/covariance_test_matrix.bin (a 2000 x 2000 lower-triangular sparse Hessian in the layout
test/symforce_covariance_utils_test.cc:20-72 reads: int32 rows, cols, nnz; inner indices; outer indices; values).

The two cases of that test (:85-128) take the bottom-right and the top-left 1000 x 1000 corner; the fixture keeps those
two corners in CSC form (lower triangle) so the GPU box, which has no /root/reference, can replay the test.

Run in the build container:  python tests/golden/make_covariance_fixture.py
"""
import os
import struct

import numpy as np

SRC = "/root/reference/test/test_data/covariance_test_matrix.bin"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(path):
    b = open(path, "rb").read()
    rows, cols, nnz = struct.unpack("<iii", b[:12])
    o = 12
    inner = np.frombuffer(b, dtype="<i4", count=nnz, offset=o)
    o += 4 * nnz
    outer = np.frombuffer(b, dtype="<i4", count=cols + 1, offset=o)
    o += 4 * (cols + 1)
    values = np.frombuffer(b, dtype="<f8", count=nnz, offset=o)
    assert o + 8 * nnz == len(b)
    return rows, cols, outer, inner, values


def corner(outer, inner, values, r0, c0, n):
    """CSC of M[r0:r0+n, c0:c0+n] (Eigen block of a compressed column-major matrix)."""
    o, i, v = [0], [], []
    for c in range(c0, c0 + n):
        for q in range(outer[c], outer[c + 1]):
            if r0 <= inner[q] < r0 + n:
                i.append(inner[q] - r0)
                v.append(values[q])
        o.append(len(i))
    return np.array(o, dtype=np.int32), np.array(i, dtype=np.int32), np.array(v)


if __name__ == "__main__":
    rows, cols, outer, inner, values = load(SRC)
    assert rows == cols == 2000
    out = {}
    for name, off in (("bottom_right", 1000), ("top_left", 0)):
        o, i, v = corner(outer, inner, values, off, off, 1000)
        out[name + "_outer"], out[name + "_inner"], out[name + "_values"] = o, i, v
    np.savez_compressed(os.path.join(HERE, "covariance_test_matrix.npz"), **out)
    print({k: x.shape for k, x in out.items()})

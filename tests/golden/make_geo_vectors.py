"""
Generates tests/golden/geo_vectors.json from the reference's generated numeric geometry package
(/root/reference/gen/python/sym) -- run in the build container only; the JSON is what the tests read.

    PYTHONPATH=/root/reference/gen/python python tests/golden/make_geo_vectors.py
"""
import json
import os

import numpy as np
import sym

EPS = 10 * np.finfo(np.float64).eps


def main():
    rng = np.random.default_rng(0x6E0)
    cases = []
    for _ in range(40):
        v, w = rng.normal(size=6), rng.normal(size=6)
        ypr = rng.uniform(-3, 3, 3)
        a = sym.Pose3.from_tangent(v, EPS)
        b = sym.Pose3.from_tangent(w, EPS)
        pt = rng.normal(size=3)
        cases.append(dict(
            v=v.tolist(), w=w.tolist(), ypr=ypr.tolist(), pt=pt.tolist(),
            a=list(a.to_storage()), b=list(b.to_storage()),
            a_inv=list(a.inverse().to_storage()),
            ab=list((a * b).to_storage()),
            a_retract_w=list(a.retract(w, EPS).to_storage()),
            a_local_b=np.asarray(a.local_coordinates(b, EPS)).reshape(-1).tolist(),
            a_tangent=np.asarray(a.to_tangent(EPS)).reshape(-1).tolist(),
            a_pt=np.asarray(a * pt).reshape(-1).tolist(),
            rot_ypr=list(sym.Rot3.from_yaw_pitch_roll(*ypr).to_storage()),
            rot_matrix=np.asarray(a.rotation().to_rotation_matrix()).reshape(-1).tolist(),
            rot_local=np.asarray(a.rotation().local_coordinates(b.rotation(), EPS)).reshape(-1).tolist(),
        ))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geo_vectors.json")
    with open(out, "w") as f:
        json.dump(dict(epsilon=EPS, cases=cases), f)
    print("wrote", out, len(cases))


if __name__ == "__main__":
    main()

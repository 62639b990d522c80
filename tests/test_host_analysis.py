"""
CPU tests of the product's host logic (no GPU): the C-ABI library loads and exports every symbol of
include/sfx.h, and the structural analysis (index maps, CSC layout, Schur match lists, multifrontal
plan) replayed in numpy reproduces the oracle's H / rhs / residual / LM step.
"""
import re
import os

import numpy as np
import pytest

from symforce_b200 import capi, desc as D, problems as P
from tests import host_emulation as E
from tests import oracle_capi as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    header = open(os.path.join(ROOT, "include", "sfx.h")).read()
    declared = set(re.findall(r"\b(sfx_[a-z_0-9]+)\s*\(", header))
    declared -= {"sfx_status"}
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        capi.SfxProblem(P.pose_smoothing())


PROBLEMS = {
    "pose_smoothing": lambda: P.pose_smoothing(),
    "rotation_smoothing": lambda: P.rotation_smoothing(),
    "frozen_keys": lambda: P.frozen_keys(),
    "robot3d": lambda: P.robot_3d_localization(),
    "ba_example": lambda: P.ba_example(),
    "bal_tiny_schur": lambda: P.bal_problem("tiny", solver=D.SOLVER_SCHUR),
    "bal_tiny_chol": lambda: P.bal_problem("tiny", solver=D.SOLVER_CHOLESKY),
    "bal_tiny_natural": lambda: _with_ordering(P.bal_problem("tiny", solver=D.SOLVER_SCHUR), D.ORDERING_NATURAL),
    "bal_tiny_block": lambda: _with_ordering(P.bal_problem("tiny", solver=D.SOLVER_CHOLESKY), D.ORDERING_METIS_BLOCK),
    "pose_graph_small": lambda: P.pose_graph_problem(n_poses=60, n_loops=15),
}


def _with_ordering(p, o):
    p.ordering = o
    return p


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_host_maps_reproduce_oracle(name):
    prob = PROBLEMS[name]()
    A = capi.analysis_json(prob)
    o = O.OracleProblem(prob)
    N, M, nnz = o.dims()
    assert (A["N"], A["M"]) == (N, M)
    outer, inner = o.hessian_pattern()
    assert np.array_equal(outer, np.array(A["csc_outer"]))
    assert np.array_equal(inner, np.array(A["csc_inner"]))
    lam = 0.37
    upd, (res, rhs, H) = E.emulate_solve_step(prob, A, lam)
    res_o, rhs_o, H_o = o.linearize()
    scale = max(1.0, np.abs(H_o).max())
    assert np.allclose(res, res_o, rtol=0, atol=1e-12 * max(1.0, np.abs(res_o).max()))
    assert np.allclose(rhs, rhs_o, rtol=0, atol=1e-11 * max(1.0, np.abs(rhs_o).max()))
    assert np.allclose(H, H_o, rtol=0, atol=1e-12 * scale)
    upd_o = o.solve_step(lam)
    assert np.allclose(upd, upd_o, rtol=1e-8, atol=1e-9 * max(1e-3, np.abs(upd_o).max()))


@pytest.mark.parametrize("seed", range(6))
def test_host_maps_random_bal_shapes(seed, monkeypatch):
    """Ragged BAL structures (random camera / point counts and track lengths, windows wider than the camera ring) through
    the same host-vs-oracle replay, with the per-slot index building forced onto several host threads."""
    rng = np.random.default_rng(1000 + seed)
    n_cams = int(rng.integers(3, 12))
    n_pts = int(rng.integers(4 * n_cams, 70))
    window = int(rng.integers(1, n_cams + 3))
    slots = 2 * min(window, (n_cams - 1) // 2) + 1  # longest track P.bal_structure can draw
    n_obs = int(rng.integers(2 * n_pts, slots * n_pts + 1))
    monkeypatch.setenv("SFX_HOST_THREADS", str(int(rng.integers(1, 5))))
    solver = D.SOLVER_SCHUR if seed % 2 == 0 else D.SOLVER_CHOLESKY
    for structure_seed in range(seed, seed + 600, 100):  # every camera must be observed at least once
        prob = P.bal_problem(n_cams=n_cams, n_pts=n_pts, n_obs=n_obs, window=window, solver=solver,
                             seed_structure=structure_seed)
        if len(np.unique(prob.meta["cam"])) == n_cams:
            break
    else:
        pytest.fail("no structure seed observes every camera")
    A = capi.analysis_json(prob)
    o = O.OracleProblem(prob)
    outer, inner = o.hessian_pattern()
    assert np.array_equal(outer, np.array(A["csc_outer"])) and np.array_equal(inner, np.array(A["csc_inner"]))
    upd, (res, rhs, H) = E.emulate_solve_step(prob, A, 0.5)
    res_o, rhs_o, H_o = o.linearize()
    assert np.allclose(H, H_o, rtol=0, atol=1e-12 * max(1.0, np.abs(H_o).max()))
    assert np.allclose(rhs, rhs_o, rtol=0, atol=1e-11 * max(1.0, np.abs(rhs_o).max()))
    upd_o = o.solve_step(0.5)
    assert np.allclose(upd, upd_o, rtol=1e-8, atol=1e-9 * max(1e-3, np.abs(upd_o).max()))


@pytest.mark.parametrize("name", ["pose_smoothing", "frozen_keys", "robot3d", "ba_example", "bal_tiny_schur",
                                  "bal_tiny_chol", "pose_graph_small"])
def test_jacobian_index_maps_reproduce_oracle(name):
    """include_jacobians (linearizer.cc:252-259, 297-313): the CSC pattern of Linearization::jacobian built by the host
    analysis is the oracle's (triplets compressed like setFromTriplets), the per-slot scatter positions fill every entry
    exactly once, and J^T J / J^T r of the result are the Hessian and rhs."""
    import scipy.sparse as sp

    prob = PROBLEMS[name]()
    A = capi.analysis_json(prob)
    o = O.OracleProblem(prob)
    outer, inner, val = o.jacobian()
    assert np.array_equal(outer, np.array(A["jac_outer"]))
    assert np.array_equal(inner, np.array(A["jac_inner"]))
    got = E.emulate_jacobian(prob, A)
    assert not np.isnan(got).any()
    assert np.array_equal(got, val)
    N, M, _ = o.dims()
    J = sp.csc_matrix((got, inner, outer), shape=(M, N))
    res, rhs, Hv = o.linearize()
    ho, hi = o.hessian_pattern()
    H = sp.csc_matrix((Hv, hi, ho), shape=(N, N)).toarray()
    assert np.allclose(np.tril((J.T @ J).toarray()), H, rtol=0, atol=1e-12 * np.abs(H).max())
    assert np.allclose(J.T @ res, rhs, rtol=0, atol=1e-12 * max(1.0, np.abs(rhs).max()))


def test_bal_nodes_merge_pose_and_intrinsics():
    prob = P.bal_problem("tiny", solver=D.SOLVER_SCHUR)
    A = capi.analysis_json(prob)
    n_cams = prob.meta["n_cams"]
    kn = A["key_node"]
    for j in range(n_cams):
        assert kn[j] == kn[n_cams + j]  # c_j and i_j share a node
    assert A["H"]["node_dim"][: n_cams] == [9] * n_cams
    assert A["schur_plan"]["n_landmarks"] == prob.meta["n_pts"]


def test_unoptimized_key_is_an_error():
    prob = P.pose_smoothing()
    # add an optimized key no factor touches
    vals = np.concatenate([prob.values, [0, 0, 0, 1, 0, 0, 0]])
    keys = np.concatenate([prob.keys, [[D.TYPE_POSE3, len(prob.values), 7, 6]]])
    bad = D.Problem(vals, keys, prob.batches)
    with pytest.raises(RuntimeError, match="not optimized by any factor"):
        capi.analysis_json(bad)
    with pytest.raises(RuntimeError, match="not optimized by any factor"):
        O.OracleProblem(bad)


def test_tile_task_lists_are_dependency_ordered():
    """The tile-DAG task generator (sfx_api.cu: build_front_tasks) is replayed sequentially against the
    tile version counters for many front shapes and panel widths: every wait condition of
    chol_large.cu must hold when its task comes up and every tile must end up final."""
    from symforce_b200 import capi

    lib = capi.load()
    for kc in (1, 2, 3, 4, 6, 8):
        for wt in range(1, 14):
            for nt in range(wt, wt + 14):
                n = lib.sfx_debug_verify_tasks(wt, nt, kc)
                assert n > 0, (kc, wt, nt, lib.sfx_last_error(None).decode())
    # the range tasks cut the task count of the final-shape level-1 fronts by ~3x
    assert lib.sfx_debug_verify_tasks(18, 41, 4) < lib.sfx_debug_verify_tasks(18, 41, 1) // 2


def test_bal_text_file_round_trip(tmp_path):
    """problems.read_bal reads the BAL text format the way the reference example does
    (bundle_adjustment_in_the_large.cc:61-118); written from a synthetic problem and read back it gives the same
    factors and keys and -- up to the sign of the camera quaternions, which the Rodrigues vector does not keep --
    the same values, hence the same linearization in the oracle."""
    a = P.bal_problem("tiny", solver=D.SOLVER_SCHUR)
    path = str(tmp_path / "tiny.bal")
    P.write_bal(path, a)
    b = P.read_bal(path)
    assert np.array_equal(a.keys, b.keys) and b.schur_num_keys == a.schur_num_keys
    for x, y in zip(a.batches[0][1:], b.batches[0][1:]):
        assert np.array_equal(x, y)
    va, vb = a.values.copy(), b.values.copy()
    for v in (va, vb):
        q = v[a.meta["cam_off"]:a.meta["cam_off"] + 10 * a.meta["n_cams"]].reshape(-1, 10)
        q[:, :4] *= np.sign(q[:, 3:4])
    assert np.allclose(va, vb, rtol=0, atol=1e-14)
    ra, ga, Ha = O.OracleProblem(a).linearize()
    rb, gb, Hb = O.OracleProblem(b).linearize()
    assert np.allclose(ra, rb, rtol=0, atol=1e-10) and np.allclose(Ha, Hb, rtol=1e-12, atol=1e-9 * np.abs(Ha).max())
    # malformed files are rejected
    with open(path, "a") as f:
        f.write("1.0\n")
    with pytest.raises(ValueError, match="expected"):
        P.read_bal(path)
    bad = str(tmp_path / "bad.bal")
    with open(bad, "w") as f:
        f.write("1 1 1\n0 5 1.0 2.0\n" + "0 " * 9 + "\n0 0 0\n")
    with pytest.raises(ValueError, match="not in the file"):
        P.read_bal(bad)


def _random_pose_graph(seed):
    """Random Pose3 graph: a chain plus random extra between factors and priors, a random subset of poses held fixed,
    optimized keys in shuffled (not storage) order, factors of the two kinds interleaved in the caller's order."""
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(5, 40))
    poses = np.concatenate([P.quat_exp(rng.normal(0, 0.5, (n, 3))), rng.normal(0, 2.0, (n, 3))], axis=1)
    vb = P.ValuesBuilder()
    pose_off = vb.add_many(poses)
    eps_off = vb.add([D.K_DEFAULT_EPSILON])
    frozen = rng.random(n) < 0.25
    frozen[int(rng.integers(0, n))] = False
    edges = [(i, i + 1) for i in range(n - 1)]
    for _ in range(int(rng.integers(0, 2 * n))):
        i, j = (int(x) for x in rng.integers(0, n, 2))
        if i != j:
            edges.append((i, j))
    edges = [(i, j) for i, j in edges if not (frozen[i] and frozen[j])]
    touched = set(i for e in edges for i in e)
    priors = [i for i in range(n) if not frozen[i] and (i not in touched or rng.random() < 0.2)]
    opt = [i for i in rng.permutation(n) if not frozen[i]]
    key_of = {int(p): k for k, p in enumerate(opt)}
    keys = [(D.TYPE_POSE3, int(pose_off[p]), 7, 6) for p in opt]
    ne, npr = len(edges), len(priors)
    meas = np.concatenate([P.quat_exp(rng.normal(0, 0.3, (ne + npr, 3))), rng.normal(0, 1.0, (ne + npr, 3))], axis=1)
    meas_off = vb.add_many(meas)
    si_off = vb.add_many(np.stack([np.diag(rng.uniform(0.5, 3.0, 6)).reshape(-1, order="F") for _ in range(ne + npr)]))
    order = rng.permutation(ne + npr)  # caller's factor index of each factor
    between = (
        D.KIND_BETWEEN_POSE3,
        np.array([[pose_off[i] for i, _ in edges], [pose_off[j] for _, j in edges], meas_off[:ne], si_off[:ne],
                  np.full(ne, eps_off)]),
        np.array([[key_of.get(i, -1) for i, _ in edges], [key_of.get(j, -1) for _, j in edges]]),
        order[:ne],
    )
    batches = [between]
    if npr:
        batches.append((
            D.KIND_PRIOR_POSE3,
            np.array([[pose_off[i] for i in priors], meas_off[ne:], si_off[ne:], np.full(npr, eps_off)]),
            np.array([[key_of[i] for i in priors]]),
            order[ne:],
        ))
    ordering = [D.ORDERING_METIS_SCALAR, D.ORDERING_METIS_BLOCK, D.ORDERING_NATURAL][seed % 3]
    return D.Problem(vb.data(), keys, batches, ordering=ordering)


@pytest.mark.parametrize("seed", range(8))
def test_host_maps_random_pose_graphs(seed, monkeypatch):
    """Generic (non-BAL) path on random structures: fixed poses inside factors, shuffled key order, two factor kinds
    interleaved by factor index, all three orderings; Hessian, rhs, LM step and the Jacobian maps against the oracle."""
    monkeypatch.setenv("SFX_HOST_THREADS", str(1 + seed % 3))
    prob = _random_pose_graph(seed)
    A = capi.analysis_json(prob)
    o = O.OracleProblem(prob)
    outer, inner = o.hessian_pattern()
    assert np.array_equal(outer, np.array(A["csc_outer"])) and np.array_equal(inner, np.array(A["csc_inner"]))
    upd, (res, rhs, H) = E.emulate_solve_step(prob, A, 0.8)
    res_o, rhs_o, H_o = o.linearize()
    assert np.allclose(res, res_o, rtol=0, atol=1e-12 * max(1.0, np.abs(res_o).max()))
    assert np.allclose(H, H_o, rtol=0, atol=1e-12 * max(1.0, np.abs(H_o).max()))
    assert np.allclose(rhs, rhs_o, rtol=0, atol=1e-11 * max(1.0, np.abs(rhs_o).max()))
    upd_o = o.solve_step(0.8)
    assert np.allclose(upd, upd_o, rtol=1e-8, atol=1e-9 * max(1e-3, np.abs(upd_o).max()))
    jo, ji, jv = o.jacobian()
    assert np.array_equal(jo, np.array(A["jac_outer"])) and np.array_equal(ji, np.array(A["jac_inner"]))
    assert np.array_equal(E.emulate_jacobian(prob, A), jv)


@pytest.mark.parametrize("depth", [0, 1, 2, 3])
@pytest.mark.parametrize("name", ["bal_tiny_schur", "bal_tiny_chol", "pose_graph_small", "robot3d", "ring"])
def test_dissect_then_sweep_ordering_reproduces_oracle(name, depth, monkeypatch):
    """The ordering candidates of choose_front_plan (symbolic.cc: nested dissection to a fixed depth, then a
    Cuthill-McKee sweep of every subdomain; cumulative, width-capped relaxed amalgamation) through the host replay:
    the fronts, extend-add maps and assembly copies of such a plan must give the oracle's step."""
    monkeypatch.setenv("SFX_ND_DEPTH", str(depth))
    monkeypatch.setenv("SFX_RELAX", "0.10")
    monkeypatch.setenv("SFX_RELAX_CUM", "1")
    monkeypatch.setenv("SFX_MAX_MERGE_W", "36" if depth % 2 else "1024")
    if name == "ring":
        # a camera ring with a covisibility window, the structure of the synthetic Final-shape problem in small
        prob = P.bal_problem(n_cams=40, n_pts=400, n_obs=1400, window=3, solver=D.SOLVER_SCHUR)
    else:
        prob = PROBLEMS[name]()
    A = capi.analysis_json(prob)
    perm = A["fronts"]["perm_nodes"]
    assert sorted(perm) == list(range(len(perm)))
    assert sorted(A["fronts"]["scalar_perm"]) == list(range(A["fronts"]["n"]))
    o = O.OracleProblem(prob)
    upd, (res, rhs, H) = E.emulate_solve_step(prob, A, 0.37)
    upd_o = o.solve_step(0.37)
    assert np.allclose(upd, upd_o, rtol=1e-8, atol=1e-9 * max(1e-3, np.abs(upd_o).max()))


def test_plan_search_on_a_camera_ring(monkeypatch):
    """choose_front_plan (sfx_api.cu) on a 320-camera ring: the METIS_NodeND plan and the dissect-then-sweep candidates
    are all planned for 296 resident CTAs, their fused task lists pass verify_fused_list (every wait condition of the
    tile-DAG kernel holds in list order), and the search never keeps a plan the model rates slower than the reference's."""
    import ctypes as C
    import re

    prob = P.bal_problem(n_cams=320, n_pts=6000, n_obs=26000, window=12, solver=D.SOLVER_SCHUR)
    lib = capi.load()
    d, keep = prob.desc(rank=0, world=1, comm=None)
    res = (C.c_int64 * 6)()
    assert lib.sfx_debug_large_plan(C.byref(d), 296, res) == 0, lib.sfx_last_error(None).decode()
    fused_T0, n_large, n_tasks, n_fused = res[0], res[1], res[2], res[3]
    assert fused_T0 == 0 and n_large >= 3 and n_fused == n_tasks > 0
    monkeypatch.setenv("SFX_ORDERING_SEARCH", "0")
    res0 = (C.c_int64 * 6)()
    assert lib.sfx_debug_large_plan(C.byref(d), 296, res0) == 0
    assert res0[0] == 0 and res0[3] > 0


def test_exclusive_blocks_start_on_a_128_byte_boundary():
    """The landmark-camera blocks only one factor writes come last in H, in observation order, and start on a multiple
    of 16 doubles: the TMA-streamed kernels copy them in 27,648-byte tiles (cp.async.bulk needs 16-byte aligned sources)."""
    for prob in (P.bal_problem("tiny", solver=D.SOLVER_SCHUR), P.bal_problem("small", solver=D.SOLVER_SCHUR),
                 P.bal_problem(n_cams=7, n_pts=31, n_obs=100, window=3, solver=D.SOLVER_SCHUR)):
        A = capi.analysis_json(prob)
        assert A["h_accum_values"] % 16 == 0
        eoff = np.array(A["schur_plan"]["r_eoff"])
        assert eoff.min() == A["h_accum_values"]
        assert np.array_equal(np.sort(eoff), A["h_accum_values"] + 27 * np.arange(eoff.shape[0]))

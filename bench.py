#!/usr/bin/env python
"""
bench.py -- LM iterations/sec (linearize + Schur + Cholesky) on synthetic BAL, the metric of
BASELINE.json, measured on the CUDA path through the C ABI (symforce_b200/lib/libsfx.so).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload final|ladybug|pose_graph|...] [--impl reference]

One "step" = one Levenberg-Marquardt iteration (damp -> Schur -> factorize -> solve -> retract ->
relinearize -> accept/reject) of sym::Optimizer on the workload.
  value : K iterations inside ONE sfx_optimize call, inputs already resident in HBM, timed on the
          device with CUDA events on the library's stream.
  e2e   : K calls of the reference-facing API the way the reference's own benchmark drives it
          (`Optimize(values, 1, ...)`, symforce/benchmarks/robot_3d_localization/
          robot_3d_localization_benchmark.cc:88-93): per step a host->device copy of the Values
          buffer from pinned host memory, one LM iteration, and the device->host read of the result
          into the same buffer (Values::Update semantics: the optimized keys' storage).
  --impl reference : the reference's own CPU algorithm (oracle/, a restatement pinned to the
          reference's KATs -- the reference C++ cannot be built here), single
          thread like the reference, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "LM iterations/sec (linearize+Schur+Cholesky) on synthetic BAL"
UNIT = "iterations/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sfx", choices=["sfx", "reference"])
    ap.add_argument("--workload", default="final")
    ap.add_argument("--cpu-baseline", type=int, default=1, help="0: skip the cpu_baseline leg")
    ap.add_argument("--cpu-workload", default=None)
    ap.add_argument("--secondary", type=int, default=1, help="0: skip the ladybug / pose_graph secondary entries")
    return ap.parse_args()


POSE_GRAPH = dict(n_poses=100000, n_loops=20000)  # BASELINE.json configs[4]


def make_problem(name):
    """BAL shapes (Schur path) or `pose_graph` (config E: 100k Pose3 poses, SparseCholeskySolver path, 1 GPU)."""
    from symforce_b200 import desc as D, problems as P

    if name == "pose_graph":
        return P.pose_graph_problem(POSE_GRAPH["n_poses"], POSE_GRAPH["n_loops"], params=never_exit_params())
    return P.bal_problem(name, solver=D.SOLVER_SCHUR, params=never_exit_params())


def workload_config(name):
    from symforce_b200 import problems as P

    if name == "pose_graph":
        return {
            "workload": f"synthetic Pose3 pose graph: {POSE_GRAPH['n_poses']} poses, {POSE_GRAPH['n_poses'] - 1} odometry + "
                        f"{POSE_GRAPH['n_loops']} loop-closure BetweenFactorPose3 + 1 PriorFactorPose3, multifrontal "
                        f"Cholesky on the full Hessian (N = 600,000), DYNAMIC lambda",
            "n_poses": POSE_GRAPH["n_poses"], "n_loops": POSE_GRAPH["n_loops"],
            "l2": "inputs larger than L2 (fronts + Hessian ~ 1.8 GB)",
        }
    s = P.BAL_SHAPES[name]
    return {
        "workload": f"synthetic BAL {name}-shape: {s['n_cams']} cams / {s['n_pts']} pts / {s['n_obs']} obs, "
                    f"Snavely reprojection + Schur + multifrontal Cholesky, DYNAMIC lambda",
        "n_cams": s["n_cams"], "n_pts": s["n_pts"], "n_obs": s["n_obs"],
        "l2": "inputs larger than L2 (block Hessian > 126 MB)" if s["n_obs"] * 27 * 8 > 126e6
              else "working set fits L2 and is NOT flushed (parity/debug workload, not the headline)",
    }


def bench_config(name, n_gpus):
    """`config` of the JSON line: identical in the sfx and the reference arm."""
    return dict(workload_config(name),
                parallelism=f"landmarks+observations sharded over {n_gpus} GPUs, NCCL reduce of S" if n_gpus > 1
                else "1 GPU")


def never_exit_params():
    """Reference BAL params (DYNAMIC lambda) with early exit disabled so exactly K iterations run."""
    from symforce_b200 import desc as D

    p = D.default_params()
    p.lambda_update_type = D.LAMBDA_DYNAMIC
    p.early_exit_min_reduction = 0.0
    p.early_exit_min_absolute_error = -1.0
    p.lambda_upper_bound = 1e300
    p.iterations = 1000
    return p


class ClockSampler:
    def __init__(self, device=0):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = False
        self.device = device
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def stop(self):
        if self.proc:
            self.proc.terminate()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_baseline(workload, max_seconds=40.0, max_iters=20, warmup_iters=0):
    """The oracle (CPU restatement of the reference algorithm, 1 thread like the reference's
    symforce/opt) on a bounded sample of the workload: LM iterations driven like the reference
    benchmark (`Optimize(values, 1)`), timed by the oracle's own phase timers; index building,
    METIS and symbolic factorization (first call) are excluded, like the GPU side's setup."""
    from symforce_b200 import desc as D, problems as P
    from tests import oracle_capi as O

    prob = make_problem(workload)
    t0 = time.time()
    o = O.OracleProblem(prob)
    iters = 0
    per_iter = []
    first = None
    while True:
        o.reset_timings()
        o.optimize(1)
        if first is None:  # records of the first LM iteration from the initial values: the in-run parity anchor
            first = [float(r.new_error) for r in o.iterations()]
        tm = o.timings()
        # one LM iteration = 1 linearize + 1 factorize + 1 solve (the first call linearizes twice)
        per_iter.append(tm["linearize_s"] / max(tm["n_linearize"], 1) + tm["factorize_s"] / max(tm["n_factorize"], 1)
                        + tm["solve_s"] / max(tm["n_factorize"], 1))
        iters += 1
        if sum(per_iter) > max_seconds / 2 or iters >= max_iters + warmup_iters:
            break
    wall = time.time() - t0
    warm = min(warmup_iters, iters - 1)
    per_iter = per_iter[warm:]
    iters -= warm
    it_s = float(np.mean(per_iter))
    return {
        "steps_run": iters, "warmup_run": warm, "first_iteration_errors": first,
        "value": 1.0 / it_s, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"{iters} LM iteration(s) of the {workload} problem on the CPU oracle ("
                  f"{'simplicial LDLT on H' if workload == 'pose_graph' else 'Schur + simplicial LDLT on S'}, "
                  f"single thread like the reference): {it_s * 1e3:.1f} ms per iteration = linearize "
                  f"{tm['linearize_s'] / max(tm['n_linearize'], 1) * 1e3:.1f} + factorize "
                  f"{tm['factorize_s'] / max(tm['n_factorize'], 1) * 1e3:.1f} + solve "
                  f"{tm['solve_s'] / max(tm['n_factorize'], 1) * 1e3:.1f} ms; one-off setup (index maps, METIS, symbolic) "
                  f"excluded; whole leg took {wall:.0f} s",
        "host_cores_available": os.cpu_count(),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.cpu_workload or args.workload
    # every step is one LM iteration of the full workload on one host core (the reference has no threads); the
    # run is bounded to ~2.5 minutes, so fewer than --steps iterations may fit: `steps` / `warmup` are the counts
    # that actually ran (`requested_steps` / `requested_warmup` echo the flags)
    cb = cpu_baseline(wl, max_seconds=240.0, max_iters=args.steps, warmup_iters=min(max(args.warmup, 0), 1))
    cfg = bench_config(wl, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": cb["steps_run"], "warmup": cb["warmup_run"], "requested_steps": args.steps,
        "requested_warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def rooflines(workload, info, ph):
    """Roofline objects of the three hot phases from the algorithmic bytes / flops of DESIGN.md section 3
     and the measured per-phase times; returns (linearize, schur, factorize, dominant)."""
    from symforce_b200 import problems as P

    is_pg = workload == "pose_graph"
    shape = None if is_pg else P.BAL_SHAPES[workload]
    n_obs, n_cams, n_pts = (shape["n_obs"], shape["n_cams"], shape["n_pts"]) if shape else (0, 0, 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload, {})
    except Exception:
        pass
    lin_bytes = 256 * n_obs + 512 * n_cams + 96 * n_pts
    if is_pg:  # 688 B per between edge + 272 B per pose
        lin_bytes = 688 * (POSE_GRAPH["n_poses"] - 1 + POSE_GRAPH["n_loops"]) + 272 * POSE_GRAPH["n_poses"]
    schur_bytes = 216 * n_obs + 72 * n_pts + 432 * n_cams + 8 * 81 * info["s_blocks"]
    fac_flops = float(info["factor_flops"])
    # MEASURED_PEAKS.json has no FP64 figure: the DMMA peak was measured once on this pool's B200 with
    # tools/micro/fp64_peak.cu (profiles/fp64_peak.json); nominal 40 TFLOP/s otherwise
    fp64_peak, fp64_src = 40.0, "nominal B200 FP64 tensor 40 TFLOP/s (no FP64 number in MEASURED_PEAKS.json)"
    try:
        fp = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
        fp64_peak = float(fp["dmma_tflops"])
        fp64_src = "profiles/fp64_peak.json: mma.sync.m8n8k4.f64 peak measured with tools/micro/fp64_peak.cu"
    except Exception:
        pass
    rl_lin = {"kernel": "linearize_kernel<between/prior> (+zero, error reduce)" if is_pg else
                        "linearize_bal_kernel (+zero, point sums, error reduce)", "bound": "hbm",
              "achieved": lin_bytes / (ph["linearize"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
              "traffic": traffic.get("linearize"), "peak_source": hbm_src}
    rl_lin["frac"] = rl_lin["achieved"] / hbm
    rl_schur = None
    if not is_pg:  # no Schur elimination on the pose-graph path
        rl_schur = {"kernel": "schur_cinv + S-product kernels", "bound": "hbm",
                    "achieved": schur_bytes / (ph["schur"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                    "traffic": traffic.get("schur"), "peak_source": hbm_src}
        rl_schur["frac"] = rl_schur["achieved"] / hbm
        # the S product gathers both whitened 3 x 9 blocks of every (landmark, camera pair) match through L2: that
        # traffic, not HBM, is what bounds it (profiles/r02_results.md); measured L2 read peak: tools/micro/l2_peak.cu
        try:
            l2 = json.load(open(os.path.join(ROOT, "profiles", "l2_peak.json")))
            l2_peak = float(l2["l2_read_gbs"])
            l2_bytes = 432 * info["schur_pairs"] + schur_bytes
            rl_schur["l2"] = {"bound": "l2", "achieved": l2_bytes / (ph["schur"] * 1e-3) / 1e9, "peak": l2_peak,
                              "unit": "GB/s", "frac": l2_bytes / (ph["schur"] * 1e-3) / 1e9 / l2_peak,
                              "bytes": "432 B per match (two 216-byte blocks) + the HBM figure",
                              "peak_source": "profiles/l2_peak.json (tools/micro/l2_peak.cu)"}
        except Exception:
            pass
    rl_fac = {"kernel": "large_factor_kernel (tile-DAG supernodal Cholesky, DMMA m8n8k4)", "bound": "tensor",
              "achieved": fac_flops / (ph["factorize"] * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
              "traffic": traffic.get("large_factor_kernel"), "peak_source": fp64_src}
    rl_fac["frac"] = rl_fac["achieved"] / fp64_peak
    # `achieved` counts the flops of the plan that ran; the plan is chosen by modelled time and may need fewer flops than
    # the reference's METIS_NodeND ordering would (sfx_get_info: plan, ref_ordering_flops)
    rl_fac["plan"] = {-1: "METIS_NodeND", 0: "METIS_NodeND, capped amalgamation"}.get(
        info.get("plan", -1), "dissect to depth %d, then sweep" % info.get("plan", -1))
    if info.get("ref_ordering_flops"):
        rl_fac["gflop_ref_ordering"] = float(info["ref_ordering_flops"]) / 1e9
        # the same time against the flops the reference's ordering needs for this factorization (not the headline: the
        # fraction above is what the hardware did)
        rl_fac["frac_ref_ordering"] = float(info["ref_ordering_flops"]) / (ph["factorize"] * 1e-3) / 1e12 / fp64_peak
    dominant = max([kv for kv in (("factorize", rl_fac), ("schur", rl_schur), ("linearize", rl_lin)) if kv[1]],
                   key=lambda kv: ph[kv[0]])[1]
    return rl_lin, rl_schur, rl_fac, dominant


def measure(workload, K, W, rank, world, local_rank, comm, sample_clocks):
    """One workload on the CUDA path: device-resident K iterations + e2e K x Optimize(values, 1); returns a dict
    (timings are this rank's; the caller takes the max over ranks)."""
    import torch
    import torch.distributed as dist

    from symforce_b200 import capi

    prob = make_problem(workload)
    t0 = time.time()
    gpu = capi.SfxProblem(prob, device=local_rank, rank=rank, world=world, comm=comm)
    setup_s = time.time() - t0
    info = gpu.info()
    pinned = torch.empty(prob.values.shape[0], dtype=torch.float64).pin_memory()
    pinned.numpy()[:] = prob.values

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # in-run parity anchor: initial error and the error after the first LM iteration from the initial values
    # (compared with the cpu_baseline leg's, which runs the same iteration on the oracle)
    gpu.set_values(pinned.numpy())
    gpu.optimize(1)
    first_errors = [float(r.new_error) for r in gpu.iterations()]
    # clocks are sampled by rank 0 only: N concurrent nvidia-smi pollers contend for the driver and for host cores.
    # The poller (50 ms period) starts before the warm-up so that it is up and reporting when the timed regions run
    # (K = 10 iterations are ~65 ms device-resident + ~95 ms end to end); it is stopped right after them.
    sampler = ClockSampler(local_rank)
    if sample_clocks:
        sampler.start()
    # warm-up
    gpu.set_values(pinned.numpy())
    gpu.optimize(W)
    # ---- device-resident: K iterations in one call ---------------------------------------------
    gpu.set_values(pinned.numpy())
    barrier()
    st = gpu.optimize(K)
    barrier()
    tm = gpu.timings()
    iters_run = tm["iterations_run"]
    assert iters_run == K, f"expected {K} iterations, ran {iters_run} (status {st.status})"
    dev_ms = tm["total_ms"]
    # ---- e2e: host buffers, one iteration per call -----------------------------------------------
    barrier()
    t0 = time.perf_counter()
    d2h_bytes = 0
    h2d_bytes = 0
    diag = [0.0, 0.0, 0.0]
    for _ in range(K):
        ta = time.perf_counter()
        h2d_bytes = gpu.set_values(pinned.numpy())
        tb = time.perf_counter()
        gpu.optimize(1)
        tc = time.perf_counter()
        # the result lands in the caller's Values buffer like sym::Optimizer::Optimize(values) does it:
        # Values::Update semantics, only the optimized keys travel back
        d2h_bytes = gpu.update_best_values(pinned.numpy())
        td = time.perf_counter()
        diag[0] += tb - ta
        diag[1] += tc - tb
        diag[2] += td - tc
    if os.environ.get("SFX_E2E_DIAG"):
        print(f"[e2e rank {rank}] per step: set_values {diag[0] / K * 1e3:.2f} ms, optimize(1) {diag[1] / K * 1e3:.2f} ms, "
              f"update_best_values {diag[2] / K * 1e3:.2f} ms", file=sys.stderr, flush=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sample_clocks else None
    gpu.close()
    ph = {"linearize": tm["linearize_ms"] / max(tm["n_linearize"], 1), "schur": tm["schur_ms"] / K,
          "factorize": tm["factorize_ms"] / K, "solve": tm["solve_ms"] / K, "update": tm["update_ms"] / K}
    return dict(dev_ms=dev_ms, e2e_s=e2e_s, clocks=clocks, info=info, setup_s=setup_s, phases=ph,
                launches=int(tm["kernel_launches"]), first_errors=first_errors,
                h2d_bytes=int(h2d_bytes) if h2d_bytes else int(prob.values.nbytes), d2h_bytes=int(d2h_bytes))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    from symforce_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    K, W = args.steps, max(args.warmup, 3)
    assert not (args.workload == "pose_graph" and world > 1), "pose graphs are single-GPU (replicas only, SURVEY.md 8e)"
    # multi-GPU: every rank is given the same problem; libsfx shards landmarks + their observations
    # over the ranks and sums the reduced camera system over NVLink once per iteration
    comm = capi.Comm(rank, world, local_rank) if world > 1 else None
    m = measure(args.workload, K, W, rank, world, local_rank, comm, sample_clocks=(rank == 0))

    t = torch.tensor([m["dev_ms"], m["e2e_s"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    value = K / (dev_ms * 1e-3)  # one LM solve shared by all ranks
    e2e = K / e2e_s

    if rank == 0:
        info, ph = m["info"], m["phases"]
        rl_lin, rl_schur, rl_fac, dominant = rooflines(args.workload, info, ph)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong",  # one fixed problem sharded over N GPUs
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.workload, world),
            "analysis": dict(reduced_dim=info["reduced_dim"], nnz_L=info["nnz_L"], supernodes=info["num_supernodes"],
                             levels=info["num_levels"], max_front=info["max_front"],
                             factor_gflop=float(info["factor_flops"]) / 1e9, setup_s=round(m["setup_s"], 2)),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_bytes"]},
            "gpu_launches": m["launches"],
            "clocks": m["clocks"],
            "phases_ms_per_iteration": ph,
            "roofline": dominant,
            "roofline_linearize": rl_lin, "roofline_schur": rl_schur, "roofline_factorize": rl_fac,
        }
        # secondary workloads (BASELINE.json configs C and E), 1 GPU only: same measurement, shorter
        if world == 1 and args.workload == "final" and args.secondary:
            sec = []
            for wl in ("ladybug", "pose_graph"):
                try:
                    ms = measure(wl, K, W, 0, 1, local_rank, None, sample_clocks=False)
                    rl = rooflines(wl, ms["info"], ms["phases"])
                    sec.append({"config": bench_config(wl, 1), "value": K / (ms["dev_ms"] * 1e-3), "unit": UNIT,
                                "ms_per_step": ms["dev_ms"] / K, "e2e": K / ms["e2e_s"],
                                "phases_ms_per_iteration": ms["phases"], "gpu_launches": ms["launches"],
                                "roofline": rl[3], "roofline_linearize": rl[0], "roofline_schur": rl[1],
                                "roofline_factorize": rl[2]})
                except Exception as e:
                    sec.append({"config": {"workload": wl}, "failed": str(e)})
            line["secondary"] = sec
        if args.cpu_baseline:
            try:
                cb = cpu_baseline(args.cpu_workload or args.workload)
                line["cpu_baseline"] = cb
                if (args.cpu_workload or args.workload) == args.workload and cb.get("first_iteration_errors"):
                    # the oracle's initial error and first-iteration error against the GPU's, same inputs
                    ce, ge = cb["first_iteration_errors"], m["first_errors"]
                    rel = [abs(a - b) / abs(b) for a, b in zip(ge, ce)]
                    line["parity_check"] = {"what": "0.5*|r|^2 at the initial values and after the first LM iteration, "
                                                    "GPU vs CPU oracle in this run", "gpu": ge, "oracle": ce,
                                            "max_rel_diff": max(rel), "tolerance": 1e-9,
                                            "ok": len(ge) == len(ce) and max(rel) <= 1e-9}
            except Exception as e:  # the checker must not break the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
        if "parity_check" in line:
            assert line["parity_check"]["ok"], f"GPU and CPU oracle disagree: {line['parity_check']}"
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

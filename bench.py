#!/usr/bin/env python
"""bench.py — TT rounding (BASELINE config 3: degree 32, n=2, rank 256 -> 128, FP64) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl xb200|reference] [--workload c3|c1]

One step = one TTTensor::round(128) of a device-resident TT with the ranks of TTTensor::random({2}x32, 256)
(ranks 2,4,..,128,256 x17,128,..,2; 18.2 MB of cores).  Prints ONE JSON line (contract in the task statement):
`value` = ms per round with the TT already resident in HBM (CUDA events on the library stream, max over ranks),
`e2e` = the same through the host-pointer API (pinned host cores -> set_component H2D -> round -> get_component D2H),
`roofline` for the dominant kernel class, `cpu_baseline` = the unmodified reference (oracle/_ref/ref_bench) timed on
this box's host cores.  `--impl reference` times the reference's own CPU path for the same workload.
Multi-GPU: a single TT rounding is a sequential chain (replicas only), so each rank rounds its own TT
(weak scaling, no data-path collective); NCCL is used for the barrier and the max over ranks only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# Hardware work queues of the CUDA context (read when the context is created, so before torch touches the device): the default
# of 8 maps the library's worker streams onto 8 queues and serialises independent TTs behind each other (config 5: 400 items/s
# with 8 queues, 1200 with 32 — see DESIGN.md, batches).  xerus_b200 sets the same default when it is imported first.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

WORKLOADS = {
    "c3": dict(d=32, n=2, r=256, target=128, name="TT rounding degree 32, n=2, rank 256 -> 128 (BASELINE configs[2])"),
    "c1": dict(d=8, n=4, r=32, target=16, name="TT rounding degree 8, n=4, rank 32 -> 16 (BASELINE configs[0])"),
}
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench")


def algorithmic_counts(dims, ranks_in, ranks_out):
    """Algorithmic flops / minimum HBM bytes of one round() with the reference operation sequence and LAWN-41 counts
    (SURVEY.md §8d).  Returns (total_flops, svd_flops, qr_flops, gemm_flops, min_bytes)."""
    d = len(dims)
    rin = [1] + list(ranks_in) + [1]
    rout = [1] + list(ranks_out) + [1]
    qr = lambda m, n: 2 * max(m, n) * min(m, n) ** 2 - (2.0 / 3) * min(m, n) ** 3
    qf = lambda m, n: 2 * m * min(m, n) ** 2 - (2.0 / 3) * min(m, n) ** 3
    f_qr = f_svd = f_gemm = 0.0
    for i in range(d - 1):                      # orthogonalisation sweep (transfer_core left -> right)
        m, r = rin[i] * dims[i], rin[i + 1]
        f_qr += qr(m, r) + qf(m, r)
        f_gemm += 2 * min(m, r) * r * dims[i + 1] * rin[i + 2]
    for e in range(d - 2, -1, -1):              # truncation sweep (round_edge right -> left)
        a, r, c2 = rin[e], rin[e + 1], rout[e + 2]
        k = rout[e + 1]
        mr, ml = dims[e + 1] * c2, a * dims[e]
        ra, rb = min(r, mr), min(r, ml)
        f_qr += qr(mr, r) + qf(mr, r) + qr(ml, r) + qf(ml, r)
        f_gemm += 2 * ra * r * rb
        f_svd += 22 * min(ra, rb) ** 3
        f_gemm += 2 * k * ra * mr + 2 * ml * rb * k
    nbytes = 4 * 8 * sum(rin[i] * dims[i] * rin[i + 1] for i in range(d))
    return f_qr + f_svd + f_gemm, f_svd, f_qr, f_gemm, nbytes


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_ref_bench(w, reps, threads=None, dump=None):
    env = dict(os.environ)
    if threads:
        env["OPENBLAS_NUM_THREADS"] = str(threads)
    cmd = [REF_BENCH, "round", str(w["d"]), str(w["n"]), str(w["r"]), str(w["target"]), str(reps)]
    if dump:
        cmd.append(dump)
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def reference_dump(w):
    """Runs the unmodified reference once on this box (`ref_bench round ... dump`: TTTensor::random with seed 0xBAADF00D, then
    round(target)) and returns its input cores, its result cores and its own summary — the checker for the timed path."""
    import tempfile
    import numpy as np
    import struct
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "dump.bin")
        info = run_ref_bench(w, 1, dump=path)
        rec = {}
        with open(path, "rb") as f:                       # record container of oracle/drivers/common.h
            assert f.read(8) == b"XBGOLD01"
            while True:
                hdr = f.read(8)
                if not hdr:
                    break
                name = f.read(struct.unpack("<Q", hdr)[0]).decode()
                nd = struct.unpack("<Q", f.read(8))[0]
                dims = struct.unpack("<%dQ" % nd, f.read(8 * nd)) if nd else ()
                n = int(np.prod(dims)) if nd else 1
                rec[name] = np.frombuffer(f.read(8 * n), dtype="<f8").copy().reshape(dims)
    d = int(rec["in.d"])
    return {"in": [rec["in.c%d" % i] for i in range(d)], "out": [rec["out.c%d" % i] for i in range(d)],
            "ranks": [int(v) for v in rec["out.ranks"]], "info": info}


def check_against_reference(xb, np, out_t, ref_dump, target):
    """`check` of the bench line: the result of the timed path against the reference's result for the same input."""
    if not ref_dump:
        return {"reference": "not available in this snapshot (oracle/_ref/ref_bench missing or --no-cpu-baseline)"}
    if "error" in ref_dump:
        return {"reference": "failed: " + ref_dump["error"]}
    ref = xb.TTTensor.from_cores(ref_dump["out"], core_position=0)
    norm_ref = ref.frob_norm()
    bond = out_t.num_components // 2
    a, b = out_t.copy(), ref.copy()
    a.move_core(bond, True); b.move_core(bond, True)
    ca, cb = a.get_component(bond), b.get_component(bond)
    sa = np.linalg.svd(ca.reshape(ca.shape[0], -1), compute_uv=False)
    sb = np.linalg.svd(cb.reshape(cb.shape[0], -1), compute_uv=False)
    return {"against": "unmodified reference (oracle/_ref/ref_bench, same input cores, seed 0xBAADF00D)",
            "ranks_equal": out_t.ranks() == ref_dump["ranks"],
            "rel_dist_vs_reference": out_t.distance(ref) / norm_ref,
            "norm_rel": abs(out_t.frob_norm() - ref_dump["info"]["norm_out"]) / ref_dump["info"]["norm_out"],
            "sigma_rel": float(np.max(np.abs(sa - sb)) / sb[0]), "sigma_bond": bond,
            "tolerance": "ranks equal; reconstructed TT and singular values <= 1e-9 relative (north_star)"}


def oracle_port_ms(w, reps):
    """Fallback CPU baseline: the numpy restatement (only when the compiled reference is not in the snapshot)."""
    import numpy as np
    from oracle import tt_oracle as O
    rng = np.random.default_rng(0)
    t = O.tt_random([w["n"]] * w["d"], w["r"], rng)
    times = []
    for _ in range(reps):
        c = t.copy()
        t0 = time.perf_counter()
        c.round(w["target"])
        times.append((time.perf_counter() - t0) * 1e3)
    return times


def als_sweep_numbers(xb, np, torch, stream, args, fp64_peak=None):
    """ALS_SPD(A, x, b, 2) = one full sweep, device resident, matrix-free CG local solves."""
    d, n, r = 16, 10, 50
    rng = np.random.default_rng(16)
    A, b = xb.TTOperator.laplace(d, n), xb.TTTensor.ones([n] * d)
    x0 = xb.TTTensor.random([n] * d, r, rng)
    variant = xb.ALSVariant(1, 0, True)
    times, energy, x = [], None, None
    for rep in range(4):
        x = x0.copy()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xb.synchronize()
        e0.record(stream)
        energy = variant(A, x, b, 2)
        e1.record(stream)
        xb.synchronize()
        if rep > 0:
            times.append(e0.elapsed_time(e1))
    residual = A.apply(x).distance(b) / b.frob_norm()
    apply_flops = 2 * 2 * (2 * n * r ** 3) + 2 * 2 * 2 * n * n * r * r          # 5 + 2 + 5 MFLOP at C2
    # time inside the local solves (CUDA events inside the library around every local step: one persistent spd_cg_kernel launch
    # per site plus the local right-hand side)
    xb.profile_enable(True)
    xp = x0.copy()
    variant(A, xp, b, 2)
    xb.synchronize()
    _, solve_launches, solve_ms = xb.profile_get("als_local_step")
    _, _, move_ms = xb.profile_get("als_move_to_next")
    xb.profile_enable(False)
    out = {"workload": "ALS_SPD, Laplace-like TTOperator d=16 n=10 (TT-rank 2), b = ones, solution rank 50, one full sweep",
           "ms_per_sweep": sum(times) / len(times), "energy": energy, "residual": residual,
           "local_solver": "matrix-free CG, %d operator applications per sweep" % variant.last_local_iterations,
           "algorithmic_flops_per_sweep": variant.last_local_iterations * apply_flops,
           "roofline": {"kernel": "spd_cg_kernel (one-site SPD local solve: a whole CG run in one cooperative cluster launch)", "bound": "tensor",
                        "unit": "TFLOP/s", "achieved": (variant.last_local_iterations * apply_flops / (solve_ms * 1e-3) / 1e12) if solve_ms > 0 else None,
                        "peak": fp64_peak, "frac": (variant.last_local_iterations * apply_flops / (solve_ms * 1e-3) / 1e12 / fp64_peak) if solve_ms > 0 and fp64_peak else None,
                        "accounting": "operator applications x (2 x 2 r_A n r^3 + 2 r_A^2 n^2 r^2) flop (SURVEY 8d: 12 MFLOP per apply at config 2)",
                        "kernel_ms_per_sweep": solve_ms, "launches_per_sweep": solve_launches, "move_and_environment_ms_per_sweep": move_ms,
                        "traffic": None,
                        "note": "latency bound: 50 CTAs, two grid barriers per CG iteration on 12 MFLOP of work (profiles/r1_cg.txt)"},
           "reference_cpu": "not runnable inside a bench run at r=50: ~1 h per sweep extrapolated (BASELINE.md section 2)"}
    if os.path.exists(REF_BENCH) and not args.no_cpu_baseline:
        try:
            res = subprocess.run([REF_BENCH, "als", "16", "10", "8", "2", "2"], capture_output=True, text=True, check=True).stdout
            ref = json.loads(res.strip().splitlines()[-1])
            xr0 = xb.TTTensor.random([n] * d, 8, np.random.default_rng(16))
            best = float("inf")
            for rep in range(3):                                       # first repetition loads the solver kernels (lazy module loading)
                xr = xr0.copy()
                xb.synchronize()
                t0 = time.perf_counter()
                variant(A, xr, b, 2)
                xb.synchronize()
                if rep > 0:
                    best = min(best, (time.perf_counter() - t0) * 1e3)
            out["reduced_rank8"] = {"reference_cpu_ms": ref["best_ms"], "xb200_ms": best,
                                    "note": "n_loc = 640 <= als_direct_max: dense reference-semantics path on the GPU"}
        except Exception as ex:
            out["reduced_rank8"] = {"error": repr(ex)}
    return out


_SETUP = {}


def setup(local_rank, world):
    """One process per GPU: device, library, NCCL process group (N > 1), the library stream as a torch stream.  Idempotent."""
    if not _SETUP:
        import torch
        import xerus_b200 as xb
        torch.cuda.set_device(local_rank)
        xb.init(local_rank)
        dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        _SETUP.update(torch=torch, xb=xb, dist=dist, stream=torch.cuda.ExternalStream(xb.stream_handle(), device=local_rank))
    return _SETUP["torch"], _SETUP["xb"], _SETUP["dist"], _SETUP["stream"]


def teardown():
    if _SETUP.get("dist") is not None:
        _SETUP["dist"].destroy_process_group()
    _SETUP.clear()


def batch_workload(args, rank, local_rank, world):
    """--workload c5: BASELINE configs[4], reduced per step: every rank processes `items` independent items
    (y_b = A x_b with x_b = random({4}x12, 64), then y_b.round(64)); items shard by index, no data-path collective.
    Returns the JSON line (rank 0) or None."""
    import numpy as np
    from xerus_b200 import parallel
    torch, xb, dist, stream = setup(local_rank, world)
    d, n, r, per_rank = 12, 4, 64, args.items
    n_items = per_rank * world
    A = xb.TTOperator.laplace(d, n)
    mine = parallel.shard_items(n_items, rank, world)
    xs = {b: xb.TTTensor.random([n] * d, r, parallel.item_rng(5, b)) for b in mine}       # inputs resident in HBM

    xb.set_option("batch_workers", args.workers)
    order = list(mine)

    def step():
        # one C-ABI call for the rank's whole share: the items run on library-owned threads / streams (no interpreter on the path)
        return dict(zip(order, xb.apply_round_batched(A, [xs[b] for b in order], r)))

    for _ in range(args.warmup):
        step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(); xb.synchronize()
    launches0 = xb.kernel_launch_count()
    # every step is timed on its own (CUDA events on the library stream); the results of a step are released between the timed
    # regions, as a caller that consumes them would (their device memory goes back to the workers' pools)
    total, res = 0.0, None
    for _ in range(args.steps):
        res = None
        xb.synchronize_all(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = step()
        e1.record(stream)
        xb.synchronize()
        total += e0.elapsed_time(e1)
    if dist is not None:
        dist.barrier()
    ms = torch.tensor([total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    summaries = parallel.gather_by_item({b: (tuple(y.ranks()), float(y.frob_norm())) for b, y in res.items()}, n_items)
    if rank == 0:
        total_s = float(ms.item()) * 1e-3
        line = {"metric": "TT mat-vec + rounding items/s (FP64)", "value": n_items * args.steps / total_s, "unit": "items/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms.item()) / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "batch of independent degree-12 rank-64 TT contractions + roundings (BASELINE configs[4]), %d items per GPU per step" % per_rank,
                           "l2": "inputs of a step (%d MB of cores per GPU) are larger than L2; results released between timed steps" % (per_rank * 3.4),
                           "parallelism": "items sharded b mod %d, no data-path collective; one xb_tt_apply_round_batched call per GPU, %d library workers (streams + plan caches)" % (world, args.workers)},
                "gpu_launches": xb.kernel_launch_count() - launches0,
                "check": {"ranks_item0": list(summaries[0][0]), "algorithmic_flops_per_item": 1.0e9}}
        return line
    return None


def dmrg_sweep_workload(args, rank, local_rank, world):
    """--workload c4sweep: BASELINE configs[3] as a sweep — one full two-site DMRG_SPD sweep (2 half-sweeps) at bond rank `--bond`
    (default 512) on the Laplace-like operator of degree 24, n = 4 with a right-hand side of full TT-rank (with b = ones the
    two-site split collapses the bond to the solution's low rank after the first site).  Local problems: 4.19 M unknowns, matrix-free
    CG on the three-factor local operator (17.5 GFLOP per application); split: SVD of a 2048 x 2048 matrix cut to rank 512.
    The reference cannot run this (local matrix 8.8 TB, and its two-site driver throws at the sweep turn).  Single GPU; one step."""
    import numpy as np
    torch, xb, dist, stream = setup(local_rank, world)
    if rank != 0:
        return None
    d, n, r = args.sweep_degree, 4, args.bond
    rng = np.random.default_rng(4)
    A = xb.TTOperator.laplace(d, n)
    b = xb.TTTensor.random([n] * d, r, rng); b *= 1.0 / b.frob_norm()
    x0 = xb.TTTensor.random([n] * d, r, rng); x0 *= 1.0 / x0.frob_norm()
    variant = xb.ALSVariant(2, 0, True, localTolerance=1e-8)
    warm = xb.TTTensor.random([n] * 6, 8, rng)                                   # loads the kernels of the path on a toy problem
    variant(xb.TTOperator.laplace(6, n), warm, xb.TTTensor.random([n] * 6, 8, rng), 2)
    out = {}
    for hs in (1, 2):
        x = x0.copy()
        xb.synchronize()
        l0 = xb.kernel_launch_count()
        t0 = time.perf_counter()
        e = variant(A, x, b, hs)
        xb.synchronize()
        out[hs] = dict(seconds=time.perf_counter() - t0, energy=e, applications=variant.last_local_iterations,
                       residual=A.apply(x).distance(b) / b.frob_norm(), ranks_max=max(x.ranks()), launches=xb.kernel_launch_count() - l0)
    apply_flops = 2 * (r * 2) * r * (n * n * r) + 2 * 2 * (r * n * r) * (2 * n) * (n * 2) + 2 * (r * n * n) * (2 * r) * r
    return {"metric": "two-site DMRG_SPD sweep seconds (FP64, bond %d)" % r, "value": out[2]["seconds"], "unit": "s", "n_gpus": 1, "steps": 1, "warmup": 1,
            "ms_per_step": out[2]["seconds"] * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (Laplace-like operator, random right-hand side and start of TT-rank %d, normalised)" % r,
            "config": {"workload": "two-site DMRG sweep, bond rank %d, degree %d, n=4 (BASELINE configs[3] as a sweep)" % (r, d),
                       "local_problem": "%d unknowns, matrix-free CG to 1e-8, SVD split of a %d x %d matrix" % (r * n * n * r, r * n, n * r),
                       "l2": "operands of one application (75 MB) smaller than L2; a sweep touches 24 x 8 MB of cores"},
            "gpu_launches": out[2]["launches"], "half_sweep_1": out[1], "full_sweep": out[2],
            "check": {"energy_monotone": out[2]["energy"] >= out[1]["energy"] - 1e-12 * abs(out[1]["energy"]), "ranks_max": out[2]["ranks_max"]},
            "algorithmic_flops_local_applies": out[2]["applications"] * apply_flops,
            "reference": "not runnable: densified local operator (als.cpp:44) would be 8.8 TB; two-site driver throws at the sweep turn (als.cpp:371,:376)"}


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written, HBM copy GB/s and bf16 TF/s of this pool's B200s); {} if absent."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def _rows_only(parallel, L, A_cores, R, v, rank, world, out):
    """this rank's row block without the exchange (profiled pass of the GEMM class)"""
    import ctypes as C
    from xerus_b200._lib import call
    s = len(A_cores)
    l, r = L.shape[0], R.shape[0]
    begin, end = parallel.slab_range(l, rank, world)
    ptrs = (C.c_void_p * s)(*[a.data_ptr() for a in A_cores])
    dims = (C.c_size_t * (4 * s))(*[int(x) for a in A_cores for x in a.shape])
    call("xb_env_apply_rows", out[begin:end].data_ptr(), L.data_ptr(), l, L.shape[1], ptrs, dims, s, R.data_ptr(), r, R.shape[1], v.data_ptr(), begin, end)


def bond_split_workload(args, rank, local_rank, world):
    """--workload c4: BASELINE configs[3], the two-site DMRG local-operator application at bond rank 512 (n = 4,
    operator rank 2), contraction split along the right bond index across the GPUs + one NCCL sum all-reduce.
    Strong scaling: the total work (17.45 GFLOP per application) is fixed.  Returns the JSON line (rank 0) or None."""
    import numpy as np
    from xerus_b200 import parallel
    torch, xb, dist, stream = setup(local_rank, world)
    r, n, a = args.bond, 4, 2
    g = torch.Generator(device="cuda").manual_seed(4)                  # same inputs on every rank (replicated operands)
    rnd = lambda *shape: torch.randn(*shape, dtype=torch.float64, device="cuda", generator=g)
    L, R, A1, A2, v = rnd(r, a, r), rnd(r, a, r), rnd(a, n, n, a), rnd(a, n, n, a), rnd(r, n, n, r)
    y = torch.empty_like(v)
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    flops = 2 * (r * a) * r * (n * n * r) + 2 * 2 * (r * n * r) * (a * n) * (n * a) + 2 * (r * n * n) * (a * r) * r

    # N > 1: the reduction is fused into the application over peer memory (xb_env_apply_fused: the last GEMM's epilogue writes
    # each rank's row block into that rank's buffer over NVLink, a reduce kernel sums the blocks); --collective nccl times the
    # plain xb_env_apply + ncclAllReduce for comparison.  Both are measured, the fused one is the line's value.
    px = parallel.PeerExchange(r * n * n, r, rank, world, dist=dist) if world > 1 else None
    state = {"y": y}

    split = getattr(args, "split", "rows")

    def step(fused, mode=None):
        mode = mode or split
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if world == 1:
                state["y"] = parallel.env_apply(L, [A1, A2], R, v, out=y)                  # nothing to split
            elif mode == "rows":
                state["y"] = (parallel.row_split_apply_fused(L, [A1, A2], R, v, px) if fused else
                              parallel.row_split_apply(L, [A1, A2], R, v, rank, world, out=y))
            elif fused:
                state["y"] = parallel.bond_split_apply_fused(L, [A1, A2], R, v, px)
            else:
                state["y"] = parallel.bond_split_apply(L, [A1, A2], R, v, rank, world, out=y)
            e1.record(stream)
        return e0, e1

    def timed(fused, mode=None):
        for _ in range(args.warmup):
            step(fused, mode)
            if dist is not None:
                xb.synchronize(); dist.barrier()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(); xb.synchronize()
        l0 = xb.kernel_launch_count()
        events = []
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()                          # the ranks enter a step together (the fused path has no collective to align them)
            events.append(step(fused, mode))
        torch.cuda.synchronize(); xb.synchronize()
        if dist is not None:
            dist.barrier()
        t = torch.tensor([sum(a_.elapsed_time(b_) for a_, b_ in events) / args.steps], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t, xb.kernel_launch_count() - l0

    collective = args.collective if args.collective != "auto" else ("fused" if world <= 4 else "nccl")
    use_fused = world > 1 and collective == "fused"
    ms_other, variants = None, {}
    if world > 1:
        # every exchange variant in the same run: {rows, bond} x {fused over peer memory, NCCL}; the line's value is --split/--collective
        for mode in ("rows", "bond"):
            for fz in (True, False):
                if mode == split and fz == use_fused:
                    continue
                t_, _ = timed(fz, mode)
                variants["%s_%s" % (mode, "fused" if fz else "nccl")] = float(t_.item())
        ms_other = variants.get("%s_%s" % (split, "nccl" if use_fused else "fused"))
        ms_other = torch.tensor([ms_other]) if ms_other is not None else None
    ms, launches = timed(use_fused)
    variants["%s_%s" % (split, "fused" if use_fused else ("nccl" if world > 1 else "single"))] = float(ms.item())
    if px is not None:
        px.check()
    y = state["y"]
    # check against the unsplit application on this rank
    ref = parallel.env_apply(L, [A1, A2], R, v)
    xb.synchronize()
    err = float((y - ref).norm() / ref.norm())
    # time spent in the GEMM class alone (CUDA events inside the library), for the roofline of the dominant kernel
    xb.profile_enable(True)
    if split == "rows" and world > 1:
        _rows_only(parallel, L, [A1, A2], R, v, rank, world, y)
    else:
        parallel.env_apply(L, [A1, A2], R, v, slab=parallel.slab_range(r, rank, world))
    xb.synchronize()
    gsc, glaunch, gms = xb.profile_get("gemm")
    _, mlaunch, mms = xb.profile_get("mid_apply")
    xb.profile_enable(False)
    # the GEMM class holds the two bond contractions (L.v and .R); the operator cores go through mid_apply_kernel (HBM bound)
    gemm_flops = 2 * (r * a) * r * (n * n * r) + 2 * (r * n * n) * (a * r) * r
    mid_bytes = 2 * 2 * 8 * (r * a * n * n * r)                           # per application: two passes, each reads and writes r*a*n*n*r doubles
    if rank == 0:
        aa = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); bb = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        torch.matmul(aa, bb)
        best = 1e9
        for _ in range(3):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(); torch.matmul(aa, bb); s1.record(); torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1))
        peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        t = float(ms.item())
        achieved = gemm_flops / world / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
        hbm_peak = measured_peaks().get("hbm_gbs") or 6540.2
        line = {"metric": "DMRG two-site local-operator application ms (FP64, bond %d)" % r, "value": t, "unit": "ms", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic (i.i.d. N(0,1) environments, operator cores and vector)",
                "config": {"workload": "two-site DMRG local apply, bond rank %d, n=4, operator rank 2 (BASELINE configs[3]); "
                                       "environment contraction split along the bond index over %d GPU(s)" % (r, world),
                           "l2": "flushed between timed iterations",
                           "split": ("left bond index: row blocks of the result, all-gather" if split == "rows" else "right (contracted) bond index: partial sums, reduction"),
                           "collective": ("none" if world == 1 else (("fused: the epilogue of the last GEMM stores the rank's row block into every rank's buffer over NVLink (xb_env_apply_rows_fused)"
                                                                      if split == "rows" else "fused: GEMM epilogue writes row blocks into the peers' buffers over NVLink + reduce kernel (xb_env_apply_fused)")
                                                                     if use_fused else ("xb_env_apply_rows + ncclAllGather" if split == "rows" else "xb_env_apply + ncclAllReduce"))),
                           "exchanged_bytes_per_rank": (int(y.numel() * 8 * (1 if split == "rows" else 2) * (world - 1) / world) if world > 1 else 0)},
                "other_collective_ms": (float(ms_other.item()) if ms_other is not None else None),
                "variants_ms": variants,
                "gpu_launches": launches, "check": {"rel_err_vs_unsplit": err},
                "whole_job_tflops": flops / (t * 1e-3) / 1e12,
                "whole_job_frac_of_peak": flops / (t * 1e-3) / 1e12 / (peak * world),      # of the N GPUs' combined peak
                "roofline": {"kernel": "gemm_f64_big_kernel (128x128 tiles, cp.async, DMMA m8n8k4)", "bound": "tensor", "unit": "TFLOP/s",
                             "achieved": achieved, "peak": peak, "frac": achieved / peak,
                             # dram bytes read + written by the 1024 x 8192 x 512 launch (ncu --set full, profiles/r1_gemm_big.txt);
                             # its operands and result are 105 MB, part of the result stays in the 126 MB L2
                             "traffic": (38317056 + 34946304) if world == 1 else None,
                             "peak_source": "cuBLAS DGEMM 8192^3 measured in this run",
                             "algorithmic_flops_per_rank": gemm_flops / world, "gemm_ms_per_apply": gms, "gemm_launches": glaunch},
                "roofline_mid_apply": {"kernel": "mid_apply_kernel (operator cores, r_A*n = 8)", "bound": "hbm", "unit": "GB/s",
                                       "achieved": (mid_bytes / world / (mms * 1e-3) / 1e9) if mms > 0 else 0.0, "peak": hbm_peak,
                                       "frac": (mid_bytes / world / (mms * 1e-3) / 1e9 / hbm_peak) if mms > 0 else 0.0,
                                       "algorithmic_bytes_per_rank": mid_bytes / world, "ms_per_apply": mms, "launches": mlaunch}}
        return line
    return None


def expected_ranks(w):
    """Ranks of TTTensor::random({n}^d, r) (reduce_to_maximal_ranks, ttNetwork.cpp:370-402) capped at the target."""
    d, n, r = w["d"], w["n"], w["r"]
    return [min(r, n ** (i + 1), n ** (d - 1 - i), w["target"]) for i in range(d - 1)]


def bench_config(w, world):
    """The `config` object, identical in both arms (the driver compares them)."""
    return {"workload": w["name"], "d": w["d"], "n": w["n"], "rank_in": w["r"], "rank_out": w["target"],
            "ranks_out": expected_ranks(w),
            "l2": "xb200 arm: flushed between timed iterations (512 MB write); reference arm: host caches as the reference leaves them",
            "parallelism": "replicas x%d (no data-path collective); reference arm: rank 0 only, host threads" % world}


def reference_arm(args, w, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count()
    threads_used = cores
    if os.path.exists(REF_BENCH):
        # the reference's own CPU implementation with the BLAS threading that is fastest for it on this box: LAPACK
        # factorisations of these sizes do not scale with threads, so both settings are run and the better
        # one is reported
        runs = {cores: run_ref_bench(w, args.steps + args.warmup), 1: run_ref_bench(w, args.steps + args.warmup, threads=1)}
        threads_used = min(runs, key=lambda k: sum(runs[k]["times_ms"][args.warmup:]))
        times = runs[threads_used]["times_ms"][args.warmup:]
        kind = "reference"
    else:
        times = oracle_port_ms(w, args.steps + args.warmup)[args.warmup:]
        kind = "port"
    ms = sum(times) / len(times)
    line = {
        "impl": "reference", "metric": "TT-round ms (FP64)", "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (TTTensor::random, seed 0xBAADF00D)",
        "config": bench_config(w, world),
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": threads_used, "kind": kind,
                         "sample": "%d x TTTensor::round(%d), unmodified xerus + OpenBLAS; faster of {1, %d} BLAS threads on %d host cores"
                                   % (len(times), w["target"], cores, cores)},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="xb200", choices=["xb200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + ["c4", "c5", "c4sweep"])
    ap.add_argument("--sweep-degree", type=int, default=24, help="degree of the TT for --workload c4sweep")
    ap.add_argument("--bond", type=int, default=512, help="bond rank for --workload c4")
    ap.add_argument("--collective", default="auto", choices=["auto", "fused", "nccl"],
                    help="--workload c4 at N > 1: exchange fused over peer memory, or NCCL; auto = fused up to 4 GPUs, NCCL all-gather at 8 "
                         "(measured: at 8 GPUs the 64-byte remote stores of the GEMM epilogue lose to NCCL's bulk all-gather)")
    ap.add_argument("--split", default="rows", choices=["rows", "bond"], help="--workload c4: split the LEFT bond index (row blocks of the result, all-gather; default) or the contracted RIGHT bond index (partial sums, reduction)")
    ap.add_argument("--items", type=int, default=8, help="items per GPU per step for --workload c5")
    ap.add_argument("--workers", type=int, default=16, help="library workers (host threads + CUDA streams) per GPU for --workload c5")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-als", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the c4 / c5 sub-records of the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "xb200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload in ("c4", "c5", "c4sweep") and args.impl == "xb200":
        line = {"c5": batch_workload, "c4": bond_split_workload, "c4sweep": dmrg_sweep_workload}[args.workload](args, rank, local_rank, world)
        if line is not None:
            print(json.dumps(line))
        teardown()
        return
    w = WORKLOADS[args.workload if args.workload in WORKLOADS else "c3"]

    if args.impl == "reference":
        reference_arm(args, w, rank, world)
        return

    import numpy as np
    torch, xb, dist, stream = setup(local_rank, world)

    # ---- input.  Rank 0 rounds the very TT the reference rounds: `ref_bench ... dump` (the unmodified reference, seed 0xBAADF00D)
    # writes TTTensor::random({n}^d, r) and its own round(target); the result of the timed path is checked against it below
    # (`check`).  Without the compiled reference in the snapshot, and on the other ranks (replicas): i.i.d. N(0,1) cores with the
    # ranks of TTTensor::random, canonicalised on the device (move_core(0)).
    dims = [w["n"]] * w["d"]
    ref_dump = None
    if rank == 0 and os.path.exists(REF_BENCH) and not args.no_cpu_baseline:
        try:
            ref_dump = reference_dump(w)
        except Exception as ex:                                    # reported in `check`, never fatal
            ref_dump = {"error": repr(ex)}
    if ref_dump and "in" in ref_dump:
        base = xb.TTTensor.from_cores(ref_dump["in"], core_position=0)
    else:
        base = xb.TTTensor.random(dims, w["r"], np.random.default_rng(0xBAADF00D + rank))
    ranks_in = base.ranks()
    host_cores = [torch.from_numpy(c).pin_memory() for c in base.cores()]     # pinned host copy for the e2e leg
    h2d_bytes = sum(c.numel() * 8 for c in host_cores)

    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")   # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        xb.synchronize()

    def one_round(t):
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            t.round(w["target"])
            e1.record(stream)
        return e0, e1

    # ---- device-resident timing -----------------------------------------------------------------------------------
    clones = [base.copy() for _ in range(args.warmup + args.steps)]
    for i in range(args.warmup):
        one_round(clones[i])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = xb.kernel_launch_count()
    t_wall0 = time.perf_counter()
    events = []
    for i in range(args.steps):
        flush.zero_()                                  # L2 flush between timed iterations (outside the timed events)
        torch.cuda.synchronize()
        events.append(one_round(clones[args.warmup + i]))
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = xb.kernel_launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in events]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    out_t = clones[-1]
    ranks_out = out_t.ranks()
    assert ranks_out == expected_ranks(w), (ranks_out, expected_ranks(w))

    # ---- end to end through the host-pointer API --------------------------------------------------------------------
    def e2e_step():
        t = xb.TTTensor.from_cores([c.numpy() for c in host_cores], core_position=0)   # H2D of every core (pinned)
        t.round(w["target"])
        res = t.cores()                                                                   # D2H of the rounded TT
        return sum(c.size * 8 for c in res)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    d2h_bytes = 0
    for _ in range(args.steps):
        d2h_bytes = e2e_step()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    # the same from pageable host arrays (what the reference's Tensor hands out: plain new[] memory, tensor.cpp:58)
    pageable = [c.numpy().copy() for c in host_cores]

    def e2e_pageable_step():
        t = xb.TTTensor.from_cores(pageable, core_position=0)
        t.round(w["target"])
        return t.cores()

    e2e_pageable_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_pageable_step()
    barrier()
    e2e_pageable_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    # ---- the partitioned workloads in the same run (BASELINE configs[3] and [4]): all ranks take part ------------------------
    sub = {}
    if not args.no_sub:
        import copy
        for name, fn, over in (("c4", bond_split_workload, dict(steps=max(3, min(args.steps, 10)), warmup=3)),
                               ("c5", batch_workload, dict(steps=3, warmup=3, items=192))):
            a2 = copy.copy(args)
            for k, v in over.items():
                setattr(a2, k, v)
            try:
                sub[name] = fn(a2, rank, local_rank, world)
            except Exception as ex:                                # reported, never fatal for the headline line
                sub[name] = {"error": repr(ex)}
            barrier()

    # ---- roofline of the dominant kernel class (extra profiled pass, CUDA events inside the library) ----------------
    total_f, svd_f, qr_f, gemm_f, min_bytes = algorithmic_counts(dims, ranks_in, ranks_out)
    prof = {}
    if rank == 0:
        xb.profile_enable(True)
        p = base.copy()
        p.round(w["target"])
        xb.synchronize()
        for k in ["svd_jacobi", "svd", "qr", "gemm"]:
            prof[k] = xb.profile_get(k)
        xb.profile_enable(False)
        # FP64 yardstick: cuBLAS DGEMM 8192^3 through torch (MEASURED_PEAKS.json carries no FP64 entry)
        a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        best = 1e9
        for _ in range(3):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(); torch.matmul(a, b); s1.record(); torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1))
        fp64_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, b

    # ---- second half of BASELINE's metric: one ALS_SPD sweep (configs[1]: Laplace-like A, d=16, n=10, r=50) ----------
    als = None
    if rank == 0 and not args.no_als:
        try:
            als = als_sweep_numbers(xb, np, torch, stream, args, fp64_peak)
        except Exception as ex:       # reported, never fatal for the bench line
            als = {"error": repr(ex)}

    if rank != 0:
        teardown()
        return

    check = check_against_reference(xb, np, out_t, ref_dump, w["target"])
    jac_scopes, jac_launches, jac_ms = prof["svd_jacobi"]
    svd_tf = svd_f / (jac_ms * 1e-3) / 1e12 if jac_ms > 0 else 0.0
    roofline = {
        "kernel": "jacobi_split_kernel (one-sided block Jacobi SVD, X and V workers, one cooperative cluster launch per SVD)", "bound": "tensor", "unit": "TFLOP/s",
        "achieved": svd_tf, "peak": fp64_peak, "frac": svd_tf / fp64_peak if fp64_peak else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch on a 256-column problem (ncu --set full, profiles/r1_jacobi_split.txt)
        "traffic": 1087744 + 312576, "traffic_unit": "bytes per launch (256-column SVD; algorithmic: 1.5 MB matrix in + factors out)",
        "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (no FP64 entry in MEASURED_PEAKS.json; nominal ~40)",
        "algorithmic_flops_per_round": svd_f, "accounting": "22*min(m,n)^3 per SVD (SURVEY §8d)",
        "launches_per_round": jac_launches, "kernel_ms_per_round": jac_ms,
        "share_of_step": jac_ms / ms_per_step if ms_per_step else None,
        "classes_ms_per_round": {k: v[2] for k, v in prof.items()},
        "whole_round": {"flops": total_f, "TFLOP/s": total_f / (ms_per_step * 1e-3) / 1e12, "min_hbm_bytes": min_bytes},
    }

    cpu = None
    if not args.no_cpu_baseline:
        try:
            if os.path.exists(REF_BENCH):
                reps = 12 if args.workload == "c3" else 200
                r1 = run_ref_bench(w, reps, threads=1)
                rN = run_ref_bench(w, reps)
                best = min((r1["median_ms"], 1), (rN["median_ms"], os.cpu_count()))
                cpu = {"value": best[0], "unit": "ms", "cores": best[1], "kind": "reference",
                       "sample": "%d x TTTensor::round(%d) on the same shape, unmodified xerus v3.0.1 + OpenBLAS; median; "
                                 "1 thread %.1f ms, %d threads %.1f ms" % (reps, w["target"], r1["median_ms"], os.cpu_count(), rN["median_ms"])}
            else:
                times = oracle_port_ms(w, 5)
                cpu = {"value": statistics.median(times), "unit": "ms", "cores": os.cpu_count(), "kind": "port",
                       "sample": "5 x numpy restatement (oracle/tt_oracle.py) of round(%d)" % w["target"]}
        except Exception as ex:   # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": "ms", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (ex,)}

    line = {
        "metric": "TT-round ms (FP64)", "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (i.i.d. N(0,1) cores, ranks of TTTensor::random)",
        "config": bench_config(w, world),
        "e2e": {"value": float(e2e_ms.item()), "unit": "ms", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "host_buffers": "pinned", "pageable_value": e2e_pageable_ms},
        "gpu_launches": launches, "wall_ms_timed_region": wall_ms, "step_ms": step_ms,
        "clocks": clocks, "check": check, "roofline": roofline, "cpu_baseline": cpu, "als": als,
        "c4": sub.get("c4"), "c5": sub.get("c5"),
    }
    print(json.dumps(line))
    teardown()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- SDDMM GFLOPS (2*nnz*K / time) of the BSMR hot path on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): the nips-shaped matrix of BASELINE.json configs[1] at K=128,
alpha=0.3, delta=0.3.  dataset/nips.mtx is not in the reference tree, so a seeded synthetic
of the same shape / nnz is used (bsmr-sddmm_b200/synth.py) unless dataset/nips.mtx exists.
At N GPUs the job is N such row blocks stacked into one matrix, reordered globally and sharded
over the ranks by work-balanced ranges of reordered row panels (bsmr_plan_set_shard: nnz plus tile count): weak scaling,
per-GPU work fixed, no collective on the data path (B is replicated before the clock starts).

A "step" = one pass of the hot path (wide row-group tcgen05 kernel + dense-block tcgen05 kernel +
residual kernel, each over its share of the nnz) over the matrix.  The inputs (27 MB) fit in L2, so L2 is flushed (a 512 MB buffer is rewritten) between
timed steps, outside the event pairs.  Times are CUDA events on the stream the kernels run on.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
# stdout carries exactly one JSON line: everything else that libraries write to fd 1 while the bench runs (NCCL prints
# its version banner there) goes to stderr; emit() puts the line on the real stdout
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

K_DEFAULT, ALPHA, DELTA = 128, 0.3, 0.3
METRIC, UNIT = "sddmm_gflops", "GFLOP/s"


def load_workload(pkg, blocks):
    """`blocks` nips-shaped row blocks stacked vertically (blocks == 1: the nips config itself)."""
    path = os.path.join(ROOT, "dataset", "nips.mtx")
    parts = []
    source = "synthetic nips-shaped (1500x12419, nnz 746316), seed 1500+i"
    for i in range(blocks):
        if i == 0 and os.path.exists(path):
            M, N, ro, ci = pkg.synth.read_mtx(path)
            source = "dataset/nips.mtx + synthetic blocks"
        else:
            M, N, ro, ci = pkg.synth.nips_like(seed=1500 + i)
        parts.append((M, N, ro.astype(np.int64), ci))
    N = parts[0][1]
    ro = [np.zeros(1, dtype=np.int64)]
    off = 0
    for _, _, r, c in parts:
        ro.append(r[1:] + off)
        off += len(c)
    ro = np.concatenate(ro).astype(np.uint32)
    ci = np.concatenate([p[3] for p in parts])
    return sum(p[0] for p in parts), N, ro, ci, source


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Host threads this process may use (its CPU affinity mask), independent of OMP_NUM_THREADS."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (sddmm_cpu, OpenMP, all host
    threads) from oracle/_ref; the oracle port when that library did not travel.  Rank 0 only."""
    if rank != 0:
        return
    pkg = entry.load_package()
    from oracle.bindings import Oracle, Ref, REF_SO
    M, N, ro, ci, source = load_workload(pkg, max(1, args.gpus))
    K = args.k
    A, B = pkg.synth.make_ab(M, N, K)
    # all the host threads the box offers: torchrun exports OMP_NUM_THREADS=1, which would otherwise pin the
    # reference's OpenMP loop to one core (the harness calls omp_set_num_threads(threads) around the call)
    threads = host_threads()
    if os.path.exists(REF_SO):
        impl, kind = Ref(), "reference"
    else:
        impl, kind = Oracle(), "port"
    steps = max(1, min(args.steps, 20))
    for _ in range(max(1, min(args.warmup, 2))):
        impl.sddmm_cpu(M, N, K, A, B, ro, ci, num_threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        impl.sddmm_cpu(M, N, K, A, B, ro, ci, num_threads=threads)
    dt = (time.perf_counter() - t0) / steps
    gflops = 2.0 * len(ci) * K / dt / 1e9
    sample = "whole workload (%d nnz, K=%d) per step, %d steps" % (len(ci), K, steps)
    line = {"impl": "reference", "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(M, N, len(ci), K, source, args.gpus),
            "cpu_baseline": {"value": gflops, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": gflops, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(M, N, nnz, K, source, gpus):
    return {"workload": "nips K=%d alpha=%.1f delta=%.1f (BASELINE.json configs[1]); %s" % (K, ALPHA, DELTA, source),
            "M": int(M), "N": int(N), "nnz": int(nnz), "K": int(K), "alpha": ALPHA, "delta": DELTA,
            "blocks": max(1, gpus), "sharding": "work-balanced (nnz + wide tiles) reordered row-panel ranges" if gpus > 1 else "none",
            "l2": "flushed between timed steps (512 MB buffer rewritten outside the event pairs)"}


def kernel_alg_bytes(K, rows_touched, cols_touched, nnz, extra_index_bytes_per_nnz=0):
    return 4 * K * (rows_touched + cols_touched) + (8 + extra_index_bytes_per_nnz) * nnz


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--k", type=int, default=K_DEFAULT)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = entry.load_package()
    K = args.k

    M, N, ro, ci, source = load_workload(pkg, world)
    nnz = len(ci)
    A, B = pkg.synth.make_ab(M, N, K)
    # an explicit (non-default) stream shared by torch and the library: the CUDA events below must be
    # recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = pkg.Context(local_rank, stream.cuda_stream)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    t0 = time.perf_counter()
    plan.reorder(ALPHA, DELTA)                     # block_size from calculateBlockSize, reference_compat reduction
    reorder_wall_ms = (time.perf_counter() - t0) * 1e3          # first call of the process: module load + allocations
    reorder_warm_ms = None
    for _ in range(3):                                          # the scratch arena settles at its high-water mark on the 2nd call
        t0 = time.perf_counter()
        plan.reorder(ALPHA, DELTA)
        dt = (time.perf_counter() - t0) * 1e3
        reorder_warm_ms = dt if reorder_warm_ms is None else min(reorder_warm_ms, dt)
    info = plan.info()
    shard_nnz = nnz
    shard_sizes = None
    if world > 1:
        p0s, p1s, shard_nnz = plan.set_shard(rank, world)
        shard_panels = (p0s, p1s)
        sizes = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([p1s - p0s, shard_nnz], dtype=torch.int64, device="cuda"))
        shard_sizes = [[int(x[0]), int(x[1])] for x in sizes]

    dA = torch.from_numpy(A).cuda()
    dB = torch.empty((N, K), dtype=torch.float32, device="cuda")
    if world > 1:
        # B is replicated: produced on rank 0, broadcast over NCCL/NVLink before the clock starts
        if rank == 0:
            dB.copy_(torch.from_numpy(B))
        dist.broadcast(dB, src=0)
    else:
        dB.copy_(torch.from_numpy(B))
    dP = torch.zeros(nnz, dtype=torch.float32, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_timed(ev0, ev1):
        flush.fill_(1)                              # L2 flush, outside the timed pair
        ev0.record(stream)
        plan.sddmm(K, dA, dB, dP, iterations=1, timed=False)
        ev1.record(stream)

    for _ in range(max(3, args.warmup)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_timed(e0, e1)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in events:
        step_timed(e0, e1)
    barrier()
    launches = ctx.launch_count() - launches0
    exec_choice = plan.execution_choice(K)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(e0.elapsed_time(e1) for e0, e1 in events))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    per_rank_ms = [total_ms / args.steps]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x.item()) / args.steps for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = 2.0 * nnz * K / (ms_per_step * 1e-3) / 1e9

    # ---- per-kernel times (cold L2), the roofline of the dominant kernel -------------------------
    wide_ms, dense_ms, res_ms = [], [], []
    for _ in range(min(args.steps, 20)):
        flush.fill_(1)
        w, a, b = plan.sddmm_profile3(K, dA, dB, dP)
        wide_ms.append(w)
        dense_ms.append(a)
        res_ms.append(b)
    wide_ms, dense_ms, res_ms = float(np.mean(wide_ms)), float(np.mean(dense_ms)), float(np.mean(res_ms))
    # back-to-back, L2-resident figure (diagnostic: what the reference's own timing loop measures)
    hot_ms = plan.sddmm(K, dA, dB, dP, iterations=100)

    # ---- e2e: host buffers through the C ABI (H2D A,B + kernels + D2H P inside every step) -------
    # (1) latency: one blocking bsmr_sddmm_host call per step; (2) throughput (the reported e2e value): the same
    # per-step work through bsmr_sddmm_host_submit / _wait, two steps in flight, so the copy-in of step i+1 and the
    # copy-out of step i-1 overlap the kernels of step i.  Every step copies its own A, B in and its own P out.
    hA = torch.from_numpy(A).pin_memory()
    hB = torch.from_numpy(B).pin_memory()
    hPs = [torch.zeros(nnz, dtype=torch.float32).pin_memory() for _ in range(2)]
    hP = hPs[0]
    e2e_steps = max(3, min(args.steps, 20))
    # N > 1: every rank copies in A and B, runs its shard and copies its P out (the shards are disjoint index sets of the
    # CSR value array; rows of a shard are scattered over the original matrix, so the copy-out is the whole array).  The
    # result is left distributed over the ranks' host buffers: no collective on this path either.
    for _ in range(2):
        plan.sddmm_host(K, hA, hB, hP)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.sddmm_host(K, hA, hB, hP)
    barrier()
    e2e_serial_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    def pipelined(n):
        for i in range(n):
            plan.sddmm_host_submit(K, hA, hB, hPs[i % 2])
        plan.sddmm_host_wait()

    pipelined(4)
    barrier()
    t0 = time.perf_counter()
    pipelined(e2e_steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms, e2e_serial_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, e2e_serial_ms = float(t[0].item()), float(t[1].item())
    e2e_value = 2.0 * nnz * K / (e2e_ms * 1e-3) / 1e9
    hP = hPs[(e2e_steps - 1) % 2]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (rank 0's shard when N > 1) ---------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # which nnz each of the three kernels computes (rank 0's shard when N > 1): the wide kernel owns whole row groups
    # (256 reordered rows = 16 panels); outside them the BSMR split applies (residual list / dense blocks)
    rows = plan.vector("reordered_rows")
    sv = plan.vector("sparse_values")
    svo = plan.vector("sparse_value_offsets")
    gw = plan.vector("group_wide").astype(bool)
    p0, p1 = (0, info["num_row_panels"]) if world == 1 else shard_panels
    row_of = np.repeat(np.arange(M, dtype=np.int64), np.diff(ro.astype(np.int64)))
    pos_of_row = np.full(M, -1, dtype=np.int64)
    pos_of_row[rows] = np.arange(len(rows))
    pos = pos_of_row[row_of]                                   # reordered position of every nnz's row
    in_shard = (pos >= p0 * 16) & (pos < p1 * 16)
    is_wide = np.zeros(nnz, dtype=bool)
    if info["num_wide_tiles"]:
        is_wide = in_shard & gw[np.clip(pos // pkg.WIDE_GROUP_ROWS, 0, len(gw) - 1)]
    is_res = np.zeros(nnz, dtype=bool)
    is_res[sv[svo[p0]:svo[p1]]] = True
    is_res &= ~is_wide
    is_blk = in_shard & ~is_wide & ~is_res

    def part_bytes(mask, extra):
        n = int(mask.sum())
        if n == 0:
            return 0
        return kernel_alg_bytes(K, len(np.unique(row_of[mask])), len(np.unique(ci[mask])), n, extra)

    parts = [("wide_sddmm_kernel (tcgen05, 256-row groups)", wide_ms, part_bytes(is_wide, 0)),
             ("dense_sddmm_kernel (tcgen05, 16x16 blocks)", dense_ms, part_bytes(is_blk, 0)),
             ("residual_rows_kernel", res_ms, part_bytes(is_res, 8))]
    step_bytes = pkg.synth.algorithmic_bytes(M, N, K, ro, ci) if world == 1 else sum(x[2] for x in parts)
    dom, dom_ms, dom_bytes = max(parts, key=lambda x: x[1])
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture of this workload, if any
    traffic = tensor_pct = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and world == 1:
        committed = json.load(open(tpath))
        traffic = committed.get("K%d" % K, {}).get(dom.split(" ")[0])
        tensor_pct = committed.get("tensor_pipe_active_pct_K%d" % K, {}).get(dom.split(" ")[0])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "tensor_pipe_active_pct": tensor_pct,     # sm__pipe_tensor_cycles_active of the committed ncu capture (profiles/)
                "algorithmic_bytes_per_launch": int(dom_bytes), "kernel_ms": dom_ms,
                "other_kernels_ms": {n: ms for n, ms, _ in parts if n != dom},
                "step": {"algorithmic_bytes": int(step_bytes), "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                         "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak}}

    # ---- CPU baseline: the reference's sddmm_cpu on this box's host cores (rank 0, N = 1 only) ----
    cpu_baseline = None
    if world == 1:
        from oracle.bindings import Oracle, Ref, REF_SO
        threads = host_threads()
        if os.path.exists(REF_SO):
            impl, kind = Ref(), "reference"
        else:
            impl, kind = Oracle(), "port"
        impl.sddmm_cpu(M, N, K, A, B, ro, ci, num_threads=threads)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            Pcpu = impl.sddmm_cpu(M, N, K, A, B, ro, ci, num_threads=threads)
        cpu_dt = (time.perf_counter() - t0) / reps
        cpu_baseline = {"value": 2.0 * nnz * K / cpu_dt / 1e9, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": "whole workload, %d passes after 1 warm-up (%.1f ms per pass)" % (reps, cpu_dt * 1e3)}
        # the bench also checks what it timed
        bad = Oracle().check_data(Pcpu, hP.numpy())
        if bad:
            raise SystemExit("bench: %d of %d values outside the reference tolerance" % (bad, nnz))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32 (tcgen05 wide groups and dense blocks) / f32 (residual), f32 accumulate", "data": "synthetic",
            "config": workload_config(M, N, nnz, K, source, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int((M + N) * K * 4) * world, "d2h_bytes_per_step": int(nnz * 4) * world,
                    "how": "bsmr_sddmm_host_submit/_wait, pinned host buffers, 2 steps in flight (copies of neighbouring steps "
                           "overlap the kernels); %d steps%s" % (e2e_steps, "" if world == 1 else "; every rank copies A, B in and "
                           "its P out (result left distributed over the ranks, bytes summed over ranks)"),
                    "blocking_call_ms": e2e_serial_ms,
                    "blocking_call_value": 2.0 * nnz * K / (e2e_serial_ms * 1e-3) / 1e9},
            "gpu_launches": int(launches),
            "ms_per_step_by_rank": per_rank_ms,
            "shards_panels_nnz": shard_sizes,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "reorder": {"row_ms": info["row_reordering_ms"], "col_ms": info["col_reordering_ms"],
                        "format_ms": info["format_build_ms"], "wall_ms": reorder_warm_ms,
                        "first_call_wall_ms": reorder_wall_ms,
                        "block_size": info["block_size"], "clusters": info["num_clusters_true"],
                        "dense_nnz": int(info["num_dense_values"]), "residual_nnz": int(info["num_sparse_values"]),
                        "dense_tiles": info["num_dense_tiles"],
                        "plan": {"row_groups": info["num_row_groups"], "wide_groups": info["num_wide_groups"],
                                 "wide_tiles": info["num_wide_tiles"], "wide_nnz": int(info["num_wide_values"]),
                                 "block_tiles": info["num_block_tiles"], "block_nnz": int(info["num_block_values"]),
                                 "residual_nnz_outside_wide": int(info["num_residual_values"]),
                                 "wide_format_ms": info["wide_format_ms"]},
                        "gflops_incl_reorder": 2.0 * nnz * K / ((ms_per_step + reorder_warm_ms) * 1e-3) / 1e9},
            "execution_plan": ("pinned by BSMR_NO_AUTOTUNE: " if os.environ.get("BSMR_NO_AUTOTUNE") else "") + {0: "wide row groups + BSMR split (three kernels)", 4: "BSMR split (dense blocks + residual)",
                               2: "CSR-order residual kernel"}.get(exec_choice, "?") + " (chosen by measurement per K)",
            "kernels": {"wide_ms_cold": wide_ms, "dense_ms_cold": dense_ms, "residual_ms_cold": res_ms, "step_ms_hot_l2": hot_ms,
                        "gflops_hot_l2": 2.0 * shard_nnz * K / (hot_ms * 1e-3) / 1e9}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

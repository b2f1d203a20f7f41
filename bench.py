#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json):

  2^20-point Goldilocks NTTs/sec (batched, 256 columns per GPU, forward + inverse per step) and
  Tip5 Merkle leaves/sec (2^24 leaves per GPU), on N B200s of one node, next to the reference
  algorithm's CPU path on the same host.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LOG2N = 20
COLS_PER_GPU = 256
MERKLE_LOG2 = 24
METRIC = "2^20-pt Goldilocks NTTs/sec (batched BFieldElement columns, forward+inverse) and Tip5 Merkle leaves/sec"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled during the timed region"""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self) -> bool:
        """fast path: NVML in-process (a sample every ~2 ms, so that even a 50 ms timed region is covered)"""
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [("nvmlClocksThrottleReasonHwSlowdown", 0x8), ("nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    ("nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), ("nvmlClocksThrottleReasonSwPowerCap", 0x4)]
            masks = [getattr(nv, name, default) for name, default in bits]
        except Exception:
            return False
        while not self.stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = get_reasons(h)
                self.samples.append([str(sm), str(mx)] + ["Active" if r & m else "Not Active" for m in masks])
            except Exception:
                pass
            self.stop.wait(0.002)
        return True

    def _run(self):
        if self._run_nvml():
            return
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=5)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_reference_leg(steps: int, warmup: int, sample_cols: int | None, merkle_log2: int):
    """The reference algorithm (oracle port: bit-reversal + radix-2 DIT + montyred, scalar Tip5,
    subtree-per-thread Merkle) with OpenMP in rayon's role, on all host threads."""
    import numpy as np

    import oracle

    o = oracle.get(native=True)
    # all host threads, like rayon's default pool; torchrun exports OMP_NUM_THREADS=1, undo that here
    o.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    cores = o.num_threads()
    n = 1 << LOG2N
    cols = sample_cols or 128  # half of the 256-column batch: ~1 s of wall time per step on 16 threads
    x = oracle.splitmix64_words(0x210001, n * cols)
    orig = x.copy()
    for _ in range(max(1, min(warmup, 1))):
        o.ntt_batch(x, n, 1, cols, False)
        o.ntt_batch(x, n, 1, cols, True)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        o.ntt_batch(x, n, 1, cols, False)
        o.ntt_batch(x, n, 1, cols, True)
        times.append(time.perf_counter() - t0)
    assert np.array_equal(x, orig)
    ntt_per_s = 2 * cols * len(times) / sum(times)
    leafs = oracle.splitmix64_words(0x210002, 5 << merkle_log2)
    o.merkle_par_new(leafs)
    t0 = time.perf_counter()
    reps = max(1, min(steps, 3))
    for _ in range(reps):
        o.merkle_par_new(leafs)
    merkle_s = (time.perf_counter() - t0) / reps
    return {
        "ntt_per_s": ntt_per_s, "ms_per_step": 1e3 * sum(times) / len(times), "cores": cores,
        "sample": f"{cols} columns x 2^{LOG2N} forward+inverse per step; Merkle par_new over 2^{merkle_log2} leaves",
        "merkle_leaves_per_s": (1 << merkle_log2) / merkle_s,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cols", type=int, default=COLS_PER_GPU)
    ap.add_argument("--merkle-log2", type=int, default=MERKLE_LOG2)
    ap.add_argument("--cpu-cols", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lde", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner)
    # is diverted to stderr for the rest of the run
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_leg(args.steps, args.warmup, args.cpu_cols, 22)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["ntt_per_s"], "unit": "NTT/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            # same workload as the GPU arm (BASELINE configs[1]); each step is a bounded sample of it
            "config": {"workload": f"batched {args.cols}x2^{LOG2N}-point BFieldElement NTT then iNTT per GPU (BASELINE configs[1]), "
                                   "reference algorithm (CPU port) on all host threads",
                       "columns_per_gpu": args.cols, "log2_n": LOG2N, "sample": r["sample"]},
            "cpu_baseline": {"value": r["ntt_per_s"], "unit": "NTT/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"], "merkle_leaves_per_s": r["merkle_leaves_per_s"]},
            "e2e": {"value": r["ntt_per_s"], "unit": "NTT/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "merkle": {"value": r["merkle_leaves_per_s"], "unit": "leaves/s"},
        }
        print(json.dumps(line), file=json_out, flush=True)
        return

    import numpy as np
    import torch

    tf = importlib.import_module("twenty-first_b200")
    dev = tf.device
    torch.cuda.set_device(local_rank)
    cuda = torch.device(f"cuda:{local_rank}")
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ["NCCL_DEBUG"] = os.environ.get("TF21_NCCL_DEBUG", "WARN")  # keep NCCL's banner off stdout
        dist.init_process_group("nccl", device_id=cuda)
    dev.init(local_rank)
    peak_gbs, peak_src = load_peaks()
    n = 1 << LOG2N
    cols = args.cols

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=cuda)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b))

    # ---- inputs, resident in HBM (each rank owns its columns / leaves: no data-path collective) ----
    gen = torch.Generator(device=cuda)
    gen.manual_seed(0x210001 + rank)
    x = torch.randint(0, 2**63 - 1, (cols * n,), dtype=torch.int64, device=cuda, generator=gen)  # < p: canonical
    x0_check = x[: 4 * n].clone()
    n_leafs = 1 << args.merkle_log2
    leafs = torch.randint(0, 2**63 - 1, (5 * n_leafs,), dtype=torch.int64, device=cuda, generator=gen)
    nodes = torch.zeros(10 * n_leafs, dtype=torch.int64, device=cuda)
    roots_all = torch.zeros(5 * world, dtype=torch.int64, device=cuda)
    cap = torch.zeros(10 * world, dtype=torch.int64, device=cuda)

    def ntt_step():
        dev.ntt_(x, n, 1, False)
        dev.ntt_(x, n, 1, True)

    def merkle_step():
        dev.merkle_build(leafs, nodes)
        if dist is not None:  # gather the tree cap (one 40-byte root per rank) and finish the top levels
            dist.all_gather_into_tensor(roots_all, nodes[5:10])
            dev.merkle_build(roots_all, cap)

    for _ in range(warmup):
        ntt_step()
        merkle_step()
    torch.cuda.synchronize()

    with ClockSampler(local_rank) as clocks:
        l0 = dev.kernel_launch_count()
        ntt_ms = timed(ntt_step, args.steps)
        l1 = dev.kernel_launch_count()
        merkle_ms = timed(merkle_step, args.steps)
        l2 = dev.kernel_launch_count()
    clock_summary = clocks.summary()
    assert torch.equal(x[: 4 * n], x0_check), "NTT -> iNTT round trip changed the data"

    transforms_per_step = 2 * cols * world
    ntt_per_s = transforms_per_step * args.steps / (ntt_ms * 1e-3)
    leaves_per_s = n_leafs * world * args.steps / (merkle_ms * 1e-3)

    # ---- roofline pass: per-launch CUDA-event timing of every kernel in the same steps ------------
    dev.profile_enable(True)
    for _ in range(args.steps):
        ntt_step()
    torch.cuda.synchronize()
    ntt_prof = dev.profile_read()
    dev.profile_enable(True)
    for _ in range(args.steps):
        dev.merkle_build(leafs, nodes)
    torch.cuda.synchronize()
    merkle_prof = dev.profile_read()
    dev.profile_enable(False)

    def by_kernel(prof):
        agg = {}
        for name, ms in prof:
            tot, cnt = agg.get(name, (0.0, 0))
            agg[name] = (tot + ms, cnt + 1)
        return agg

    # DRAM traffic per launch of the dominant kernels, from the committed `ncu --set full` capture of the
    # same shapes (profiles/r01_ncu_traffic.json, written by tools/ncu_traffic.py); null if absent
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic_db = json.load(f)

    def traffic_of(kernel_name, shape_key):
        e = traffic_db.get(shape_key, {})
        for k, v in e.items():
            if k.split("<")[0] == kernel_name.split("<")[0]:
                return v
        return None

    ntt_agg = by_kernel(ntt_prof)
    ntt_kernel_ms = sum(t for t, _ in ntt_agg.values())
    # one batched transform (all its pass launches) is the roofline unit: 2*8*n bytes per column
    launches_per_transform = max(1, len(ntt_prof) // (2 * args.steps))
    ntt_alg_bytes = 16 * n * cols
    ntt_avg_ms = ntt_kernel_ms / (2 * args.steps)
    ntt_achieved = ntt_alg_bytes / (ntt_avg_ms * 1e-3) / 1e9
    dominant = max(ntt_agg.items(), key=lambda kv: kv[1][0])
    merkle_agg = by_kernel(merkle_prof)
    merkle_kernel_ms = sum(t for t, _ in merkle_agg.values()) / args.steps
    merkle_alg_bytes = 80 * n_leafs
    merkle_achieved = merkle_alg_bytes / (merkle_kernel_ms * 1e-3) / 1e9

    # ---- secondary workload: BASELINE configs[3], coset LDE 2^22 -> 2^26 XFieldElement (rank 0 timing) ----
    lde = None
    if not args.no_lde:
        li, lo = 22, 26
        vals = torch.randint(0, 2**63 - 1, (3 << li,), dtype=torch.int64, device=cuda, generator=gen)
        out = torch.zeros(3 << lo, dtype=torch.int64, device=cuda)
        g7 = tf.BFieldElement.generator()
        for _ in range(2):
            dev.coset_lde(vals, 3, g7, 1 << lo, g7, out)
        lde_steps = max(1, min(args.steps, 3))
        lde_ms = timed(lambda: dev.coset_lde(vals, 3, g7, 1 << lo, g7, out), lde_steps) / lde_steps
        lde_bytes = 24 * ((1 << li) + (1 << lo))
        lde = {"workload": "coset LDE 2^22 -> 2^26 XFieldElement per GPU (BASELINE configs[3])", "ms": lde_ms,
               "algorithmic_bytes": lde_bytes,
               "roofline": {"bound": "hbm", "achieved": lde_bytes / (lde_ms * 1e-3) / 1e9, "peak": peak_gbs,
                            "unit": "GB/s", "frac": lde_bytes / (lde_ms * 1e-3) / 1e9 / peak_gbs}}
        del vals, out

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ------------------------
    e2e = None
    if not args.no_e2e:
        e2e_cols = cols
        host = torch.empty(e2e_cols * n, dtype=torch.int64).pin_memory()
        host.copy_(x[: e2e_cols * n])
        host_np = host.numpy().view(np.uint64)
        api = importlib.import_module("twenty-first_b200.api")
        e2e_steps = max(1, min(args.steps, 2))
        api.ntt_batch(host_np, n, 1, False)
        api.ntt_batch(host_np, n, 1, True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            api.ntt_batch(host_np, n, 1, False)
            api.ntt_batch(host_np, n, 1, True)
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        bytes_dir = 2 * e2e_cols * n * 8  # two calls per step, each copies the whole batch in and out
        e2e = {"value": 2 * e2e_cols * world * e2e_steps / (e2e_ms * 1e-3), "unit": "NTT/s",
               "h2d_bytes_per_step": bytes_dir, "d2h_bytes_per_step": bytes_dir, "steps": e2e_steps,
               "api": "tf21_ntt / tf21_intt (host pointers, pinned)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_leg(2, 1, args.cpu_cols, 22)
        cpu = {"value": r["ntt_per_s"], "unit": "NTT/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "merkle_leaves_per_s": r["merkle_leaves_per_s"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": ntt_per_s, "unit": "NTT/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ntt_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"batched {cols}x2^{LOG2N}-point BFieldElement NTT then iNTT per GPU (BASELINE configs[1]), "
                                   f"device resident, bit-exact round trip checked",
                       "columns_per_gpu": cols, "log2_n": LOG2N, "l2_policy": "inputs (2 GiB per GPU) larger than L2",
                       "transforms_per_step": transforms_per_step},
            "roofline": {"bound": "hbm", "achieved": ntt_achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": ntt_achieved / peak_gbs,
                         "traffic": (lambda a, b: (a + b) if (a and b) else None)(
                             traffic_of("ntt1024_col_kernel", f"ntt20_cols{cols}"),
                             traffic_of("ntt1024_row_kernel", f"ntt20_cols{cols}")),
                         "traffic_note": "dram read+write bytes of one col-pass launch + one row-pass launch (ncu)",
                         "peak_source": peak_src,
                         "kernel": "one batched 2^20 transform = " + " + ".join(sorted(ntt_agg)),
                         "launches_per_transform": launches_per_transform,
                         "algorithmic_bytes_per_launch_group": ntt_alg_bytes,
                         "avg_ms_per_launch_group": ntt_avg_ms,
                         "dominant_kernel": dominant[0],
                         "per_kernel_ms_avg": {k: t / c for k, (t, c) in ntt_agg.items()}},
            "merkle": {"value": leaves_per_s, "unit": "leaves/s", "ms_per_step": merkle_ms / args.steps,
                       "leaves_per_gpu": n_leafs,
                       "roofline": {"bound": "hbm", "achieved": merkle_achieved, "peak": peak_gbs, "unit": "GB/s",
                                    "frac": merkle_achieved / peak_gbs,
                                    "traffic": traffic_of("tip5_hash10_kernel", f"merkle{args.merkle_log2}"),
                                    "traffic_note": "dram bytes of the leaf-level tip5_hash10_kernel launch (ncu)",
                                    "algorithmic_bytes": merkle_alg_bytes, "kernel_ms": merkle_kernel_ms,
                                    "per_kernel_ms_total": {k: t / args.steps for k, (t, c) in merkle_agg.items()}}},
            "lde": lde, "cpu_baseline": cpu, "e2e": e2e, "clocks": clock_summary,
            "gpu_launches": int(l2 - l0), "gpu_launches_ntt": int(l1 - l0),
        }
        print(json.dumps(line), file=json_out, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

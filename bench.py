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

    def _init_nvml(self):
        """NVML is initialised on the calling thread BEFORE the timed region (nvmlInit alone can take 100 ms and more:
        a sampler that initialises inside its thread misses most of a 0.2 s region)"""
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
            get_reasons(h)  # first call of each entry point done here
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self._nvml = (nv, h, mx, get_reasons, masks)
        except Exception:
            self._nvml = None

    def _run_nvml(self) -> bool:
        """fast path: NVML in-process (a sample every ~2 ms, so that even a 50 ms timed region is covered)"""
        if self._nvml is None:
            return False
        nv, h, mx, get_reasons, masks = self._nvml
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
        self._init_nvml()
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
    cols = sample_cols or COLS_PER_GPU  # the whole 256-column batch of configs[1]: ~2 s of wall time per step on 16 threads
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
    # Tip5 in both forms the reference ships: the scalar build (mds_generated, tip5/mod.rs:175-506) and, when this host
    # has avx512f/bw/ifma/vbmi, its AVX-512 build (tip5/avx512.rs); the faster one is the Merkle baseline
    merkle_rates = {}
    for impl in ("scalar", "avx512"):
        if o.tip5_set_impl(impl) != impl:
            continue
        o.merkle_par_new(leafs)
        t0 = time.perf_counter()
        reps = max(1, min(steps, 3))
        for _ in range(reps):
            o.merkle_par_new(leafs)
        merkle_rates[impl] = (1 << merkle_log2) * reps / (time.perf_counter() - t0)
    o.tip5_set_impl("scalar")
    merkle_impl = max(merkle_rates, key=merkle_rates.get)
    # BASELINE configs[0]: one 2^10-point BFieldElement ntt -> intt round trip, single thread (the reference's own
    # CPU-runnable case; nearest reference bench shapes are 2^7 / 2^18, benches/ntt.rs:19-21)
    v = oracle.splitmix64_words(0x210000, 1 << 10)
    v0 = v.copy()
    o.ntt(v, 1)
    o.intt(v, 1)
    assert np.array_equal(v, v0)
    reps0 = 2000
    big = np.tile(v0, reps0)
    t0 = time.perf_counter()
    o.ntt_batch(big, 1 << 10, 1, reps0, False, 1)
    o.ntt_batch(big, 1 << 10, 1, reps0, True, 1)
    cfg0_us = (time.perf_counter() - t0) / reps0 * 1e6
    assert np.array_equal(big[: 1 << 10], v0)
    return {
        "config0_roundtrip_us": cfg0_us,
        "ntt_per_s": ntt_per_s, "ms_per_step": 1e3 * sum(times) / len(times), "cores": cores,
        "sample": f"{cols} columns x 2^{LOG2N} forward+inverse per step; Merkle par_new over 2^{merkle_log2} leaves",
        "merkle_leaves_per_s": merkle_rates[merkle_impl], "merkle_tip5_impl": merkle_impl,
        "merkle_leaves_per_s_by_tip5_impl": merkle_rates,
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
    ap.add_argument("--no-cfg4", action="store_true")
    ap.add_argument("--cols-total", type=int, default=1024, help="BASELINE configs[4]: columns over all ranks (N > 1)")
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
        r = cpu_reference_leg(args.steps, args.warmup, args.cpu_cols or args.cols, 22)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["ntt_per_s"], "unit": "NTT/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            # same workload as the GPU arm (BASELINE configs[1]); each step is a bounded sample of it
            "config": {"workload": f"batched {args.cols}x2^{LOG2N}-point BFieldElement NTT then iNTT per GPU (BASELINE configs[1]), "
                                   "reference algorithm (CPU port) on all host threads",
                       "columns_per_gpu": args.cols, "log2_n": LOG2N, "sample": r["sample"]},
            "cpu_baseline": {"value": r["ntt_per_s"], "unit": "NTT/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"], "merkle_leaves_per_s": r["merkle_leaves_per_s"],
                             "merkle_tip5_impl": r["merkle_tip5_impl"], "merkle_leaves_per_s_by_tip5_impl": r["merkle_leaves_per_s_by_tip5_impl"],
                             "config0_2^10_ntt_intt_roundtrip_us_1thread": r["config0_roundtrip_us"]},
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
        # NCCL_DEBUG is left as the launcher set it: stdout was re-pointed at stderr above, so NCCL's banner and
        # INFO lines cannot reach the JSON line
        dist.init_process_group("nccl", device_id=cuda)
    dev.init(local_rank)
    peak_gbs, peak_src = load_peaks()
    n = 1 << LOG2N
    cols = args.cols
    P = 0xFFFFFFFF00000001

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

    def canonical_words(count, seed):
        """uniform raw words in [0, p) -- the whole range, including [2^63, p) -- as int64 bit patterns"""
        g = torch.Generator(device=cuda)
        g.manual_seed(seed)
        v = torch.randint(-(2**63), 2**63 - 1, (count,), dtype=torch.int64, device=cuda, generator=g)
        # as u64 the values >= p are the int64 values in [-(2^32 - 1), -1]; v - p = v + 2^32 - 1 (mod 2^64)
        hi = (v < 0) & (v >= -(2**32 - 1))
        return torch.where(hi, v + (2**32 - 1), v)

    def column_words(first_col, n_cols, seed):
        """SplitMix64 stream per global column index (element k of column c = stream seed ^ (c << 32)):
        a shard regenerates exactly its own columns whatever the number of ranks"""
        out = torch.empty(n_cols * n, dtype=torch.int64, device=cuda)
        k = torch.arange(1, n + 1, dtype=torch.int64, device=cuda)
        GOLD, M1, M2 = -7046029254386353131, -4658895280553007687, -7723592293110705685  # u64 constants as int64

        def lsr(z, sh):
            return (z >> sh) & ((1 << (64 - sh)) - 1)

        for c in range(n_cols):
            z = k * GOLD + (seed ^ ((first_col + c) << 32))
            z = (z ^ lsr(z, 30)) * M1
            z = (z ^ lsr(z, 27)) * M2
            z = z ^ lsr(z, 31)
            hi = (z < 0) & (z >= -(2**32 - 1))
            out[c * n:(c + 1) * n] = torch.where(hi, z + (2**32 - 1), z)
        return out

    def u64_sum_mod_p(t):
        """sum of raw words mod p of an int64 tensor holding u64 bit patterns (exact, via 32-bit halves)"""
        lo = (t & 0xFFFFFFFF).sum().item()
        hi = ((t >> 32) & 0xFFFFFFFF).sum().item()
        return (lo + (hi << 32)) % P

    # ---- inputs, resident in HBM (each rank owns its columns / leaves: no data-path collective) ----
    x = canonical_words(cols * n, 0x210001 + rank)
    x0_check = x[: 4 * n].clone()
    n_leafs = 1 << args.merkle_log2
    leafs = canonical_words(5 * n_leafs, 0x210002 + 1000 * rank)
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

    # ---- N-rank Merkle root against the oracle (outside the timed region) ---------------------------------
    # The timed build's root when the whole tree is small enough for the CPU checker, else the same code path
    # (local build -> cap all-gather -> top levels) over 2^20 leaves per rank.
    merkle_parity = None
    if dist is not None:
        full = n_leafs * world <= (1 << 26)
        v_leafs = leafs if full else canonical_words(5 << 20, 0x210012 + 1000 * rank)
        v_nodes = nodes if full else torch.zeros(10 << 20, dtype=torch.int64, device=cuda)
        dev.merkle_build(v_leafs, v_nodes)
        dist.all_gather_into_tensor(roots_all, v_nodes[5:10])
        dev.merkle_build(roots_all, cap)
        root = cap[5:10].cpu().numpy().view(np.uint64)
        all_leafs = [torch.empty_like(v_leafs) for _ in range(world)] if rank == 0 else None
        dist.gather(v_leafs, all_leafs, dst=0)
        if rank == 0:
            import oracle

            o = oracle.get(native=True)
            o.set_num_threads(len(os.sched_getaffinity(0)))
            host = torch.cat([t.cpu() for t in all_leafs]).numpy().view(np.uint64)
            rc, want = o.merkle_par_frugal_root(host)
            merkle_parity = {"ok": bool(rc == 0 and np.array_equal(root, want)), "ranks": world,
                             "leaves_total": int(host.size // 5), "timed_tree": full,
                             "checker": "oracle.merkle_par_frugal_root over the concatenated leaves of all ranks"}
            assert merkle_parity["ok"], "N-rank Merkle root differs from the oracle"
            del host
        del all_leafs

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

    # DRAM traffic per launch of the dominant kernels: NOT measured in this run (ncu replays kernels) but read from
    # the committed `ncu --set full` capture of the same shapes (profiles/r0X_ncu_traffic.json, tools/ncu_traffic.py)
    traffic_db, traffic_src = {}, None
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic_db = json.load(f)
            traffic_src = "profiles/" + name
            break

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
    dominant = max(ntt_agg.items(), key=lambda kv: kv[1][0] / kv[1][1])  # the longest average launch
    ntt_traffic = [traffic_of(k, f"ntt20_cols{cols}") for k in sorted({kk.split("<")[0] for kk in ntt_agg})]
    merkle_agg = by_kernel(merkle_prof)
    merkle_kernel_ms = sum(t for t, _ in merkle_agg.values()) / args.steps
    merkle_alg_bytes = 80 * n_leafs
    merkle_achieved = merkle_alg_bytes / (merkle_kernel_ms * 1e-3) / 1e9

    # ---- BASELINE configs[4]: 1024 x 2^20 columns sharded by column over the ranks (N > 1) ---------------
    config4 = None
    if dist is not None and not args.no_cfg4:
        total_cols = args.cols_total
        per = total_cols // world
        del x
        torch.cuda.empty_cache()
        first = rank * per
        y = column_words(first, per, 0x210004)
        col_sums = [u64_sum_mod_p(y[c * n:(c + 1) * n]) for c in range(0, per, max(1, per // 8))]
        sums = torch.zeros(world, dtype=torch.int64, device=cuda)
        my_sum = torch.zeros(1, dtype=torch.int64, device=cuda)

        def cfg4_step():
            dev.ntt_(y, n, 1, False)
            dev.ntt_(y, n, 1, True)
            dist.all_gather_into_tensor(sums, my_sum)  # the only collective of configs[4]: per-shard checksums

        for _ in range(2):
            cfg4_step()
        c4_steps = max(1, min(args.steps, 5))
        c4_ms = timed(cfg4_step, c4_steps)
        # per-shard 64-bit checksum of the FORWARD transform (wrapping sum of all words) + the DC property
        # X[0] = sum_j x[j] (mod p) on a sample of columns; the checksum of checksums does not depend on N
        dev.ntt_(y, n, 1, False)
        my_sum.copy_(y.sum().reshape(1))
        dc_ok = all((int(y[c * n].item()) % (1 << 64)) % P == s_ for c, s_ in zip(range(0, per, max(1, per // 8)), col_sums))
        dist.all_gather_into_tensor(sums, my_sum)
        dev.ntt_(y, n, 1, True)
        rt_ok = bool(torch.equal(y[:n], column_words(first, 1, 0x210004)))
        ok = torch.tensor([int(dc_ok and rt_ok)], device=cuda)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        shard_sums = [int(v) % (1 << 64) for v in sums.cpu().tolist()]
        config4 = {"workload": f"{total_cols} x 2^{LOG2N}-point BFieldElement NTT then iNTT sharded by column over {world} "
                               f"B200 ({per} columns per GPU), per-shard checksum all-gather over NCCL (BASELINE configs[4])",
                   "value": 2 * total_cols * c4_steps / (c4_ms * 1e-3), "unit": "NTT/s", "ms_per_step": c4_ms / c4_steps,
                   "columns_per_gpu": per, "checksum_of_checksums": f"{sum(shard_sums) % (1 << 64):016x}",
                   "shard_checksums": [f"{v:016x}" for v in shard_sums],
                   "parity": {"dc_term_equals_column_sum": bool(ok.item()), "round_trip": bool(ok.item())}}
        assert ok.item() == 1, "configs[4] parity properties failed"
        del y
        torch.cuda.empty_cache()
        x = canonical_words(cols * n, 0x210001 + rank)

    # ---- secondary workload: BASELINE configs[3], coset LDE 2^22 -> 2^26 XFieldElement (rank 0 timing) ----
    lde = None
    if not args.no_lde:
        li, lo = 22, 26
        vals = canonical_words(3 << li, 0x210003 + rank)
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
        # the same calls on PAGEABLE memory (a plain Vec<BFieldElement> / numpy array): through the library's pinned
        # staging ring (csrc/host_stage.cuh); reported next to the pinned figure, not instead of it
        pg = np.empty(e2e_cols * n, dtype=np.uint64)
        pg[:] = host_np
        api.ntt_batch(pg, n, 1, False)
        api.ntt_batch(pg, n, 1, True)
        barrier()
        t0 = time.perf_counter()
        api.ntt_batch(pg, n, 1, False)
        api.ntt_batch(pg, n, 1, True)
        torch.cuda.synchronize()
        pg_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e["pageable"] = {"value": 2 * e2e_cols * world / (pg_ms * 1e-3), "unit": "NTT/s",
                           "round_trip_ok": bool(np.array_equal(pg, host_np)),
                           "api": "tf21_ntt / tf21_intt (host pointers, pageable memory, pinned staging ring)"}
        del pg

    # ---- Merkle e2e: host leaves in, host node array out, through tf21_merkle_build (N = 1 only: it is a per-process call) --
    merkle_e2e = None
    if not args.no_e2e and world == 1:
        h_leafs = torch.empty(5 * n_leafs, dtype=torch.int64).pin_memory()
        h_leafs.copy_(leafs)
        h_nodes = torch.empty(10 * n_leafs, dtype=torch.int64).pin_memory()
        ln, nn = h_leafs.numpy().view(np.uint64), h_nodes.numpy().view(np.uint64)
        B = importlib.import_module("twenty-first_b200._binding")
        import ctypes

        def host_build(a, b):
            B.check(B.lib.tf21_merkle_build(ctypes.c_void_p(a.ctypes.data), n_leafs, ctypes.c_void_p(b.ctypes.data)))

        host_build(ln, nn)
        t0 = time.perf_counter()
        host_build(ln, nn)
        pin_s = time.perf_counter() - t0
        root_ok = bool(np.array_equal(nn[5:10], nodes[5:10].cpu().numpy().view(np.uint64)))
        pl, pn = np.array(ln), np.empty_like(nn)
        host_build(pl, pn)
        t0 = time.perf_counter()
        host_build(pl, pn)
        pg_s = time.perf_counter() - t0
        merkle_e2e = {"value": n_leafs / pin_s, "unit": "leaves/s", "api": "tf21_merkle_build (host pointers, pinned)",
                      "h2d_bytes": 40 * n_leafs, "d2h_bytes": 40 * n_leafs,
                      "note": "the leaf half of the node array is copied on the host, only inner nodes cross PCIe",
                      "root_equals_device_build": root_ok,
                      "pageable": {"value": n_leafs / pg_s, "unit": "leaves/s",
                                   "nodes_equal_pinned_run": bool(np.array_equal(pn, nn))}}
        del h_leafs, h_nodes, pl, pn

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_leg(2, 1, args.cpu_cols or args.cols, 22)
        cpu = {"value": r["ntt_per_s"], "unit": "NTT/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "merkle_leaves_per_s": r["merkle_leaves_per_s"], "merkle_tip5_impl": r["merkle_tip5_impl"],
               "merkle_leaves_per_s_by_tip5_impl": r["merkle_leaves_per_s_by_tip5_impl"],
               "config0_2^10_ntt_intt_roundtrip_us_1thread": r["config0_roundtrip_us"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": ntt_per_s, "unit": "NTT/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ntt_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"batched {cols}x2^{LOG2N}-point BFieldElement NTT then iNTT per GPU (BASELINE configs[1]), "
                                   f"device resident, bit-exact round trip checked",
                       "columns_per_gpu": cols, "log2_n": LOG2N, "l2_policy": "inputs (2 GiB per GPU) larger than L2",
                       "inputs": "uniform canonical raw words over the whole of [0, p)",
                       "transforms_per_step": transforms_per_step},
            "roofline": {"bound": "hbm", "achieved": ntt_achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": ntt_achieved / peak_gbs,
                         "traffic": sum(ntt_traffic) if ntt_traffic and all(ntt_traffic) else None,
                         "traffic_note": "dram read+write bytes of the launches of one batched transform, from the committed "
                                         f"ncu --set full capture {traffic_src} of this shape (not re-measured in this run)",
                         "peak_source": peak_src,
                         "kernel": "one batched 2^20 transform = " + " + ".join(sorted(ntt_agg)),
                         "launches_per_transform": launches_per_transform,
                         "algorithmic_bytes_per_launch_group": ntt_alg_bytes,
                         "avg_ms_per_launch_group": ntt_avg_ms,
                         "dominant_kernel": dominant[0],
                         "per_kernel_ms_avg": {k: t / c for k, (t, c) in ntt_agg.items()}},
            "merkle": {"value": leaves_per_s, "unit": "leaves/s", "ms_per_step": merkle_ms / args.steps,
                       "leaves_per_gpu": n_leafs, "n_rank_root_parity": merkle_parity, "e2e": merkle_e2e,
                       "roofline": {"bound": "hbm", "achieved": merkle_achieved, "peak": peak_gbs, "unit": "GB/s",
                                    "frac": merkle_achieved / peak_gbs,
                                    "traffic": traffic_of("tip5_hash10_kernel", f"merkle{args.merkle_log2}"),
                                    "traffic_note": f"dram bytes of the leaf-level launch, committed capture {traffic_src}",
                                    "algorithmic_bytes": merkle_alg_bytes, "kernel_ms": merkle_kernel_ms,
                                    "per_kernel_ms_total": {k: t / args.steps for k, (t, c) in merkle_agg.items()}}},
            "config4": config4, "lde": lde, "cpu_baseline": cpu, "e2e": e2e, "clocks": clock_summary,
            "gpu_launches": int(l2 - l0), "gpu_launches_ntt": int(l1 - l0),
        }
        print(json.dumps(line), file=json_out, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

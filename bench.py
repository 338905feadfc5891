#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 bzip2 encoder core (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # banzai's CPU algorithm (oracle port)

One "step" = one pass of the hot path (banzai::encode, lib/lib.rs:84) over one batch: the
1 GiB mixed synthetic corpus at level 9 (BASELINE.json configs[1]).  With N > 1 (torchrun, one
rank per GPU) every rank encodes its own 1 GiB object — blocks are independent, there is no
data-path collective — so per-GPU work is fixed ("weak" scaling) and `value` is the whole-job
aggregate.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import corpus  # noqa: E402

METRIC = "encode MB/s (level 9)"
UNIT = "MB/s"
WORKLOADS = {
    # name: (corpus kind, bytes, level, seed)
    "mixed-1GiB-L9": ("mixed", 1 << 30, 9, corpus.SEED_C2),
    "text-10MB-L9": ("text", 10 * 1000 * 1000, 9, corpus.SEED_C1),
    "text-4GiB-L1": ("text", 4 << 30, 1, corpus.SEED_C4),
    "random-4GiB-L9": ("random", 4 << 30, 9, corpus.SEED_C5),
    "mixed-256MiB-L9": ("mixed", 256 << 20, 9, corpus.SEED_C2),
    "random-1GiB-L9": ("random", 1 << 30, 9, corpus.SEED_C5),
    "text-1GiB-L1": ("text", 1 << 30, 1, corpus.SEED_C4),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mixed-1GiB-L9", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--radix-bits", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=1, help="concurrent block batches per GPU")
    ap.add_argument("--set", action="append", default=[], help="ctx tunable key=value (repeatable)")
    return ap.parse_args()


def dist_env():
    from banzai_b200 import dist as D
    return D.env()


# ---------------------------------------------------------------------------------- clocks

CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
               "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
               "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={CLOCK_QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smmax, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smmax.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smmax), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------- reference arm

def oracle_mt_encode(shards, level):
    """banzai's CPU algorithm (oracle port), one thread per shard (ctypes releases the GIL)."""
    from oracle import pyoracle as O
    O.lib()
    outs = [None] * len(shards)

    def work(i):
        outs[i] = len(O.encode(shards[i], level))

    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(shards))]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0, outs


def run_reference(args, kind, size, level, seed):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_thread = 12 * 1000 * 1000            # ~2-3 s of CPU per thread per step
    sample = min(size, per_thread * cores)
    data = corpus.by_name(kind, sample, seed)
    n_sh = min(cores, max(1, sample // per_thread))
    bounds = np.linspace(0, sample, n_sh + 1).astype(np.int64)
    shards = [data[bounds[i]:bounds[i + 1]] for i in range(n_sh)]
    for _ in range(min(args.warmup, 1)):
        oracle_mt_encode(shards, level)
    times = []
    for _ in range(args.steps):
        dt, _ = oracle_mt_encode(shards, level)
        times.append(dt)
    total = sum(times)
    value = sample * args.steps / total / 1e6
    desc = (f"first {sample} bytes of the workload split into {n_sh} contiguous shards, one oracle "
            f"thread per shard (banzai restatement in C, -O3 -march=native)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "level": level, "bytes_per_gpu": size,
                   "note": "reference arm = CPU restatement of banzai (no Rust toolchain in the image)"},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": n_sh, "kind": "port",
                         "sample": desc},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- B200 arm

def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(workload):
    p = os.path.join(ROOT, "profiles", "bwt_traffic.json")
    try:
        t = json.load(open(p))
        return t.get(workload)
    except Exception:
        return None


def run_b200(args, kind, size, level, seed):
    rank, world, local = dist_env()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)

    import banzai_b200
    from banzai_b200 import _ffi
    from banzai_b200 import dist as D
    group = D.Group(backend="nccl", device=torch.device("cuda", local))
    lib = _ffi.lib

    ctx = banzai_b200.Context(devices=[local] * args.lanes)
    if args.radix_bits:
        ctx.set("bwt_radix_bits", args.radix_bits)
    for kv in args.set:
        k, v = kv.split("=")
        ctx.set(k, int(v))

    # pinned host input (the e2e arm copies from it every step), synthetic corpus per rank
    h_in = lib.bnz_host_alloc(size)
    if not h_in:
        raise SystemExit("bnz_host_alloc failed")
    h_arr = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(size,))
    corpus.by_name(kind, size, D.object_seed(seed, rank), out=h_arr)
    d_in = lib.bnz_device_alloc(ctx._h, size + 64)
    out_cap = size + size // 8 + (64 << 20)        # incompressible input grows by ~0.4 %
    d_out = lib.bnz_device_alloc(ctx._h, out_cap)
    if not d_in or not d_out:
        raise SystemExit("device allocation failed")
    ctx._check(lib.bnz_memcpy_h2d(ctx._h, d_in, h_in, size))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            group.barrier()
            torch.cuda.synchronize()

    max_over_ranks = group.max_over_ranks

    # ---- device-resident arm ("value")
    for _ in range(args.warmup):
        out_len = ctx.encode_device(d_in, h_in, size, level, d_out, out_cap)
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    stats = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_len = ctx.encode_device(d_in, h_in, size, level, d_out, out_cap)
        stats.append(ctx.stats())
    sync_all()
    t_dev = max_over_ranks(time.perf_counter() - t0)

    # ---- end-to-end arm: pinned host input -> finished .bz2 in host memory, through the C ABI
    for _ in range(args.warmup):
        o, n = ctx.encode_ptr(h_in, size, level)
        ctx.free_out(o)
    sync_all()
    t0 = time.perf_counter()
    e2e_stats = []
    for _ in range(args.steps):
        o, n = ctx.encode_ptr(h_in, size, level)
        e2e_stats.append(ctx.stats())
        ctx.free_out(o)
    sync_all()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()

    if rank == 0:
        peak, peak_src = load_peak()
        bwt_ms = statistics.mean(s["bwt_ms"] for s in stats)
        alg = statistics.mean(s["bwt_algorithmic_bytes"] for s in stats)
        achieved = alg / (bwt_ms * 1e-3) / 1e9
        step_ms = t_dev / args.steps * 1e3
        launches = sum(s["kernel_launches"] for s in stats) + sum(s["kernel_launches"] for s in e2e_stats)
        line = {
            "metric": METRIC, "value": round(D.aggregate_throughput(size, world, t_dev, args.steps), 1), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "level": level, "bytes_per_gpu": size,
                       "blocks_per_gpu": stats[-1]["n_blocks"], "corpus": kind,
                       "l2": "input (and every per-stage array) is far larger than the 126 MB L2",
                       "sharding": "one independent object per GPU, no collective",
                       "compressed_bytes_per_gpu": int(out_len),
                       "bwt_radix_bits": stats[-1]["bwt_radix_bits"]},
            "e2e": {"value": round(D.aggregate_throughput(size, world, t_e2e, args.steps), 1), "unit": UNIT,
                    "h2d_bytes_per_step": int(e2e_stats[-1]["h2d_bytes"]),
                    "d2h_bytes_per_step": int(e2e_stats[-1]["d2h_bytes"])},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "bwt_sort_kernel", "bound": "hbm", "achieved": round(achieved, 1),
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": load_traffic(args.workload),
                         "algorithmic_bytes_per_launch": int(alg), "kernel_ms": round(bwt_ms, 3),
                         "share_of_step": round(bwt_ms / step_ms, 3)},
            "stage_ms": {k: round(statistics.mean(s[k] for s in stats), 3)
                         for k in ("rle_ms", "bwt_ms", "mtf_ms", "huff_ms", "pack_ms", "total_ms")},
            "bwt": {"rounds_avg": round(stats[-1]["bwt_rounds_total"] / max(1, stats[-1]["n_blocks"]), 2),
                    "rounds_max": stats[-1]["bwt_max_rounds"],
                    "sum_active_over_n": round(stats[-1]["bwt_sum_active"] / max(1, stats[-1]["bwt_n"]), 3),
                    "tied_blocks": stats[-1]["bwt_tied_blocks"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle as O
            sample = min(size, 48 * 1000 * 1000)
            t0 = time.perf_counter()
            O.encode(h_arr[:sample], level)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": round(sample / dt / 1e6, 3), "unit": UNIT, "cores": 1,
                                    "kind": "port",
                                    "sample": f"first {sample} bytes of the same workload, single "
                                              f"thread, C restatement of banzai (oracle)"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)

    lib.bnz_device_free(ctx._h, d_in)
    lib.bnz_device_free(ctx._h, d_out)
    lib.bnz_host_free(h_in)
    ctx.close()
    group.close()


def main():
    args = parse_args()
    kind, size, level, seed = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, kind, size, level, seed)
    else:
        run_b200(args, kind, size, level, seed)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 bzip2 encoder core (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # banzai's CPU algorithm (oracle port)

One "step" = one pass of the hot path (banzai::encode, lib/lib.rs:84) over ONE object: the 1 GiB
mixed synthetic corpus at level 9 (BASELINE.json configs[1], "1 GiB ... across 8xB200").

N > 1 (torchrun, one rank per GPU): the ONE stream is sharded block-wise over the N GPUs
("scaling": "strong").  `banzai::encode` is a single call of a single host program, so rank 0
drives all N devices through one context (bnz_ctx_create(n_gpus=N): every device uploads and cuts
its own 1/N byte range, the ranges exchange two scalars through host memory, the compressed shards
are concatenated at bit offsets — no data-path collective); the other ranks hold the timing
barriers.  Outside the timed region the N-GPU stream is byte-compared with the 1-GPU stream
("parity_check").  The round-1 measurement — every rank encodes its own object on its own GPU —
is kept as the extra key "replicas".

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import hashlib
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
    "mixed-8GiB-L9": ("mixed", 8 << 30, 9, corpus.SEED_C2),
    "text-10MB-L9": ("text", 10 * 1000 * 1000, 9, corpus.SEED_C1),
    "text-4GiB-L1": ("text", 4 << 30, 1, corpus.SEED_C4),
    "random-4GiB-L9": ("random", 4 << 30, 9, corpus.SEED_C5),
    "mixed-256MiB-L9": ("mixed", 256 << 20, 9, corpus.SEED_C2),
    "random-1GiB-L9": ("random", 1 << 30, 9, corpus.SEED_C5),
    "text-1GiB-L1": ("text", 1 << 30, 1, corpus.SEED_C4),
    "ab-64MiB-L9": ("ab", 64 << 20, 9, 0),
    "period1000-64MiB-L9": ("period1000", 64 << 20, 9, corpus.SEED_C3),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mixed-1GiB-L9", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-replicas", action="store_true", help="skip the one-object-per-GPU extra arm (N > 1)")
    ap.add_argument("--set", action="append", default=[], help="ctx tunable key=value (repeatable)")
    return ap.parse_args()


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)"""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def make_corpus(kind, size, seed, out=None):
    if kind == "ab":
        a = corpus.periodic(size, b"ab")
    elif kind == "period1000":
        a = corpus.periodic(size, corpus.random_bytes(1000, seed=seed))
    else:
        return corpus.by_name(kind, size, seed, out=out)
    if out is not None:
        out[:size] = a
        return out[:size]
    return a


def config_of(args, kind, size, level):
    """identical in both arms (the driver compares them)"""
    return {"workload": args.workload, "level": level, "bytes": size, "corpus": kind,
            "object": "one stream, sharded block-wise over the GPUs",
            "l2": "input and every per-stage array are far larger than the 126 MB L2 (no flush needed)"}


# ---------------------------------------------------------------------------------- clocks

CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
               "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
               "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_indices):
        self.rows = []
        self.proc = None
        self.idx = ",".join(str(i) for i in gpu_indices)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={CLOCK_QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", self.idx],
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

def run_reference(args, kind, size, level, seed):
    """banzai's own CPU algorithm on the box's host cores: the C restatement (oracle/), every block
    on its own thread (oracle.encode_mt), the WHOLE workload per step.  No GPU code is loaded."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import pyoracle as O
    cores = os.cpu_count() or 1
    data = make_corpus(kind, size, seed)
    for _ in range(args.warmup):
        O.encode_mt(data, level, cores, digest=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sha, out_len, nb = O.encode_mt(data, level, cores, digest=True)
    total = time.perf_counter() - t0
    value = size * args.steps / total / 1e6
    desc = (f"the whole workload ({size} bytes, {nb} blocks) per step; C restatement of banzai "
            f"(-O3 -march=native, literal SA-IS), sequential cut chain + one worker thread per block "
            f"on {cores} host threads; no Rust toolchain exists in the image, so banzai itself cannot run")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_of(args, kind, size, level),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "cores": cores,
        "stream": {"sha256": sha, "bytes": out_len, "blocks": nb},
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


def sha_of(ptr, n):
    return hashlib.sha256((C.c_uint8 * n).from_address(ptr.value if hasattr(ptr, "value") else ptr)).hexdigest()


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
    cpu_group = D.CpuGate(group)           # gloo: idle ranks wait on the CPU, not in a spinning NCCL kernel
    lib = _ffi.lib
    n_gpus = world if world > 1 else max(1, args.gpus)
    if n_gpus > torch.cuda.device_count():
        raise SystemExit(f"bench.py: --gpus {n_gpus} but {torch.cuda.device_count()} visible")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            cpu_group.wait()               # rank 0 arrives here only after its timed work
            group.barrier()
            torch.cuda.synchronize()

    def tune(ctx):
        for kv in args.set:
            k, v = kv.split("=")
            ctx.set(k, int(v))

    def timed(fn, steps, warmup, active=True):
        """W warm-up + K timed calls of fn() on the active ranks, barrier + synchronize on both
        sides, wall time = max over ranks"""
        if active:
            for _ in range(warmup):
                fn()
        sync_all()
        t0 = time.perf_counter()
        if active:
            for _ in range(steps):
                fn()
        sync_all()
        return group.max_over_ranks(time.perf_counter() - t0)

    line = None
    h_in = None
    # ------------------------------------------------------------------ strong: one stream over n_gpus devices
    if rank == 0:
        h_in = lib.bnz_host_alloc(size)
        if not h_in:
            raise SystemExit("bnz_host_alloc failed")
        h_arr = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(size,))
        make_corpus(kind, size, seed, out=h_arr)
        ctx = banzai_b200.Context(devices=list(range(n_gpus)))
        tune(ctx)
        res_stats, e2e_stats, outs = [], [], {}

        def step_resident():
            o, n = ctx.encode_ptr(h_in, size, level)
            res_stats.append(ctx.stats())
            outs["n"] = n
            ctx.free_out(o)

        def step_e2e():
            o, n = ctx.encode_ptr(h_in, size, level)
            e2e_stats.append(ctx.stats())
            ctx.free_out(o)
    sampler = ClockSampler(list(range(n_gpus)) if rank == 0 else [local])

    # "value": the input is resident in HBM when the timed region starts (every device keeps its
    # byte range from the previous call: reuse_input), the finished stream lands in host memory
    if rank == 0:
        ctx.set("reuse_input", 1)
        step_resident()                    # makes the input resident (outside the timed region)
        res_stats.clear()
    if rank == 0:
        sampler.start()
    t_res = timed(step_resident if rank == 0 else None, args.steps, args.warmup, active=rank == 0)
    # "e2e": pinned host input -> finished .bz2 in host memory; H2D and D2H inside every step
    if rank == 0:
        ctx.set("reuse_input", 0)
    t_e2e = timed(step_e2e if rank == 0 else None, args.steps, args.warmup, active=rank == 0)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        res_stats = res_stats[-args.steps:]
        e2e_stats = e2e_stats[-args.steps:]
        peak, peak_src = load_peak()
        bwt_ms = statistics.mean(s["bwt_ms"] for s in res_stats)
        alg = statistics.mean(s["bwt_algorithmic_bytes"] for s in res_stats)
        achieved = alg / (bwt_ms * 1e-3) / 1e9
        step_ms = t_res / args.steps * 1e3
        launches = sum(s["kernel_launches"] for s in res_stats) + sum(s["kernel_launches"] for s in e2e_stats)
        last = res_stats[-1]
        line = {
            "metric": METRIC, "value": round(size * args.steps / t_res / 1e6, 1), "unit": UNIT,
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_of(args, kind, size, level),
            "e2e": {"value": round(size * args.steps / t_e2e / 1e6, 1), "unit": UNIT,
                    "ms_per_step": round(t_e2e / args.steps * 1e3, 3),
                    "h2d_bytes_per_step": int(e2e_stats[-1]["h2d_bytes"]),
                    "d2h_bytes_per_step": int(e2e_stats[-1]["d2h_bytes"])},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "bwt_sort_kernel", "bound": "hbm", "achieved": round(achieved, 1),
                         "peak": peak * n_gpus, "peak_source": peak_src + (f" x {n_gpus} GPUs" if n_gpus > 1 else ""),
                         "unit": "GB/s", "frac": round(achieved / (peak * n_gpus), 4),
                         "traffic": load_traffic(args.workload) if n_gpus == 1 else None,
                         "algorithmic_bytes_per_launch": int(alg), "kernel_ms": round(bwt_ms, 3),
                         "share_of_step": round(bwt_ms / step_ms, 3)},
            "details": {"blocks": last["n_blocks"], "devices_used": last["n_devices"],
                        "compressed_bytes": int(outs["n"]),
                        "value_is": "input resident in HBM (reuse_input), stream delivered to host memory",
                        "timing": "wall clock around K synchronous calls, barrier + device synchronize on both "
                                  "sides, max over ranks; stage_ms are CUDA events on the library's streams",
                        "device_ms_per_step": round(statistics.mean(s["total_ms"] for s in res_stats), 3)},
            "stage_ms": {k: round(statistics.mean(s[k] for s in res_stats), 3)
                         for k in ("h2d_ms", "rle_ms", "bwt_ms", "mtf_ms", "huff_ms", "pack_ms", "d2h_ms", "total_ms")},
            "bwt": {"rounds_avg": round(last["bwt_rounds_total"] / max(1, last["n_blocks"]), 2),
                    "rounds_max": last["bwt_max_rounds"],
                    "sum_active_over_n": round(last["bwt_sum_active"] / max(1, last["bwt_n"]), 3),
                    "tied_blocks": last["bwt_tied_blocks"]},
        }
        # ---- outside the timed region: the stream itself
        o, n = ctx.encode_ptr(h_in, size, level)
        sha_n = sha_of(o, n)
        ctx.free_out(o)
        parity = {"sha256": sha_n, "bytes": int(n)}
        if n_gpus > 1:
            with banzai_b200.Context(devices=[0]) as one:
                tune(one)
                o1, n1 = one.encode_ptr(h_in, size, level)
                parity["n_gpu_stream_equals_1_gpu_stream"] = bool(n1 == n and sha_of(o1, n1) == sha_n)
                one.free_out(o1)
        if n_gpus == 1 and not args.no_cpu_baseline:
            # the checker and the CPU baseline in one leg: the oracle (all host threads) encodes the
            # same workload; its stream must be byte-identical to the GPU's
            from oracle import pyoracle as O
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            sha_o, len_o, nb_o = O.encode_mt(h_arr, level, cores, digest=True)
            dt = time.perf_counter() - t0
            parity["oracle_identical"] = bool(sha_o == sha_n and len_o == n)
            line["cpu_baseline"] = {"value": round(size / dt / 1e6, 3), "unit": UNIT, "cores": cores,
                                    "kind": "port",
                                    "sample": f"the whole workload once ({size} bytes, {nb_o} blocks): C "
                                              f"restatement of banzai, one worker thread per block on {cores} "
                                              f"host threads"}
            sample = min(size, 24 * 1000 * 1000)
            t0 = time.perf_counter()
            O.encode(h_arr[:sample], level)
            line["cpu_single_thread"] = {"value": round(sample / (time.perf_counter() - t0) / 1e6, 3), "unit": UNIT,
                                         "cores": 1, "sample": f"first {sample} bytes, one thread"}
        else:
            line["cpu_baseline"] = None
        line["parity_check"] = parity
        ctx.close()

    # ------------------------------------------------------------------ replicas: one object per GPU (extra key)
    if world > 1 and not args.no_replicas:
        if rank != 0:
            h_in = lib.bnz_host_alloc(size)
            if not h_in:
                raise SystemExit("bnz_host_alloc failed")
            h_arr = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(size,))
            make_corpus(kind, size, D.object_seed(seed, rank), out=h_arr)
        rctx = banzai_b200.Context(devices=[local])
        tune(rctx)

        def step_rep():
            o, n = rctx.encode_ptr(h_in, size, level)
            rctx.free_out(o)

        rctx.set("reuse_input", 1)
        step_rep()
        t_rep = timed(step_rep, args.steps, args.warmup)
        rctx.set("reuse_input", 0)
        t_rep_e2e = timed(step_rep, args.steps, args.warmup)
        if rank == 0:
            line["replicas"] = {"what": "every rank encodes its own object on its own GPU (weak scaling, round-1 method)",
                                "value": round(D.aggregate_throughput(size, world, t_rep, args.steps), 1),
                                "e2e": round(D.aggregate_throughput(size, world, t_rep_e2e, args.steps), 1), "unit": UNIT}
        rctx.close()

    if rank == 0:
        print(json.dumps(line), flush=True)
    if h_in:
        lib.bnz_host_free(h_in)
    cpu_group.close()
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

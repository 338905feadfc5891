"""CPU, world_size 2 over gloo: the host logic of the one-process-per-GPU path (rank env, distinct
per-rank objects, barrier, max-over-ranks timing, whole-job aggregation).  The data path itself
has no collective."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json, time
sys.path.insert(0, %r)
import numpy as np
import corpus
from banzai_b200 import dist as D
from oracle import pyoracle as O
g = D.Group(backend="gloo")
assert g.world == 2 and g.rank in (0, 1)
seed = D.object_seed(corpus.SEED_C2, g.rank)
data = corpus.mixed(300000, seed)
g.barrier()
t0 = time.perf_counter()
out = O.encode(data, 9)            # stands in for the per-rank GPU encode on the CPU box
elapsed = time.perf_counter() - t0 + 0.05 * g.rank
tmax = g.max_over_ranks(elapsed)
total_out = g.sum_over_ranks(len(out))
g.barrier()
print(json.dumps({"rank": g.rank, "elapsed": elapsed, "tmax": tmax, "sha": hash(out) & 0xffff,
                  "first": int(data[:64].sum()), "total_out": total_out,
                  "value": D.aggregate_throughput(data.size, g.world, tmax, 1)}))
g.close()
""" % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_gloo():
    import json
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        so, se = p.communicate(timeout=240)
        assert p.returncode == 0, se[-2000:]
        outs.append(json.loads(so.strip().splitlines()[-1]))
    a, b = sorted(outs, key=lambda d: d["rank"])
    assert a["first"] != b["first"]                     # distinct objects per rank
    assert abs(a["tmax"] - b["tmax"]) < 1e-9            # both see the max
    assert a["tmax"] >= max(a["elapsed"], b["elapsed"]) - 1e-9
    assert a["total_out"] == b["total_out"] > 0
    assert abs(a["value"] - 2 * 300000 / a["tmax"] / 1e6) < 1e-6


def test_single_process_defaults():
    from banzai_b200 import dist as D
    env = {k: os.environ.pop(k, None) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        g = D.Group()
        assert (g.rank, g.world, g.local) == (0, 1, 0)
        assert g.max_over_ranks(1.5) == 1.5
        assert D.object_seed(7, 0) == 7 and D.object_seed(7, 1) != 7
        assert D.aggregate_throughput(10**6, 4, 2.0, 3) == 6.0
    finally:
        for k, v in env.items():
            if v is not None:
                os.environ[k] = v


GATE_WORKER = r"""
import os, sys, json, time
sys.path.insert(0, %r)
from banzai_b200 import dist as D
g = D.Group(backend="gloo")
gate = D.CpuGate(g)
# the one-stream measurement of bench.py: rank 0 works, the other ranks wait in the CPU gate
def sync_all():
    gate.wait()
    g.barrier()
sync_all()
t0 = time.perf_counter()
if g.rank == 0:
    time.sleep(0.4)                 # stands in for rank 0 driving every GPU
sync_all()
tmax = g.max_over_ranks(time.perf_counter() - t0)
print(json.dumps({"rank": g.rank, "tmax": tmax}))
gate.close()
g.close()
""" % ROOT


def test_idle_ranks_wait_in_cpu_gate():
    import json
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", GATE_WORKER], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        so, se = p.communicate(timeout=240)
        assert p.returncode == 0, se[-2000:]
        outs.append(json.loads(so.strip().splitlines()[-1]))
    for o in outs:
        assert 0.39 < o["tmax"] < 5.0          # the idle rank was held until rank 0 finished


def test_reference_arm_runs_without_the_product_library():
    """bench.py --impl reference: whole workload through the block-parallel oracle driver, and the
    product package is never imported on that arm"""
    import json
    code = ("import sys; sys.argv=['bench.py','--impl','reference','--workload','text-10MB-L9','--steps','1',"
            "'--warmup','0']; sys.path.insert(0, %r); import bench; bench.main(); "
            "assert not any(m.startswith('banzai_b200') for m in sys.modules), 'product imported'" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["scaling"] == "strong"
    assert line["config"]["workload"] == "text-10MB-L9" and line["config"]["bytes"] == 10 * 1000 * 1000
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and line["gpu_launches"] == 0
    assert line["stream"]["blocks"] >= 11

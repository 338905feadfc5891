"""wall time and stage times of bnz_encode from a pinned host buffer (the bench's e2e path) against the
knobs of the piecewise upload: python tools/e2e_probe.py"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus, banzai_b200
from banzai_b200 import _ffi
n = 1 << 30
data = corpus.mixed(n, seed=corpus.SEED_C2)
hp = _ffi.lib.bnz_host_alloc(n)
C.memmove(hp, data.ctypes.data, n)
for sets in ({"h2d_pieces": 3, "piece_blocks_per_sm_x16": 7}, {"h2d_pieces": 3, "piece_blocks_per_sm_x16": 5},
             {"h2d_pieces": 3, "piece_blocks_per_sm_x16": 9}, {"h2d_pieces": 3, "piece_blocks_per_sm_x16": 12},
             {"h2d_pieces": 2, "piece_blocks_per_sm_x16": 7}, {"h2d_pieces": 4, "piece_blocks_per_sm_x16": 7},
             {"h2d_pieces": 3, "piece_blocks_per_sm_x16": 7, "bwt_cluster_below": 100}):
    ctx = banzai_b200.Context(n_gpus=1)
    for k, v in sets.items(): ctx.set(k, v)
    ts = []
    for rep in range(6):
        t0 = time.perf_counter(); out, olen = ctx.encode_ptr(hp, n, 9); dt = time.perf_counter() - t0
        ctx.free_out(out)
        ts.append(dt)
    ts = sorted(ts[2:])
    st = ctx.stats()
    print(sets, "wall min %.1f median %.1f ms" % (ts[0] * 1e3, ts[len(ts) // 2] * 1e3), flush=True)
    ctx.close()

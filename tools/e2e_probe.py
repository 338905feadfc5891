"""wall time and stage times of bnz_encode from a pinned host buffer (the bench's e2e path)"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus, banzai_b200
from banzai_b200 import _ffi
n = 1 << 30
data = corpus.mixed(n, seed=corpus.SEED_C2) if hasattr(corpus, "SEED_C2") else corpus.mixed(n)
hp = _ffi.lib.bnz_host_alloc(n)
C.memmove(hp, data.ctypes.data, n)
for sets in ({"h2d_overlap": 0}, {"h2d_pieces": 2}, {"h2d_pieces": 3}, {"h2d_pieces": 4}, {"h2d_pieces": 6}):
    ctx = banzai_b200.Context(n_gpus=1)
    for k, v in sets.items(): ctx.set(k, v)
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter(); out, olen = ctx.encode_ptr(hp, n, 9); dt = time.perf_counter() - t0
        ctx.free_out(out)
        best = min(best, dt)
    st = ctx.stats()
    print(sets, "wall %.1f ms" % (best * 1e3), {k: round(st[k], 1) for k in ("h2d_ms", "rle_ms", "bwt_ms", "mtf_ms", "huff_ms", "pack_ms", "d2h_ms", "total_ms")}, flush=True)
    ctx.close()

"""radix-pass phase times (thread PROF_TID of every CTA, default 37 = warp 1 lane 5) from the profiling
build tools/micro/libbanzai_prof.so (`make -C banzai_b200/csrc prof [PROF_TID=n]`): tma wait / rank / B1 wait / scan (B1->B2) / place (B2->B3) / store"""
import sys, os, ctypes as C
os.environ["BANZAI_B200_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro", "libbanzai_prof.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import corpus, banzai_b200
from banzai_b200 import _ffi
kind = sys.argv[1]; nb = int(sys.argv[2])
blk = 899999
data = corpus.by_name(kind, nb * blk)
blocks = [data[i * blk:(i + 1) * blk] for i in range(nb)]
names = ["zero+loop", "tma wait", "rank", "B1 wait", "scan B1->B2", "place B2->B3", "store"]
for cps in (2, 1):
    ctx = banzai_b200.Context(n_gpus=1)
    ctx.set("bwt_cluster", 0); ctx.set("bwt_ctas_per_sm", cps)
    ctx.stage_bwt(blocks, 9)
    out = (C.c_ulonglong * 8)()
    _ffi.lib.bnz_prof_read(out)
    ctx.stage_bwt(blocks, 9)
    _ffi.lib.bnz_prof_read(out)
    st = ctx.stats()
    tot = sum(out[:7])
    print("ctas/sm", cps, "bwt_ms %.1f" % st["bwt_ms"], "radix cyc %.2f G" % (st["bwt_cyc_radix"] / 1e9), "tile-loop cyc %.2f G" % (tot / 1e9))
    tiles = st["bwt_sum_active_passes"] / 4096
    for n, v in zip(names, out[:7]):
        print("   %-14s %5.1f%%  %7.0f cycles/tile" % (n, 100.0 * v / tot, v / tiles))
    ctx.close()

import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus, banzai_b200
kind = sys.argv[1]; nb = int(sys.argv[2])
blk = 899999
data = corpus.by_name(kind, nb * blk)
blocks = [data[i * blk:(i + 1) * blk] for i in range(nb)]
ctx = banzai_b200.Context(n_gpus=1)
cl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx.set("bwt_cluster", cl)
for kv in sys.argv[4:]:
    k, v = kv.split("="); ctx.set(k, int(v))
best = 1e9
for _ in range(3):
    ctx.stage_bwt(blocks, 9); best = min(best, ctx.stats()["bwt_ms"])
st = ctx.stats()
print(os.environ.get("BANZAI_B200_LIB", "default"), kind, nb, "cluster", cl, " ".join(sys.argv[4:]), "bwt_ms %.2f" % best, "cyc build/radix/rerank %.2f/%.2f/%.2f G" % (st["bwt_cyc_build"] / 1e9, st["bwt_cyc_radix"] / 1e9, st["bwt_cyc_rerank"] / 1e9))

"""How much does the dynamic-queue tail cost?  Per-block cycle counts from the one-CTA kernel, then
simulate 296 workers: index order (what the kernel does) vs longest-first."""
import sys, os, heapq
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import corpus, banzai_b200
from banzai_b200 import _ffi

blk = 899999
nb = 1098
data = corpus.mixed(nb * blk + 1000)
ctx = banzai_b200.Context(n_gpus=1)
ctx.set("bwt_cluster", 0)
blocks = [data[i * blk:(i + 1) * blk] for i in range(nb)]
for lpt in (0, 1, 0, 1):
    ctx.set("bwt_lpt", lpt)
    res = ctx.stage_bwt(blocks, 9, with_stats=True)
    print("lpt", lpt, "bwt_ms", ctx.stats()["bwt_ms"])
ctx.set("bwt_lpt", 0)
res = ctx.stage_bwt(blocks, 9, with_stats=True)
cost = np.array([r[3]["cycles"] for r in res], dtype=float)
proxy = np.array([r[3]["sum_active"] * 5 + 900000 * r[3]["rounds"] * 0.3 for r in res], dtype=float)
print("corr(cycles, proxy)", np.corrcoef(cost, proxy)[0, 1])
def sim(order, workers=296):
    h = [0.0] * workers
    heapq.heapify(h)
    for i in order:
        t = heapq.heappop(h)
        heapq.heappush(h, t + cost[i])
    return max(h)
base = sim(range(nb))
lpt = sim(np.argsort(-cost))
print("ideal", cost.sum() / 296, "index order", base, "longest first", lpt, "tail overhead %.1f%% -> %.1f%%" % ((base / (cost.sum() / 296) - 1) * 100, (lpt / (cost.sum() / 296) - 1) * 100))
print("cost min/median/max", cost.min(), np.median(cost), cost.max())
rounds = np.array([r[3]["rounds"] for r in res])
print("rounds hist", np.bincount(rounds))

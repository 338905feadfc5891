import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus, banzai_b200
blk = 899999
nb = 1098
data = corpus.mixed(nb * blk + 1000)
ctx = banzai_b200.Context(n_gpus=1)
ctx.set("bwt_cluster", 0); ctx.set("bwt_lpt", 0)
blocks = [data[i * blk:(i + 1) * blk] for i in range(nb)]
res = ctx.stage_bwt(blocks, 9, with_stats=True)
cyc = np.array([r[3]["cycles"] for r in res], dtype=float)
rounds = np.array([r[3]["rounds"] for r in res])
sa = np.array([r[3]["sum_active"] for r in res], dtype=float)
ctx.set("bwt_lpt", 1)
res2 = ctx.stage_bwt(blocks, 9, with_stats=True)
score = np.array([r[3]["score"] for r in res2], dtype=float)
cyc2 = np.array([r[3]["cycles"] for r in res2], dtype=float)
print("lpt run bwt_ms", ctx.stats()["bwt_ms"], "corr(score, sum_active) %.3f corr(score, cycles_lpt0) %.3f" % (np.corrcoef(score, sa)[0, 1], np.corrcoef(score, cyc)[0, 1]))
print("scores first 30:", score[:30].astype(int).tolist())
print("sum cycles lpt0 %.3g lpt1 %.3g" % (cyc.sum(), cyc2.sum()))
import heapq
def sim(order, cost, workers=296):
    h = [0.0] * workers; heapq.heapify(h)
    for i in order:
        t = heapq.heappop(h); heapq.heappush(h, t + cost[i])
    return max(h)
order1 = np.argsort(-score, kind="stable")
print("sim index-order with lpt0 cycles: %.0fM; sim score-order with lpt1 cycles: %.0fM; ideal lpt0 %.0fM lpt1 %.0fM" % (
    sim(range(nb), cyc) / 1e6, sim(order1, cyc2) / 1e6, cyc.sum() / 296e6, cyc2.sum() / 296e6))
for name, sel in (("text", rounds == 4), ("binary", rounds == 5), ("source", rounds >= 13)):
    print(name, sel.sum(), "mean Mcyc lpt0 %.1f lpt1 %.1f" % (cyc[sel].mean() / 1e6, cyc2[sel].mean() / 1e6))
sys.exit(0)
idx = np.arange(0, nb, 37)
def rep_frac(b, w, step=1):
    a = np.lib.stride_tricks.sliding_window_view(b, w)[::step]
    v = np.ascontiguousarray(a).view(np.dtype((np.void, w))).ravel()
    _, cnt = np.unique(v, return_counts=True)
    return (cnt[cnt > 1].sum()) / v.size
for w in (8, 16, 24, 48, 96):
    f = np.array([rep_frac(blocks[i], w, 4) for i in idx])
    print("window", w, "corr with cycles %.3f" % np.corrcoef(f, cyc[idx])[0, 1], "with rounds %.3f" % np.corrcoef(f, rounds[idx])[0, 1], "with sum_active %.3f" % np.corrcoef(f, sa[idx])[0, 1])
print("corr(cycles, sum_active) %.3f corr(cycles, rounds) %.3f" % (np.corrcoef(cyc, sa)[0, 1], np.corrcoef(cyc, rounds)[0, 1]))
for k in range(0, 30):
    print(k, rounds[k], int(sa[k]), int(cyc[k] / 1e6), end=" | ")

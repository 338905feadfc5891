"""one-CTA-per-block kernel vs cluster kernel by number of blocks on the device (mixed corpus, level 9)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import corpus, banzai_b200
blk = 899999
data = corpus.mixed(600 * blk)
ctx = banzai_b200.Context(n_gpus=1)
for nb in (12, 37, 74, 137, 148, 200, 296, 400, 600):
    blocks = [data[i * blk:(i + 1) * blk] for i in range(nb)]
    res = []
    for clu in (0, 4, 8, 16):
        ctx.set("bwt_cluster", clu)
        best = 1e9
        for _ in range(2):
            ctx.stage_bwt(blocks, 9)
            best = min(best, ctx.stats()["bwt_ms"])
        res.append(f"c{clu}={best:.1f}ms")
    print(nb, " ".join(res), flush=True)

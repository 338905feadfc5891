"""8-bit vs 10-bit digits at 1 and 2 CTAs per SM (one-CTA-per-block kernel)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import corpus, banzai_b200
kind = sys.argv[1]; nb = int(sys.argv[2])
blk = 899999
data = corpus.by_name(kind, nb * blk)
blocks = [data[i * blk:(i + 1) * blk] for i in range(nb)]
for bits, cps in ((8, 2), (8, 1), (10, 1), (8, 2)):
    ctx = banzai_b200.Context(n_gpus=1)
    ctx.set("bwt_cluster", 0); ctx.set("bwt_radix_bits", bits); ctx.set("bwt_ctas_per_sm", cps)
    best = 1e9
    for _ in range(2):
        ctx.stage_bwt(blocks, 9); best = min(best, ctx.stats()["bwt_ms"])
    st = ctx.stats()
    print(kind, nb, "bits", bits, "ctas/sm", cps, "bwt_ms %.2f" % best, "cyc build/radix/rerank %.2f/%.2f/%.2f G" % (st["bwt_cyc_build"] / 1e9, st["bwt_cyc_radix"] / 1e9, st["bwt_cyc_rerank"] / 1e9))
    ctx.close()

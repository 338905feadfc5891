import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import corpus, banzai_b200
kind = sys.argv[1]; nb = int(sys.argv[2]); cps = int(sys.argv[3]); clu = int(sys.argv[4]) if len(sys.argv) > 4 else 8
blk = 899999
data = corpus.by_name(kind, nb * blk)
blocks = [data[i * blk:(i + 1) * blk] for i in range(nb)]
ctx = banzai_b200.Context(n_gpus=1)
ctx.set("bwt_radix_bits", 8); ctx.set("bwt_ctas_per_sm", cps); ctx.set("bwt_cluster", clu)
ctx.stage_bwt(blocks, 9)
print(ctx.stats()["bwt_ms"])

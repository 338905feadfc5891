"""Quick BWT-stage timing on the GPU box (not the bench): python tools/bwt_perf.py [kind] [nblocks] [level]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus, banzai_b200

kind = sys.argv[1] if len(sys.argv) > 1 else "text"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 296
level = int(sys.argv[3]) if len(sys.argv) > 3 else 9
blk = 100000 * level - 1
data = corpus.by_name(kind, nb * blk) if kind != "ab" else corpus.periodic(nb * blk, b"ab")
blocks = [data[i * blk:(i + 1) * blk] for i in range(nb)]
ctx = banzai_b200.Context(n_gpus=1)
for bits, cps, clu, thr in ((8, 0, 0, 512), (8, 1, 8, 512), (8, 0, 8, 512), (8, 0, 4, 512), (8, 0, 16, 512), (8, 0, 8, 1024)):
    if True:
        ctx.set("bwt_radix_bits", bits)
        ctx.set("bwt_ctas_per_sm", cps)
        ctx.set("bwt_cluster", clu)
        ctx.set("bwt_threads", thr)
        best = None
        for it in range(3):
            ctx.stage_bwt(blocks, level)
            st = ctx.stats()
            if best is None or st["bwt_ms"] < best["bwt_ms"]:
                best = st
        gbs = best["bwt_algorithmic_bytes"] / best["bwt_ms"] / 1e6
        print(f"{kind} L{level} nb={nb} cluster={clu} thr={thr} ctas/sm={cps}: bwt {best['bwt_ms']:.2f} ms, "
              f"{best['bwt_n'] / best['bwt_ms'] / 1e6:.2f} GB/s input, alg {gbs:.0f} GB/s "
              f"({gbs / 6550 * 100:.1f}% of 6550), rounds avg {best['bwt_rounds_total'] / nb:.1f} max {best['bwt_max_rounds']}, "
              f"sum_active/n {best['bwt_sum_active'] / best['bwt_n']:.2f}, passes/rec {best['bwt_sum_active_passes'] / best['bwt_sum_active']:.2f}, "
              f"cyc build/radix/rerank {best['bwt_cyc_build'] / 1e9:.2f}/{best['bwt_cyc_radix'] / 1e9:.2f}/{best['bwt_cyc_rerank'] / 1e9:.2f} G",
              flush=True)

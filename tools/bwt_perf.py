"""Quick BWT-stage timing on the GPU box (not the bench): python tools/bwt_perf.py [kind] [nblocks] [level] [modes]
modes: comma list of cluster sizes (0 = one CTA per block), default "0,8" """
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import corpus, banzai_b200

kind = sys.argv[1] if len(sys.argv) > 1 else "mixed"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 296
level = int(sys.argv[3]) if len(sys.argv) > 3 else 9
modes = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "0,8").split(",")]
sets = [kv.split("=") for kv in sys.argv[5:]]
blk = 100000 * level - 1
if kind == "ab":
    data = corpus.periodic(nb * blk, b"ab")
elif kind == "period1000":
    data = corpus.periodic(nb * blk, corpus.random_bytes(1000, seed=corpus.SEED_C3))
else:
    data = corpus.by_name(kind, nb * blk)
blocks = [data[i * blk:(i + 1) * blk] for i in range(nb)]
ctx = banzai_b200.Context(n_gpus=1)
for k, v in sets:
    ctx.set(k, int(v))
for clu in modes:
    ctx.set("bwt_cluster", clu)
    best = None
    for it in range(3):
        ctx.stage_bwt(blocks, level)
        st = ctx.stats()
        if best is None or st["bwt_ms"] < best["bwt_ms"]:
            best = st
    gbs = best["bwt_algorithmic_bytes"] / best["bwt_ms"] / 1e6
    print(f"{kind} L{level} nb={nb} cluster={clu}: bwt {best['bwt_ms']:.2f} ms, "
          f"{best['bwt_n'] / best['bwt_ms'] / 1e6:.2f} GB/s input, alg {gbs:.0f} GB/s "
          f"({gbs / 6550 * 100:.1f}% of 6550), rounds avg {best['bwt_rounds_total'] / nb:.1f} max {best['bwt_max_rounds']}, "
          f"sum_active/n {best['bwt_sum_active'] / best['bwt_n']:.2f} (in smem {best['bwt_sum_tile'] / best['bwt_n']:.2f}), "
          f"HBM passes/n {best['bwt_sum_active_passes'] / best['bwt_n']:.2f}, "
          f"Mcyc/block build/radix/rerank/tile/final {best['bwt_cyc_build'] / 1e6 / nb:.1f}/{best['bwt_cyc_radix'] / 1e6 / nb:.1f}/"
          f"{best['bwt_cyc_rerank'] / 1e6 / nb:.1f}/{best['bwt_cyc_tile'] / 1e6 / nb:.1f}/{best['bwt_cyc_final'] / 1e6 / nb:.1f}", flush=True)

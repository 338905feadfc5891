"""Workload for compute-sanitizer (tools/calls/sanitize.sh): the smoke encode (every kernel of the
path once at levels 1 and 9, both sort kernels, the streaming front end, the sharded path) and a
batch of 700 small blocks through both sort kernels.  Everything is checked against the oracle."""
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import banzai_b200  # noqa: E402
import corpus  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
data = corpus.mixed(300000)
if which in ("all", "encode"):
    with banzai_b200.Context(n_gpus=1) as ctx:
        for level in (1, 9):
            assert ctx.encode_bytes(data, level) == O.encode(data, level)
        ctx.set("bwt_cluster", 0)
        assert ctx.encode_bytes(data, 9) == O.encode(data, 9)
        sink = io.BytesIO()
        ctx.encode_stream(io.BytesIO(data.tobytes()), sink, 9)
        assert sink.getvalue() == O.encode(data, 9)
    with banzai_b200.Context(devices=[0, 0]) as ctx:
        assert ctx.encode_bytes(data, 1) == O.encode(data, 1)
    with banzai_b200.Context(n_gpus=1) as ctx:            # self-verification kernels, literal Huffman loop
        ctx.set("verify", 1)
        ctx.set("huff_literal", 1)
        for cl in (0, 8):
            ctx.set("bwt_cluster", cl)
            assert ctx.encode_bytes(data, 9) == O.encode(data, 9)
        periodic = (b"ab" * 200000)[:399999] + bytes(150000)
        assert ctx.encode_bytes(periodic, 5) == O.encode(periodic, 5)
if which in ("all", "bwt"):
    nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 700
    src = corpus.mixed(nblk * 3000)
    blocks = [src[i * 3000:(i + 1) * 3000].tobytes() for i in range(nblk)]
    blocks += [b"abcdefg" * 500, bytes(4000), (b"ab" * 3000)[:4999], (b"ab" * 40000)[:79999],
               (corpus.random_bytes(1000, seed=3).tobytes() * 80)[:79999]]
    with banzai_b200.Context(n_gpus=1) as ctx:
        for cl in (0, 8):
            ctx.set("bwt_cluster", cl)
            got = ctx.stage_bwt(blocks, 1)
            for blk, g in zip(blocks, got):
                ebw, eptr, _ = O.bwt(blk)
                assert g[1] == eptr and bytes(g[0]) == bytes(ebw)
print("sanitize target ok:", which)

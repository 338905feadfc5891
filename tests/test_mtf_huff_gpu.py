"""GPU parity: K5 (`mtf::mtf_and_rle`, reference lib/mtf.rs:14) and K6-K8 (`huffman::encode`,
reference lib/huffman.rs:313) vs the oracle, stage by stage.  Bit-exact."""
import json
import os

import numpy as np
import pytest

import corpus
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


@pytest.fixture(scope="module")
def ctx():
    import banzai_b200
    c = banzai_b200.Context(n_gpus=1)
    yield c
    c.close()


def _bwt_blocks(datas):
    out = []
    for d in datas:
        bw, _, has = O.bwt(d)
        out.append((bw, has))
    return out


def _check_mtf(ctx, blocks):
    got = ctx.stage_mtf([b for b, _ in blocks], [h for _, h in blocks])
    for (bw, has), (syms, num_syms, freqs) in zip(blocks, got):
        es, en, ef = O.mtf_and_rle(bw, has)
        assert num_syms == en
        assert syms.size == es.size
        assert (syms == es).all()
        assert (freqs.astype(np.uint64) == ef).all()
    return got


def test_mtf_reference_vector(ctx):
    k = KATS["mtf_smoke"]
    buf = np.array(k["input"], dtype=np.uint8)
    has = np.zeros(256, np.uint8)
    has[buf] = 1
    got = ctx.stage_mtf([buf], [has])
    assert got[0][0].tolist() == k["expected"] and got[0][1] == 42


def test_mtf_small_and_edge(ctx):
    raw = [b"x", b"aa", b"ab", bytes(5000), b"ab" * 3000, bytes(range(256)) * 9,
           bytes(range(255, -1, -1)) * 5, b"abcdefg" * 1000, bytes([0] * 1023 + [1] + [0] * 1025 + [2]),
           bytes([5] * 1024), bytes([5] * 1025), bytes([5] * 2048 + [6])]
    # mtf operates on arbitrary byte strings: feed them directly with their own has_byte
    blocks = []
    for r in raw:
        a = np.frombuffer(r, dtype=np.uint8)
        has = np.zeros(256, np.uint8)
        has[a] = 1
        blocks.append((a, has))
    _check_mtf(ctx, blocks)


@pytest.mark.parametrize("kind", ["text", "source", "binary", "random"])
def test_mtf_full_blocks(ctx, kind):
    data = corpus.by_name(kind, 899999 + 300000)
    _check_mtf(ctx, _bwt_blocks([data[:899999], data[899999:]]))


def _check_huff(ctx, mtfs):
    got = ctx.stage_huffman([m[0] for m in mtfs], [m[1] for m in mtfs], [m[2] for m in mtfs])
    for (syms, num_syms, freqs), (bits, bit_len, tables, nt) in zip(mtfs, got):
        ent, etab, esel = O.huffman_model(syms, num_syms, freqs)
        ebits, ebit_len = O.huffman_encode(syms, num_syms, freqs)
        assert nt == ent
        assert (tables == etab).all()
        assert bit_len == ebit_len
        assert bytes(bits) == bytes(ebits)


@pytest.mark.parametrize("kind", ["text", "source", "binary", "random"])
def test_huffman_full_blocks(ctx, kind):
    data = corpus.by_name(kind, 899999 + 200000)
    mtfs = []
    for d in (data[:899999], data[899999:]):
        bw, _, has = O.bwt(d)
        s, n, f = O.mtf_and_rle(bw, has)
        mtfs.append((s, n, f))
    _check_huff(ctx, mtfs)


def test_huffman_small_alphabets_and_tiny_blocks(ctx):
    mtfs = []
    for raw in (b"a", b"aaaa", b"hello world", bytes(1000), b"ab" * 500, bytes(range(256)),
                bytes(range(199)) * 3, bytes(range(198)) * 3, bytes(range(197)) * 3, b"abcdefg" * 1000):
        bw, _, has = O.bwt(raw)
        s, n, f = O.mtf_and_rle(bw, has)
        mtfs.append((s, n, f))
    _check_huff(ctx, mtfs)


def test_huffman_literal_refinement_loop_on_device(ctx):
    """bnz_ctx_set("huff_literal", 1): the device runs huffman::encode's four refinement rounds
    literally (lib/huffman.rs:399-460: per-group cost and strict-< argmin over all tables,
    table_freqs accumulation, rebuild, selectors of the last round) instead of their closed form.
    Tables, selectors and bits must be the oracle's (whose loop is the literal one)."""
    mtfs = []
    for kind in ("text", "binary", "random"):
        data = corpus.by_name(kind, 899999 + 100000)
        for d in (data[:899999], data[899999:]):
            bw, _, has = O.bwt(d)
            mtfs.append(O.mtf_and_rle(bw, has))
    for raw in (b"a", b"hello world", bytes(1000), bytes(range(199)) * 3, bytes(range(198)) * 3):
        bw, _, has = O.bwt(raw)
        mtfs.append(O.mtf_and_rle(bw, has))
    ctx.set("huff_literal", 1)
    try:
        _check_huff(ctx, mtfs)
    finally:
        ctx.set("huff_literal", 0)


@pytest.mark.parametrize("level", [1, 9])
def test_stream_with_literal_refinement(level):
    import banzai_b200
    data = corpus.mixed(6 * 1000 * 1000 + 17)
    want = O.encode_mt(data, level)
    with banzai_b200.Context(n_gpus=1) as c:
        c.set("huff_literal", 1)
        assert c.encode_bytes(data, level) == want
    with banzai_b200.Context(devices=[0, 0]) as c:
        c.set("huff_literal", 1)
        assert c.encode_bytes(data, level) == want

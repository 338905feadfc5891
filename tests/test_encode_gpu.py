"""GPU parity, whole stream: bnz_encode through the C ABI must be byte-identical to the oracle's
restatement of `banzai::encode` (reference lib/lib.rs:84-132), and decode with libbz2."""
import bz2
import hashlib
import io
import json
import os

import numpy as np
import pytest

import corpus
from oracle import pyoracle as O
from tests.golden.make_vectors import vector_input

pytestmark = pytest.mark.gpu

VECS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_vectors.json")))


@pytest.fixture(scope="module")
def ctx():
    import banzai_b200
    c = banzai_b200.Context(n_gpus=1)
    yield c
    c.close()


@pytest.mark.parametrize("name", sorted(VECS.keys() - {"_comment"}, key=lambda s: int(s[1:])))
def test_survey_vectors(ctx, name):
    v = VECS[name]
    out = ctx.encode_bytes(vector_input(name), v["level"])
    if "hex" in v:
        assert out.hex() == v["hex"]
    else:
        assert len(out) == v["len"] and hashlib.sha256(out).hexdigest() == v["sha256"]


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_levels_mixed(ctx, level):
    data = corpus.mixed(3 * 1000 * 1000 + 17)
    got = ctx.encode_bytes(data, level)
    assert got == O.encode(data, level)
    assert bz2.decompress(got) == data.tobytes()


@pytest.mark.parametrize("kind", ["text", "source", "binary", "random"])
def test_corpora_level9(ctx, kind):
    data = corpus.by_name(kind, 4 * 1000 * 1000)
    got = ctx.encode_bytes(data, 9)
    assert got == O.encode(data, 9)


def test_config0_10mb_text_level9(ctx):
    """BASELINE.json configs[0]: 10 MB English-like text at level 9, byte-identical"""
    data = corpus.text(10 * 1000 * 1000, seed=corpus.SEED_C1)
    got = ctx.encode_bytes(data, 9)
    want = O.encode(data, 9)
    assert got == want
    assert ctx.stats()["n_blocks"] == 12


def test_degenerate_and_periodic(ctx):
    unit = corpus.random_bytes(1000, seed=corpus.SEED_C3).tobytes()
    cases = [bytes(3 * 1000 * 1000), b"ab" * 1000000, b"abcdefg" * 300000, unit * 2000,
             b"abcdefg" * 1000, unit * 1000, b"aa", b"aaaab" * 400000, bytes(range(256)) * 4000]
    for data in cases:
        for level in (1, 9):
            got = ctx.encode_bytes(data, level)
            assert got == O.encode(data, level), (len(data), level)
            assert bz2.decompress(got) == data


def test_tiny_inputs(ctx):
    rng = np.random.default_rng(9)
    for n in list(range(0, 40)) + [255, 256, 257, 1023, 1024, 1025, 4095, 4097]:
        data = rng.integers(0, 4, n).astype(np.uint8).tobytes()
        assert ctx.encode_bytes(data, 9) == O.encode(data, 9), n


def test_repeated_calls_reuse_context(ctx):
    a = corpus.text(500000, seed=1)
    b = corpus.random_bytes(2000000, seed=2)
    for _ in range(2):
        assert ctx.encode_bytes(a, 9) == O.encode(a, 9)
        assert ctx.encode_bytes(b, 1) == O.encode(b, 1)


def test_level_out_of_range_is_einval(ctx):
    import banzai_b200
    for level in (0, 10, -1):
        with pytest.raises(banzai_b200.BanzaiError):
            ctx.encode_bytes(b"abc", level)


def test_reference_api_mirror_encode_and_encode_file(tmp_path):
    """encode(reader, writer, level) / encode_file(in, out) keep the reference's contract"""
    import banzai_b200
    data = corpus.source(700000).tobytes()
    sink = io.BytesIO()
    n = banzai_b200.encode(io.BytesIO(data), sink, 5)
    assert n == len(data)
    assert sink.getvalue() == O.encode(data, 5)
    src, dst = tmp_path / "in.bin", tmp_path / "out.bz2"
    src.write_bytes(data)
    assert banzai_b200.encode_file(str(src), str(dst)) == len(data)
    assert dst.read_bytes() == O.encode(data, 9)


def test_size_independent_properties_at_scale(ctx):
    """64 MiB mixed corpus at level 9: too big for a quick oracle run, so check properties:
    libbz2 decodes it back (every block CRC + the combined CRC verify) and a second run is
    bit-identical."""
    data = corpus.mixed(64 << 20)
    out1 = ctx.encode_bytes(data, 9)
    assert bz2.decompress(out1) == data.tobytes()
    out2 = ctx.encode_bytes(data, 9)
    assert out1 == out2


def test_device_resident_entry():
    """bnz_encode_device (input and stream stay in device memory) gives the bytes of the plain call;
    it needs a one-device context and a 16-byte aligned input, and says so"""
    import ctypes as C

    import banzai_b200
    from banzai_b200 import _ffi
    from banzai_b200.api import BanzaiError
    data = corpus.mixed(9 * 1000 * 1000)
    want = O.encode(data, 9)
    lib = _ffi.lib
    n = data.size
    with banzai_b200.Context(devices=[0]) as c:
        assert c.encode_bytes(data, 9) == want
        d_in = lib.bnz_device_alloc(c._h, n + 64)
        cap = len(want) + 8                       # the documented minimum: stream size + 8
        d_out = lib.bnz_device_alloc(c._h, cap)
        c._check(lib.bnz_memcpy_h2d(c._h, d_in, data.ctypes.data_as(C.c_void_p), n))
        olen = c.encode_device(d_in, data.ctypes.data_as(C.c_void_p), n, 9, d_out, cap)
        back = np.zeros(olen, np.uint8)
        c._check(lib.bnz_memcpy_d2h(c._h, back.ctypes.data_as(C.c_void_p), d_out, olen))
        assert back.tobytes() == want
        with pytest.raises(BanzaiError):          # misaligned device input
            c.encode_device(C.c_void_p(d_in + 1), data.ctypes.data_as(C.c_void_p), n - 1, 9, d_out, cap)
        with pytest.raises(BanzaiError):          # output buffer too small
            c.encode_device(d_in, data.ctypes.data_as(C.c_void_p), n, 9, d_out, len(want))
        assert c.encode_bytes(data, 9) == want    # the context is still usable
        lib.bnz_device_free(c._h, d_in)
        lib.bnz_device_free(c._h, d_out)
    with banzai_b200.Context(devices=[0, 0]) as c2:
        d_in = lib.bnz_device_alloc(c2._h, n + 64)
        d_out = lib.bnz_device_alloc(c2._h, n)
        with pytest.raises(BanzaiError):
            c2.encode_device(d_in, data.ctypes.data_as(C.c_void_p), n, 9, d_out, n)
        lib.bnz_device_free(c2._h, d_in)
        lib.bnz_device_free(c2._h, d_out)


def test_streaming_batches_do_not_change_the_stream():
    """inputs above max_batch_bytes are encoded batch by batch (bounded device memory); block cuts
    and bit offsets must chain across batches exactly"""
    import banzai_b200
    rng = np.random.default_rng(4)
    mixed = corpus.mixed(23 * 1000 * 1000 + 5)
    runs = np.repeat(rng.integers(0, 3, 4000), rng.integers(1, 30000, 4000)).astype(np.uint8)[:30 * 1000 * 1000]
    zeros_then_text = np.concatenate([np.zeros(20 << 20, np.uint8), corpus.text(3000000), np.zeros(9 << 20, np.uint8)])
    with banzai_b200.Context(n_gpus=1) as one, banzai_b200.Context(n_gpus=1) as ctx:
        ctx.set("max_batch_bytes", 4 << 20)
        for data, level in ((mixed, 9), (mixed, 1), (runs, 9), (zeros_then_text, 9), (zeros_then_text, 1)):
            want = one.encode_bytes(data, level)
            got = ctx.encode_bytes(data, level)
            assert got == want, (data.size, level)
        small = corpus.text(100000)
        assert ctx.encode_bytes(small, 9) == O.encode(small, 9)
        assert ctx.encode_bytes(mixed[:9000000], 3) == O.encode(mixed[:9000000], 3)


@pytest.mark.parametrize("sets", [{}, {"mtf_overlap": 95, "mtf_groups": 6}, {"mtf_overlap": 0, "crc_low_prio": 0}],
                         ids=["default", "many-lists", "serial"])
def test_overlap_of_sort_tail_with_mtf_and_crc(sets):
    """more blocks than the one-CTA-per-block sort has CTAs (and than the cluster threshold): the
    MTF of finished blocks and the block CRCs are queued beside the sort on low-priority streams
    (DESIGN §5 "Overlap").  Whatever the schedule, the stream is the oracle's."""
    import banzai_b200
    data = corpus.mixed(52 * 1000 * 1000)
    with banzai_b200.Context(n_gpus=1) as c:
        for k, v in sets.items():
            c.set(k, v)
        got = c.encode_bytes(data, 1)
        assert c.stats()["n_blocks"] > 450
        again = c.encode_bytes(data, 1)
    assert got == again
    assert got == O.encode(data, 1)
    assert bz2.decompress(got) == data.tobytes()


@pytest.mark.parametrize("level,size", [(1, 40 * 1000 * 1000), (9, 30 * 1000 * 1000)])
def test_upload_in_pieces(level, size):
    """one GPU, host input: the input is uploaded and cut in two pieces and the first piece's
    blocks are sorted while the second arrives.  Forced here at small sizes (the
    automatic mode needs ~750 MB at level 9); the stream must be the plain path's and the oracle's."""
    import banzai_b200
    data = corpus.mixed(size + 4321)
    want = O.encode(data, level)
    with banzai_b200.Context(n_gpus=1) as c:
        c.set("h2d_overlap", 0)
        plain = c.encode_bytes(data, level)
        c.set("h2d_overlap", 2)
        got = c.encode_bytes(data, level)
        assert c.stats()["n_devices"] >= 2          # several lanes on the same GPU
        again = c.encode_bytes(data, level)
    assert plain == want
    assert got == want
    assert again == want
    # long runs across the piece boundary (8 MiB): the second piece's tables continue the first's
    runs = bytes(7 * 1000 * 1000) + b"\x01" * (3 * 1000 * 1000) + bytes(range(256)) * 40000 + bytes(9 * 1000 * 1000)
    for pieces in (2, 3, 5):
        with banzai_b200.Context(n_gpus=1) as c:
            c.set("h2d_overlap", 2)
            c.set("h2d_pieces", pieces)
            assert c.encode_bytes(runs, 1) == O.encode(runs, 1)
    # a first piece without a single complete block (level 9 zeros: one block swallows 46 MB)
    zeros = bytes(70 * 1000 * 1000) + b"tail"
    with banzai_b200.Context(n_gpus=1) as c:
        c.set("h2d_overlap", 2)
        assert c.encode_bytes(zeros, 9) == O.encode(zeros, 9)


def test_upload_in_pieces_automatic_mode():
    """the automatic mode (>= 256 MiB and at least four first pieces' worth of input; here 300 MB
    at level 3, three lanes) gives the bytes of the single-upload path"""
    import banzai_b200
    data = corpus.mixed(300 * 1000 * 1000)
    with banzai_b200.Context(n_gpus=1) as c:
        got = c.encode_bytes(data, 3)
        assert c.stats()["n_devices"] == 3          # three lanes on the one GPU
        c.set("h2d_overlap", 0)
        plain = c.encode_bytes(data, 3)
        assert c.stats()["n_devices"] == 1
    assert got == plain
    assert got == O.encode_mt(data, 3)

"""Self-verification (csrc/verify.cu): with "verify" on, every block's RLE1 image is
decoded back to its input bytes and its (BWT, origPtr) is inverted back to the RLE1 image on the
device, the cut chain and the CRCs are re-derived on the host — the role libbz2 plays for the
reference in fuzz/fuzz_targets/round_trip.rs:8-22.  A damaged intermediate must be caught."""
import pytest

import corpus
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


def _inputs():
    unit = corpus.random_bytes(1000, seed=corpus.SEED_C3).tobytes()
    return [
        (corpus.mixed(7 * 1000 * 1000 + 123).tobytes(), 9),
        (corpus.mixed(3 * 1000 * 1000).tobytes(), 1),
        (b"abcdefg" * 1000, 9),                      # identical rotations (one backward chain)
        (unit * 1000, 9),                            # period | n, full-size block
        (bytes(3 * 1000 * 1000) + b"tail", 2),       # long runs: RLE1 count bytes, truncated runs at the cuts
        (b"aaaab" * 300000, 1),
        (b"a", 9), (b"", 9), (b"hello world", 1),
    ]


@pytest.mark.parametrize("cluster", [0, 8])
def test_verify_passes_and_leaves_the_stream_unchanged(cluster):
    import banzai_b200
    with banzai_b200.Context(n_gpus=1) as ctx:
        ctx.set("verify", 1)
        ctx.set("bwt_cluster", cluster)
        for data, level in _inputs():
            assert ctx.encode_bytes(data, level) == O.encode_mt(data, level), (len(data), level)


def test_verify_on_the_sharded_and_piecewise_paths():
    import banzai_b200
    data = corpus.mixed(30 * 1000 * 1000 + 77)
    want = O.encode_mt(data, 1)
    with banzai_b200.Context(devices=[0, 0, 0]) as ctx:
        ctx.set("verify", 1)
        assert ctx.encode_bytes(data, 1) == want
    with banzai_b200.Context(n_gpus=1) as ctx:
        ctx.set("verify", 1)
        ctx.set("h2d_overlap", 2)
        assert ctx.encode_bytes(data, 1) == want
        assert ctx.stats()["n_devices"] >= 2


@pytest.mark.parametrize("cluster", [0, 8])
@pytest.mark.parametrize("what", [1, 2, 3, 4], ids=["bwt-byte", "origptr", "rle1-byte", "crc"])
def test_corrupted_intermediate_is_caught(what, cluster):
    import banzai_b200
    from banzai_b200 import _ffi
    from banzai_b200.api import BanzaiError
    data = corpus.mixed(5 * 1000 * 1000)
    want = O.encode_mt(data, 9)
    with banzai_b200.Context(n_gpus=1) as ctx:
        ctx.set("bwt_cluster", cluster)
        ctx.set("verify", 1)
        ctx.set("verify_corrupt", what)
        with pytest.raises(BanzaiError) as e:
            ctx.encode_bytes(data, 9)
        assert e.value.code == _ffi.EVERIFY
        ctx.set("verify_corrupt", 0)
        assert ctx.encode_bytes(data, 9) == want          # the context is still usable
        # without "verify" the same damage goes unnoticed by the encoder (libbz2 would reject the stream)
        if what in (1, 2, 3):
            ctx.set("verify", 0)
            ctx.set("verify_corrupt", what)
            assert ctx.encode_bytes(data, 9) == want      # (the hook only acts inside the verification step)

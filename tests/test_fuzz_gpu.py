"""GPU fuzzing in the spirit of the reference's fuzz targets (fuzz/fuzz_targets/encode.rs: no crash
at level 1; round_trip.rs: libbz2 decodes to the input) — plus the stronger check that the stream
is byte-identical to the oracle.  Inputs are biased towards what breaks bzip2 encoders: runs around
the 4/255/256 thresholds, tiny alphabets, periodic data, block-capacity boundaries."""
import bz2

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import banzai_b200
    c = banzai_b200.Context(n_gpus=1)
    yield c
    c.close()


run_piece = st.tuples(st.integers(0, 255), st.sampled_from([1, 2, 3, 4, 5, 6, 254, 255, 256, 257, 509, 510, 511, 1000]))
pieces = st.lists(st.one_of(
    st.binary(min_size=0, max_size=300),
    run_piece.map(lambda t: bytes([t[0]]) * t[1]),
    st.tuples(st.binary(min_size=1, max_size=9), st.integers(1, 400)).map(lambda t: t[0] * t[1]),
), min_size=0, max_size=12).map(b"".join)


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
@given(data=pieces, level=st.sampled_from([1, 9]))
def test_small_inputs_match_oracle_and_round_trip(ctx, data, level):
    got = ctx.encode_bytes(data, level)
    assert got == O.encode(data, level)
    assert bz2.decompress(got) == data


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 32 - 1), tail=pieces)
def test_block_boundary_inputs(ctx, seed, tail):
    """level 1: ~one block of noise followed by a fuzzed tail, so the cut lands inside the tail"""
    rng = np.random.default_rng(seed)
    head = rng.integers(0, 256, 99999 - int(rng.integers(0, 60))).astype(np.uint8).tobytes()
    data = head + tail + bytes([7]) * int(rng.integers(0, 600)) + tail
    got = ctx.encode_bytes(data, 1)
    assert got == O.encode(data, 1)


@settings(max_examples=10, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 32 - 1), mid=pieces, cuts=st.lists(st.integers(1, 3_000_000), min_size=1, max_size=12))
def test_streaming_front_end_any_write_sizes(seed, mid, cuts):
    """bnz_stream_write with fuzzed write sizes over ~3 minimum-size windows at level 1 (the
    reference's reader-chunk-dependent truncation, SURVEY A-Q4, must not exist here): the stream
    equals the whole-buffer call, which equals the oracle."""
    import ctypes as C
    import banzai_b200
    from banzai_b200 import _ffi
    rng = np.random.default_rng(seed)
    n = int(rng.integers(11_000_000, 13_000_000))
    body = rng.integers(0, 4, n).astype(np.uint8)                     # small alphabet: runs cross windows
    body[rng.integers(0, n, 2000)] = 200
    zero_at = int(rng.integers(0, n - 6_000_000))
    body[zero_at:zero_at + int(rng.integers(0, 6_000_000))] = 0        # a run longer than the headroom of a window
    data = body.tobytes()[:n // 2] + mid + body.tobytes()[n // 2:]
    with banzai_b200.Context(n_gpus=1) as c:
        c.set("stream_window_bytes", 1 << 16)                         # clamped to the minimum (~5.1 MB at level 1)
        chunks = []
        cb = _ffi.SINK_FN(lambda u, p, k: chunks.append(bytes((C.c_ubyte * k).from_address(p))) or 0)
        h = C.c_void_p()
        assert _ffi.lib.bnz_stream_open(c._h, 1, cb, None, C.byref(h)) == _ffi.OK
        pos, i = 0, 0
        while pos < len(data):
            k = min(cuts[i % len(cuts)], len(data) - pos)
            assert _ffi.lib.bnz_stream_write(h, data[pos:pos + k], k) == _ffi.OK
            pos += k
            i += 1
        used = C.c_size_t()
        assert _ffi.lib.bnz_stream_finish(h, C.byref(used)) == _ffi.OK
        _ffi.lib.bnz_stream_close(h)
        assert used.value == len(data)
        got = b"".join(chunks)
        assert got == c.encode_bytes(data, 1)
    assert got == O.encode(data, 1)

"""GPU parity of the streaming front end (bnz_stream_*): whatever the window size
and the sizes of the caller's reads, the stream bytes are those of `banzai::encode` over the whole
input (reference lib/lib.rs:84-132; the reference refills its reader in lib/rle.rs:43-91)."""
import bz2
import ctypes as C
import io
import os
import subprocess

import numpy as np
import pytest

import corpus
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ChunkyReader:
    """reader without readinto that returns ragged short reads, like a pipe"""

    def __init__(self, data, seed):
        self.data, self.pos = data, 0
        self.rng = np.random.default_rng(seed)

    def read(self, n=-1):
        if n is None or n < 0:
            n = len(self.data) - self.pos
        n = min(n, int(self.rng.integers(1, 3 << 20)))
        out = self.data[self.pos:self.pos + n]
        self.pos += len(out)
        return out


@pytest.fixture()
def ctx():
    import banzai_b200
    c = banzai_b200.Context(n_gpus=1)
    yield c
    c.close()


def test_many_windows_level1_against_oracle(ctx):
    """30 MB mixed corpus at level 1 through ~6 MB windows: block cuts, the bit phase carried
    across windows and the footer must reproduce the oracle's stream"""
    data = corpus.mixed(30 * 1000 * 1000 + 123).tobytes()
    ctx.set("stream_window_bytes", 1 << 16)          # clamped to one block's worst-case input (~5 MB)
    sink = io.BytesIO()
    n = ctx.encode_stream(ChunkyReader(data, 1), sink, 1)
    assert n == len(data)
    want = O.encode(data, 1)
    assert sink.getvalue() == want
    # readinto path, other read sizes
    sink2 = io.BytesIO()
    assert ctx.encode_stream(io.BytesIO(data), sink2, 1, read_size=777777) == len(data)
    assert sink2.getvalue() == want


def test_windows_level9_equal_whole_buffer_call(ctx):
    """level 9, 150 MB, 46 MB windows (the minimum): same bytes as one bnz_encode"""
    data = corpus.mixed(150 * 1000 * 1000).tobytes()
    whole = ctx.encode_bytes(data, 9)
    ctx.set("stream_window_bytes", 1 << 16)
    sink = io.BytesIO()
    assert ctx.encode_stream(io.BytesIO(data), sink, 9) == len(data)
    assert sink.getvalue() == whole
    assert bz2.decompress(whole) == data


@pytest.mark.parametrize("level", [1, 9])
def test_long_runs_leave_large_tails(ctx, level):
    """all-zero and long-run inputs: one block swallows up to 51x its capacity in input, so the
    partial block carried between windows is as large as the headroom allows"""
    n = 48 * 1000 * 1000 if level == 1 else 200 * 1000 * 1000
    data = bytearray(n)
    data[n // 3:n // 3 + 1000] = bytes(range(250)) * 4
    data[n // 2:] = b"\x07" * (n - n // 2)
    data = bytes(data)
    ctx.set("stream_window_bytes", 1 << 16)
    sink = io.BytesIO()
    assert ctx.encode_stream(ChunkyReader(data, 2), sink, level) == n
    assert sink.getvalue() == ctx.encode_bytes(data, level)
    assert bz2.decompress(sink.getvalue()) == data


@pytest.mark.parametrize("data", [b"", b"a", b"hello world", b"aaaa" * 1000], ids=["empty", "a", "hello", "runs"])
def test_tiny_streams(ctx, data):
    sink = io.BytesIO()
    assert ctx.encode_stream(io.BytesIO(data), sink, 3) == len(data)
    assert sink.getvalue() == O.encode(data, 3)


def test_default_window_single_job(ctx):
    data = corpus.text(5 * 1000 * 1000).tobytes()
    sink = io.BytesIO()
    assert ctx.encode_stream(io.BytesIO(data), sink, 9) == len(data)
    assert sink.getvalue() == O.encode(data, 9)


def test_sink_failure_is_eio_and_context_survives(ctx):
    import banzai_b200
    from banzai_b200 import _ffi

    class Broken:
        def write(self, b):
            raise OSError("disk full")

    data = corpus.text(300000).tobytes()
    with pytest.raises(OSError):
        ctx.encode_stream(io.BytesIO(data), Broken(), 9)
    # raw ABI: a failing sink gives BNZ_EIO from finish
    cb = _ffi.SINK_FN(lambda u, p, n: 1)
    h = C.c_void_p()
    assert _ffi.lib.bnz_stream_open(ctx._h, 9, cb, None, C.byref(h)) == _ffi.OK
    # one open stream per context
    h2 = C.c_void_p()
    assert _ffi.lib.bnz_stream_open(ctx._h, 9, cb, None, C.byref(h2)) == _ffi.EINVAL
    assert _ffi.lib.bnz_stream_write(h, data, len(data)) == _ffi.OK
    assert _ffi.lib.bnz_stream_finish(h, None) == _ffi.EIO
    _ffi.lib.bnz_stream_close(h)
    assert _ffi.lib.bnz_stream_open(ctx._h, 0, cb, None, C.byref(h)) == _ffi.EINVAL      # lib.rs:89
    # the context is usable afterwards
    assert ctx.encode_bytes(data, 9) == O.encode(data, 9)
    # abandoning a stream mid-way is allowed
    assert _ffi.lib.bnz_stream_open(ctx._h, 1, _ffi.SINK_FN(lambda u, p, n: 0), None, C.byref(h)) == _ffi.OK
    assert _ffi.lib.bnz_stream_write(h, data, len(data)) == _ffi.OK
    _ffi.lib.bnz_stream_close(h)
    assert isinstance(banzai_b200.BanzaiError(_ffi.EIO).args[0], str)


def test_encode_file_and_cli_pipe(tmp_path):
    """bnz_encode_file (lib/lib.rs:141-153) and `bnz -c -` on a pipe stream through the same path"""
    from banzai_b200 import _ffi
    import banzai_b200
    data = corpus.source(12 * 1000 * 1000).tobytes()
    src, dst = tmp_path / "in.bin", tmp_path / "out.bz2"
    src.write_bytes(data)
    with banzai_b200.Context(n_gpus=1) as c:
        c.set("stream_window_bytes", 1 << 16)
        used = C.c_size_t()
        assert _ffi.lib.bnz_encode_file(c._h, str(src).encode(), str(dst).encode(), C.byref(used)) == _ffi.OK
        assert used.value == len(data)
        want = c.encode_bytes(data, 9)
        assert dst.read_bytes() == want
        assert _ffi.lib.bnz_encode_file(c._h, str(tmp_path / "missing").encode(), str(dst).encode(), None) == _ffi.EIO
    exe = os.path.join(ROOT, "banzai_b200", "bnz")
    p = subprocess.run([exe, "-c", "-"], input=data, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert p.returncode == 0, p.stderr
    assert p.stdout == want


def test_final_window_large_enough_for_the_piecewise_upload(ctx):
    """a (final) window of >= 256 MiB takes the piecewise-upload path inside the stream's worker
    thread; same bytes as the whole-buffer call"""
    data = corpus.mixed(300 * 1000 * 1000).tobytes()
    sink = io.BytesIO()
    assert ctx.encode_stream(io.BytesIO(data), sink, 1) == len(data)
    assert ctx.stats()["n_devices"] == 3
    ctx.set("h2d_overlap", 0)
    assert sink.getvalue() == ctx.encode_bytes(data, 1)

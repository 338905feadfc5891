"""Multi-GPU: blocks are sharded over the devices of one process, no collective; the
stream must not depend on the number of devices.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import bz2

import pytest

import corpus
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_ngpu() < 2")
@pytest.mark.parametrize("level", [1, 9])
def test_two_devices_same_bytes_as_oracle(level):
    import banzai_b200
    data = corpus.mixed(7 * 1000 * 1000 + 123)
    want = O.encode(data, level)
    with banzai_b200.Context(n_gpus=2) as ctx:
        got = ctx.encode_bytes(data, level)
        assert ctx.stats()["n_devices"] == 2
    assert got == want


@pytest.mark.skipif("_ngpu() < 2")
def test_all_devices_large_and_tiny():
    import banzai_b200
    n = _ngpu()
    with banzai_b200.Context(n_gpus=0) as ctx, banzai_b200.Context(n_gpus=1) as one:
        data = corpus.mixed(40 << 20)
        a = ctx.encode_bytes(data, 9)
        assert a == one.encode_bytes(data, 9)
        assert bz2.decompress(a) == data.tobytes()
        assert 2 <= ctx.stats()["n_devices"] <= min(n, ctx.stats()["n_blocks"])
        for tiny in (b"", b"a", b"hello world", bytes(2000000)):
            assert ctx.encode_bytes(tiny, 9) == O.encode(tiny, 9)
        # runs crossing shard boundaries
        zeros = bytes(60 << 20) + b"ab" * 3000000 + bytes(30 << 20)
        assert ctx.encode_bytes(zeros, 1) == one.encode_bytes(zeros, 1)


@pytest.mark.skipif("_ngpu() < 2")
def test_streaming_front_end_on_two_devices():
    """bnz_stream_* windows sharded over two GPUs: same bytes as one bnz_encode on one GPU"""
    import io
    import banzai_b200
    data = corpus.mixed(60 * 1000 * 1000).tobytes()
    with banzai_b200.Context(n_gpus=1) as one:
        want = one.encode_bytes(data, 2)
    with banzai_b200.Context(n_gpus=2) as ctx:
        ctx.set("stream_window_bytes", 1 << 16)          # minimum-size windows (~10 MB at level 2)
        sink = io.BytesIO()
        assert ctx.encode_stream(io.BytesIO(data), sink, 2) == len(data)
        assert ctx.stats()["n_devices"] == 2
    assert sink.getvalue() == want
    assert bz2.decompress(want) == data


@pytest.mark.skipif("_ngpu() < 2")
def test_upload_in_pieces_on_a_device_other_than_0():
    """the lanes of the piecewise upload run in their own host threads, which must select the
    context's device (one process per GPU under torchrun: LOCAL_RANK > 0)"""
    import banzai_b200
    data = corpus.mixed(30 * 1000 * 1000 + 77)
    want = O.encode(data, 1)
    with banzai_b200.Context(devices=[1]) as ctx:
        ctx.set("h2d_overlap", 2)
        assert ctx.encode_bytes(data, 1) == want
        import io
        sink = io.BytesIO()
        ctx.set("stream_window_bytes", 1 << 16)
        assert ctx.encode_stream(io.BytesIO(data.tobytes()), sink, 1) == len(data)
        assert sink.getvalue() == want

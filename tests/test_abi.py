"""CPU: the C-ABI library loads and exports every symbol include/banzai_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "banzai_b200.h")).read()
    return sorted(set(re.findall(r"BNZ_API\s+[^;(]*?\b(bnz_\w+)\s*\(", hdr)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("bnz_ctx_create", "bnz_encode", "bnz_free", "bnz_stage_bwt", "bnz_stage_rle1",
                 "bnz_stage_mtf", "bnz_stage_huffman", "bnz_strerror"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "banzai_b200", "libbanzai_b200.so"))
    for name in _declared():
        assert hasattr(lib, name), name
    from banzai_b200 import _ffi
    assert sorted(_ffi.EXPORTS) == _declared()


def test_error_strings_and_arg_checks_without_gpu():
    from banzai_b200 import _ffi
    assert _ffi.lib.bnz_strerror(0) == b"ok"
    assert b"level" in _ffi.lib.bnz_strerror(1)
    assert _ffi.lib.bnz_max_compressed_size(1000) > 1000
    h = ctypes.c_void_p()
    assert _ffi.lib.bnz_ctx_create(ctypes.byref(h), -1) == _ffi.EINVAL

"""CPU: libpa_b200.so loads, exports every symbol include/pa_b200.h declares, and rejects bad
arguments with the documented status codes before touching a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cabi():
    import importlib.util
    pkg = os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200")
    spec = importlib.util.spec_from_file_location("pa_build", os.path.join(pkg, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from llm_decoder import _cabi
    return _cabi


def test_header_symbols_exported(cabi):
    hdr = open(os.path.join(ROOT, "include", "pa_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    handle = C.CDLL(cabi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, f"declared in pa_b200.h but not exported: {missing}"
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)


def test_version_and_error_strings(cabi):
    lib = cabi.lib()
    assert lib.pa_version() >= 100
    assert lib.pa_error_string(0) == b"ok"
    for st in (-1, -2, -3, -4):
        assert b"pa_b200" in lib.pa_error_string(st)


def test_invalid_arguments_rejected_without_device(cabi):
    lib = cabi.lib()
    assert lib.pa_page_table_clear(None, 10, None) == -1
    assert lib.pa_page_table_update(None, 10, None, None, 1, None) == -1
    assert lib.pa_quantize_i8(None, 16, 1.0, None, None) == -1
    assert lib.pa_quantize_i8(None, 0, 1.0, None, None) == 0          # empty input is a no-op
    assert lib.pa_kv_gather(None, None, None, 1, 1, 1, 1, 16, 64, 2, None, 1, 0, None) == -1
    # decode: null q; unsupported head_dim; unsupported tile_size; zero temperature
    common = [0x1000, 1, 1, 1, 1, None, None, 1, 16]
    assert lib.pa_paged_decode_f16(None, 0x1000, 0x1000, 0x1000, *common, 128, 16, 1.0, None, None, None, 0, None) == -1
    assert lib.pa_paged_decode_f16(0x1000, 0x1000, 0x1000, 0x1000, *common, 96, 16, 1.0, None, None, None, 0, None) == -2
    assert lib.pa_paged_decode_f16(0x1000, 0x1000, 0x1000, 0x1000, *common, 128, 8, 1.0, None, None, None, 0, None) == -2
    assert lib.pa_paged_decode_f16(0x1000, 0x1000, 0x1000, 0x1000, *common, 128, 16, 0.0, None, None, None, 0, None) == -1
    assert lib.pa_decode_workspace_bytes(64, 32, 128, 256, 16) > 0
    assert lib.pa_decode_workspace_bytes(-1, 32, 128, 256, 16) == 0


def test_missing_library_fails_loudly(cabi, monkeypatch):
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", "/nonexistent/libpa_b200.so")
    with pytest.raises(cabi.PAError):
        cabi.lib()


def test_packed_linear_geometry_and_argument_rules_need_no_gpu(cabi):
    """pa_linear_pack_bytes is pure geometry ([ceil(N/128)][ceil(K/32)][32][128] f32) and pa_linear_f32_packed rejects
    what its kernel cannot take before touching the device (K % 4 != 0 -> PA_ERR_UNSUPPORTED; null pointers and unknown
    activations -> PA_ERR_INVALID_ARG)."""
    lib = cabi.lib()
    assert lib.pa_linear_pack_bytes(32, 128) == 32 * 128 * 4
    assert lib.pa_linear_pack_bytes(33, 129) == 2 * 2 * 32 * 128 * 4
    assert lib.pa_linear_pack_bytes(4096, 11008) == 86 * 128 * 32 * 128 * 4
    assert lib.pa_linear_pack_bytes(0, 5) == 0
    # (the pointers are never dereferenced: the argument rules come first)
    assert lib.pa_linear_f32_packed(0x1000, 0x1000, None, 4, 6, 8, 0, 0x2000, None, 0, None) == -2
    assert lib.pa_linear_f32_packed(None, 0x1000, None, 4, 8, 8, 0, 0x2000, None, 0, None) == -1
    assert lib.pa_linear_f32_packed(0x1000, 0x1000, None, 4, 8, 8, 7, 0x2000, None, 0, None) == -1

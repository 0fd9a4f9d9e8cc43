"""The C-ABI library loads without a GPU and exports every symbol include/b200kv.h declares."""
import ctypes
import os
import re

from b200kv import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200kv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200kv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 35
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/b200kv.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_strerror():
    l = _lib.lib()
    assert l.b200kv_abi_version() == 1
    assert l.b200kv_strerror(0) == b"ok"
    assert b"pool full" in l.b200kv_strerror(_lib.ENOSPC)


def test_engine_fails_loudly_without_gpu():
    """No CPU fallback: creating an engine without a CUDA device is an error, not a slow path."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from b200kv import B200KVError, KVEngine, KVGeometry
    with pytest.raises(B200KVError) as ei:
        KVEngine(KVGeometry(2, 2, 64, 8), None, 0, staging_bytes=0)
    assert ei.value.code == _lib.ENODEV


def test_chunk_bytes_formula():
    from b200kv import FMT_FP8, FMT_RAW, KVGeometry
    g = KVGeometry(32, 8, 128, 8192)
    assert g.chunk_bytes == 32 * 2 * 256 * 8 * 128 * 2 == 32 << 20          # 32 MiB
    assert g.payload_bytes_per_token == 131072
    g8 = KVGeometry(32, 8, 128, 8192, fmt=FMT_FP8)
    assert g8.chunk_bytes == (16 << 20) + 2048                               # 16 MiB + 2 KiB scales
    assert KVGeometry(32, 8, 128, 8192, fmt=FMT_RAW).stride == 32768

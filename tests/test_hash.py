"""XXH64 / chunk keys: product (C ABI) vs oracle (pure Python and C) vs the xxhash wheel."""
import json
import os

import numpy as np
import pytest

import b200kv
from oracle import kv_oracle as ko
from tests import oracle_c

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "xxh64_vectors.json")))


def test_xxh64_golden_vectors():
    for v in VEC:
        data = bytes.fromhex(v["hex"])
        assert ko.xxh64(data, v["seed"]) == v["digest"], "oracle differs from xxhash wheel fixture"
        assert b200kv.xxh64(data, v["seed"]) == v["digest"], "libb200kv differs from xxhash wheel fixture"
        buf = np.frombuffer(data, dtype=np.uint8)
        got = oracle_c.lib().oracle_xxh64(buf.ctypes.data if len(buf) else None, len(buf), v["seed"])
        assert got == v["digest"]


def test_xxh64_live_wheel():
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(1)
    for n in (0, 1, 31, 32, 33, 1024, 4096 + 7):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert b200kv.xxh64(data, 99) == xxhash.xxh64(data, seed=99).intdigest()


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 512, 700, 2048, 2049])
@pytest.mark.parametrize("partial", [True, False])
def test_chunk_keys_match_oracle(n, partial):
    rng = np.random.default_rng(n)
    toks = rng.integers(0, 128256, n).astype(np.int32)
    want = ko.chunk_keys(toks, 256, 7, partial)
    got = b200kv.chunk_keys(toks, 256, 7, partial)
    assert [int(k) for k in got] == want
    out = np.zeros(max(1, (n + 255) // 256), dtype=np.uint64)
    k = oracle_c.lib().oracle_chunk_keys(toks.ctypes.data, n, 256, 7, int(partial), out.ctypes.data)
    assert [int(x) for x in out[:k]] == want


def test_chunk_keys_are_prefix_chained():
    a = np.arange(1024, dtype=np.int32)
    b = a.copy()
    b[300] += 1  # differs inside chunk 1
    ka, kb = b200kv.chunk_keys(a, 256, 0), b200kv.chunk_keys(b, 256, 0)
    assert ka[0] == kb[0] and all(ka[i] != kb[i] for i in (1, 2, 3))
    assert b200kv.chunk_keys(a, 256, 1)[0] != ka[0]  # seed namespaces keys

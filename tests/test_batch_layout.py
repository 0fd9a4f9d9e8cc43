"""Host-side layout of a step-batched op (KVEngine._batch_layout): what b200kv_store_batch_async /
b200kv_load_batch_async are handed — requests back to back, every request's last chunk padded to C slot entries,
one key and one token count per chunk, request boundaries as chunk offsets.  CPU only (no engine is created)."""
import numpy as np

from b200kv import KVGeometry, chunk_keys
from b200kv.engine import KVEngine

C = 64


def mk():
    eng = object.__new__(KVEngine)               # the layout code needs the geometry and the key seed only
    eng.geom = KVGeometry(2, 2, 8, 64, 16, C)
    eng.key_seed = 1234
    return eng


def test_layout_of_three_requests_with_ragged_tails_and_a_skipped_prefix():
    eng = mk()
    rng = np.random.default_rng(0)
    toks = [rng.integers(0, 1000, n).astype(np.int32) for n in (C + 5, 2 * C, 3 * C + 1)]
    sms = [np.arange(len(t), dtype=np.int64) + 1000 * i for i, t in enumerate(toks)]
    keys, ct, first, sm = eng._batch_layout([(toks[0], sms[0], 0), (toks[1], sms[1], 1), (toks[2], sms[2], 2)])
    assert list(first) == [0, 2, 3, 5]                               # 2 + 1 + 2 chunks
    assert list(ct) == [C, 5, C, C, 1]
    assert len(sm) == 5 * C
    # chunk c's slots sit at [c*C, c*C + ct[c]); padding is -1 and never read
    assert list(sm[0:C]) == list(sms[0][:C]) and list(sm[C:C + 5]) == list(sms[0][C:]) and (sm[C + 5:2 * C] == -1).all()
    assert list(sm[2 * C:3 * C]) == list(sms[1][C:])                 # request 1 starts at its chunk 1
    assert list(sm[3 * C:4 * C]) == list(sms[2][2 * C:3 * C]) and sm[4 * C] == sms[2][3 * C] and (sm[4 * C + 1:] == -1).all()
    # keys are the requests' own prefix-chained keys from the first participating chunk on
    want = np.concatenate([chunk_keys(toks[0], C, 1234)[0:], chunk_keys(toks[1], C, 1234)[1:], chunk_keys(toks[2], C, 1234)[2:]])
    assert np.array_equal(keys, want)
    assert keys.dtype == np.uint64 and ct.dtype == np.int32 and first.dtype == np.int32 and sm.dtype == np.int64


def test_store_batch_rejects_unaligned_offsets_and_skips_finished_requests():
    import pytest
    eng = mk()
    t = np.arange(2 * C, dtype=np.int32)
    with pytest.raises(ValueError):
        eng.store_batch([(t, np.arange(2 * C), 10)])
    assert eng.store_batch([(t, np.arange(2 * C), 2 * C)]) == 0      # nothing beyond the offset: no op at all
    got, ticket = eng.retrieve_batch([(t, np.arange(2 * C), 2 * C)])
    assert list(got) == [0] and ticket == 0

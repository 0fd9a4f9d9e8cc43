"""Model-based test of the pool index (hypothesis): random sequences of the writer / reader / lookup
protocol against a small Python model of the documented semantics — presence, token counts, LRU eviction
order (touch on commit, acquire and lookup hit), pins, and the structural self-check after every step."""
from collections import OrderedDict

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from b200kv import B200KVError, KVPool, _lib

SLOT, N_SLOTS = 64, 4
KEYS = st.integers(min_value=1, max_value=9)
OPS = st.lists(st.one_of(
    st.tuples(st.just("put"), KEYS, st.integers(1, 256)),
    st.tuples(st.just("reserve"), KEYS, st.integers(1, 256)),
    st.tuples(st.just("commit"), KEYS),
    st.tuples(st.just("abort"), KEYS),
    st.tuples(st.just("acquire"), KEYS),
    st.tuples(st.just("release"), KEYS),
    st.tuples(st.just("lookup"), st.lists(KEYS, min_size=1, max_size=4, unique=True)),
), min_size=1, max_size=60)


class Model:
    def __init__(self):
        self.ready = OrderedDict()      # key -> n_tokens, oldest first (LRU order)
        self.writing = {}               # key -> n_tokens
        self.pins = {}

    def used(self):
        return len(self.ready) + len(self.writing)

    def evictable(self):
        for k in self.ready:
            if self.pins.get(k, 0) == 0:
                return k
        return None

    def reserve(self, k, n):
        if k in self.ready or k in self.writing:
            return _lib.EEXIST
        if self.used() >= N_SLOTS:
            v = self.evictable()
            if v is None:
                return _lib.ENOSPC
            del self.ready[v]
        self.writing[k] = n
        return 0


def code_of(fn, *a):
    try:
        fn(*a)
        return 0
    except B200KVError as e:
        return e.code


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(OPS)
def test_pool_follows_the_model(ops):
    p = KVPool(None, N_SLOTS * SLOT, SLOT, _lib.POOL_CREATE)
    m = Model()
    try:
        for op in ops:
            kind = op[0]
            if kind in ("put", "reserve"):
                _, k, n = op
                want = m.reserve(k, n)
                assert code_of(p.reserve, k, n) == want, op
                if want == 0 and kind == "put":
                    p.commit(k)
                    m.ready[k] = m.writing.pop(k)
            elif kind == "commit":
                k = op[1]
                want = 0 if k in m.writing else _lib.ENOENT
                assert code_of(p.commit, k) == want, op
                if want == 0:
                    m.ready[k] = m.writing.pop(k)
            elif kind == "abort":
                k = op[1]
                want = 0 if k in m.writing else _lib.ENOENT
                assert code_of(p.abort, k) == want, op
                m.writing.pop(k, None)
            elif kind == "acquire":
                k = op[1]
                if k in m.ready:
                    slot, n, _ = p.acquire(k)
                    assert n == m.ready[k], op
                    m.pins[k] = m.pins.get(k, 0) + 1
                    m.ready.move_to_end(k)
                else:
                    assert code_of(p.acquire, k) == _lib.ENOENT, op
            elif kind == "release":
                k = op[1]
                want = 0 if (k in m.ready and m.pins.get(k, 0) > 0) else _lib.ENOENT
                assert code_of(p.release, k) == want, op
                if want == 0:
                    m.pins[k] -= 1
            else:
                keys = op[1]
                ct = np.array([m.ready.get(k, m.writing.get(k, 7)) for k in keys], np.int32)
                hits = 0
                for k in keys:                       # longest READY prefix, each hit touched
                    if k not in m.ready:
                        break
                    m.ready.move_to_end(k)
                    hits += 1
                got = p.lookup(np.array(keys, np.uint64), ct)
                assert got[0] == hits, op
            assert p.check(), op
            assert p.stats()["n_used"] == m.used(), op
    finally:
        p.close()

"""Host pool + index (CPU): writer/reader protocol, LRU, leases, pins, cross-process sharing."""
import multiprocessing as mp
import time

import numpy as np
import pytest

from b200kv import B200KVError, KVPool, _lib

SLOT = 4096


def mk(n_slots=4, name=None):
    return KVPool(name, n_slots * SLOT, SLOT, _lib.POOL_CREATE)


def put(pool, key, n_tok=256, fill=None, owner=0):
    slot = pool.reserve(key, n_tok, 0, owner)
    if fill is not None:
        pool.slot_view(slot)[:] = fill
    pool.commit(key)
    return slot


def test_reserve_commit_lookup_prefix_semantics():
    p = mk()
    keys = np.array([11, 22, 33], dtype=np.uint64)
    ct = np.array([256, 256, 100], dtype=np.int32)
    assert p.lookup(keys, ct) == (0, 0)
    put(p, 11)
    put(p, 33, 100)
    assert p.lookup(keys, ct) == (1, 256)          # stops at the first miss (22)
    s = p.reserve(22, 256)
    assert p.lookup(keys, ct) == (1, 256)          # WRITING chunks are invisible
    p.commit(22)
    assert p.lookup(keys, ct) == (3, 612)
    ct_bad = np.array([256, 256, 101], dtype=np.int32)
    assert p.lookup(keys, ct_bad) == (2, 512)      # partial chunk must match its token count
    st = p.stats()
    assert st["n_used"] == 3 and st["n_stored_chunks"] == 3 and st["n_slots"] == 4
    assert s < 4
    p.close()


def test_duplicate_reserve_and_abort():
    p = mk()
    p.reserve(5, 256)
    with pytest.raises(B200KVError) as ei:
        p.reserve(5, 256)
    assert ei.value.code == _lib.EEXIST
    p.abort(5)
    assert p.stats()["n_used"] == 0
    put(p, 5)
    with pytest.raises(B200KVError):
        p.commit(5)                                # already READY
    p.close()


def test_lru_eviction_order_and_touch():
    p = mk(3)
    for k in (1, 2, 3):
        put(p, k)
    p.lookup(np.array([1], np.uint64), np.array([256], np.int32))   # touch 1 -> LRU order 2,3,1
    put(p, 4)                                                        # evicts 2
    one = lambda k: p.lookup(np.array([k], np.uint64), np.array([256], np.int32))[0]
    assert one(2) == 0 and one(3) == 1 and one(1) == 1 and one(4) == 1
    assert p.stats()["n_evicted_chunks"] == 1
    p.close()


def test_pins_and_leases_block_eviction():
    p = mk(2)
    put(p, 1)
    put(p, 2)
    slot, n, fmt = p.acquire(1)
    assert n == 256 and fmt == 0
    p.lookup(np.array([2], np.uint64), np.array([256], np.int32), lease_ms=60000)
    with pytest.raises(B200KVError) as ei:
        p.reserve(3, 256)                          # 1 pinned, 2 leased -> nothing evictable
    assert ei.value.code == _lib.ENOSPC and p.stats()["n_dropped_chunks"] == 1
    p.release(1)
    put(p, 3)                                      # evicts 1 (unpinned), not the leased 2
    assert p.lookup(np.array([2], np.uint64), np.array([256], np.int32))[0] == 1
    assert p.lookup(np.array([1], np.uint64), np.array([256], np.int32))[0] == 0
    p.close()


def test_lease_expires():
    p = mk(1)
    put(p, 1)
    p.lookup(np.array([1], np.uint64), np.array([256], np.int32), lease_ms=30)
    with pytest.raises(B200KVError):
        p.reserve(2, 256)
    time.sleep(0.06)
    put(p, 2)
    p.close()


def test_many_keys_tombstone_rebuild():
    p = mk(8)
    rng = np.random.default_rng(0)
    live = []
    for i in range(2000):
        k = int(rng.integers(1, 2 ** 63))
        put(p, k)
        live.append(k)
        live = live[-8:]
    for k in live:
        assert p.lookup(np.array([k], np.uint64), np.array([256], np.int32))[0] == 1
    assert p.stats()["n_used"] == 8
    assert p.clear() and p.stats()["n_used"] == 0
    p.close()


def test_owner_lookup():
    p = mk()
    put(p, 7, owner=3)
    put(p, 8, owner=5)
    hits, owners = p.lookup_owner(np.array([7, 8, 9], np.uint64))
    assert hits == 2 and list(owners) == [3, 5]
    p.close()


def _child(name, q):
    try:
        c = KVPool(name, 0, SLOT, _lib.POOL_ATTACH)
        hit = c.lookup(np.array([42, 43], np.uint64), np.array([256, 7], np.int32))
        slot, n, _ = c.acquire(43)
        data = bytes(c.slot_view(slot)[:8])
        c.release(43)
        put(c, 44, 256, fill=9)
        c.close()
        q.put((hit, n, data))
    except Exception as e:  # pragma: no cover
        q.put(repr(e))


def test_shared_segment_across_processes(shm_name):
    """Scheduler-role and worker-role connectors (and other replicas) share one index."""
    p = KVPool(shm_name, 4 * SLOT, SLOT, _lib.POOL_CREATE)
    put(p, 42, 256, fill=1)
    put(p, 43, 7, fill=5)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    proc = ctx.Process(target=_child, args=(shm_name, q))
    proc.start()
    res = q.get(timeout=60)
    proc.join(60)
    assert res == ((2, 263), 7, b"\x05" * 8), res
    slot, n, _ = p.acquire(44)                     # written by the other process
    assert n == 256 and int(p.slot_view(slot)[0]) == 9
    p.release(44)
    with pytest.raises(B200KVError):
        KVPool(shm_name, 4 * SLOT, SLOT, _lib.POOL_CREATE)      # exclusive create
    with pytest.raises(B200KVError):
        KVPool(shm_name, 0, SLOT * 2, _lib.POOL_ATTACH)          # geometry mismatch
    p.close()


def test_bad_arguments():
    with pytest.raises(B200KVError):
        KVPool(None, 10, 4096, _lib.POOL_CREATE)                 # pool smaller than a slot
    with pytest.raises(B200KVError):
        KVPool("no-leading-slash", 8192, 4096, _lib.POOL_CREATE)
    with pytest.raises(B200KVError):
        KVPool("/b200kv-does-not-exist", 0, 0, _lib.POOL_ATTACH)


# ---- failure detection: processes that die mid-protocol ------------------------------------------
def test_stale_writer_and_stale_pin_are_reclaimed(monkeypatch):
    """A writer that died between reserve and commit, and a reader that died between acquire and
    release, must not wedge the pool: both are recognised by age (B200KV_POOL_STALE_MS)."""
    monkeypatch.setenv("B200KV_POOL_STALE_MS", "400")
    p = mk(2)
    p.reserve(1, 256)                      # never committed
    put(p, 2)
    p.acquire(2)                           # never released
    for key in (1, 3):                     # young: the key is somebody's / nothing can be evicted
        with pytest.raises(B200KVError) as ei:
            p.reserve(key, 256)
        assert ei.value.code == (_lib.EEXIST if key == 1 else _lib.ENOSPC)
    time.sleep(0.6)
    put(p, 1, fill=7)                      # the dead writer's key is taken over
    put(p, 3)                              # the dead reader's pin no longer protects chunk 2
    keys = np.array([2], np.uint64)
    assert p.lookup(keys, np.array([256], np.int32))[0] == 0
    st = p.stats()
    assert st["n_reclaimed_chunks"] == 2 and st["n_used"] == 2 and p.check()
    slot, _, _ = p.acquire(1)
    assert int(p.slot_view(slot)[0]) == 7
    p.release(1)
    p.close()


def _hammer(name, ready):
    c = KVPool(name, 0, SLOT, _lib.POOL_ATTACH)
    k = 10_000
    while True:                            # killed from outside, possibly inside a critical section
        k += 1
        if k == 10_200:
            ready.set()
        try:
            c.reserve(k, 256)
            c.commit(k)
            c.acquire(k)
            c.release(k)
        except B200KVError:
            pass


def test_pool_survives_a_process_killed_at_any_point(shm_name):
    """SIGKILL a process that is reserving/committing/evicting in a tight loop, many times: whatever
    it held must not break the others (b200kv_pool_check verifies lists and table)."""
    import os
    import signal
    p = KVPool(shm_name, 8 * SLOT, SLOT, _lib.POOL_CREATE)
    ctx = mp.get_context("spawn")
    for round_ in range(4):
        ready = ctx.Event()
        proc = ctx.Process(target=_hammer, args=(shm_name, ready))
        proc.start()
        assert ready.wait(60)
        time.sleep(0.003 * (round_ + 1))
        os.kill(proc.pid, signal.SIGKILL)  # exactly the process started above
        proc.join(30)
        assert p.check(), f"round {round_}: inconsistent after kill"
        put(p, 500 + round_, fill=round_)  # the pool still takes writes and serves them
        slot, n, _ = p.acquire(500 + round_)
        assert int(p.slot_view(slot)[0]) == round_
        p.release(500 + round_)
    assert p.stats()["n_used"] <= 8 and p.check()
    p.close()


def _die_in_reserve(name):
    os_env_key = 777
    c = KVPool(name, 0, SLOT, _lib.POOL_ATTACH)
    put(c, 5, fill=5)
    c.reserve(os_env_key, 256)             # B200KV_POOL_TEST_DIE_KEY=777: _exit(9) with the lock held


def test_lock_owner_death_rebuilds_the_index(shm_name, monkeypatch):
    """Deterministic version: the child exits INSIDE reserve's critical section (fault injection),
    slot written and hashed but the counters not yet updated.  The next caller gets EOWNERDEAD,
    rebuilds free list / LRU / table from the slot array, and carries on."""
    p = KVPool(shm_name, 4 * SLOT, SLOT, _lib.POOL_CREATE)
    put(p, 1, fill=1)
    put(p, 2, fill=2)
    monkeypatch.setenv("B200KV_POOL_TEST_DIE_KEY", "777")
    ctx = mp.get_context("spawn")
    proc = ctx.Process(target=_die_in_reserve, args=(shm_name,))
    proc.start()
    proc.join(60)
    assert proc.exitcode == 9
    monkeypatch.delenv("B200KV_POOL_TEST_DIE_KEY")
    keys = np.array([1, 2, 5], np.uint64)
    ct = np.array([256] * 3, np.int32)
    assert p.lookup(keys, ct)[0] == 3      # first call after the death: recovers, then answers
    st = p.stats()
    assert st["n_recoveries"] == 1 and st["n_used"] == 4 and p.check()   # 1, 2, 5 READY + 777 WRITING
    with pytest.raises(B200KVError) as ei:
        p.reserve(777, 256)                # the dead writer's reservation is still young
    assert ei.value.code == _lib.EEXIST
    for k in (1, 2, 5):
        slot, _, _ = p.acquire(k)
        assert int(p.slot_view(slot)[0]) == k
        p.release(k)
    p.close()


def _store_and_exit(name):
    c = KVPool(name, 4 * SLOT, SLOT, _lib.POOL_CREATE_OR_ATTACH)
    put(c, 900, 256, fill=9)
    put(c, 901, 77, fill=8)
    # no close(), no unlink: the engine process just goes away


def test_named_pool_survives_engine_restart(shm_name):
    """Warm restart (SURVEY.md §5 checkpoint/resume, for this path): with B200KV_POOL_NAME the KV a
    replica offloaded outlives the replica — a restarted engine attaches to the same segment and finds
    its chunks, sized and tagged as they were."""
    ctx = mp.get_context("spawn")
    p1 = ctx.Process(target=_store_and_exit, args=(shm_name,))
    p1.start()
    p1.join(60)
    assert p1.exitcode == 0
    again = KVPool(shm_name, 4 * SLOT, SLOT, _lib.POOL_CREATE_OR_ATTACH)     # what the restarted connector does
    assert again.lookup(np.array([900, 901], np.uint64), np.array([256, 77], np.int32)) == (2, 333)
    slot, n, _ = again.acquire(901)
    assert n == 77 and int(again.slot_view(slot)[0]) == 8
    again.release(901)
    with pytest.raises(B200KVError):                                          # another model geometry: refused
        KVPool(shm_name, 4 * SLOT, 2 * SLOT, _lib.POOL_CREATE_OR_ATTACH)
    again.close()


def test_sweep_removes_only_segments_nobody_holds():
    """A per-engine segment whose processes died (SIGKILL: no atexit) is removed by the next engine's start-up
    sweep; a segment somebody still has open — in this or another process — is left alone."""
    import multiprocessing as mp
    import os
    import signal
    import time
    from b200kv import KVPool, _lib
    tag = f"b200kv-swp{os.getpid()}-"
    live, dead = "/" + tag + "live", "/" + tag + "dead"

    def holder(name, ready):
        import sys
        sys.path[:0] = [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "production-stack_b200")]
        from b200kv import KVPool as KP, _lib as L
        keep = KP(name, 4 * 4096, 4096, L.POOL_CREATE)   # noqa: F841  (the reference keeps the segment open, i.e. locked)
        ready.set()
        time.sleep(60)

    ctx = mp.get_context("fork")
    ready = ctx.Event()
    p = ctx.Process(target=holder, args=(dead, ready))
    p.start()
    assert ready.wait(30)
    mine = KVPool(live, 4 * 4096, 4096, _lib.POOL_CREATE)
    try:
        assert KVPool.sweep("/" + tag, 0) == 0                       # both held
        os.kill(p.pid, signal.SIGKILL)
        p.join(10)
        assert os.path.exists("/dev/shm" + dead)                     # the killed process could not clean up
        assert KVPool.sweep("/" + tag, 3600) == 0                    # too young for a cautious sweep
        assert KVPool.sweep("/" + tag, 0) == 1
        assert not os.path.exists("/dev/shm" + dead) and os.path.exists("/dev/shm" + live)
        assert mine.check()
    finally:
        mine.close()
        KVPool.unlink(live)
        KVPool.unlink(dead)

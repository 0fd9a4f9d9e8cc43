"""Remote chunk tier (CPU): the `lm://` cache server and its client (SURVEY.md §8f rank 3;
helm/templates/deployment-cache-server.yaml:62-65, deployment-vllm-multi.yaml:338-345).
Chunks travel socket <-> pool slot inside libb200kv.so; bytes, format tag and token count must
survive the round trip exactly."""
import os
import subprocess
import sys
import threading
import time

import numpy as np
import pytest

from b200kv import KVPool, _lib
from b200kv.engine import chunk_keys
from b200kv.remote import RemoteClient, RemoteServer, RemoteTier, parse_remote_url

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SLOT = 64 * 1024


def mk_pool(n_slots=8):
    return KVPool(None, n_slots * SLOT, SLOT, _lib.POOL_CREATE)


def put_local(pool, key, n_tok, seed, fmt=0, owner=0):
    slot = pool.reserve(int(key), n_tok, fmt, owner)
    data = np.random.default_rng(seed).integers(0, 256, SLOT, dtype=np.uint8)
    pool.slot_view(slot)[:] = data
    pool.commit(int(key))
    return data


def read_local(pool, key):
    slot, n_tok, fmt = pool.acquire(int(key))
    out = pool.slot_view(slot).copy()
    pool.release(int(key))
    return out, n_tok, fmt


@pytest.fixture
def server():
    s = RemoteServer("127.0.0.1", 0, 6 * SLOT)
    yield s
    s.stop()


def test_parse_remote_url():
    assert parse_remote_url("lm://vllm-cache-server-service:80") == ("vllm-cache-server-service", 80)
    assert parse_remote_url("lm://10.0.0.3:8080/") == ("10.0.0.3", 8080)
    assert parse_remote_url("cache:81") == ("cache", 81)
    assert parse_remote_url("") is None and parse_remote_url(None) is None
    for bad in ("redis://h:1", "lm://nohost", "lm://h:port"):
        with pytest.raises(ValueError):
            parse_remote_url(bad)


def test_put_get_roundtrip_is_bit_exact(server):
    a, b = mk_pool(), mk_pool()
    c = RemoteClient("127.0.0.1", server.port)
    assert c.ping()
    want = {}
    for i, (key, n_tok, fmt) in enumerate([(101, 256, 0), (102, 256, 1 | (1 << 8)), (103, 40, 0)]):
        want[key] = (put_local(a, key, n_tok, i, fmt, owner=7), n_tok, fmt)
        assert c.put(a, key, 7) == 0
    assert c.put(a, 101, 7) == _lib.EEXIST                  # second upload: nothing is sent
    assert c.put(a, 999, 7) == _lib.ENOENT                  # not in the local pool
    assert c.exists(np.array([101, 102, 103, 104], np.uint64)) == 3
    assert c.exists(np.array([104, 101], np.uint64)) == 0   # prefix semantics
    for key, (data, n_tok, fmt) in want.items():
        assert c.get(b, key, 9) == 0
        got, n2, f2 = read_local(b, key)
        assert np.array_equal(got, data) and n2 == n_tok and f2 == fmt
    assert c.get(b, 101, 9) == 0                            # already local: fine
    assert c.get(b, 555, 9) == _lib.ENOENT
    up, down = c.traffic()
    assert up == 3 * SLOT and down == 3 * SLOT
    st = server.stats()
    assert st["n_put"] == 3 and st["n_get"] == 4 and st["n_get_miss"] == 1
    assert c.stats()["n_used"] == 3
    c.close()


def test_server_evicts_lru_and_keeps_one_pool_per_geometry(monkeypatch):
    # the server's budget (12 slots' worth) is split over the 2 geometries it is told to expect
    monkeypatch.setenv("B200KV_SERVER_GEOMETRIES", "2")
    server = RemoteServer("127.0.0.1", 0, 12 * SLOT)
    try:
        _geometry_checks(server)
    finally:
        server.stop()


def _geometry_checks(server):
    a = mk_pool(16)
    c = RemoteClient("127.0.0.1", server.port)
    for k in range(1, 9):                                   # server holds 6 slots
        put_local(a, k, 256, k)
        assert c.put(a, k, 0) == 0
    assert c.stats()["n_used"] == 6 and c.stats()["n_evicted_chunks"] == 2
    assert c.exists(np.array([1], np.uint64)) == 0 and c.exists(np.array([8, 7, 3], np.uint64)) == 3
    # a second model on the same server (the chart deploys ONE cache server for all modelSpecs): its chunks
    # have another size and live in a pool of their own, with their own LRU
    other = KVPool(None, 4 * 2 * SLOT, 2 * SLOT, _lib.POOL_CREATE)
    slot = other.reserve(77, 256, 0, 0)
    other.slot_view(slot)[:] = 5
    other.commit(77)
    assert c.put(other, 77, 0) == 0
    assert c.exists(np.array([77], np.uint64)) == 1 and c.exists(np.array([8, 7, 3], np.uint64)) == 3
    assert c.stats()["n_used"] == 7                          # 6 chunks of the first model + 1 of the second
    dst = KVPool(None, 2 * 2 * SLOT, 2 * SLOT, _lib.POOL_CREATE)
    assert c.get(dst, 77, 0) == 0
    s2, n2, _ = dst.acquire(77)
    assert n2 == 256 and int(dst.slot_view(s2)[-1]) == 5
    dst.release(77)
    # asking for a chunk of the other geometry is refused (payload drained, connection still usable)
    assert c.get(other, 8, 0) == _lib.EINVAL and c.get(a, 77, 0) == _lib.EINVAL
    assert c.ping()
    # a third slot size is not given a pool of its own (the budget is the server's, not each geometry's)
    third = KVPool(None, 2 * 4 * SLOT, 4 * SLOT, _lib.POOL_CREATE)
    slot = third.reserve(99, 256, 0, 0)
    third.slot_view(slot)[:] = 1
    third.commit(99)
    assert c.put(third, 99, 0) != 0 and c.exists(np.array([99], np.uint64)) == 0
    assert c.ping() and c.exists(np.array([77], np.uint64)) == 1
    c.close()


def test_concurrent_clients(server):
    pools = [mk_pool(4) for _ in range(4)]
    errs = []

    def work(i):
        try:
            c = RemoteClient("127.0.0.1", server.port)
            data = put_local(pools[i], 1000 + i, 256, 50 + i)
            assert c.put(pools[i], 1000 + i, i) == 0
            dst = mk_pool(2)
            assert c.get(dst, 1000 + i, 0) == 0
            assert np.array_equal(read_local(dst, 1000 + i)[0], data)
            c.close()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs


def test_tier_push_then_prefetch_into_another_pool(server):
    C_, seed = 16, 1234
    toks = np.arange(5 * C_ + 3, dtype=np.int32)
    keys = chunk_keys(toks, C_, seed, True)
    prod_pool, cons_pool = mk_pool(8), mk_pool(8)
    data = [put_local(prod_pool, k, min(C_, len(toks) - i * C_), i) for i, k in enumerate(keys)]
    prod = RemoteTier(prod_pool, "127.0.0.1", server.port, C_, seed, owner=1)
    prod.push(keys)
    assert prod.flush(10)
    assert prod.stats()["pushed_chunks"] == len(keys)

    cons = RemoteTier(cons_pool, "127.0.0.1", server.port, C_, seed, owner=2)
    put_local(cons_pool, keys[0], C_, 0)                    # first chunk already local: only the rest is fetched
    assert cons.prefetch_state("r1", toks) == RemoteTier.PENDING
    t0 = time.time()
    while cons.prefetch_state("r1", toks) == RemoteTier.PENDING:
        assert time.time() - t0 < 10
        time.sleep(0.005)
    assert cons.stats()["fetched_chunks"] == len(keys) - 1
    n_local, owners = cons_pool.lookup_owner(keys)
    assert n_local == len(keys) and list(owners[1:]) == [2] * (len(keys) - 1)
    for i, k in enumerate(keys):
        assert np.array_equal(read_local(cons_pool, k)[0], data[i])
    assert cons.prefetch_state("r2", toks) == RemoteTier.DONE          # everything local: no job
    other = np.arange(1000, 1000 + 2 * C_, dtype=np.int32)
    assert cons.prefetch_state("r3", other) == RemoteTier.PENDING      # one EXISTS round trip, nothing to fetch
    t0 = time.time()
    while cons.prefetch_state("r3", other) == RemoteTier.PENDING:
        assert time.time() - t0 < 10
        time.sleep(0.002)
    assert cons.stats()["fetched_chunks"] == len(keys) - 1
    prod.close()
    cons.close()


def test_tier_degrades_when_server_is_gone():
    s = RemoteServer("127.0.0.1", 0, 4 * SLOT)
    port = s.port
    s.stop()
    pool = mk_pool(4)
    tier = RemoteTier(pool, "127.0.0.1", port, 16, 1, retry_s=60.0, timeout_ms=500)
    toks = np.arange(32, dtype=np.int32)
    assert tier.prefetch_state("a", toks) == RemoteTier.PENDING
    t0 = time.time()
    while tier.prefetch_state("a", toks) == RemoteTier.PENDING:
        assert time.time() - t0 < 10
        time.sleep(0.005)
    assert tier.stats()["errors"] == 1
    assert tier.prefetch_state("b", toks) == RemoteTier.DONE            # backing off: requests are not stalled
    put_local(pool, 5, 16, 0)
    tier.push([5])
    assert tier.flush(5) and tier.stats()["push_skipped"] == 1
    tier.close()


def test_pending_push_waits_for_commit(server):
    pool = mk_pool(4)
    tier = RemoteTier(pool, "127.0.0.1", server.port, 16, 1)
    slot = pool.reserve(42, 16, 0, 0)            # store in flight: reserved, not yet committed
    tier.push([42])
    time.sleep(0.05)
    assert tier.stats()["pushed_chunks"] == 0
    pool.slot_view(slot)[:] = 7
    pool.commit(42)
    assert tier.flush(5) and tier.stats()["pushed_chunks"] == 1
    dst = mk_pool(2)
    c = RemoteClient("127.0.0.1", server.port)
    assert c.get(dst, 42, 0) == 0 and int(read_local(dst, 42)[0][0]) == 7
    c.close()
    tier.close()


def test_lmcache_server_entry_point_and_probe(tmp_path):
    """`lmcache_server <host> <port>` exactly as the chart execs it, from the compat tree."""
    env = dict(os.environ, B200KV_SERVER_GB="0.01")
    exe = os.path.join(ROOT, "production-stack_b200", "compat", "bin", "lmcache_server")
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    proc = subprocess.Popen([exe, "127.0.0.1", str(port)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                            text=True, start_new_session=True)
    try:
        line = proc.stdout.readline()
        assert "listening" in line, line
        probe = [sys.executable, "-m", "b200kv.server", "--probe", "127.0.0.1", str(port)]
        penv = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "production-stack_b200"))
        assert subprocess.run(probe, env=penv, timeout=60).returncode == 0
        pool = mk_pool(2)
        put_local(pool, 9, 256, 1)
        c = RemoteClient("127.0.0.1", port)
        assert c.put(pool, 9, 0) == 0 and c.exists(np.array([9], np.uint64)) == 1
        c.close()
    finally:
        proc.terminate()               # exactly the process started above
        assert proc.wait(timeout=20) == 0
    assert subprocess.run([sys.executable, "-m", "b200kv.server", "--probe", "127.0.0.1", str(port)],
                          env=dict(os.environ, PYTHONPATH=os.path.join(ROOT, "production-stack_b200")),
                          timeout=60, stderr=subprocess.DEVNULL).returncode == 1


def test_server_aborts_half_received_put_and_rejects_garbage(server):
    """A client that dies in the middle of an upload must not leave a half-written chunk behind, and a
    peer speaking another protocol is dropped without harming the server."""
    import socket
    import struct
    s = socket.create_connection(("127.0.0.1", server.port))
    # header: magic, version, op=PUT(2), key, fmt, n_tokens, owner, status, length, slot_bytes  (48 bytes, LE)
    hdr = struct.pack("<IHHQIiIiQQ", 0x564B3242, 1, 2, 4242, 0, 256, 0, 0, SLOT, SLOT)
    assert len(hdr) == 48
    s.sendall(hdr)
    rep = s.recv(48, socket.MSG_WAITALL)
    assert struct.unpack("<IHHQIiIiQQ", rep)[7] == 0          # "send it"
    s.sendall(b"x" * (SLOT // 2))
    s.close()                                                 # dies half way
    g = socket.create_connection(("127.0.0.1", server.port))
    g.sendall(b"GET / HTTP/1.1\r\nHost: x\r\n\r\n" + b"\0" * 32)
    try:
        assert g.recv(64) == b""                              # connection dropped (FIN) ...
    except ConnectionResetError:
        pass                                                  # ... or reset, because unread bytes were pending
    g.close()
    c = RemoteClient("127.0.0.1", server.port)
    t0 = time.time()
    while c.stats()["n_used"] != 0:                           # the reservation is aborted once the socket closes
        assert time.time() - t0 < 5
        time.sleep(0.01)
    assert c.exists(np.array([4242], np.uint64)) == 0
    a = mk_pool(2)
    put_local(a, 4242, 256, 9)
    assert c.put(a, 4242, 0) == 0                             # the key can be stored afterwards
    c.close()

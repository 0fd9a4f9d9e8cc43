"""The lmcache compat shim's controller against the calls the reference router makes
(src/vllm_router/routers/routing_logic.py:276-316, 378-428): ZMQ registration from a worker-side
client, LookupMsg -> layout_info[instance][1], QueryInstMsg -> instance_id.  CPU only."""
import asyncio
import os
import socket
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "production-stack_b200", "compat"))

from lmcache.v1.cache_controller import controller_manager  # noqa: E402
from lmcache.v1.cache_controller.message import LookupMsg, QueryInstMsg  # noqa: E402

from b200kv import KVPool, _lib, chunk_keys  # noqa: E402
from b200kv.controller_client import ControllerClient  # noqa: E402

SLOT = 4096


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def fill(pool, toks, seed, owner=0):
    for i, k in enumerate(chunk_keys(toks, 256, seed)):
        pool.reserve(int(k), min(256, len(toks) - i * 256), 0, owner)
        pool.commit(int(k))


def test_controller_lookup_and_query_inst(shm_name):
    port = free_port()
    mgr = controller_manager.LMCacheControllerManager({"pull": f"0.0.0.0:{port}", "reply": None},
                                                      health_check_interval=1, lmcache_worker_timeout=30)
    loop = asyncio.new_event_loop()
    th = threading.Thread(target=loop.run_forever, daemon=True)
    th.start()
    task = asyncio.run_coroutine_threadsafe(mgr.start_all(), loop)   # exactly what KvawareRouter does
    name_a, name_b = shm_name + "a", shm_name + "b"
    pa = KVPool(name_a, 16 * SLOT, SLOT, _lib.POOL_CREATE)
    pb = KVPool(name_b, 16 * SLOT, SLOT, _lib.POOL_CREATE)
    try:
        toks = np.arange(3000, dtype=np.int32)
        fill(pa, toks[:1024], seed=11)          # instance A holds 4 chunks of the prompt
        fill(pb, toks[:2300], seed=22)          # instance B holds 8 full chunks + a partial tail
        ca = ControllerClient(f"127.0.0.1:{port}", "pod-a", name_a, 11, 256, heartbeat_s=0.2, ip="10.0.0.1")
        cb = ControllerClient(f"127.0.0.1:{port}", "pod-b", name_b, 22, 256, heartbeat_s=0.2, ip="10.0.0.2")
        t0 = time.time()
        while len(mgr.workers) < 2 and time.time() - t0 < 10:
            time.sleep(0.05)
        assert set(mgr.workers) == {"pod-a", "pod-b"}

        def ask(msg):
            return asyncio.run_coroutine_threadsafe(mgr.handle_orchestration_message(msg), loop).result(10)

        ret = ask(LookupMsg(tokens=list(toks[:2300]), event_id="e1"))
        first = list(ret.layout_info.keys())[0]                       # the router takes the first key
        assert first == "pod-b" and ret.layout_info[first][1] == 2300
        ret = ask(LookupMsg(tokens=list(toks[:1500]), event_id="e2"))  # B's partial tail does not apply here
        assert list(ret.layout_info.items())[0] == ("pod-b", ("LocalCPUBackend", 1280))
        ret = ask(LookupMsg(tokens=[5, 6, 7], event_id="e3"))
        assert ret.layout_info == {}
        assert ask(QueryInstMsg(ip="10.0.0.1", event_id="q")).instance_id == "pod-a"
        assert ask(QueryInstMsg(ip="10.9.9.9", event_id="q")).instance_id is None
        ca.close()
        cb.close()
    finally:
        mgr.stop()
        try:
            task.result(5)
        except Exception:
            pass
        loop.call_soon_threadsafe(loop.stop)
        pa.close()
        pb.close()
        KVPool.unlink(name_a)
        KVPool.unlink(name_b)


def test_shared_pool_attribution_by_owner_tag(shm_name):
    """BASELINE config 3: one pool shared by all replicas; a lookup credits an instance only with
    the prefix it stored itself."""
    mgr = controller_manager.LMCacheControllerManager({"pull": "0.0.0.0:1", "reply": None})
    from lmcache.v1.cache_controller.message import RegisterMsg
    p = KVPool(shm_name, 32 * SLOT, SLOT, _lib.POOL_CREATE)
    toks = np.arange(1024, dtype=np.int32)
    keys = chunk_keys(toks, 256, 5)
    for i, k in enumerate(keys):                       # chunks 0,1 by replica 7; chunks 2,3 by replica 9
        p.reserve(int(k), 256, 0, 7 if i < 2 else 9)
        p.commit(int(k))
    mgr.register(RegisterMsg("r7", "10.0.0.7", shm_name, 5, 256, owner_tag=7))
    mgr.register(RegisterMsg("r9", "10.0.0.9", shm_name, 5, 256, owner_tag=9))
    ret = asyncio.run(mgr.handle_orchestration_message(LookupMsg(tokens=list(toks))))
    assert ret.layout_info == {"r7": ("LocalCPUBackend", 512)}
    p.close()

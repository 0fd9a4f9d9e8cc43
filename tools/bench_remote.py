"""Loopback throughput of the cache-server tier (CPU only, no GPU needed): Llama-3-8B chunks
(32 MiB RAW / 16 MiB + 2 KiB FP8) pushed from one pool to the server and fetched into another,
with 1..N connections.  Every pool is cycled once before the timed pass (production pools are pinned,
i.e. resident; a first pass over fresh anonymous memory measures page faults).  Prints one JSON line.

    python tools/bench_remote.py [--chunks 24] [--conns 1,2,4] [--fp8]
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "production-stack_b200"))
from b200kv import KVPool, _lib  # noqa: E402
from b200kv.remote import RemoteClient, RemoteServer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=16)
    ap.add_argument("--conns", default="1,2,4")
    ap.add_argument("--fp8", action="store_true")
    args = ap.parse_args()
    slot = (16 << 20) + 2048 if args.fp8 else 32 << 20
    n = args.chunks
    src = KVPool(None, 2 * n * slot, slot, _lib.POOL_CREATE)
    rng = np.random.default_rng(0)
    for base in (1000, 2000):                 # 1000.. = warm-up set, 2000.. = timed set
        for k in range(n):
            s = src.reserve(base + k, 256, 0, 0)
            src.slot_view(s)[:: 4096] = rng.integers(0, 256, (slot + 4095) // 4096, dtype=np.uint8)   # touch every page
            src.commit(base + k)
    out = {"chunk_bytes": slot, "chunks": n, "host_cores": len(os.sched_getaffinity(0)), "transport": "TCP loopback",
           "runs": []}
    for conns in [int(x) for x in args.conns.split(",")]:
        srv = RemoteServer("127.0.0.1", 0, n * slot)          # exactly n slots: the timed set reuses touched slots
        dst = KVPool(None, n * slot, slot, _lib.POOL_CREATE)
        clients = [RemoteClient("127.0.0.1", srv.port) for _ in range(conns)]

        def run(fn):
            ts = [threading.Thread(target=fn, args=(i,)) for i in range(conns)]
            t0 = time.perf_counter()
            [t.start() for t in ts]
            [t.join() for t in ts]
            return time.perf_counter() - t0

        def mk(op, base):
            def fn(i):
                for k in range(i, n, conns):
                    rc = clients[i].put(src, base + k, 0) if op == "put" else clients[i].get(dst, base + k, 0)
                    assert rc == 0, rc
            return fn

        run(mk("put", 1000))
        run(mk("get", 1000))
        dst.clear()
        t_put = run(mk("put", 2000))
        t_get = run(mk("get", 2000))
        s0, _, _ = dst.acquire(2000)
        s1, _, _ = src.acquire(2000)
        assert np.array_equal(dst.slot_view(s0), src.slot_view(s1))
        src.release(2000)
        out["runs"].append({"connections": conns, "put_GBps": round(n * slot / t_put / 1e9, 2),
                            "get_GBps": round(n * slot / t_get / 1e9, 2)})
        for c in clients:
            c.close()
        srv.stop()
        dst.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

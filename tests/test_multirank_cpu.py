"""N>1 host logic on CPU: 2 processes, gloo backend.  Covers the shared pinned-host pool
(BASELINE.json config 3), the IPC-descriptor handshake and the session sharding bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

SLOT = 8192


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shm, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "production-stack_b200")]
    from b200kv import KVPool, _lib, chunk_keys
    from b200kv.peers import all_gather_bytes, shard_sessions
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank opens the same segment; exactly one creates it
        pool = KVPool(shm, 8 * SLOT, SLOT, _lib.POOL_CREATE_OR_ATTACH)
        toks = (np.arange(700) * (rank + 3)).astype(np.int32)
        keys = chunk_keys(toks, 256, 5)
        for i, k in enumerate(keys):
            slot = pool.reserve(int(k), min(256, 700 - i * 256), 0, owner=rank)
            pool.slot_view(slot)[:] = rank + 1
            pool.commit(int(k))
        dist.barrier()
        # the other rank's chunks are visible (shared index) with the right owner and payload
        other = (rank + 1) % world
        otoks = (np.arange(700) * (other + 3)).astype(np.int32)
        hit = pool.lookup_tokens(otoks, 256, 5)
        okeys = chunk_keys(otoks, 256, 5)
        hc, owners = pool.lookup_owner(okeys)
        slot, n, _ = pool.acquire(int(okeys[-1]))
        payload_ok = bool((pool.slot_view(slot) == other + 1).all()) and n == 700 - 512
        pool.release(int(okeys[-1]))
        # IPC descriptor handshake (bytes are opaque here; on GPU they are b200kv_ipc_desc[2L])
        descs = all_gather_bytes(bytes([rank]) * 80 * 4)
        dist.barrier()
        st = pool.stats()
        pool.close()
        q.put((rank, hit, hc, [int(o) for o in owners], payload_ok, [d[0] for d in descs], len(descs[0]),
               list(shard_sessions(16, rank, world)), st["n_used"]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_share_pool_and_exchange_descriptors(shm_name):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shm_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, hit, hc, owners, payload_ok, firsts, dlen, sessions, used in res:
        other = (rank + 1) % world
        assert hit == 700 and hc == 3 and owners == [other] * 3 and payload_ok
        assert firsts == [0, 1] and dlen == 320
        assert sessions == list(range(rank * 8, rank * 8 + 8))
        assert used == 6


def test_shard_sessions_covers_everything():
    from b200kv.peers import shard_sessions
    for n, w in [(16, 1), (16, 2), (16, 8), (10, 4), (3, 8)]:
        got = [s for r in range(w) for s in shard_sessions(n, r, w)]
        assert got == list(range(n))

"""Cache-server tier with the real CUDA engine (run on a B200 with -m gpu): engine A stores a prompt's
KV, its chunks are uploaded from A's pinned pool to the server, engine B — another pool, other pages —
prefetches them into ITS pinned pool and loads them; B's pages must equal what the oracle says A's
contained, bit for bit (RAW) / code for code (FP8)."""
import time

import numpy as np
import pytest

from oracle import kv_oracle as ko

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from b200kv import FMT_FP8, FMT_RAW, KVEngine, KVGeometry, KVPool  # noqa: E402
from b200kv.remote import RemoteServer, RemoteTier  # noqa: E402
from tests.test_gpu_kernels import SMALL, bits_of, mk_host_layers, need_gpu, to_dev  # noqa: E402


@pytest.mark.parametrize("fmt", [FMT_RAW, FMT_FP8])
def test_store_upload_prefetch_retrieve_across_two_engines(fmt):
    need_gpu()
    p = SMALL
    rng = np.random.default_rng(77)
    host = mk_host_layers(rng, p["L"], p["NB"], p["bs"], p["H"], p["D"])
    dev_a = to_dev(host)
    dev_b = [torch.zeros_like(t) for t in dev_a]
    geom = KVGeometry(p["L"], p["H"], p["D"], p["NB"], p["bs"], p["C"], 2, 0, fmt)
    cb = geom.chunk_bytes
    pool_a, pool_b = KVPool(None, 6 * cb, cb, 1), KVPool(None, 6 * cb, cb, 1)
    eng_a = KVEngine(geom, pool_a, 0, staging_bytes=4 * cb)
    eng_b = KVEngine(geom, pool_b, 0, staging_bytes=4 * cb)
    eng_a.register_kv_caches(dev_a)
    eng_b.register_kv_caches(dev_b)
    srv = RemoteServer("127.0.0.1", 0, 8 * cb)
    tier_a = tier_b = None
    try:
        n = 2 * p["C"] + 40
        toks = rng.integers(0, 128256, n).astype(np.int32)
        nblk = (n + p["bs"] - 1) // p["bs"]
        sm_a = ko.slot_mapping_from_blocks(rng.permutation(p["NB"])[:nblk], p["bs"], n)
        sm_b = ko.slot_mapping_from_blocks(rng.permutation(p["NB"])[:nblk], p["bs"], n)
        eng_a.wait(eng_a.store(toks, None, sm_a))
        tier_a = RemoteTier(pool_a, "127.0.0.1", srv.port, p["C"], eng_a.key_seed, owner=1)
        tier_a.push(eng_a._keys(toks))
        assert tier_a.flush(30) and tier_a.stats()["pushed_chunks"] == 3

        assert eng_b.lookup(toks) == 0
        tier_b = RemoteTier(pool_b, "127.0.0.1", srv.port, p["C"], eng_b.key_seed, owner=2)
        t0 = time.time()
        while tier_b.prefetch_state("req", toks) == RemoteTier.PENDING:
            assert time.time() - t0 < 30
            time.sleep(0.002)
        assert tier_b.stats()["fetched_chunks"] == 3 and eng_b.lookup(toks) == n
        ret = eng_b.retrieve(toks, None, sm_b)
        torch.cuda.synchronize()
        assert ret.all()

        o_dst = [np.zeros_like(x) for x in host]
        oe = ko.OracleEngine(p["C"], "fp8" if fmt == FMT_FP8 else "raw")
        oe.store(toks, np.ones(n, bool), host, sm_a)
        oe.retrieve(toks, np.ones(n, bool), o_dst, sm_b)
        for got, want in zip(dev_b, o_dst):
            assert np.array_equal(bits_of(got), want)
        assert srv.stats()["n_put"] == 3 and srv.stats()["n_get"] == 3
    finally:
        for t in (tier_a, tier_b):
            if t is not None:
                t.close()
        srv.stop()
        eng_a.close()
        eng_b.close()

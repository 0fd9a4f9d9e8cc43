"""Device chunk tier on the GPU (BASELINE.json configs[3]): an owner process keeps stored chunks in HBM
and publishes the buffer over CUDA IPC; a consumer process — another engine, other pages, EMPTY host
pool — finds them through the shm index, pins them, and scatters straight from the owner's HBM (P2P loads
when the owner sits on another GPU).  Pages must equal the owner's, bit for bit (RAW) / code for code (FP8)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L, NB, BS, H, D, C = 4, 128, 16, 8, 128, 256
N_TOK = 3 * C + 40


def _mk(dev, seed, fmt):
    from b200kv import KVEngine, KVGeometry, KVPool
    g = torch.Generator(device=f"cuda:{dev}").manual_seed(seed)
    caches = [torch.randn((2, NB, BS, H, D), generator=g, device=f"cuda:{dev}", dtype=torch.float32).bfloat16()
              for _ in range(L)]
    geom = KVGeometry(L, H, D, NB, BS, C, 2, 0, fmt)
    pool = KVPool(None, 4 * geom.chunk_bytes, geom.chunk_bytes, 1)
    eng = KVEngine(geom, pool, dev, staging_bytes=4 * geom.chunk_bytes)
    eng.register_kv_caches(caches)
    return caches, eng


def _owner(dev, eid, fmt, ready, stop, shm_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "production-stack_b200")]
    import torch
    import b200kv.device_tier as dt
    import b200kv.pd as pd
    from b200kv.adapter import ReqMeta, SaveSpec, WorkerState
    dt.SHM_DIR = pd.SHM_DIR = shm_dir
    torch.cuda.set_device(dev)
    caches, eng = _mk(dev, 42, fmt)
    w = WorkerState(eng, BS, C)
    w.tiers = dt.TierSet(eid, eng.geom.chunk_bytes, fmt, importer=eng.tier_import)
    w.local_tier = dt.LocalTier(eng, eid, 8, dev, fmt, owner=1)
    w.tiers.add_local(w.local_tier)
    toks = np.arange(N_TOK, dtype=np.int32) * 3 % 50000
    blocks = [int(b) for b in np.random.default_rng(1).permutation(NB)[: (N_TOK + BS - 1) // BS]]
    w.save([ReqMeta("r", toks, blocks, is_last_prefill=True, save_spec=SaveSpec(0, True))], stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    w.reap()                      # gathers done -> chunks committed, visible to peers
    ready.set()
    stop.wait(120)
    w.local_tier.close()
    eng.close()


@pytest.mark.parametrize("fmt", [0, 1])
def test_consumer_scatters_from_the_owners_hbm(tmp_path, fmt):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected but no CUDA device")
    sys.path[:0] = [ROOT, os.path.join(ROOT, "production-stack_b200")]
    import torch.multiprocessing as mp

    import b200kv.device_tier as dt
    import b200kv.pd as pd
    from b200kv.adapter import LoadSpec, ReqMeta, WorkerState
    from oracle import kv_oracle as ko
    shm_dir = str(tmp_path)
    dt.SHM_DIR = pd.SHM_DIR = shm_dir
    owner_dev = 1 if torch.cuda.device_count() > 1 else 0
    ctx = mp.get_context("spawn")
    ready, stop = ctx.Event(), ctx.Event()
    eid = f"own{os.getpid()}f{fmt}"
    proc = ctx.Process(target=_owner, args=(owner_dev, eid, fmt, ready, stop, shm_dir))
    proc.start()
    try:
        assert ready.wait(180), "owner did not come up"
        caches, eng = _mk(0, 7, fmt)
        for t in caches:
            t.zero_()
        w = WorkerState(eng, BS, C)
        w.tiers = dt.TierSet("cons", eng.geom.chunk_bytes, fmt, importer=eng.tier_import, refresh_s=0.0)
        toks = np.arange(N_TOK, dtype=np.int32) * 3 % 50000
        blocks = [int(b) for b in np.random.default_rng(2).permutation(NB)[: (N_TOK + BS - 1) // BS]]
        assert eng.lookup(toks) == 0                                        # nothing in the consumer's host pool
        w.start_load([ReqMeta("r", toks, blocks, load_spec=LoadSpec(0, N_TOK, True))], stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert w.stats.num_tier_peer_tokens == N_TOK and w.stats.num_loaded_tokens == N_TOK
        assert eng.stats()["h2d_bytes"] < (1 << 20)                         # tables only: no chunk crossed PCIe
        w.reap()
        assert not w._tier_pins
        # expected: what the owner's pages held (same seed), through the oracle's store/retrieve
        g = torch.Generator(device="cuda:0").manual_seed(42)
        src = [torch.randn((2, NB, BS, H, D), generator=g, device="cuda:0", dtype=torch.float32).bfloat16().cpu()
               .view(torch.int16).numpy().view(np.uint16) for _ in range(L)]
        oe = ko.OracleEngine(C, "fp8" if fmt else "raw")
        sm_o = ko.slot_mapping_from_blocks([int(b) for b in np.random.default_rng(1).permutation(NB)[: (N_TOK + BS - 1) // BS]],
                                           BS, N_TOK)
        sm_c = ko.slot_mapping_from_blocks(blocks, BS, N_TOK)
        dst = [np.zeros_like(x) for x in src]
        oe.store(toks, np.ones(N_TOK, bool), src, sm_o)
        oe.retrieve(toks, np.ones(N_TOK, bool), dst, sm_c)
        for got, want in zip(caches, dst):
            assert np.array_equal(got.cpu().view(torch.int16).numpy().view(np.uint16), want)
        eng.close()
    finally:
        stop.set()
        proc.join(60)

"""Device chunk tier, host logic on CPU (BASELINE.json configs[3]): two replicas with PRIVATE host pools;
what replica A stored is loaded by replica B straight from A's device tier.  The CUDA engine is replaced
by a stand-in whose "HBM" is a dict and whose gather/scatter move bytes with the oracle; indices, pins,
discovery files and the worker/scheduler logic are the real ones."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest

from b200kv import KVPool, _lib
from b200kv.adapter import LoadSpec, ReqMeta, SaveSpec, WorkerState
from b200kv.device_tier import LocalTier, TierSet, combined_prefix_tokens, index_name, tier_path
from b200kv.engine import chunk_keys
from oracle import kv_oracle as ko

BS, C = 16, 64
HBM: dict[int, np.ndarray] = {}      # "device memory" shared by the fake engines (one box)


class Ev:
    def __init__(self, done=True):
        self.done = done

    def query(self):
        return self.done


class FakeEngine:
    next_base = 0x7000_0000_0000

    def __init__(self, layers, key_seed=5):
        self.layers, self.key_seed = layers, key_seed
        self.geom = NS(chunk_bytes=4096, chunk_tokens=C)
        self.oe = ko.OracleEngine(C)
        self.calls = []

    def _keys(self, tokens):
        return chunk_keys(tokens, C, self.key_seed, True)

    # host path
    def store(self, tokens, mask, slot_mapping, offset=0, stream=None):
        self.oe.store(np.asarray(tokens), mask, self.layers, slot_mapping, offset)
        return 1

    def retrieve(self, tokens, mask, slot_mapping, stream=None, return_ticket=False, layers_per_group=0):
        self.calls.append(("retrieve", len(tokens), int((~mask).sum())))
        ret = self.oe.retrieve(np.asarray(tokens), mask, self.layers, slot_mapping)
        return (ret, 3) if return_ticket else ret

    def poll(self, t):
        return True

    # tier
    def tier_create(self, n_slots):
        FakeEngine.next_base += 1 << 32
        return FakeEngine.next_base

    def tier_export(self):
        return b"\0" * 80

    def tier_import(self, desc):
        raise AssertionError("replaced per test")

    def gather_chunks(self, sm, ptrs, stream=None):
        self.calls.append(("gather_chunks", len(sm), len(ptrs)))
        for i, ptr in enumerate(ptrs):
            seg = np.asarray(sm[i * C:(i + 1) * C])
            HBM[int(ptr)] = ko.gather_tokens(self.layers, seg).copy()

    def scatter_chunks(self, sm, ptrs, stream=None):
        self.calls.append(("scatter_chunks", len(sm), len(ptrs)))
        for i, ptr in enumerate(ptrs):
            seg = np.asarray(sm[i * C:(i + 1) * C])
            ko.scatter_tokens(self.layers, HBM[int(ptr)][:, :, :len(seg)], seg)


def mk_layers(seed):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 2 ** 16, (2, 64, BS, 2, 8), dtype=np.uint16) for _ in range(2)]


@pytest.fixture
def box(monkeypatch, tmp_path):
    import b200kv.device_tier as dt
    import b200kv.pd as pd
    monkeypatch.setattr(dt, "SHM_DIR", str(tmp_path))
    monkeypatch.setattr(pd, "SHM_DIR", str(tmp_path))
    HBM.clear()
    made = []
    yield made
    for name in made:
        KVPool.unlink(name)


def replica(box, eid, layers, n_slots=4, with_tier=True):
    eng = FakeEngine(layers)
    w = WorkerState(eng, BS, C)
    bases = {}
    w.tiers = TierSet(eid, 4096, 0, importer=lambda desc: bases["peer"], refresh_s=0.0)
    w.event_factory = lambda s: Ev(True)
    if with_tier:
        w.local_tier = LocalTier(eng, eid, n_slots, 0, 0, owner=1, event_factory=lambda s: Ev(True))
        w.tiers.add_local(w.local_tier)
        box.append(index_name(eid))
    return eng, w, bases


def save_meta(rid, tokens, blocks):
    return ReqMeta(rid, np.asarray(tokens, np.int32), list(blocks), is_last_prefill=True, save_spec=SaveSpec(0, True))


def load_meta(rid, tokens, blocks, n, vllm_cached=0):
    return ReqMeta(rid, np.asarray(tokens, np.int32), list(blocks), load_spec=LoadSpec(vllm_cached, n, True))


def test_peer_replica_loads_from_the_owners_device_tier(box):
    eid = f"{os.getpid()}"
    la, lb = mk_layers(1), mk_layers(2)
    ea, wa, _ = replica(box, "A" + eid, la)
    eb, wb, bases_b = replica(box, "B" + eid, lb, with_tier=True)
    bases_b["peer"] = wa.local_tier.base                      # what cudaIpcOpenMemHandle would return
    n = 3 * C + 10
    toks = list(range(500, 500 + n))
    blocks_a = list(range(10, 10 + 13))
    wa.save([save_meta("r", toks, blocks_a)])
    assert ea.calls[-1] == ("gather_chunks", n, 4)            # four chunks (the last one partial) into the tier
    keys = ea._keys(np.asarray(toks, np.int32))
    ct = np.array([C, C, C, 10], np.int32)
    assert not wa.tiers.presence(keys, ct).any()              # not visible before the gather has completed
    wa.reap()
    assert wa.tiers.presence(keys, ct).all() and wa.local_tier.stored_chunks == 4

    # replica B: nothing in its host pool, the scheduler's combined lookup still finds the whole prompt
    sched_tiers = TierSet("B" + eid, 4096, None, refresh_s=0.0)
    sched_tiers.refresh()
    host_b = KVPool(None, 8 * 16, 16, _lib.POOL_CREATE)
    assert combined_prefix_tokens(host_b, sched_tiers, keys, ct, 0) == n
    assert combined_prefix_tokens(host_b, None, keys, ct, 0) == 0

    blocks_b = list(range(30, 30 + 13))
    wb.start_load([load_meta("r", toks, blocks_b, n)])
    assert eb.calls == [("scatter_chunks", n, 4)]             # no PCIe, no host pool: four chunk pointers into A's HBM
    sm_a = ko.slot_mapping_from_blocks(blocks_a, BS, n)
    sm_b = ko.slot_mapping_from_blocks(blocks_b, BS, n)
    for x, y in zip(la, lb):
        assert np.array_equal(x.reshape(2, -1, 2, 8)[:, sm_a], y.reshape(2, -1, 2, 8)[:, sm_b])
    assert wb.stats.num_tier_peer_tokens == n and wb.stats.num_tier_local_tokens == 0 and wb.stats.num_loaded_tokens == n
    # while B reads, A cannot recycle those slots: a new 4-chunk store finds no room ...
    other = list(range(9000, 9000 + 4 * C))
    wa.save([save_meta("q", other, list(range(40, 56)))])
    wa.reap()
    assert wa.tiers.presence(keys, ct).all()
    # ... until B's scatter has finished and its pins are dropped
    wb.reap()
    wa.save([save_meta("q2", other, list(range(40, 56)))])
    wa.reap()
    assert not wa.tiers.presence(keys, ct).all()
    sched_tiers.close()


def test_mixed_sources_device_host_device(box):
    eid = f"m{os.getpid()}"
    la, lb = mk_layers(3), mk_layers(4)
    ea, wa, _ = replica(box, "A" + eid, la, n_slots=8)
    eb, wb, bases_b = replica(box, "B" + eid, lb, with_tier=False)
    bases_b["peer"] = wa.local_tier.base
    n = 4 * C
    toks = np.arange(100, 100 + n, dtype=np.int32)
    blocks_a = list(range(0, 16))
    wa.save([save_meta("r", toks, blocks_a)])
    wa.reap()
    keys = ea._keys(toks)
    # chunk 1 leaves A's tier; B's own host pool happens to hold chunks 0..1 (an earlier turn)
    wa.local_tier.index.acquire(int(keys[1]))
    wa.local_tier.index.release(int(keys[1]))
    assert wa.local_tier.index.clear() is True or True
    for k, c in zip(keys, range(4)):                          # re-create the tier with chunk 1 missing
        if c != 1:
            slot = wa.local_tier.index.reserve(int(k), C, 0, 1)
            HBM[wa.local_tier.base + slot * 4096] = ko.gather_tokens(la, ko.slot_mapping_from_blocks(blocks_a, BS, n)[c * C:(c + 1) * C])
            wa.local_tier.index.commit(int(k))
    sm_a = ko.slot_mapping_from_blocks(blocks_a, BS, n)
    eb.oe.store(toks[:2 * C], np.ones(2 * C, bool), la, sm_a[:2 * C])    # same bytes as A computed
    blocks_b = list(range(20, 36))
    wb.start_load([load_meta("r", toks, blocks_b, n)])
    assert eb.calls == [("scatter_chunks", C, 1), ("retrieve", 2 * C, C), ("scatter_chunks", 2 * C, 2)]
    sm_b = ko.slot_mapping_from_blocks(blocks_b, BS, n)
    for x, y in zip(la, lb):
        assert np.array_equal(x.reshape(2, -1, 2, 8)[:, sm_a], y.reshape(2, -1, 2, 8)[:, sm_b])
    assert wb.stats.num_tier_peer_tokens == 3 * C and wb.stats.num_loaded_tokens == n


def test_no_tier_hit_keeps_the_ordinary_path_and_vanished_peers_are_forgotten(box, tmp_path):
    eid = f"v{os.getpid()}"
    la, lb = mk_layers(5), mk_layers(6)
    ea, wa, _ = replica(box, "A" + eid, la)
    eb, wb, bases_b = replica(box, "B" + eid, lb, with_tier=False)
    bases_b["peer"] = wa.local_tier.base
    toks = np.arange(0, 2 * C, dtype=np.int32)
    sm = ko.slot_mapping_from_blocks(list(range(8)), BS, 2 * C)
    eb.oe.store(toks, np.ones(2 * C, bool), lb, sm)
    wb.start_load([load_meta("r", toks, list(range(8)), 2 * C)], layers_per_group=2)
    assert eb.calls == [("retrieve", 2 * C, 0)] and wb.layer_loads and wb.layer_loads[0][0] == 3   # layer-wise host load
    assert "A" + eid in wb.tiers.views
    wa.local_tier.close()                                       # replica A goes away
    assert not os.path.exists(tier_path("A" + eid))
    wb.tiers.refresh(force=True)
    assert "A" + eid not in wb.tiers.views

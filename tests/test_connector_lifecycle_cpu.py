"""The whole plugin lifecycle on CPU: a SCHEDULER-role and a WORKER-role B200KVConnector (the real
classes, as vLLM constructs them) share one shm pool; only the CUDA engine is replaced by a stand-in
with KVEngine's signatures that moves bytes with the oracle.  Two turns of one conversation:

  turn 1: miss -> forward -> wait_for_save stores the prompt
  turn 2: lookup hits turn 1's whole chunks -> start_load_kv (layer-wise) -> wait_for_layer_load per
          layer -> pages equal what turn 1 computed -> wait_for_save stores only the new chunks

plus the stats the worker reports (lmcache:* series), the "hooks never ran" safety net, and shutdown
(KVConnectorBase_V1 call order: vllm/v1/worker/kv_connector_model_runner_mixin.py:85-119)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest

vllm = pytest.importorskip("vllm")
import torch  # noqa: E402

from b200kv import _lib  # noqa: E402
from b200kv.engine import chunk_keys  # noqa: E402
from oracle import kv_oracle as ko  # noqa: E402

BS, C = 16, 64
L, H, D, NB = 4, 2, 64, 64


def fake_vllm_config(engine_id):
    from vllm.config import KVTransferConfig
    ktc = KVTransferConfig(kv_connector="B200KVConnector", kv_connector_module_path="b200kv.connector",
                           kv_role="kv_both", engine_id=engine_id)
    mc = NS(model="synth-llama", dtype=torch.bfloat16, get_num_layers=lambda pc: L, get_num_kv_heads=lambda pc: H,
            get_head_size=lambda: D)
    return NS(kv_transfer_config=ktc, model_config=mc, parallel_config=NS(tensor_parallel_size=1, rank=0),
              cache_config=NS(block_size=BS, cache_dtype="auto"),
              scheduler_config=NS(disable_hybrid_kv_cache_manager=False))


class OracleEngine:
    """KVEngine's surface as the connector uses it; bytes move with the oracle, presence is mirrored
    into the real pool index so the scheduler role's lookups see it."""

    instances = []

    def __init__(self, geom, pool, device=0, staging_bytes=0, owner=0, variant=0, stages=0, ctas_per_sm=0,
                 key_seed=None, numa_policy=0):
        self.geom, self.pool, self.owner = geom, pool, owner
        self.key_seed = geom.key_seed() if key_seed is None else key_seed
        self.oe = ko.OracleEngine(geom.chunk_tokens)
        self.calls = []
        OracleEngine.instances.append(self)

    def numa_placement(self):
        return "stand-in"

    def register_kv_caches(self, tensors, layout=None):
        self.layers = [t.view(torch.int16).numpy().view(np.uint16) for t in tensors]   # share memory

    def _keys(self, tokens):
        return chunk_keys(tokens, self.geom.chunk_tokens, self.key_seed, True)

    def store(self, tokens, mask=None, slot_mapping=None, offset=0, stream=None, keys=None):
        self.calls.append(("store", len(tokens), offset))
        self.oe.store(np.asarray(tokens), mask, self.layers, slot_mapping, offset)
        Ct = self.geom.chunk_tokens
        for c, k in enumerate(self._keys(tokens)):
            if c * Ct < offset:
                continue
            try:
                self.pool.reserve(int(k), min(Ct, len(tokens) - c * Ct), 0, self.owner)
                self.pool.commit(int(k))
            except _lib.B200KVError as e:
                assert e.code == _lib.EEXIST
        return 1

    def retrieve(self, tokens, mask=None, slot_mapping=None, stream=None, return_ticket=False, layers_per_group=0):
        self.calls.append(("retrieve", len(tokens), int((~mask).sum()), layers_per_group))
        ret = self.oe.retrieve(np.asarray(tokens), mask, self.layers, slot_mapping)
        return (ret, 5) if return_ticket else ret

    def wait_layer(self, ticket, layer, stream=None):
        self.calls.append(("wait_layer", ticket, layer))

    def poll(self, ticket):
        return True

    def export_ipc(self):
        raise RuntimeError("no CUDA IPC on CPU")

    def wait_all(self):
        pass

    def close(self):
        self.calls.append(("close",))


def blocks_ns(ids):
    return NS(get_block_ids=lambda: (list(ids),))


def sched_output(new=(), cached=None, num_sched=None, finished=()):
    cached = cached or NS(req_ids=[], new_block_ids=[], resumed_req_ids=set(), all_token_ids={})
    return NS(scheduled_new_reqs=list(new), scheduled_cached_reqs=cached, num_scheduled_tokens=num_sched or {},
              finished_req_ids=set(finished))


def run_step(worker, meta, layer_names, call_layer_hooks=True):
    fc = NS(cudagraph_runtime_mode=NS(name="NONE"))
    worker.bind_connector_metadata(meta)
    worker.start_load_kv(fc)
    if call_layer_hooks:
        for name in layer_names:
            worker.wait_for_layer_load(name)
            worker.save_kv_layer(name, None, None)
    worker.wait_for_save()
    fin = worker.get_finished(set())
    bad = worker.get_block_ids_with_load_errors()
    stats = worker.get_kv_connector_stats()
    worker.clear_connector_metadata()
    return fin, bad, stats


def test_two_turn_conversation_through_both_roles(monkeypatch):
    from vllm.distributed.kv_transfer.kv_connector.v1.base import KVConnectorRole

    import b200kv.connector as bc
    monkeypatch.setattr(bc, "KVEngine", OracleEngine)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: None)
    monkeypatch.setenv("LMCACHE_MAX_LOCAL_CPU_SIZE", "0.05")
    monkeypatch.setenv("LMCACHE_CHUNK_SIZE", str(C))
    monkeypatch.setenv("B200KV_LAYER_GROUP", "2")
    monkeypatch.setenv("B200KV_ASYNC_LOAD", "0")
    monkeypatch.setenv("B200KV_LAYERWISE", "1")
    OracleEngine.instances.clear()
    cfg = fake_vllm_config(f"life{os.getpid()}x{os.urandom(3).hex()}")
    sched = bc.B200KVConnector(cfg, KVConnectorRole.SCHEDULER, None)
    worker = bc.B200KVConnector(cfg, KVConnectorRole.WORKER, None)
    pool_file = "/dev/shm" + sched._pool_name
    try:
        g = torch.Generator().manual_seed(0)
        caches = {f"model.layers.{i}.self_attn.attn": torch.randn((2, NB, BS, H, D), generator=g).bfloat16()
                  for i in range(L)}
        worker.register_kv_caches(caches)
        names = list(caches)
        eng = OracleEngine.instances[-1]
        assert worker._pdw is None                 # IPC publication failed softly: offload still works

        # ---------------- turn 1: 200-token prompt, nothing cached ----------------
        p1 = [int(x) for x in np.random.default_rng(1).integers(0, 30000, 200)]
        r1 = NS(request_id="r1", prompt_token_ids=p1, num_tokens=200, all_token_ids=p1, kv_transfer_params=None,
                sampling_params=None)
        assert sched.get_num_new_matched_tokens(r1, 0) == (0, False)
        b1 = list(range(3, 16))                    # 13 blocks = 208 slots
        sched.update_state_after_alloc(r1, blocks_ns(b1), 0)
        meta = sched.build_connector_meta(sched_output(
            [NS(req_id="r1", prompt_token_ids=p1, block_ids=(b1,), num_computed_tokens=0, sampling_params=None)],
            num_sched={"r1": 200}))
        want = [c.clone() for c in caches.values()]          # what "the forward pass" left in the pages
        fin, bad, stats = run_step(worker, meta, names)
        assert fin == (None, None) and bad == set()
        assert eng.calls[-1] == ("store", 200, 0)
        assert stats.data["num_stored_tokens"] == 200 and stats.data["num_requested_tokens"] == 200
        assert stats.data["num_hit_tokens"] == 0 and stats.data["local_cache_usage_bytes"] == 4 * worker._engine.geom.chunk_bytes
        assert sched.request_finished(r1, b1) == (False, None)
        sched.build_connector_meta(sched_output(finished=["r1"]))

        # ---------------- turn 2: same history + 70 new tokens, other blocks ----------------
        p2 = p1 + [int(x) for x in np.random.default_rng(2).integers(0, 30000, 70)]
        r2 = NS(request_id="r2", prompt_token_ids=p2, num_tokens=270, all_token_ids=p2, kv_transfer_params=None,
                sampling_params=None)
        n, is_async = sched.get_num_new_matched_tokens(r2, 0)
        assert (n, is_async) == (192, False)       # three whole chunks; turn 1's 8-token tail has another key now
        b2 = list(range(27, 44))
        sched.update_state_after_alloc(r2, blocks_ns(b2), 192)
        meta = sched.build_connector_meta(sched_output(
            [NS(req_id="r2", prompt_token_ids=p2, block_ids=(b2,), num_computed_tokens=192, sampling_params=None)],
            num_sched={"r2": 78}))
        n_calls = len(eng.calls)
        fin, bad, stats = run_step(worker, meta, names)
        assert bad == set()
        new_calls = eng.calls[n_calls:]
        assert new_calls[0] == ("retrieve", 192, 0, 2)                                  # layer-wise, groups of 2
        assert [c for c in new_calls if c[0] == "wait_layer"] == [("wait_layer", 5, 0), ("wait_layer", 5, 2)]
        assert new_calls[-1] == ("store", 270, 192)                                     # only the new chunks
        sm1 = ko.slot_mapping_from_blocks(b1, BS, 192)
        sm2 = ko.slot_mapping_from_blocks(b2, BS, 192)
        for t, w in zip(caches.values(), want):
            got = t.view(torch.int16).reshape(2, NB * BS, H, D)[:, sm2]
            exp = w.view(torch.int16).reshape(2, NB * BS, H, D)[:, sm1]
            assert torch.equal(got, exp)                                                # turn 1's KV, bit for bit
        d = stats.data
        assert d["num_loaded_tokens"] == 192 and d["num_hit_tokens"] == 192 and d["num_requested_tokens"] == 270
        assert d["num_stored_tokens"] == 78 and d["retrieve_calls"] == 1
        red = stats.reduce()
        assert red["num_hit_tokens"] == 192 and red["num_loaded_tokens"] == 192

        # ---------------- a step whose per-layer hooks never run (full CUDA graph replay) ----------------
        r3 = NS(request_id="r3", prompt_token_ids=p2, num_tokens=270, all_token_ids=p2, kv_transfer_params=None,
                sampling_params=None)
        assert sched.get_num_new_matched_tokens(r3, 0)[0] == 269          # full hit: the last token is recomputed
        b3 = list(range(44, 61))
        sched.update_state_after_alloc(r3, blocks_ns(b3), 269)
        meta = sched.build_connector_meta(sched_output(
            [NS(req_id="r3", prompt_token_ids=p2, block_ids=(b3,), num_computed_tokens=269, sampling_params=None)],
            num_sched={"r3": 1}))
        fin, bad, stats = run_step(worker, meta, names, call_layer_hooks=False)
        assert bad == set(b3) and worker.cfg.layerwise is False           # reported for recompute, path switched off
        # the same step when vLLM says it replays ONE full graph: chunk-wise load from the start
        worker.cfg.layerwise = True
        worker.bind_connector_metadata(sched.build_connector_meta(sched_output()))
        worker.start_load_kv(NS(cudagraph_runtime_mode=NS(name="FULL")))
        worker.clear_connector_metadata()
        assert os.path.exists(pool_file)
    finally:
        worker.shutdown()
        sched.shutdown()
    assert not os.path.exists(pool_file)           # a per-engine segment goes away with its engine
    assert ("close",) in OracleEngine.instances[-1].calls

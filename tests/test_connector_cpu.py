"""Plugin-slot checks that need no GPU: vLLM's own factory resolves the connector by module path,
HMA support is advertised, the scheduler role works against the shm index, the config surface
honours the reference's env names, and the `lmcache`-named alias resolves (chart-level drop-in)."""
import os
import sys
from types import SimpleNamespace as NS

import numpy as np
import pytest

vllm = pytest.importorskip("vllm")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_vllm_config(engine_id, extra=None, role="kv_both"):
    import torch
    from vllm.config import KVTransferConfig
    ktc = KVTransferConfig(kv_connector="B200KVConnector", kv_connector_module_path="b200kv.connector",
                           kv_role=role, engine_id=engine_id, kv_connector_extra_config=extra or {})
    mc = NS(model="synth-llama", dtype=torch.bfloat16, get_num_layers=lambda pc: 4, get_num_kv_heads=lambda pc: 2,
            get_head_size=lambda: 64)
    return NS(kv_transfer_config=ktc, model_config=mc, parallel_config=NS(tensor_parallel_size=1, rank=0),
              cache_config=NS(block_size=16, cache_dtype="auto"),
              scheduler_config=NS(disable_hybrid_kv_cache_manager=False))


def test_factory_resolves_class_and_hma():
    from vllm.distributed.kv_transfer.kv_connector.factory import KVConnectorFactory
    from vllm.distributed.kv_transfer.kv_connector.v1.base import supports_hma
    cfg = fake_vllm_config("abi")
    cls = KVConnectorFactory.get_connector_class(cfg.kv_transfer_config)
    from b200kv.connector import B200KVConnector
    assert cls is B200KVConnector and supports_hma(cls)
    assert cls.get_required_kvcache_layout(cfg) == "NHD"


def test_scheduler_role_lookup_and_metadata(monkeypatch):
    from vllm.distributed.kv_transfer.kv_connector.v1.base import KVConnectorRole

    from b200kv import KVPool, chunk_keys
    from b200kv.connector import B200KVConnector, B200KVConnectorMetadata, geometry_from_vllm
    monkeypatch.setenv("LMCACHE_MAX_LOCAL_CPU_SIZE", "0.05")
    monkeypatch.setenv("LMCACHE_CHUNK_SIZE", "64")
    monkeypatch.setenv("B200KV_ASYNC_LOAD", "0")        # the synchronous (LMCache-like) flow; async: test_adapter.py
    eid = f"t{os.getpid()}x{os.urandom(3).hex()}"
    cfg = fake_vllm_config(eid)
    conn = B200KVConnector(cfg, KVConnectorRole.SCHEDULER, None)
    try:
        geom = geometry_from_vllm(cfg, conn.cfg)
        assert geom.chunk_tokens == 64 and geom.chunk_bytes == 4 * 2 * 64 * 2 * 64 * 2
        prompt = list(range(200))
        req = NS(request_id="r1", prompt_token_ids=prompt, num_tokens=200, all_token_ids=prompt)
        assert conn.get_num_new_matched_tokens(req, 0) == (0, False)
        # a worker (here: the test) commits the first two chunks into the shared index
        seed = geom.key_seed("synth-llama", 1, 0)
        keys = chunk_keys(np.asarray(prompt, np.int32), 64, seed)
        for k in keys[:2]:
            conn._pool.reserve(int(k), 64, 0, 0)
            conn._pool.commit(int(k))
        assert conn.get_num_new_matched_tokens(req, 0) == (128, False)
        assert conn.get_num_new_matched_tokens(req, 64) == (64, False)
        conn.update_state_after_alloc(req, None, 64)
        so = NS(scheduled_new_reqs=[NS(req_id="r1", prompt_token_ids=prompt, block_ids=(list(range(13)),),
                                       num_computed_tokens=128)],
                scheduled_cached_reqs=NS(req_ids=[], new_block_ids=[], resumed_req_ids=set(), all_token_ids={}),
                num_scheduled_tokens={"r1": 72}, finished_req_ids=set())
        meta = conn.build_connector_meta(so)
        assert isinstance(meta, B200KVConnectorMetadata) and len(meta.requests) == 1
        m = meta.requests[0]
        assert m.load_spec.can_load and m.load_spec.external_cached_tokens == 128 and m.load_spec.vllm_cached_tokens == 64
        import pickle
        assert pickle.loads(pickle.dumps(meta)).requests[0].block_ids == list(range(13))   # crosses processes
        # this engine's own lookup traffic rides along (3 lookups of a 200-token prompt: 0 + 128 + 128 hit)
        assert pickle.loads(pickle.dumps(meta)).sched_counters == (3, 256, 600)
        assert conn.request_finished(req, []) == (False, None)
        assert conn.request_finished_all_groups(req, ([],)) == (False, None)
    finally:
        conn.shutdown()
        KVPool.unlink("/b200kv-eng-" + eid)


def test_config_surface_uses_reference_env_names():
    from b200kv.config import B200KVConfig
    from b200kv import FMT_FP8, FMT_RAW
    env = {"LMCACHE_CHUNK_SIZE": "128", "LMCACHE_LOCAL_CPU": "True", "LMCACHE_MAX_LOCAL_CPU_SIZE": "60",
           "LMCACHE_REMOTE_SERDE": "cachegen", "LMCACHE_LMCACHE_INSTANCE_ID": "pod-7", "LMCACHE_ENABLE_CONTROLLER": "True",
           "LMCACHE_CONTROLLER_PULL_URL": "router:9000", "LMCACHE_LMCACHE_WORKER_HEARTBEAT_TIME": "3",
           "LMCACHE_MAX_LOCAL_DISK_SIZE": "10", "LMCACHE_USE_EXPERIMENTAL": "True"}
    c = B200KVConfig.from_env(env)
    assert (c.chunk_size, c.max_local_cpu_size_gb, c.fmt, c.instance_id) == (128, 60.0, FMT_FP8, "pod-7")
    assert c.enable_controller and c.controller_pull_url == "router:9000" and c.worker_heartbeat_s == 3.0
    assert c.pool_bytes == 60 << 30
    c2 = B200KVConfig.from_env({}).apply_extra({"b200kv.format": "fp8", "lmcache.chunk_size": 512, "unrelated": 1})
    assert c2.fmt == FMT_FP8 and c2.chunk_size == 512
    assert B200KVConfig.from_env({"LMCACHE_REMOTE_SERDE": "naive"}).fmt == FMT_RAW
    with pytest.raises(ValueError):
        B200KVConfig.from_env({"B200KV_FORMAT": "int4"})


def test_lmcache_named_alias_resolves():
    sys.path.insert(0, os.path.join(ROOT, "production-stack_b200", "compat"))
    try:
        from lmcache.integration.vllm.vllm_v1_adapter import LMCacheConnectorV1Impl
        import inspect
        assert list(inspect.signature(LMCacheConnectorV1Impl.__init__).parameters)[1:] == ["vllm_config", "role", "parent"]
        for name in ("register_kv_caches", "start_load_kv", "wait_for_layer_load", "save_kv_layer", "wait_for_save",
                     "get_finished", "get_block_ids_with_load_errors", "get_kv_events", "get_num_new_matched_tokens",
                     "update_state_after_alloc", "build_connector_meta", "request_finished"):
            assert callable(getattr(LMCacheConnectorV1Impl, name))   # what lmcache_connector.py:120-354 forwards
    finally:
        sys.path.pop(0)


def test_vllm_builtin_lmcache_wrapper_drives_this_engine(monkeypatch):
    """The chart's literal `{"kv_connector":"LMCacheConnectorV1"}`: vLLM's own wrapper class
    (lmcache_connector.py:72-354) imports `lmcache.integration.vllm.vllm_v1_adapter` — here the compat
    tree — and forwards the scheduler calls to this engine."""
    from vllm.config import KVTransferConfig
    from vllm.distributed.kv_transfer.kv_connector.factory import KVConnectorFactory
    from vllm.distributed.kv_transfer.kv_connector.v1.base import KVConnectorRole

    from b200kv import KVPool, chunk_keys
    from b200kv.connector import geometry_from_vllm
    monkeypatch.syspath_prepend(os.path.join(ROOT, "production-stack_b200", "compat"))
    monkeypatch.setenv("LMCACHE_MAX_LOCAL_CPU_SIZE", "0.05")
    monkeypatch.setenv("LMCACHE_CHUNK_SIZE", "64")
    for mod in [m for m in sys.modules if m == "lmcache" or m.startswith("lmcache.")]:
        monkeypatch.delitem(sys.modules, mod)
    cfg = fake_vllm_config(f"w{os.getpid()}x{os.urandom(3).hex()}")
    cfg.kv_transfer_config = KVTransferConfig(kv_connector="LMCacheConnectorV1", kv_role="kv_both",
                                              engine_id=cfg.kv_transfer_config.engine_id)
    cls = KVConnectorFactory.get_connector_class(cfg.kv_transfer_config)
    assert cls.__name__ == "LMCacheConnectorV1" and cls.__module__.startswith("vllm.")
    conn = cls(cfg, KVConnectorRole.SCHEDULER, None)
    inner = conn._lmcache_engine._inner
    try:
        assert type(conn._lmcache_engine).__module__ == "lmcache.integration.vllm.vllm_v1_adapter"
        prompt = list(range(200))
        req = NS(request_id="r1", prompt_token_ids=prompt, num_tokens=200, all_token_ids=prompt)
        assert conn.get_num_new_matched_tokens(req, 0) == (0, False)
        geom = geometry_from_vllm(cfg, inner.cfg)
        for k in chunk_keys(np.asarray(prompt, np.int32), 64, geom.key_seed("synth-llama", 1, 0))[:2]:
            inner._pool.reserve(int(k), 64, 0, 0)
            inner._pool.commit(int(k))
        assert conn.get_num_new_matched_tokens(req, 0) == (128, False)
        conn.update_state_after_alloc(req, None, 128)
        so = NS(scheduled_new_reqs=[NS(req_id="r1", prompt_token_ids=prompt, block_ids=(list(range(13)),),
                                       num_computed_tokens=128)],
                scheduled_cached_reqs=NS(req_ids=[], new_block_ids=[], resumed_req_ids=set(), all_token_ids={}),
                num_scheduled_tokens={"r1": 72}, finished_req_ids=set())
        meta = conn.build_connector_meta(so)
        assert meta.requests[0].load_spec.can_load and meta.requests[0].load_spec.external_cached_tokens == 128
        assert conn.request_finished(req, []) == (False, None)
        # the wrapper passes update_state_after_alloc no block ids: what would need them is declined on this
        # route (a remote prefill is then an ordinary request: pool lookup / recompute), loads stay in-step
        assert inner._pd.blocks_known_at_alloc is False and inner._sched.async_load is False
        pd_req = NS(request_id="r2", prompt_token_ids=prompt, num_tokens=200, all_token_ids=prompt,
                    kv_transfer_params={"do_remote_prefill": True, "remote_engine_id": "P", "remote_block_ids": [[1, 2, 3]],
                                        "remote_num_tokens": 200})
        assert conn.get_num_new_matched_tokens(pd_req, 0) == (128, False)      # the pool's answer, not 199 "remote" tokens
    finally:
        inner.shutdown()
        KVPool.unlink(inner._pool_name)


def test_lmcache_prometheus_series_names_and_aggregation():
    """The series production-stack's Grafana dashboard queries
    (/root/reference/helm/dashboards/lmcache-dashboard.json:204,295,389,453,517)."""
    from prometheus_client import CollectorRegistry, Counter, Gauge, Histogram, generate_latest

    from b200kv.connector import B200KVConnector
    from b200kv.metrics import B200KVStats
    reg = CollectorRegistry()

    def bind(cls):
        return lambda **kw: cls(registry=reg, **{k: v for k, v in kw.items() if k != "multiprocess_mode"})

    cfg = fake_vllm_config("prom")
    pm = B200KVConnector.build_prom_metrics(cfg, {Gauge: bind(Gauge), Counter: bind(Counter), Histogram: bind(Histogram)},
                                            ["model_name", "engine"], {0: ["synth", "0"]})
    a = B200KVConnector.build_kv_connector_stats({"num_hit_tokens": 512, "num_requested_tokens": 600,
                                                  "num_loaded_tokens": 512, "retrieve_seconds": 0.004,
                                                  "retrieve_calls": 2, "retrieve_bytes": 512 * 131072,
                                                  "local_cache_usage_bytes": 1 << 30})
    b = B200KVStats({"num_hit_tokens": 8, "num_requested_tokens": 8, "num_stored_tokens": 256})
    assert not a.is_empty() and B200KVStats().is_empty()
    agg = a.aggregate(b)
    assert agg.data["num_hit_tokens"] == 520 and agg.reduce()["retrieve_GBps"] > 1
    pm.observe(agg.data, 0)
    text = generate_latest(reg).decode()
    for series in ("lmcache:num_hit_tokens_total", "lmcache:num_requested_tokens_total", "lmcache:local_cache_usage",
                   "lmcache:retrieve_speed_sum", "lmcache:retrieve_speed_count"):
        assert series in text, series
    assert 'lmcache:num_hit_tokens_total{engine="0",model_name="synth"} 520.0' in text


def test_config_file_and_per_request_skip_save(tmp_path):
    from types import SimpleNamespace as NS

    from b200kv.adapter import SchedulerState, request_skip_save
    from b200kv.config import B200KVConfig
    f = tmp_path / "lmcache.yaml"
    f.write_text("chunk_size: 128\nlocal_cpu: true\nmax_local_cpu_size: 12\nremote_serde: cachegen\n")
    c = B200KVConfig.from_env({"LMCACHE_CONFIG_FILE": str(f), "LMCACHE_MAX_LOCAL_CPU_SIZE": "40"})
    assert c.chunk_size == 128 and c.fmt == 1 and c.max_local_cpu_size_gb == 40.0     # env beats file
    req = NS(req_id="r", prompt_token_ids=list(range(300)), block_ids=([1] * 19,), num_computed_tokens=0,
             sampling_params=NS(extra_args={"kv_transfer_params": {"lmcache.skip_save": True}}))
    assert request_skip_save(req)
    sched = SchedulerState(lambda t: 0, 16, 256, False)
    out = NS(scheduled_new_reqs=[req], scheduled_cached_reqs=NS(req_ids=[], new_block_ids=[], resumed_req_ids=set(), all_token_ids={}),
             num_scheduled_tokens={"r": 300}, finished_req_ids=set())
    assert sched.build_meta(out) == []                                                 # nothing saved, nothing loaded


def test_scheduler_waits_for_remote_prefetch_then_hits(monkeypatch):
    """LMCACHE_REMOTE_URL=lm://...: a prompt whose chunks sit only on the cache server makes
    get_num_new_matched_tokens answer (None, False) — vLLM's "ask again" — until they have landed in
    the LOCAL pool; afterwards it is an ordinary local hit (helm/templates/deployment-vllm-multi.yaml:
    338-345; KVConnectorBase_V1.get_num_new_matched_tokens, base.py:453-486)."""
    import time

    from vllm.distributed.kv_transfer.kv_connector.v1.base import KVConnectorRole

    from b200kv import KVPool, _lib, chunk_keys
    from b200kv.connector import B200KVConnector, geometry_from_vllm
    from b200kv.remote import RemoteClient, RemoteServer
    srv = RemoteServer("127.0.0.1", 0, 64 << 20)
    monkeypatch.setenv("LMCACHE_MAX_LOCAL_CPU_SIZE", "0.05")
    monkeypatch.setenv("LMCACHE_CHUNK_SIZE", "64")
    monkeypatch.setenv("B200KV_ASYNC_LOAD", "0")
    monkeypatch.setenv("LMCACHE_REMOTE_URL", f"lm://127.0.0.1:{srv.port}")
    cfg = fake_vllm_config(f"t{os.getpid()}r{os.urandom(3).hex()}")
    conn = B200KVConnector(cfg, KVConnectorRole.SCHEDULER, None)
    try:
        geom = geometry_from_vllm(cfg, conn.cfg)
        prompt = list(range(200))
        keys = chunk_keys(np.asarray(prompt, np.int32), 64, geom.key_seed("synth-llama", 1, 0))
        # another replica stored chunks 0..2 and uploaded them
        other = KVPool(None, 8 * geom.chunk_bytes, geom.chunk_bytes, _lib.POOL_CREATE)
        c = RemoteClient("127.0.0.1", srv.port)
        for i, k in enumerate(keys[:3]):
            slot = other.reserve(int(k), 64, 0, 5)
            other.slot_view(slot)[:] = i + 1
            other.commit(int(k))
            assert c.put(other, int(k), 5) == 0
        req = NS(request_id="r1", prompt_token_ids=prompt, num_tokens=200, all_token_ids=prompt)
        assert conn.get_num_new_matched_tokens(req, 0) == (None, False)
        t0 = time.time()
        while True:
            n, is_async = conn.get_num_new_matched_tokens(req, 0)
            if n is not None:
                break
            assert time.time() - t0 < 10
            time.sleep(0.005)
        assert (n, is_async) == (192, False)
        slot, n_tok, fmt = conn._pool.acquire(int(keys[1]))
        assert n_tok == 64 and int(conn._pool.slot_view(slot)[0]) == 2     # the bytes the other replica stored
        conn._pool.release(int(keys[1]))
        assert conn.get_num_new_matched_tokens(req, 0) == (192, False)     # local now: answered at once
        assert conn.request_finished(req, []) == (False, None)
        c.close()
        other.close()
    finally:
        conn.shutdown()
        KVPool.unlink(conn._pool_name)
        srv.stop()


def test_shared_pool_name_is_per_geometry(monkeypatch):
    """Two models on one box with the same B200KV_POOL_NAME: one segment per chunk geometry."""
    from b200kv.config import B200KVConfig
    from b200kv.connector import pool_name_for
    cfg = B200KVConfig(pool_name="box")
    assert pool_name_for(None, cfg, 0x2000000) == "/box-2000000"
    assert pool_name_for(None, cfg, 0x1000800) == "/box-1000800"
    assert pool_name_for(NS(kv_transfer_config=NS(engine_id="e-1/x")), B200KVConfig(), 123) == "/b200kv-eng-e-1x"


@pytest.mark.parametrize("backend,layout", [("flashinfer", "HND"), ("flashinfer", "NHD"), ("flash_attn", "NHD"),
                                            ("flash_attn", "HND")])
def test_layout_detection_on_tensors_built_like_vllms_model_runner(backend, layout):
    """The physical KV layouts the engine must understand are whatever vLLM allocates: the attention
    backend's `get_kv_cache_shape` permuted by its `get_kv_cache_stride_order`, then viewed back in the logical
    order (vllm/v1/worker/gpu_model_runner.py:6935-6990; backends: v1/attention/backends/flashinfer.py,
    flash_attn.py).  Built here with vLLM's own functions (CPU tensors) and handed to `paged_layout_of`."""
    import torch
    from vllm.v1.attention.backends import utils as U

    from b200kv import _lib
    from b200kv.engine import paged_layout_of
    if backend == "flashinfer":
        from vllm.v1.attention.backends.flashinfer import FlashInferBackend as B
    else:
        from vllm.v1.attention.backends.flash_attn import FlashAttentionBackend as B
    NB, bs, H, D = 12, 16, 8, 128
    U.set_kv_cache_layout(layout)
    try:
        U.get_kv_cache_layout.cache_clear() if hasattr(U.get_kv_cache_layout, "cache_clear") else None
        shape = B.get_kv_cache_shape(NB, bs, H, D)
        order = B.get_kv_cache_stride_order()
    finally:
        U.set_kv_cache_layout(None)
        U.get_kv_cache_layout.cache_clear() if hasattr(U.get_kv_cache_layout, "cache_clear") else None
    phys = tuple(shape[i] for i in order)
    inv = [order.index(i) for i in range(len(order))]
    raw = torch.arange(int(np.prod(shape)), dtype=torch.int32).to(torch.bfloat16)
    t = raw.view(phys).permute(*inv)                          # exactly what the model runner registers
    assert tuple(t.shape) == tuple(shape)
    k, v, stride, nb, h, d, tile = paged_layout_of(t, bs)
    assert (nb, h, d) == (NB, H, D)
    assert tile == (_lib.LAYOUT_HND if layout == "HND" else _lib.LAYOUT_NHD)
    es = t.element_size()
    # K and V planes of block b, and the element (token j, head g, dim e) inside a tile, where vLLM put them
    for b, j, g, e in [(0, 0, 0, 0), (3, 5, 2, 7), (NB - 1, bs - 1, H - 1, D - 1)]:
        for kv, base in ((0, k), (1, v)):
            want = (t[kv, b, j, g, e] if tuple(shape)[0] == 2 else t[b, kv, j, g, e]).data_ptr()
            inner = (g * bs * D + j * D + e) if tile == _lib.LAYOUT_HND else (j * H * D + g * D + e)
            assert base + b * stride + inner * es == want, (backend, layout, kv, b, j, g, e)


def test_several_kv_cache_groups_are_refused():
    from vllm.distributed.kv_transfer.kv_connector.v1.base import KVConnectorRole

    from b200kv.connector import B200KVConnector
    with pytest.raises(ValueError, match="single KV-cache group"):
        B200KVConnector(fake_vllm_config("hyb"), KVConnectorRole.SCHEDULER, NS(kv_cache_groups=[object(), object()]))

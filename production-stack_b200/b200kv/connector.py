"""vLLM v1 KV-connector plugin: ``B200KVConnector``.

Loaded with
``--kv-transfer-config '{"kv_connector":"B200KVConnector","kv_connector_module_path":"b200kv.connector","kv_role":"kv_both"}'``
(vllm/distributed/kv_transfer/kv_connector/factory.py:96-128) — the slot production-stack fills
with ``LMCacheConnectorV1`` (helm/templates/deployment-vllm-multi.yaml:194-207,
operator/internal/controller/vllmruntime_controller.go:536-543).  Constructed twice by vLLM:
role SCHEDULER in the scheduler process (no CUDA) and role WORKER in each worker.

Lifecycle contract: KVConnectorBase_V1 (vllm/.../v1/base.py:171-674); the behaviour of every
method mirrors the LMCache adapter cited in b200kv/adapter.py.  The two roles share nothing but
the POSIX-shm pool index (b200kv_pool): no ZMQ lookup server, no PYTHONHASHSEED coupling.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass, field
from typing import TYPE_CHECKING, Any

import torch
from vllm.distributed.kv_transfer.kv_connector.v1.base import (KVConnectorBase_V1, KVConnectorMetadata,
                                                               KVConnectorRole, SupportsHMA)

from . import _lib
from .adapter import ReqMeta, SchedulerState, WorkerState, request_identity
from .config import B200KVConfig
from .engine import KVEngine, KVGeometry, KVPool, paged_layout_of, xxh64
from .pd import PDMeta, PDScheduler, PDWorker, first_group, publish_ipc, unpublish_ipc

if TYPE_CHECKING:
    from vllm.config import VllmConfig
    from vllm.forward_context import ForwardContext
    from vllm.v1.core.kv_cache_manager import KVCacheBlocks
    from vllm.v1.core.sched.output import SchedulerOutput
    from vllm.v1.kv_cache_interface import KVCacheConfig
    from vllm.v1.request import Request

logger = logging.getLogger("b200kv")

_TRACE_PATH = os.environ.get("B200KV_TRACE")   # file to append per-call host timings to (debug only)


def _traced(fn):
    if not _TRACE_PATH:
        return fn
    import functools
    import time as _t

    @functools.wraps(fn)
    def wrap(self, *a, **kw):
        t0 = _t.perf_counter()
        try:
            return fn(self, *a, **kw)
        finally:
            dt = (_t.perf_counter() - t0) * 1e3
            if dt > 0.05:
                with open(_TRACE_PATH, "a") as f:
                    f.write(f"{_t.time():.6f} {self._role.name} {fn.__name__} {dt:.3f} ms\n")
    return wrap


@dataclass
class B200KVConnectorMetadata(KVConnectorMetadata):
    requests: list[ReqMeta] = field(default_factory=list)
    pd: PDMeta = field(default_factory=PDMeta)   # disaggregated-prefill pulls / held requests
    # this engine's cumulative (lookups, hit tokens, requested tokens): the worker reports them as the
    # lmcache:* series, so replicas sharing one pool each report their own traffic, not the box's
    sched_counters: tuple = (0, 0, 0)


def geometry_from_vllm(vllm_config, cfg: B200KVConfig, n_blocks: int = 1) -> KVGeometry:
    """LMCache's kv_shape = (num_layer, 2, chunk, num_kv_head, head_size)
    (vllm_v1_adapter.py:471-477), derivable in both roles from the vLLM config alone."""
    mc, pc, cc = vllm_config.model_config, vllm_config.parallel_config, vllm_config.cache_config
    dt = str(cc.cache_dtype)
    if dt in ("auto", "None"):
        elem = torch.empty((), dtype=mc.dtype).element_size()
    else:
        elem = 1 if "fp8" in dt else 2
    return KVGeometry(n_layers=mc.get_num_layers(pc), n_kv_heads=mc.get_num_kv_heads(pc),
                      head_dim=mc.get_head_size(), n_blocks=n_blocks, block_tokens=cc.block_size,
                      chunk_tokens=cfg.chunk_size, elem_bytes=elem, block_stride_bytes=0, fmt=cfg.fmt)


def owner_tag_of(instance_id: str) -> int:
    """Stable 31-bit tag of LMCACHE_LMCACHE_INSTANCE_ID (not Python's salted hash())."""
    return (xxh64(instance_id.encode(), 0) & 0x7FFFFFFF) or 1


# per-engine segments carry vLLM's random engine id: nothing ever re-attaches to one whose engine died, so the
# scheduler role sweeps the unheld ones of earlier incarnations at start-up (k8s keeps a memory-backed emptyDir
# across container restarts inside a pod: a crash loop would otherwise fill /dev/shm)
ENGINE_POOL_PREFIX = "/b200kv-eng-"


def pool_name_for(vllm_config, cfg: B200KVConfig, chunk_bytes: int = 0) -> str:
    """One segment per engine unless B200KV_POOL_NAME names a shared one (BASELINE.json
    config 3: "shared pinned-host KV pool" across the replicas of a box).  A shared name is
    suffixed with the chunk size: replicas of DIFFERENT models on one box (the chart's modelSpec is
    a list) given the same name get one segment per geometry instead of failing to attach."""
    if cfg.pool_name:
        base = cfg.pool_name if cfg.pool_name.startswith("/") else "/" + cfg.pool_name
        return f"{base}-{chunk_bytes:x}" if chunk_bytes else base
    eid = vllm_config.kv_transfer_config.engine_id or "engine"
    return ENGINE_POOL_PREFIX + "".join(ch for ch in str(eid) if ch.isalnum() or ch in "-_")[:48]


class B200KVConnector(KVConnectorBase_V1, SupportsHMA):
    def __init__(self, vllm_config: "VllmConfig", role: KVConnectorRole,
                 kv_cache_config: "KVCacheConfig | None" = None):
        super().__init__(vllm_config=vllm_config, role=role, kv_cache_config=kv_cache_config)
        ktc = vllm_config.kv_transfer_config
        groups = getattr(kv_cache_config, "kv_cache_groups", None)
        if groups is not None and len(groups) > 1:
            # one paged cache per layer with ONE block table per request is what this engine moves; with several
            # KV-cache groups (hybrid / sliding-window models) block ids differ per group — refuse rather than
            # store pages under the wrong ids
            raise ValueError(f"B200KVConnector supports a single KV-cache group, the model has {len(groups)}: "
                             "start vLLM with --disable-hybrid-kv-cache-manager")
        self.cfg = B200KVConfig.from_env().apply_extra(ktc.kv_connector_extra_config)
        if role == KVConnectorRole.SCHEDULER and getattr(ktc, "kv_load_failure_policy", "recompute") == "fail":
            # vLLM's default (config/kv_transfer.py:70).  This connector never raises on the data path: a short or
            # failed load is REPORTED (get_block_ids_with_load_errors) so that the blocks are recomputed — under the
            # default policy vLLM fails such a request instead.
            logger.warning('b200kv: kv_load_failure_policy is "fail": a chunk that is evicted between lookup and load, an '
                           'expired P/D lease or an unreachable peer will FAIL the request; add '
                           '"kv_load_failure_policy":"recompute" to --kv-transfer-config to recompute instead')
        self.kv_role = ktc.kv_role
        self._block_size = vllm_config.cache_config.block_size
        self._chunk = self.cfg.chunk_size
        if self._chunk % self._block_size:
            raise ValueError(f"chunk size {self._chunk} must be a multiple of block size {self._block_size}")
        self._discard_partial = bool(ktc.get_from_extra_config("discard_partial_chunks", False)) \
            or not self.cfg.save_unfull_chunk
        geom = geometry_from_vllm(vllm_config, self.cfg)
        pc = vllm_config.parallel_config
        self._model = str(vllm_config.model_config.model)
        self._world = pc.tensor_parallel_size
        self._pool_name = pool_name_for(vllm_config, self.cfg, geom.chunk_bytes)
        if role == KVConnectorRole.SCHEDULER:
            try:
                age = int(self.cfg.extra.get("sweep_min_age_s", 60))
                n_swept = KVPool.sweep(ENGINE_POOL_PREFIX, age) + KVPool.sweep("/b200kv-dev-", age)   # pools + tier indices
                if n_swept:
                    logger.warning("b200kv: removed %d pool segment(s) left in /dev/shm by engines that died", n_swept)
            except Exception as e:     # hygiene only
                logger.debug("b200kv: segment sweep skipped (%s)", e)
        self._pool = KVPool(self._pool_name, self.cfg.pool_bytes, geom.chunk_bytes, _lib.POOL_CREATE_OR_ATTACH)
        self._engine: KVEngine | None = None
        self._worker: WorkerState | None = None
        self._sched: SchedulerState | None = None
        self._controller = None
        self._engine_id = str(ktc.engine_id or "engine")
        self._pd: PDScheduler | None = None
        self._pdw: PDWorker | None = None
        self._remote_computed: dict[str, int] = {}
        self._remote = None     # cache-server tier (LMCACHE_REMOTE_URL), b200kv/remote.py
        self._sched_counters: tuple = (0, 0, 0)   # worker: last counters received from this engine's scheduler
        if role == KVConnectorRole.SCHEDULER:
            self._pd = PDScheduler(self._engine_id, self._block_size,
                                   lease_s=float(self.cfg.extra.get("pd_lease_s", 120.0)), tp_size=self._world)
            seed = self._key_seed(0)
            lease = self.cfg.lookup_lease_ms
            chunk = self._chunk
            pool = self._pool
            include_partial = not self._discard_partial

            self._tiers = None
            if self.cfg.device_tier_gb > 0:      # index-only view of every replica's device tier
                from .device_tier import TierSet
                self._tiers = TierSet(self._engine_id, geom.chunk_bytes, None)
            tiers = self._tiers

            tier_off = os.environ.get("B200KV_TIER_DISABLE_FILE") or None

            def lookup(token_ids):
                if tiers is None or (tier_off and os.path.exists(tier_off)):   # kill switch: what the host pool holds
                    return pool.lookup_tokens(token_ids, chunk, seed, lease, include_partial)
                from .device_tier import combined_prefix_tokens
                from .engine import chunk_keys
                import numpy as np
                toks = np.asarray(token_ids, dtype=np.int32)
                keys = chunk_keys(toks, chunk, seed, include_partial)
                ct = np.minimum(chunk, len(toks) - np.arange(len(keys)) * chunk).astype(np.int32)
                tiers.refresh()
                return combined_prefix_tokens(pool, tiers, keys, ct, lease)

            self._sched = SchedulerState(lookup, self._block_size, self._chunk, self._discard_partial,
                                         self.cfg.save_decode_cache, self.kv_role, async_load=self.cfg.async_load,
                                         priority_limit=self.cfg.priority_limit)
            if self.kv_role != "kv_producer":     # the scheduler fetches every TP rank's chunks of a prompt
                self._remote = self._make_remote([self._key_seed(r) for r in range(self._world)])
            if not self.cfg.pool_name:
                # a per-engine segment dies with its engine (vLLM's engine ids are random: nobody would
                # ever attach to it again); a named, shared one stays for the other replicas of the box
                import atexit
                atexit.register(KVPool.unlink, self._pool_name)
        logger.info("b200kv connector role=%s kv_role=%s pool=%s (%.1f GB, chunk %d, fmt %s)",
                    role.name, self.kv_role, self._pool_name, self.cfg.max_local_cpu_size_gb, self._chunk,
                    ("raw", "fp8", "q4")[self.cfg.fmt])

    def _make_remote(self, key_seeds):
        """LMCACHE_REMOTE_URL=lm://host:port (deployment-vllm-multi.yaml:338-345): the cache-server
        tier behind the local pinned pool.  A malformed URL is a fatal misconfiguration."""
        from .remote import RemoteTier, parse_remote_url
        hp = parse_remote_url(self.cfg.remote_url)
        if hp is None:
            return None
        logger.info("b200kv remote tier: %s:%d", *hp)
        return RemoteTier(self._pool, hp[0], hp[1], self._chunk, key_seeds, owner=owner_tag_of(self.cfg.instance_id),
                          include_partial=not self._discard_partial, wait_s=self.cfg.remote_wait_ms / 1e3)

    def _key_seed(self, rank: int) -> int:
        """Same namespace in both roles: derived from the vLLM config only (the tile layout is a
        per-process constant of the attention backend, so it need not enter the scheduler's seed;
        chunks are additionally tagged with their format in the pool)."""
        return geometry_from_vllm(self._vllm_config, self.cfg).key_seed(self._model, self._world, rank)

    # ------------------------------------------------------------------ class-level hooks
    @classmethod
    def get_required_kvcache_layout(cls, vllm_config: "VllmConfig") -> str | None:
        # Both tile orders are supported; NHD is preferred when the backend leaves the choice
        # (on Blackwell vLLM's FlashInfer backend imposes HND, selector.py:124-133).
        return "NHD"

    # ------------------------------------------------------------------ worker side
    def register_kv_caches(self, kv_caches: dict[str, torch.Tensor]):
        """KVConnectorBase_V1.register_kv_caches (base.py:251) — adapter :786-795."""
        tensors = list(kv_caches.values())
        if not tensors:
            raise ValueError("no KV cache tensors to register")
        t0 = tensors[0]
        k, v, stride, nb, h, d, tile_layout = paged_layout_of(t0, self._block_size)
        geom0 = geometry_from_vllm(self._vllm_config, self.cfg)
        geom = KVGeometry(n_layers=len(tensors), n_kv_heads=h, head_dim=d, n_blocks=nb,
                          block_tokens=self._block_size, chunk_tokens=self._chunk,
                          elem_bytes=t0.element_size(), block_stride_bytes=stride, fmt=self.cfg.fmt,
                          layout=tile_layout)
        if geom.chunk_bytes != geom0.chunk_bytes:
            raise ValueError("KV cache tensors do not match the model geometry the pool was sized for "
                             f"({geom.chunk_bytes} vs {geom0.chunk_bytes} bytes per chunk)")
        rank = getattr(self._vllm_config.parallel_config, "rank", 0)
        self._rank = rank
        self._engine = KVEngine(geom, self._pool, device=t0.device.index or 0,
                                staging_bytes=self.cfg.staging_mb << 20, owner=owner_tag_of(self.cfg.instance_id),
                                variant=self.cfg.variant, key_seed=self._key_seed(rank),
                                # a pool shared by the box's replicas is interleaved over the sockets, a
                                # per-engine pool lives on the GPU's own node
                                numa_policy=_lib.NUMA_INTERLEAVE if self.cfg.pool_name else _lib.NUMA_LOCAL)
        logger.info("b200kv pool pages: %s", self._engine.numa_placement())
        self._engine.register_kv_caches(tensors)
        self._layer_index = {name: i for i, name in enumerate(kv_caches.keys())}
        self._layer_hooks_seen = 0
        self._worker = WorkerState(self._engine, self._block_size, self._chunk, self.kv_role,
                                   owner_tag=owner_tag_of(self.cfg.instance_id) if self.cfg.pool_name else 0)
        if self.cfg.device_tier_gb > 0:
            from .device_tier import LocalTier, TierSet
            fmt_tag = self.cfg.fmt | (tile_layout << 8)
            try:
                n_slots = max(1, int(self.cfg.device_tier_gb * (1 << 30)) // geom.chunk_bytes)
                self._worker.tiers = TierSet(self._engine_id, geom.chunk_bytes, fmt_tag, importer=self._engine.tier_import)
                self._worker.tier_disable_file = os.environ.get("B200KV_TIER_DISABLE_FILE") or None
                if self.kv_role != "kv_consumer":
                    self._worker.local_tier = LocalTier(self._engine, self._engine_id, n_slots, t0.device.index or 0,
                                                        fmt_tag, owner_tag_of(self.cfg.instance_id))
                    self._worker.tiers.add_local(self._worker.local_tier)
                logger.info("b200kv device tier: %d chunk slots in HBM", n_slots)
            except Exception as e:
                logger.warning("b200kv: device tier unavailable (%s)", e)
                self._worker.tiers = self._worker.local_tier = None
        if self.kv_role != "kv_consumer":
            self._remote = self._make_remote(self._engine.key_seed)
            if self._remote is not None:
                self._worker.on_stored = self._remote.push     # upload what this worker stores
        try:
            tp_rank = self._tp_rank()
            publish_ipc(self._engine_id, self._engine, t0.device.index or 0, tp_rank, self._world)   # peers may pull from us
            self._pdw = PDWorker(self._engine, self._engine_id, self._block_size, rank=tp_rank, tp_size=self._world)
        except Exception as e:  # e.g. VMM-allocated cache (sleep mode): P/D pull unavailable, offload still works
            logger.warning("b200kv: cannot publish CUDA-IPC descriptors (%s); peer pull disabled", e)
        if self.cfg.enable_controller and self.cfg.controller_pull_url and rank == 0:
            from .controller_client import ControllerClient
            self._controller = ControllerClient(
                self.cfg.controller_pull_url, self.cfg.instance_id, self._pool_name, self._engine.key_seed,
                self._chunk, owner_tag=owner_tag_of(self.cfg.instance_id) if self.cfg.pool_name else 0,
                include_partial=not self._discard_partial, heartbeat_s=self.cfg.worker_heartbeat_s,
                ip=self.cfg.advertise_ip)
        logger.info("b200kv registered %d layers, %d blocks, stride %d", len(tensors), nb, stride)

    def _tp_rank(self) -> int:
        """This worker's tensor-parallel rank (every rank of an engine shares the engine id)."""
        try:
            from vllm.distributed.parallel_state import get_tensor_model_parallel_rank
            return int(get_tensor_model_parallel_rank())
        except Exception:
            return int(getattr(self._vllm_config.parallel_config, "rank", 0) or 0) % max(self._world, 1)

    def _metas(self) -> list[ReqMeta]:
        md = self._get_connector_metadata()
        assert isinstance(md, B200KVConnectorMetadata)
        return md.requests

    @_traced
    def start_load_kv(self, forward_context: "ForwardContext", **kwargs: Any) -> None:
        if self._worker is None:
            return
        stream = torch.cuda.current_stream()
        md = self._get_connector_metadata()
        if isinstance(md, B200KVConnectorMetadata) and md.sched_counters[0] >= self._sched_counters[0]:
            self._sched_counters = tuple(md.sched_counters)
        if self._pdw is not None and isinstance(md, B200KVConnectorMetadata) and (md.pd.pulls or md.pd.held):
            self._pdw.start_pulls(md.pd, stream=stream)
        self._layer_hooks_seen = 0
        # A step replayed as ONE full CUDA graph never runs the per-layer hooks (base.py:591-611): use
        # the chunk-wise load there (e.g. a full-prompt hit that only recomputes its last token).
        full_graph = getattr(getattr(forward_context, "cudagraph_runtime_mode", None), "name", "NONE") == "FULL"
        self._worker.start_load(self._metas(), stream=stream,
                                layers_per_group=self.cfg.layer_group if (self.cfg.layerwise and not full_graph) else 0)

    def wait_for_layer_load(self, layer_name: str) -> None:
        """Chunk-wise loads are ordered before the forward pass by a stream wait.  Layer-wise loads
        (B200KV_LAYERWISE=1; LMCache `use_layerwise`, adapter :907-929): the compute stream waits
        here, once per layer group, while later groups are still crossing PCIe."""
        w = self._worker
        if w is None or not w.layer_loads:
            return
        idx = self._layer_index.get(layer_name)
        if idx is None:
            return
        self._layer_hooks_seen += 1
        if idx % self.cfg.layer_group == 0:
            w.wait_layer(idx, torch.cuda.current_stream())

    def save_kv_layer(self, layer_name: str, kv_layer: torch.Tensor, attn_metadata, **kwargs: Any) -> None:
        return  # whole-request store in wait_for_save: CUDA-graph replay skips per-layer hooks (base.py:591-611)

    @_traced
    def wait_for_save(self):
        if self._worker is None:
            return
        if self._worker.layer_loads and self._layer_hooks_seen == 0:
            # the per-layer hooks never ran in this step: do not trust the loaded pages, and stop
            # using the layer-wise path in this process
            logger.error("b200kv: wait_for_layer_load was not called during a step with layer-wise loads; "
                         "falling back to chunk-wise loads")
            self._worker.abandon_layer_loads()
            self.cfg.layerwise = False
        self._worker.layer_loads = []
        self._worker.save(self._metas(), stream=torch.cuda.current_stream())

    @_traced
    def get_finished(self, finished_req_ids: set[str]) -> tuple[set[str] | None, set[str] | None]:
        recv = None
        if self._worker is not None:
            self._worker.reap()
            recv = self._worker.poll_async_loads() or None    # detached pool loads that have landed
        sent = None
        if self._pdw is not None:
            sent, _pulled = self._pdw.poll(block=True)   # P/D pulls are synchronous for the scheduler
            sent = sent or None
        return sent, recv

    def get_block_ids_with_load_errors(self) -> set[int]:
        bad = self._worker.take_load_errors() if self._worker is not None else set()
        if self._pdw is not None:
            bad |= self._pdw.take_failed_blocks()
        return bad

    # ------------------------------------------------------------------ stats (lmcache:* series)
    @_traced
    def get_kv_connector_stats(self):
        """Worker side: deltas since the previous call (base.py:403)."""
        if self._worker is None or self._pool is None:
            return None
        from .metrics import B200KVStats
        ws = self._worker.stats
        ps = self._pool.stats()
        cur = {"num_stored_tokens": ws.num_stored_tokens, "num_loaded_tokens": ws.num_loaded_tokens,
               "retrieve_seconds": ws.retrieve_seconds, "retrieve_calls": ws.retrieve_calls,
               "load_shortfalls": ws.num_load_shortfalls, "num_foreign_loaded_tokens": ws.num_foreign_loaded_tokens,
               "num_tier_local_tokens": ws.num_tier_local_tokens, "num_tier_peer_tokens": ws.num_tier_peer_tokens,
               # the scheduler's counters are per engine, not per TP rank: rank 0 reports them
               "num_hit_tokens": self._sched_counters[1] if getattr(self, "_rank", 0) == 0 else 0,
               "num_requested_tokens": self._sched_counters[2] if getattr(self, "_rank", 0) == 0 else 0,
               "retrieve_bytes": ws.num_loaded_tokens * self._engine.geom.payload_bytes_per_token,
               "store_bytes": ws.num_stored_tokens * self._engine.geom.payload_bytes_per_token}
        prev = getattr(self, "_stats_prev", {})
        delta = {k: v - prev.get(k, 0) for k, v in cur.items()}
        self._stats_prev = cur
        if not any(delta.values()):
            return None
        delta["local_cache_usage_bytes"] = ps["n_used"] * ps["slot_bytes"]
        delta["local_cache_capacity_bytes"] = ps["n_slots"] * ps["slot_bytes"]
        return B200KVStats(delta)

    @classmethod
    def build_kv_connector_stats(cls, data: dict[str, Any] | None = None):
        from .metrics import B200KVStats
        return B200KVStats(data=data) if data is not None else B200KVStats()

    @classmethod
    def build_prom_metrics(cls, vllm_config, metric_types, labelnames, per_engine_labelvalues):
        from .metrics import B200KVPromMetrics
        return B200KVPromMetrics(vllm_config, metric_types, labelnames, per_engine_labelvalues)

    def shutdown(self):
        if self._remote is not None:
            self._remote.close()
            self._remote = None
        if self._controller is not None:
            self._controller.close()
            self._controller = None
        if self._pdw is not None:
            unpublish_ipc(self._engine_id, self._pdw.rank)
            self._pdw = None
        if self._pd is not None:
            self._pd.close()
        if self._worker is not None and self._worker.tiers is not None:
            if self._worker.local_tier is not None:
                self._worker.local_tier.close()
            self._worker.tiers.close()
            self._worker.tiers = self._worker.local_tier = None
        if getattr(self, "_tiers", None) is not None:
            self._tiers.close()
            self._tiers = None
        if self._engine is not None:
            self._engine.wait_all()
            self._engine.close()
            self._engine = None
        if self._pool is not None:
            self._pool.close()
            self._pool = None
            if self._sched is not None and not self.cfg.pool_name:
                KVPool.unlink(self._pool_name)    # mappings of the workers stay valid until they exit

    # ------------------------------------------------------------------ scheduler side
    @_traced
    def get_num_new_matched_tokens(self, request: "Request", num_computed_tokens: int) -> tuple[int | None, bool]:
        assert self._sched is not None and self._pd is not None
        remote = self._pd.remote_prefill_tokens(request, num_computed_tokens)
        if remote is not None:   # decode side of a disaggregated request: pull from the prefiller
            self._remote_computed[request.request_id] = num_computed_tokens
            return remote, False
        # multimodal items, cache_salt, LoRA adapter and lmcache.tag.* enter the chunk keys (adapter :1168-1172)
        ident = self._sched.identities.get(request.request_id) or request_identity(request)
        prompt = request.prompt_token_ids or []
        if self._remote is not None and self._remote.prefetch_state(
                request.request_id, prompt if ident is None else ident.apply(prompt)) == self._remote.PENDING:
            return None, False   # chunks are on their way from the cache server: ask again next step
        n = self._sched.num_new_matched_tokens(request.request_id, prompt,
                                               request.num_tokens, num_computed_tokens,
                                               int(getattr(request, "priority", 0) or 0), identity=ident)
        return n, bool(self._sched.async_load and n > 0)

    @_traced
    def update_state_after_alloc(self, request: "Request", blocks: "KVCacheBlocks", num_external_tokens: int):
        assert self._sched is not None and self._pd is not None
        if request.request_id in self._remote_computed:
            local = first_group(blocks.get_block_ids()) if blocks is not None else []
            self._pd.after_alloc(request, local, num_external_tokens, self._remote_computed.pop(request.request_id))
            self._sched.unfinished[request.request_id] = request
            return
        self._sched.after_alloc(request, num_external_tokens,
                                blocks.get_block_ids() if blocks is not None else None)

    @_traced
    def build_connector_meta(self, scheduler_output: "SchedulerOutput") -> KVConnectorMetadata:
        assert self._sched is not None and self._pd is not None
        sc = self._sched
        return B200KVConnectorMetadata(sc.build_meta(scheduler_output), self._pd.build_meta(),
                                       (sc.num_lookups, sc.num_hit_tokens, sc.num_requested_tokens))

    @_traced
    def request_finished(self, request: "Request", block_ids: list[int]) -> tuple[bool, dict[str, Any] | None]:
        # Offload: the gather that reads a request's pages is ordered before any later forward pass
        # on the compute stream, so blocks may be freed immediately.  Disaggregated prefill: keep
        # the pages (delay_free) until the decoder has pulled them (b200kv/pd.py).
        self._remote_computed.pop(request.request_id, None)   # looked up as a remote prefill but never allocated
        if self._remote is not None:
            self._remote.forget(request.request_id)
        if self._pd is not None:
            ok = True
            st = getattr(request, "status", None)
            if st is not None:
                ok = getattr(st, "name", str(st)) in ("FINISHED_LENGTH_CAPPED", "FINISHED_STOPPED")
            delay, params = self._pd.request_finished(request, block_ids, ok)
            if delay:
                return True, params
        return False, None

    def update_connector_output(self, connector_output):
        if self._pd is not None:
            self._pd.sending_finished(getattr(connector_output, "finished_sending", None))

    def request_finished_all_groups(self, request: "Request", block_ids: tuple[list[int], ...]):
        return self.request_finished(request, block_ids[0] if block_ids else [])

    def reset_cache(self) -> bool | None:
        return self._pool.clear() if self._pool is not None else None

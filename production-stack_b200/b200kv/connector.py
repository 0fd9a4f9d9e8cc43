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
from dataclasses import dataclass, field
from typing import TYPE_CHECKING, Any

import torch
from vllm.distributed.kv_transfer.kv_connector.v1.base import (KVConnectorBase_V1, KVConnectorMetadata,
                                                               KVConnectorRole, SupportsHMA)

from . import _lib
from .adapter import ReqMeta, SchedulerState, WorkerState
from .config import B200KVConfig
from .engine import KVEngine, KVGeometry, KVPool, paged_layout_of, xxh64

if TYPE_CHECKING:
    from vllm.config import VllmConfig
    from vllm.forward_context import ForwardContext
    from vllm.v1.core.kv_cache_manager import KVCacheBlocks
    from vllm.v1.core.sched.output import SchedulerOutput
    from vllm.v1.kv_cache_interface import KVCacheConfig
    from vllm.v1.request import Request

logger = logging.getLogger("b200kv")


@dataclass
class B200KVConnectorMetadata(KVConnectorMetadata):
    requests: list[ReqMeta] = field(default_factory=list)


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


def pool_name_for(vllm_config, cfg: B200KVConfig) -> str:
    """One segment per engine unless B200KV_POOL_NAME names a shared one (BASELINE.json
    config 3: "shared pinned-host KV pool" across the replicas of a box)."""
    if cfg.pool_name:
        return cfg.pool_name if cfg.pool_name.startswith("/") else "/" + cfg.pool_name
    eid = vllm_config.kv_transfer_config.engine_id or "engine"
    return "/b200kv-" + "".join(ch for ch in str(eid) if ch.isalnum() or ch in "-_")[:48]


class B200KVConnector(KVConnectorBase_V1, SupportsHMA):
    def __init__(self, vllm_config: "VllmConfig", role: KVConnectorRole,
                 kv_cache_config: "KVCacheConfig | None" = None):
        super().__init__(vllm_config=vllm_config, role=role, kv_cache_config=kv_cache_config)
        ktc = vllm_config.kv_transfer_config
        self.cfg = B200KVConfig.from_env().apply_extra(ktc.kv_connector_extra_config)
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
        self._pool_name = pool_name_for(vllm_config, self.cfg)
        self._pool = KVPool(self._pool_name, self.cfg.pool_bytes, geom.chunk_bytes, _lib.POOL_CREATE_OR_ATTACH)
        self._engine: KVEngine | None = None
        self._worker: WorkerState | None = None
        self._sched: SchedulerState | None = None
        self._controller = None
        if role == KVConnectorRole.SCHEDULER:
            seed = self._key_seed(0)
            lease = self.cfg.lookup_lease_ms
            chunk = self._chunk
            pool = self._pool
            include_partial = not self._discard_partial

            def lookup(token_ids):
                return pool.lookup_tokens(token_ids, chunk, seed, lease, include_partial)

            self._sched = SchedulerState(lookup, self._block_size, self._chunk, self._discard_partial,
                                         self.cfg.save_decode_cache, self.kv_role)
        logger.info("b200kv connector role=%s kv_role=%s pool=%s (%.1f GB, chunk %d, fmt %s)",
                    role.name, self.kv_role, self._pool_name, self.cfg.max_local_cpu_size_gb, self._chunk,
                    "fp8" if self.cfg.fmt else "raw")

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
        self._engine = KVEngine(geom, self._pool, device=t0.device.index or 0,
                                staging_bytes=self.cfg.staging_mb << 20, owner=owner_tag_of(self.cfg.instance_id),
                                variant=self.cfg.variant, key_seed=self._key_seed(rank))
        self._engine.register_kv_caches(tensors)
        self._worker = WorkerState(self._engine, self._block_size, self._chunk, self.kv_role)
        if self.cfg.enable_controller and self.cfg.controller_pull_url and rank == 0:
            from .controller_client import ControllerClient
            self._controller = ControllerClient(
                self.cfg.controller_pull_url, self.cfg.instance_id, self._pool_name, self._engine.key_seed,
                self._chunk, owner_tag=owner_tag_of(self.cfg.instance_id) if self.cfg.pool_name else 0,
                include_partial=not self._discard_partial, heartbeat_s=self.cfg.worker_heartbeat_s)
        logger.info("b200kv registered %d layers, %d blocks, stride %d", len(tensors), nb, stride)

    def _metas(self) -> list[ReqMeta]:
        md = self._get_connector_metadata()
        assert isinstance(md, B200KVConnectorMetadata)
        return md.requests

    def start_load_kv(self, forward_context: "ForwardContext", **kwargs: Any) -> None:
        if self._worker is None:
            return
        self._worker.start_load(self._metas(), stream=torch.cuda.current_stream())

    def wait_for_layer_load(self, layer_name: str) -> None:
        return  # loads are ordered before the forward pass on the compute stream (stream wait)

    def save_kv_layer(self, layer_name: str, kv_layer: torch.Tensor, attn_metadata, **kwargs: Any) -> None:
        return  # whole-request store in wait_for_save: CUDA-graph replay skips per-layer hooks (base.py:591-611)

    def wait_for_save(self):
        if self._worker is None:
            return
        self._worker.save(self._metas(), stream=torch.cuda.current_stream())

    def get_finished(self, finished_req_ids: set[str]) -> tuple[set[str] | None, set[str] | None]:
        if self._worker is not None:
            self._worker.reap()
        return None, None

    def get_block_ids_with_load_errors(self) -> set[int]:
        return self._worker.take_load_errors() if self._worker is not None else set()

    def shutdown(self):
        if self._controller is not None:
            self._controller.close()
            self._controller = None
        if self._engine is not None:
            self._engine.wait_all()
            self._engine.close()
            self._engine = None
        if self._pool is not None:
            self._pool.close()
            self._pool = None

    # ------------------------------------------------------------------ scheduler side
    def get_num_new_matched_tokens(self, request: "Request", num_computed_tokens: int) -> tuple[int | None, bool]:
        assert self._sched is not None
        n = self._sched.num_new_matched_tokens(request.request_id, request.prompt_token_ids or [],
                                               request.num_tokens, num_computed_tokens)
        return n, False

    def update_state_after_alloc(self, request: "Request", blocks: "KVCacheBlocks", num_external_tokens: int):
        assert self._sched is not None
        self._sched.after_alloc(request, num_external_tokens)

    def build_connector_meta(self, scheduler_output: "SchedulerOutput") -> KVConnectorMetadata:
        assert self._sched is not None
        return B200KVConnectorMetadata(self._sched.build_meta(scheduler_output))

    def request_finished(self, request: "Request", block_ids: list[int]) -> tuple[bool, dict[str, Any] | None]:
        # the gather that reads a request's pages is ordered before any later forward pass on the
        # compute stream, so blocks may be freed immediately (no delay_free)
        return False, None

    def request_finished_all_groups(self, request: "Request", block_ids: tuple[list[int], ...]):
        return self.request_finished(request, block_ids[0] if block_ids else [])

    def reset_cache(self) -> bool | None:
        return self._pool.clear() if self._pool is not None else None

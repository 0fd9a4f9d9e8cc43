"""LMCacheControllerManager stand-in for the router's kv-aware dispatch.

Contract used by the router (src/vllm_router/routers/routing_logic.py:276-316):
    m = LMCacheControllerManager({"pull": "0.0.0.0:P", "reply": ... , ["heartbeat": ...]},
                                 health_check_interval=5, lmcache_worker_timeout=30)
    await m.start_all()                      # runs for the life of the router (own loop/thread)
    ret = await m.handle_orchestration_message(LookupMsg | QueryInstMsg)

Workers (b200kv connectors with LMCACHE_ENABLE_CONTROLLER) PUSH a RegisterMsg + heartbeats to the
"pull" URL (LMCACHE_CONTROLLER_PULL_URL, helm/templates/deployment-vllm-multi.yaml:359-362).  A
lookup then reads the workers' chunk index directly: on one box every replica's pool is a POSIX
shm segment, so the controller attaches to it (CPU only) instead of mirroring KV admit/evict
events the way LMCache's controller does.
"""
from __future__ import annotations

import asyncio
import json
import logging
import time

import numpy as np

from .message import (HeartbeatMsg, LookupMsg, LookupRetMsg, QueryInstMsg, QueryInstRetMsg, RegisterMsg)

logger = logging.getLogger("b200kv.controller")


class _Worker:
    def __init__(self, reg: RegisterMsg):
        self.reg = reg
        self.last_seen = time.monotonic()
        self.pool = None

    def open_pool(self):
        if self.pool is None:
            from b200kv import KVPool, _lib
            self.pool = KVPool(self.reg.pool_name, 0, 0, _lib.POOL_ATTACH)
        return self.pool


class LMCacheControllerManager:
    def __init__(self, controller_urls: dict, health_check_interval: int = 5, lmcache_worker_timeout: int = 30):
        self.controller_urls = controller_urls
        self.health_check_interval = health_check_interval
        self.worker_timeout = lmcache_worker_timeout
        self.workers: dict[str, _Worker] = {}
        self._stop = False
        self._sock = None
        self._ctx = None

    # ------------------------------------------------------------------ registration (direct)
    def register(self, reg: RegisterMsg):
        """Also callable in-process (tests, single-process deployments)."""
        w = self.workers.get(reg.instance_id)
        if w is None or w.reg != reg:
            self.workers[reg.instance_id] = _Worker(reg)
            logger.info("registered instance %s ip=%s pool=%s", reg.instance_id, reg.ip, reg.pool_name)
        else:
            w.last_seen = time.monotonic()

    def _on_wire(self, raw: bytes):
        try:
            d = json.loads(raw.decode())
            kind = d.pop("type")
            if kind == "register":
                self.register(RegisterMsg(**d))
            elif kind == "heartbeat":
                w = self.workers.get(HeartbeatMsg(**d).instance_id)
                if w is not None:
                    w.last_seen = time.monotonic()
        except Exception as e:  # a malformed message must not kill the router's controller thread
            logger.warning("bad controller message: %r", e)

    # ------------------------------------------------------------------ router-facing API
    async def start_all(self):
        import zmq
        import zmq.asyncio
        self._ctx = zmq.asyncio.Context.instance()
        self._sock = self._ctx.socket(zmq.PULL)
        url = self.controller_urls.get("pull")
        self._sock.bind(url if "://" in url else f"tcp://{url}")
        last_reap = time.monotonic()
        try:
            while not self._stop:
                if await self._sock.poll(timeout=200):
                    self._on_wire(await self._sock.recv())
                now = time.monotonic()
                if now - last_reap > self.health_check_interval:
                    last_reap = now
                    for iid in [i for i, w in self.workers.items() if now - w.last_seen > self.worker_timeout]:
                        logger.warning("instance %s timed out", iid)
                        self.workers.pop(iid)
        finally:
            self._sock.close(0)

    def stop(self):
        self._stop = True

    async def handle_orchestration_message(self, msg):
        if isinstance(msg, LookupMsg):
            return self._lookup(msg)
        if isinstance(msg, QueryInstMsg):
            for iid, w in self.workers.items():
                if w.reg.ip == msg.ip:
                    return QueryInstRetMsg(iid, msg.event_id)
            return QueryInstRetMsg(None, msg.event_id)
        raise TypeError(f"unsupported orchestration message {type(msg).__name__}")

    def _lookup(self, msg: LookupMsg) -> LookupRetMsg:
        from b200kv import chunk_keys
        toks = np.asarray(msg.tokens or [], dtype=np.int32)
        best: tuple[int, str] | None = None
        for iid, w in self.workers.items():
            try:
                pool = w.open_pool()
            except Exception:
                continue
            keys = chunk_keys(toks, w.reg.chunk_tokens, w.reg.key_seed, w.reg.include_partial)
            if len(keys) == 0:
                continue
            hits, owners = pool.lookup_owner(keys)
            if w.reg.owner_tag:  # shared pool: count only the prefix this instance itself stored
                own = 0
                for o in owners:
                    if int(o) != w.reg.owner_tag:
                        break
                    own += 1
                hits = own
            matched = min(hits * w.reg.chunk_tokens, len(toks))
            if matched > 0 and (best is None or matched > best[0]):
                best = (matched, iid)
        layout = {} if best is None else {best[1]: ("LocalCPUBackend", best[0])}
        return LookupRetMsg(layout, msg.event_id)

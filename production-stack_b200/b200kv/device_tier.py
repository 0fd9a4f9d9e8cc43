"""Device chunk tier (BASELINE.json configs[3]: "kv-aware routing with peer-GPU KV pull over NVLink (no
host hop)").

Every engine may keep the chunks it stored most recently in HBM as well — `B200KV_DEVICE_TIER_GB` of
chunk-format slots, LRU — and exports that buffer to the other replicas of the box over CUDA IPC.  A
replica that needs a chunk takes it from the nearest copy:

    its own device tier  (HBM -> HBM scatter)
    a peer's device tier (the same scatter kernels reading the peer mapping: P2P loads over NVSwitch)
    the pinned host pool (PCIe)            -> b200kv.engine / the host path, unchanged
    the cache server                       -> b200kv.remote

Reading a peer's *live paged cache* would be unsafe (vLLM recycles pages at any step); tier slots are
only recycled through the tier's own index, a `b200kv_pool` in shm (`/b200kv-dev-<engine>`, 16-byte
dummy slots: the payload lives in HBM) in which a consumer PINS the chunks it is reading — pins of a
consumer that died expire by age (b200kv_pool.cpp).  Discovery on one box: every worker publishes
`/dev/shm/b200kv-tier-<engine>.json` (index name, geometry, IPC handle); scheduler processes attach to
the indices only (no CUDA), workers also map the buffers.
"""
from __future__ import annotations

import base64
import json
import logging
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import B200KVError
from .engine import KVPool
from .pd import SHM_DIR, _safe

logger = logging.getLogger("b200kv")

INDEX_SLOT = 16   # bytes of a dummy slot in a tier's index pool


def tier_path(engine_id: str) -> str:
    return os.path.join(SHM_DIR, f"b200kv-tier-{_safe(engine_id)}.json")


def index_name(engine_id: str) -> str:
    return "/b200kv-dev-" + _safe(engine_id)[:48]


class LocalTier:
    """This engine's own tier: device buffer + index, and the publication for the peers."""

    def __init__(self, engine, engine_id: str, n_slots: int, device: int, fmt_tag: int, owner: int = 0,
                 event_factory=None):
        self.engine, self.engine_id, self.n_slots = engine, engine_id, n_slots
        self.chunk_bytes = engine.geom.chunk_bytes
        self.fmt_tag, self.owner = fmt_tag, owner
        self.base = engine.tier_create(n_slots)
        KVPool.unlink(index_name(engine_id))      # a stale index of a crashed predecessor
        self.index = KVPool(index_name(engine_id), n_slots * INDEX_SLOT, INDEX_SLOT, _lib.POOL_CREATE)
        self._event_factory = event_factory
        self._pending: list[tuple[object, list[int]]] = []     # (event, keys) waiting for their gather
        doc = {"engine_id": engine_id, "device": device, "pid": os.getpid(), "n_slots": n_slots,
               "chunk_bytes": self.chunk_bytes, "fmt_tag": fmt_tag, "index": index_name(engine_id),
               "desc": base64.b64encode(engine.tier_export()).decode()}
        tmp = tier_path(engine_id) + f".{os.getpid()}.tmp"
        with open(tmp, "w") as f:
            json.dump(doc, f)
        os.replace(tmp, tier_path(engine_id))
        self.stored_chunks = 0

    def _event(self, stream):
        if self._event_factory is not None:
            return self._event_factory(stream)
        import torch
        ev = torch.cuda.Event()
        ev.record(stream if stream is not None else torch.cuda.current_stream())
        return ev

    def put(self, keys, chunk_tokens, slot_mapping, chunk: int, stream=None) -> int:
        """Gather the chunks of `slot_mapping` (chunk-aligned start, one key per chunk) into the tier.
        Chunks already there are skipped; when no slot can be freed the rest is dropped (the host pool
        has them anyway).  They become visible to peers once their gather has completed (poll())."""
        new: list[tuple[int, int]] = []      # (chunk index, slot)
        for i, k in enumerate(keys):
            try:
                new.append((i, self.index.reserve(int(k), int(chunk_tokens[i]), self.fmt_tag, self.owner)))
            except B200KVError as e:
                if e.code == _lib.EEXIST:
                    continue
                break                       # ENOSPC: everything is pinned or being written
        # one launch per run of consecutive new chunks
        s = 0
        while s < len(new):
            e = s
            while e + 1 < len(new) and new[e + 1][0] == new[e][0] + 1:
                e += 1
            c0, c1 = new[s][0], new[e][0] + 1
            sm = np.asarray(slot_mapping[c0 * chunk: min(c1 * chunk, len(slot_mapping))])
            ptrs = [self.base + slot * self.chunk_bytes for _, slot in new[s:e + 1]]
            seg_keys = [int(keys[i]) for i, _ in new[s:e + 1]]
            try:
                self.engine.gather_chunks(sm, ptrs, stream)
                self._pending.append((self._event(stream), seg_keys))
            except Exception as ex:
                logger.error("b200kv device tier: gather failed (%s)", ex)
                for k in seg_keys:
                    self.index.abort(k)
            s = e + 1
        return len(new)

    def poll(self):
        """Commit the chunks whose gather kernel has finished."""
        still = []
        for ev, keys in self._pending:
            if ev.query():
                for k in keys:
                    self.index.commit(k)
                self.stored_chunks += len(keys)
            else:
                still.append((ev, keys))
        self._pending = still

    def close(self):
        try:
            os.unlink(tier_path(self.engine_id))
        except OSError:
            pass
        self.index.clear()
        self.index.close()
        KVPool.unlink(index_name(self.engine_id))


@dataclass
class TierView:
    engine_id: str
    index: KVPool
    base: int | None          # mapped device address (workers), None in a scheduler process
    chunk_bytes: int
    fmt_tag: int
    local: bool


class TierSet:
    """Every tier this process can see: its own engine's first, then the peers' (discovered in SHM_DIR)."""

    def __init__(self, my_engine_id: str, chunk_bytes: int, fmt_tag: int | None, importer=None, refresh_s: float = 2.0):
        """fmt_tag None (scheduler role: the tile layout is a worker-side fact) = do not compare it; the
        key namespace already separates formats and models."""
        self.my_engine_id, self.chunk_bytes, self.fmt_tag = my_engine_id, chunk_bytes, fmt_tag
        self.refresh_s, self._last_refresh = refresh_s, 0.0
        self.importer = importer      # desc bytes -> mapped base (worker role); None = index only
        self.views: dict[str, TierView] = {}
        self._bad: set[str] = set()

    def add_local(self, tier: LocalTier):
        self.views[tier.engine_id] = TierView(tier.engine_id, tier.index, tier.base, tier.chunk_bytes, tier.fmt_tag, True)

    def refresh(self, force: bool = False):
        import time
        now = time.monotonic()
        if not force and now - self._last_refresh < self.refresh_s:
            return
        self._last_refresh = now
        try:
            names = [n for n in os.listdir(SHM_DIR) if n.startswith("b200kv-tier-") and n.endswith(".json")]
        except OSError:
            return
        seen = set()
        for n in names:
            try:
                with open(os.path.join(SHM_DIR, n)) as f:
                    doc = json.load(f)
                eid = doc["engine_id"]
                seen.add(eid)
                if eid in self.views or eid in self._bad:
                    continue
                if doc["chunk_bytes"] != self.chunk_bytes or (self.fmt_tag is not None and doc["fmt_tag"] != self.fmt_tag):
                    self._bad.add(eid)            # another model / format: never compatible
                    continue
                index = KVPool(doc["index"], 0, INDEX_SLOT, _lib.POOL_ATTACH)
                base = None
                if self.importer is not None and eid != self.my_engine_id:
                    base = self.importer(base64.b64decode(doc["desc"]))
                self.views[eid] = TierView(eid, index, base, doc["chunk_bytes"], doc["fmt_tag"], eid == self.my_engine_id)
            except Exception as e:                # torn file, vanished peer, IPC refused: skip it
                logger.debug("b200kv device tier: cannot use %s (%s)", n, e)
        for eid in [e for e, v in self.views.items() if e not in seen and not v.local]:
            self.views.pop(eid).index.close()     # the peer unpublished its tier

    def _ordered(self):
        return sorted(self.views.values(), key=lambda v: not v.local)

    def presence(self, keys, chunk_tokens, lease_ms: int = 0) -> np.ndarray:
        out = np.zeros(len(keys), dtype=bool)
        for v in self._ordered():
            if out.all():
                break
            try:
                out |= v.index.contains(keys, chunk_tokens, lease_ms)
            except B200KVError:
                continue
        return out

    def resolve(self, keys, chunk_tokens) -> list:
        """Per key: (view, slot) of the nearest tier holding it — PINNED there — or None."""
        out: list = [None] * len(keys)
        for v in self._ordered():
            if v.base is None:
                continue
            for i, k in enumerate(keys):
                if out[i] is not None:
                    continue
                try:
                    slot, n_tok, fmt = v.index.acquire(int(k))
                except B200KVError:
                    continue
                if n_tok != int(chunk_tokens[i]) or (self.fmt_tag is not None and fmt != self.fmt_tag):
                    v.index.release(int(k))
                    continue
                out[i] = (v, slot)
        return out

    @staticmethod
    def release(pins):
        for view, key in pins:
            try:
                view.index.release(int(key))
            except B200KVError:
                pass

    def close(self):
        for v in self.views.values():
            if not v.local:
                v.index.close()
        self.views = {}


def combined_prefix_tokens(host_pool: KVPool, tiers: TierSet | None, keys, chunk_tokens, lease_ms: int) -> int:
    """Tokens of the longest prefix of chunks each of which is in the host pool or in some device tier
    (the scheduler's lookup when tiers exist; LMCache's lookup is prefix-shaped too, adapter :1187-1191)."""
    if len(keys) == 0:
        return 0
    present = host_pool.contains(keys, chunk_tokens, lease_ms)
    if tiers is not None and not present.all():
        present |= tiers.presence(keys, chunk_tokens, lease_ms)
    n = int(np.argmin(present)) if not present.all() else len(keys)
    return int(np.sum(np.asarray(chunk_tokens[:n], dtype=np.int64)))

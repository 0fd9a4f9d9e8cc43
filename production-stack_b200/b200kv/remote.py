"""Remote chunk tier: `LMCACHE_REMOTE_URL=lm://host:port` (helm/templates/deployment-vllm-multi.yaml:
338-345) and the cache server the chart deploys beside the engines
(helm/templates/deployment-cache-server.yaml:62-65 — `lmcache_server 0.0.0.0 <port>`).

* `RemoteServer` / `python -m b200kv.server host port` — the server (C++ threads inside libb200kv.so;
  chunks live in a b200kv pool with the usual LRU).
* `RemoteClient` — one connection (ctypes over the C ABI `b200kv_remote_*`).
* `RemoteTier` — what the connector uses:
    - scheduler role: `prefetch_state(req_id, tokens)` — when the local pool holds less than the whole
      prompt, ask the server which further chunks it has and fetch them INTO THE LOCAL PINNED POOL on
      a background thread; the connector answers vLLM "ask me again" (`(None, False)`,
      KVConnectorBase_V1.get_num_new_matched_tokens, base.py:453-486) until that is done, then the
      ordinary local lookup sees the chunks and the ordinary load path moves them to the GPU.
    - worker role: `push(keys)` — after a store, upload the new chunks in the background.
No byte of a chunk passes through Python: sockets <-> pool slots inside the library.
"""
from __future__ import annotations

import ctypes as C
import logging
import queue
import threading
import time
from concurrent.futures import Future, ThreadPoolExecutor

import numpy as np

from . import _lib
from ._lib import B200KVError, PoolStats, check, lib
from .engine import KVPool, chunk_keys

logger = logging.getLogger("b200kv")

EIO = -5


def parse_remote_url(url: str | None) -> tuple[str, int] | None:
    """`lm://host:port` (also `b200kv://`, or bare host:port).  None/"" -> no remote tier."""
    if not url:
        return None
    u = url.strip()
    for scheme in ("lm://", "b200kv://", "tcp://"):
        if u.startswith(scheme):
            u = u[len(scheme):]
            break
    else:
        if "://" in u:
            raise ValueError(f"LMCACHE_REMOTE_URL={url!r}: only lm://host:port is supported")
    u = u.rstrip("/")
    host, sep, port = u.rpartition(":")
    if not sep or not host or not port.isdigit():
        raise ValueError(f"LMCACHE_REMOTE_URL={url!r}: expected lm://host:port")
    return host, int(port)


class RemoteServer:
    """In-process cache server (threads live in the library)."""

    def __init__(self, host: str = "0.0.0.0", port: int = 0, pool_bytes: int = 8 << 30):
        h = C.c_void_p()
        check(lib().b200kv_server_start(host.encode(), port, pool_bytes, C.byref(h)), "b200kv_server_start")
        self._h = h
        self.port = lib().b200kv_server_port(h)

    def stats(self) -> dict:
        a = (C.c_uint64 * 5)()
        check(lib().b200kv_server_get_stats(self._h, a), "b200kv_server_get_stats")
        return dict(zip(("n_put", "n_get", "n_get_miss", "bytes_in", "bytes_out"), (int(x) for x in a)))

    def stop(self):
        if self._h:
            lib().b200kv_server_stop(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.stop()
        except Exception:
            pass


class RemoteClient:
    """One TCP connection to a server; calls are serialised inside the library."""

    def __init__(self, host: str, port: int, timeout_ms: int = 5000):
        h = C.c_void_p()
        check(lib().b200kv_remote_connect(host.encode(), port, timeout_ms, C.byref(h)), "b200kv_remote_connect")
        self._h = h

    def ping(self) -> bool:
        return lib().b200kv_remote_ping(self._h) == 0

    def exists(self, keys) -> int:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        n = C.c_int32(0)
        check(lib().b200kv_remote_exists(self._h, keys.ctypes.data_as(C.POINTER(C.c_uint64)), len(keys), C.byref(n)),
              "b200kv_remote_exists")
        return n.value

    def put(self, pool: KVPool, key: int, owner: int = 0) -> int:
        """0 sent, -EEXIST server already has it, -ENOENT not READY locally, other < 0 = failure."""
        return lib().b200kv_remote_put(self._h, pool.handle, C.c_uint64(int(key)), owner)

    def get(self, pool: KVPool, key: int, owner: int = 0) -> int:
        return lib().b200kv_remote_get(self._h, pool.handle, C.c_uint64(int(key)), owner)

    def stats(self) -> dict:
        st = PoolStats()
        check(lib().b200kv_remote_stats(self._h, C.byref(st)), "b200kv_remote_stats")
        return {f: int(getattr(st, f)) for f, _ in PoolStats._fields_}

    def traffic(self) -> tuple[int, int]:
        up, down = C.c_uint64(0), C.c_uint64(0)
        lib().b200kv_remote_traffic(self._h, C.byref(up), C.byref(down))
        return up.value, down.value

    def close(self):
        if self._h:
            lib().b200kv_remote_close(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class RemoteTier:
    PENDING, DONE = "pending", "done"

    def __init__(self, pool: KVPool, host: str, port: int, chunk: int, key_seed, owner: int = 0,
                 include_partial: bool = True, workers: int = 2, wait_s: float = 2.0, timeout_ms: int = 5000,
                 retry_s: float = 5.0, io_conns: int = 4):
        """key_seed: one seed, or one per tensor-parallel rank (the scheduler prefetches every rank's
        chunks of a prompt; each worker pushes its own)."""
        self.pool, self.host, self.port = pool, host, port
        self.key_seeds = [int(x) for x in (key_seed if isinstance(key_seed, (list, tuple)) else [key_seed])]
        self.chunk, self.owner, self.include_partial = chunk, owner, include_partial
        self.wait_s, self.timeout_ms, self.retry_s = wait_s, timeout_ms, retry_s
        self._tls = threading.local()
        self._exec = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="b200kv-remote")
        # chunk transfers of ONE prefetch run on several connections (a TCP stream moves ~2 GB/s on the
        # build box's loopback and the rate scales with connections: profiles/remote_loopback_r01.json)
        self._io = ThreadPoolExecutor(max_workers=max(1, io_conns), thread_name_prefix="b200kv-remote-io")
        self._io_conns = max(1, io_conns)
        self._jobs: dict[str, tuple[Future, float]] = {}
        self._asked: set[str] = set()                # requests whose one prefetch has been answered
        self._down_until = 0.0                       # server unreachable: do not stall requests on it
        self._push_q: queue.Queue = queue.Queue()
        self._pusher: threading.Thread | None = None
        self._closed = False
        self._inflight = 0                           # chunks queued or being uploaded
        self._inflight_mu = threading.Lock()
        self.fetched_chunks = 0
        self.pushed_chunks = 0
        self.push_skipped = 0
        self.errors = 0

    # ---- connections (one per thread) -------------------------------------------------------------
    def _client(self) -> RemoteClient:
        c = getattr(self._tls, "client", None)
        if c is None:
            c = RemoteClient(self.host, self.port, self.timeout_ms)
            self._tls.client = c
        return c

    def _drop_client(self):
        c = getattr(self._tls, "client", None)
        if c is not None:
            c.close()
            self._tls.client = None

    def _failed(self, what: str, err):
        self.errors += 1
        self._down_until = time.monotonic() + self.retry_s
        self._drop_client()
        logger.warning("b200kv remote tier %s:%d: %s failed (%s); local tier only for %.0f s",
                       self.host, self.port, what, err, self.retry_s)

    # ---- scheduler role ---------------------------------------------------------------------------
    def _get_some(self, keys) -> int:
        c = self._client()
        got = 0
        for k in keys:
            rc = c.get(self.pool, int(k), self.owner)
            if rc == EIO:
                self._drop_client()
                raise B200KVError(rc, "b200kv_remote_get")
            if rc != 0:              # evicted meanwhile / no local room: nothing behind it is useful
                break
            got += 1
        return got

    def _fetch(self, missing: list[np.ndarray]) -> int:
        try:
            c = self._client()
            got = 0
            for keys in missing:
                n = c.exists(keys)
                lanes = min(self._io_conns, n)
                futs = [self._io.submit(self._get_some, keys[i:n:lanes]) for i in range(lanes)]
                got += sum(f.result() for f in futs)
            self.fetched_chunks += got
            return got
        except Exception as e:
            self._failed("prefetch", e)
            return 0

    def prefetch_state(self, req_id: str, tokens) -> str:
        """PENDING while chunks of this prompt are on their way from the server, else DONE.
        Never raises; a dead server degrades to the local tier."""
        job = self._jobs.get(req_id)
        now = time.monotonic()
        if job is not None:
            fut, t0 = job
            if fut.done() or now - t0 > self.wait_s:
                self._jobs.pop(req_id, None)      # late chunks still land in the pool for the next turn
                self._asked.add(req_id)           # one prefetch per request, however often vLLM asks
                return self.DONE
            return self.PENDING
        if self._closed or now < self._down_until or req_id in self._asked:
            return self.DONE
        missing = []
        for seed in self.key_seeds:
            keys = chunk_keys(tokens, self.chunk, seed, self.include_partial)
            n_local = self.pool.lookup_owner(keys)[0] if len(keys) else 0
            if n_local < len(keys):
                missing.append(keys[n_local:].copy())
        if not missing:
            return self.DONE
        self._jobs[req_id] = (self._exec.submit(self._fetch, missing), now)
        return self.PENDING

    def forget(self, req_id: str):
        self._jobs.pop(req_id, None)
        self._asked.discard(req_id)

    # ---- worker role ------------------------------------------------------------------------------
    def push(self, keys):
        """Upload these chunks once they are READY in the local pool (the D2H of a store is still in
        flight when this is called)."""
        if self._closed:
            return
        if self._pusher is None:
            self._pusher = threading.Thread(target=self._push_loop, name="b200kv-remote-push", daemon=True)
            self._pusher.start()
        keys = [int(k) for k in keys]
        with self._inflight_mu:
            self._inflight += len(keys)
        for k in keys:
            self._push_q.put((k, time.monotonic()))

    def _settled(self):
        with self._inflight_mu:
            self._inflight -= 1

    def _push_loop(self):
        while True:
            item = self._push_q.get()
            if item is None:
                return
            key, t0 = item
            if time.monotonic() < self._down_until:
                self.push_skipped += 1
                self._settled()
                continue
            try:
                rc = self._client().put(self.pool, key, self.owner)
                if rc == _lib.ENOENT and time.monotonic() - t0 < 10.0:
                    time.sleep(0.002)                       # not committed yet: try again shortly
                    self._push_q.put((key, t0))
                    continue
                if rc == 0:
                    self.pushed_chunks += 1
                elif rc in (_lib.ENOENT, _lib.EEXIST, _lib.ENOSPC):
                    self.push_skipped += 1
                else:
                    raise B200KVError(rc, "b200kv_remote_put")
            except Exception as e:
                self._failed("push", e)
            self._settled()

    def flush(self, timeout_s: float = 30.0) -> bool:
        """Wait until the push queue is drained (tests, shutdown)."""
        t0 = time.monotonic()
        while time.monotonic() - t0 < timeout_s:
            with self._inflight_mu:
                if self._inflight <= 0:
                    return True
            time.sleep(0.005)
        return False

    def stats(self) -> dict:
        return {"fetched_chunks": self.fetched_chunks, "pushed_chunks": self.pushed_chunks,
                "push_skipped": self.push_skipped, "errors": self.errors}

    def close(self):
        self._closed = True
        if self._pusher is not None:
            self._push_q.put(None)
            self._pusher.join(timeout=5)
            self._pusher = None
        self._exec.shutdown(wait=False, cancel_futures=True)
        self._io.shutdown(wait=False, cancel_futures=True)

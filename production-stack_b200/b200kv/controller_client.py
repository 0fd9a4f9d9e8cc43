"""Worker-side registration with the in-router controller (the compat shim in
production-stack_b200/compat/lmcache/v1/cache_controller).  Mirrors LMCache's worker that
connects to LMCACHE_CONTROLLER_PULL_URL and heartbeats every
LMCACHE_LMCACHE_WORKER_HEARTBEAT_TIME seconds (helm/templates/deployment-vllm-multi.yaml:346-382).
Wire format: JSON over a ZMQ PUSH socket (both ends belong to this repository)."""
from __future__ import annotations

import json
import socket
import threading


def local_ip() -> str:
    try:
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        s.connect(("10.255.255.255", 1))
        ip = s.getsockname()[0]
        s.close()
        return ip
    except Exception:
        return "127.0.0.1"


class ControllerClient:
    def __init__(self, pull_url: str, instance_id: str, pool_name: str, key_seed: int, chunk_tokens: int,
                 owner_tag: int = 0, include_partial: bool = True, heartbeat_s: float = 10.0, ip: str | None = None):
        import zmq
        self._zmq = zmq
        self._ctx = zmq.Context.instance()
        self._sock = self._ctx.socket(zmq.PUSH)
        self._sock.setsockopt(zmq.LINGER, 0)
        self._sock.setsockopt(zmq.SNDHWM, 16)
        self._sock.connect(pull_url if "://" in pull_url else f"tcp://{pull_url}")
        self._reg = {"type": "register", "instance_id": instance_id, "ip": ip or local_ip(), "pool_name": pool_name,
                     "key_seed": int(key_seed), "chunk_tokens": int(chunk_tokens), "owner_tag": int(owner_tag),
                     "include_partial": bool(include_partial)}
        self._hb = {"type": "heartbeat", "instance_id": instance_id}
        self._stop = threading.Event()
        self._period = heartbeat_s
        self._t = threading.Thread(target=self._run, name="b200kv-controller-client", daemon=True)
        self._t.start()

    def _send(self, obj):
        try:
            self._sock.send(json.dumps(obj).encode(), flags=self._zmq.NOBLOCK)
        except self._zmq.Again:
            pass  # controller not up yet: registration is re-sent with every heartbeat

    def _run(self):
        while not self._stop.is_set():
            self._send(self._reg)   # idempotent on the controller; doubles as re-registration
            self._send(self._hb)
            self._stop.wait(self._period)

    def close(self):
        self._stop.set()
        self._t.join(timeout=2)
        self._sock.close(0)

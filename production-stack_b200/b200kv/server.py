"""The cache server the chart deploys as `lmcache_server <host> <port>`
(helm/templates/deployment-cache-server.yaml:62-65), backed by libb200kv.so:

    python -m b200kv.server 0.0.0.0 8080            # serve until SIGTERM/SIGINT
    python -m b200kv.server --probe 127.0.0.1 8080  # liveness probe: exit 0 iff the server answers a PING

Capacity: `B200KV_SERVER_GB` (default 20) of host memory for chunks, LRU-evicted.  The chart's probe
(`health_probe.py` from the LMCache image, deployment-cache-server.yaml:80-85) speaks LMCache's wire
format; with this server set `cacheserverSpec.livenessProbe` to the `--probe` form above.
"""
from __future__ import annotations

import argparse
import logging
import os
import signal
import sys
import threading

from .remote import RemoteClient, RemoteServer

logger = logging.getLogger("b200kv")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="lmcache_server", description=__doc__.split("\n\n")[0])
    ap.add_argument("host")
    ap.add_argument("port", type=int)
    ap.add_argument("device", nargs="?", default="cpu", help="accepted for lmcache_server compatibility; ignored")
    ap.add_argument("--probe", action="store_true", help="liveness probe instead of serving")
    ap.add_argument("--gb", type=float, default=float(os.environ.get("B200KV_SERVER_GB", "20")))
    args = ap.parse_args(argv)
    if args.probe:
        try:
            c = RemoteClient(args.host, args.port, timeout_ms=3000)
            ok = c.ping()
            c.close()
        except Exception as e:
            print(f"probe failed: {e}", file=sys.stderr)
            return 1
        return 0 if ok else 1
    logging.basicConfig(level=os.environ.get("LMCACHE_LOG_LEVEL", "INFO").upper())
    srv = RemoteServer(args.host, args.port, int(args.gb * (1 << 30)))
    logger.info("b200kv cache server listening on %s:%d (%.1f GB)", args.host, srv.port, args.gb)
    print(f"b200kv cache server listening on {args.host}:{srv.port}", flush=True)
    stop = threading.Event()
    for sig in (signal.SIGTERM, signal.SIGINT):
        signal.signal(sig, lambda *_: stop.set())
    while not stop.wait(30.0):
        logger.info("b200kv cache server: %s", srv.stats())
    srv.stop()
    return 0


if __name__ == "__main__":
    sys.exit(main())

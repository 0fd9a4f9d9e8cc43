"""The multi-replica end-to-end runner (tools/e2e/run_multi.py) in its orchestration dry run: mock
OpenAI backends instead of `vllm serve`, the UNMODIFIED reference router in front, the multi-round-QA
driver through it.  Keeps the round-2 runbook executable; no GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_ROUTER = any(os.path.isdir(os.path.join(p, "vllm_router"))
                  for p in ("/root/reference/src", os.path.join(ROOT, "baseline", "_ref")))
pytestmark = pytest.mark.skipif(not HAVE_ROUTER, reason="reference router not present on this machine")


def run(tmp_path, *extra):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "e2e", "run_multi.py"), "--mock", "--num-users", "3",
           "--num-rounds", "2", "--qps", "8", "--shared-system-prompt", "20", "--user-history-prompt", "20",
           "--log-dir", str(tmp_path), *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def test_round_robin_with_cache_server_mode(tmp_path):
    (r,) = run(tmp_path, "--routing", "roundrobin", "--modes", "remote")
    assert r["requests"] == 6 and r["failed"] == 0 and "vllm_router" in os.listdir(r["router"])
    assert "listening" in open(tmp_path / "cache_server_remote.log").read()


def test_orchestrated_prefill_decode_routing(tmp_path):
    (r,) = run(tmp_path, "--routing", "pd", "--replicas", "4", "--modes", "private")
    assert r["requests"] == 6 and r["failed"] == 0
    log = open(tmp_path / "router_private.log").read()
    assert log.count("Prefill endpoint: http://127.0.0.1:810") == 8 + 6  # warm-up + run: every request went P then D
    assert "Decode endpoint: http://127.0.0.1:8102" in log and "Decode endpoint: http://127.0.0.1:8103" in log

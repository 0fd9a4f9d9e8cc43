"""bench.py driver contract, CPU-checkable part: the reference arm runs without a GPU, prints ONE
JSON line with the keys the driver reads, and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ, ORACLE_THREADS="4", BENCH_REF_BLOCKS="1024", BENCH_REF_SESSIONS="2", **(env_extra or {}))   # CI-sized paged cache
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                           "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_json_line():
    p = run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "kv_offload_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 4 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    p = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import pytest
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def test_numa_maps_reader_and_slot_map_helpers():
    """bench.py's own small helpers: the /proc/self/numa_maps reader behind `per_rank[].pages_per_node` and the
    input-table builder that replaced the oracle import in the GPU arm."""
    import mmap
    import ctypes
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    buf = mmap.mmap(-1, 8 << 20)
    buf[:] = b"\x01" * (8 << 20)                      # touch every page
    addr = ctypes.addressof(ctypes.c_char.from_buffer(buf))
    nodes = bench.numa_nodes_of(addr + 4096)
    if os.path.exists("/proc/self/numa_maps"):
        assert nodes and sum(nodes.values()) >= (8 << 20) // 4096 // 2 and all(k.startswith("N") for k in nodes)
    sm = bench.slot_map(np.array([5, 2]), 20)
    assert list(sm[:3]) == [80, 81, 82] and list(sm[16:20]) == [32, 33, 34, 35] and sm.dtype == np.int64

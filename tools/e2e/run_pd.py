"""Disaggregated prefill end to end (BASELINE.json configs[4]: `--prefill 4 --decode 4` on 8 B200s; default
1P + 1D on 2): prefill `vllm serve`s on GPUs 0..P-1 and decode `vllm serve`s on GPUs P..P+D-1, all with
B200KVConnector, requests paired round-robin like the router does (routing_logic.py:619-641), and the
two-step flow the unmodified router performs in `route_orchestrated_disaggregated_request`
(/root/reference/src/vllm_router/services/request_service/request.py:755-908):

  1. POST prefill  {..., max_tokens: 1, stream: false,
                    kv_transfer_params: {do_remote_decode: true, do_remote_prefill: false, ...}}
  2. take `kv_transfer_params` from the prefill response, set remote_host, POST it with the
     original request to the decode engine (streamed).

Measured per request: prefill latency, decode-side TTFT (= KV hand-off over NVLink + 1-token
prefill + first decode step), and whether the decoded text equals what a single engine produces
for the same prompt (greedy).  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import asyncio
import hashlib
import json
import os
import signal
import statistics
import subprocess
import sys
import time

import aiohttp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from run_e2e import wait_ready  # noqa: E402

MODEL = "synth-llama3-8b"     # served model name (tools/e2e/run_scale.py serves under the directory name)


def start_server(gpu: int, port: int, model_dir: str, max_len: int, log_path: str, extra: list[str]):
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = str(gpu)
    env["PYTHONPATH"] = os.path.join(ROOT, "production-stack_b200") + os.pathsep + env.get("PYTHONPATH", "")
    env.update(LMCACHE_LOCAL_CPU="True", LMCACHE_MAX_LOCAL_CPU_SIZE="20", LMCACHE_CHUNK_SIZE="256")
    cfg = {"kv_connector": "B200KVConnector", "kv_connector_module_path": "b200kv.connector", "kv_role": "kv_both",
               "kv_load_failure_policy": "recompute"}
    cmd = [sys.executable, "-m", "vllm.entrypoints.openai.api_server", "--model", model_dir,
           "--served-model-name", "synth-llama3-8b", "--load-format", "dummy", "--dtype", "bfloat16",
           "--max-model-len", str(max_len), "--no-enable-prefix-caching", "--gpu-memory-utilization", "0.8",
           "--port", str(port), "--seed", "0", "--host", "127.0.0.1", "--kv-transfer-config", json.dumps(cfg)] + extra
    log = open(log_path, "w")
    return subprocess.Popen(cmd, env=env, stdout=log, stderr=subprocess.STDOUT, start_new_session=True), log


async def stream_completion(session, url, body):
    t0 = time.time()
    first, text = None, []
    async with session.post(url, json=body) as r:
        r.raise_for_status()
        async for raw in r.content:
            line = raw.decode().strip()
            if not line.startswith("data:") or line[5:].strip() == "[DONE]":
                continue
            obj = json.loads(line[5:])
            for ch in obj.get("choices") or []:
                if first is None:
                    first = time.time()
                text.append(ch.get("text") or "")
    return (first or time.time()) - t0, "".join(text)


async def drive(args, p_urls, d_urls, ref_urls=None):
    """ref_urls: engines WITHOUT a KV connection to the P/D engines for the single-engine reference (when P and D
    share one pinned pool, a reference run on D would put the prompt's KV where P's prefill finds it)."""
    rows = []
    timeout = aiohttp.ClientTimeout(total=600)
    sem = asyncio.Semaphore(max(1, args.concurrency))
    async with aiohttp.ClientSession(timeout=timeout) as s:
        async def one(i):
            async with sem:
                rows.append(await one_request(args, s, i, p_urls[i % len(p_urls)], d_urls[i % len(d_urls)],
                                              ref_urls[i % len(ref_urls)] if ref_urls else None))
        await asyncio.gather(*[one(i) for i in range(args.requests)])
    return rows


async def one_request(args, s, i, p_url, d_url, ref_url=None):
    prompt = f"request {i} " + " ".join(["hi"] * args.prompt_words)
    base = {"model": MODEL, "prompt": prompt, "max_tokens": args.max_tokens, "temperature": 0}
    # reference: one engine does everything (decode engine, no hand-off)
    ref_ttft, ref_text = await stream_completion(s, (ref_url or d_url) + "/v1/completions", dict(base, stream=True))
    # step 1: prefill
    t0 = time.time()
    pre = dict(base, max_tokens=1, stream=False,
               kv_transfer_params={"do_remote_decode": True, "do_remote_prefill": False, "remote_engine_id": None,
                                   "remote_block_ids": None, "remote_host": None, "remote_port": None})
    async with s.post(p_url + "/v1/completions", json=pre) as r:
        r.raise_for_status()
        pdata = await r.json()
    t_prefill = time.time() - t0
    ktp = pdata.get("kv_transfer_params") or {}
    if ktp:
        ktp["remote_host"] = "127.0.0.1"
    # step 2: decode with the hand-off
    d_ttft, d_text = await stream_completion(s, d_url + "/v1/completions",
                                             dict(base, stream=True, kv_transfer_params=ktp))
    return {"prefill_s": t_prefill, "decode_ttft_s": d_ttft, "single_engine_ttft_s": ref_ttft,
            "handoff_params": bool(ktp), "n_remote_blocks": len((ktp.get("remote_block_ids") or [])),
            "same_text": hashlib.sha1(d_text.encode()).hexdigest() == hashlib.sha1(ref_text.encode()).hexdigest(),
            "prompt_tokens": pdata.get("usage", {}).get("prompt_tokens"), "prefill": p_url, "decode": d_url}


def metrics_of(port):
    import urllib.request
    with urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=5) as r:
        txt = r.read().decode()
    out = {}
    for ln in txt.splitlines():
        if ln and not ln.startswith("#") and ("external_kv_transfer" in ln or "kv_cache_usage" in ln or "gpu_cache_usage" in ln):
            out[ln.split(" ")[0]] = float(ln.split(" ")[-1])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--model-dir", default="/tmp/llama3-8b-synth")
    ap.add_argument("--prompt-words", type=int, default=8000)
    ap.add_argument("--max-model-len", type=int, default=8704)
    ap.add_argument("--max-tokens", type=int, default=64)
    ap.add_argument("--requests", type=int, default=6)
    ap.add_argument("--prefill", type=int, default=1, help="number of prefill engines (GPUs 0..P-1)")
    ap.add_argument("--decode", type=int, default=1, help="number of decode engines (GPUs P..P+D-1)")
    ap.add_argument("--concurrency", type=int, default=1, help="requests in flight")
    ap.add_argument("--mock", action="store_true", help="orchestration dry run with tools/mock_backend.py (no GPU)")
    ap.add_argument("--extra", default="")
    ap.add_argument("--log-dir", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    os.makedirs(args.log_dir, exist_ok=True)
    if not args.mock:
        subprocess.run([sys.executable, os.path.join(HERE, "make_model.py"), args.model_dir, "--layers", str(args.layers),
                        "--max-len", str(args.max_model_len)], check=True, stdout=subprocess.DEVNULL)
    extra = args.extra.split() if args.extra else []
    procs = []
    res = {}
    try:
        n = args.prefill + args.decode
        ports = [8021 + i for i in range(n)]
        for i, port in enumerate(ports):
            kind = "prefill" if i < args.prefill else "decode"
            log_path = os.path.join(args.log_dir, f"vllm_pd_{kind}{i}.log")
            if args.mock:
                log = open(log_path, "w")
                procs.append((subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "mock_backend.py"), "--port", str(port),
                                                "--model", "synth-llama3-8b"], stdout=log, stderr=subprocess.STDOUT,
                                               start_new_session=True), log))
            else:
                procs.append(start_server(i, port, args.model_dir, args.max_model_len, log_path, extra))
        ok = all(wait_ready(port, pr[0], 1200) for port, pr in zip(ports, procs))
        p_urls = [f"http://127.0.0.1:{p}" for p in ports[:args.prefill]]
        d_urls = [f"http://127.0.0.1:{p}" for p in ports[args.prefill:]]
        if not ok:
            res["error"] = "servers not ready"
        else:
            rows = asyncio.run(drive(args, p_urls, d_urls))
            time.sleep(2)
            res = {"requests": len(rows), "prompt_tokens": rows[0]["prompt_tokens"],
                   "handoff_params_returned": sum(r["handoff_params"] for r in rows),
                   "remote_blocks": rows[0]["n_remote_blocks"],
                   "prefill_p50_ms": statistics.median(r["prefill_s"] for r in rows) * 1e3,
                   "decode_ttft_p50_ms": statistics.median(r["decode_ttft_s"] for r in rows) * 1e3,
                   "single_engine_ttft_p50_ms": statistics.median(r["single_engine_ttft_s"] for r in rows) * 1e3,
                   "same_text": f"{sum(r['same_text'] for r in rows)}/{len(rows)}",
                   "engines": {"prefill": p_urls, "decode": d_urls}, "concurrency": args.concurrency,
                   "rows": rows, "prefill_metrics": metrics_of(ports[0]), "decode_metrics": metrics_of(ports[args.prefill])}
    finally:
        for proc, log in procs:
            try:
                os.killpg(proc.pid, signal.SIGTERM)   # exactly the groups started above
                proc.wait(timeout=60)
            except Exception:
                try:
                    os.killpg(proc.pid, signal.SIGKILL)
                except Exception:
                    pass
            log.close()
    print(json.dumps(res))
    with open(os.path.join(args.log_dir, "pd_results.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

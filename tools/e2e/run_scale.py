"""The headline metric in ONE box session: multi-round-QA p50 TTFT and tokens/s at 1 / 2 / 4 / 8 replicas
behind the UNMODIFIED reference router and driven by the UNMODIFIED benchmarks/multi-round-qa harness
(BASELINE.json configs[1..4]), with every engine started once.

GPU time is the scarce resource (an 8-GPU box is charged 8x), so instead of restarting `vllm serve` per
experiment this runner boots, per GPU g, two engines that stay up for the whole session:

  none_g  no connector (every turn re-prefills its context)                 port 8100+g on 127.0.0.(g+1)
  kv_g    B200KVConnector, ONE pinned pool for the box (B200KV_POOL_NAME),   port 8200+g on 127.0.0.(g+1)
          device chunk tier on, registered with the router's controller

(each with half of the GPU's memory budget; only one of the two carries traffic at a time) and then only
swaps ROUTERS — cheap — between experiments.  Experiments of one wave run concurrently on disjoint GPUs:

  wave 1   N=1 none | N=1 kv | N=2 none | N=2 kv          (GPUs 0 | 1 | 2-3 | 4-5)     prefixaware, harness
  wave 2   N=4 none | N=4 kv                              (GPUs 0-3 | 4-7)
  wave 3   N=8 none ; wave 4  N=8 kv                                                   configs[2]
  then     N=G none ; N=G kv with --routing-logic session (balanced; the reference tutorial's setting)
  wave 5   N=8 kv, --routing-logic kvaware, /v1/completions driver                     configs[3]
  wave 6   N=8 kv, roundrobin (every turn lands on another replica): device tier on -> peer-HBM pull over
           NVLink; wave 7: the same with the tier switched off at run time -> shared host pool (one PCIe hop)
  wave 8   4 prefill + 4 decode: disaggregated_prefill_orchestrated through the router (non-streamed) and the
           same two-step flow issued directly (streamed: decode-side TTFT), 8K-token prompts          configs[4]

With fewer GPUs (--gpus 2 / 4) the plan shrinks to what fits.  User ids of different experiments never
collide (--init-user-id), so a later experiment cannot hit KV an earlier one stored, except for the shared
system prompt — which is what a warm deployment looks like.

    python tools/e2e/run_scale.py --gpus 8 --seconds 40 --qps-per-replica 12
    python tools/e2e/run_scale.py --gpus 2 --mock          # orchestration dry run, no GPU
"""
from __future__ import annotations

import argparse
import asyncio
import csv
import json
import os
import signal
import statistics
import subprocess
import sys
import threading
import time
import urllib.request

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import mrqa_driver  # noqa: E402
from run_e2e import harness_path, wait_ready  # noqa: E402
from run_multi import killpg, router_path, scrape  # noqa: E402

PKG = os.path.join(ROOT, "production-stack_b200")


def host_of(g: int) -> str:
    return f"127.0.0.{g + 1}"


class Box:
    def __init__(self, args):
        self.a = args
        self.procs: list[tuple[str, subprocess.Popen]] = []
        self.logs = []
        self.pool_name = f"/b200kv-box-{os.getpid()}"
        self.tier_off_file = os.path.join(args.log_dir, "tier.off")
        self.model_name = args.model_dir          # served under its directory name: the kv-aware router loads the tokenizer from it
        self.lock = threading.Lock()

    # ---------------------------------------------------------------- engines
    def engine_cmd(self, kind: str, g: int):
        a = self.a
        port = (8100 if kind == "none" else 8200) + g
        if a.mock:
            return [sys.executable, os.path.join(ROOT, "tools", "mock_backend.py"), "--host", host_of(g), "--port", str(port),
                    "--model", self.model_name], dict(os.environ), port
        env = dict(os.environ)
        env["CUDA_VISIBLE_DEVICES"] = str(g)
        env["PYTHONPATH"] = PKG + os.pathsep + env.get("PYTHONPATH", "")
        cargs = []
        if kind == "kv":
            env.update(LMCACHE_LOCAL_CPU="True", LMCACHE_MAX_LOCAL_CPU_SIZE=str(a.cpu_gb * a.gpus), LMCACHE_CHUNK_SIZE="256",
                       B200KV_FORMAT=a.format, B200KV_POOL_NAME=self.pool_name, B200KV_DEVICE_TIER_GB=str(a.tier_gb),
                       B200KV_TIER_DISABLE_FILE=self.tier_off_file, LMCACHE_LMCACHE_INSTANCE_ID=f"replica-{g}",
                       LMCACHE_ENABLE_CONTROLLER="True", LMCACHE_CONTROLLER_PULL_URL="127.0.0.1:9000",
                       LMCACHE_LMCACHE_WORKER_HEARTBEAT_TIME="2", B200KV_ADVERTISE_IP=host_of(g),
                       B200KV_PD_TRACE=os.path.join(a.log_dir, f"pd_trace_{g}.jsonl"))
            # vLLM's default policy FAILS a request whose load came up short (config/kv_transfer.py:70); the connector's
            # contract is "report the blocks, vLLM recomputes them"
            cfg = {"kv_connector": "B200KVConnector", "kv_connector_module_path": "b200kv.connector", "kv_role": "kv_both",
                   "kv_load_failure_policy": "recompute"}
            cargs = ["--kv-transfer-config", json.dumps(cfg)]
        cmd = [sys.executable, "-m", "vllm.entrypoints.openai.api_server", "--model", a.model_dir,
               "--served-model-name", self.model_name, "--load-format", "dummy", "--dtype", "bfloat16",
               "--max-model-len", str(a.max_model_len), "--no-enable-prefix-caching",
               "--gpu-memory-utilization", str(a.gpu_mem_util), "--port", str(port), "--seed", "0",
               # an explicit KV-cache size: vLLM then skips its free-memory profiling, which two engines starting on
               # one GPU at the same time would disturb for each other (gpu_worker.py:370-388)
               "--kv-cache-memory-bytes", str(int(a.kv_cache_gb * (1 << 30))),
               "--host", host_of(g)] + cargs + (a.extra.split() if a.extra else [])
        return cmd, env, port

    def start_engines(self, kinds=("none", "kv")):
        a = self.a
        started = []
        for g in range(a.gpus):
            for kind in kinds:
                cmd, env, port = self.engine_cmd(kind, g)
                log = open(os.path.join(a.log_dir, f"vllm_{kind}_{g}.log"), "w")
                self.logs.append(log)
                p = subprocess.Popen(cmd, env=env, stdout=log, stderr=subprocess.STDOUT, start_new_session=True)
                self.procs.append((f"{kind}_{g}", p))
                started.append((kind, g, port, p))
            if g == 0 and not a.mock and a.stagger_s > 0:
                time.sleep(a.stagger_s)       # the first pair warms the page cache / compile cache for the rest
        t0 = time.time()
        ok = all(wait_ready(port, p, a.startup_timeout, host_of(g)) for kind, g, port, p in started)
        return ok, time.time() - t0

    # ---------------------------------------------------------------- routers
    def start_router(self, port: int, kind: str, gpus: list[int], routing: str, tag: str, labels: list[str] | None = None):
        rpath = router_path()
        env = dict(os.environ)
        extra = [os.path.join(PKG, "compat"), PKG] if routing == "kvaware" else []
        env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "stubs"), rpath, *extra, env.get("PYTHONPATH", "")])
        env["HF_HUB_OFFLINE"] = "1"
        base = 8100 if kind == "none" else 8200
        backends = ",".join(f"http://{host_of(g)}:{base + g}" for g in gpus)
        logic = "disaggregated_prefill_orchestrated" if routing == "pd" else routing
        cmd = [sys.executable, "-m", "vllm_router.app", "--host", "127.0.0.1", "--port", str(port),
               "--service-discovery", "static", "--static-backends", backends,
               "--static-models", ",".join([self.model_name] * len(gpus)), "--routing-logic", logic]
        if routing in ("session", "kvaware"):
            cmd += ["--session-key", "x-user-id"]
        if routing == "kvaware":
            cmd += ["--lmcache-controller-port", "9000", "--kv-aware-threshold", str(self.a.kv_aware_threshold)]
        if routing == "pd":
            cmd += ["--static-model-labels", ",".join(labels), "--prefill-model-labels", "p", "--decode-model-labels", "d"]
        log = open(os.path.join(self.a.log_dir, f"router_{tag}.log"), "w")
        p = subprocess.Popen(cmd, env=env, stdout=log, stderr=subprocess.STDOUT, start_new_session=True)
        with self.lock:
            self.logs.append(log)
            self.procs.append((f"router_{tag}", p))
        if not wait_ready(port, p, 180):
            raise RuntimeError(f"router {tag} not ready")
        return p

    def stop(self, proc):
        killpg(proc)
        with self.lock:
            self.procs = [(n, p) for n, p in self.procs if p is not proc]

    def shutdown(self):
        for _, p in reversed(self.procs):
            killpg(p)
        for lg in self.logs:
            try:
                lg.close()
            except Exception:
                pass
        import glob
        for f in glob.glob("/dev/shm" + self.pool_name + "*"):
            try:
                os.unlink(f)
            except OSError:
                pass


# -------------------------------------------------------------------------------------------------- traffic
def summarize_rows(rows: list[dict]) -> dict:
    if not rows:
        return {"requests": 0}
    ttft = sorted(float(r["ttft"]) for r in rows)
    later = sorted(float(r["ttft"]) for r in rows if int(float(r["question_id"])) > 1)
    gen = sum(float(r["generation_tokens"]) for r in rows)
    span = max(float(r["finish_time"]) for r in rows) - min(float(r["launch_time"]) for r in rows)
    return {"requests": len(rows), "ttft_p50_ms": statistics.median(ttft) * 1e3, "ttft_mean_ms": statistics.fmean(ttft) * 1e3,
            "ttft_p90_ms": ttft[int(0.9 * (len(ttft) - 1))] * 1e3,
            "ttft_p50_later_turns_ms": statistics.median(later) * 1e3 if later else None,
            "output_tokens_per_s": gen / span if span > 0 else None,
            "prompt_tokens_mean": statistics.fmean(float(r["prompt_tokens"]) for r in rows), "span_s": span}


def run_harness(box: Box, base_url: str, n_rep: int, tag: str, init_uid: int, qps_per_replica: float | None = None,
                seconds: float | None = None) -> dict:
    """The unmodified harness; p50 from the CSV it writes (it prints the mean, multi-round-qa.py:497,526).  One harness
    process is a single asyncio loop: from 8 replicas on, the load is offered by two processes (disjoint user ids,
    half of the users and of the qps each) and their CSVs are merged."""
    a = box.a
    hp = harness_path()
    qps = (qps_per_replica or a.qps_per_replica) * n_rep
    users = max(2, int(round(a.users_per_replica * n_rep)))
    k = max(1, n_rep // 4)
    procs, csvs = [], []
    t0 = time.time()
    for j in range(k):
        out_csv = os.path.join(a.log_dir, f"harness_{tag}_{j}.csv" if k > 1 else f"harness_{tag}.csv")
        csvs.append(out_csv)
        cmd = [sys.executable, hp, "--num-users", str(max(2, users // k)), "--num-rounds", str(a.num_rounds), "--qps", str(qps / k),
               "--shared-system-prompt", str(a.shared_system_prompt), "--user-history-prompt", str(a.user_history_prompt),
               "--answer-len", str(a.answer_len), "--model", box.model_name, "--base-url", base_url,
               "--time", str(int(seconds or a.seconds)), "--request-with-user-id", "--init-user-id", str(init_uid + j * 5000),
               "--output", out_csv]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=a.log_dir))
    tails, rc = [], 0
    for p in procs:
        try:
            out, _ = p.communicate(timeout=(seconds or a.seconds) + 240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        rc = rc or p.returncode
        tails.append((out or "")[-400:])
    rows = []
    for c in csvs:
        if os.path.exists(c):
            rows += list(csv.DictReader(open(c)))
    res = summarize_rows(rows)
    res.update(driver="unmodified multi-round-qa.py" + (f" x{k} processes" if k > 1 else ""), harness_exit=rc,
               wall_s=time.time() - t0, qps_offered=qps, users=users)
    if rc:
        res["harness_tail"] = tails
    return res


def run_driver(box: Box, base_url: str, n_rep: int, init_uid: int, api: str, stream: bool = True, rounds: int | None = None,
               users_per_replica: float | None = None, qps_per_replica: float | None = None, **shape) -> dict:
    a = box.a
    users = max(2, int(round((users_per_replica or a.users_per_replica) * n_rep)))
    d = argparse.Namespace(base_url=base_url, model=box.model_name, api=api, stream=stream, num_users=users,
                           num_rounds=rounds or a.num_rounds, qps=(qps_per_replica or a.qps_per_replica) * n_rep,
                           shared_system_prompt=shape.get("S", a.shared_system_prompt),
                           user_history_prompt=shape.get("U", a.user_history_prompt), answer_len=a.answer_len,
                           init_user_id=init_uid, output=None)
    rows, s = asyncio.run(mrqa_driver.run(d))
    out = {"driver": f"tools/e2e/mrqa_driver.py ({api}, stream={stream})", "requests": s["requests"], "failed": s["failed"],
           "ttft_p50_ms": (s["ttft_p50_s"] or 0) * 1e3, "ttft_p90_ms": (s["ttft_p90_s"] or 0) * 1e3,
           "ttft_p50_later_turns_ms": (s["ttft_p50_later_turns_s"] or 0) * 1e3,
           "ttft_p50_first_turn_ms": (s["ttft_p50_first_turn_s"] or 0) * 1e3,
           "output_tokens_per_s": s["output_tokens_per_s"], "prompt_tokens_mean": s["mean_prompt_tokens"], "wall_s": s["wall_s"],
           "qps_offered": d.qps, "users": users}
    if not stream:
        out["note"] = "non-streamed: ttft_* carry total request latency"
    return out


def kv_counters(box: Box, gpus: list[int]) -> dict:
    tot: dict[str, float] = {}
    for g in gpus:
        for k, v in scrape(8200 + g, ("b200kv", "lmcache", "external_prefix_cache"), host_of(g)).items():
            name = k.split("{")[0]
            tot[name] = tot.get(name, 0.0) + v
    return tot


def delta(a: dict, b: dict) -> dict:
    return {k: b.get(k, 0.0) - a.get(k, 0.0) for k in b if b.get(k, 0.0) != a.get(k, 0.0)}


# -------------------------------------------------------------------------------------------------- experiments
def experiment(box: Box, name: str, kind: str, gpus: list[int], routing: str, rport: int, uid: int, traffic: str = "harness",
               results: list | None = None, **kw):
    res = {"experiment": name, "engines": kind, "replicas": len(gpus), "gpus": gpus, "routing": routing}
    router = None
    try:
        labels = kw.pop("labels", None)
        router = box.start_router(rport, kind, gpus, routing, name, labels)
        if routing == "kvaware":
            time.sleep(5)      # two heartbeats: every replica is registered with the router's controller
        before = kv_counters(box, gpus) if kind == "kv" and not box.a.mock else {}
        base = f"http://127.0.0.1:{rport}/v1"
        if traffic == "harness" and harness_path():
            res.update(run_harness(box, base, len(gpus), name, uid, **kw))
        else:
            res.update(run_driver(box, base, len(gpus), uid, **kw))
        if kind == "kv" and not box.a.mock:
            time.sleep(1.0)
            res["connector_counters"] = delta(before, kv_counters(box, gpus))
        res["router_counters"] = scrape(rport, ("vllm:num_incoming_requests", "vllm:current_qps"))
    except Exception as e:
        res["error"] = repr(e)
    finally:
        if router is not None:
            box.stop(router)
    print(json.dumps(res), flush=True)
    if results is not None:
        results.append(res)
    return res


def wave(box: Box, results: list, specs: list[dict]):
    ths = [threading.Thread(target=experiment, kwargs=dict(box=box, results=results, **s)) for s in specs]
    for t in ths:
        t.start()
    for t in ths:
        t.join()


def pd_direct(box: Box, p_gpus: list[int], d_gpus: list[int], results: list):
    """The router's two-step flow issued directly (streamed, so the decode-side TTFT is visible), plus the hand-off
    breakdown from the engines' own event traces (B200KV_PD_TRACE)."""
    import run_pd
    a = box.a
    args = argparse.Namespace(prompt_words=a.pd_prompt_words, max_tokens=a.answer_len, requests=a.pd_requests,
                              concurrency=len(d_gpus))
    run_pd.MODEL = box.model_name
    p_urls = [f"http://{host_of(g)}:{8200 + g}" for g in p_gpus]
    d_urls = [f"http://{host_of(g)}:{8200 + g}" for g in d_gpus]
    res = {"experiment": "pd_direct", "prefill_gpus": p_gpus, "decode_gpus": d_gpus}
    try:
        t_start = time.time()
        ref_urls = [f"http://{host_of(g)}:{8100 + g}" for g in d_gpus]    # connector-less engines on the decode GPUs
        rows = asyncio.run(run_pd.drive(args, p_urls, d_urls, ref_urls))
        res.update(requests=len(rows), prompt_tokens=rows[0]["prompt_tokens"],
                   handoff_params_returned=sum(r["handoff_params"] for r in rows),
                   prefill_p50_ms=statistics.median(r["prefill_s"] for r in rows) * 1e3,
                   decode_ttft_p50_ms=statistics.median(r["decode_ttft_s"] for r in rows) * 1e3,
                   single_engine_ttft_p50_ms=statistics.median(r["single_engine_ttft_s"] for r in rows) * 1e3,
                   same_text=f"{sum(r['same_text'] for r in rows)}/{len(rows)}", concurrency=args.concurrency)
        # hand-off breakdown per request: decode request sent -> blocks allocated -> pull issued -> pull done -> first token
        ev: dict[str, dict] = {}
        for g in d_gpus:
            path = os.path.join(a.log_dir, f"pd_trace_{g}.jsonl")
            if os.path.exists(path):
                for ln in open(path):
                    e = json.loads(ln)
                    if e["t"] >= t_start and e["event"] in ("decode_alloc", "pull_issued", "pull_done", "pull_failed", "pull_invalidated"):
                        ev.setdefault(e["req"], {})[e["event"]] = e["t"]
        segs = {"alloc_to_pull_issued_ms": [], "pull_issued_to_done_ms": []}
        for e in ev.values():
            if "decode_alloc" in e and "pull_issued" in e:
                segs["alloc_to_pull_issued_ms"].append((e["pull_issued"] - e["decode_alloc"]) * 1e3)
            if "pull_issued" in e and "pull_done" in e:
                segs["pull_issued_to_done_ms"].append((e["pull_done"] - e["pull_issued"]) * 1e3)
        res["handoff_breakdown_p50"] = {k: statistics.median(v) for k, v in segs.items() if v}
        res["pull_events"] = {"requests_traced": len(ev), "failed": sum("pull_failed" in e for e in ev.values()),
                              "invalidated": sum("pull_invalidated" in e for e in ev.values())}
        res["rows"] = rows[:4]
    except Exception as e:
        res["error"] = repr(e)
    print(json.dumps(res), flush=True)
    results.append(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--model-dir", default="/tmp/llama3-8b-synth")
    ap.add_argument("--max-model-len", type=int, default=8704)
    ap.add_argument("--gpu-mem-util", type=float, default=0.42)
    ap.add_argument("--kv-cache-gb", type=float, default=40.0, help="paged KV cache per engine (two engines share a GPU)")
    ap.add_argument("--cpu-gb", type=float, default=30.0, help="pinned pool GB per replica (the box pool is gpus x this)")
    ap.add_argument("--tier-gb", type=float, default=8.0)
    ap.add_argument("--format", default="raw")
    ap.add_argument("--startup-timeout", type=float, default=1200)
    ap.add_argument("--stagger-s", type=float, default=0.0)
    ap.add_argument("--seconds", type=float, default=40.0, help="harness --time per experiment")
    ap.add_argument("--qps-per-replica", type=float, default=12.0)
    ap.add_argument("--users-per-replica", type=float, default=24.0)
    ap.add_argument("--num-rounds", type=int, default=8)
    ap.add_argument("--shared-system-prompt", type=int, default=1024)
    ap.add_argument("--user-history-prompt", type=int, default=3072)
    ap.add_argument("--answer-len", type=int, default=64)
    ap.add_argument("--kv-aware-threshold", type=int, default=2000)
    ap.add_argument("--pd-prompt-words", type=int, default=8000)
    ap.add_argument("--pd-requests", type=int, default=16)
    ap.add_argument("--qps-sweep", default="", help="comma list of per-replica qps for an N=1 none-vs-kv sweep (2+ GPUs)")
    ap.add_argument("--skip", default="", help="comma list of waves to skip: scale,session,kvaware,cross,pd,sweep")
    ap.add_argument("--extra", default="")
    ap.add_argument("--mock", action="store_true")
    ap.add_argument("--log-dir", default=os.path.join(ROOT, "gpurun_out", "scale"))
    a = ap.parse_args()
    a.log_dir = os.path.abspath(a.log_dir)      # the harness runs with another cwd
    os.makedirs(a.log_dir, exist_ok=True)
    skip = set(a.skip.split(",")) if a.skip else set()
    if True:      # also for --mock: the kv-aware router loads the tokenizer from this directory
        subprocess.run([sys.executable, os.path.join(HERE, "make_model.py"), a.model_dir, "--layers", str(a.layers),
                        "--max-len", str(max(a.max_model_len, 8192))], check=True, stdout=subprocess.DEVNULL)
    box = Box(a)
    results: list = []
    meta = {"gpus": a.gpus, "format": a.format, "workload": {k: getattr(a, k) for k in (
        "seconds", "qps_per_replica", "users_per_replica", "num_rounds", "shared_system_prompt", "user_history_prompt",
        "answer_len")}, "router": router_path(), "harness": harness_path()}
    t_all = time.time()
    try:
        ok, startup = box.start_engines()
        meta["engine_startup_s"] = startup
        if not ok:
            meta["error"] = "engines not ready"
            raise SystemExit(1)
        G = a.gpus
        uid = [10000]

        def nu():
            uid[0] += 10000
            return uid[0]

        sweep = [float(x) for x in a.qps_sweep.split(",")] if a.qps_sweep else []

        def sweep_pair(q, g_none, g_kv, port):
            return [dict(name=f"sweep_q{q:g}_none", kind="none", gpus=[g_none], routing="roundrobin", rport=port, uid=nu(),
                         qps_per_replica=q), dict(name=f"sweep_q{q:g}_kv", kind="kv", gpus=[g_kv], routing="roundrobin",
                                                  rport=port + 1, uid=nu(), qps_per_replica=q)]

        if "scale" not in skip:
            # N = 1, 2 side by side, then 4, then 8 — none and kv on disjoint GPUs wherever they fit; on an 8-GPU box the
            # spare GPUs of the first waves carry the N=1 qps sweep
            if G >= 6:
                specs = [dict(name="n1_none", kind="none", gpus=[0], routing="prefixaware", rport=8090, uid=nu()),
                         dict(name="n1_kv", kind="kv", gpus=[1], routing="prefixaware", rport=8091, uid=nu()),
                         dict(name="n2_none", kind="none", gpus=[2, 3], routing="prefixaware", rport=8092, uid=nu()),
                         dict(name="n2_kv", kind="kv", gpus=[4, 5], routing="prefixaware", rport=8093, uid=nu())]
                if G >= 8 and sweep:
                    specs += sweep_pair(sweep.pop(0), 6, 7, 8094)
                wave(box, results, specs)
                if sweep and "sweep" not in skip:      # the remaining sweep points, up to four at a time
                    specs = []
                    for j, q in enumerate(sweep[:G // 2]):
                        specs += sweep_pair(q, 2 * j, 2 * j + 1, 8090 + 2 * j)
                    sweep = sweep[G // 2:]
                    wave(box, results, specs)
            elif G >= 2:
                wave(box, results, [dict(name="n1_none", kind="none", gpus=[0], routing="prefixaware", rport=8090, uid=nu()),
                                    dict(name="n1_kv", kind="kv", gpus=[1], routing="prefixaware", rport=8091, uid=nu())])
                for kind in ("none", "kv"):
                    wave(box, results, [dict(name=f"n2_{kind}", kind=kind, gpus=[0, 1], routing="prefixaware", rport=8092, uid=nu())])
            else:
                for kind in ("none", "kv"):
                    wave(box, results, [dict(name=f"n1_{kind}", kind=kind, gpus=[0], routing="prefixaware", rport=8090, uid=nu())])
            if G >= 8:
                wave(box, results, [dict(name="n4_none", kind="none", gpus=[0, 1, 2, 3], routing="prefixaware", rport=8090, uid=nu()),
                                    dict(name="n4_kv", kind="kv", gpus=[4, 5, 6, 7], routing="prefixaware", rport=8091, uid=nu())])
            elif G >= 4:
                for kind in ("none", "kv"):
                    wave(box, results, [dict(name=f"n4_{kind}", kind=kind, gpus=[0, 1, 2, 3], routing="prefixaware", rport=8090, uid=nu())])
            if G >= 8:
                allg = list(range(8))
                for kind in ("none", "kv"):
                    wave(box, results, [dict(name=f"n8_{kind}", kind=kind, gpus=allg, routing="prefixaware", rport=8090, uid=nu())])
        if "session" not in skip and G >= 2:
            # the routing the reference's own multi-GPU benchmark uses (tutorials/08-benchmark-multi-round-qa-multi-gpu.md:
            # 71-72): consistent hash of x-user-id — balanced, and a conversation stays on its replica.  (The prefix-aware
            # router sends a workload with a shared system prompt to ONE replica: profiles/scale_8gpu_r02.json.)
            for kind in ("none", "kv"):
                wave(box, results, [dict(name=f"n{G}_{kind}_session", kind=kind, gpus=list(range(G)), routing="session",
                                         rport=8090, uid=nu())])
        allg = list(range(G))
        if "sweep" not in skip and G >= 2:
            for q in sweep:      # whatever did not fit beside the scaling waves
                wave(box, results, sweep_pair(q, 0, 1, 8090))
        if "kvaware" not in skip and G >= 2:
            # configs[3]: kv-aware routing through the compat controller; /v1/completions so the router's lookup sees the
            # engine's tokens
            wave(box, results, [dict(name=f"n{G}_kv_kvaware", kind="kv", gpus=allg, routing="kvaware", rport=8090, uid=nu(),
                                     traffic="driver", api="completions", rounds=4, users_per_replica=8)])
        if "cross" not in skip and G >= 2:
            # every turn of a conversation lands on ANOTHER replica: its history comes from the peer's HBM (device tier,
            # NVLink, no host hop) — or, with the tier switched off, from the shared pinned pool (one PCIe hop)
            for name, off in ((f"n{G}_kv_roundrobin_tier", False), (f"n{G}_kv_roundrobin_hostpool", True)):
                if off:
                    open(box.tier_off_file, "w").close()
                kind = "none" if off is None else "kv"
                wave(box, results, [dict(name=name, kind=kind, gpus=allg, routing="roundrobin", rport=8090, uid=nu(),
                                         traffic="driver", api="chat", rounds=4, users_per_replica=8)])
                if off and os.path.exists(box.tier_off_file):
                    os.unlink(box.tier_off_file)
        if "pd" not in skip and G >= 2:
            n_p = G // 2
            p_gpus, d_gpus = list(range(n_p)), list(range(n_p, G))
            pd_direct(box, p_gpus, d_gpus, results)
            wave(box, results, [dict(name=f"pd_{n_p}p{G - n_p}d_router", kind="kv", gpus=allg, routing="pd", rport=8090, uid=nu(),
                                     labels=["p"] * n_p + ["d"] * (G - n_p), traffic="driver", api="completions", stream=False,
                                     rounds=1, users_per_replica=2, qps_per_replica=1.0, S=1000, U=a.pd_prompt_words - 1000)])
    finally:
        meta["total_s"] = time.time() - t_all
        with open(os.path.join(a.log_dir, "scale_results.json"), "w") as f:
            json.dump({"meta": meta, "results": results}, f, indent=1)
        print(json.dumps({"meta": meta}), flush=True)
        box.shutdown()


if __name__ == "__main__":
    main()

"""Kernel microbenchmark in the shape SURVEY.md §8(d) asks for: Llama-3-8B KV (L32 H8 D128 bf16,
block 16, NB=8192 -> 16 GiB of pages), source block table = first n/16 entries of a seeded
permutation, n_tok in {256, 2048, 8192, 32768}, 20 warm-up + 100 timed launches with CUDA events,
median reported.  Every iteration uses ANOTHER window of the permutation, so consecutive launches
never touch the same pages (small sizes would otherwise be timed out of the 126 MB L2).

Rows: RAW / FP8 x NHD / HND x store (gather) / retrieve (scatter) through the C ABI
(`b200kv_gather` / `b200kv_scatter`, device-resident chunk buffer), and — the same-GPU baseline
SURVEY.md §8(d) calls 2(b) — a plain-PyTorch restatement of LMCache's gather/scatter
(`index_select` per layer into the (L, 2, n, H, D) object; `index_copy_` back).

    python tools/microbench.py [--iters 100] [--out gpurun_out/microbench.json]
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "production-stack_b200"))
import b200kv  # noqa: E402
from b200kv import FMT_FP8, FMT_Q4, FMT_RAW, KVEngine, KVGeometry  # noqa: E402

L, H, D, BS, C, NB = 32, 8, 128, 16, 256, 8192
SIZES = (256, 2048, 8192, 32768)
P_BF16 = L * 2 * H * D * 2                       # 131072 B / token
ALGO = {FMT_RAW: 2 * P_BF16, FMT_FP8: P_BF16 + P_BF16 // 2 + 8,      # §8(d): HBM read + HBM write per token
        FMT_Q4: P_BF16 + P_BF16 * 9 // 32}                          # 4 bits + a bf16 scale per 32 elements
FMT_NAME = {FMT_RAW: "raw", FMT_FP8: "fp8", FMT_Q4: "q4"}


def slots_of(blocks: np.ndarray) -> np.ndarray:
    return (blocks.astype(np.int64)[:, None] * BS + np.arange(BS, dtype=np.int64)[None, :]).reshape(-1)


def windows(perm: np.ndarray, n_tok: int, count: int):
    nblk = n_tok // BS
    n_win = max(1, min(count, len(perm) // nblk))
    return [perm[i * nblk:(i + 1) * nblk] for i in range(n_win)]


def timed(fn, n_warm: int, n_iter: int, kernel_ms=None):
    for i in range(n_warm):
        fn(i)
    torch.cuda.synchronize()
    call, kern = [], []
    for i in range(n_iter):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(n_warm + i)
        e1.record()
        e1.synchronize()
        call.append(e0.elapsed_time(e1))
        if kernel_ms is not None:
            kern.append(kernel_ms())
    return statistics.median(call), (statistics.median(kern) if kern else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.json"))
    ap.add_argument("--no-torch-baseline", action="store_true")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        peak_kind = "MEASURED_PEAKS.json"
    except Exception:
        peak, peak_kind = 6650.0, "fallback (B200_PROFILING.md)"
    dev = torch.device("cuda:0")
    perm = torch.randperm(NB, generator=torch.Generator().manual_seed(1234)).numpy()
    dperm = torch.randperm(NB, generator=torch.Generator().manual_seed(4321)).numpy()
    rows = []

    def add(n_tok, fmt, layout, op, call_ms, kern_ms, impl):
        algo = ALGO[fmt] * n_tok
        r = {"n_tok": n_tok, "format": FMT_NAME[fmt], "tile": layout, "op": op, "impl": impl,
             "call_ms": round(call_ms, 4), "kernel_ms": None if kern_ms is None else round(kern_ms, 4),
             "algo_bytes": algo, "call_GBps": round(algo / call_ms / 1e6, 1),
             "kernel_GBps": None if kern_ms is None else round(algo / kern_ms / 1e6, 1)}
        r["frac_of_hbm_peak"] = round((r["kernel_GBps"] or r["call_GBps"]) / peak, 3)
        rows.append(r)
        print(json.dumps(r), flush=True)

    for layout in ("NHD", "HND"):
        g = torch.Generator(device=dev).manual_seed(0)
        if layout == "NHD":
            caches = [torch.randn((2, NB, BS, H, D), generator=g, device=dev, dtype=torch.float32).bfloat16()
                      for _ in range(L)]
            stride = 0
        else:   # vLLM FlashInfer on Blackwell: logical (NB, 2, bs, H, D) over physical (NB, 2, H, bs, D)
            caches = [torch.randn((NB, 2, H, BS, D), generator=g, device=dev, dtype=torch.float32).bfloat16()
                      .permute(0, 1, 3, 2, 4) for _ in range(L)]
            stride = 2 * BS * H * D * 2
        lay = b200kv._lib.LAYOUT_NHD if layout == "NHD" else b200kv._lib.LAYOUT_HND
        for fmt in (FMT_RAW, FMT_FP8, FMT_Q4):
            geom = KVGeometry(L, H, D, NB, BS, C, 2, stride, fmt, lay)
            eng = KVEngine(geom, None, 0, staging_bytes=0)
            eng.register_kv_caches(caches)
            buf = torch.empty((max(SIZES) // C) * geom.chunk_bytes, dtype=torch.uint8, device=dev)
            for n_tok in SIZES:
                src = [slots_of(w) for w in windows(perm, n_tok, 64)]
                dst = [slots_of(w) for w in windows(dperm, n_tok, 64)]
                eng.gather(src[0], buf.data_ptr())
                c, k = timed(lambda i: eng.gather(src[i % len(src)], buf.data_ptr()), args.warmup, args.iters,
                             lambda: eng.last_kernel_ms(0))
                add(n_tok, fmt, layout, "store(gather)", c, k, "b200kv")
                c, k = timed(lambda i: eng.scatter(dst[i % len(dst)], buf.data_ptr()), args.warmup, args.iters,
                             lambda: eng.last_kernel_ms(1))
                add(n_tok, fmt, layout, "retrieve(scatter)", c, k, "b200kv")
            # one engine op per step: 16 requests x 256 tokens handed over together (b200kv_store_batch_async /
            # b200kv_load_batch_async build exactly this table) vs one op per request
            n_req, small = 16, 256
            wins = windows(perm, small, 512)
            dwins = windows(dperm, small, 512)

            def group(ws, i):
                return [slots_of(ws[(i * n_req + j) % len(ws)]) for j in range(n_req)]

            def per_request(i, ws, fn):
                for j, sm in enumerate(group(ws, i)):
                    fn(sm, buf.data_ptr() + j * geom.chunk_bytes)

            for op, ws, fn, which in (("store(gather)", wins, eng.gather, 0), ("retrieve(scatter)", dwins, eng.scatter, 1)):
                c1, _ = timed(lambda i: per_request(i, ws, fn), 5, max(10, args.iters // 4))
                c2, k2 = timed(lambda i: fn(np.concatenate(group(ws, i)), buf.data_ptr()), 5, max(10, args.iters // 4),
                               lambda: eng.last_kernel_ms(which))
                algo = ALGO[fmt] * small * n_req
                r = {"n_tok": small, "requests_per_step": n_req, "format": FMT_NAME[fmt], "tile": layout, "op": op,
                     "impl": "b200kv", "one_op_per_request_call_ms": round(c1, 4), "one_op_per_step_call_ms": round(c2, 4),
                     "one_op_per_step_kernel_ms": round(k2, 4), "per_request_call_us_batched": round(c2 / n_req * 1e3, 2),
                     "per_request_call_us_unbatched": round(c1 / n_req * 1e3, 2),
                     "kernel_GBps_batched": round(algo / k2 / 1e6, 1), "frac_of_hbm_peak": round(algo / k2 / 1e6 / peak, 3)}
                rows.append(r)
                print(json.dumps(r), flush=True)
            eng.close()
            del buf
        if layout == "NHD" and not args.no_torch_baseline:
            # baseline 2(b): LMCache's gather/scatter restated in PyTorch, same pages, same GPU
            for n_tok in SIZES:
                nblk = n_tok // BS
                obj = torch.empty((L, 2, nblk, BS, H, D), dtype=torch.bfloat16, device=dev)
                src = [torch.from_numpy(w.astype(np.int64)).to(dev) for w in windows(perm, n_tok, 64)]
                dst = [torch.from_numpy(w.astype(np.int64)).to(dev) for w in windows(dperm, n_tok, 64)]

                def t_store(i):
                    idx = src[i % len(src)]
                    for l in range(L):
                        torch.index_select(caches[l], 1, idx, out=obj[l])

                def t_load(i):
                    idx = dst[i % len(dst)]
                    for l in range(L):
                        caches[l].index_copy_(1, idx, obj[l])

                c, _ = timed(t_store, args.warmup, args.iters)
                add(n_tok, FMT_RAW, layout, "store(gather)", c, None, "torch index_select x32")
                c, _ = timed(t_load, args.warmup, args.iters)
                add(n_tok, FMT_RAW, layout, "retrieve(scatter)", c, None, "torch index_copy_ x32")
                del obj
        del caches
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"peak_GBps": peak, "peak_kind": peak_kind, "iters": args.iters, "warmup": args.warmup,
                   "geometry": {"L": L, "H": H, "D": D, "block": BS, "chunk": C, "NB": NB}, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()

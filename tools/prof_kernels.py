"""Short kernel-only driver for ncu: RAW gather/scatter and FP8 pack/unpack on the bench's
32768-token wave (device-resident), `reps` launches each.  Usage (under gpurun):
  ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tools/prof_kernels.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "production-stack_b200"))
from b200kv import FMT_FP8, FMT_Q4, FMT_RAW, KVEngine, KVGeometry  # noqa: E402
from oracle import kv_oracle as ko  # noqa: E402  (slot-mapping helper only)

L, H, D, BS, C, NB = 32, 8, 128, 16, 256, 8192
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
variant = int(os.environ.get("PROF_VARIANT", "0"))
hnd = os.environ.get("PROF_HND", "0") == "1"      # tiles as vLLM's FlashInfer backend lays them out on B200
NB = int(os.environ.get("PROF_NB", NB))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
if hnd:
    caches = [torch.randn((NB, 2, H, BS, D), generator=g, device=dev, dtype=torch.float32).bfloat16().permute(0, 1, 3, 2, 4)
              for _ in range(L)]
else:
    caches = [torch.randn((2, NB, BS, H, D), generator=g, device=dev, dtype=torch.float32).bfloat16() for _ in range(L)]
perm = torch.randperm(NB, generator=torch.Generator().manual_seed(1234)).numpy()
dperm = torch.randperm(NB, generator=torch.Generator().manual_seed(4321)).numpy()
sm = ko.slot_mapping_from_blocks(perm[: tokens // BS], BS, tokens)
dm = ko.slot_mapping_from_blocks(dperm[: tokens // BS], BS, tokens)
fmts = {"raw": FMT_RAW, "fp8": FMT_FP8, "q4": FMT_Q4}
for fmt in [fmts[f] for f in os.environ.get("PROF_FORMATS", "raw,fp8").split(",")]:
    geom = KVGeometry(L, H, D, NB, BS, C, 2, 2 * BS * H * D * 2 if hnd else 0, fmt, 1 if hnd else 0)
    eng = KVEngine(geom, None, 0, staging_bytes=0, variant=variant)
    eng.register_kv_caches(caches)
    buf = torch.empty((tokens // C) * geom.chunk_bytes, dtype=torch.uint8, device=dev)
    for _ in range(reps):
        eng.gather(sm, buf.data_ptr())
        eng.scatter(dm, buf.data_ptr())
    torch.cuda.synchronize()
    print("fmt", fmt, "gather ms", eng.last_kernel_ms(0), "scatter ms", eng.last_kernel_ms(1))
    eng.close()

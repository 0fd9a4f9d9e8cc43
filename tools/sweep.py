"""Launch-shape sweep of the RAW bulk-copy kernel (device-resident gather/scatter, 32768 tokens).
Prints one line per configuration: kernel ms from the library's CUDA events (median of reps)."""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "production-stack_b200"))
from b200kv import FMT_RAW, KVEngine, KVGeometry  # noqa: E402
from oracle import kv_oracle as ko  # noqa: E402

L, H, D, BS, C, NB = 32, 8, 128, 16, 256, 8192
tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
caches = [torch.randn((2, NB, BS, H, D), generator=g, device=dev, dtype=torch.float32).bfloat16() for _ in range(L)]
perm = torch.randperm(NB, generator=torch.Generator().manual_seed(1234)).numpy()
dperm = torch.randperm(NB, generator=torch.Generator().manual_seed(4321)).numpy()
sm = ko.slot_mapping_from_blocks(perm[: tokens // BS], BS, tokens)
dm = ko.slot_mapping_from_blocks(dperm[: tokens // BS], BS, tokens)
geom = KVGeometry(L, H, D, NB, BS, C, 2, 0, FMT_RAW)
buf = torch.empty((tokens // C) * geom.chunk_bytes, dtype=torch.uint8, device=dev)
algo = tokens * 262144


def run(variant, env):
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        eng = KVEngine(geom, None, 0, staging_bytes=0, variant=variant)
    except Exception as e:
        print(variant, env, "ERR", e)
        return
    eng.register_kv_caches(caches)
    gs, ss = [], []
    for i in range(6):
        eng.gather(sm, buf.data_ptr())
        eng.scatter(dm, buf.data_ptr())
        torch.cuda.synchronize()
        if i >= 2:
            gs.append(eng.last_kernel_ms(0))
            ss.append(eng.last_kernel_ms(1))
    gm, s_ = float(np.median(gs)), float(np.median(ss))
    print(f"variant={variant} {env} gather {gm:.4f} ms {algo/gm/1e6:.0f} GB/s | scatter {s_:.4f} ms {algo/s_/1e6:.0f} GB/s", flush=True)
    eng.close()


run(1, {})
for (S, LAG), cps, kb in itertools.product([(2, 1), (3, 1), (3, 2), (4, 2), (4, 3), (6, 3), (6, 4)], [1, 2, 3, 4], [32, 16]):
    if (256 + S * kb * 1024) * cps > 227 * 1024:
        continue
    run(0, {"B200KV_STAGES": S, "B200KV_LAG": LAG, "B200KV_CTAS_PER_SM": cps, "B200KV_STAGE_KB": kb})

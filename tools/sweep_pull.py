"""Launch-shape sweep of the peer-pull kernel on 2 GPUs in ONE process (peer access, no IPC):
GPU 0 pulls a 32768-token wave out of GPU 1's pages.  One line per configuration."""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "production-stack_b200")]
from b200kv import FMT_RAW, KVEngine, KVGeometry  # noqa: E402
from oracle import kv_oracle as ko  # noqa: E402

L, H, D, BS, C, NB = 32, 8, 128, 16, 256, 4096
tokens = 32768
g1 = torch.Generator(device="cuda:1").manual_seed(0)
remote = [torch.randn((2, NB, BS, H, D), generator=g1, device="cuda:1", dtype=torch.float32).bfloat16() for _ in range(L)]
local = [torch.zeros((2, NB, BS, H, D), device="cuda:0", dtype=torch.bfloat16) for _ in range(L)]
perm = torch.randperm(NB, generator=torch.Generator().manual_seed(1234)).numpy()
src = ko.slot_mapping_from_blocks(perm[: tokens // BS], BS, tokens)
dst = ko.slot_mapping_from_blocks(perm[::-1][: tokens // BS].copy(), BS, tokens)
geom = KVGeometry(L, H, D, NB, BS, C)
payload = tokens * 131072


def run(variant, env):
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        eng = KVEngine(geom, None, 0, staging_bytes=0, variant=variant)
    except Exception as e:
        print(variant, env, "ERR", e)
        return
    eng.register_kv_caches(local)
    eng.import_peer_ptrs(1, 1, [t[0].data_ptr() for t in remote], [t[1].data_ptr() for t in remote])
    ms = []
    for i in range(5):
        eng.wait(eng.peer_pull(1, src, dst))
        if i >= 2:
            ms.append(eng.last_kernel_ms(2))
    m = float(np.median(ms))
    print(f"variant={variant} {env} pull {m:.3f} ms {payload / m / 1e6:.0f} GB/s", flush=True)
    eng.close()


run(1, {})
for (S, LAG), cps, kb in itertools.product([(2, 1), (3, 2), (4, 2), (4, 3), (6, 3), (6, 4)], [1, 2, 3, 4, 6], [32, 16]):
    if (256 + S * kb * 1024) * cps > 227 * 1024:
        continue
    run(0, {"B200KV_STAGES": S, "B200KV_LAG": LAG, "B200KV_CTAS_PER_SM": cps, "B200KV_STAGE_KB": kb})

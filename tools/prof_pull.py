"""Short single-process 2-GPU driver for an ncu capture of the peer-pull kernel: GPU 0 pulls a
32768-token wave (4 GiB) out of GPU 1's pages, `reps` launches.
  ncu --set full --clock-control none -k regex:kv_bulk_copy -o gpurun_out/prof_pull python tools/prof_pull.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "production-stack_b200")]
from b200kv import KVEngine, KVGeometry  # noqa: E402
from oracle import kv_oracle as ko  # noqa: E402

L, H, D, BS, C, NB = 32, 8, 128, 16, 256, 4096
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tokens = 32768
g1 = torch.Generator(device="cuda:1").manual_seed(0)
remote = [torch.randn((2, NB, BS, H, D), generator=g1, device="cuda:1", dtype=torch.float32).bfloat16() for _ in range(L)]
local = [torch.zeros((2, NB, BS, H, D), device="cuda:0", dtype=torch.bfloat16) for _ in range(L)]
perm = torch.randperm(NB, generator=torch.Generator().manual_seed(1234)).numpy()
src = ko.slot_mapping_from_blocks(perm[: tokens // BS], BS, tokens)
dst = ko.slot_mapping_from_blocks(perm[::-1][: tokens // BS].copy(), BS, tokens)
eng = KVEngine(KVGeometry(L, H, D, NB, BS, C), None, 0, staging_bytes=0)
eng.register_kv_caches(local)
eng.import_peer_ptrs(1, 1, [t[0].data_ptr() for t in remote], [t[1].data_ptr() for t in remote])
for _ in range(reps):
    eng.wait(eng.peer_pull(1, src, dst))
print("pull kernel ms", eng.last_kernel_ms(2), "GB/s", tokens * 131072 / eng.last_kernel_ms(2) / 1e6)
eng.close()

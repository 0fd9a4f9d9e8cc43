"""Single-process 2-GPU driver for an ncu capture of the device-tier read path (BASELINE.json configs[3]): chunks
held in GPU 1's HBM (what a peer replica's device tier holds) are scattered into GPU 0's pages by GPU 0's kernel,
reading over NVLink — no host hop.  `reps` launches of a 32768-token wave.
  ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -k regex:kv_ python tools/prof_tier.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "production-stack_b200")]
from b200kv import FMT_FP8, FMT_RAW, KVEngine, KVGeometry  # noqa: E402

L, H, D, BS, C, NB = 32, 8, 128, 16, 256, 4096
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fmt = FMT_FP8 if os.environ.get("PROF_FORMAT", "raw") == "fp8" else FMT_RAW
tokens = 32768
g1 = torch.Generator(device="cuda:1").manual_seed(0)
owner_pages = [torch.randn((2, NB, BS, H, D), generator=g1, device="cuda:1", dtype=torch.float32).bfloat16() for _ in range(L)]
local_pages = [torch.zeros((2, NB, BS, H, D), device="cuda:0", dtype=torch.bfloat16) for _ in range(L)]
perm = torch.randperm(NB, generator=torch.Generator().manual_seed(1234)).numpy()


def slots(blocks):
    return (np.asarray(blocks, dtype=np.int64)[:, None] * BS + np.arange(BS, dtype=np.int64)[None, :]).reshape(-1)


src, dst = slots(perm[: tokens // BS]), slots(perm[::-1][: tokens // BS].copy())
geom = KVGeometry(L, H, D, NB, BS, C, 2, 0, fmt)
owner = KVEngine(geom, None, 1, staging_bytes=0)
owner.register_kv_caches(owner_pages)
n_chunks = tokens // C
tier = torch.empty(n_chunks * geom.chunk_bytes, dtype=torch.uint8, device="cuda:1")        # the owner's device tier
owner.gather(src, tier.data_ptr(), stream=torch.cuda.current_stream(torch.device("cuda:1")))
torch.cuda.synchronize(1)
probe = torch.empty(16, dtype=torch.uint8, device="cuda:0")
probe.copy_(tier[:16])                       # torch enables peer access 0 <-> 1 on first cross-device copy
cons = KVEngine(geom, None, 0, staging_bytes=0)
cons.register_kv_caches(local_pages)
# in-process stand-in for the CUDA-IPC import a real consumer does: maps GPU 1 into GPU 0's context (peer access)
cons.import_peer_ptrs(1, 1, [t[0].data_ptr() for t in owner_pages], [t[1].data_ptr() for t in owner_pages])
ptrs = np.array([tier.data_ptr() + i * geom.chunk_bytes for i in range(n_chunks)], dtype=np.uint64)
torch.cuda.set_device(0)
for _ in range(reps):
    cons.scatter_chunks(dst, ptrs)
    torch.cuda.synchronize(0)
ms = cons.last_kernel_ms(1)
payload = tokens * geom.payload_bytes_per_token
print("tier scatter kernel ms", ms, "NVLink GB/s", payload / ms / 1e6)
# correctness: the consumer's pages now hold what the owner's pages held (RAW: bit for bit)
if fmt == FMT_RAW:
    si, di = torch.from_numpy(src).cuda(0), torch.from_numpy(dst).cuda(0)
    ok = all(torch.equal(a.view(2, NB * BS, H * D)[:, di].view(torch.int16),
                         b.to("cuda:0").view(2, NB * BS, H * D)[:, si].view(torch.int16))
             for a, b in zip(local_pages[:4], owner_pages[:4]))
    print("bit-exact vs owner pages (4 layers checked):", ok)
owner.close()
cons.close()

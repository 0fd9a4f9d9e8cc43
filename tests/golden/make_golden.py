"""Generates the committed fixtures in tests/golden/ (run once in the build container):

    python tests/golden/make_golden.py

Sources of truth (none of them is this repository's code):
* xxh64_vectors.json  — the `xxhash` wheel (what /root/reference/src/vllm_router/prefix/hashtrie.py
  :56-57 calls).
* e4m3_vectors.npz    — `ml_dtypes.float8_e4m3fn` casts, cross-checked with torch.float8_e4m3fn.
* paged_gather.npz    — torch evaluation of vLLM's slot-mapping semantics, written exactly as in
  vllm/distributed/kv_transfer/kv_connector/v1/example_connector.py:247-248 (extract) and
  :154-159 (inject); that file is the executable spec the LMCache adapter's gather/scatter obeys.
The lmcache wheel itself is absent from this image, so KV-byte parity with it stays unpinned.
"""
import json
import os

import ml_dtypes
import numpy as np
import torch
import xxhash

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20260921)

# ---- xxh64 ------------------------------------------------------------------------------------
vec = []
for n in list(range(0, 40)) + [63, 64, 65, 127, 128, 255, 256, 1000, 1024]:
    data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    for seed in (0, 1, 123, 0xDEADBEEFCAFEF00D):
        vec.append({"hex": data.hex(), "seed": seed, "digest": xxhash.xxh64(data, seed=seed).intdigest()})
json.dump(vec, open(os.path.join(HERE, "xxh64_vectors.json"), "w"))

# ---- e4m3 -------------------------------------------------------------------------------------
# every bf16-representable magnitude class + exact ties between adjacent e4m3 values
dec = np.arange(256, dtype=np.uint8).view(ml_dtypes.float8_e4m3fn).astype(np.float32)
fin = np.sort(dec[np.isfinite(dec) & (dec >= 0)])
ties = (fin[:-1] + fin[1:]) / 2
near = np.concatenate([np.nextafter(ties, np.float32(0)).astype(np.float32), ties.astype(np.float32),
                       np.nextafter(ties, np.float32(1e9)).astype(np.float32)])
rand = (rng.standard_normal(4096) * np.exp(rng.uniform(-12, 6, 4096))).astype(np.float32)
x = np.concatenate([fin, near, rand, np.float32([0.0, 448.0, 2.0 ** -9, 2.0 ** -10, 2.0 ** -11, 1e-30])])
x = np.concatenate([x, -x]).astype(np.float32)
x = x[np.abs(x) <= 448.0]  # ml_dtypes overflows to NaN above 448; satfinite is tested separately
codes = x.astype(ml_dtypes.float8_e4m3fn).view(np.uint8)
tcodes = torch.from_numpy(x).to(torch.float8_e4m3fn).view(torch.uint8).numpy()
assert np.array_equal(codes, tcodes), "ml_dtypes and torch disagree on e4m3 rounding"
np.savez_compressed(os.path.join(HERE, "e4m3_vectors.npz"), x=x, codes=codes, decode=dec)

# ---- paged gather / scatter ---------------------------------------------------------------------
L, NB, bs, H, D = 3, 12, 4, 2, 8
g = torch.Generator().manual_seed(0)
layers = [torch.randint(-2 ** 15, 2 ** 15 - 1, (2, NB, bs, H, D), generator=g, dtype=torch.int16) for _ in range(L)]
block_ids = torch.randperm(NB, generator=g)[:7]
n_tok = 7 * bs - 3  # ragged tail
slot_mapping = (torch.arange(bs)[None, :] + block_ids[:, None] * bs).flatten()[:n_tok]  # adapter :368-375
gathered = torch.stack([l.reshape(2, NB * bs, -1)[:, slot_mapping, ...] for l in layers])  # (L,2,n,H*D)
dst_blocks = torch.randperm(NB, generator=g)[:7]
dst_map = (torch.arange(bs)[None, :] + dst_blocks[:, None] * bs).flatten()[:n_tok]
dst_layers = [torch.full_like(l, -1) for l in layers]
for l, d in enumerate(dst_layers):
    dv = d.reshape(2, NB * bs, -1)
    dv[:, dst_map, ...] = gathered[l]
np.savez_compressed(os.path.join(HERE, "paged_gather.npz"),
                    layers=np.stack([l.numpy() for l in layers]), slot_mapping=slot_mapping.numpy(),
                    gathered=gathered.numpy().reshape(L, 2, n_tok, H, D), dst_map=dst_map.numpy(),
                    scattered=np.stack([d.numpy() for d in dst_layers]), block_ids=block_ids.numpy())
print("golden fixtures written to", HERE)

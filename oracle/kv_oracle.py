"""CPU oracle for the KV offload / transfer hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``production-stack_b200``)
never does: it fails loudly when its CUDA library is missing.

PARITY STATUS: **parity unpinned for KV bytes** — the algorithm lives in the third-party
``lmcache`` wheel (pin ``lmcache==0.3.11``, /root/reference/pyproject.toml:49-52) which is neither
vendored under /root/reference (.SUBMODULES.json:8 is empty) nor installable offline, and the
reference's own tests never assert a KV byte, chunk key or hit count (SURVEY.md §4, §8c:
tests/e2e/test-routing.py:471-475 returns True unconditionally).  What IS pinned, by
``tests/golden/`` fixtures generated with ``tests/golden/make_golden.py``:

* gather / scatter indexing  — vLLM's executable spec of the slot-mapping semantics
  (``layer.reshape(2, NB*bs, -1)[:, slot_mapping]``, vllm/.../v1/example_connector.py:247-248 and
  the inverse at :154-159), evaluated with torch;
* XXH64                      — the ``xxhash`` wheel the reference's prefix router uses
  (/root/reference/src/vllm_router/prefix/hashtrie.py:56-57);
* e4m3 rounding              — ``ml_dtypes.float8_e4m3fn`` / ``torch.float8_e4m3fn`` casts;
* save / load planning and the engine calling convention (``plan_save``, ``num_new_matched_tokens``
  here; ``b200kv.adapter`` in the product) — vLLM's vendored LMCache adapter EXECUTED
  (``tests/golden/make_adapter_golden.py`` -> ``adapter_plan_vectors.json``, ``adapter_flow_vectors.json``).

Each function cites the file:line whose behaviour it restates.  [vllm-0.22] means the vLLM wheel
in this image (its vendored LMCache adapter is the only executable description of how the
reference drives the engine).
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1

# ---------------------------------------------------------------------------------------------
# XXH64 (published algorithm), pure Python — small inputs only
# ---------------------------------------------------------------------------------------------
_P1 = 11400714785074694791
_P2 = 14029467366897019727
_P3 = 1609587929392839161
_P4 = 9650029242287828579
_P5 = 2870177450012600261


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & MASK64


def _round(acc: int, inp: int) -> int:
    acc = (acc + inp * _P2) & MASK64
    return (_rotl(acc, 31) * _P1) & MASK64


def _merge(h: int, v: int) -> int:
    h ^= _round(0, v)
    return (h * _P1 + _P4) & MASK64


def xxh64(data: bytes, seed: int = 0) -> int:
    """XXH64 digest; same function as xxhash.xxh64(...).intdigest() used by
    /root/reference/src/vllm_router/prefix/hashtrie.py:56-57."""
    n = len(data)
    p = 0
    if n >= 32:
        v1 = (seed + _P1 + _P2) & MASK64
        v2 = (seed + _P2) & MASK64
        v3 = seed & MASK64
        v4 = (seed - _P1) & MASK64
        while p <= n - 32:
            v1 = _round(v1, int.from_bytes(data[p:p + 8], "little"))
            v2 = _round(v2, int.from_bytes(data[p + 8:p + 16], "little"))
            v3 = _round(v3, int.from_bytes(data[p + 16:p + 24], "little"))
            v4 = _round(v4, int.from_bytes(data[p + 24:p + 32], "little"))
            p += 32
        h = (_rotl(v1, 1) + _rotl(v2, 7) + _rotl(v3, 12) + _rotl(v4, 18)) & MASK64
        for v in (v1, v2, v3, v4):
            h = _merge(h, v)
    else:
        h = (seed + _P5) & MASK64
    h = (h + n) & MASK64
    while p + 8 <= n:
        h ^= _round(0, int.from_bytes(data[p:p + 8], "little"))
        h = (_rotl(h, 27) * _P1 + _P4) & MASK64
        p += 8
    if p + 4 <= n:
        h ^= (int.from_bytes(data[p:p + 4], "little") * _P1) & MASK64
        h = (_rotl(h, 23) * _P2 + _P3) & MASK64
        p += 4
    while p < n:
        h ^= (data[p] * _P5) & MASK64
        h = (_rotl(h, 11) * _P1) & MASK64
        p += 1
    h ^= h >> 33
    h = (h * _P2) & MASK64
    h ^= h >> 29
    h = (h * _P3) & MASK64
    h ^= h >> 32
    return h


def chunk_keys(tokens, chunk_tokens: int = 256, seed: int = 0, include_partial: bool = True):
    """Prefix-chained chunk keys.  LMCache hashes each `chunk_size`-token chunk chained on the
    previous chunk's hash so a key identifies the whole prefix ([vllm-0.22]
    lmcache_integration/vllm_v1_adapter.py:1187-1191 lookup; :329-333 partial-chunk rule).  This
    build defines key[i] = XXH64(int32-LE bytes of chunk i, seed=key[i-1]), key[-1] = seed."""
    toks = np.asarray(tokens, dtype=np.int32)
    keys = []
    prev = seed & MASK64
    n = len(toks)
    full = n // chunk_tokens
    for i in range(full):
        prev = xxh64(toks[i * chunk_tokens:(i + 1) * chunk_tokens].tobytes(), prev)
        keys.append(prev)
    if include_partial and n % chunk_tokens:
        prev = xxh64(toks[full * chunk_tokens:].tobytes(), prev)
        keys.append(prev)
    return keys


# ---------------------------------------------------------------------------------------------
# slot mapping / gather / scatter
# ---------------------------------------------------------------------------------------------
def slot_mapping_from_blocks(block_ids, block_size: int, n_tokens: int) -> np.ndarray:
    """slot_mapping[i] = block_ids[i // bs] * bs + i % bs
    ([vllm-0.22] lmcache_integration/vllm_v1_adapter.py:368-375)."""
    b = np.asarray(block_ids, dtype=np.int64)
    sm = (b[:, None] * block_size + np.arange(block_size, dtype=np.int64)[None, :]).reshape(-1)
    return sm[:n_tokens]


def gather_tokens(kv_layers, slot_mapping: np.ndarray) -> np.ndarray:
    """Paged cache -> contiguous (L, 2, n_tok, H, D).  kv_layers: list of L arrays shaped
    (2, NB, bs, H, D) (FlashAttention layout, [vllm-0.22] v1/attention/backends/flash_attn.py:
    140-149).  Per layer: layer.reshape(2, NB*bs, -1)[:, slot_mapping]
    ([vllm-0.22] v1/example_connector.py:247-248); stacked into LMCache's kv_shape
    (L, 2, chunk, H, D) ([vllm-0.22] lmcache_integration/vllm_v1_adapter.py:471-477)."""
    out = []
    for layer in kv_layers:
        two, nb, bs, h, d = layer.shape
        flat = layer.reshape(2, nb * bs, h, d)
        out.append(flat[:, slot_mapping])
    return np.stack(out, axis=0)


def scatter_tokens(kv_layers, chunk: np.ndarray, slot_mapping: np.ndarray) -> None:
    """Inverse of gather_tokens, in place: dst.reshape(2, NB*bs, -1)[:, slot_mapping] = src
    ([vllm-0.22] v1/example_connector.py:154-159)."""
    for l, layer in enumerate(kv_layers):
        two, nb, bs, h, d = layer.shape
        flat = layer.reshape(2, nb * bs, h, d)  # view: layers are C-contiguous
        assert np.shares_memory(flat, layer)
        flat[:, slot_mapping] = chunk[l]


# ---------------------------------------------------------------------------------------------
# fp8 (e4m3fn) chunk codec
# ---------------------------------------------------------------------------------------------
def bf16_bits_to_f32(u16: np.ndarray) -> np.ndarray:
    return (u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits_rn(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 round-to-nearest-even (what __float2bfloat16_rn does for finite values)."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = (b + 0x7FFF + ((b >> 16) & 1)) >> 16
    return r.astype(np.uint16)


def f32_to_e4m3_satfinite(x: np.ndarray) -> np.ndarray:
    """fp32 -> e4m3fn byte, round-to-nearest-even, saturate-to-finite (PTX
    cvt.rn.satfinite.e4m3x2.f32).  e4m3fn: 1-4-3, bias 7, max finite 448 (0x7E), 0x7F = NaN,
    subnormal quantum 2^-9."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    bits = x.view(np.uint32)
    sign = ((bits >> 24) & 0x80).astype(np.uint8)
    a = np.abs(x).astype(np.float64)
    nan = np.isnan(a)
    a = np.where(nan, 0.0, np.minimum(a, 448.0))
    m, ex = np.frexp(a)  # a = m * 2**ex, m in [0.5, 1)
    e = ex - 1
    normal = a >= 2.0 ** -6
    # normal: mantissa grid 2**(e-3)
    q = np.rint(np.ldexp(a, 3 - e))  # in [8, 16]
    bump = q == 16
    e_n = np.where(bump, e + 1, e)
    q_n = np.where(bump, 8, q)
    code_n = ((e_n + 7).astype(np.int64) << 3) | (q_n.astype(np.int64) - 8)
    # subnormal: grid 2**-9; rint==8 lands on the smallest normal (code 0x08) by construction
    code_s = np.rint(a * 512.0).astype(np.int64)
    code = np.where(normal, code_n, code_s).astype(np.uint8)
    code = np.where(nan, np.uint8(0x7F), code)
    return (code | sign).astype(np.uint8)


def e4m3_decode_table() -> np.ndarray:
    t = np.zeros(256, dtype=np.float32)
    for c in range(256):
        s = -1.0 if c & 0x80 else 1.0
        e = (c >> 3) & 0xF
        m = c & 7
        if e == 0xF and m == 7:
            v = np.nan
        elif e == 0:
            v = m * 2.0 ** -9
        else:
            v = (8 + m) * 2.0 ** (e - 10)
        t[c] = s * v
    return t


_E4M3_DEC = e4m3_decode_table()


def fp8_pack_chunk(chunk_bits: np.ndarray):
    """(L, 2, n, H, D) bf16 bit patterns (uint16) -> (codes uint8 same shape, scales f32 (L,2,H)).
    One scale per (chunk, layer, K/V, head): scale = absmax/448, codes = e4m3(x * (448/absmax))
    (north_star "per-chunk scale").  absmax == 0 -> scale 1."""
    L, two, n, H, D = chunk_bits.shape
    mag = (chunk_bits & 0x7FFF).astype(np.uint16)
    amax_bits = mag.max(axis=(2, 4)) if n else np.zeros((L, two, H), np.uint16)  # (L, 2, H)
    amax = bf16_bits_to_f32(amax_bits)
    zero = amax_bits == 0
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(zero, np.float32(1.0), np.float32(448.0) / amax).astype(np.float32)
        scale = np.where(zero, np.float32(1.0), amax / np.float32(448.0)).astype(np.float32)
    x = bf16_bits_to_f32(chunk_bits)
    y = (x * inv[:, :, None, :, None]).astype(np.float32)
    return f32_to_e4m3_satfinite(y), scale


def fp8_unpack_chunk(codes: np.ndarray, scales: np.ndarray) -> np.ndarray:
    """Inverse: bf16_rn(float(e4m3) * scale) as uint16 bit patterns."""
    x = _E4M3_DEC[codes] * scales[:, :, None, :, None].astype(np.float32)
    return f32_to_bf16_bits_rn(x.astype(np.float32))


def fp8_tolerance(x: np.ndarray, scale_per_elem: np.ndarray) -> np.ndarray:
    """Stated tolerance of the fp8 path:
    |x - x̂| <= max(2^-4 |x|, 2^-10 * absmax) + 2^-8 |x|   with absmax = 448 * scale."""
    ax = np.abs(x)
    return np.maximum(ax * 2.0 ** -4, 2.0 ** -10 * 448.0 * scale_per_elem) + ax * 2.0 ** -8


# ---------------------------------------------------------------------------------------------
# q4 group-wise chunk codec (SURVEY.md §8f-4 "sub-8-bit codec"; north_star "optional CacheGen
# token-chunk quantisation").  SPECIFIED HERE FIRST; the CUDA kernels (kv_q4_store/load_kernel) follow it.
#   * group = Q4_GROUP consecutive elements along D of one (token, head)
#   * scale = bf16_rn(absmax / 7); inv = fl32(1 / scale); codes = rint(x * inv) (exact product, ties to even; within +-7) as two's-complement nibbles,
#     element 2i in the low nibble of byte i; absmax == 0 -> scale 1, codes 0
#   * x_hat = bf16_rn(code * scale)        ->  4 + 16/Q4_GROUP = 4.5 bits per element
# ---------------------------------------------------------------------------------------------
Q4_GROUP = 32


def q4_pack_chunk(chunk_bits: np.ndarray):
    """(L, 2, n, H, D) bf16 bit patterns -> (codes uint8 (L,2,n,H,D/2), scales uint16 bf16 bits (L,2,n,H,D/G))."""
    L, two, n, H, D = chunk_bits.shape
    assert D % Q4_GROUP == 0
    x = bf16_bits_to_f32(chunk_bits).reshape(L, two, n, H, D // Q4_GROUP, Q4_GROUP)
    amax = np.abs(x).max(axis=-1)
    s_bits = np.where(amax == 0, f32_to_bf16_bits_rn(np.float32(1.0)),
                      f32_to_bf16_bits_rn((amax / np.float32(7.0)).astype(np.float32)))
    s = bf16_bits_to_f32(s_bits)
    # one IEEE division per group, then rint of the EXACT product x * inv (a float64 holds it exactly: 24 + 24
    # mantissa bits) — what one fused multiply-add against 1.5 * 2^23 yields on the GPU.  |x * inv| <= 7 (1 + 2^-8).
    inv = (np.float32(1.0) / s).astype(np.float32)
    q = np.clip(np.rint(x.astype(np.float64) * inv[..., None].astype(np.float64)), -7, 7).astype(np.int8)
    q = q.reshape(L, two, n, H, D)
    nib = (q & 0xF).astype(np.uint8)
    codes = (nib[..., 0::2] | (nib[..., 1::2] << 4)).astype(np.uint8)
    return codes, s_bits.astype(np.uint16)


def q4_unpack_chunk(codes: np.ndarray, scale_bits: np.ndarray) -> np.ndarray:
    L, two, n, H, Dh = codes.shape
    lo = (codes & 0xF).astype(np.int8)
    hi = (codes >> 4).astype(np.int8)
    lo = np.where(lo > 7, lo - 16, lo)
    hi = np.where(hi > 7, hi - 16, hi)
    q = np.empty((L, two, n, H, 2 * Dh), dtype=np.float32)
    q[..., 0::2], q[..., 1::2] = lo, hi
    s = bf16_bits_to_f32(scale_bits)
    x = (q.reshape(L, two, n, H, -1, Q4_GROUP) * s[..., None]).astype(np.float32)
    return f32_to_bf16_bits_rn(x.reshape(L, two, n, H, 2 * Dh))


def q4_tolerance(x: np.ndarray) -> np.ndarray:
    """|x - x_hat| <= scale/2 + 2^-8 |x| with scale <= absmax_group/7 * (1 + 2^-8): half a quantisation
    step of the group plus the bf16 roundings of scale and result."""
    L, two, n, H, D = x.shape
    amax = np.abs(x.reshape(L, two, n, H, D // Q4_GROUP, Q4_GROUP)).max(axis=-1, keepdims=True)
    step = np.broadcast_to(amax / 7.0 * (1 + 2.0 ** -7), (L, two, n, H, D // Q4_GROUP, Q4_GROUP)).reshape(x.shape)
    return step / 2 + np.abs(x) * 2.0 ** -7


# ---------------------------------------------------------------------------------------------
# engine semantics (store / retrieve / lookup as the adapter drives them)
# ---------------------------------------------------------------------------------------------
class OracleEngine:
    """Restates the observable behaviour of LMCacheEngine.store / .retrieve / lookup under the
    calling conventions of [vllm-0.22] lmcache_integration/vllm_v1_adapter.py
    (store :1115-1123, retrieve :882-889, lookup :1187-1191)."""

    def __init__(self, chunk_tokens: int = 256, fmt: str = "raw", seed: int = 0,
                 capacity_chunks: int | None = None):
        self.C = chunk_tokens
        self.fmt = fmt
        self.seed = seed
        self.capacity = capacity_chunks
        self.pool: dict[int, tuple] = {}  # key -> (n_tok, payload...) in LRU order

    def _touch(self, key):
        v = self.pool.pop(key)
        self.pool[key] = v

    def lookup(self, tokens) -> int:
        """Number of tokens in the longest stored prefix, whole chunks only."""
        hit = 0
        n = len(tokens)
        for i, k in enumerate(chunk_keys(tokens, self.C, self.seed, True)):
            n_tok = min(self.C, n - i * self.C)
            if k in self.pool and self.pool[k][0] == n_tok:
                self._touch(k)
                hit += n_tok
            else:
                break
        return hit

    def store(self, tokens, mask, kv_layers, slot_mapping, offset: int = 0) -> int:
        """mask is False exactly on the chunk-aligned prefix [0, offset) (adapter :1084-1091).
        Chunks already present are skipped.  Returns number of chunks written."""
        n = len(tokens)
        assert len(slot_mapping) == n and offset % self.C == 0
        assert int(np.count_nonzero(~np.asarray(mask))) == offset
        keys = chunk_keys(tokens, self.C, self.seed, True)
        wrote = 0
        for c in range(offset // self.C, len(keys)):
            s, e = c * self.C, min((c + 1) * self.C, n)
            if keys[c] in self.pool:
                continue
            bits = gather_tokens(kv_layers, np.asarray(slot_mapping[s:e]))
            if self.fmt == "fp8":
                payload = fp8_pack_chunk(bits)
            elif self.fmt == "q4":
                payload = q4_pack_chunk(bits)
            else:
                payload = (bits.copy(),)
            if self.capacity is not None and len(self.pool) >= self.capacity:
                self.pool.pop(next(iter(self.pool)))  # LRU
            self.pool[keys[c]] = (e - s, *payload)
            wrote += 1
        return wrote

    def retrieve(self, tokens, mask, kv_layers, slot_mapping) -> np.ndarray:
        """Scatter every stored chunk after the masked (chunk-aligned) prefix until the first
        miss; returns the bool mask of tokens written (adapter :882-905)."""
        n = len(tokens)
        skip = int(np.count_nonzero(~np.asarray(mask)))
        assert skip % self.C == 0
        ret = np.zeros(n, dtype=bool)
        keys = chunk_keys(tokens, self.C, self.seed, True)
        for c in range(skip // self.C, len(keys)):
            s, e = c * self.C, min((c + 1) * self.C, n)
            ent = self.pool.get(keys[c])
            if ent is None or ent[0] != e - s:
                break
            self._touch(keys[c])
            bits = (fp8_unpack_chunk(ent[1], ent[2]) if self.fmt == "fp8"
                    else q4_unpack_chunk(ent[1], ent[2]) if self.fmt == "q4" else ent[1])
            scatter_tokens(kv_layers, bits, np.asarray(slot_mapping[s:e]))
            ret[s:e] = True
        return ret


def num_new_matched_tokens(hit_tokens: int, num_computed_tokens: int, request_num_tokens: int) -> int:
    """The adapter's post-lookup arithmetic ([vllm-0.22] vllm_v1_adapter.py:1205-1228): tokens to
    load beyond vLLM's own prefix hit; a full-prompt hit recomputes the last token."""
    need = hit_tokens - num_computed_tokens
    if hit_tokens == request_num_tokens:
        need -= 1
    return max(need, 0)


def plan_save(input_token_len: int, prompt_len: int, num_saved_tokens: int, chunk_tokens: int,
              discard_partial_chunks: bool, is_decode_phase: bool = False,
              save_decode_cache: bool = False, skip_save: bool = False):
    """Which tokens a step saves ([vllm-0.22] vllm_v1_adapter.py:292-338): returns
    (skip_leading_tokens aligned down to a chunk, num_tokens_to_save) or None when nothing is
    saved this step."""
    is_last_prefill = input_token_len == prompt_len
    chunk_boundary = -(-(num_saved_tokens + 1) // chunk_tokens) * chunk_tokens
    skip = (skip_save or (num_saved_tokens > 0 and input_token_len < chunk_boundary)
            or (is_decode_phase and not save_decode_cache))
    if skip:
        return None
    n_save = (input_token_len // chunk_tokens * chunk_tokens
              if (not is_last_prefill or discard_partial_chunks) else input_token_len)
    lead = num_saved_tokens // chunk_tokens * chunk_tokens
    if lead >= n_save:
        return None
    return lead, n_save

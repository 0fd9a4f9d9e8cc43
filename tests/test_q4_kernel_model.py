"""A line-by-line Python model of kv_q4_store_kernel / kv_q4_load_kernel's INDEX MATH (b200kv_kernels.cuh)
run on numpy byte arrays standing in for the paged cache: the records it produces must be the ones the
oracle specifies, for NHD and HND tiles.  Catches addressing / packing mistakes before the kernels meet a
GPU (the arithmetic on values is the oracle's own here)."""
import numpy as np
import pytest

from oracle import kv_oracle as ko

L, NB, BS, H, D, C = 2, 12, 16, 4, 64, 64
TB = H * D * 2                      # token bytes (bf16)
ROW = D * 2                         # head_bytes
RV = ROW // 16
VPT = TB // 16
REC = H * D // 2 + H * (D // 32) * 2
SLAB = C * REC


def paged_bytes(host, hnd):
    """Per plane one flat uint8 array laid out like the device tensor; block stride = BS*TB."""
    planes = []
    for l in range(L):
        for kv in range(2):
            t = host[l][kv]                                     # (NB, BS, H, D) uint16
            if hnd:
                t = t.transpose(0, 2, 1, 3)                     # (NB, H, BS, D)
            planes.append(np.ascontiguousarray(t).view(np.uint8).reshape(-1))
    return planes


def src_off(slot, v, hnd):                                      # q4_src_addr
    blk, off = divmod(slot, BS)
    if not hnd:
        return blk * BS * TB + off * TB + v * 16                # paged_addr + v*16
    h, c = divmod(v, RV)
    return blk * BS * TB + h * BS * ROW + off * ROW + c * 16    # paged_addr_hnd + c*16


def model_store(planes, runs, n_chunks, hnd):
    out = np.zeros(n_chunks * 2 * L * SLAB, dtype=np.uint8)
    codes_bytes = VPT * 4
    for (a, b, n) in runs:
        c, t0 = divmod(b, C)
        for plane in range(2 * L):
            slab = c * 2 * L * SLAB + plane * SLAB
            nvec = n * VPT
            for base in range(0, nvec, 256):
                vecs = {}
                for tid in range(256):
                    idx = base + tid
                    if idx < nvec:
                        t, v = divmod(idx, VPT)
                        o = src_off(a + t, v, hnd)
                        vecs[tid] = (t, v, planes[plane][o:o + 16].view(np.uint16).copy())
                for tid, (t, v, x) in vecs.items():
                    quad = [vecs[q][2] for q in range(tid & ~3, (tid & ~3) + 4)]        # shfl_xor 1, 2
                    grp = np.concatenate(quad)
                    amax_bits = int((grp & 0x7FFF).max())
                    amax = ko.bf16_bits_to_f32(np.array([amax_bits], np.uint16))[0]
                    sb = ko.f32_to_bf16_bits_rn(np.array([amax / np.float32(7.0)], np.float32))[0] if amax_bits else \
                        ko.f32_to_bf16_bits_rn(np.array([1.0], np.float32))[0]
                    s = ko.bf16_bits_to_f32(np.array([sb], np.uint16))[0]
                    xf = ko.bf16_bits_to_f32(x)
                    q = np.clip(np.rint(xf.astype(np.float64) * np.float64(np.float32(1.0) / s)), -7, 7).astype(np.int64)
                    packed = 0
                    for i in range(4):
                        packed |= ((int(q[2 * i]) & 0xF) | ((int(q[2 * i + 1]) & 0xF) << 4)) << (8 * i)
                    rec = slab + (t0 + t) * REC
                    out[rec + v * 4: rec + v * 4 + 4] = np.frombuffer(int(packed).to_bytes(4, "little"), np.uint8)
                    if (v & 3) == 0:
                        out[rec + codes_bytes + (v >> 2) * 2: rec + codes_bytes + (v >> 2) * 2 + 2] = \
                            np.array([sb], np.uint16).view(np.uint8)
    return out


def runs_of(sm):
    runs, cur = [], None
    for i, s in enumerate(sm):
        s = int(s)
        if cur and s == cur[0] + cur[2] and s % BS != 0 and i % C != 0 and i % BS != 0:
            cur[2] += 1
        else:
            if cur:
                runs.append(tuple(cur))
            cur = [s, i, 1]
    runs.append(tuple(cur))
    return runs


def expected(host_layers, sm, n_chunks):
    out = np.zeros(n_chunks * 2 * L * SLAB, dtype=np.uint8)
    for c in range(n_chunks):
        seg = sm[c * C:(c + 1) * C]
        codes, scales = ko.q4_pack_chunk(ko.gather_tokens(host_layers, seg))
        for l in range(L):
            for kv in range(2):
                for t in range(len(seg)):
                    o = c * 2 * L * SLAB + (2 * l + kv) * SLAB + t * REC
                    out[o:o + H * D // 2] = codes[l, kv, t].reshape(-1)
                    out[o + H * D // 2:o + REC] = scales[l, kv, t].reshape(-1).view(np.uint8)
    return out


@pytest.mark.parametrize("hnd", [False, True])
@pytest.mark.parametrize("n_tok", [1, 17, 64, 100])
def test_q4_store_index_math_matches_the_oracle_layout(hnd, n_tok):
    rng = np.random.default_rng(n_tok)
    host = [ko.f32_to_bf16_bits_rn((rng.standard_normal((2, NB, BS, H, D)) * 3).astype(np.float32)) for _ in range(L)]
    sm = ko.slot_mapping_from_blocks(rng.permutation(NB)[: (n_tok + BS - 1) // BS], BS, n_tok)
    n_chunks = (n_tok + C - 1) // C
    got = model_store(paged_bytes(host, hnd), runs_of(sm), n_chunks, hnd)
    assert np.array_equal(got, expected(host, sm, n_chunks))

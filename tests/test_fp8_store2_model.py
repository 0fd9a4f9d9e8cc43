"""Python model of kv_fp8_store2_kernel's ADDRESSING (b200kv_kernels.cuh, B200KV_FP8_2PASS): per-token
source addresses from the run list, the thread -> (token group, column) mapping of both passes, and the
output index for NHD and HND chunks.  With the oracle's arithmetic on the values the produced chunk must be
the oracle's FP8 chunk (codes and scales), for block-structured and token-granular mappings."""
import numpy as np
import pytest

from oracle import kv_oracle as ko

L, NB, BS, H, D, C, K = 1, 24, 16, 4, 64, 256, 8        # K = cluster size
TB, HB = H * D * 2, D * 2
VPT, RV, W = TB // 16, HB // 16, C // K


def planes_of(host, hnd):
    out = []
    for kv in range(2):
        t = host[0][kv]
        if hnd:
            t = t.transpose(0, 2, 1, 3)
        out.append(np.ascontiguousarray(t).view(np.uint8).reshape(-1))
    return out


def s_src(slot, hnd):
    blk, off = divmod(slot, BS)
    return blk * BS * TB + (off * HB if hnd else off * TB)          # paged_addr_hnd(.., h=0) / paged_addr


def runs_of(sm):
    runs, cur = [], None
    for i, s in enumerate(sm):
        s = int(s)
        if cur and s == cur[0] + cur[2] and s % BS != 0 and i % C != 0 and i % BS != 0:
            cur[2] += 1
        else:
            if cur:
                runs.append(cur)
            cur = [s, i, 1]
    return runs + [cur]


def model(planes, sm, hnd):
    n = len(sm)
    n_chunks = (n + C - 1) // C
    slab = C * TB // 2
    codes = np.zeros(n_chunks * 2 * slab, np.uint8)
    scales = np.zeros((n_chunks, 2, H), np.float32)
    head_stride = BS * HB if hnd else HB
    groups = 256 // VPT if VPT < 256 else 1
    for c in range(n_chunks):
        for plane in range(2):
            amax = np.zeros(H, np.uint16)
            src_of = {}
            for rank in range(K):                                  # one CTA of the cluster each
                win_lo = c * C + rank * W
                n_valid = max(0, min(W, min((c + 1) * C, n) - win_lo))
                src = {}
                for a, b, ln in runs_of(sm):
                    for t in range(max(b, win_lo), min(b + ln, win_lo + n_valid)):
                        src[t - win_lo] = s_src(a + (t - b), hnd)
                src_of[rank] = (src, n_valid)
                for tid in range(256):                              # pass 1
                    grp = tid // VPT
                    if grp >= groups:
                        continue
                    col = tid - grp * VPT
                    while col < VPT:
                        h, cc = divmod(col, RV)
                        for t in range(grp, n_valid, groups):
                            o = src[t] + h * head_stride + cc * 16
                            amax[h] = max(amax[h], int((planes[plane][o:o + 16].view(np.uint16) & 0x7FFF).max()))
                        col += 256 if groups == 1 else VPT
            am = ko.bf16_bits_to_f32(amax)
            inv = np.where(amax == 0, np.float32(1), np.float32(448) / np.where(amax == 0, 1, am)).astype(np.float32)
            scales[c, plane] = np.where(amax == 0, np.float32(1), am / np.float32(448))
            for rank in range(K):                                  # pass 2
                src, n_valid = src_of[rank]
                out0 = c * 2 * slab + plane * slab + rank * W * (TB // 2)
                for tid in range(256):
                    grp = tid // VPT
                    if grp >= groups:
                        continue
                    col = tid - grp * VPT
                    while col < VPT:
                        h, cc = divmod(col, RV)
                        for t in range(grp, n_valid, groups):
                            o = src[t] + h * head_stride + cc * 16
                            x = ko.bf16_bits_to_f32(planes[plane][o:o + 16].view(np.uint16))
                            q = ko.f32_to_e4m3_satfinite((x * inv[h]).astype(np.float32))
                            idx = (((t // BS) * H + h) * BS + (t % BS)) * RV + cc if hnd else t * VPT + col
                            codes[out0 + idx * 8: out0 + idx * 8 + 8] = q
                        col += 256 if groups == 1 else VPT
    return codes, scales


def expected(host, sm, hnd):
    n = len(sm)
    n_chunks = (n + C - 1) // C
    slab = C * TB // 2
    codes = np.zeros(n_chunks * 2 * slab, np.uint8)
    scales = np.zeros((n_chunks, 2, H), np.float32)
    for c in range(n_chunks):
        seg = sm[c * C:(c + 1) * C]
        q, s = ko.fp8_pack_chunk(ko.gather_tokens(host, seg))       # (1,2,n,H,D), (1,2,H)
        scales[c] = s[0]
        for kv in range(2):
            for t in range(len(seg)):
                for h in range(H):
                    if hnd:       # tiles verbatim: [tile][H][bs][D]
                        o = c * 2 * slab + kv * slab + (((t // BS) * H + h) * BS + t % BS) * D
                    else:
                        o = c * 2 * slab + kv * slab + t * H * D + h * D
                    codes[o:o + D] = q[0, kv, t, h]
    return codes, scales


@pytest.mark.parametrize("hnd", [False, True])
@pytest.mark.parametrize("n_tok,granular", [(1, False), (40, False), (256, False), (300, False), (40, True)])
def test_two_pass_store_addressing(hnd, n_tok, granular):
    rng = np.random.default_rng(n_tok + 7 * hnd)
    host = [ko.f32_to_bf16_bits_rn((rng.standard_normal((2, NB, BS, H, D)) * 2).astype(np.float32))]
    if granular:
        sm = rng.permutation(NB * BS)[:n_tok].astype(np.int64)
    else:
        sm = ko.slot_mapping_from_blocks(rng.permutation(NB)[: (n_tok + BS - 1) // BS], BS, n_tok)
    got_c, got_s = model(planes_of(host, hnd), sm, hnd)
    want_c, want_s = expected(host, sm, hnd)
    assert np.array_equal(got_s, want_s)
    assert np.array_equal(got_c, want_c)

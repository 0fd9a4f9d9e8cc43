"""Pin the oracle: golden fixtures (tests/golden/make_golden.py) and Python-vs-C agreement."""
import os

import numpy as np
import pytest

from oracle import kv_oracle as ko
from tests import oracle_c

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_gather_scatter_match_vllm_slot_mapping_semantics():
    z = np.load(os.path.join(G, "paged_gather.npz"))
    layers = [z["layers"][i].view(np.uint16) for i in range(z["layers"].shape[0])]
    bs = layers[0].shape[2]
    n = len(z["slot_mapping"])
    sm = ko.slot_mapping_from_blocks(z["block_ids"], bs, n)
    assert np.array_equal(sm, z["slot_mapping"])                       # adapter :368-375
    got = ko.gather_tokens(layers, sm)
    assert np.array_equal(got, z["gathered"].view(np.uint16))          # example_connector :247-248
    dst = [np.full_like(l, 0xFFFF) for l in layers]
    ko.scatter_tokens(dst, got, z["dst_map"])
    assert np.array_equal(np.stack(dst), z["scattered"].view(np.uint16))  # example_connector :154-159


def test_e4m3_encoder_matches_ml_dtypes_and_torch():
    z = np.load(os.path.join(G, "e4m3_vectors.npz"))
    assert np.array_equal(ko.f32_to_e4m3_satfinite(z["x"]), z["codes"])
    dec = ko.e4m3_decode_table()
    fin = np.isfinite(z["decode"])
    assert np.array_equal(dec[fin], z["decode"][fin]) and np.isnan(dec[~fin]).all()


def test_e4m3_satfinite_and_nan():
    x = np.float32([449, 464, 465, 1e9, np.inf, -500, -np.inf, np.nan])
    want = np.uint8([0x7E, 0x7E, 0x7E, 0x7E, 0x7E, 0xFE, 0xFE, 0x7F])
    got = ko.f32_to_e4m3_satfinite(x)
    assert np.array_equal(got[:7], want[:7]) and (got[7] & 0x7F) == 0x7F


def test_bf16_rounding_matches_torch():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(20000) * np.exp(rng.uniform(-20, 20, 20000))).astype(np.float32)
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(ko.f32_to_bf16_bits_rn(x), want)


def _layers(rng, L, NB, bs, H, D, scale_spread=True):
    out = []
    for l in range(L):
        x = rng.standard_normal((2, NB, bs, H, D)).astype(np.float32)
        if scale_spread:
            x *= np.exp(rng.uniform(-3, 3, (2, 1, 1, H, 1))).astype(np.float32)
        out.append(ko.f32_to_bf16_bits_rn(x).reshape(2, NB, bs, H, D))
    return out


@pytest.mark.parametrize("fmt", ["raw", "fp8"])
@pytest.mark.parametrize("n_tok", [1, 15, 16, 100, 256, 257, 600])
def test_c_oracle_matches_python_oracle(fmt, n_tok):
    rng = np.random.default_rng(n_tok)
    L, NB, bs, H, D, C_ = 3, 64, 16, 4, 32, 256
    layers = _layers(rng, L, NB, bs, H, D)
    blocks = rng.permutation(NB)[: (n_tok + bs - 1) // bs]
    sm = ko.slot_mapping_from_blocks(blocks, bs, n_tok)
    chunks, cb, so = oracle_c.gather(layers, sm, C_, fmt)
    for c in range((n_tok + C_ - 1) // C_):
        s, e = c * C_, min((c + 1) * C_, n_tok)
        bits = ko.gather_tokens(layers, sm[s:e])
        raw = chunks[c * cb:(c + 1) * cb]
        if fmt == "raw":
            got = raw.view(np.uint16).reshape(L, 2, C_, H, D)[:, :, : e - s]
            assert np.array_equal(got, bits)
        else:
            codes, scales = ko.fp8_pack_chunk(bits)
            got = raw[:so].reshape(L, 2, C_, H, D)[:, :, : e - s]
            assert np.array_equal(got, codes)
            assert np.array_equal(raw[so:so + L * 2 * H * 4].view(np.float32).reshape(L, 2, H), scales)
    dst_c = [np.zeros_like(l) for l in layers]
    dst_p = [np.zeros_like(l) for l in layers]
    dblocks = rng.permutation(NB)[: (n_tok + bs - 1) // bs]
    dm = ko.slot_mapping_from_blocks(dblocks, bs, n_tok)
    oracle_c.scatter(dst_c, dm, C_, chunks, cb, so, fmt)
    oe = ko.OracleEngine(C_, fmt)
    toks = np.arange(n_tok, dtype=np.int32)
    oe.store(toks, np.ones(n_tok, bool), layers, sm)
    assert oe.retrieve(toks, np.ones(n_tok, bool), dst_p, dm).all()
    for a, b in zip(dst_c, dst_p):
        assert np.array_equal(a, b)


def test_fp8_roundtrip_within_stated_tolerance():
    rng = np.random.default_rng(0)
    L, n, H, D = 2, 256, 8, 128
    x = (rng.standard_normal((L, 2, n, H, D)) * np.exp(rng.uniform(-4, 4, (L, 2, 1, H, 1)))).astype(np.float32)
    bits = ko.f32_to_bf16_bits_rn(x).reshape(x.shape)
    xb = ko.bf16_bits_to_f32(bits)
    codes, scales = ko.fp8_pack_chunk(bits)
    back = ko.bf16_bits_to_f32(ko.fp8_unpack_chunk(codes, scales))
    tol = ko.fp8_tolerance(xb, np.broadcast_to(scales[:, :, None, :, None], xb.shape))
    assert (np.abs(back - xb) <= tol).all()
    assert (scales > 0).all()
    zeros = np.zeros((1, 2, 4, 2, 8), np.uint16)
    c0, s0 = ko.fp8_pack_chunk(zeros)
    assert (c0 == 0).all() and (s0 == 1).all()


def test_engine_semantics_mask_offset_prefix_lru():
    rng = np.random.default_rng(5)
    L, NB, bs, H, D, C_ = 2, 128, 16, 2, 16, 64
    layers = _layers(rng, L, NB, bs, H, D, False)
    n = 3 * C_ + 10
    toks = rng.integers(0, 1000, n).astype(np.int32)
    sm = ko.slot_mapping_from_blocks(rng.permutation(NB)[: (n + bs - 1) // bs], bs, n)
    oe = ko.OracleEngine(C_)
    assert oe.lookup(toks) == 0
    mask = np.ones(n, bool)
    mask[:C_] = False
    assert oe.store(toks, mask, layers, sm, offset=C_) == 3     # chunk 0 skipped by the mask
    assert oe.lookup(toks) == 0                                   # prefix semantics: chunk 0 missing
    assert oe.store(toks, np.ones(n, bool), layers, sm) == 1      # only chunk 0 is new
    assert oe.lookup(toks) == n
    assert oe.lookup(toks[: 2 * C_ + 5]) == 2 * C_                # different partial tail -> miss
    dst = [np.zeros_like(l) for l in layers]
    ret = oe.retrieve(toks, mask, dst, sm)
    assert not ret[:C_].any() and ret[C_:].all()
    assert not np.stack(dst).reshape(L, 2, NB * bs, H, D)[:, :, sm[:C_]].any()   # masked prefix untouched
    assert np.array_equal(ko.gather_tokens(dst, sm[C_:]), ko.gather_tokens(layers, sm[C_:]))


def test_adapter_arithmetic():
    # full-prompt hit recomputes the last token (adapter :1205-1209)
    assert ko.num_new_matched_tokens(512, 0, 512) == 511
    assert ko.num_new_matched_tokens(512, 256, 600) == 256
    assert ko.num_new_matched_tokens(256, 256, 600) == 0
    # save planning (adapter :292-338)
    assert ko.plan_save(600, 600, 0, 256, discard_partial_chunks=False) == (0, 600)
    assert ko.plan_save(600, 600, 0, 256, discard_partial_chunks=True) == (0, 512)
    assert ko.plan_save(600, 900, 0, 256, discard_partial_chunks=False) == (0, 512)   # chunked prefill
    assert ko.plan_save(700, 900, 512, 256, discard_partial_chunks=False) is None     # below next boundary
    assert ko.plan_save(900, 900, 512, 256, discard_partial_chunks=False) == (512, 900)
    assert ko.plan_save(901, 900, 900, 256, False, is_decode_phase=True) is None


def test_q4_groupwise_codec_spec():
    """The sub-8-bit format is specified in the oracle before any kernel exists: size,
    tolerance, exactness on representable inputs, idempotence of re-quantisation."""
    rng = np.random.default_rng(11)
    x = (rng.standard_normal((2, 2, 37, 4, 128)) * np.exp(rng.uniform(-4, 4, (2, 2, 37, 4, 1)))).astype(np.float32)
    x[0, 0, 0, 0, :32] = 0.0                                    # an all-zero group
    bits = ko.f32_to_bf16_bits_rn(x)
    xb = ko.bf16_bits_to_f32(bits)
    codes, scales = ko.q4_pack_chunk(bits)
    assert codes.shape == (2, 2, 37, 4, 64) and scales.shape == (2, 2, 37, 4, 4) and scales.dtype == np.uint16
    assert codes.nbytes + scales.nbytes == bits.nbytes * 576 // 2048           # 4.5 bits per element
    back = ko.bf16_bits_to_f32(ko.q4_unpack_chunk(codes, scales))
    assert np.all(np.abs(back - xb) <= ko.q4_tolerance(xb))
    assert not back[0, 0, 0, 0, :32].any()
    codes2, scales2 = ko.q4_pack_chunk(ko.q4_unpack_chunk(codes, scales))       # idempotent
    assert np.array_equal(ko.q4_unpack_chunk(codes2, scales2), ko.q4_unpack_chunk(codes, scales))
    # values that are exact multiples of a bf16-exact step round-trip exactly
    q = rng.integers(-7, 8, (1, 2, 5, 4, 128)).astype(np.float32)
    q[..., ::32] = 7.0                                           # pins absmax = 7 * step in every group
    y = ko.f32_to_bf16_bits_rn(q * np.float32(0.125))
    assert np.array_equal(ko.q4_unpack_chunk(*ko.q4_pack_chunk(y)), y)

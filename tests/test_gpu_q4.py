"""Q4 (group-wise 4-bit) format on the GPU, against its oracle specification (oracle/kv_oracle.py
q4_pack_chunk / q4_unpack_chunk): the stored records — code bytes and bf16 scales — and the pages after a
retrieve must equal the oracle's, bit for bit, for NHD and HND tiles.  EXPERIMENTAL format."""
import numpy as np
import pytest

from oracle import kv_oracle as ko

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import b200kv  # noqa: E402
from b200kv import FMT_Q4, KVEngine, KVGeometry, KVPool  # noqa: E402
from tests.test_gpu_kernels import SMALL, bits_of, logical_bits, mk_host_layers, need_gpu, to_dev, to_dev_hnd  # noqa: E402


def expected_chunks(host, sm, C, n_chunks, chunk_bytes):
    """Chunk bytes per the spec: per plane a slab of C token records [H*D/2 codes][H*D/32 bf16 scales]."""
    L = len(host)
    H, D = host[0].shape[3], host[0].shape[4]
    rec = H * D // 2 + H * (D // 32) * 2
    out = np.zeros(n_chunks * chunk_bytes, dtype=np.uint8)
    for c in range(n_chunks):
        seg = sm[c * C:(c + 1) * C]
        codes, scales = ko.q4_pack_chunk(ko.gather_tokens(host, seg))          # (L,2,n,H,D/2), (L,2,n,H,D/32)
        for l in range(L):
            for kv in range(2):
                plane = 2 * l + kv
                for t in range(len(seg)):
                    o = c * chunk_bytes + plane * C * rec + t * rec
                    out[o:o + H * D // 2] = codes[l, kv, t].reshape(-1)
                    out[o + H * D // 2:o + rec] = scales[l, kv, t].reshape(-1).view(np.uint8)
    return out, rec


@pytest.mark.parametrize("hnd", [False, True])
@pytest.mark.parametrize("n_tok", [1, 17, 256, 300, 700])
def test_q4_gather_scatter_vs_oracle(hnd, n_tok):
    need_gpu()
    p = SMALL
    rng = np.random.default_rng(4000 + n_tok)
    host = mk_host_layers(rng, p["L"], p["NB"], p["bs"], p["H"], p["D"])
    dev = to_dev_hnd(host) if hnd else to_dev(host)
    geom = KVGeometry(p["L"], p["H"], p["D"], p["NB"], p["bs"], p["C"], 2, 2 * p["bs"] * p["H"] * p["D"] * 2 if hnd else 0,
                      FMT_Q4, b200kv._lib.LAYOUT_HND if hnd else b200kv._lib.LAYOUT_NHD)
    eng = KVEngine(geom, None, 0, staging_bytes=0)
    eng.register_kv_caches(dev)
    nb = (n_tok + p["bs"] - 1) // p["bs"]
    sm = ko.slot_mapping_from_blocks(rng.permutation(p["NB"])[:nb], p["bs"], n_tok)
    dm = ko.slot_mapping_from_blocks(rng.permutation(p["NB"])[:nb], p["bs"], n_tok)
    n_chunks = (n_tok + p["C"] - 1) // p["C"]
    buf = torch.zeros(n_chunks * geom.chunk_bytes, dtype=torch.uint8, device="cuda:0")
    eng.gather(sm, buf.data_ptr())
    torch.cuda.synchronize()
    want, rec = expected_chunks(host, sm, p["C"], n_chunks, geom.chunk_bytes)
    assert rec * p["C"] * 2 * p["L"] <= geom.chunk_bytes
    assert np.array_equal(buf.cpu().numpy(), want)
    for t in dev:
        t.zero_()
    eng.scatter(dm, buf.data_ptr())
    torch.cuda.synchronize()
    dst = [np.zeros_like(l) for l in host]
    for c in range(n_chunks):
        seg_s, seg_d = sm[c * p["C"]:(c + 1) * p["C"]], dm[c * p["C"]:(c + 1) * p["C"]]
        ko.scatter_tokens(dst, ko.q4_unpack_chunk(*ko.q4_pack_chunk(ko.gather_tokens(host, seg_s))), seg_d)
    for a, b in zip(dev, dst):
        assert np.array_equal(logical_bits(a) if hnd else bits_of(a), b)
    eng.close()


def test_q4_store_retrieve_through_pool_and_tolerance():
    need_gpu()
    p = SMALL
    rng = np.random.default_rng(5)
    host = mk_host_layers(rng, p["L"], p["NB"], p["bs"], p["H"], p["D"])
    dev = to_dev(host)
    geom = KVGeometry(p["L"], p["H"], p["D"], p["NB"], p["bs"], p["C"], 2, 0, FMT_Q4)
    assert geom.chunk_bytes * 32 <= KVGeometry(p["L"], p["H"], p["D"], p["NB"], p["bs"], p["C"]).chunk_bytes * 9 + 32 * 256
    pool = KVPool(None, 6 * geom.chunk_bytes, geom.chunk_bytes, 1)
    eng = KVEngine(geom, pool, 0, staging_bytes=4 * geom.chunk_bytes)
    eng.register_kv_caches(dev)
    n = 2 * p["C"] + 37
    toks = rng.integers(0, 128256, n).astype(np.int32)
    sm = ko.slot_mapping_from_blocks(rng.permutation(p["NB"])[: (n + 15) // 16], 16, n)
    dm = ko.slot_mapping_from_blocks(rng.permutation(p["NB"])[: (n + 15) // 16], 16, n)
    eng.wait(eng.store(toks, None, sm))
    for t in dev:
        t.zero_()
    assert eng.retrieve(toks, None, dm).all()
    torch.cuda.synchronize()
    oe = ko.OracleEngine(p["C"], "q4")
    dst = [np.zeros_like(l) for l in host]
    oe.store(toks, np.ones(n, bool), host, sm)
    oe.retrieve(toks, np.ones(n, bool), dst, dm)
    for a, b, src in zip(dev, dst, host):
        assert np.array_equal(bits_of(a), b)
        x = ko.bf16_bits_to_f32(ko.gather_tokens([src], sm))
        y = ko.bf16_bits_to_f32(ko.gather_tokens([bits_of(a)], dm))
        assert np.all(np.abs(x - y) <= ko.q4_tolerance(x))
    eng.close()

"""Cross-process peer pull over CUDA IPC (the disaggregated-prefill data path): a producer
process publishes its paged cache, a consumer process maps it and pulls blocks with the in-kernel
P2P kernel.  Both on GPU 0 when only one GPU is visible, producer on GPU 1 otherwise."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L, NB, BS, H, D = 4, 128, 16, 8, 128


def _producer(dev_index, eid, ready, stop, shm_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "production-stack_b200")]
    os.environ["B200KV_SHM_DIR"] = shm_dir
    import torch
    from b200kv import KVEngine, KVGeometry, pd
    pd.SHM_DIR = shm_dir
    torch.cuda.set_device(dev_index)
    g = torch.Generator(device=f"cuda:{dev_index}").manual_seed(42)
    caches = [torch.randn((2, NB, BS, H, D), generator=g, device=f"cuda:{dev_index}", dtype=torch.float32).bfloat16()
              for _ in range(L)]
    eng = KVEngine(KVGeometry(L, H, D, NB, BS, 256), None, dev_index, staging_bytes=0)
    eng.register_kv_caches(caches)
    torch.cuda.synchronize()
    pd.publish_ipc(eid, eng, dev_index)
    ready.set()
    stop.wait(120)
    pd.unpublish_ipc(eid)
    eng.close()


def test_pull_from_another_process(tmp_path):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected but no CUDA device")
    sys.path[:0] = [ROOT, os.path.join(ROOT, "production-stack_b200")]
    import torch.multiprocessing as mp

    from b200kv import KVEngine, KVGeometry, pd
    from oracle import kv_oracle as ko
    shm_dir = str(tmp_path)
    pd.SHM_DIR = shm_dir
    prod_dev = 1 if torch.cuda.device_count() > 1 else 0
    ctx = mp.get_context("spawn")
    ready, stop = ctx.Event(), ctx.Event()
    eid = f"prod-{os.getpid()}"
    proc = ctx.Process(target=_producer, args=(prod_dev, eid, ready, stop, shm_dir))
    proc.start()
    try:
        assert ready.wait(180), "producer did not come up"
        local = [torch.zeros((2, NB, BS, H, D), device="cuda:0", dtype=torch.bfloat16) for _ in range(L)]
        eng = KVEngine(KVGeometry(L, H, D, NB, BS, 256), None, 0, staging_bytes=0)
        eng.register_kv_caches(local)
        worker = pd.PDWorker(eng, "cons", BS)
        rng = np.random.default_rng(0)
        n_tok = 1000
        remote_blocks = [int(x) for x in rng.permutation(NB)[: (n_tok + BS - 1) // BS]]
        local_blocks = [int(x) for x in rng.permutation(NB)[: (n_tok + BS - 1) // BS]]
        spec = pd.PullSpec("req-d", eid, "req-p", remote_blocks, local_blocks, n_tok, skip_tokens=32)
        worker.start_pulls(pd.PDMeta([spec], {}))
        assert worker.take_failed_blocks() == set()
        for _ in range(2000):
            if worker.poll()[1]:
                break
            torch.cuda.synchronize()
        else:
            pytest.fail("pull never completed")
        assert os.path.exists(os.path.join(pd.done_dir(eid), "req-p"))      # producer may free its blocks
        # expected bytes: regenerate the producer's cache with the same seed on this device
        g = torch.Generator(device="cuda:0").manual_seed(42)
        want = [torch.randn((2, NB, BS, H, D), generator=g, device="cuda:0", dtype=torch.float32).bfloat16()
                for _ in range(L)]
        src = ko.slot_mapping_from_blocks(remote_blocks, BS, n_tok)[32:]
        dst = ko.slot_mapping_from_blocks(local_blocks, BS, n_tok)[32:]
        si, di = torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda()
        for a, w in zip(local, want):
            fa, fw = a.view(2, NB * BS, H * D), w.view(2, NB * BS, H * D)
            assert torch.equal(fa[:, di].view(torch.int16), fw[:, si].view(torch.int16))
            untouched = torch.ones(NB * BS, dtype=torch.bool, device="cuda:0")
            untouched[di] = False
            assert not fa[:, untouched].any()
        assert eng.stats()["p2p_bytes"] == (n_tok - 32) * 2 * L * H * D * 2
        eng.close()
    finally:
        stop.set()
        proc.join(60)

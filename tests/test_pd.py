"""Disaggregated-prefill hand-off state machine (CPU, fake engine) — the kv_transfer_params
round trip the unmodified router performs (/root/reference/src/vllm_router/services/
request_service/request.py:771-778, 823-829)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest

from b200kv import pd

BS = 16


class FakeEngine:
    geom = NS(n_blocks=64, stride=4096, n_layers=2, n_kv_heads=2, head_dim=8, block_tokens=BS, elem_bytes=2, layout=0)

    def __init__(self):
        self.pulls, self.done = [], set()

    def export_ipc(self):
        return b"\x07" * (72 * 4)

    def import_peer(self, pid, device, descs, stride, n_blocks):
        self.imported = (pid, device, len(descs), stride, n_blocks)

    def peer_pull(self, peer, src, dst, stream=None):
        self.pulls.append((peer, np.asarray(src).copy(), np.asarray(dst).copy()))
        return len(self.pulls)

    def poll(self, ticket):
        return ticket in self.done


@pytest.fixture
def shm(tmp_path, monkeypatch):
    monkeypatch.setattr(pd, "SHM_DIR", str(tmp_path))
    return tmp_path


def test_prefill_to_decode_round_trip(shm):
    # ---------------- prefill replica ("P") ----------------
    p_eng = FakeEngine()
    pd.publish_ipc("P-engine", p_eng, device=3)
    p_sched = pd.PDScheduler("P-engine", BS, lease_s=60)
    p_work = pd.PDWorker(p_eng, "P-engine", BS)
    prompt = list(range(100))
    # the router's prefill request (request.py:771-778)
    p_req = NS(request_id="cmpl-1", prompt_token_ids=prompt, num_computed_tokens=100,
               kv_transfer_params={"do_remote_decode": True, "do_remote_prefill": False, "remote_engine_id": None,
                                   "remote_block_ids": None, "remote_host": None, "remote_port": None})
    p_blocks = [40, 41, 42, 43, 44, 45, 46]
    delay, params = p_sched.request_finished(p_req, p_blocks)
    assert delay is True
    assert params["do_remote_prefill"] is True and params["remote_engine_id"] == "P-engine"
    assert params["remote_block_ids"] == p_blocks and params["remote_num_tokens"] == 100
    meta = p_sched.build_meta()
    assert meta.held.keys() == {"cmpl-1"}
    p_work.start_pulls(meta)
    assert p_work.poll() == (set(), set())               # nobody pulled yet: blocks stay held
    # ---------------- router: copies params, fills remote_host (request.py:823-829) -------------
    params = dict(params, remote_host="10.0.0.5")
    # ---------------- decode replica ("D") ----------------
    d_eng = FakeEngine()
    d_sched = pd.PDScheduler("D-engine", BS)
    d_work = pd.PDWorker(d_eng, "D-engine", BS)
    d_req = NS(request_id="cmpl-1-d", prompt_token_ids=prompt, kv_transfer_params=params)
    ext = d_sched.remote_prefill_tokens(d_req, num_computed_tokens=32)
    assert ext == 99 - 32                                  # last prompt token recomputed locally
    d_blocks = [5, 6, 7, 8, 9, 10, 11]
    d_sched.after_alloc(d_req, d_blocks, ext, 32)
    assert d_req.kv_transfer_params["do_remote_prefill"] is False
    assert d_sched.remote_prefill_tokens(d_req, 32) is None   # consumed: second call is an ordinary request
    dmeta = d_sched.build_meta()
    assert len(dmeta.pulls) == 1 and dmeta.pulls[0].skip_tokens == 32 and dmeta.pulls[0].n_tokens == 99
    d_work.start_pulls(dmeta)
    assert d_eng.imported == (0, 3, 72 * 4, 4096, 64)
    peer, src, dst = d_eng.pulls[0]
    assert len(src) == 99 - 32
    assert src[0] == 42 * BS and dst[0] == 7 * BS          # token 32 = first token of the third block
    assert src[-1] == 46 * BS + 2 and dst[-1] == 11 * BS + 2
    assert d_work.poll() == (set(), set())                 # pull still in flight
    d_eng.done.add(1)
    assert d_work.poll() == (set(), {"cmpl-1-d"})
    # ---------------- back on P: the marker releases the held blocks ----------------
    sent, _ = p_work.poll()
    assert sent == {"cmpl-1"}
    p_sched.sending_finished(sent)
    assert p_sched._held == {} and p_work.poll() == (set(), set())
    pd.unpublish_ipc("P-engine")
    assert not os.path.exists(pd.ipc_path("P-engine"))


def test_lease_expiry_and_failures(shm):
    eng = FakeEngine()
    sched = pd.PDScheduler("P2", BS, lease_s=0.0)
    work = pd.PDWorker(eng, "P2", BS)
    req = NS(request_id="r", prompt_token_ids=[1] * 20, num_computed_tokens=20, kv_transfer_params={"do_remote_decode": True})
    assert sched.request_finished(req, [1, 2])[0] is True
    work.start_pulls(sched.build_meta())
    assert work.poll()[0] == {"r"}                          # lease expired without a consumer
    # not a P/D request, aborted request, no blocks -> nothing held
    assert sched.request_finished(NS(request_id="x", prompt_token_ids=[1], kv_transfer_params=None), [1]) == (False, None)
    assert sched.request_finished(req, [1], finished_ok=False) == (False, None)
    assert sched.request_finished(req, []) == (False, None)
    # consumer whose peer is unknown: blocks reported for recompute, never an exception
    d_sched = pd.PDScheduler("D2", BS)
    d_req = NS(request_id="d", prompt_token_ids=list(range(40)),
               kv_transfer_params={"do_remote_prefill": True, "remote_engine_id": "nowhere", "remote_block_ids": [[3, 4, 5]],
                                   "remote_request_id": "r9", "remote_num_tokens": 40})
    n = d_sched.remote_prefill_tokens(d_req, 0)
    assert n == 39
    d_sched.after_alloc(d_req, [7, 8, 9], n, 0)
    work.start_pulls(d_sched.build_meta())
    assert work.take_failed_blocks() == {7, 8, 9} and eng.pulls == []


class WaitableEngine(FakeEngine):
    def wait(self, ticket):
        self.done.add(ticket)


def _prefill(shm, eid="P3", lease_s=60.0, tp=1):
    eng = FakeEngine()
    pd.publish_ipc(eid, eng, device=0, rank=0, tp_size=tp)
    sched = pd.PDScheduler(eid, BS, lease_s=lease_s, tp_size=tp)
    req = NS(request_id="p-req", prompt_token_ids=list(range(64)), num_computed_tokens=64,
             kv_transfer_params={"do_remote_decode": True})
    delay, params = sched.request_finished(req, [10, 11, 12, 13])
    assert delay and params["remote_lease_id"] and params["tp_size"] == tp
    assert os.path.exists(pd.lease_path(eid, params["remote_lease_id"]))
    return sched, params


def test_expired_lease_turns_a_pull_into_reported_load_errors(shm):
    """The producer frees held blocks when its lease runs out; a consumer that pulls later (or is still
    pulling) must report its blocks for recompute, not decode on recycled pages."""
    sched, params = _prefill(shm)
    d_eng = WaitableEngine()
    d_sched, d_work = pd.PDScheduler("D3", BS), pd.PDWorker(d_eng, "D3", BS)
    d_req = NS(request_id="d-req", prompt_token_ids=list(range(64)), kv_transfer_params=dict(params))
    n = d_sched.remote_prefill_tokens(d_req, 0)
    d_sched.after_alloc(d_req, [1, 2, 3, 4], n, 0)
    meta = d_sched.build_meta()
    assert meta.pulls[0].remote_lease_id == params["remote_lease_id"]
    # (a) lease gone while the pull is in flight: detected by the check AFTER the pull
    d_work.start_pulls(meta)
    assert len(d_eng.pulls) == 1 and d_work.take_failed_blocks() == set()
    sched.sending_finished({"p-req"})                      # producer: lease expired -> file removed, blocks freed next
    assert not os.path.exists(pd.lease_path("P3", params["remote_lease_id"]))
    assert d_work.poll(block=True)[1] == {"d-req"}         # still reported as finished ...
    assert d_work.take_failed_blocks() == {1, 2, 3, 4}     # ... with its blocks invalid
    # (b) lease already gone before the pull starts: no pull is issued at all
    d_work.start_pulls(meta)
    assert len(d_eng.pulls) == 1 and d_work.take_failed_blocks() == {1, 2, 3, 4} and d_work.n_lease_failures == 2


def test_live_lease_pull_is_clean_and_lease_removed_before_blocks_are_freed(shm):
    sched, params = _prefill(shm, "P4")
    d_eng = WaitableEngine()
    d_sched, d_work = pd.PDScheduler("D4", BS), pd.PDWorker(d_eng, "D4", BS)
    d_req = NS(request_id="d", prompt_token_ids=list(range(64)), kv_transfer_params=dict(params))
    n = d_sched.remote_prefill_tokens(d_req, 0)
    d_sched.after_alloc(d_req, [1, 2, 3, 4], n, 0)
    d_work.start_pulls(d_sched.build_meta())
    assert d_work.poll(block=True)[1] == {"d"} and d_work.take_failed_blocks() == set()
    assert os.path.exists(os.path.join(pd.done_dir("P4"), "p-req"))
    sched.sending_finished({"p-req"})
    assert not os.path.exists(pd.lease_path("P4", params["remote_lease_id"]))


def test_tensor_parallel_ranks_have_their_own_control_files_and_sizes_must_match(shm):
    e0, e1 = FakeEngine(), FakeEngine()
    p0 = pd.publish_ipc("PT", e0, device=0, rank=0, tp_size=2)
    p1 = pd.publish_ipc("PT", e1, device=1, rank=1, tp_size=2)
    assert p0 != p1 and os.path.isdir(pd.done_dir("PT", 0)) and os.path.isdir(pd.done_dir("PT", 1))
    # decode rank 1 maps prefill rank 1's cache (device 1), not whoever wrote last
    d1 = FakeEngine()
    assert pd.PeerResolver(d1, rank=1, tp_size=2).resolve("PT") == 0 and d1.imported[1] == 1
    d0 = FakeEngine()
    assert pd.PeerResolver(d0, rank=0, tp_size=2).resolve("PT") == 0 and d0.imported[1] == 0
    # a TP=1 decoder must not pull half of the heads from a TP=2 prefiller
    with pytest.raises(ValueError):
        pd.PeerResolver(FakeEngine(), rank=0, tp_size=1).resolve("PT")
    sched1 = pd.PDScheduler("D", BS, tp_size=1)
    req = NS(request_id="x", prompt_token_ids=list(range(40)),
             kv_transfer_params={"do_remote_prefill": True, "remote_engine_id": "PT", "remote_block_ids": [[1, 2, 3]], "tp_size": 2})
    assert sched1.remote_prefill_tokens(req, 0) is None
    # each rank's marker releases that rank only
    w0, w1 = pd.PDWorker(e0, "PT", BS, rank=0, tp_size=2), pd.PDWorker(e1, "PT", BS, rank=1, tp_size=2)
    held = pd.PDMeta([], {"r": 1e18})
    w0.start_pulls(held)
    w1.start_pulls(held)
    open(os.path.join(pd.done_dir("PT", 1), "r"), "w").close()
    assert w0.poll()[0] == set() and w1.poll()[0] == {"r"}


def test_fewer_local_blocks_than_promised_tokens_is_a_load_failure(shm):
    _, params = _prefill(shm, "P5")
    d_eng = WaitableEngine()
    d_work = pd.PDWorker(d_eng, "D5", BS)
    spec = pd.PullSpec("d", "P5", "p-req", [10, 11, 12, 13], [7], 63, remote_lease_id=params["remote_lease_id"])
    d_work.start_pulls(pd.PDMeta([spec], {}))
    assert d_eng.pulls == [] and d_work.take_failed_blocks() == {7}

"""Scheduler/worker state machine of the connector against fake SchedulerOutputs (CPU).
The worker half is driven with an engine stand-in that forwards to the oracle engine, so the
end-to-end request flow (lookup -> alloc -> load -> save) is checked token for token."""
from types import SimpleNamespace as NS

import numpy as np

from b200kv.adapter import (LoadSpec, ReqMeta, RequestTracker, SchedulerState, WorkerState, first_group,
                            make_req_meta)
from oracle import kv_oracle as ko

BS, C = 16, 64


class OracleBackedEngine:
    """Test stand-in with KVEngine's call signatures, backed by oracle.OracleEngine."""

    def __init__(self, layers):
        self.oe = ko.OracleEngine(C)
        self.layers = layers
        self.calls = []

    def store(self, tokens, mask, slot_mapping, offset=0, stream=None):
        self.calls.append(("store", len(tokens), offset))
        self.oe.store(tokens, mask, self.layers, slot_mapping, offset)
        return 1

    def retrieve(self, tokens, mask, slot_mapping, stream=None, return_ticket=False, layers_per_group=0):
        if layers_per_group:
            self.calls.append(("retrieve_layerwise", len(tokens), layers_per_group))
            return self.oe.retrieve(tokens, mask, self.layers, slot_mapping), 91
        self.calls.append(("retrieve", len(tokens), int((~mask).sum()), stream))
        ret = self.oe.retrieve(tokens, mask, self.layers, slot_mapping)
        return (ret, 77) if return_ticket else ret

    done = True

    def poll(self, t):
        return self.done

    def wait_layer(self, ticket, layer, stream=None):
        self.calls.append(("wait_layer", ticket, layer))


def sched_out(new=(), cached=None, num_sched=None, finished=()):
    cached = cached or NS(req_ids=[], new_block_ids=[], resumed_req_ids=set(), all_token_ids={})
    return NS(scheduled_new_reqs=list(new), scheduled_cached_reqs=cached,
              num_scheduled_tokens=num_sched or {}, finished_req_ids=set(finished))


def new_req(rid, prompt, blocks, computed=0):
    return NS(req_id=rid, prompt_token_ids=list(prompt), block_ids=(list(blocks),), num_computed_tokens=computed)


def test_first_group_shapes():
    assert first_group(([1, 2], [9])) == [1, 2]
    assert first_group([3, 4]) == [3, 4]
    assert first_group(None) == [] and first_group(()) == []


def test_save_planning_matches_oracle_plan_save():
    for prompt_len, n_in, saved, discard, decode in [
        (600, 600, 0, False, False), (600, 600, 0, True, False), (900, 600, 0, False, False),
        (900, 700, 512, False, False), (900, 900, 512, False, False), (900, 901, 900, False, True),
        (64, 64, 0, False, False), (10, 10, 0, False, False), (10, 10, 0, True, False)]:
        tr = RequestTracker("r", prompt_len, list(range(n_in)), list(range(100)), num_saved_tokens=saved,
                            is_decode_phase=decode)
        m = make_req_meta(tr, BS, 256, None, discard)
        want = ko.plan_save(n_in, prompt_len, saved, 256, discard, decode)
        if want is None:
            assert m is None or not m.save_spec.can_save or m.save_spec.skip_leading_tokens // 256 * 256 >= len(m.token_ids)
        else:
            lead, n_save = want
            assert m is not None and m.save_spec.can_save
            assert m.save_spec.skip_leading_tokens // 256 * 256 == lead and len(m.token_ids) == n_save


def test_request_flow_miss_then_hit_with_last_token_rule():
    rng = np.random.default_rng(0)
    layers = [rng.integers(0, 2 ** 16, (2, 64, BS, 2, 8), dtype=np.uint16) for _ in range(2)]
    eng = OracleBackedEngine(layers)
    lookup = lambda toks: eng.oe.lookup(toks)
    sched = SchedulerState(lookup, BS, C, discard_partial_chunks=False)
    worker = WorkerState(eng, BS, C)
    prompt = list(rng.integers(0, 1000, 3 * C + 10))           # 202 tokens
    blocks = list(range(10, 10 + 13))
    # --- turn 1: nothing cached ---------------------------------------------------------------
    req = NS(request_id="a", prompt_token_ids=prompt, num_tokens=len(prompt), all_token_ids=prompt)
    assert sched.num_new_matched_tokens("a", prompt, len(prompt), 0) == 0
    sched.after_alloc(req, 0)
    metas = sched.build_meta(sched_out([new_req("a", prompt, blocks)], num_sched={"a": len(prompt)}))
    assert len(metas) == 1 and metas[0].is_last_prefill and metas[0].load_spec is None
    worker.start_load(metas)
    worker.save(metas)
    assert eng.calls == [("store", len(prompt), 0)]
    assert eng.oe.lookup(prompt) == len(prompt)                 # partial tail chunk saved too
    # decode steps never save (adapter :313-318)
    cached = NS(req_ids=["a"], new_block_ids=[None], resumed_req_ids=set(), all_token_ids={})
    req.all_token_ids = prompt + [5]
    assert sched.build_meta(sched_out(cached=cached, num_sched={"a": 1})) == []
    sched.build_meta(sched_out(finished=["a"]))
    assert "a" not in sched.trackers
    # --- turn 2: same prompt -> full hit, last token recomputed (adapter :1205-1209) ------------
    req2 = NS(request_id="b", prompt_token_ids=prompt, num_tokens=len(prompt), all_token_ids=prompt)
    need = sched.num_new_matched_tokens("b", prompt, len(prompt), 0)
    assert need == len(prompt) - 1
    sched.after_alloc(req2, need)
    blocks2 = list(range(40, 53))
    metas = sched.build_meta(sched_out([new_req("b", prompt, blocks2, computed=need)], num_sched={"b": 1}))
    m = metas[0]
    assert m.load_spec is not None and m.load_spec.can_load and m.load_spec.external_cached_tokens == len(prompt)
    eng.calls.clear()
    before = [l.copy() for l in layers]
    worker.start_load(metas)
    assert eng.calls == [("retrieve", len(prompt), 0, None)]
    sm_src = ko.slot_mapping_from_blocks(blocks, BS, len(prompt))
    sm_dst = ko.slot_mapping_from_blocks(blocks2, BS, len(prompt))
    assert np.array_equal(ko.gather_tokens(layers, sm_dst), ko.gather_tokens(before, sm_src))
    worker.save(metas)                                          # everything already saved -> no store
    assert [c for c in eng.calls if c[0] == "store"] == []
    assert worker.take_load_errors() == set()
    # --- turn 3: longer prompt sharing 2 chunks, vLLM prefix cache already holds 1 chunk ---------
    prompt3 = prompt[: 2 * C] + list(rng.integers(1000, 2000, 100))
    need3 = sched.num_new_matched_tokens("c", prompt3, len(prompt3), C)
    assert need3 == C                                           # 2 chunks hit - 1 chunk computed
    req3 = NS(request_id="c", prompt_token_ids=prompt3, num_tokens=len(prompt3), all_token_ids=prompt3)
    sched.after_alloc(req3, need3)
    metas = sched.build_meta(sched_out([new_req("c", prompt3, list(range(20, 35)), computed=2 * C)],
                                       num_sched={"c": len(prompt3) - 2 * C}))
    eng.calls.clear()
    worker.start_load(metas)
    assert eng.calls == [("retrieve", 2 * C, C, None)]          # first chunk masked (vLLM has it)
    worker.save(metas)
    assert eng.calls[-1] == ("store", len(prompt3), 2 * C)      # only the new tail is stored
    assert sched.num_lookups == 3 and sched.num_hit_tokens == len(prompt) + 2 * C


def test_short_load_reports_error_blocks():
    class ShortEngine(OracleBackedEngine):
        def retrieve(self, tokens, mask, slot_mapping, stream=None):
            r = np.zeros(len(tokens), bool)
            r[:C] = True                                        # only the first chunk arrives
            return r

    eng = ShortEngine([np.zeros((2, 8, BS, 1, 8), np.uint16)])
    w = WorkerState(eng, BS, C)
    m = ReqMeta("x", np.arange(3 * C, dtype=np.int32), list(range(100, 112)), load_spec=LoadSpec(0, 3 * C, True))
    w.start_load([m])
    assert w.take_load_errors() == set(range(100 + C // BS, 112))
    assert w.take_load_errors() == set()


def test_consumer_never_saves_and_producer_never_looks_up():
    sched = SchedulerState(lambda t: 128, BS, C, False, kv_role="kv_producer")
    assert sched.num_new_matched_tokens("p", list(range(200)), 200, 0) == 0
    eng = OracleBackedEngine([np.zeros((2, 8, BS, 1, 8), np.uint16)])
    w = WorkerState(eng, BS, C, kv_role="kv_consumer")
    from b200kv.adapter import SaveSpec
    w.save([ReqMeta("x", np.arange(C, dtype=np.int32), [0, 1, 2, 3], True, SaveSpec(0, True))])
    assert eng.calls == []


def test_chunked_prefill_saves_whole_chunks_incrementally():
    eng = OracleBackedEngine([np.zeros((2, 64, BS, 1, 8), np.uint16)])
    sched = SchedulerState(lambda t: 0, BS, C, False)
    w = WorkerState(eng, BS, C)
    prompt = list(range(1, 3 * C + 21))                         # 212 tokens, prefill in 100-token steps
    req = NS(request_id="q", prompt_token_ids=prompt, num_tokens=len(prompt), all_token_ids=prompt)
    sched.num_new_matched_tokens("q", prompt, len(prompt), 0)
    sched.after_alloc(req, 0)
    w.save(sched.build_meta(sched_out([new_req("q", prompt, list(range(7)))], num_sched={"q": 100})))
    assert eng.calls == [("store", C, 0)]                       # 100 tokens -> one whole chunk
    cached = NS(req_ids=["q"], new_block_ids=[(list(range(7, 14)),)], resumed_req_ids=set(), all_token_ids={})
    w.save(sched.build_meta(sched_out(cached=cached, num_sched={"q": 100})))
    assert eng.calls[-1] == ("store", 3 * C, C)                 # 200 tokens -> chunks 1,2 new
    cached = NS(req_ids=["q"], new_block_ids=[None], resumed_req_ids=set(), all_token_ids={})
    w.save(sched.build_meta(sched_out(cached=cached, num_sched={"q": 12})))
    # reference rule (adapter :313-316): once something is saved, a step that does not reach the
    # next chunk boundary saves nothing — even the last prefill's partial tail
    assert len(eng.calls) == 2


def test_async_load_flow_request_waits_for_remote_kvs():
    """B200KV_ASYNC_LOAD (default): get_num_new_matched_tokens -> (n, True); the load is emitted
    for a request that is NOT in the SchedulerOutput, runs detached, is reported through
    finished_recving, and the request is then scheduled as new with num_computed_tokens set
    (vllm/v1/core/sched/scheduler.py:587-660, 763-781, 2069-2125)."""
    rng = np.random.default_rng(3)
    layers = [rng.integers(0, 2 ** 16, (2, 64, BS, 2, 8), dtype=np.uint16) for _ in range(2)]
    eng = OracleBackedEngine(layers)
    sched = SchedulerState(lambda t: eng.oe.lookup(t), BS, C, False, async_load=True)
    worker = WorkerState(eng, BS, C)
    prompt = list(rng.integers(0, 1000, 3 * C))                 # exactly 3 chunks
    sm0 = ko.slot_mapping_from_blocks(list(range(12)), BS, len(prompt))
    eng.oe.store(np.asarray(prompt, np.int32), np.ones(len(prompt), bool), layers, sm0)
    req = NS(request_id="a", prompt_token_ids=prompt, num_tokens=len(prompt), all_token_ids=prompt)
    need = sched.num_new_matched_tokens("a", prompt, len(prompt), 0)
    assert need == len(prompt)                                  # full hit: vLLM drops the last token itself
    blocks = (list(range(30, 42)),)
    sched.after_alloc(req, need, blocks)                        # request goes to WAITING_FOR_REMOTE_KVS
    metas = sched.build_meta(sched_out())                       # ... and is absent from the SchedulerOutput
    assert len(metas) == 1 and metas[0].async_load and metas[0].save_spec is None
    eng.calls.clear()
    eng.done = False
    worker.start_load(metas)
    assert eng.calls == [("retrieve", len(prompt), 0, "detached")]
    assert worker.poll_async_loads() == set()                   # still in flight
    eng.done = True
    assert worker.poll_async_loads() == {"a"}                   # -> finished_recving
    sm1 = ko.slot_mapping_from_blocks(blocks[0], BS, len(prompt))
    assert np.array_equal(ko.gather_tokens(layers, sm1), ko.gather_tokens(layers, sm0))
    # second update_state_after_alloc (num_external_tokens = 0), then scheduled as a new request
    sched.after_alloc(req, 0, blocks)
    metas = sched.build_meta(sched_out([new_req("a", prompt, blocks[0], computed=len(prompt) - 1)], num_sched={"a": 1}))
    eng.calls.clear()
    worker.start_load(metas)
    worker.save(metas)
    assert eng.calls == []                                      # nothing to load again, nothing new to store


def test_engine_failures_never_raise_on_the_data_path():
    """SURVEY §8b "Errors": a failed load becomes load-error blocks (vLLM recomputes), a failed store
    becomes a future miss."""
    from b200kv.adapter import SaveSpec

    class Broken(OracleBackedEngine):
        def retrieve(self, *a, **k):
            raise RuntimeError("cuda went away")

        def store(self, *a, **k):
            raise RuntimeError("pool is gone")

    w = WorkerState(Broken([np.zeros((2, 8, BS, 1, 8), np.uint16)]), BS, C)
    m = ReqMeta("x", np.arange(2 * C, dtype=np.int32), list(range(50, 58)), True, SaveSpec(0, True), LoadSpec(0, 2 * C, True))
    w.start_load([m])
    assert w.take_load_errors() == set(range(50, 58))
    w.save([m])                                   # logged, skipped
    assert w.pending_tickets == []
    a = ReqMeta("y", np.arange(C, dtype=np.int32), [1, 2, 3, 4], load_spec=LoadSpec(0, C, True), async_load=True)
    w.start_load([a])
    assert w.poll_async_loads() == {"y"} and w.take_load_errors() == {1, 2, 3, 4}


def test_layerwise_worker_flow_and_safety_net():
    """LMCache `use_layerwise`: start_load issues a layer-wise retrieve, wait_for_layer_load forwards
    per layer; if the hooks never ran the loaded blocks are reported for recompute."""
    rng = np.random.default_rng(8)
    layers = [rng.integers(0, 2 ** 16, (2, 64, BS, 2, 8), dtype=np.uint16) for _ in range(2)]
    eng = OracleBackedEngine(layers)
    w = WorkerState(eng, BS, C)
    toks = np.arange(2 * C, dtype=np.int32)
    sm = ko.slot_mapping_from_blocks(list(range(8)), BS, 2 * C)
    eng.oe.store(toks, np.ones(2 * C, bool), layers, sm)
    m = ReqMeta("r", toks, list(range(20, 28)), load_spec=LoadSpec(C, 2 * C, True))
    w.start_load([m], layers_per_group=4)
    assert eng.calls[-1] == ("retrieve_layerwise", 2 * C, 4)
    assert w.layer_loads == [(91, list(range(24, 28)))]        # the blocks past vLLM's own prefix hit
    w.wait_layer(4)
    assert eng.calls[-1] == ("wait_layer", 91, 4)
    w.abandon_layer_loads()
    assert w.take_load_errors() == {24, 25, 26, 27} and w.layer_loads == []
    w.start_load([])                                            # a new step forgets the previous tickets
    assert w.layer_loads == []


def test_foreign_tokens_counts_chunks_of_other_owners(shm_name):
    """Shared pool (BASELINE.json config 3): tokens served from chunks another replica stored."""
    from b200kv.engine import KVPool, chunk_keys

    pool = KVPool(shm_name, 64 * 1024, 1024, 3)
    toks = np.arange(3 * C + 10, dtype=np.int32)
    keys = chunk_keys(toks, C, 7, True)
    for i, (k, owner) in enumerate(zip(keys, (11, 22, 22, 11))):
        pool.reserve(int(k), min(C, len(toks) - i * C), 0, owner)
        pool.commit(int(k))
    eng = NS(pool=pool, _keys=lambda t: chunk_keys(t, C, 7, True))
    w = WorkerState(eng, BS, C, owner_tag=11)
    assert w._foreign_tokens(toks, 0, len(toks)) == 2 * C            # chunks 1 and 2 belong to owner 22
    assert w._foreign_tokens(toks, C + 8, 3 * C) == 2 * C - 8        # clipped to the loaded range
    assert w._foreign_tokens(toks, 0, C) == 0
    assert WorkerState(eng, BS, C, owner_tag=22)._foreign_tokens(toks, 0, len(toks)) == C + 10
    pool.close()
    KVPool.unlink(shm_name)


def test_on_stored_hook_gets_the_new_chunk_keys_only():
    """Remote tier upload hook: after each store the worker hands over the keys of exactly the chunks
    this store added (not the leading ones an earlier step already saved)."""
    from b200kv.engine import chunk_keys

    eng = OracleBackedEngine([np.zeros((2, 64, BS, 1, 8), np.uint16)])
    eng._keys = lambda t: chunk_keys(t, C, 99, True)
    sched = SchedulerState(lambda t: 0, BS, C, False)
    w = WorkerState(eng, BS, C)
    seen = []
    w.on_stored = lambda keys: seen.append([int(k) for k in keys])
    prompt = list(range(1, 3 * C + 21))
    want = [int(k) for k in chunk_keys(np.asarray(prompt, np.int32), C, 99, True)]
    req = NS(request_id="q", prompt_token_ids=prompt, num_tokens=len(prompt), all_token_ids=prompt)
    sched.num_new_matched_tokens("q", prompt, len(prompt), 0)
    sched.after_alloc(req, 0)
    w.save(sched.build_meta(sched_out([new_req("q", prompt, list(range(7)))], num_sched={"q": 100})))
    cached = NS(req_ids=["q"], new_block_ids=[(list(range(7, 14)),)], resumed_req_ids=set(), all_token_ids={})
    w.save(sched.build_meta(sched_out(cached=cached, num_sched={"q": 100})))
    assert seen == [want[:1], want[1:3]]
    w.on_stored = lambda keys: 1 / 0          # a failing hook never fails the step
    w2 = WorkerState(eng, BS, C)
    w2.on_stored = w.on_stored
    p2 = list(range(500, 500 + C))
    r2 = NS(request_id="z", prompt_token_ids=p2, num_tokens=C, all_token_ids=p2)
    sched.num_new_matched_tokens("z", p2, C, 0)
    sched.after_alloc(r2, 0)
    w2.save(sched.build_meta(sched_out([new_req("z", p2, list(range(20, 24)))], num_sched={"z": C})))
    assert eng.calls[-1] == ("store", C, 0)


def test_priority_limit_serves_but_does_not_save_low_priority_requests():
    """LMCACHE_PRIORITY_LIMIT (adapter :1163, :1332-1337): a request whose priority value exceeds the
    limit may load from the pool but its KV is not stored."""
    eng = OracleBackedEngine([np.zeros((2, 64, BS, 1, 8), np.uint16)])
    sched = SchedulerState(lambda t: 0, BS, C, False, priority_limit=1)
    w = WorkerState(eng, BS, C)
    for rid, prio, stored in (("lo", 5, False), ("hi", 1, True), ("dflt", 0, True)):
        prompt = list(range(100, 100 + 2 * C))
        req = NS(request_id=rid, prompt_token_ids=prompt, num_tokens=len(prompt), all_token_ids=prompt)
        sched.num_new_matched_tokens(rid, prompt, len(prompt), 0, priority=prio)
        sched.after_alloc(req, 0)
        n = len(eng.calls)
        w.save(sched.build_meta(sched_out([new_req(rid, prompt, list(range(8)))], num_sched={rid: len(prompt)})))
        assert (len(eng.calls) > n) == stored, rid


def test_preempted_request_restarts_on_its_new_blocks():
    """A request preempted mid-prefill comes back in `scheduled_cached_reqs` with `resumed_req_ids`, ALL of
    its block ids replaced and `num_computed_tokens` reset (vllm/v1/core/sched/output.py:112-126).  What the
    connector saves afterwards must be addressed through the new blocks, and nothing already saved is saved
    twice."""
    eng = OracleBackedEngine([np.zeros((2, 64, BS, 1, 8), np.uint16)])
    sched = SchedulerState(lambda t: 0, BS, C, False)
    w = WorkerState(eng, BS, C)
    prompt = list(range(1, 3 * C + 10))                                     # 202 tokens
    req = NS(request_id="p", prompt_token_ids=prompt, num_tokens=len(prompt), all_token_ids=prompt)
    sched.num_new_matched_tokens("p", prompt, len(prompt), 0)
    sched.after_alloc(req, 0)
    old_blocks = list(range(0, 7))
    metas = sched.build_meta(sched_out([new_req("p", prompt, old_blocks)], num_sched={"p": 100}))
    w.save(metas)
    assert eng.calls == [("store", C, 0)]                                    # chunk 0 saved from the old blocks
    new_blocks = list(range(20, 33))
    cached = NS(req_ids=["p"], new_block_ids=[(new_blocks,)], resumed_req_ids={"p"}, all_token_ids={},
                num_computed_tokens=[0])
    metas = sched.build_meta(sched_out(cached=cached, num_sched={"p": len(prompt)}))   # the whole prompt again
    (m,) = metas
    assert m.block_ids == new_blocks and len(m.token_ids) == len(prompt) and m.is_last_prefill
    assert list(m.slot_mapping(BS)[:3]) == [20 * BS, 20 * BS + 1, 20 * BS + 2]
    w.save(metas)
    assert eng.calls[-1] == ("store", len(prompt), C)                         # chunks 1.. only; chunk 0 is not re-saved
    # and it decodes normally afterwards
    req.all_token_ids = prompt + [7]
    cached = NS(req_ids=["p"], new_block_ids=[None], resumed_req_ids=set(), all_token_ids={}, num_computed_tokens=[len(prompt)])
    assert sched.build_meta(sched_out(cached=cached, num_sched={"p": 1})) == []


# ---- what besides the tokens decides the KV: multimodal items, cache_salt, LoRA, lmcache.tag.* ----------------
def _mm(ident, offset, length):
    return NS(identifier=ident, mm_position=NS(offset=offset, length=length))


def test_multimodal_items_enter_the_chunk_keys_like_the_reference_placeholder_rewrite():
    """adapter :198, :344-350, :1168-1172: same text + different image must not share KV after the image;
    chunks wholly BEFORE the placeholder are still shared (causal attention)."""
    from b200kv import chunk_keys
    from b200kv.adapter import request_identity
    prompt = list(range(100, 100 + 4 * C))
    req_a = NS(request_id="a", prompt_token_ids=prompt, mm_features=[_mm("deadbeef" * 8, 2 * C + 5, 20)])
    req_b = NS(request_id="b", prompt_token_ids=prompt, mm_features=[_mm("0badf00d" * 8, 2 * C + 5, 20)])
    req_a2 = NS(request_id="a2", prompt_token_ids=prompt, mm_features=[_mm("deadbeef" * 8, 2 * C + 5, 20)])
    ka, kb, ka2 = (chunk_keys(request_identity(r).apply(prompt), C, 7) for r in (req_a, req_b, req_a2))
    assert list(ka) == list(ka2)
    assert list(ka[:2]) == list(kb[:2]) == list(chunk_keys(prompt, C, 7)[:2])     # before the image: shared, = plain text
    assert ka[2] != kb[2] and ka[3] != kb[3]                                         # from the image on: distinct
    assert request_identity(NS(request_id="t", prompt_token_ids=prompt)) is None    # plain text: untouched
    # the window form used for decode-time tokens agrees with the whole-sequence form
    ident = request_identity(req_a)
    whole = ident.apply(prompt)
    assert np.array_equal(ident.apply(prompt[2 * C:3 * C], start=2 * C), whole[2 * C:3 * C])


def test_cache_salt_lora_and_tags_isolate_requests():
    from b200kv.adapter import request_identity
    prompt = list(range(3 * C))
    plain = np.asarray(prompt, dtype=np.int32)
    seen = [plain.tobytes()]
    for kw in (dict(cache_salt="tenant-a"), dict(cache_salt="tenant-b"), dict(lora_request=NS(lora_name="adapter-1", lora_int_id=1)),
               dict(sampling_params=NS(extra_args={"kv_transfer_params": {"lmcache.tag.user": "u1"}})),
               dict(sampling_params=NS(extra_args={"kv_transfer_params": {"lmcache.tag.user": "u2"}}))):
        ident = request_identity(NS(request_id="r", prompt_token_ids=prompt, **kw))
        assert ident is not None and ident.salt
        kt = ident.apply(prompt)
        assert kt.dtype == np.int32 and (kt >= 0).all() and kt.tobytes() not in seen
        seen.append(kt.tobytes())
    # a window of the sequence (tokens appended during decode) is salted like the same positions of the whole
    whole = ident.apply(prompt)
    for lo_ in (0, 1, C, C + 7):
        assert np.array_equal(ident.apply(prompt[lo_:lo_ + 33], start=lo_), whole[lo_:lo_ + 33])
    # lmcache.skip_save alone is a save rule, not an identity
    assert request_identity(NS(request_id="r", prompt_token_ids=prompt,
                               sampling_params=NS(extra_args={"kv_transfer_params": {"lmcache.skip_save": True}}))) is None


def test_scheduler_and_worker_use_key_tokens_end_to_end():
    """Image A stored; same text with image B misses from the placeholder's chunk on; with image A again: full hit."""
    from b200kv.adapter import request_identity
    rng = np.random.default_rng(3)
    layers = [rng.integers(0, 2 ** 16, (2, 64, BS, 2, 8), dtype=np.uint16) for _ in range(2)]
    eng = OracleBackedEngine(layers)
    sched = SchedulerState(lambda toks: eng.oe.lookup(toks), BS, C, discard_partial_chunks=False)
    worker = WorkerState(eng, BS, C)
    prompt = list(rng.integers(0, 1000, 3 * C))
    feat_a, feat_b = [_mm("aa" * 32, C + 3, 10)], [_mm("bb" * 32, C + 3, 10)]

    def turn(rid, feats, blocks):
        req = NS(request_id=rid, prompt_token_ids=prompt, num_tokens=len(prompt), all_token_ids=prompt, mm_features=feats)
        need = sched.num_new_matched_tokens(rid, prompt, len(prompt), 0, identity=request_identity(req))
        sched.after_alloc(req, need)
        nr = new_req(rid, prompt, blocks, computed=need)
        metas = sched.build_meta(sched_out([nr], num_sched={rid: len(prompt) - need}))
        worker.start_load(metas)
        worker.save(metas)
        sched.build_meta(sched_out(finished=[rid]))
        return need

    assert turn("a", feat_a, list(range(0, 12))) == 0
    assert turn("b", feat_b, list(range(12, 24))) == C            # only the chunk before the image is reusable
    assert turn("a2", feat_a, list(range(24, 36))) == 3 * C - 1   # full hit minus the recomputed last token
    assert not sched.identities                                    # dropped with the finished requests


# ---- one engine op per step: the requests of a step are handed to the engine together --------------------------
class BatchingEngine(OracleBackedEngine):
    """OracleBackedEngine + KVEngine's batch calls (same semantics, per request, through the oracle)."""

    def store_batch(self, reqs, stream=None):
        self.calls.append(("store_batch", [(len(t), off) for t, _sm, off in reqs]))
        for t, sm, off in reqs:
            mask = np.ones(len(t), bool)
            mask[:off] = False
            self.oe.store(t, mask, self.layers, sm, off)
        return 5

    def retrieve_batch(self, reqs, stream=None, layers_per_group=0):
        self.calls.append(("retrieve_batch", [(len(t), skip) for t, _sm, skip in reqs], layers_per_group))
        got = []
        for t, sm, skip in reqs:
            mask = np.ones(len(t), bool)
            mask[:skip] = False
            got.append(int(self.oe.retrieve(t, mask, self.layers, sm).sum()))
        return np.asarray(got, dtype=np.int64), 9


def test_requests_of_one_step_share_one_engine_op():
    rng = np.random.default_rng(21)
    layers = [rng.integers(0, 2 ** 16, (2, 96, BS, 2, 8), dtype=np.uint16) for _ in range(2)]
    ref_layers = [l.copy() for l in layers]
    eng, ref = BatchingEngine(layers), OracleBackedEngine(ref_layers)
    w, wr = WorkerState(eng, BS, C), WorkerState(ref, BS, C)
    prompts = [list(rng.integers(0, 1000, n)) for n in (2 * C + 5, C, 3 * C + 17)]
    blocks = [list(range(0, 9)), list(range(9, 13)), list(range(13, 27))]
    from b200kv.adapter import SaveSpec
    metas = lambda: [ReqMeta(f"r{i}", np.asarray(p, np.int32), b, is_last_prefill=True, save_spec=SaveSpec(0, True))  # noqa: E731
                     for i, (p, b) in enumerate(zip(prompts, blocks))]
    w.save(metas())
    wr.save(metas())
    assert [c[0] for c in eng.calls] == ["store_batch"] and len(eng.calls[0][1]) == 3      # ONE op for three requests
    assert [c[0] for c in ref.calls] == ["store"] * 3
    assert w.stats.num_stored_tokens == wr.stats.num_stored_tokens == sum(map(len, prompts))
    assert w.pending_tickets == [5]
    for p in prompts:
        assert eng.oe.lookup(p) == ref.oe.lookup(p) == len(p)
    # load them into other pages, the third one after a masked (vLLM-cached) first chunk; a fourth request misses
    for l in layers + ref_layers:
        l[:] = 0
    dst = [list(range(40, 49)), list(range(49, 53)), list(range(53, 67)), list(range(70, 74))]
    miss = list(rng.integers(2000, 3000, C + 3))
    load_metas = lambda: [ReqMeta(f"l{i}", np.asarray(p, np.int32), b, load_spec=LoadSpec(C if i == 2 else 0, len(p), True))  # noqa: E731
                          for i, (p, b) in enumerate(zip(prompts + [miss], dst))]
    w.start_load(load_metas(), layers_per_group=2)
    wr.start_load(load_metas(), layers_per_group=2)
    assert eng.calls[-1][0] == "retrieve_batch" and eng.calls[-1][1] == [(len(prompts[0]), 0), (C, 0), (len(prompts[2]), C),
                                                                       (len(miss), 0)] and eng.calls[-1][2] == 2
    assert w.stats.num_loaded_tokens == wr.stats.num_loaded_tokens == sum(map(len, prompts)) - C
    assert w.stats.num_load_shortfalls == wr.stats.num_load_shortfalls == 1
    assert w.take_load_errors() == wr.take_load_errors() == set(dst[3][:(len(miss) + BS - 1) // BS])
    assert len(w.layer_loads) == 1 and w.layer_loads[0][0] == 9            # one ticket the forward pass waits on per layer
    for a, b in zip(layers, ref_layers):
        assert np.array_equal(a, b)

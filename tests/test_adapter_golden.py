"""Row A4 (which tokens a step saves / loads, slot mapping, last-prefill and decode rules) pinned
against vLLM's vendored LMCache adapter: tests/golden/adapter_plan_vectors.json holds what
`RequestTracker` + `ReqMeta.from_request_tracker` of vllm_v1_adapter.py produced, step by step, for 160
randomised request histories (generator: tests/golden/make_adapter_golden.py).  The same histories are
replayed through b200kv.adapter and must agree on every decision."""
import json
import os

import numpy as np

from b200kv.adapter import LoadSpec, RequestTracker, make_req_meta
from b200kv.engine import xxh64
from oracle import kv_oracle as ko

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.environ.get("GOLDEN_OUT") or os.path.join(HERE, "golden")   # GOLDEN_OUT: a larger, uncommitted sweep


def test_planning_matches_the_vendored_lmcache_adapter_step_by_step():
    doc = json.load(open(os.path.join(GOLDEN, "adapter_plan_vectors.json")))
    assert "vllm_v1_adapter.py" in doc["source"] and len(doc["scenarios"]) >= 160
    n_meta = n_load = 0
    for sc in doc["scenarios"]:
        bs, chunk = sc["block_size"], sc["chunk"]
        tokens = list(range(1, sc["prompt_len"] + sc["n_decode"] + 1))
        tr = None
        for st in sc["steps"]:
            if st["kind"] == "new":
                # RequestTracker.from_new_request (:150-210): tokens up to what this step computes,
                # num_saved_tokens = LMCache hit when it will be loaded
                tr = RequestTracker(f"s{sc['id']}", sc["prompt_len"], tokens[: sc["first_tokens"]], list(st["new_blocks"]),
                                    num_saved_tokens=sc["lmcache_hit"] if sc["can_load"] else 0, skip_save=sc["skip_save"])
                spec = LoadSpec(sc["vllm_hit"], sc["lmcache_hit"], sc["can_load"]) if sc["lmcache_hit"] > sc["vllm_hit"] else None
            else:
                cur = len(tr.token_ids)
                tr.update(tokens[cur: cur + st["new_tokens"]], (st["new_blocks"],) if st["new_blocks"] else None)
                spec = None
            # the oracle's own restatement of the rule (oracle.plan_save) is pinned by the same vectors
            plan = ko.plan_save(len(tr.token_ids), sc["prompt_len"], tr.num_saved_tokens, chunk, sc["discard_partial_chunks"],
                                tr.is_decode_phase, sc["save_decode_cache"], tr.skip_save)
            ref_saves = st["meta"] is not None and st["meta"]["save"][1]
            if plan is None:
                assert (not ref_saves) or st["num_saved_tokens_after"] <= st["meta"]["save"][0] // chunk * chunk, (sc["id"], "oracle")
            else:
                assert ref_saves and plan == (st["meta"]["save"][0] // chunk * chunk, st["num_saved_tokens_after"]), (sc["id"], "oracle")
            m = make_req_meta(tr, bs, chunk, spec, sc["discard_partial_chunks"], sc["save_decode_cache"])
            where = (sc["id"], sc["steps"].index(st))
            assert tr.num_saved_tokens == st["num_saved_tokens_after"], where
            want = st["meta"]
            assert (m is None) == (want is None), where
            if m is None:
                continue
            n_meta += 1
            assert m.is_last_prefill == want["is_last_prefill"], where
            assert [m.save_spec.skip_leading_tokens, m.save_spec.can_save] == want["save"], where
            got_load = None if m.load_spec is None else [m.load_spec.vllm_cached_tokens, m.load_spec.external_cached_tokens]
            assert got_load == want["load"], where
            # tokens: the reference carries input[:num_tokens_to_save]; this build also carries the tokens of
            # a load that reaches past them (it loads from the metadata, the reference from request state)
            n_ref = want["n_tokens"]
            if want["load"] is None:
                assert len(m.token_ids) == n_ref, where
            else:
                n_load += 1
                assert len(m.token_ids) >= n_ref, where
            sm = m.slot_mapping(bs)[:n_ref].astype("<i8")
            assert len(sm) == want["slot_mapping_len"] and xxh64(sm.tobytes(), 0) == want["slot_mapping_xxh64"], where
    assert n_meta > 150 and n_load > 20


def test_scheduler_flows_match_the_vendored_connector_impl():
    """Whole request flows (several interleaved requests: lookup -> alloc -> chunked prefill -> decode ->
    finish; kv_both / kv_producer / kv_consumer) through SchedulerState, against what
    LMCacheConnectorV1Impl.get_num_new_matched_tokens / update_state_after_alloc / build_connector_meta
    produced for the same SchedulerOutputs (tests/golden/adapter_flow_vectors.json)."""
    from types import SimpleNamespace as NS

    from b200kv.adapter import SchedulerState, WorkerState

    class Recorder:
        """KVEngine's store / retrieve signatures; records what the reference's engine would be told."""

        def __init__(self):
            self.calls = []

        def retrieve(self, tokens, mask, slot_mapping, stream=None, return_ticket=False, layers_per_group=0):
            self.calls.append(["retrieve", len(tokens), int((~mask).sum()), xxh64(np.asarray(slot_mapping, "<i8").tobytes(), 0)])
            return mask.copy()

        def store(self, tokens, mask, slot_mapping, offset=0, stream=None):
            self.calls.append(["store", len(tokens), int((~mask).sum()), xxh64(np.asarray(slot_mapping, "<i8").tobytes(), 0),
                               int(offset)])
            return 1

        def poll(self, t):
            return True

    doc = json.load(open(os.path.join(GOLDEN, "adapter_flow_vectors.json")))
    assert "LMCacheConnectorV1Impl" in doc["source"] and len(doc["flows"]) >= 60
    n_meta = n_need = n_calls = 0
    for fl in doc["flows"]:
        bs, chunk = fl["block_size"], fl["chunk"]
        hits, serial, reqs = {}, {}, {}
        sched = SchedulerState(lambda toks: hits[int(toks[0])], bs, chunk, fl["discard_partial_chunks"], False, fl["kv_role"])
        rec = Recorder()
        worker = WorkerState(rec, bs, chunk, fl["kv_role"])
        for si, st in enumerate(fl["steps"]):
            new = []
            for n in st["new"]:
                tag = serial.setdefault(n["rid"], 100000 + len(serial))
                toks = [tag] + list(range(1, n["prompt_len"] + n["n_decode"]))
                hits[tag] = n["hit"]
                req = NS(request_id=n["rid"], prompt_token_ids=toks[: n["prompt_len"]], num_tokens=n["prompt_len"],
                         all_token_ids=toks)
                reqs[n["rid"]] = req
                need = sched.num_new_matched_tokens(n["rid"], req.prompt_token_ids, n["prompt_len"], n["num_computed_before"])
                assert need == n["need"], (fl["id"], si, n["rid"])
                if fl["kv_role"] != "kv_producer":       # the oracle's restatement of the same arithmetic
                    assert ko.num_new_matched_tokens(n["hit"], n["num_computed_before"], n["prompt_len"]) == n["need"]
                n_need += need > 0
                sched.after_alloc(req, need)
                new.append(NS(req_id=n["rid"], prompt_token_ids=req.prompt_token_ids, block_ids=(list(n["blocks"]),),
                              num_computed_tokens=n["num_computed"], sampling_params=None))
            cached = NS(req_ids=[c["rid"] for c in st["cached"]],
                        new_block_ids=[(list(c["new_blocks"]),) if c["new_blocks"] else None for c in st["cached"]],
                        resumed_req_ids=set(), all_token_ids={})
            num_sched = {n["rid"]: n["n_sched"] for n in st["new"]}
            num_sched.update({c["rid"]: c["n_sched"] for c in st["cached"]})
            metas = sched.build_meta(NS(scheduled_new_reqs=new, scheduled_cached_reqs=cached, num_scheduled_tokens=num_sched,
                                        finished_req_ids=set(st["finished"])))
            assert [m.req_id for m in metas] == [w["rid"] for w in st["metas"]], (fl["id"], si)
            for m, w in zip(metas, st["metas"]):
                where = (fl["id"], si, m.req_id)
                n_meta += 1
                assert m.is_last_prefill == w["is_last_prefill"], where
                assert [m.save_spec.skip_leading_tokens, m.save_spec.can_save] == w["save"], where
                got_load = None if m.load_spec is None else [m.load_spec.vllm_cached_tokens, m.load_spec.external_cached_tokens]
                assert got_load == w["load"], where
                n_ref = w["n_tokens"]
                assert len(m.token_ids) == n_ref if w["load"] is None else len(m.token_ids) >= n_ref, where
                sm = m.slot_mapping(bs)[:n_ref].astype("<i8")
                assert len(sm) == w["slot_mapping_len"] and xxh64(sm.tobytes(), 0) == w["slot_mapping_xxh64"], where
            # the worker half: start_load_kv (:798-905) and wait_for_save (:1033-1128) must tell the engine the
            # same thing (token count, masked prefix, slot mapping, offset).  The reference also issues stores
            # whose every token is masked out (offset == length): no-ops, which this build does not issue.
            if st["engine_calls"] is not None:
                rec.calls = []
                worker.start_load(metas)
                worker.save(metas)
                want_calls = [c for c in st["engine_calls"] if not (c[0] == "store" and c[4] >= c[1])]
                assert rec.calls == want_calls, (fl["id"], si)
                n_calls += len(want_calls)
    assert n_meta >= 160 and n_need > 10 and n_calls > 50

"""Golden traces of the save/load PLANNING rules, produced by executing vLLM's
vendored LMCache adapter itself —

    vllm/distributed/kv_transfer/kv_connector/v1/lmcache_integration/vllm_v1_adapter.py
      RequestTracker (:120-245)  and  ReqMeta.from_request_tracker (:270-399)

— the file SURVEY.md §8c names as *the* executable spec of this path.  The `lmcache` wheel it imports
at module level is absent from this image; those imports are satisfied by empty stand-in modules
(nothing of them is executed by the two classes above, which are plain Python + torch).

    python tests/golden/make_adapter_golden.py        # writes adapter_plan_vectors.json

Each scenario is one request stepped through the scheduler: a first step (new request, optionally with
an LMCache hit / vLLM prefix hit) and then further steps (chunked prefill, decode).  Per step the
fixture records what the reference produced: meta or None, token count, slot mapping, is_last_prefill,
SaveSpec, LoadSpec and the tracker's num_saved_tokens afterwards (slot mappings as length + XXH64 of
their little-endian int64 bytes, to keep the fixture small).
"""
import importlib.abc
import importlib.machinery
import json
import logging
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import xxhash

HERE = os.path.dirname(os.path.abspath(__file__))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return MagicMock(name=f"{self.__name__}.{name}")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == "lmcache" or fullname.startswith("lmcache."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        if module.__name__ == "lmcache.utils":
            module._lmcache_nvtx_annotate = lambda f: f
        if module.__name__ == "lmcache.logging":
            module.init_logger = lambda name: logging.getLogger(name)


def main():
    sys.meta_path.insert(0, _Finder())
    from vllm.distributed.kv_transfer.kv_connector.v1.lmcache_integration import vllm_v1_adapter as A
    rng = np.random.default_rng(20260921)
    scenarios = []
    for sid in range(int(os.environ.get("GOLDEN_SCENARIOS", "160"))):
        bs = 16
        chunk = int(rng.choice([64, 256]))
        prompt_len = int(rng.choice([1, 15, 16, chunk - 1, chunk, chunk + 1, 2 * chunk, 2 * chunk + 37,
                                     int(rng.integers(1, 5 * chunk))]))
        discard = bool(rng.integers(0, 2))
        save_decode = bool(rng.integers(0, 4) == 0)
        skip_save = bool(rng.integers(0, 8) == 0)
        # LMCache hit (whole chunks, maybe everything) and vLLM prefix hit (block aligned, <= lmcache hit)
        lm_hit = int(rng.choice([0, 0, (prompt_len // chunk) * chunk, int(rng.integers(0, prompt_len // chunk + 1)) * chunk]))
        vllm_hit = int(rng.integers(0, lm_hit // bs + 1)) * bs if lm_hit else 0
        can_load = bool(lm_hit > vllm_hit and rng.integers(0, 5) != 0)
        # prefill schedule: whole prompt at once or in pieces
        computed = max(vllm_hit, lm_hit if can_load else vllm_hit)
        computed = min(computed, prompt_len - 1) if prompt_len > 1 else 0
        remaining = prompt_len - computed
        pieces = []
        while remaining > 0:
            p = remaining if rng.integers(0, 2) else int(rng.integers(1, remaining + 1))
            pieces.append(p)
            remaining -= p
        n_decode = int(rng.integers(0, 4))
        blocks = [int(b) for b in rng.permutation(4096)[: (prompt_len + n_decode + bs - 1) // bs + 2]]
        tokens = [int(t) for t in rng.integers(0, 32000, prompt_len + n_decode)]

        steps = []
        n_first = computed + pieces[0]
        nblk_first = (n_first + bs - 1) // bs
        new_req = types.SimpleNamespace(req_id=f"s{sid}", prompt_token_ids=tokens[:prompt_len],
                                        block_ids=[blocks[:nblk_first]], sampling_params=None, mm_features=None,
                                        mm_hashes=None, mm_positions=None, mm_kwargs=None)
        tracker = A.RequestTracker.from_new_request(None, new_req, n_first, lm_hit if can_load else 0, skip_save)
        load_spec = A.LoadSpec(vllm_cached_tokens=vllm_hit, lmcache_cached_tokens=lm_hit, can_load=can_load) \
            if lm_hit > vllm_hit else None

        def record(kind, new_tokens, new_blocks, ls):
            m = A.ReqMeta.from_request_tracker(tracker, bs, lmcache_chunk_size=chunk, load_spec=ls,
                                               discard_partial_chunks=discard, save_decode_cache=save_decode)
            out = {"kind": kind, "new_tokens": new_tokens, "new_blocks": new_blocks, "meta": None,
                   "num_saved_tokens_after": tracker.num_saved_tokens}
            if m is not None:
                sm = np.asarray(m.slot_mapping.tolist(), dtype="<i8")
                out["meta"] = {"n_tokens": len(m.token_ids), "slot_mapping_xxh64": xxhash.xxh64(sm.tobytes()).intdigest(),
                               "slot_mapping_len": int(len(sm)),
                               "is_last_prefill": bool(m.is_last_prefill),
                               "save": [int(m.save_spec.skip_leading_tokens), bool(m.save_spec.can_save)],
                               "load": None if m.load_spec is None else
                               [int(m.load_spec.vllm_cached_tokens), int(m.load_spec.lmcache_cached_tokens)]}
            steps.append(out)

        record("new", n_first, blocks[:nblk_first], load_spec)
        have_tok, have_blk = n_first, nblk_first
        for p in pieces[1:] + [1] * n_decode:
            need_blk = (have_tok + p + bs - 1) // bs
            nb = blocks[have_blk:need_blk]
            tracker.update(tokens[have_tok:have_tok + p], (nb,) if nb else None)
            have_tok, have_blk = have_tok + p, max(have_blk, need_blk)
            record("cached", p, nb, None)
        scenarios.append({"id": sid, "block_size": bs, "chunk": chunk, "prompt_len": prompt_len, "n_decode": n_decode,
                          "discard_partial_chunks": discard, "save_decode_cache": save_decode, "skip_save": skip_save,
                          "lmcache_hit": lm_hit, "vllm_hit": vllm_hit, "can_load": can_load,
                          "first_tokens": n_first, "steps": steps})
    import vllm
    doc = {"source": "vllm %s lmcache_integration/vllm_v1_adapter.py RequestTracker + ReqMeta.from_request_tracker, "
                     "executed by tests/golden/make_adapter_golden.py" % vllm.__version__, "scenarios": scenarios}
    with open(os.path.join(os.environ.get("GOLDEN_OUT", HERE), "adapter_plan_vectors.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    n_steps = sum(len(s["steps"]) for s in scenarios)
    n_meta = sum(st["meta"] is not None for s in scenarios for st in s["steps"])
    print(f"{len(scenarios)} scenarios, {n_steps} steps, {n_meta} with a meta")


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


# ------------------------------------------------------------------------------------------------
# Second fixture: the scheduler half of the connector as a whole —
#   LMCacheConnectorV1Impl.get_num_new_matched_tokens (:1141-1228), update_state_after_alloc
#   (:1231-1293), build_connector_meta (:1296-1407) — driven with fake Requests / SchedulerOutputs
#   over several interleaved requests.  The instance is created without __init__ (which would build
#   an LMCache engine) and given exactly the attributes those three methods read; its lookup client
#   answers from a scripted table.            python tests/golden/make_adapter_golden.py flows
# ------------------------------------------------------------------------------------------------
def main_flows():
    sys.meta_path.insert(0, _Finder())
    from vllm.distributed.kv_transfer.kv_connector.v1.lmcache_integration import vllm_v1_adapter as A
    rng = np.random.default_rng(4242)
    NS = types.SimpleNamespace
    import copy

    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self       # the worker methods move slot mappings to the GPU
    flows = []
    for fid in range(int(os.environ.get("GOLDEN_FLOWS", "60"))):
        bs = 16
        chunk = int(rng.choice([64, 256]))
        discard = bool(rng.integers(0, 2))
        kv_role = str(rng.choice(["kv_both", "kv_both", "kv_both", "kv_consumer", "kv_producer"]))
        hits: dict[str, int] = {}

        class Lookup:
            def lookup(self, token_ids, lookup_id=None, request_configs=None):
                return hits[tuple(token_ids[:4]).__repr__()]

            def clear_lookup_status(self, rid):
                pass

        impl = object.__new__(A.LMCacheConnectorV1Impl)
        impl.kv_role = kv_role
        impl.lookup_client = Lookup()
        impl._requests_priority, impl.load_specs, impl._unfinished_requests, impl._request_trackers = {}, {}, {}, {}
        impl.skip_last_n_tokens, impl.async_loading, impl._lookup_requests_in_step = 0, False, []
        impl.force_skip_save, impl.config = False, NS(priority_limit=None)
        impl._block_size, impl._lmcache_chunk_size = bs, chunk
        impl._discard_partial_chunks, impl._save_decode_cache = discard, False

        # plan the requests of this flow
        reqs = []
        next_block = 0
        for r in range(int(rng.integers(1, 4))):
            plen = int(rng.choice([chunk, chunk + 5, 2 * chunk, 3 * chunk + 17, int(rng.integers(2, 4 * chunk))]))
            n_dec = int(rng.integers(0, 3))
            toks = [int(t) for t in rng.integers(0, 32000, plen + n_dec)]
            hit = int(rng.choice([0, plen, (plen // chunk) * chunk, int(rng.integers(0, plen // chunk + 1)) * chunk]))
            if hit == plen and plen % chunk and discard:
                hit = (plen // chunk) * chunk          # a partial tail is never stored when partial chunks are discarded
            vhit = int(rng.integers(0, hit // bs + 1)) * bs if rng.integers(0, 2) else 0
            vhit = min(vhit, max(plen - 1, 0) // bs * bs)
            nblk = (plen + n_dec + bs - 1) // bs + 1
            blocks = list(range(next_block, next_block + nblk))
            next_block += nblk
            reqs.append(dict(rid=f"f{fid}r{r}", plen=plen, n_dec=n_dec, toks=toks, hit=hit, vhit=vhit, blocks=blocks,
                             start=int(rng.integers(0, 3)), done=0, have_blk=0, state="waiting"))
            hits[tuple(toks[:4]).__repr__()] = hit
        steps = []
        for step in range(40):
            new, cached_ids, cached_blocks, num_sched, finished = [], [], [], {}, []
            step_rec = {"new": [], "cached": [], "finished": []}
            for q in reqs:
                if q["state"] == "finished_pending":
                    finished.append(q["rid"])
                    q["state"] = "gone"
            for q in reqs:
                if q["state"] == "waiting" and step >= q["start"]:
                    request = NS(request_id=q["rid"], prompt_token_ids=q["toks"][:q["plen"]], num_tokens=q["plen"],
                                 all_token_ids=q["toks"], priority=0, sampling_params=None, mm_features=None,
                                 kv_transfer_params=None)
                    q["request"] = request
                    need = impl.get_num_new_matched_tokens(request, q["vhit"])
                    impl.update_state_after_alloc(request, need)
                    computed = q["vhit"] + need
                    n_sched = q["plen"] - computed if rng.integers(0, 2) else int(rng.integers(1, q["plen"] - computed + 1))
                    nb = (computed + n_sched + bs - 1) // bs
                    new.append(NS(req_id=q["rid"], prompt_token_ids=q["toks"][:q["plen"]], block_ids=[q["blocks"][:nb]],
                                  num_computed_tokens=computed, sampling_params=None, mm_features=None))
                    num_sched[q["rid"]] = n_sched
                    q.update(done=computed + n_sched, have_blk=nb, state="running")
                    step_rec["new"].append({"rid": q["rid"], "prompt_len": q["plen"], "n_decode": q["n_dec"], "hit": q["hit"],
                                            "num_computed_before": q["vhit"], "need": need, "num_computed": computed,
                                            "n_sched": n_sched, "blocks": q["blocks"][:nb]})
                elif q["state"] == "running":
                    total = q["plen"] + q["n_dec"]
                    if q["done"] >= total:
                        q["state"] = "finished_pending"
                        continue
                    left_prefill = q["plen"] - q["done"]
                    n_sched = 1 if left_prefill <= 0 else (left_prefill if rng.integers(0, 2) else int(rng.integers(1, left_prefill + 1)))
                    nb = (q["done"] + n_sched + bs - 1) // bs
                    newb = q["blocks"][q["have_blk"]:nb]
                    cached_ids.append(q["rid"])
                    cached_blocks.append((newb,) if newb else None)
                    num_sched[q["rid"]] = n_sched
                    q.update(done=q["done"] + n_sched, have_blk=max(nb, q["have_blk"]))
                    step_rec["cached"].append({"rid": q["rid"], "n_sched": n_sched, "new_blocks": newb})
            step_rec["finished"] = finished
            so = NS(finished_req_ids=set(finished), scheduled_new_reqs=new, num_scheduled_tokens=num_sched,
                    scheduled_cached_reqs=NS(req_ids=cached_ids, new_block_ids=cached_blocks))
            meta = impl.build_connector_meta(so)
            step_rec["metas"] = [{
                "rid": m.req_id, "n_tokens": len(m.token_ids), "is_last_prefill": bool(m.is_last_prefill),
                "save": [int(m.save_spec.skip_leading_tokens), bool(m.save_spec.can_save)],
                "load": None if m.load_spec is None else [int(m.load_spec.vllm_cached_tokens), int(m.load_spec.lmcache_cached_tokens)],
                "slot_mapping_len": int(len(m.slot_mapping)),
                "slot_mapping_xxh64": xxhash.xxh64(np.asarray(m.slot_mapping.tolist(), dtype="<i8").tobytes()).intdigest(),
            } for m in meta.requests]
            # ---- the worker half on a pickled copy of that metadata: which engine calls does
            # start_load_kv (:798-905) / wait_for_save (:1033-1128) make? -----------------------------------
            step_rec["engine_calls"] = None
            if kv_role != "kv_producer":            # a producer needs a DisaggSpec (NIXL push), not modelled here
                calls = []

                def digest(sm):
                    return xxhash.xxh64(np.asarray(sm.tolist(), dtype="<i8").tobytes()).intdigest()

                class Engine:
                    def post_init(self, **kw):
                        pass

                    def lookup_unpin(self, ids):
                        pass

                    def retrieve(self, tokens, mask, kvcaches=None, slot_mapping=None, **kw):
                        calls.append(["retrieve", len(tokens), int((~mask).sum()), digest(slot_mapping)])
                        return mask.clone()

                    def store(self, token_ids, mask=None, kvcaches=None, slot_mapping=None, offset=0, **kw):
                        calls.append(["store", len(token_ids), int((~mask).sum()), digest(slot_mapping), int(offset)])

                w = object.__new__(A.LMCacheConnectorV1Impl)
                wmeta = copy.deepcopy(meta)
                w.kv_role, w._lmcache_chunk_size = kv_role, chunk
                w.kv_caches = {"layer0": torch.zeros(1)}
                w._parent = NS(_get_connector_metadata=lambda: wmeta)
                w.lmcache_engine, w._stats_monitor = Engine(), MagicMock()
                w.use_layerwise, w.enable_blending, w.layerwise_storers, w.current_layer = False, False, [], 0
                w.start_load_kv(NS(attn_metadata=object()))
                w.wait_for_save()
                step_rec["engine_calls"] = calls
            steps.append(step_rec)
            if all(q["state"] == "gone" for q in reqs):
                break
        flows.append({"id": fid, "block_size": bs, "chunk": chunk, "discard_partial_chunks": discard, "kv_role": kv_role,
                      "steps": steps})
    import vllm
    doc = {"source": "vllm %s lmcache_integration/vllm_v1_adapter.py LMCacheConnectorV1Impl.get_num_new_matched_tokens / "
                     "update_state_after_alloc / build_connector_meta, executed by tests/golden/make_adapter_golden.py flows"
                     % vllm.__version__, "flows": flows}
    with open(os.path.join(os.environ.get("GOLDEN_OUT", HERE), "adapter_flow_vectors.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print(len(flows), "flows,", sum(len(f["steps"]) for f in flows), "steps,",
          sum(len(s["metas"]) for f in flows for s in f["steps"]), "metas")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "flows":
    main_flows()

"""`LMCacheConnectorV1Impl` as vLLM's built-in wrapper imports it
(vllm/distributed/kv_transfer/kv_connector/v1/lmcache_connector.py:105-113): ctor
(vllm_config, role, parent).  With this module on PYTHONPATH the chart's literal
`{"kv_connector":"LMCacheConnectorV1"}` (helm/templates/deployment-vllm-multi.yaml:198) loads the
b200kv engine — the chart-level drop-in of SURVEY.md §8b.  The wrapper forwards only a subset of
the plugin API and is not SupportsHMA, so `--disable-hybrid-kv-cache-manager` stays necessary on
that route; the native `B200KVConnector` entry point has neither limit.
"""
from __future__ import annotations

from b200kv.connector import B200KVConnector


class LMCacheConnectorV1Impl:
    def __init__(self, vllm_config, role, parent):
        self._parent = parent
        self._inner = B200KVConnector(vllm_config, role, getattr(parent, "_kv_cache_config", None))
        # The wrapper hands update_state_after_alloc no block ids (lmcache_connector.py:253-262).  What needs
        # them at allocation time is therefore declined on this route — disaggregated-prefill pulls and loads
        # detached from the forward step — and such requests are served by the pool lookup or recomputed,
        # instead of being promised tokens that have nowhere to land.
        if self._inner._pd is not None:
            self._inner._pd.blocks_known_at_alloc = False
        if self._inner._sched is not None:
            self._inner._sched.async_load = False
        self._inner.cfg.async_load = False

    def _sync_meta(self):
        self._inner._connector_metadata = self._parent._connector_metadata

    # worker side
    def register_kv_caches(self, kv_caches):
        return self._inner.register_kv_caches(kv_caches)

    def start_load_kv(self, forward_context, **kwargs):
        self._sync_meta()
        return self._inner.start_load_kv(forward_context, **kwargs)

    def wait_for_layer_load(self, layer_name):
        return self._inner.wait_for_layer_load(layer_name)   # layer-wise loads are the default

    def save_kv_layer(self, layer_name, kv_layer, attn_metadata, **kwargs):
        return None

    def wait_for_save(self):
        self._sync_meta()
        return self._inner.wait_for_save()

    def get_finished(self, finished_req_ids):
        return self._inner.get_finished(finished_req_ids)

    def get_block_ids_with_load_errors(self):
        return self._inner.get_block_ids_with_load_errors()

    def get_kv_events(self):
        return []

    def shutdown(self):
        return self._inner.shutdown()

    # scheduler side (the wrapper unpacks the tuple itself: lmcache_connector.py returns (n, False))
    def get_num_new_matched_tokens(self, request, num_computed_tokens):
        return self._inner.get_num_new_matched_tokens(request, num_computed_tokens)[0]

    def update_state_after_alloc(self, request, num_external_tokens):
        return self._inner.update_state_after_alloc(request, None, num_external_tokens)

    def build_connector_meta(self, scheduler_output):
        return self._inner.build_connector_meta(scheduler_output)

    def request_finished(self, request, block_ids):
        return self._inner.request_finished(request, block_ids)

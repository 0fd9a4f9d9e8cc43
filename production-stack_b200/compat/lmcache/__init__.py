"""`lmcache` compatibility shim provided by b200kv.

production-stack's unmodified router imports LMCache's controller classes by name
(src/vllm_router/routers/routing_logic.py:33-40) and the Helm chart selects the connector by the
name `LMCacheConnectorV1` (helm/templates/deployment-vllm-multi.yaml:198,204).  Putting
`production-stack_b200/compat` on PYTHONPATH makes both resolve to this engine — it is NOT the
LMCache project and implements only the surface production-stack touches (SURVEY.md §8b "second
boundary", §8f-1).
"""
__version__ = "0.0.0+b200kv.shim"

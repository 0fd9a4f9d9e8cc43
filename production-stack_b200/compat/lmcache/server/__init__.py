"""`lmcache.server` under LMCache's name: the cache server production-stack's chart starts as
`lmcache_server <host> <port>` (helm/templates/deployment-cache-server.yaml:62-65) — here the
b200kv server (production-stack_b200/b200kv/server.py)."""

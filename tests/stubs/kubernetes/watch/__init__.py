"""Test-only stub: the reference router imports `kubernetes` at module import time
(src/vllm_router/service_discovery.py:27); static discovery never calls into it."""


class _Missing:
    def __getattr__(self, name):
        raise RuntimeError("kubernetes is not available in this environment (test stub)")


def __getattr__(name):
    return _Missing()

"""The multimodal part of the chunk-key identity against vLLM's vendored LMCache helpers, EXECUTED here
(lmcache_integration/utils.py: extract_mm_features :169-209, apply_mm_hashes_to_token_ids :73-89 — what the
adapter calls before lookup :1168-1172 and before store :344-350).  The reference writes 16 bits of the item's hash
into the placeholder range; this repo writes 128 bits of its identifier, so the token streams differ by design.  What
must agree is the PARTITION they induce: two requests get the same LMCache token stream up to position p exactly when
they get the same b200kv key tokens up to p (barring the reference's own 16-bit collisions, excluded by construction
here), i.e. the same chunks are shared and the same chunks are kept apart."""
import importlib.abc
import importlib.machinery
import logging
import sys
import types
from types import SimpleNamespace as NS
from unittest.mock import MagicMock

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytest.importorskip("vllm")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return MagicMock(name=f"{self.__name__}.{name}")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """`lmcache` (absent wheel) as empty stand-ins: the two helpers under test are plain Python + torch."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname == "lmcache" or fullname.startswith("lmcache."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        if module.__name__ == "lmcache.logging":
            module.init_logger = lambda name: logging.getLogger(name)


@pytest.fixture(scope="module")
def ref_utils():
    finder = _Finder()
    saved = {k: v for k, v in sys.modules.items() if k == "lmcache" or k.startswith("lmcache.")}
    for k in saved:
        del sys.modules[k]
    sys.meta_path.insert(0, finder)
    try:
        from vllm.distributed.kv_transfer.kv_connector.v1.lmcache_integration import utils as U
        yield U
    finally:
        sys.meta_path.remove(finder)
        for k in [k for k in sys.modules if k == "lmcache" or k.startswith("lmcache.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_same_requests_are_shared_and_kept_apart_as_in_the_reference(ref_utils):
    from b200kv.adapter import request_identity
    rng = np.random.default_rng(5)
    n = 400
    base = rng.integers(10, 30000, n).astype(np.int64)
    # distinct in their low 16 bits as well, so that the reference itself tells them apart
    hashes = ["ab" * 30 + f"{(0x1000 + 7919 * i) & 0xFFFF:04x}" for i in range(6)]    # int(s, 16) & 0xFFFF = the last 4 digits
    reqs = []
    for i in range(40):
        items = []
        pos = int(rng.integers(0, 60))
        for _ in range(int(rng.integers(0, 3))):
            length = int(rng.integers(4, 40))
            if pos + length >= n:
                break
            items.append(NS(identifier=hashes[int(rng.integers(0, len(hashes)))], mm_position=NS(offset=pos, length=length)))
            pos += length + int(rng.integers(1, 80))
        reqs.append(NS(request_id=f"r{i}", prompt_token_ids=list(base), mm_features=items or None))
    ref_streams, our_streams = [], []
    for r in reqs:
        t = torch.tensor(r.prompt_token_ids)
        h, p = ref_utils.extract_mm_features(r)
        if h:
            ref_utils.apply_mm_hashes_to_token_ids(t, h, p)
        ref_streams.append(t.numpy())
        ident = request_identity(r)
        our_streams.append(np.asarray(r.prompt_token_ids, dtype=np.int32) if ident is None else ident.apply(r.prompt_token_ids))

    def common_prefix(a, b):
        d = np.flatnonzero(a != b)
        return len(a) if len(d) == 0 else int(d[0])

    checked = 0
    for i in range(len(reqs)):
        for j in range(i + 1, len(reqs)):
            assert common_prefix(ref_streams[i], ref_streams[j]) == common_prefix(our_streams[i], our_streams[j]), (i, j)
            checked += 1
    assert checked == 40 * 39 // 2
    # and a text-only request is left exactly as it is by both
    plain = NS(request_id="p", prompt_token_ids=list(base), mm_features=None)
    assert request_identity(plain) is None and ref_utils.extract_mm_features(plain) == ([], [])

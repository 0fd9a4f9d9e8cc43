"""Test-only stand-in for the `uhashring` wheel (absent offline) so the UNMODIFIED reference router
can be imported in plumbing tests.  Consistent hashing with virtual nodes; only the calls the
router makes (src/vllm_router/routers/routing_logic.py:213-249) are provided."""
import bisect
import hashlib


class HashRing:
    def __init__(self, nodes=None, vnodes=160, **kw):
        self._vn = vnodes
        self._ring, self._keys, self._nodes = {}, [], set()
        for n in nodes or []:
            self.add_node(n)

    @staticmethod
    def _h(s):
        return int(hashlib.md5(str(s).encode()).hexdigest()[:16], 16)

    def add_node(self, node, conf=None):
        if node in self._nodes:
            return
        self._nodes.add(node)
        for i in range(self._vn):
            k = self._h(f"{node}-{i}")
            self._ring[k] = node
            bisect.insort(self._keys, k)

    def remove_node(self, node):
        if node not in self._nodes:
            return
        self._nodes.discard(node)
        for i in range(self._vn):
            k = self._h(f"{node}-{i}")
            if self._ring.pop(k, None) is not None:
                self._keys.remove(k)

    def get_nodes(self):
        return list(self._nodes)

    def get_node(self, key):
        if not self._keys:
            return None
        i = bisect.bisect(self._keys, self._h(key)) % len(self._keys)
        return self._ring[self._keys[i]]

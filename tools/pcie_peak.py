"""Measured host-link ceilings for the e2e roofline: pinned H2D, D2H and both at once, for
cudaHostAlloc memory (torch pinned) and for a POSIX-shm segment pinned with cudaHostRegister
(what the pool uses).  Prints one JSON line."""
import json
import mmap
import os

import torch

GiB = 1 << 30
n = GiB
dev = torch.device("cuda:0")
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def bw(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure(h_a, h_b, tag):
    out = {}

    def h2d():
        d_a.copy_(h_a, non_blocking=True)

    def d2h():
        h_b.copy_(d_b, non_blocking=True)

    def both():
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(s1):
            s1.wait_event(ev)
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            s2.wait_event(ev)
            h_b.copy_(d_b, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)

    out[f"{tag}_h2d_GBps"] = n / bw(h2d) / 1e6
    out[f"{tag}_d2h_GBps"] = n / bw(d2h) / 1e6
    out[f"{tag}_duplex_total_GBps"] = 2 * n / bw(both) / 1e6
    return out


res = {}
res.update(measure(torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(), "hostalloc"))
# shm + cudaHostRegister
fd = os.open("/dev/shm/b200kv-pcie-probe", os.O_CREAT | os.O_RDWR, 0o600)
os.ftruncate(fd, 2 * n)
mm = mmap.mmap(fd, 2 * n)
t = torch.frombuffer(mm, dtype=torch.uint8)
t.fill_(1)
rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), 2 * n, 1)
res["hostregister_rc"] = int(rc)
res.update(measure(t[:n], t[n:], "shmregistered"))
torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
os.unlink("/dev/shm/b200kv-pcie-probe")
print(json.dumps(res))

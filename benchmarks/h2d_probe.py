import torch, time
dev = torch.device("cuda:0")
host = torch.empty((64, 768, 1024, 3), dtype=torch.uint8).pin_memory()
print("view pinned:", host[3].is_pinned())
d = torch.empty_like(host, device=dev)
s = torch.cuda.Stream()
def timeit(fn, n=5):
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        t=time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter()-t)
    return min(ts)*1e3
print("one 151MB copy ms:", timeit(lambda: d.copy_(host, non_blocking=True)))
def per_image():
    for i in range(64): d[i].copy_(host[i], non_blocking=True)
print("64 x 2.36MB copies ms:", timeit(per_image))
import sys; sys.path.insert(0, "/root/repo")
import fsd_b200
from fsd_b200 import ops
pool = ops.ImagePool(64, 768, 1024, dev)
def up():
    for i in range(64): pool.upload(i, host[i], non_blocking=True)
print("pool.upload x64 ms:", timeit(up))
ph = torch.empty((64,768,1024,3),dtype=torch.uint8)
print("pageable one copy ms:", timeit(lambda: d.copy_(ph, non_blocking=True)))

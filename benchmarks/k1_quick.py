import torch, numpy as np, time
import fsd_b200.ops as ops
dev = torch.device("cuda:0")
N=32
pool = ops.ImagePool(N, 768, 1024, dev)
pool.buf.random_(0,256)
boxes=[[0,0],[410,0],[512,0],[0,256],[410,256],[512,256]]
ent = torch.tensor([[i,x,y] for i in range(N) for x,y in boxes], dtype=torch.int32, device=dev)
entf = torch.tensor([[i,0,0] for i in range(N)], dtype=torch.int32, device=dev)
out = torch.empty((N*6,3,1024,1024), dtype=torch.float16, device=dev)
outf = torch.empty((N,3,768,1024), dtype=torch.float16, device=dev)
for _ in range(3):
    ops.gather_letterbox(pool, ent, 512,512, out=out); ops.gather_letterbox(pool, entf, 1024,768, out=outf)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e2=torch.cuda.Event(enable_timing=True)
ts=[];tf=[]
for _ in range(10):
    e0.record(); ops.gather_letterbox(pool, ent, 512,512, out=out); e1.record(); ops.gather_letterbox(pool, entf, 1024,768, out=outf); e2.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); tf.append(e1.elapsed_time(e2))
bytes_s = N*(6*3*1024*1024*2) + N*768*1024*3
bytes_f = N*(768*1024*3*2 + 768*1024*3)
print("slices: ms", min(ts), "GB/s", bytes_s/min(ts)/1e6, " full: ms", min(tf), "GB/s", bytes_f/min(tf)/1e6)

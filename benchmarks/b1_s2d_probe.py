"""Probe: layer 1 of YOLO11n (Conv 16->32, 3x3, stride 2 on [E,16,512,512]) as cuDNN runs it today vs the algebraically
identical 2x2 stride-1 convolution on the space-to-depth tensor [E,64,257,257] (zero top row / left column)."""
import torch, time
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
E = 96
x = torch.randn((E, 16, 512, 512), device=dev).half().contiguous(memory_format=torch.channels_last)
w = (torch.randn((32, 16, 3, 3), device=dev) * 0.1).half().contiguous(memory_format=torch.channels_last)
s = torch.zeros((E, 64, 257, 257), device=dev).half().contiguous(memory_format=torch.channels_last)
# s[e, (dy*2+dx)*16+c, Y+1, X+1] = x[e, c, 2Y+dy, 2X+dx]
for dy in range(2):
    for dx in range(2):
        s[:, (dy * 2 + dx) * 16:(dy * 2 + dx + 1) * 16, 1:, 1:] = x[:, :, dy::2, dx::2]
w2 = torch.zeros((32, 64, 2, 2), device=dev).half()
kmap = {0: (0, 1), 1: (1, 0), 2: (1, 1)}  # k -> (a, d)
for ky in range(3):
    a, dy = kmap[ky]
    for kx in range(3):
        b, dx = kmap[kx]
        w2[:, (dy * 2 + dx) * 16:(dy * 2 + dx + 1) * 16, a, b] = w[:, :, ky, kx]
w2 = w2.contiguous(memory_format=torch.channels_last)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
ref = torch.nn.functional.conv2d(x, w, None, 2, 1)
alt = torch.nn.functional.conv2d(s, w2, None, 1, 0)
print("shapes", tuple(ref.shape), tuple(alt.shape), "max abs diff", float((ref.float()-alt.float()).abs().max()), "ref max", float(ref.abs().max()))
print("3x3 s2 on [E,16,512,512]  ms:", t(lambda: torch.nn.functional.conv2d(x, w, None, 2, 1)))
print("2x2 s1 on [E,64,257,257]  ms:", t(lambda: torch.nn.functional.conv2d(s, w2, None, 1, 0)))

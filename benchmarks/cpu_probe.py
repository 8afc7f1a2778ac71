import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import fsd_b200
from fsd_b200.backbones.yolo11_pose import build_yolo11n_pose
from fsd_b200.synthetic import make_image
from oracle.yolo11_pose_plain import build_plain_yolo11n_pose
from oracle import letterbox as olb
torch.set_num_threads(os.cpu_count())
print("threads", torch.get_num_threads(), "cpus", os.cpu_count())
a = build_yolo11n_pose(); b = build_plain_yolo11n_pose(); c = build_plain_yolo11n_pose(state_dict=a.state_dict())
img = make_image(0, 768, 1024)[0]
x = olb.preprocess(np.ascontiguousarray(img[:512, :512]), imgsz=1024, half=False).float()
xr = torch.rand(1, 3, 1024, 1024)
for flush in (False, True):
    torch.set_flush_denormal(flush)
    for name, m in (("product", a), ("plain-own-init", b), ("plain-product-weights", c)):
        for tag, inp in (("image", x), ("rand", xr)):
            with torch.no_grad():
                m(inp)
                t = time.time()
                for _ in range(3): m(inp)
                print(f"flush={flush} {name:22s} {tag:6s} {(time.time()-t)/3*1e3:8.1f} ms", flush=True)

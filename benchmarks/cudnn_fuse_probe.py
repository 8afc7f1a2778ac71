"""Probe: cuDNN convolution + fsd_bias_act (two passes, today's path) against ONE cuDNN runtime-fusion graph
(conv -> bias -> SiLU [-> + residual]) built through the cudnn frontend, for every convolution shape the YOLO11n-pose
backbone sends down the library path at a 1024x1024 network input.  Prints one JSON line per shape and a total.

    python benchmarks/cudnn_fuse_probe.py [N]      (N = network inputs per call, default 32)
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

import fsd_b200  # noqa: F401
from fsd_b200 import ops
from fsd_b200.backbones.yolo11_pose import Conv, YOLO11Pose


def shapes():
    m = YOLO11Pose().eval()
    rows = []

    def hook(mod, inp, out):
        c, x = mod.conv, inp[0]
        rows.append((c.in_channels, c.out_channels, c.kernel_size[0], c.stride[0], c.groups, x.shape[2], x.shape[3],
                     isinstance(mod.act, torch.nn.SiLU)))

    for mod in m.modules():
        if isinstance(mod, Conv):
            mod.register_forward_hook(hook)
    with torch.no_grad():
        m(torch.zeros(1, 3, 256, 256))
    out = {}
    for cin, cout, k, s, g, h, w, act in rows:
        h, w = h * 4, w * 4
        if cin == 3 or (k == 1 and cin <= 64 and cout <= 64 and h * w >= 1024):
            continue  # the stem and the small 1x1 layers have their own kernels
        key = (cin, cout, k, s, g, h, w, act)
        out[key] = out.get(key, 0) + 1
    return out


def time_graph(fn, iters=10, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (iters * reps)


def build_fused(handle, x, w, b, out, act, k, s, g_):
    import cudnn

    graph = cudnn.pygraph(handle=handle, io_data_type=cudnn.data_type.HALF, intermediate_data_type=cudnn.data_type.FLOAT,
                          compute_data_type=cudnn.data_type.FLOAT)
    X, W, B = graph.tensor_like(x), graph.tensor_like(w), graph.tensor_like(b)
    y = graph.conv_fprop(image=X, weight=W, padding=[k // 2, k // 2], stride=[s, s], dilation=[1, 1])
    y = graph.bias(input=y, bias=B)
    if act == 1:  # (frontend 1.18: the keyword names of swish's beta / compute type are swapped in the binding)
        y = graph.swish(input=y, swish_beta=cudnn.data_type.FLOAT, compute_data_type=1.0)
    elif act == 2:
        sg = graph.sigmoid(input=y)
        y = graph.mul(a=y, b=sg)
    y.set_output(True).set_data_type(cudnn.data_type.HALF).set_dim(list(out.shape)).set_stride(list(out.stride()))
    graph.validate()
    graph.build_operation_graph()
    graph.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK])
    graph.check_support()
    graph.build_plans(cudnn.build_plan_policy.HEURISTICS_CHOICE)
    ws = torch.empty(max(1, graph.get_workspace_size()), dtype=torch.uint8, device=x.device)
    pack = {X: x, W: w, B: b, y: out}
    return graph, pack, ws


def main():
    import cudnn

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    handle = cudnn.create_handle()
    print(json.dumps({"cudnn_backend": cudnn.backend_version(), "frontend": cudnn.__version__, "inputs": n}), flush=True)
    tot_now = tot_fused = 0.0
    with torch.no_grad(), ops.cudnn_benchmark():
        for (cin, cout, k, s, g_, h, w, act), count in sorted(shapes().items(), key=lambda kv: -kv[0][5] * kv[0][6] * kv[0][1]):
            x = torch.randn((n, cin, h, w), device=dev, dtype=torch.half).contiguous(memory_format=torch.channels_last)
            wt = (torch.randn((cout, cin // g_, k, k), device=dev, dtype=torch.half) * 0.05).contiguous(memory_format=torch.channels_last)
            b = torch.randn(cout, device=dev, dtype=torch.half) * 0.1
            row = {"cin": cin, "cout": cout, "k": k, "s": s, "groups": g_, "h": h, "w": w, "act": int(act), "count": count}

            def now():
                y = F.conv2d(x, wt, None, s, k // 2, 1, g_)
                return ops.bias_act(y, b, "silu" if act else "none")

            ref = now()
            row["now_us"] = time_graph(now)
            out = torch.empty_like(ref)
            try:
                stream = torch.cuda.current_stream().cuda_stream
                cudnn.set_stream(handle=handle, stream=stream)
                try:
                    graph, pack, ws = build_fused(handle, x, wt, b.view(1, -1, 1, 1), out, 1 if act else 0, k, s, g_)
                    row["act_form"] = "swish" if act else "none"
                except Exception as e1:
                    if not act:
                        raise
                    row["swish_error"] = str(e1)[-120:]
                    graph, pack, ws = build_fused(handle, x, wt, b.view(1, -1, 1, 1), out, 2, k, s, g_)
                    row["act_form"] = "sigmoid*mul"

                def fused():
                    cudnn.set_stream(handle=handle, stream=torch.cuda.current_stream().cuda_stream)
                    graph.execute(pack, ws, handle=handle)

                fused()
                torch.cuda.synchronize()
                row["max_abs_diff"] = float((out.float() - ref.float()).abs().max())
                row["fused_us"] = time_graph(fused)
                try:
                    row["plan"] = graph.get_plan_name_at_index(0)[:60]
                except Exception:
                    pass
            except Exception as e:  # unsupported pattern / engine: the two-pass path stays
                row["fused_error"] = str(e)[:160]
                row["fused_us"] = None
            bytes_ = (x.numel() + ref.numel()) * 2
            row["now_gbs"] = bytes_ / row["now_us"] / 1e3
            if row["fused_us"]:
                row["fused_gbs"] = bytes_ / row["fused_us"] / 1e3
            tot_now += row["now_us"] * count
            tot_fused += (row["fused_us"] if row["fused_us"] and row["fused_us"] < row["now_us"] else row["now_us"]) * count
            print(json.dumps(row), flush=True)
    print(json.dumps({"total_now_ms": tot_now / 1e3, "total_best_of_both_ms": tot_fused / 1e3, "inputs": n}), flush=True)


if __name__ == "__main__":
    main()

"""Weight loading for the two PyTorch backbones, with the reference's failure behaviour: a checkpoint is loaded or the
call raises — nothing continues on random weights unless the caller asks for it (`allow_random_init=True`, tests/bench).

The reference loads its detector with `ultralytics.YOLO(model_path)` (utils/yolo_wrapper.py:55; checkpoints such as
models/yolo11s-pose-default/.../best.pt, eval/eval_official_widerface.py:48) and its enhancer through
`RealESRGANer(model_path=...)` (utils/enhancer.py:131-156).  ultralytics is not installed here, and its `best.pt` is a
pickle of the whole `PoseModel` object graph.  `read_ultralytics_checkpoint` reads it WITHOUT ultralytics and without
executing pickled code: it lists the pickle's globals, refuses anything outside `ultralytics.*` / `torch.nn.modules.*`,
maps those to inert stub classes for torch's weights-only unpickler, and walks the resulting object graph for tensors.
Conv+BatchNorm pairs (ultralytics `Conv`: conv bias-free + bn, eps 1e-3) are folded into the fused `Conv2d(bias)` layers
backbones/yolo11_pose.py is built from, by module index (`model.N.` -> b0..b10, h13..h22, head).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Tuple

import torch

# ultralytics yolo11-pose.yaml layer index -> attribute of backbones.yolo11_pose.YOLO11Pose (11/12/14/15/18/21 are
# Upsample / Concat: no parameters)
_LAYER_NAMES = {0: "b0", 1: "b1", 2: "b2", 3: "b3", 4: "b4", 5: "b5", 6: "b6", 7: "b7", 8: "b8", 9: "b9", 10: "b10",
                13: "h13", 16: "h16", 17: "h17", 19: "h19", 20: "h20", 22: "h22", 23: "head"}
_ALLOWED_PREFIXES = ("ultralytics.", "torch.nn.modules.")
# scale letter -> (depth, width, max_channels, c3k everywhere) of yolo11.yaml
_SCALES = {"n": (0.50, 0.25, 1024, False), "s": (0.50, 0.50, 1024, False), "m": (0.50, 1.00, 512, True),
           "l": (1.00, 1.00, 512, True), "x": (1.00, 1.50, 512, True)}


class CheckpointError(ValueError):
    pass


def _stub(full_name: str):
    module, _, name = full_name.rpartition(".")
    return type(name, (object,), {"__module__": module, "__doc__": "inert stand-in: holds the pickled attributes only"})


def read_ultralytics_checkpoint(path: str):
    """-> the unpickled checkpoint dict whose module objects are inert stubs (attributes only, no code)."""
    names = torch.serialization.get_unsafe_globals_in_checkpoint(path)
    foreign = [n for n in names if not n.startswith(_ALLOWED_PREFIXES)]
    if foreign:
        raise CheckpointError(f"{path}: refusing to unpickle globals outside ultralytics.* / torch.nn.modules.*: {foreign[:5]}")
    stubs = [_stub(n) for n in names]
    with torch.serialization.safe_globals(stubs):
        return torch.load(path, map_location="cpu", weights_only=True)


def _is_module(obj) -> bool:
    return hasattr(obj, "_modules") and hasattr(obj, "_parameters")


def _harvest(obj, prefix: str, out: Dict[str, torch.Tensor], kinds: Dict[str, str]):
    """state_dict of a stub module graph (parameters + buffers) and the class name of every sub-module."""
    kinds[prefix.rstrip(".")] = type(obj).__name__
    for k, v in list(getattr(obj, "_parameters", {}).items()) + list(getattr(obj, "_buffers", {}).items()):
        if v is not None:
            out[prefix + k] = (v.data if isinstance(v, torch.nn.Parameter) else v).detach().float()
    for k, m in getattr(obj, "_modules", {}).items():
        if m is not None and _is_module(m):
            _harvest(m, f"{prefix}{k}.", out, kinds)
    if type(obj).__name__ == "BatchNorm2d":
        out[prefix + "eps"] = torch.tensor(float(getattr(obj, "eps", 1e-3)))


def fold_conv_bn(w: torch.Tensor, b, gamma, beta, mean, var, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Conv2d(bias=b) followed by BatchNorm2d in eval mode == Conv2d(w', b') (ultralytics fuse_conv_and_bn)."""
    scale = gamma / torch.sqrt(var + eps)
    w2 = w * scale.view(-1, 1, 1, 1)
    b2 = beta + ((b if b is not None else torch.zeros_like(mean)) - mean) * scale
    return w2, b2


def fused_state_dict(flat: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """ultralytics names (`model.N.….conv.weight` + `.bn.*`) -> YOLO11Pose names with BatchNorm folded in."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, w in flat.items():
        if not key.startswith("model."):
            continue
        idx, _, rest = key[len("model."):].partition(".")
        if not idx.isdigit() or int(idx) not in _LAYER_NAMES:
            continue
        base = _LAYER_NAMES[int(idx)] + "."
        if rest.startswith("dfl."):
            continue  # fixed arange(16) weights: the decode kernel owns them
        if rest.endswith(".bn.weight") or ".bn." in rest or rest.startswith("bn."):
            continue  # consumed together with the convolution below
        if rest.endswith("conv.weight"):
            stem = key[: -len("conv.weight")]
            bn = stem + "bn."
            bias = flat.get(stem + "conv.bias")
            if bn + "weight" in flat:
                w, bias = fold_conv_bn(w, bias, flat[bn + "weight"], flat[bn + "bias"], flat[bn + "running_mean"],
                                       flat[bn + "running_var"], float(flat.get(bn + "eps", torch.tensor(1e-3))))
            elif bias is None:
                raise CheckpointError(f"{key}: convolution without bias and without a BatchNorm to fold")
            out[base + rest] = w
            out[base + rest[: -len("weight")] + "bias"] = bias
        elif rest.endswith("conv.bias"):
            continue
        else:
            out[base + rest] = w  # the head's plain nn.Conv2d layers (cv2/cv3/cv4 .2.weight/.bias)
    return out


def yolo11_pose_from_ultralytics(path: str):
    """Build a YOLO11Pose of the checkpoint's own scale / nc / kpt_shape and load the folded weights (strict)."""
    from .backbones.yolo11_pose import YOLO11Pose

    ckpt = read_ultralytics_checkpoint(path)
    root = ckpt.get("ema") or ckpt.get("model") if isinstance(ckpt, dict) else ckpt
    if root is None or not _is_module(root):
        raise CheckpointError(f"{path}: no pickled model under 'ema' / 'model'")
    flat, kinds = {}, {}
    _harvest(root, "", flat, kinds)
    sd = fused_state_dict(flat)
    if "b0.conv.weight" not in sd or "head.cv3.0.2.weight" not in sd:
        raise CheckpointError(f"{path}: not a YOLO11-pose graph (layers model.0 / model.23 missing)")
    c0 = int(sd["b0.conv.weight"].shape[0])
    n_b2 = len([k for k in kinds if k.startswith("model.2.m.") and k.count(".") == 3])
    letter = {16: "n", 32: "s", 96: "x"}.get(c0) or ("l" if n_b2 >= 2 else "m")
    depth, width, max_ch, c3k_all = _SCALES[letter]
    nc = int(sd["head.cv3.0.2.weight"].shape[0])
    nk = int(sd["head.cv4.0.2.weight"].shape[0])
    kpt_shape = tuple(getattr(getattr(root, "_modules", {}).get("model", None)._modules.get("23"), "kpt_shape", (nk // 3, 3)))
    model = YOLO11Pose(nc=nc, kpt_shape=kpt_shape, depth=depth, width=width, max_channels=max_ch, c3k_all=c3k_all)
    try:
        model.load_state_dict(sd, strict=True)
    except RuntimeError as e:
        raise CheckpointError(f"{path}: weights do not fit YOLO11{letter}-pose (nc={nc}, kpt_shape={kpt_shape}): {e}") from e
    names = getattr(root, "names", None)
    return model.eval(), dict(scale=letter, nc=nc, kpt_shape=kpt_shape, names=names if isinstance(names, dict) else {0: "face"})


def load_yolo(path: str):
    """-> (YOLO11Pose, info).  Accepts this package's own `YOLO.save()` file ({'model': state_dict, 'arch': {...}}) or an
    ultralytics checkpoint; raises FileNotFoundError / CheckpointError otherwise (the reference raises as well)."""
    from .backbones.yolo11_pose import YOLO11Pose

    if not os.path.isfile(path):
        raise FileNotFoundError(f"YOLO weights not found: {path}")
    try:
        state = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        state = None  # a pickled ultralytics model: needs the stub reader
    if isinstance(state, dict) and isinstance(state.get("model"), dict):
        arch = state.get("arch") or {}
        model = YOLO11Pose(**{k: (tuple(v) if k == "kpt_shape" else v) for k, v in arch.items()})
        try:
            model.load_state_dict(state["model"], strict=True)
        except RuntimeError as e:
            raise CheckpointError(f"{path}: state_dict does not fit YOLO11Pose({arch}): {e}") from e
        return model.eval(), dict(scale=arch.get("scale", "?"), nc=model.nc, kpt_shape=model.kpt_shape, names={0: "face"})
    return yolo11_pose_from_ultralytics(path)


def load_rrdbnet_state(path: str) -> dict:
    """Real-ESRGAN .pth: the state dict under 'params_ema' (preferred, as RealESRGANer does) or 'params', or a bare one."""
    if not os.path.isfile(path):
        raise FileNotFoundError(f"Real-ESRGAN weights not found: {path}")
    state = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(state, dict):
        for key in ("params_ema", "params"):
            if key in state:
                return state[key]
    return state

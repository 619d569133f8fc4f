"""Algorithmic work of the hot path (2 x MAC of the reference's fp32 ops, unpadded, counted once
whatever the kernel executes) -- used for roofline fractions."""
from __future__ import annotations

from .configs import CONFIGS


def res2net_stage_macs(name: str, length: int = 64600):
    """MACs per utterance of the fork's Res2Net+SE model (AASIST.py:603-669): split convs, conv_cat, downsample."""
    from .model import res2net_splits
    cfg = CONFIGS[name]
    f = cfg["filts"]
    taps = cfg["first_conv"] + 1 if cfg["first_conv"] % 2 == 0 else cfg["first_conv"]
    t = length - taps + 1
    out = {"sinc": f[0] * taps * t}
    w = t // 3
    for i, (ci, co) in enumerate([f[1], f[2], f[3], f[4], f[4], f[4]]):
        sizes, _ = res2net_splits((ci, co), cfg.get("res2net_width", 14), cfg.get("res2net_scale", 8))
        out[f"enc{i}.splits"] = sum(n * n * 9 for n in sizes) * 23 * w
        out[f"enc{i}.conv_cat"] = co * ci * 9 * 23 * w
        out[f"enc{i}.ds"] = co * ci * 3 * 23 * w if ci != co else 0
        w //= 3
    out["graph"] = 11.6e6
    return out


def stage_macs(name: str, length: int = 64600):
    """MACs per utterance by stage: {'sinc', 'enc{i}.conv1', 'enc{i}.conv2', 'enc{i}.ds', 'graph'}."""
    cfg = CONFIGS[name]
    if "res2net_width" in cfg:
        return res2net_stage_macs(name, length)
    f = cfg["filts"]
    taps = cfg["first_conv"] + 1 if cfg["first_conv"] % 2 == 0 else cfg["first_conv"]
    t = length - taps + 1
    out = {"sinc": f[0] * taps * t}
    w = t // 3
    chans = [f[1], f[2], f[3], f[4], f[4], f[4]]
    n_enc = 2 if name == "RawGAT-ST" else 1
    for i, (ci, co) in enumerate(chans):
        out[f"enc{i}.conv1"] = n_enc * co * ci * 6 * 24 * w
        out[f"enc{i}.conv2"] = n_enc * co * co * 6 * 23 * w
        out[f"enc{i}.ds"] = n_enc * (co * ci * 3 * 23 * w if ci != co else 0)
        w //= 3
    out["graph"] = {"AASIST": 11.6e6, "AASIST-L": 2.55e6, "RawGAT-ST": 3.25e6}[name]
    return out


def kernel_flops(name: str, kernel: str, batch: int, length: int = 64600) -> float:
    """Algorithmic FLOPs per forward of all launches reported under `kernel`."""
    m = stage_macs(name, length)
    if kernel.startswith("enc") and ".res2_" in kernel:     # Res2Net: "enc{i}.res2_split_convs_f32" / "...conv_cat..."
        blk = kernel.split(".")[0]
        macs = m[f"{blk}.splits"] if "split" in kernel else m[f"{blk}.conv_cat"] + m[f"{blk}.ds"]
        return 2.0 * macs * batch
    if kernel.startswith("enc") and "." in kernel:          # "enc{i}.conv1[_tc]" / "enc{i}.conv2[_tc]"
        blk, conv = kernel.split(".")[0], kernel.split(".")[1]
        if conv.startswith("fused"):   # whole block in one kernel
            macs = m[f"{blk}.conv1"] + m[f"{blk}.conv2"] + m[f"{blk}.ds"]
        else:
            macs = m[f"{blk}.conv1"] if conv.startswith("conv1") else m[f"{blk}.conv2"] + m[f"{blk}.ds"]
        return 2.0 * macs * batch
    if kernel.startswith("sinc_frontend"):
        macs = m["sinc"]
    elif kernel.startswith("res2_conv_cat"):
        macs = sum(v for k, v in m.items() if k.endswith(".conv_cat"))
    elif kernel.startswith("res2_split"):
        macs = sum(v for k, v in m.items() if k.endswith(".splits"))
    elif kernel.startswith("res2_gate"):
        macs = sum(v for k, v in m.items() if k.endswith(".ds"))
    elif kernel.startswith("conv1") and "[1->C]" in kernel:
        macs = m["enc0.conv1"]
    elif kernel.startswith("conv1"):
        macs = sum(m[f"enc{i}.conv1"] for i in range(1, 6))
    elif kernel.startswith("conv2"):
        macs = sum(m[f"enc{i}.conv2"] + m[f"enc{i}.ds"] for i in range(6))
    elif kernel.startswith("block0_conv1"):
        macs = m["enc0.conv1"]
    elif "graph" in kernel:
        macs = m["graph"]
    else:
        macs = 0.0
    return 2.0 * macs * batch

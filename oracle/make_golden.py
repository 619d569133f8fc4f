"""Generate the golden fixtures under tests/golden/ by running the REFERENCE itself.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where the reference is
mounted read-only at /root/reference (it does not exist on the GPU box, so nothing
at test/bench time imports it -- only the committed .npz fixtures travel).

The reference classes are imported unmodified:
  * fork ``models/AASIST.py::Model`` (forward dataflow, CONV filter bank, graph layers),
    with ``.encoder`` rebuilt from ``models/RawNetGatSpoofST.py::Residual_block``
    -- the (2,3)-kernel block the shipped checkpoints were trained with;
  * ``models/RawNetGatSpoofST.py::Model`` for the RawGAT-ST baseline.  No checkpoint
    ships for it; its weights are the class's own init under ``torch.manual_seed(1234)``
    with BN statistics/affines randomised (so BN folding is exercised), saved to
    ``aasist_b200/weights/RawGATST_seed1234.pth``.

Usage:  python oracle/make_golden.py            (about one minute on 8 cores)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("AASIST_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import aasist_oracle as O  # noqa: E402  (input generators + config dicts only)

GOLD = os.path.join(ROOT, "tests", "golden")
WDIR = os.path.join(ROOT, "aasist_b200", "weights")
POOLS = ["pool_S", "pool_T", "pool_hS1", "pool_hT1", "pool_hS2", "pool_hT2"]


def reference_aasist(name: str):
    sys.path.insert(0, REF)
    from models.AASIST import Model
    from models.RawNetGatSpoofST import Residual_block as RB23
    conf = {"AASIST": "AASIST.conf", "AASIST-L": "AASIST-L.conf"}[name]
    mc = json.load(open(f"{REF}/config/{conf}"))["model_config"]
    assert mc == O.CONFIGS[name], "restated config differs from the reference's"
    f = mc["filts"]
    m = Model(mc)
    m.encoder = nn.Sequential(
        nn.Sequential(RB23(f[1], first=True)), nn.Sequential(RB23(f[2])),
        nn.Sequential(RB23(f[3])), nn.Sequential(RB23(f[4])),
        nn.Sequential(RB23(f[4])), nn.Sequential(RB23(f[4])))
    sd = torch.load(f"{REF}/models/weights/{name}.pth", map_location="cpu")
    m.load_state_dict(sd, strict=True)
    m.eval()
    return m, mc


def reference_rawgat():
    sys.path.insert(0, REF)
    from models.RawNetGatSpoofST import Model
    mc = json.load(open(f"{REF}/config/RawGATST_baseline.conf"))["model_config"]
    assert mc == O.CONFIGS["RawGAT-ST"]
    torch.manual_seed(1234)
    m = Model(mc)
    g = torch.Generator().manual_seed(1234)
    for mod in m.modules():
        if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
            n = mod.num_features
            mod.running_mean.copy_(0.2 * torch.randn(n, generator=g))
            mod.running_var.copy_(0.5 + torch.rand(n, generator=g))
            mod.weight.data.copy_(0.7 + 0.6 * torch.rand(n, generator=g))
            mod.bias.data.copy_(0.1 * torch.randn(n, generator=g))
    # first_bn of the trained models scales by ~121; random init would starve the
    # encoder of signal on 0.05-amplitude inputs, so give it a comparable gain.
    m.first_bn.running_var.fill_(6e-5)
    m.first_bn.running_mean.fill_(3e-3)
    m.eval()
    os.makedirs(WDIR, exist_ok=True)
    torch.save(m.state_dict(), os.path.join(WDIR, "RawGATST_seed1234.pth"))
    return m, mc


def run_with_taps(m, x, pools):
    taps = {}
    hooks = []

    def grab(name):
        def fn(_mod, _inp, out):
            # clone at hook time: the reference applies nn.SELU(inplace=True) to some of
            # these tensors right after the hooked module returns
            taps[name] = out.detach().clone()
        return fn

    def grab_pool(name, mod):
        def fn(_mod, inp, _out):
            h = inp[0]
            w = mod.proj(h)
            s = torch.sigmoid(w)
            k = _out.shape[1]
            taps[name + ".weights"] = w.squeeze(-1)
            taps[name + ".scores"] = s.squeeze(-1)
            taps[name + ".idx"] = torch.topk(s, k, dim=1)[1].squeeze(-1)
        return fn

    for name, mod in m.named_modules():
        if name in pools:
            hooks.append(mod.register_forward_hook(grab_pool(name, mod)))
        elif name.startswith("GAT_layer") or name == "first_bn" or name == "conv_time":
            hooks.append(mod.register_forward_hook(grab(name)))
        elif name.startswith("encoder") and name.count(".") == 1:
            hooks.append(mod.register_forward_hook(grab(name)))
    with torch.no_grad():
        last_hidden, output = m(x)
    for h in hooks:
        h.remove()
    taps["last_hidden"], taps["output"] = last_hidden, output
    return taps


def pack(taps, x, bank, pools, enc_prefixes, slim=False):
    out = {
        "x_head": x[:, :8].numpy(), "x_sum": x.double().sum(dim=1).numpy(),
        "bank": bank.numpy(),
        "last_hidden": taps["last_hidden"].numpy(), "output": taps["output"].numpy(),
    }
    # frontend: tap is first_bn's output (pre-SELU clone) -> store the post-SELU sample grid
    z = torch.nn.functional.selu(taps["first_bn"])
    out["frontend_sample"] = z[:, 0, :, ::211].numpy()
    out["frontend_absmean"] = z.abs().mean(dim=(1, 2, 3)).numpy()
    for pre in enc_prefixes:
        for i in range(6):
            e = taps[f"{pre}.{i}"]
            step = max(1, e.shape[3] // 16)
            out[f"{pre}.{i}.sample"] = e[:, ::5, :, ::step].numpy()
            out[f"{pre}.{i}.absmean"] = e.abs().mean(dim=(1, 2, 3)).numpy()
        out[f"{pre}.5.full"] = taps[f"{pre}.5"].numpy()
    for k in taps:
        # slim (long-utterance fixtures): keep the layer outputs, drop the (B,N,N,D) sub-module taps
        if k.startswith("GAT_layer") and not (slim and "." in k):
            out[k] = taps[k].numpy()
    for p in pools:
        out[p + ".weights"] = taps[p + ".weights"].numpy()
        out[p + ".idx"] = taps[p + ".idx"].numpy().astype(np.int32)
    return out


def reference_pad():
    """The reference's own `pad` (data_utils.py:45-52).  data_utils imports soundfile, which is not
    installed, so the function is compiled from its source in place instead of importing the module."""
    import ast
    src = open(f"{REF}/data_utils.py").read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "pad"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), f"{REF}/data_utils.py", "exec"), ns)
    return ns["pad"]


PAD_CASES = [(7, 101), (1000, 102), (21533, 103), (64599, 104), (64600, 105), (64601, 106), (100000, 107)]


def make_pad_golden():
    pad = reference_pad()
    out = {}
    for n, seed in PAD_CASES:
        x = O.white_noise(1, n, seed)[0].numpy()
        y = np.asarray(pad(x, 64600), dtype=np.float32)
        assert y.shape == (64600,)
        out[f"len{n}.sample"] = y[::499]
        out[f"len{n}.tail"] = y[-16:]
        out[f"len{n}.sum"] = np.array(y.astype(np.float64).sum())
    np.savez_compressed(os.path.join(GOLD, "pad.npz"), **out)
    print("pad golden written")


# BASELINE.json configs[4] (input-length sweep up to 256 000 samples = 116 temporal nodes)
LONG_CASES = [
    ("speech128k", O.speech_like, 2, 128000, 17),
    ("speech192k", O.speech_like, 2, 192000, 19),
    ("speech256k", O.speech_like, 2, 256000, 23),
]


def main():
    if "--only-pad" in sys.argv:
        return make_pad_golden()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 8)
    cases = [
        # (tag, generator, n_utt, length, seed)
        ("white", O.white_noise, 4, 64600, 1234),
        ("speech", O.speech_like, 4, 64600, 7),
        ("speech16k", O.speech_like, 2, 16000, 11),
        ("speech96k", O.speech_like, 2, 96000, 13),
    ]
    only_long = "--only-long" in sys.argv        # add the long fixtures without rewriting the others
    summary = {}
    for name in ("AASIST", "AASIST-L"):
        m, mc = reference_aasist(name)
        nparam = sum(p.numel() for p in m.parameters())
        summary[name] = {"n_params": nparam}
        bank = m.conv_time.band_pass.clone()
        for tag, gen, n, L, seed in (LONG_CASES if only_long else cases + LONG_CASES):
            x = gen(n, L, seed)
            taps = run_with_taps(m, x, POOLS)
            d = pack(taps, x, bank, POOLS, ["encoder"], slim=L > 100000)
            d["meta"] = np.array(json.dumps({"model": name, "input": tag, "n": n, "L": L, "seed": seed,
                                             "n_params": nparam, "torch": torch.__version__,
                                             "numpy": np.__version__}))
            np.savez_compressed(os.path.join(GOLD, f"{name}_{tag}.npz"), **d)
            print(name, tag, "logits[0] =", taps["output"][0].tolist())
    if only_long:
        return
    m, mc = reference_rawgat()
    nparam = sum(p.numel() for p in m.parameters())
    summary["RawGAT-ST"] = {"n_params": nparam}
    bank = m.conv_time.band_pass.clone()
    rpools = ["pool_T", "pool_S", "pool_ST"]
    for tag, gen, n, L, seed in cases[:2]:
        x = gen(n, L, seed)
        taps = run_with_taps(m, x, rpools)
        d = pack(taps, x, bank, rpools, ["encoder_T", "encoder_S"])
        d["meta"] = np.array(json.dumps({"model": "RawGAT-ST", "input": tag, "n": n, "L": L, "seed": seed,
                                         "n_params": nparam, "torch": torch.__version__,
                                         "numpy": np.__version__}))
        np.savez_compressed(os.path.join(GOLD, f"RawGAT-ST_{tag}.npz"), **d)
        print("RawGAT-ST", tag, "logits[0] =", taps["output"][0].tolist())
    make_pad_golden()
    json.dump(summary, open(os.path.join(GOLD, "summary.json"), "w"), indent=1)
    print(summary)


if __name__ == "__main__":
    main()

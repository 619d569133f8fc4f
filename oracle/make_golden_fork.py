"""Golden fixtures for the fork-only rows of SURVEY 8(f), produced by running the REFERENCE itself.

TEST INFRASTRUCTURE ONLY; runs in the build container (reference mounted at /root/reference).  Imports,
unmodified:
  * fork ``models/AASIST.py::Model`` (Res2Net+SE encoder, SpeakerConditioningModule)   -> fork_AASIST2*.npz
  * the same class on the shipped AASIST checkpoint with ``Freq_aug=True``             -> fork_freqaug.npz
  * ``models/AASIST_Robust.py::Model``                                                 -> fork_robust*.npz
  * ``data_utils.py::pad_sequence`` / ``dynamic_chunk_size`` (compiled from source:
    the module imports soundfile, which is not installed)                              -> fork_staging.npz
No checkpoints exist for the fork models: weights are the classes' own init under torch.manual_seed(1234)
with randomised BN statistics, saved next to the shipped checkpoints (aasist_b200/weights/*_seed1234.pth).

Usage:  python oracle/make_golden_fork.py
"""
from __future__ import annotations

import ast
import json
import os
import random
import sys

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("AASIST_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
from oracle import aasist_oracle as O  # noqa: E402
from oracle import aasist2_oracle as O2  # noqa: E402
from oracle.make_golden import POOLS, reference_aasist  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
WDIR = os.path.join(ROOT, "aasist_b200", "weights")


def randomise_bn(m: nn.Module, seed: int) -> None:
    g = torch.Generator().manual_seed(seed)
    for mod in m.modules():
        if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
            n = mod.num_features
            mod.running_mean.copy_(0.2 * torch.randn(n, generator=g))
            mod.running_var.copy_(0.5 + torch.rand(n, generator=g))
            mod.weight.data.copy_(0.7 + 0.6 * torch.rand(n, generator=g))
            mod.bias.data.copy_(0.1 * torch.randn(n, generator=g))
    # the trained first_bn scales the pooled |sinc| output by ~121; keep a comparable gain
    m.first_bn.running_var.fill_(6e-5)
    m.first_bn.running_mean.fill_(3e-3)


def pool_hooks(m, pools, taps):
    hooks = []

    def grab_pool(name, mod):
        def fn(_mod, inp, out):
            w = mod.proj(inp[0])
            s = torch.sigmoid(w)
            taps[name + ".weights"] = w.squeeze(-1).detach().clone()
            taps[name + ".idx"] = torch.topk(s, out.shape[1], dim=1)[1].squeeze(-1)
        return fn

    def grab(name):
        def fn(_mod, _inp, out):
            taps[name] = out.detach().clone()
        return fn

    for name, mod in m.named_modules():
        if name in pools:
            hooks.append(mod.register_forward_hook(grab_pool(name, mod)))
        elif name.startswith("encoder") and name.count(".") == 1:
            hooks.append(mod.register_forward_hook(grab(name)))
        elif name == "first_bn":
            hooks.append(mod.register_forward_hook(grab(name)))
    return hooks


def pack_common(taps, x, pools):
    out = {"x_head": x[:, :8].numpy(), "x_sum": x.double().sum(dim=1).numpy()}
    z = torch.nn.functional.selu(taps["first_bn"])
    out["frontend_absmean"] = z.abs().mean(dim=(1, 2, 3)).numpy()
    for i in range(6):
        e = taps[f"encoder.{i}"]
        step = max(1, e.shape[3] // 16)
        out[f"encoder.{i}.sample"] = e[:, ::5, :, ::step].numpy()
        out[f"encoder.{i}.absmean"] = e.abs().mean(dim=(1, 2, 3)).numpy()
    out["encoder.5.full"] = taps["encoder.5"].numpy()
    for p in pools:
        out[p + ".weights"] = taps[p + ".weights"].numpy()
        out[p + ".idx"] = taps[p + ".idx"].numpy().astype(np.int32)
    return out


# ------------------------------------------------------------------------------------------------
def make_res2net():
    from models.AASIST import Model
    for cname, fname in (("AASIST2", "AASIST2.conf"), ("AASIST2-small", None)):
        mc = O2.CONFIGS[cname]
        if fname:
            ref_mc = json.load(open(f"{REF}/config/{fname}"))["model_config"]
            assert ref_mc == mc, "restated AASIST2 config differs from the reference's"
        torch.manual_seed(1234)
        m = Model(dict(mc))
        randomise_bn(m, 1234)
        g = torch.Generator().manual_seed(99)
        for mod in m.modules():                       # SE gates away from 0.5, biases away from 0
            if isinstance(mod, nn.Linear) and mod.bias is None:
                mod.weight.data.mul_(3.0)
        m.eval()
        torch.save(m.state_dict(), os.path.join(WDIR, f"{cname}_seed1234.pth"))
        nparam = sum(p.numel() for p in m.parameters())
        for tag, n, L, seed in (("speech", 4, 64600, 7), ("speech24k", 3, 24000, 29)):
            x = O.speech_like(n, L, seed)
            emb = torch.randn(n, mc["spk_emb_dim"], generator=torch.Generator().manual_seed(seed + 1))
            taps = {}
            hooks = pool_hooks(m, POOLS, taps)
            with torch.no_grad():
                lh, out = m(x)
            d = pack_common(taps, x, POOLS)
            d["last_hidden"], d["output"] = lh.numpy(), out.numpy()
            with torch.no_grad():
                lh_s, out_s = m(x, speaker_embedding=emb)
            for h in hooks:
                h.remove()
            d["spk_embedding"] = emb.numpy()
            d["spk.last_hidden"], d["spk.output"] = lh_s.numpy(), out_s.numpy()
            d["meta"] = np.array(json.dumps({"model": cname, "input": "speech", "n": n, "L": L, "seed": seed,
                                             "n_params": nparam, "n_tensors": len(m.state_dict()),
                                             "torch": torch.__version__}))
            np.savez_compressed(os.path.join(GOLD, f"fork_{cname}_{tag}.npz"), **d)
            print(cname, tag, "logits[0] =", out[0].tolist(), "with speaker:", out_s[0].tolist())
        # utterance-level conditioning cannot run in the reference (Linear(2*g1) applied to 5*g1+g1 features)
        mc_u = dict(mc, conditioning_level="utterance")
        mu = Model(mc_u).eval()
        try:
            with torch.no_grad():
                mu(O.speech_like(1, 16000, 1), speaker_embedding=torch.zeros(1, mc["spk_emb_dim"]))
            raise AssertionError("expected the reference to fail")
        except RuntimeError as e:
            print(cname, "utterance-level conditioning -> RuntimeError:", str(e)[:80])


def make_freq_aug():
    m, mc = reference_aasist("AASIST")
    x = O.speech_like(3, 64600, 41)
    d = {"x_head": x[:, :8].numpy()}
    masks = []
    for seed in (1, 2, 3, 4, 5, 6):
        np.random.seed(seed)
        random.seed(seed)
        with torch.no_grad():
            lh, out = m(x, Freq_aug=True)
        f = m.conv_time.filters[:, 0, :]
        zero = (f.abs().sum(dim=1) == 0).nonzero().flatten().tolist()
        a0, a = (zero[0], len(zero)) if zero else (0, 0)
        assert zero == list(range(a0, a0 + a))
        np.random.seed(seed)
        random.seed(seed)
        got = O2.draw_freq_mask(f.shape[0])
        assert a == 0 or got == (a0, a), (got, a0, a)
        masks.append([seed, got[0], got[1]])
        d[f"seed{seed}.output"], d[f"seed{seed}.last_hidden"] = out.numpy(), lh.numpy()
        print("freq_aug seed", seed, "mask rows", got, "logits[0]", out[0].tolist())
    d["meta"] = np.array(json.dumps({"model": "AASIST", "n": 3, "L": 64600, "seed": 41, "masks": masks}))
    np.savez_compressed(os.path.join(GOLD, "fork_freqaug.npz"), **d)


def make_robust():
    from models.AASIST_Robust import Model
    mc = O2.CONFIGS["AASIST-Robust"]
    ref_mc = json.load(open(f"{REF}/config/AASIST-Robust.conf"))["model_config"]
    assert dict(ref_mc, first_conv=70) == mc
    torch.manual_seed(1234)
    m = Model(dict(mc))
    randomise_bn(m, 4321)
    m.ensemble_weight.data.copy_(torch.tensor([0.3, -0.4]))
    m.eval()
    torch.save(m.state_dict(), os.path.join(WDIR, "AASIST-Robust_seed1234.pth"))
    nparam = sum(p.numel() for p in m.parameters())
    pools = ["pool_S", "pool_T", "pool_hS", "pool_hT"]
    for tag, n, L, seed in (("nt1", 2, 600000, 51), ("nt3", 2, 1700000, 53)):
        x = O.speech_like(n, L, seed)
        taps = {}
        hooks = pool_hooks(m, pools, taps)
        with torch.no_grad():
            ens, logits = m(x)
        for h in hooks:
            h.remove()
        d = pack_common(taps, x, pools)
        d["ensemble"], d["logits"] = ens.numpy(), logits.numpy()
        d["meta"] = np.array(json.dumps({"model": "AASIST-Robust", "input": "speech", "n": n, "L": L, "seed": seed,
                                         "n_params": nparam, "n_tensors": len(m.state_dict())}))
        np.savez_compressed(os.path.join(GOLD, f"fork_robust_{tag}.npz"), **d)
        print("robust", tag, "NT =", taps["encoder.5"].shape[3], "ensemble[0] =", ens[0].tolist(), "logits[0] =",
              logits[0].tolist())
    # the errors a drop-in has to reproduce: the shipped config (42 bands vs pos_S 23) and short inputs
    errs = {}
    for fc, L in ((128, 600000), (70, 64600)):
        mm = Model(dict(mc, first_conv=fc)).eval()
        try:
            with torch.no_grad():
                mm(torch.zeros(1, L))
            errs[f"{fc}/{L}"] = "ok"
        except RuntimeError as e:
            errs[f"{fc}/{L}"] = "RuntimeError: " + str(e)[:90]
    print("robust reference errors:", errs)


def _ref_functions(*names):
    src = open(f"{REF}/data_utils.py").read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=fns, type_ignores=[]), f"{REF}/data_utils.py", "exec"), ns)
    return [ns[n] for n in names]


STAGING_LENGTHS = [(16000, 201), (16001, 202), (40003, 203), (7, 204), (95998, 205), (64600, 206)]
CHUNK_CASES = [(5000, 301), (30000, 302), (120000, 303), (96000, 304), (15999, 305)]


def make_staging():
    pad_sequence, dynamic_chunk_size = _ref_functions("pad_sequence", "dynamic_chunk_size")
    d = {}
    seqs = [O.white_noise(1, n, seed)[0] for n, seed in STAGING_LENGTHS]
    X, y, dur = pad_sequence([(s, i % 2, float(s.numel()) / 16000) for i, s in enumerate(seqs)])
    d["pad_sequence.shape"] = np.array(X.shape)
    d["pad_sequence.sample"] = X[:, ::997].numpy()
    d["pad_sequence.rowsum"] = X.double().sum(dim=1).numpy()
    d["pad_sequence.tail"] = X[:, -8:].numpy()
    # dynamic_chunk_size: record the reference's draws and its outputs under a fixed numpy seed
    for n, seed in CHUNK_CASES:
        x = O.white_noise(1, n, seed)[0].numpy()
        np.random.seed(seed)
        y, duration = dynamic_chunk_size(x, 16000, 96000)
        np.random.seed(seed)
        target = np.random.randint(16000, 96000 + 1)
        start = np.random.randint(0, n - target + 1) if n >= target else 0
        assert y.shape[0] == target and np.array_equal(y, O2.chunk(x, target, start))
        d[f"chunk{n}.target_start"] = np.array([target, start])
        d[f"chunk{n}.sample"] = y[::499].astype(np.float32)
        d[f"chunk{n}.sum"] = np.array(y.astype(np.float64).sum())
    assert np.array_equal(O2.pad_sequence([s.numpy() for s in seqs]), X.numpy())
    np.savez_compressed(os.path.join(GOLD, "fork_staging.npz"), **d)
    print("staging golden written; pad_sequence ->", tuple(X.shape))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 8)
    what = sys.argv[1:] or ["res2net", "freqaug", "robust", "staging"]
    if "staging" in what:
        make_staging()
    if "freqaug" in what:
        make_freq_aug()
    if "res2net" in what:
        make_res2net()
    if "robust" in what:
        make_robust()


if __name__ == "__main__":
    main()

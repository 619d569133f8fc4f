"""CPU emulation of reduced-product tensor-core modes (SURVEY 7.2 "fast mode").

TEST/ANALYSIS TOOL (imports oracle/): for each candidate mode the encoder convolutions of chosen blocks are
evaluated with one operand rounded to fp16 (= dropping one of the two correction products of the f16x3
scheme) or both (= a single product), everything else fp32.  Reports logits max-abs error and ordered
top-k agreement against the unmodified oracle on speech-like utterances.  A mode may only ship when it
holds logits <= 1e-3 and 100 % ordered-index match.

    python tools/precision_emulation.py [n_utt]
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import aasist_oracle as O  # noqa: E402
from tests.util import load_sd, pools_of  # noqa: E402


def r16(t):
    return t.half().float()


def make_block(mode_for_block):
    """mode per block index: None (exact) | 'w' (weights fp16: drops a_hi*w_lo) | 'a' (activations fp16:
    drops a_lo*w_hi) | 'aw' (single product)."""
    def block(x, sd, prefix):
        i = int(prefix.split(".")[1])
        mode = mode_for_block.get(i)
        rw = (lambda t: r16(t)) if mode and "w" in mode else (lambda t: t)
        ra = (lambda t: r16(t)) if mode and "a" in mode else (lambda t: t)
        # conv1 with bn2 folded (as the kernels do), then SELU
        g = sd[prefix + ".bn2.weight"].double() / torch.sqrt(sd[prefix + ".bn2.running_var"].double() + 1e-5)
        w1 = (sd[prefix + ".conv1.weight"].double() * g.view(-1, 1, 1, 1)).float()
        b1 = (sd[prefix + ".conv1.bias"].double() * g + sd[prefix + ".bn2.bias"].double()
              - sd[prefix + ".bn2.running_mean"].double() * g).float()
        out = F.selu(F.conv2d(ra(x), rw(w1), b1, padding=(1, 1)))
        out = F.conv2d(ra(out), rw(sd[prefix + ".conv2.weight"]), sd[prefix + ".conv2.bias"], padding=(0, 1))
        identity = x
        if (prefix + ".conv_downsample.weight") in sd:
            identity = F.conv2d(ra(x), rw(sd[prefix + ".conv_downsample.weight"]),
                                sd[prefix + ".conv_downsample.bias"], padding=(0, 1))
        return F.max_pool2d(out + identity, (1, 3))
    return block


MODES = {
    "exact(folded)": {},
    "w16 blocks2-5": {i: "w" for i in range(2, 6)},
    "a16 blocks2-5": {i: "a" for i in range(2, 6)},
    "single blocks2-5": {i: "aw" for i in range(2, 6)},
    "w16 blocks3-5": {i: "w" for i in range(3, 6)},
    "single blocks3-5": {i: "aw" for i in range(3, 6)},
    "w16 blocks1-5": {i: "w" for i in range(1, 6)},
    "w16 blocks0-5": {i: "w" for i in range(0, 6)},
    "a16 blocks0-5": {i: "a" for i in range(0, 6)},
    "single blocks0-5": {i: "aw" for i in range(0, 6)},
}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    torch.set_num_threads(os.cpu_count() or 8)
    orig = O.residual_block
    rows = []
    for name in ("AASIST", "AASIST-L"):
        sd, cfg = load_sd(name), O.CONFIGS[name]
        x = O.speech_like(n, 64600, 777)
        ref = {}
        O.residual_block = orig
        O.forward(name, sd, cfg, x, ref)
        for mode, spec in MODES.items():
            O.residual_block = make_block(spec)
            taps = {}
            O.forward(name, sd, cfg, x, taps)
            err = (taps["output"] - ref["output"]).abs().max().item()
            herr = (taps["last_hidden"] - ref["last_hidden"]).abs().max().item()
            mism = pos = utt_bad = 0
            bad = torch.zeros(n, dtype=torch.bool)
            for p in pools_of(name):
                a, b = ref[p + ".idx"], taps[p + ".idx"]
                mism += (a != b).sum().item()
                pos += a.numel()
                bad |= (a != b).any(dim=1)
            row = {"model": name, "mode": mode, "logit_err": err, "hidden_err": herr,
                   "topk_mismatch_positions": mism, "positions": pos, "utterances_with_mismatch": int(bad.sum()),
                   "n_utt": n}
            rows.append(row)
            print(json.dumps(row), flush=True)
    O.residual_block = orig
    return rows


if __name__ == "__main__":
    main()

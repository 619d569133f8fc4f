"""CPU oracle for the fork-only parts of the path.

TEST INFRASTRUCTURE ONLY (see oracle/aasist_oracle.py for the rules): functional torch-fp32 restatements of

  * the fork's Res2Net+SE encoder block          models/AASIST.py:506-669
  * SpeakerConditioningModule                      models/AASIST.py:325-415
  * the fork's ``Model.forward`` with them         models/AASIST.py:806-921
  * ``Freq_aug`` filter masking                    models/AASIST.py:484-490
  * the fork's 3x3 ``Residual_block``              models/AASIST.py:672-725
  * AASIST-Robust ``Model.forward`` (eval)         models/AASIST_Robust.py:198-303
  * ``pad_sequence`` / ``dynamic_chunk_size``      data_utils.py:68-119

Pinned by ``oracle/make_golden_fork.py`` (imports the unmodified reference classes / functions, runs them
under fixed seeds, commits ``tests/golden/fork_*.npz``); ``tests/test_oracle_fork.py`` checks this file
against those fixtures.  No checkpoint exists for any of these models: weights are the reference classes'
own seeded initialisation with randomised BN statistics (same recipe as RawGAT-ST).
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import aasist_oracle as O

Tensor = torch.Tensor

# model_config of config/AASIST2.conf:21-35 (the fork's own configuration) and a small variant that
# exercises the other code paths (scale-group chaining, no attention, remainder splits).
CONFIGS: Dict[str, dict] = {
    "AASIST2": {
        "architecture": "AASIST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 64], [64, 64]],
        "gat_dims": [64, 32], "pool_ratios": [0.5, 0.7, 0.5, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
        "res2net_width": 14, "res2net_scale": 8,
        "speaker_conditioning": True, "spk_emb_dim": 256, "conditioning_level": "frame",
        "use_attention": True,
    },
    "AASIST2-small": {
        "architecture": "AASIST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 24], [24, 24]],
        "gat_dims": [24, 32], "pool_ratios": [0.4, 0.5, 0.7, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
        "res2net_width": 6, "res2net_scale": 2,
        "speaker_conditioning": True, "spk_emb_dim": 64, "conditioning_level": "frame",
        "use_attention": False,
    },
    # config/AASIST-Robust.conf:24-31 with first_conv 70 instead of 128: with 128 the reference builds 42
    # spectral bands and fails at `e_S + pos_S` (pos_S is (1,23,C), AASIST_Robust.py:126,237) for every input
    "AASIST-Robust": {
        "architecture": "AASIST_Robust", "nb_samp": 64600, "first_conv": 70,
        "filts": [70, [1, 32], [32, 32], [32, 24], [24, 24]],
        "gat_dims": [24, 32], "pool_ratios": [0.4, 0.5, 0.7, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
    },
}


# --------------------------------------------------------------------------- #
# Res2Net split bookkeeping            models/AASIST.py:528-574
# --------------------------------------------------------------------------- #
def res2net_splits(nb_filts, width: int = 14, scale: int = 8) -> Tuple[List[int], int]:
    """-> (split_sizes, effective scale) exactly as ``Res2NetBlock.__init__`` computes them."""
    w = min(width, nb_filts[0])                                       # :530
    s = min(scale, w)                                                 # :531
    base = max(1, nb_filts[0] // w)                                   # :544
    rem = nb_filts[0] - base * (w - 1)                                # :545
    return [max(1, base if i < w - 1 else rem) for i in range(w)], s  # :551-565


def se_layer(x: Tensor, sd: dict, prefix: str) -> Tensor:
    """SELayer.forward (models/AASIST.py:518-522): global average -> FC -> ReLU -> FC -> sigmoid -> scale."""
    b, c = x.shape[:2]
    y = x.mean(dim=(2, 3))                                            # AdaptiveAvgPool2d(1)
    y = F.relu(F.linear(y, sd[prefix + ".fc.0.weight"]))
    y = torch.sigmoid(F.linear(y, sd[prefix + ".fc.2.weight"]))
    return x * y.view(b, c, 1, 1)


def res2net_block(x: Tensor, sd: dict, prefix: str, nb_filts, width: int, scale: int, first: bool) -> Tensor:
    """Res2NetBlock.forward (models/AASIST.py:603-669).  Unlike the (2,3) Residual_block, bn1+SELU on the
    input IS live here (:611-613)."""
    identity = x
    if not first:
        x = F.selu(O._bn_eval(x, sd, prefix + ".bn1", 1))             # :611-613
    sizes, sc = res2net_splits(nb_filts, width, scale)
    spx = torch.split(x, sizes, dim=1)                                # :627
    outs = []
    sp = None
    for i in range(len(sizes)):                                       # :631-643
        if i == 0 or i % sc != 0:
            sp = spx[i]
        else:
            sp = sp + spx[i]                                          # previous split's conv OUTPUT + this split
        sp = F.conv2d(sp, sd[f"{prefix}.convs.{i}.weight"], sd[f"{prefix}.convs.{i}.bias"], padding=(1, 1))
        outs.append(sp)
    out = torch.cat(outs, dim=1)                                      # :650
    out = F.selu(O._bn_eval(out, sd, prefix + ".bn2", 1))             # :653-654
    out = F.conv2d(out, sd[prefix + ".conv_cat.weight"], sd[prefix + ".conv_cat.bias"], padding=(1, 1))  # :655
    out = se_layer(out, sd, prefix + ".se")                           # :658
    if (prefix + ".conv_downsample.weight") in sd:                    # :661-662
        identity = F.conv2d(identity, sd[prefix + ".conv_downsample.weight"],
                            sd[prefix + ".conv_downsample.bias"], padding=(0, 1))
    return F.max_pool2d(out + identity, (1, 3))                       # :665-668


def residual_block_3x3(x: Tensor, sd: dict, prefix: str) -> Tensor:
    """The fork's 3x3 Residual_block (models/AASIST.py:703-725): conv1 3x3 pad 1 -> bn2 -> SELU -> conv2 3x3
    pad 1 -> + identity / conv_downsample -> MaxPool2d((1,3)).  bn1+SELU is dead code here too (:706-712)."""
    out = F.conv2d(x, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"], padding=(1, 1))
    out = F.selu(O._bn_eval(out, sd, prefix + ".bn2", 1))
    out = F.conv2d(out, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"], padding=(1, 1))
    identity = x
    if (prefix + ".conv_downsample.weight") in sd:
        identity = F.conv2d(x, sd[prefix + ".conv_downsample.weight"],
                            sd[prefix + ".conv_downsample.bias"], padding=(0, 1))
    return F.max_pool2d(out + identity, (1, 3))


# --------------------------------------------------------------------------- #
# SpeakerConditioningModule              models/AASIST.py:325-415
# --------------------------------------------------------------------------- #
def speaker_conditioning(features: Tensor, emb: Tensor, sd: dict, prefix: str, level: str,
                         use_attention: bool) -> Tensor:
    spk = O._linear(emb, sd, prefix + ".proj")                        # :382
    if level == "frame":
        n = features.size(1)
        spk = spk.unsqueeze(1).expand(-1, n, -1)                      # :389
        if use_attention:
            cat = torch.cat([features, spk], dim=2)                   # :393
            a = torch.tanh(O._linear(cat, sd, prefix + ".attention.0"))
            a = F.softmax(O._linear(a, sd, prefix + ".attention.2"), dim=1)   # :350-355 softmax over frames
            ctx = a * spk                                             # :397
            return F.relu(O._linear(torch.cat([features, ctx], dim=2), sd, prefix + ".fusion.0"))  # :400
        return F.relu(O._linear(torch.cat([features, spk], dim=2), sd, prefix + ".fusion.0"))      # :403
    if features.dim() == 3:                                           # :407-409
        features = features.mean(dim=1)
    return F.relu(O._linear(torch.cat([features, spk], dim=1), sd, prefix + ".fusion.0"))          # :412


def mask_filterbank(bank: Tensor, a0: int, a: int) -> Tensor:
    """Freq_aug (models/AASIST.py:486-490): rows [a0, a0+a) of the band-pass bank are zeroed."""
    out = bank.clone()
    out[a0:a0 + a, :] = 0
    return out


def draw_freq_mask(n_filters: int) -> Tuple[int, int]:
    """The reference's RNG draws, in its order (models/AASIST.py:487-489): numpy's global generator for the
    width, Python's ``random`` for the start.  Returns (A0, A)."""
    a = int(np.random.uniform(0, 20))
    a0 = random.randint(0, n_filters - a)
    return a0, a


# --------------------------------------------------------------------------- #
# fork Model.forward                      models/AASIST.py:806-921
# --------------------------------------------------------------------------- #
def aasist2_forward(sd: dict, cfg: dict, x: Tensor, speaker_embedding: Optional[Tensor] = None,
                    taps: Optional[dict] = None, bank: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    filts, ratios, temps = cfg["filts"], cfg["pool_ratios"], cfg["temperatures"]
    width, scale = cfg.get("res2net_width", 14), cfg.get("res2net_scale", 8)      # :739-740
    if bank is None:
        bank = O.sinc_filterbank(filts[0], cfg["first_conv"])
    T = taps if taps is not None else {}
    blocks = [filts[1], filts[2], filts[3], filts[4], filts[4], filts[4]]
    with torch.no_grad():
        e = O.frontend(x, bank, sd)
        T["frontend"] = e
        for i, nb in enumerate(blocks):                               # :766-772
            e = res2net_block(e, sd, f"encoder.{i}.0", nb, width, scale, first=(i == 0))
            T[f"encoder.{i}"] = e
        e_S = torch.max(torch.abs(e), dim=3)[0].transpose(1, 2) + sd["pos_S"]
        e_T = torch.max(torch.abs(e), dim=2)[0].transpose(1, 2)
        gat_S = O.gat_layer(e_S, sd, "GAT_layer_S", temps[0])
        out_S = O.graph_pool(gat_S, sd, "pool_S", ratios[0], 1, T)
        gat_T = O.gat_layer(e_T, sd, "GAT_layer_T", temps[1])
        out_T = O.graph_pool(gat_T, sd, "pool_T", ratios[1], 1, T)
        branches = []
        for br, (l1, l2) in (("1", ("HtrgGAT_layer_ST11", "HtrgGAT_layer_ST12")),
                             ("2", ("HtrgGAT_layer_ST21", "HtrgGAT_layer_ST22"))):
            oT, oS, m = O.htrg_gat_layer(out_T, out_S, sd["master" + br], sd, l1, temps[2])
            oS = O.graph_pool(oS, sd, "pool_hS" + br, ratios[2], 1, T)
            oT = O.graph_pool(oT, sd, "pool_hT" + br, ratios[2], 1, T)
            aT, aS, am = O.htrg_gat_layer(oT, oS, m, sd, l2, temps[2])
            branches.append((oT + aT, oS + aS, m + am))
        (T1, S1, m1), (T2, S2, m2) = branches
        out_T, out_S, master = torch.max(T1, T2), torch.max(S1, S2), torch.max(m1, m2)
        spk = bool(cfg.get("speaker_conditioning", False)) and speaker_embedding is not None
        level = cfg.get("conditioning_level", "frame")
        if spk and level == "frame":                                  # :895-900
            att = cfg.get("use_attention", True)
            out_T = speaker_conditioning(out_T, speaker_embedding, sd, "spk_cond_gat", level, att)
            out_S = speaker_conditioning(out_S, speaker_embedding, sd, "spk_cond_gat", level, att)
        T_max = torch.max(torch.abs(out_T), dim=1)[0]
        T_avg = torch.mean(out_T, dim=1)
        S_max = torch.max(torch.abs(out_S), dim=1)[0]
        S_avg = torch.mean(out_S, dim=1)
        last_hidden = torch.cat([T_max, T_avg, S_max, S_avg, master.squeeze(1)], dim=1)
        if spk and level == "utterance":                              # :913-916 (shape error in the reference)
            last_hidden = speaker_conditioning(last_hidden, speaker_embedding, sd, "spk_cond_gat", level, True)
        output = O._linear(last_hidden, sd, "out_layer")
        T["last_hidden"], T["output"] = last_hidden, output
    return last_hidden, output


# --------------------------------------------------------------------------- #
# AASIST-Robust Model.forward (eval)      models/AASIST_Robust.py:198-303
# --------------------------------------------------------------------------- #
ROBUST_TAPS, ROBUST_STRIDE = 1024, 256                               # AASIST_Robust.py:96-102


def robust_forward(sd: dict, cfg: dict, x: Tensor, taps: Optional[dict] = None,
                   bank: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Returns ``(ensemble_logits, logits)`` (AASIST_Robust.py:303).  Eval mode: the Gaussian-noise layer and the
    feature-denoising branch are training-only (:203-204, :230-235) and ``ensemble = softmax(w)[0]*logits +
    softmax(w)[1]*aux`` (:296-301)."""
    filts, ratios, temps = cfg["filts"], cfg["pool_ratios"], cfg["temperatures"]
    if bank is None:
        bank = O.sinc_filterbank(cfg["first_conv"], ROBUST_TAPS)      # CONV(out_channels=first_conv, kernel 1024)
    T = taps if taps is not None else {}
    with torch.no_grad():
        if x.dim() == 2:
            x = x.unsqueeze(1)
        y = F.conv1d(x, bank.view(bank.shape[0], 1, bank.shape[1]), stride=ROBUST_STRIDE)   # :217
        y = F.max_pool2d(torch.abs(y.unsqueeze(1)), (3, 3))           # :218-219
        e = F.selu(O._bn_eval(y, sd, "first_bn", 1))                  # :220-221
        T["frontend"] = e
        for i in range(6):
            e = residual_block_3x3(e, sd, f"encoder.{i}.0")           # :224
            T[f"encoder.{i}"] = e
        e_flat = e.mean(dim=(2, 3))                                   # :227
        e_S = torch.max(torch.abs(e), dim=3)[0].transpose(1, 2) + sd["pos_S"]    # :237-238
        gat_S = O.gat_layer(e_S, sd, "GAT_layer_S", temps[0])
        out_S = O.graph_pool(gat_S, sd, "pool_S", ratios[0], 1, T)
        e_T = torch.max(torch.abs(e), dim=2)[0].transpose(1, 2)       # :244-245
        gat_T = O.gat_layer(e_T, sd, "GAT_layer_T", temps[1])
        out_T = O.graph_pool(gat_T, sd, "pool_T", ratios[1], 1, T)
        oT, oS, m = O.htrg_gat_layer(out_T, out_S, sd["master1"], sd, "HtrgGAT_layer_ST1", temps[2])  # :254
        oS = O.graph_pool(oS, sd, "pool_hS", ratios[2], 1, T)         # :257
        oT = O.graph_pool(oT, sd, "pool_hT", ratios[3], 1, T)         # :258
        aT, aS, am = O.htrg_gat_layer(oT, oS, m, sd, "HtrgGAT_layer_ST2", temps[3])   # :261
        oT, oS = oT + aT, oS + aS                                     # :264-266
        T_max = torch.max(torch.abs(oT), dim=1)[0]
        T_avg = torch.mean(oT, dim=1)
        S_max = torch.max(torch.abs(oS), dim=1)[0]
        S_avg = torch.mean(oS, dim=1)
        out = torch.cat([T_max, T_avg, S_max, S_avg], dim=1)          # :283
        logits = O._linear(out, sd, "out_layer")                      # :287
        aux = O._linear(e_flat, sd, "aux_out_layer")                  # :290
        w = F.softmax(sd["ensemble_weight"], dim=0)                   # :293
        ensemble = w[0] * logits + w[1] * aux                         # :301
        T["hidden"], T["logits"], T["aux"], T["ensemble"] = out, logits, aux, ensemble
    return ensemble, logits


# --------------------------------------------------------------------------- #
# input staging                            data_utils.py:55-119
# --------------------------------------------------------------------------- #
def pad_sequence(seqs: List[np.ndarray]) -> np.ndarray:
    """Zero-pad to the batch maximum rounded up to a multiple of 4 (data_utils.py:100-119)."""
    max_len = max(s.shape[0] for s in seqs)
    max_len = ((max_len + 3) // 4) * 4                                # :110
    out = np.zeros((len(seqs), max_len), dtype=np.float32)            # :114
    for i, s in enumerate(seqs):
        n = min(s.shape[0], max_len)                                  # :117
        out[i, :n] = s[:n]
    return out


def chunk(x: np.ndarray, target_len: int, start: int) -> np.ndarray:
    """The deterministic part of ``dynamic_chunk_size`` / ``pad_random`` (data_utils.py:55-97) once the random
    draws (target length, crop start) are fixed: crop ``x[start:start+target]`` when long enough, else repeat-tile."""
    n = x.shape[0]
    if n >= target_len:
        return x[start:start + target_len]
    return np.tile(x, int(target_len / n) + 1)[:target_len]

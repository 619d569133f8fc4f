"""CPU oracle for the AASIST / RawGAT-ST utterance-scoring forward pass.

TEST INFRASTRUCTURE ONLY.  This file is a functional restatement (torch fp32 on
the CPU + numpy for the filter bank) of the reference's eval-mode forward.  It is
imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs -- never by the product path in
``aasist_b200/`` (which fails loudly when the CUDA library is missing).

Pinning: the reference ships no golden vectors or tests, so
this oracle is pinned against *outputs of the reference itself run in the build
container*: ``oracle/make_golden.py`` imports the reference classes from
``/root/reference`` (fork ``models/AASIST.py::Model`` with the checkpoint-compatible
``models/RawNetGatSpoofST.py::Residual_block`` encoder), runs them on
seeded inputs and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those fixtures, and
against the strict checkpoint-load + parameter-count facts (297 866 / 85 306,
reference README.md:63).

Every function cites the reference lines it restates (paths relative to the
reference repository root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SELU_ALPHA = 1.6732632423543772
SELU_SCALE = 1.0507009873554805
BN_EPS = 1e-5

# model_config dictionaries of the three configurations on the hot path
# (restated from config/AASIST.conf:13-21, config/AASIST-L.conf:13-21,
#  config/RawGATST_baseline.conf:12-17).
CONFIGS: Dict[str, dict] = {
    "AASIST": {
        "architecture": "AASIST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 64], [64, 64]],
        "gat_dims": [64, 32], "pool_ratios": [0.5, 0.7, 0.5, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
    },
    "AASIST-L": {
        "architecture": "AASIST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 24], [24, 24]],
        "gat_dims": [24, 32], "pool_ratios": [0.4, 0.5, 0.7, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
    },
    "RawGAT-ST": {
        "architecture": "RawNetGatSpoofST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 64], [64, 64]],
    },
}


# --------------------------------------------------------------------------- #
# a1: sinc filter bank                         models/AASIST.py:419-482
# --------------------------------------------------------------------------- #
def sinc_filterbank(n_filters: int, first_conv: int, sample_rate: int = 16000) -> Tensor:
    """Mel-spaced Hamming-windowed sinc band-pass bank, (n_filters, K) fp32.

    Restates ``CONV.__init__`` (models/AASIST.py:448-482) with the dtype chain the
    reference's mixed numpy/torch expression produces under numpy>=2:
    the sinc argument ``2*f*n/sr`` is float32 (torch tensor times a scalar),
    ``np.sinc`` of it is float32, the scaled difference is float64, and the final
    product with the float32 Hamming window is float32.
    """
    K = first_conv + 1 if first_conv % 2 == 0 else first_conv        # :449-450
    nfft = 512
    f = int(sample_rate / 2) * np.linspace(0, 1, int(nfft / 2) + 1)   # :461
    fmel = 2595 * np.log10(1 + f / 700)                               # :421,462
    edges_mel = np.linspace(np.min(fmel), np.max(fmel), n_filters + 1)  # :463-465
    edges_hz = 700 * (10 ** (edges_mel / 2595) - 1)                   # :425,466
    n = np.arange(-(K - 1) / 2, (K - 1) / 2 + 1).astype(np.float32)   # :469-470 (float32 tensor)
    window = np.hamming(K).astype(np.float32)                         # :481
    bank = np.zeros((n_filters, K), dtype=np.float32)
    sr32 = np.float32(sample_rate)
    for i in range(n_filters):                                        # :472-482
        fmin, fmax = edges_hz[i], edges_hz[i + 1]
        arg_hi = (n * np.float32(2 * fmax)) / sr32                    # fp32 tensor arithmetic
        arg_lo = (n * np.float32(2 * fmin)) / sr32
        h_hi = (2 * fmax / sample_rate) * np.sinc(arg_hi)             # float64 * float32 -> float64
        h_lo = (2 * fmin / sample_rate) * np.sinc(arg_lo)
        ideal = (h_hi - h_lo).astype(np.float32)                      # Tensor(hideal)
        bank[i] = window * ideal                                      # fp32 product
    return torch.from_numpy(bank)


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def _bn_eval(x: Tensor, sd: dict, prefix: str, channel_dim: int) -> Tensor:
    """Eval-mode batch norm with running statistics (nn.BatchNorm{1,2}d.eval()).
    For (B,N,D) inputs the reference flattens to (B*N, D) first (AASIST.py:99-105)."""
    args = (sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
            sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.1, BN_EPS)
    if channel_dim == 1:
        return F.batch_norm(x, *args)
    shape = x.shape
    return F.batch_norm(x.reshape(-1, shape[-1]), *args).view(shape)


def _linear(x: Tensor, sd: dict, prefix: str) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


# --------------------------------------------------------------------------- #
# a2 + a3: sinc conv, abs, 3x3 max-pool, first_bn, SELU
#                                   models/AASIST.py:484-503 and :816-831
# --------------------------------------------------------------------------- #
def frontend(x: Tensor, bank: Tensor, sd: dict) -> Tensor:
    """(B,L) -> (B,1,23,floor((L-K+1)/3)).  models/AASIST.py:816-831."""
    if x.dim() == 2:
        x = x.unsqueeze(1)                                            # :816-817
    y = F.conv1d(x, bank.view(bank.shape[0], 1, bank.shape[1]))      # :497-503
    y = y.unsqueeze(1)                                                # :826
    y = F.max_pool2d(torch.abs(y), (3, 3))                            # :829
    y = _bn_eval(y, sd, "first_bn", 1)                                # :830
    return F.selu(y)                                                  # :831


# --------------------------------------------------------------------------- #
# a4: (2,3) Residual_block            models/RawNetGatSpoofST.py:225-278
# --------------------------------------------------------------------------- #
def residual_block(x: Tensor, sd: dict, prefix: str) -> Tensor:
    """conv1 k(2,3) pad(1,1) -> bn2 -> SELU -> conv2 k(2,3) pad(0,1) -> + identity
    (conv_downsample k(1,3) pad(0,1) when channel counts differ) -> MaxPool2d((1,3)).
    ``bn1``/SELU on the input is dead code in the reference (its result is
    overwritten at :265) and is therefore not applied."""
    out = F.conv2d(x, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"],
                   padding=(1, 1))                                    # :265
    out = F.selu(_bn_eval(out, sd, prefix + ".bn2", 1))               # :268-269
    out = F.conv2d(out, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"],
                   padding=(0, 1))                                    # :271
    identity = x
    if (prefix + ".conv_downsample.weight") in sd:                    # :273-274
        identity = F.conv2d(x, sd[prefix + ".conv_downsample.weight"],
                            sd[prefix + ".conv_downsample.bias"], padding=(0, 1))
    out = out + identity                                              # :276
    return F.max_pool2d(out, (1, 3))                                  # :277


def encoder(x: Tensor, sd: dict, prefix: str = "encoder", taps: Optional[dict] = None) -> Tensor:
    for i in range(6):                                                # AASIST.py:766-772 nesting
        x = residual_block(x, sd, f"{prefix}.{i}.0")
        if taps is not None:
            taps[f"{prefix}.{i}"] = x
    return x


# --------------------------------------------------------------------------- #
# a6: GraphAttentionLayer             models/AASIST.py:17-110
#     (RawGAT-ST variant without temperature: RawNetGatSpoofST.py:10-94)
# --------------------------------------------------------------------------- #
def gat_layer(x: Tensor, sd: dict, prefix: str, temp: float = 1.0) -> Tensor:
    pair = x.unsqueeze(2) * x.unsqueeze(1)                            # :69-73  (B,N,N,D)
    att = torch.tanh(_linear(pair, sd, prefix + ".att_proj"))         # :82
    att = torch.matmul(att, sd[prefix + ".att_weight"])               # :84     (B,N,N,1)
    att = att / temp                                                  # :87
    att = F.softmax(att, dim=-2)                                      # :89
    agg = torch.matmul(att.squeeze(-1), x)                            # :94
    out = _linear(agg, sd, prefix + ".proj_with_att") + \
        _linear(x, sd, prefix + ".proj_without_att")                  # :94-97
    out = _bn_eval(out, sd, prefix + ".bn", 2)                        # :99-105
    return F.selu(out)                                                # :58


# --------------------------------------------------------------------------- #
# a8: HtrgGraphAttentionLayer         models/AASIST.py:113-282
# --------------------------------------------------------------------------- #
def htrg_gat_layer(x1: Tensor, x2: Tensor, master: Tensor, sd: dict, prefix: str,
                   temp: float) -> Tuple[Tensor, Tensor, Tensor]:
    n1, n2 = x1.size(1), x2.size(1)                                   # :155-156
    x = torch.cat([_linear(x1, sd, prefix + ".proj_type1"),
                   _linear(x2, sd, prefix + ".proj_type2")], dim=1)   # :158-161
    if master.size(0) != x.size(0):
        master = master.expand(x.size(0), -1, -1)
    # node-to-node attention                                          # :225-255
    pair = x.unsqueeze(2) * x.unsqueeze(1)
    t = torch.tanh(_linear(pair, sd, prefix + ".att_proj"))           # :232
    board = torch.zeros_like(t[..., :1])
    board[:, :n1, :n1] = torch.matmul(t[:, :n1, :n1], sd[prefix + ".att_weight11"])   # :237
    board[:, n1:, n1:] = torch.matmul(t[:, n1:, n1:], sd[prefix + ".att_weight22"])   # :239
    board[:, :n1, n1:] = torch.matmul(t[:, :n1, n1:], sd[prefix + ".att_weight12"])   # :241
    board[:, n1:, :n1] = torch.matmul(t[:, n1:, :n1], sd[prefix + ".att_weight12"])   # :243
    att = F.softmax(board / temp, dim=-2)                             # :251-253
    # master update (uses x after proj_type*, and the *input* master)  # :187-223, :263-269
    tm = torch.tanh(_linear(x * master, sd, prefix + ".att_projM"))   # :213-214
    am = F.softmax(torch.matmul(tm, sd[prefix + ".att_weightM"]) / temp, dim=-2)  # :216-221
    new_master = _linear(torch.matmul(am.squeeze(-1).unsqueeze(1), x), sd,
                         prefix + ".proj_with_attM") + \
        _linear(master, sd, prefix + ".proj_without_attM")            # :265-269
    # node projection                                                 # :257-261
    out = _linear(torch.matmul(att.squeeze(-1), x), sd, prefix + ".proj_with_att") + \
        _linear(x, sd, prefix + ".proj_without_att")
    out = F.selu(_bn_eval(out, sd, prefix + ".bn", 2))                # :179-180
    return out[:, :n1], out[:, n1:n1 + n2], new_master                # :182-185


# --------------------------------------------------------------------------- #
# a7: GraphPool                        models/AASIST.py:285-322
#     (RawGAT-ST: at least 2 nodes, RawNetGatSpoofST.py:126)
# --------------------------------------------------------------------------- #
def pooled_node_count(n_nodes: int, ratio: float, min_nodes: int = 1) -> int:
    """``max(int(n_nodes * k), 1)`` in Python double arithmetic (models/AASIST.py:315)."""
    return max(int(n_nodes * ratio), min_nodes)


def graph_pool(h: Tensor, sd: dict, prefix: str, ratio: float, min_nodes: int = 1,
               taps: Optional[dict] = None) -> Tensor:
    weights = _linear(h, sd, prefix + ".proj")                        # :296
    scores = torch.sigmoid(weights)                                   # :297
    k = pooled_node_count(h.size(1), ratio, min_nodes)                # :315
    _, idx = torch.topk(scores, k, dim=1)                             # :316
    if taps is not None:
        taps[prefix + ".weights"] = weights.squeeze(-1)
        taps[prefix + ".scores"] = scores.squeeze(-1)
        taps[prefix + ".idx"] = idx.squeeze(-1)
    hs = h * scores                                                   # :319
    return torch.gather(hs, 1, idx.expand(-1, -1, h.size(2)))         # :317,320


# --------------------------------------------------------------------------- #
# Model.forward                         models/AASIST.py:806-921
# --------------------------------------------------------------------------- #
def aasist_forward(sd: dict, cfg: dict, x: Tensor, taps: Optional[dict] = None,
                   bank: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Eval-mode ``Model.forward(x)`` (Freq_aug=False, speaker_embedding=None).

    ``sd`` is the shipped state_dict (models/weights/AASIST*.pth); ``cfg`` the
    ``model_config`` dict.  Returns ``(last_hidden (B,5*gat_dims[1]), output (B,2))``.
    ``taps`` (optional dict) receives every intermediate the parity tests compare.
    """
    filts, gat_dims = cfg["filts"], cfg["gat_dims"]
    ratios, temps = cfg["pool_ratios"], cfg["temperatures"]
    if bank is None:
        bank = sinc_filterbank(filts[0], cfg["first_conv"])
    T = taps if taps is not None else {}
    with torch.no_grad():
        z = frontend(x, bank, sd)                                     # :816-831
        T["frontend"] = z
        e = encoder(z, sd, "encoder", T)                              # :838
        e_S = torch.max(torch.abs(e), dim=3)[0].transpose(1, 2) + sd["pos_S"]   # :841-842
        e_T = torch.max(torch.abs(e), dim=2)[0].transpose(1, 2)       # :848-849
        T["e_S"], T["e_T"] = e_S, e_T
        gat_S = gat_layer(e_S, sd, "GAT_layer_S", temps[0])           # :844
        out_S = graph_pool(gat_S, sd, "pool_S", ratios[0], 1, T)      # :845
        gat_T = gat_layer(e_T, sd, "GAT_layer_T", temps[1])           # :851
        out_T = graph_pool(gat_T, sd, "pool_T", ratios[1], 1, T)      # :852
        T["gat_S"], T["gat_T"], T["out_S"], T["out_T"] = gat_S, gat_T, out_S, out_T

        branches = []
        for br, (l1, l2) in (("1", ("HtrgGAT_layer_ST11", "HtrgGAT_layer_ST12")),
                             ("2", ("HtrgGAT_layer_ST21", "HtrgGAT_layer_ST22"))):
            oT, oS, m = htrg_gat_layer(out_T, out_S, sd["master" + br], sd, l1, temps[2])  # :859,872
            T[f"{l1}.T"], T[f"{l1}.S"], T[f"{l1}.M"] = oT, oS, m
            oS = graph_pool(oS, sd, "pool_hS" + br, ratios[2], 1, T)  # :862,874
            oT = graph_pool(oT, sd, "pool_hT" + br, ratios[2], 1, T)  # :863,875
            aT, aS, am = htrg_gat_layer(oT, oS, m, sd, l2, temps[2])  # :865,877
            branches.append((oT + aT, oS + aS, m + am))               # :867-869, :879-881
        (T1, S1, m1), (T2, S2, m2) = branches
        out_T, out_S, master = torch.max(T1, T2), torch.max(S1, S2), torch.max(m1, m2)  # :890-892
        T_max = torch.max(torch.abs(out_T), dim=1)[0]                 # :903
        T_avg = torch.mean(out_T, dim=1)                              # :904
        S_max = torch.max(torch.abs(out_S), dim=1)[0]                 # :906
        S_avg = torch.mean(out_S, dim=1)                              # :907
        last_hidden = torch.cat([T_max, T_avg, S_max, S_avg, master.squeeze(1)], dim=1)  # :909-910
        output = _linear(last_hidden, sd, "out_layer")                # :919
        T["last_hidden"], T["output"] = last_hidden, output
    return last_hidden, output


# --------------------------------------------------------------------------- #
# a11: RawGAT-ST Model.forward         models/RawNetGatSpoofST.py:324-356
# --------------------------------------------------------------------------- #
def rawgat_st_forward(sd: dict, cfg: dict, x: Tensor, taps: Optional[dict] = None,
                      bank: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    filts = cfg["filts"]
    if bank is None:
        bank = sinc_filterbank(filts[0], cfg["first_conv"])
    T = taps if taps is not None else {}
    with torch.no_grad():
        z = frontend(x, bank, sd)                                     # :326-335
        T["frontend"] = z
        e_T = encoder(z, sd, "encoder_T", T)                          # :337
        e_T = torch.max(torch.abs(e_T), dim=3)[0]                     # :338 max along time
        T["e_T"] = e_T.transpose(1, 2)
        gat_T = gat_layer(e_T.transpose(1, 2), sd, "GAT_layer_T")     # :339
        pool_T = graph_pool(gat_T, sd, "pool_T", 0.64, 2, T)          # :340
        out_T = _linear(pool_T.transpose(1, 2), sd, "proj_T")         # :341
        e_S = encoder(z, sd, "encoder_S", T)                          # :343
        e_S = torch.max(torch.abs(e_S), dim=2)[0]                     # :344 max along freq
        T["e_S"] = e_S.transpose(1, 2)
        gat_S = gat_layer(e_S.transpose(1, 2), sd, "GAT_layer_S")     # :345
        pool_S = graph_pool(gat_S, sd, "pool_S", 0.81, 2, T)          # :346
        out_S = _linear(pool_S.transpose(1, 2), sd, "proj_S")         # :347
        T["gat_T"], T["gat_S"] = gat_T, gat_S
        g = torch.mul(out_T, out_S)                                   # :349
        T["gat_ST_in"] = g.transpose(1, 2)
        gat_ST = gat_layer(g.transpose(1, 2), sd, "GAT_layer_ST")     # :351
        pool_ST = graph_pool(gat_ST, sd, "pool_ST", 0.64, 2, T)       # :352
        proj_ST = _linear(pool_ST, sd, "proj_ST").flatten(1)          # :353
        output = _linear(proj_ST, sd, "out_layer")                    # :354
        T["gat_ST"], T["last_hidden"], T["output"] = gat_ST, proj_ST, output
    return proj_ST, output


def forward(kind: str, sd: dict, cfg: dict, x: Tensor, taps: Optional[dict] = None,
            bank: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    if cfg.get("architecture", kind) == "RawNetGatSpoofST" or kind == "RawGAT-ST":
        return rawgat_st_forward(sd, cfg, x, taps, bank)
    return aasist_forward(sd, cfg, x, taps, bank)


# --------------------------------------------------------------------------- #
# synthetic inputs -- deterministic across machines
# --------------------------------------------------------------------------- #
def white_noise(n_utt: int, length: int, seed: int, scale: float = 0.05) -> Tensor:
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.standard_normal((n_utt, length)) * scale).astype(np.float32))


def speech_like(n_utt: int, length: int, seed: int) -> Tensor:
    """Harmonic stack with random envelope/vibrato + weak noise; gives well separated
    temporal GraphPool scores (white noise yields near ties)."""
    g = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(length, dtype=np.float64) / 16000.0
    out = np.empty((n_utt, length), dtype=np.float32)
    for u in range(n_utt):
        f0 = 80.0 + 200.0 * g.random()
        knots = g.random(12)
        env = np.interp(np.linspace(0, 11, length), np.arange(12), knots)
        vib = 1.0 + 0.05 * np.sin(2 * np.pi * (3.0 + 4.0 * g.random()) * t)
        phi = 2 * np.pi * np.cumsum(f0 * vib) / 16000.0
        sig = np.zeros(length)
        for h in range(1, 25):
            sig += np.sin(h * phi + 2 * np.pi * g.random()) / h
        out[u] = (0.1 * env * sig + 0.003 * g.standard_normal(length)).astype(np.float32)
    return torch.from_numpy(out)


def pad(x: np.ndarray, max_len: int = 64600) -> np.ndarray:
    """Repeat-tile / crop one utterance to max_len samples (data_utils.py:45-52)."""
    x_len = x.shape[0]
    if x_len >= max_len:                                              # :47-48
        return x[:max_len]
    num_repeats = int(max_len / x_len) + 1                            # :50
    return np.tile(x, num_repeats)[:max_len]                          # :51 (tile of the 1-D signal)


def n_params(sd: dict) -> int:
    """Trainable parameter count of a state_dict (excludes BN running stats)."""
    return sum(v.numel() for k, v in sd.items()
               if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))


# --------------------------------------------------------------------------- #
# tie policy helper
# --------------------------------------------------------------------------- #
def compare_topk(ref_scores: Tensor, ref_idx: Tensor, got_idx: Tensor,
                 near_gap: float = 0.0) -> Tuple[int, int, int]:
    """Compare ordered top-k indices under the stated tie policy.

    Positions inside a run of oracle scores whose neighbours differ by <= ``near_gap``
    (0.0 = exact fp32 ties only) may hold any permutation of that run's indices; all
    other positions must match exactly.  The run is extended over the k / k+1
    boundary (a tie there changes the selected set).  Returns
    ``(n_mismatch, n_positions, n_tie_positions)``.
    """
    ref_scores, ref_idx, got_idx = ref_scores.cpu(), ref_idx.cpu().long(), got_idx.cpu().long()
    B, k = ref_idx.shape
    mism = ties = 0
    order = torch.argsort(ref_scores, dim=1, descending=True, stable=True)
    sorted_sc = torch.gather(ref_scores, 1, order)
    for b in range(B):
        s = sorted_sc[b].tolist()
        n = len(s)
        # group id per sorted position
        grp = [0] * n
        for p in range(1, n):
            grp[p] = grp[p - 1] + (0 if (s[p - 1] - s[p]) <= near_gap else 1)
        members: Dict[int, set] = {}
        for p in range(n):
            members.setdefault(grp[p], set()).add(int(order[b, p]))
        for p in range(k):
            allowed = members[grp[p]]
            if len(allowed) > 1:
                ties += 1
                if int(got_idx[b, p]) not in allowed:
                    mism += 1
            elif int(got_idx[b, p]) != int(ref_idx[b, p]):
                mism += 1
    return mism, B * k, ties

"""oracle/aasist2_oracle.py (Res2Net+SE encoder, speaker conditioning, Freq_aug, AASIST-Robust, staging) against
fixtures produced by the REFERENCE classes/functions themselves (oracle/make_golden_fork.py).  CPU only."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import aasist2_oracle as O2
from oracle import aasist_oracle as O
from tests.util import GOLD, WDIR, load_sd

AASIST_POOLS = ["pool_S", "pool_T", "pool_hS1", "pool_hT1", "pool_hS2", "pool_hT2"]
ROBUST_POOLS = ["pool_S", "pool_T", "pool_hS", "pool_hT"]


def load_fork(name):
    g = dict(np.load(os.path.join(GOLD, f"fork_{name}.npz")))
    meta = json.loads(str(g.pop("meta")))
    return g, meta


def fork_sd(model):
    return torch.load(os.path.join(WDIR, f"{model}_seed1234.pth"), map_location="cpu")


def _check_taps(taps, g, pools):
    for i in range(6):
        e = taps[f"encoder.{i}"]
        step = max(1, e.shape[3] // 16)
        ref = g[f"encoder.{i}.sample"]
        assert np.abs(e[:, ::5, :, ::step].numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), i
    for p in pools:
        assert np.abs(taps[p + ".weights"].numpy() - g[p + ".weights"]).max() <= 5e-5, p
        mism, _, _ = O.compare_topk(torch.from_numpy(g[p + ".weights"]), torch.from_numpy(g[p + ".idx"]),
                                    taps[p + ".idx"], near_gap=1e-5)
        assert mism == 0, p


@pytest.mark.parametrize("model,tag", [("AASIST2", "speech"), ("AASIST2", "speech24k"),
                                       ("AASIST2-small", "speech"), ("AASIST2-small", "speech24k")])
def test_res2net_model_oracle_matches_reference(model, tag):
    g, meta = load_fork(f"{model}_{tag}")
    x = O.speech_like(meta["n"], meta["L"], meta["seed"])
    assert np.array_equal(x[:, :8].numpy(), g["x_head"])
    sd, cfg = fork_sd(model), O2.CONFIGS[model]
    assert len(sd) == meta["n_tensors"]
    assert O.n_params(sd) == meta["n_params"]
    torch.set_num_threads(8)
    taps = {}
    lh, out = O2.aasist2_forward(sd, cfg, x, None, taps)
    assert np.abs(out.numpy() - g["output"]).max() <= 5e-5
    assert np.abs(lh.numpy() - g["last_hidden"]).max() <= 5e-5
    _check_taps(taps, g, AASIST_POOLS)
    emb = torch.from_numpy(g["spk_embedding"])
    lh_s, out_s = O2.aasist2_forward(sd, cfg, x, emb)
    assert np.abs(out_s.numpy() - g["spk.output"]).max() <= 5e-5
    assert np.abs(lh_s.numpy() - g["spk.last_hidden"]).max() <= 5e-5
    assert np.abs(out_s.numpy() - out.numpy()).max() > 1e-3            # the conditioning does something


def test_res2net_split_bookkeeping():
    assert O2.res2net_splits([1, 32], 14, 8) == ([1], 1)
    assert O2.res2net_splits([32, 32], 14, 8) == ([2] * 13 + [6], 8)
    assert O2.res2net_splits([64, 64], 14, 8) == ([4] * 13 + [12], 8)
    assert O2.res2net_splits([24, 24], 14, 8) == ([1] * 13 + [11], 8)
    assert O2.res2net_splits([32, 24], 6, 2) == ([5] * 5 + [7], 2)


def test_utterance_level_conditioning_fails_like_the_reference():
    sd, cfg = fork_sd("AASIST2"), dict(O2.CONFIGS["AASIST2"], conditioning_level="utterance")
    with pytest.raises(RuntimeError):              # reference: mat1 and mat2 shapes cannot be multiplied
        O2.aasist2_forward(sd, cfg, O.speech_like(1, 16000, 1), torch.zeros(1, 256))


def test_freq_aug_oracle_matches_reference():
    g, meta = load_fork("freqaug")
    x = O.speech_like(meta["n"], meta["L"], meta["seed"])
    sd, cfg = load_sd("AASIST"), O.CONFIGS["AASIST"]
    bank = O.sinc_filterbank(70, 128)
    torch.set_num_threads(8)
    for seed, a0, a in meta["masks"]:
        np.random.seed(seed)
        random.seed(seed)
        assert O2.draw_freq_mask(70) == (a0, a)                        # same draws as AASIST.py:487-489
        lh, out = O.forward("AASIST", sd, cfg, x, None, O2.mask_filterbank(bank, a0, a))
        assert np.abs(out.numpy() - g[f"seed{seed}.output"]).max() <= 5e-5, seed
        assert np.abs(lh.numpy() - g[f"seed{seed}.last_hidden"]).max() <= 5e-5, seed


@pytest.mark.parametrize("tag", ["nt1", "nt3"])
def test_robust_oracle_matches_reference(tag):
    g, meta = load_fork(f"robust_{tag}")
    x = O.speech_like(meta["n"], meta["L"], meta["seed"])
    sd, cfg = fork_sd("AASIST-Robust"), O2.CONFIGS["AASIST-Robust"]
    assert len(sd) == meta["n_tensors"]
    torch.set_num_threads(8)
    taps = {}
    ens, logits = O2.robust_forward(sd, cfg, x, taps)
    assert np.abs(ens.numpy() - g["ensemble"]).max() <= 5e-5
    assert np.abs(logits.numpy() - g["logits"]).max() <= 5e-5
    _check_taps(taps, g, ROBUST_POOLS)


def test_staging_oracle_matches_reference_functions():
    from oracle.make_golden_fork import CHUNK_CASES, STAGING_LENGTHS
    g = np.load(os.path.join(GOLD, "fork_staging.npz"))
    seqs = [O.white_noise(1, n, seed)[0].numpy() for n, seed in STAGING_LENGTHS]
    X = O2.pad_sequence(seqs)
    assert list(X.shape) == g["pad_sequence.shape"].tolist() == [6, 96000]
    assert np.array_equal(X[:, ::997], g["pad_sequence.sample"])
    assert np.array_equal(X[:, -8:], g["pad_sequence.tail"])
    assert np.array_equal(X.astype(np.float64).sum(axis=1), g["pad_sequence.rowsum"])
    for n, seed in CHUNK_CASES:
        x = O.white_noise(1, n, seed)[0].numpy()
        target, start = g[f"chunk{n}.target_start"].tolist()
        y = O2.chunk(x, target, start)
        assert np.array_equal(y[::499], g[f"chunk{n}.sample"])
        assert y.astype(np.float64).sum() == float(g[f"chunk{n}.sum"])

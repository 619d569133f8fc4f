"""The CPU oracle (oracle/aasist_oracle.py) against fixtures produced by the REFERENCE
classes themselves (oracle/make_golden.py, run in the build container).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import aasist_oracle as O
from tests.util import golden_input, load_golden, load_sd, pools_of

CASES = [("AASIST", "white"), ("AASIST", "speech"), ("AASIST", "speech16k"), ("AASIST", "speech96k"),
         ("AASIST", "speech128k"), ("AASIST", "speech256k"),
         ("AASIST-L", "white"), ("AASIST-L", "speech"), ("AASIST-L", "speech16k"), ("AASIST-L", "speech96k"),
         ("AASIST-L", "speech192k"),
         ("RawGAT-ST", "white"), ("RawGAT-ST", "speech")]


def test_param_counts_match_reference_readme():
    # reference README.md:63 publishes 85,306 for AASIST-L; 297,866 measured on the reference class
    assert O.n_params(load_sd("AASIST")) == 297866
    assert O.n_params(load_sd("AASIST-L")) == 85306
    assert O.n_params(load_sd("RawGAT-ST")) == 437034


def test_filterbank_bit_exact_vs_reference():
    g, _ = load_golden("AASIST", "white")
    bank = O.sinc_filterbank(70, 128).numpy()
    assert bank.shape == (70, 129)
    assert np.array_equal(bank, g["bank"])
    # exactly symmetric taps (correlation == convolution)
    assert np.array_equal(bank, bank[:, ::-1])


@pytest.mark.parametrize("model,tag", CASES)
def test_oracle_matches_reference_outputs(model, tag):
    g, meta = load_golden(model, tag)
    x = golden_input(meta)
    assert np.array_equal(x[:, :8].numpy(), g["x_head"])          # generator is reproducible
    assert np.allclose(x.double().sum(dim=1).numpy(), g["x_sum"], rtol=0, atol=1e-9)
    sd, cfg = load_sd(model), O.CONFIGS[model]
    torch.set_num_threads(8)
    taps = {}
    last_hidden, output = O.forward(model, sd, cfg, x, taps)
    # same ATen kernels, same operation order -> agreement to fp32 rounding
    assert np.abs(output.numpy() - g["output"]).max() <= 2e-5
    assert np.abs(last_hidden.numpy() - g["last_hidden"]).max() <= 2e-5
    assert np.abs(taps["frontend"][:, 0, :, ::211].numpy() - g["frontend_sample"]).max() <= 1e-5
    prefixes = ["encoder_T", "encoder_S"] if model == "RawGAT-ST" else ["encoder"]
    for pre in prefixes:
        e = taps[f"{pre}.5"].numpy()
        assert np.abs(e - g[f"{pre}.5.full"]).max() <= 1e-4 * max(1.0, np.abs(e).max())
    for p in pools_of(model):
        w = taps[p + ".weights"].numpy()
        assert np.abs(w - g[p + ".weights"]).max() <= 5e-5
        mism, total, ties = O.compare_topk(torch.from_numpy(g[p + ".weights"]),
                                           torch.from_numpy(g[p + ".idx"]), taps[p + ".idx"],
                                           near_gap=1e-6)
        assert mism == 0, (p, mism, total, ties)


def test_pooled_node_count_python_double_semantics():
    # int(N*k) in Python double arithmetic (models/AASIST.py:315)
    assert O.pooled_node_count(29, 0.7) == 20
    assert O.pooled_node_count(14, 0.7) == 9
    assert O.pooled_node_count(9, 0.7) == 6
    assert O.pooled_node_count(23, 0.5) == 11
    assert O.pooled_node_count(1, 0.4) == 1
    assert O.pooled_node_count(23, 0.64, 2) == 14 and O.pooled_node_count(29, 0.81, 2) == 23


def test_compare_topk_tie_policy():
    s = torch.tensor([[0.9, 0.5, 0.5, 0.1]])
    ref = torch.tensor([[0, 1]])
    assert O.compare_topk(s, ref, torch.tensor([[0, 2]]))[0] == 0     # tie across the k boundary
    assert O.compare_topk(s, ref, torch.tensor([[1, 0]]))[0] == 2     # strict order violated
    assert O.compare_topk(s, ref, torch.tensor([[0, 3]]))[0] == 1


def test_pad_oracle_matches_reference_pad():
    # fixtures produced by the reference's own `pad` (data_utils.py:45-52), see oracle/make_golden.py
    import os
    from oracle.make_golden import PAD_CASES
    from tests.util import GOLD
    g = np.load(os.path.join(GOLD, "pad.npz"))
    for n, seed in PAD_CASES:
        x = O.white_noise(1, n, seed)[0].numpy()
        y = O.pad(x, 64600)
        assert y.shape == (64600,) and y.dtype == np.float32
        assert np.array_equal(y[::499], g[f"len{n}.sample"])
        assert np.array_equal(y[-16:], g[f"len{n}.tail"])
        assert y.astype(np.float64).sum() == float(g[f"len{n}.sum"])

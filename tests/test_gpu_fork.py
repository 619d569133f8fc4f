"""The fork-only rows of SURVEY 8(f) on the GPU, through the C ABI, against goldens produced by the REFERENCE
classes/functions (oracle/make_golden_fork.py) and the CPU oracle (oracle/aasist2_oracle.py):
Res2Net+SE encoder, speaker conditioning, Freq_aug, AASIST-Robust, pad_sequence / dynamic chunks, the pipelined
scoring loop, and a GPU-eager (torch CUDA fp32, TF32 off) second witness."""
import json
import os
import random

import numpy as np
import pytest
import torch

import aasist_b200
from oracle import aasist2_oracle as O2
from oracle import aasist_oracle as O
from tests.test_oracle_fork import AASIST_POOLS, ROBUST_POOLS, fork_sd, load_fork
from tests.util import GOLD, load_sd

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _g():
    from tests import gpu_util
    return gpu_util


def _pools(m, L):
    g = _g()
    return g.split_pools(m.last_topk, m.last_pool_weights, m.topk_layout(L))


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
@pytest.mark.parametrize("model,tag", [("AASIST2", "speech"), ("AASIST2", "speech24k"),
                                       ("AASIST2-small", "speech"), ("AASIST2-small", "speech24k")])
def test_res2net_model_full_forward_against_reference_golden(model, tag, precision):
    """Fork `Model(d_args)` with the Res2Net+SE encoder (AASIST.py:766-772), with and without a speaker embedding
    (AASIST.py:895-900).  precision f16x3 = tensor-core sinc front end + fp32 Res2Net kernels."""
    g = _g()
    gold, meta = load_fork(f"{model}_{tag}")
    x = O.speech_like(meta["n"], meta["L"], meta["seed"]).to(g.DEV)
    m = g.native_model(model, precision)
    m.record_topk = True
    try:
        lh, out = m(x)
        torch.cuda.synchronize()
        pools = _pools(m, meta["L"])
        lh_s, out_s = m(x, speaker_embedding=torch.from_numpy(gold["spk_embedding"]))
        torch.cuda.synchronize()
    finally:
        m.record_topk = False
    err = np.abs(out.cpu().numpy() - gold["output"]).max()
    herr = np.abs(lh.cpu().numpy() - gold["last_hidden"]).max()
    serr = np.abs(out_s.cpu().numpy() - gold["spk.output"]).max()
    sherr = np.abs(lh_s.cpu().numpy() - gold["spk.last_hidden"]).max()
    rep = g.check_pools(pools, gold, AASIST_POOLS)
    print(json.dumps({"case": f"{model}/{tag}/{precision}", "logit_err": float(err), "hidden_err": float(herr),
                      "spk_logit_err": float(serr), "spk_hidden_err": float(sherr),
                      "strict_mismatch": sum(r["strict_mismatch"] for r in rep.values()),
                      "near_tie_positions": sum(r["near_tie_positions"] for r in rep.values())}))
    tol = TOL if precision == "fp32" else 2e-4
    assert err <= tol and herr <= tol and serr <= tol and sherr <= tol, (err, herr, serr, sherr)
    for p, r in rep.items():
        assert r["weights_err"] <= 2e-4 and r["mismatch_outside_near_ties"] == 0, (p, r)


@pytest.mark.parametrize("model", ["AASIST2", "AASIST2-small"])
def test_res2net_blocks_stagewise_against_oracle(model):
    g = _g()
    x = O.speech_like(2, 24000, 5)
    taps = {}
    torch.set_num_threads(8)
    O2.aasist2_forward(fork_sd(model), O2.CONFIGS[model], x, None, taps)
    m = g.native_model(model, "fp32")
    f = O2.CONFIGS[model]["filts"]
    chans = [f[1], f[2], f[3], f[4], f[4], f[4]]
    inp = taps["frontend"]
    for i in range(6):
        ref = taps[f"encoder.{i}"]
        out = g.stage_block(m, 0, i, inp.to(g.DEV), chans[i][1])
        err = (out.cpu() - ref).abs().max().item()
        assert err <= 2e-5 * max(1.0, ref.abs().max().item()), (i, err)
        inp = ref


def test_res2net_batch_invariance_and_chunking():
    g = _g()
    m = g.native_model("AASIST2-small", "fp32")
    x = O.speech_like(3, 20000, 77).to(g.DEV)
    big = x.repeat(12, 1)                                  # 36 utterances: crosses the 32-utterance fp32 pass
    out = m(big)[1]
    ref = m(x)[1]
    assert torch.equal(out.view(12, 3, 2), ref.unsqueeze(0).expand(12, 3, 2))
    assert torch.equal(m(x)[1], ref)                       # deterministic (fixed-order SE reduction)


def test_speaker_conditioning_contract():
    g = _g()
    m = g.native_model("AASIST2", "fp32")
    x = O.speech_like(2, 16000, 3).to(g.DEV)
    with pytest.raises(RuntimeError):                      # wrong embedding width
        m(x, speaker_embedding=torch.zeros(2, 100))
    plain = g.native_model("AASIST", "fp32")               # no module: the embedding is ignored (AASIST.py:895)
    assert torch.equal(plain(x, speaker_embedding=torch.zeros(2, 256))[1], plain(x)[1])
    cfg = dict(aasist_b200.CONFIGS["AASIST2"], conditioning_level="utterance")
    mu = aasist_b200.Model(cfg, precision="fp32")
    mu.load_state_dict(load_sd("AASIST2"), strict=True)
    mu = mu.to(g.DEV).eval()
    mu(x)                                                  # without an embedding it runs
    with pytest.raises(RuntimeError, match="cannot be multiplied"):   # the reference's own failure (:913-916)
        mu(x, speaker_embedding=torch.zeros(2, 256))


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
def test_freq_aug_matches_reference_under_the_same_seeds(precision):
    """Model.forward(x, Freq_aug=True): the mask rows are drawn from numpy's / Python's global generators like
    AASIST.py:487-489 and zeroed in the device filter image."""
    g = _g()
    gold, meta = load_fork("freqaug")
    x = O.speech_like(meta["n"], meta["L"], meta["seed"]).to(g.DEV)
    m = g.native_model("AASIST", precision)
    clean = m(x)[1].clone()
    tol = TOL if precision == "fp32" else 2e-4
    for seed, a0, a in meta["masks"]:
        np.random.seed(seed)
        random.seed(seed)
        lh, out = m(x, Freq_aug=True)
        assert m.last_freq_mask == (a0, a)
        err = np.abs(out.cpu().numpy() - gold[f"seed{seed}.output"]).max()
        herr = np.abs(lh.cpu().numpy() - gold[f"seed{seed}.last_hidden"]).max()
        assert err <= tol and herr <= tol, (seed, err, herr)
    assert torch.equal(m(x)[1], clean)                     # the mask does not leak into later calls


@pytest.mark.parametrize("tag", ["nt1", "nt3"])
def test_robust_model_against_reference_golden(tag):
    g = _g()
    gold, meta = load_fork(f"robust_{tag}")
    x = O.speech_like(meta["n"], meta["L"], meta["seed"]).to(g.DEV)
    m = g.native_model("AASIST-Robust", "fp32")
    m.record_topk = True
    try:
        ens, logits = m(x)
        torch.cuda.synchronize()
        pools = _pools(m, meta["L"])
    finally:
        m.record_topk = False
    assert ens.shape == (meta["n"], 2) and logits.shape == (meta["n"], 2)
    e1 = np.abs(ens.cpu().numpy() - gold["ensemble"]).max()
    e2 = np.abs(logits.cpu().numpy() - gold["logits"]).max()
    rep = g.check_pools(pools, gold, ROBUST_POOLS)
    print(json.dumps({"case": f"robust/{tag}", "ensemble_err": float(e1), "logit_err": float(e2)}))
    assert e1 <= TOL and e2 <= TOL, (e1, e2)
    for p, r in rep.items():
        assert r["weights_err"] <= 1e-4 and r["mismatch_outside_near_ties"] == 0, (p, r)


def test_robust_blocks_stagewise_and_reference_errors():
    g = _g()
    x = O.speech_like(1, 600000, 9)
    taps = {}
    torch.set_num_threads(8)
    O2.robust_forward(fork_sd("AASIST-Robust"), O2.CONFIGS["AASIST-Robust"], x, taps)
    m = g.native_model("AASIST-Robust", "fp32")
    front = g.stage_frontend_raw(m, x.to(g.DEV), taps["frontend"].shape)
    assert (front.cpu() - taps["frontend"]).abs().max().item() <= 2e-4 * max(1.0, taps["frontend"].abs().max().item())
    inp = taps["frontend"]
    for i in range(6):
        ref = taps[f"encoder.{i}"]
        out = g.stage_block(m, 0, i, inp.to(g.DEV), ref.shape[1])
        assert (out.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), i
        inp = ref
    with pytest.raises(RuntimeError, match="too short|too small"):     # reference: max_pool2d output size too small
        m(torch.zeros(1, 64600, device=g.DEV))
    bad = aasist_b200.RobustModel(dict(aasist_b200.CONFIGS["AASIST-Robust"], first_conv=128), precision="fp32")
    bad = bad.to(g.DEV).eval()
    with pytest.raises(RuntimeError, match="must match the size of tensor b"):   # AASIST_Robust.py:237-238
        bad(torch.zeros(1, 600000, device=g.DEV))


def test_pad_sequence_and_dynamic_chunks_bit_exact():
    from oracle.make_golden_fork import CHUNK_CASES, STAGING_LENGTHS
    g = _g()
    gold = np.load(os.path.join(GOLD, "fork_staging.npz"))
    m = g.native_model("AASIST-L")
    seqs = [O.white_noise(1, n, seed)[0] for n, seed in STAGING_LENGTHS]
    X, y, dur = m.pad_sequence([(s, i % 2, float(s.numel()) / 16000) for i, s in enumerate(seqs)])
    Xc = X.cpu().numpy()
    assert list(Xc.shape) == gold["pad_sequence.shape"].tolist()
    assert np.array_equal(Xc, O2.pad_sequence([s.numpy() for s in seqs]))           # byte work: bit-exact
    assert np.array_equal(Xc[:, ::997], gold["pad_sequence.sample"])
    assert np.array_equal(Xc[:, -8:], gold["pad_sequence.tail"])
    assert y.tolist() == [0, 1, 0, 1, 0, 1] and dur.shape == (6,)
    # dynamic_chunk_size: same numpy draws as the reference, crop / tile on the device
    for n, seed in CHUNK_CASES:
        x = O.white_noise(1, n, seed)[0]
        np.random.seed(seed)
        out, durations = m.dynamic_chunks([x], 16000, 96000)
        target, start = gold[f"chunk{n}.target_start"].tolist()
        ref = O2.chunk(x.numpy(), target, start)
        got = out.cpu().numpy()[0]
        assert abs(float(durations[0]) - target / 16000) < 1e-6
        assert got.shape[0] == ((target + 3) // 4) * 4
        assert np.array_equal(got[:target], ref) and not got[target:].any()
        assert np.array_equal(got[:target][::499], gold[f"chunk{n}.sample"])


@pytest.mark.parametrize("pinned", [True, False])
def test_pipelined_host_scoring_equals_device_forward(pinned):
    """aasist_score_begin/submit/finish (double-buffered H2D under the previous forward, one wait at the end) gives
    bit-identical scores to per-batch device forwards, including a ragged last batch."""
    g = _g()
    from aasist_b200.scoring import score_utterances
    m = g.native_model("AASIST-L", "f16x3")
    x = O.speech_like(11, 64600, 61)
    ref = torch.cat([m(x[i:i + 4].to(g.DEV))[1] for i in range(0, 11, 4)])[:, 1]
    src = x.pin_memory() if pinned else x
    scores = score_utterances(m, src, 11, batch_size=4)
    assert scores.is_cuda and torch.equal(scores, ref)
    # the stream API directly, with the hidden vectors, twice in a row (buffers are reused)
    for _ in range(2):
        m.score_begin(11, 4, 64600)
        for i in range(0, 11, 4):
            m.score_submit(src[i:i + 4])
        hid, out = m.score_finish(want_hidden=True)
        assert not out.is_cuda and torch.equal(out[:, 1], ref.cpu()) and hid.shape == (11, 160)


def test_gpu_eager_reference_is_a_second_witness():
    """the reference's real deployment is torch on CUDA.  The oracle's forward run on the GPU in fp32
    (TF32 off) on FRESH utterances -- no golden involved -- against the tensor-core path: logits within tolerance and
    ordered GraphPool indices equal outside near-ties."""
    g = _g()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    name = "AASIST"
    x = O.speech_like(48, 64600, 4242)
    sd = {k: v.to(g.DEV) for k, v in load_sd(name).items()}
    bank = O.sinc_filterbank(70, 128).to(g.DEV)
    taps = {}
    ref_h, ref_o = O.forward(name, sd, O.CONFIGS[name], x.to(g.DEV), taps, bank)
    m = g.native_model(name, "f16x3")
    m.record_topk = True
    try:
        lh, out = m(x.to(g.DEV))
        torch.cuda.synchronize()
        pools = _pools(m, 64600)
    finally:
        m.record_topk = False
    err = (out - ref_o).abs().max().item()
    herr = (lh - ref_h).abs().max().item()
    ref = {k: v.cpu() for k, v in taps.items() if k.endswith((".weights", ".idx"))}
    rep = g.check_pools(pools, ref, AASIST_POOLS)
    print(json.dumps({"witness": "torch-cuda-fp32", "n": 48, "logit_err": err, "hidden_err": herr,
                      "strict_mismatch": sum(r["strict_mismatch"] for r in rep.values()),
                      "exact_tie_positions": sum(r["exact_tie_positions"] for r in rep.values()),
                      "near_tie_positions": sum(r["near_tie_positions"] for r in rep.values())}))
    assert err <= 2e-4 and herr <= 2e-4, (err, herr)
    for p, r in rep.items():
        assert r["mismatch_outside_near_ties"] == 0, (p, r)

"""Tensor-core path (precision="f16x3": tcgen05, fp16 hi/lo split operands, fp32 accumulate in TMEM)
against the CPU oracle / golden fixtures.  Tolerance on logits: north_star allows 1e-3 for
reduced-precision tensor paths; the split scheme is expected (CPU emulation) to reach ~1e-5."""
import json

import numpy as np
import pytest
import torch

from oracle import aasist_oracle as O
from tests.util import golden_input, load_golden, pools_of

pytestmark = pytest.mark.gpu

TC_LOGIT_TOL = 2e-4
CASES = [("AASIST", "white"), ("AASIST", "speech"), ("AASIST", "speech16k"), ("AASIST", "speech96k"),
         ("AASIST", "speech128k"), ("AASIST", "speech192k"), ("AASIST", "speech256k"),
         ("AASIST-L", "white"), ("AASIST-L", "speech"), ("AASIST-L", "speech16k"), ("AASIST-L", "speech96k"),
         ("AASIST-L", "speech128k"), ("AASIST-L", "speech256k"),
         ("RawGAT-ST", "white"), ("RawGAT-ST", "speech")]


def _g():
    from tests import gpu_util
    return gpu_util


@pytest.mark.parametrize("name", ["AASIST", "AASIST-L"])
def test_tc_sinc_frontend_stage(name):
    g = _g()
    x = O.speech_like(2, 64600, 3)
    x[1] *= 8.0                                   # louder utterance: exercises the fp16 operand range
    taps = g.oracle_taps(name, x)
    out = g.stage_frontend(g.native_model(name, "f16x3"), x.to(g.DEV))
    ref = taps["frontend"]
    err = (out.cpu() - ref).abs().max().item()
    print(json.dumps({"model": name, "frontend_abs_err": err, "ref_max": ref.abs().max().item()}))
    # first_bn multiplies the pooled |sinc| output by ~121: 1e-4 here is ~1e-6 on the conv output
    assert err <= 2e-4 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("name", ["AASIST", "AASIST-L", "RawGAT-ST"])
def test_tc_encoder_blocks_stagewise(name):
    g = _g()
    x = O.speech_like(2, 64600, 4)
    taps = g.oracle_taps(name, x)
    m = g.native_model(name, "f16x3")
    f = O.CONFIGS[name]["filts"]
    chans = [f[1], f[2], f[3], f[4], f[4], f[4]]
    report = {}
    # RawGAT-ST runs the same tensor-core kernels over two encoders with different weights
    for e, pre in enumerate(["encoder_T", "encoder_S"] if name == "RawGAT-ST" else ["encoder"]):
        inp = taps["frontend"]
        for i in range(6):
            ref = taps[f"{pre}.{i}"]
            out = g.stage_block(m, e, i, inp.to(g.DEV), chans[i][1])
            err = (out.cpu() - ref).abs().max().item()
            report[f"{pre}.{i}"] = err / max(1.0, ref.abs().max().item())
            inp = ref
    print(json.dumps({"model": name, "block_rel_err": report}))
    for i, e in report.items():
        # fp16-pair operands: ~2^-21 relative per product, fp32 accumulation
        assert e <= 2e-5, (i, report)


@pytest.mark.parametrize("name,tag", CASES)
def test_tc_full_forward_against_golden(name, tag):
    g = _g()
    gold, meta = load_golden(name, tag)
    x = golden_input(meta).to(g.DEV)
    m = g.native_model(name, "f16x3")
    m.record_topk = True
    try:
        last_hidden, output = m(x)
        torch.cuda.synchronize()
        pools = g.split_pools(m.last_topk, m.last_pool_weights, m.topk_layout(meta["L"]))
    finally:
        m.record_topk = False
    err = np.abs(output.cpu().numpy() - gold["output"]).max()
    herr = np.abs(last_hidden.cpu().numpy() - gold["last_hidden"]).max()
    rep = g.check_pools(pools, gold, pools_of(name))
    print(json.dumps({"case": f"{name}/{tag}", "logit_err": float(err), "hidden_err": float(herr), "pools": rep}))
    assert err <= TC_LOGIT_TOL and herr <= TC_LOGIT_TOL, (err, herr)
    for p, r in rep.items():
        assert r["weights_err"] <= 2e-4, (p, r)
        assert r["mismatch_outside_near_ties"] == 0, (p, r)
    if tag.startswith("speech"):
        # on speech-like input every ordered index equals the reference's wherever its
        # neighbouring scores differ at all in fp32 (only exact ties may permute) -- on the path that ships
        assert sum(r["strict_mismatch"] for r in rep.values()) == 0, rep


def _tiled_golden(name, tag, B):
    gold, meta = load_golden(name, tag)
    x = golden_input(meta)
    reps = (B + meta["n"] - 1) // meta["n"]
    return gold, meta, x.repeat(reps, 1)[:B].contiguous()


@pytest.mark.parametrize("name,B", [("AASIST", 640), ("AASIST", 1030), ("AASIST-L", 1030)])
def test_tc_bench_batch_and_pass_boundary_against_golden(name, B):
    """The benchmarked shape and beyond: B > 512 crosses the encoder-pass boundary (csrc/encoder_tc.cu processes
    at most 512 utterances / 40 GB per pass).  The four golden speech utterances are tiled to B rows and EVERY row
    is compared with the reference's golden logits, hidden vector and ordered GraphPool indices."""
    g = _g()
    gold, meta, x = _tiled_golden(name, "speech", B)
    n = meta["n"]
    m = g.native_model(name, "f16x3")
    m.record_topk = True
    try:
        last_hidden, output = m(x.to(g.DEV))
        torch.cuda.synchronize()
        topk, weights = m.last_topk.clone(), m.last_pool_weights.clone()
    finally:
        m.record_topk = False
    ref_out = np.tile(gold["output"], ((B + n - 1) // n, 1))[:B]
    ref_hid = np.tile(gold["last_hidden"], ((B + n - 1) // n, 1))[:B]
    err = np.abs(output.cpu().numpy() - ref_out).max()
    herr = np.abs(last_hidden.cpu().numpy() - ref_hid).max()
    print(json.dumps({"case": f"{name}/speech x{B}", "logit_err": float(err), "hidden_err": float(herr)}))
    assert err <= TC_LOGIT_TOL and herr <= TC_LOGIT_TOL, (err, herr)
    # every copy of an utterance scores bit-identically wherever it sits in the batch / pass
    assert torch.equal(output[n:2 * n], output[:n])
    for r in range(0, B - n + 1, n):
        assert torch.equal(output[r:r + n], output[:n]), r
        assert torch.equal(topk[r:r + n], topk[:n]), r
    pools = g.split_pools(topk[:n], weights[:n], m.topk_layout(meta["L"]))
    rep = g.check_pools(pools, gold, pools_of(name))
    assert sum(r["strict_mismatch"] for r in rep.values()) == 0, rep


@pytest.mark.parametrize("B,pinned", [(130, True), (130, False), (512, True), (512, False)])
def test_tc_forward_host_two_piece_path_equals_device_forward(B, pinned):
    """aasist_forward_host splits B > 128 into 128 + rest on two streams (csrc/api.cu) -- the path the bench's e2e
    figure times: bit-equal to the device forward and within tolerance of the golden logits."""
    g = _g()
    gold, meta, x = _tiled_golden("AASIST", "speech", B)
    m = g.native_model("AASIST", "f16x3")
    lh_d, out_d = m(x.to(g.DEV))
    xh = x.pin_memory() if pinned else x
    lh_h, out_h = m.score_host(xh)
    assert torch.equal(out_d.cpu(), out_h) and torch.equal(lh_d.cpu(), lh_h)
    n = meta["n"]
    ref_out = np.tile(gold["output"], ((B + n - 1) // n, 1))[:B]
    assert np.abs(out_h.numpy() - ref_out).max() <= TC_LOGIT_TOL


def test_tc_batch_invariance_and_fp32_agreement():
    g = _g()
    m = g.native_model("AASIST", "f16x3")
    m32 = g.native_model("AASIST", "fp32")
    x = O.speech_like(5, 64600, 21).to(g.DEV)
    out = m(x)[1]
    assert torch.equal(out, m(x)[1])                                   # deterministic
    single = torch.cat([m(x[i:i + 1])[1] for i in range(5)])
    assert torch.equal(out, single)                                    # shard invariance
    assert (out - m32(x)[1]).abs().max().item() <= TC_LOGIT_TOL
    m(torch.zeros(1, 2315, device=g.DEV))                              # shortest valid length
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2000, device=g.DEV))


def test_tc_larger_batch_crosses_chunk_boundary():
    g = _g()
    m = g.native_model("AASIST-L", "f16x3")
    x = O.white_noise(3, 64600, 5).to(g.DEV)
    big = x.repeat(50, 1)                                              # 150 utterances > one 128-chunk
    out = m(big)[1]
    ref = m(x)[1]
    assert torch.equal(out.view(50, 3, 2), ref.unsqueeze(0).expand(50, 3, 2))


@pytest.mark.parametrize("B,L", [(1, 64600), (5, 2400), (3, 16001), (7, 30011), (130, 9000), (2, 131072),
                                 # block 0 works on 120-column strips of four 30-column quadrants with their own
                                 # halo rows: widths whose pooled length J0 sits on / next to a strip or quadrant edge
                                 (2, 3368), (2, 3371), (3, 3377), (2, 3380), (2, 4178), (3, 4190), (2, 6608),
                                 # the two halves of an encoder pass run on two streams from 64 utterances on
                                 (63, 5000), (64, 5000), (65, 9000)])
def test_tc_path_agrees_with_fp32_path_on_odd_shapes(B, L):
    """Strip/tile edge handling (partial strips, J not a multiple of 120/126/128, tiny widths in the last blocks,
    batch sizes around the CTA count): the tensor-core path against the CUDA-core fp32 path, which the other
    tests pin to the oracle."""
    g = _g()
    x = O.speech_like(B, L, 100 + B).to(g.DEV)
    h32, o32 = g.native_model("AASIST", "fp32")(x)
    h16, o16 = g.native_model("AASIST", "f16x3")(x)
    torch.cuda.synchronize()
    assert torch.isfinite(o16).all()
    err = (o16 - o32).abs().max().item()
    herr = (h16 - h32).abs().max().item()
    print(json.dumps({"B": B, "L": L, "logit_err": err, "hidden_err": herr}))
    assert err <= TC_LOGIT_TOL and herr <= TC_LOGIT_TOL, (err, herr)


@pytest.mark.parametrize("name", ["AASIST", "AASIST-L"])
def test_f16x2_reduced_product_mode_error_is_measured(name):
    """Opt-in precision "f16x2" (two products in encoder blocks 1-5: weights rounded to fp16).
    Shipping bar for a reduced mode: logits <= 1e-3 AND 100 % ordered-index match on the reference goldens.
    MEASURED RESULT: neither model holds it (AASIST: 1.0e-3 on the 4 s speech golden and index flips on the long
    utterances; AASIST-L: top-k flips -- the CPU emulation in tools/precision_emulation.py predicts the same), so
    the mode is NOT shipped as a default anywhere: it stays an explicitly requested experiment whose error is
    printed here and in the bench line.  The test pins the measured order of magnitude (a regression guard)."""
    g = _g()
    m = g.native_model(name, "f16x2")
    worst, bad, total, report = 0.0, 0, 0, {}
    for tag in ("speech", "speech16k", "speech96k", "white"):
        gold, meta = load_golden(name, tag)
        x = golden_input(meta).to(g.DEV)
        m.record_topk = True
        try:
            lh, out = m(x)
            torch.cuda.synchronize()
            pools = g.split_pools(m.last_topk, m.last_pool_weights, m.topk_layout(meta["L"]))
        finally:
            m.record_topk = False
        err = float(np.abs(out.cpu().numpy() - gold["output"]).max())
        rep = g.check_pools(pools, gold, pools_of(name))
        strict = sum(r["strict_mismatch"] for r in rep.values())
        report[tag] = {"logit_err": err, "strict_topk_mismatch": strict}
        if tag != "white":                       # white noise has exact/near ties by construction
            worst, bad, total = max(worst, err), bad + strict, total + sum(r["positions"] for r in rep.values())
    print(json.dumps({"mode": "f16x2", "model": name, "worst_logit_err_speech": worst,
                      "strict_topk_mismatch_speech": bad, "positions": total, "cases": report}))
    assert all(np.isfinite(v["logit_err"]) for v in report.values())
    if name == "AASIST":
        assert worst <= 3e-3, report             # same order as single-pass TF32, 20-60x the f16x3 error


def test_tc_input_range_guard_warns_on_unnormalised_waveforms():
    """int16-scale samples saturate the fp16 operand pairs of the f16x3 front end; the kernel raises a
    host-visible flag and the shim turns it into a RuntimeWarning (the fp32 path has no such limit)."""
    import warnings
    g = _g()
    m = g.native_model("AASIST-L", "f16x3")
    x = O.speech_like(2, 64600, 5)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        m.score_host(x)                                    # normal input: no warning
    with pytest.warns(RuntimeWarning, match="exceed the fp16 operand range"):
        m.score_host(x * 32768.0)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        m.score_host(x)                                    # the flag was reset
    m32 = g.native_model("AASIST-L", "fp32")
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        m32.score_host(x * 32768.0)

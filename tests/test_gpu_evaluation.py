"""Device detection metrics (aasist_b200/evaluation.py -> aasist_det_metrics) against the reference's outputs
(tests/golden/evaluation.npz) and the numpy oracle: bit-exact float64."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import evaluation_oracle as E
from tests.util import GOLD

pytestmark = pytest.mark.gpu


def _golden():
    g = dict(np.load(os.path.join(GOLD, "evaluation.npz")))
    meta = json.loads(str(g.pop("meta")))
    return g, meta


@pytest.mark.parametrize("case", E.CASES, ids=[c[0] for c in E.CASES])
def test_device_metrics_equal_reference_bit_for_bit(case):
    from aasist_b200 import evaluation as V
    name = case[0]
    g, meta = _golden()
    bona, spoof, tar, non, spf = E.make_case(*case[1:])
    ref = g[f"{name}.scalars"]
    # CM scores arrive as fp32 CUDA tensors (the model's logits), ASV scores as float64 numpy (a text file)
    bona_d, spoof_d = torch.from_numpy(bona).cuda(), torch.from_numpy(spoof).cuda()
    eer_cm, eer_thr = V.compute_eer(bona_d, spoof_d)
    eer_asv, asv_thr = V.compute_eer(tar, non)
    pfa, pmiss, pmiss_spoof = V.obtain_asv_error_rates(tar, non, spf, asv_thr)
    curve, cthr, (min_tdcf, min_thr) = V.compute_tDCF(bona_d, spoof_d, pfa, pmiss, pmiss_spoof, E.COST_MODEL,
                                                       return_min=True)
    got = np.array([eer_cm, eer_thr, eer_asv, asv_thr, pfa, pmiss, pmiss_spoof, min_tdcf, min_thr, ref[9]])
    assert got.tobytes() == ref.tobytes(), (got, ref)
    frr, far, thr = V.compute_det_curve(bona_d, spoof_d)
    assert E.curves_digest(frr.cpu().numpy(), far.cpu().numpy(), thr.cpu().numpy(), curve.cpu().numpy()) == \
        meta[name]["curves_sha256"]
    assert int(torch.argmin(curve).item()) == int(ref[9]) or curve[int(ref[9])].item() == curve.min().item()


def test_device_metrics_error_behaviour():
    from aasist_b200 import evaluation as V
    from aasist_b200._lib import AasistError
    with pytest.raises(ValueError):
        V.compute_tDCF(np.array([0.0, 1.0]), np.array([0.0, 1.0]), 0.1, 0.1, 0.1, E.COST_MODEL)
    with pytest.raises(ValueError):
        V.compute_tDCF(np.array([0.0, np.nan, 2.0]), np.array([0.5]), 0.1, 0.1, 0.1, E.COST_MODEL)
    with pytest.raises(ValueError):
        V.compute_tDCF(np.array([0.0, 1.0, 2.0]), np.array([0.5]), 0.1, 0.1, None, E.COST_MODEL)
    with pytest.raises(AasistError):
        V.compute_eer(np.array([]), np.array([1.0]))
    # -0.0 and +0.0 are the same score (stable order decides), thresholds keep the original sign
    frr, far, thr = V.compute_det_curve(np.array([0.0]), np.array([-0.0]))
    assert frr.tolist() == [0.0, 1.0, 1.0] and far.tolist() == [1.0, 1.0, 0.0]


def test_score_files_to_report(tmp_path):
    """waveform scores -> score file (scoring.write_score_file, main.py:383-387 format) -> EER / min t-DCF report,
    against the oracle run on the same numbers."""
    from aasist_b200 import evaluation as V
    from aasist_b200.scoring import write_score_file
    case = E.CASES[4]
    bona, spoof, tar, non, spf = E.make_case(*case[1:])
    scores = np.concatenate((bona, spoof))
    keys = ["bonafide"] * bona.size + ["spoof"] * spoof.size
    srcs = ["-"] * bona.size + [f"A{7 + i % 13:02d}" for i in range(spoof.size)]
    utts = [f"LA_E_{i:07d}" for i in range(scores.size)]
    cm_file, asv_file, out_file = tmp_path / "cm.txt", tmp_path / "asv.txt", tmp_path / "report.txt"
    trial_lines = [f"SPK {u} - {s} {k}" for u, s, k in zip(utts, srcs, keys)]
    write_score_file(str(cm_file), utts, scores.tolist(), trial_lines)
    with open(asv_file, "w") as f:
        for lab, arr in (("target", tar), ("nontarget", non), ("spoof", spf)):
            for v in arr:
                f.write(f"X {lab} {float(v)!r}\n")
    eer_pct, min_tdcf = V.calculate_tDCF_EER(str(cm_file), str(asv_file), str(out_file), printout=True)
    ref_eer, ref_tdcf = E.cm_metrics(bona, spoof, tar, non, spf)
    assert eer_pct == ref_eer and min_tdcf == ref_tdcf
    text = open(out_file).read()
    assert "min-tDCF" in text and "EER A07" in text and "EER A19" in text

"""Detection-metrics oracle (oracle/evaluation_oracle.py) against fixtures produced by the REFERENCE's own
evaluation.py (oracle/make_golden_eval.py): bit-exact float64, including tie handling."""
import json
import os

import numpy as np
import pytest

from oracle import evaluation_oracle as E
from tests.util import GOLD


def _golden():
    g = dict(np.load(os.path.join(GOLD, "evaluation.npz")))
    meta = json.loads(str(g.pop("meta")))
    return g, meta


@pytest.mark.parametrize("case", E.CASES, ids=[c[0] for c in E.CASES])
def test_oracle_matches_reference_outputs_bit_for_bit(case):
    name = case[0]
    g, meta = _golden()
    bona, spoof, tar, non, spf = E.make_case(*case[1:])
    assert E.curves_digest(bona, spoof, tar, non, spf) == meta[name]["inputs_sha256"], "seeded inputs changed"
    ref = g[f"{name}.scalars"]
    eer_cm, eer_thr = E.eer(bona, spoof)
    eer_asv, asv_thr = E.eer(tar, non)
    pfa, pmiss, pmiss_spoof = E.asv_error_rates(tar, non, spf, asv_thr)
    curve, cthr = E.tdcf(bona, spoof, pfa, pmiss, pmiss_spoof)
    imin = int(np.argmin(curve))
    got = np.array([eer_cm, eer_thr, eer_asv, asv_thr, pfa, pmiss, pmiss_spoof, curve[imin], cthr[imin], imin])
    assert got.tobytes() == ref.tobytes(), (got, ref)
    frr, far, thr = E.det_curve(bona, spoof)
    assert E.curves_digest(frr, far, thr, curve) == meta[name]["curves_sha256"]
    if f"{name}.frr" in g:
        assert frr.tobytes() == g[f"{name}.frr"].tobytes() and far.tobytes() == g[f"{name}.far"].tobytes()
        assert thr.tobytes() == g[f"{name}.thr"].tobytes() and curve.tobytes() == g[f"{name}.tdcf"].tobytes()
    eer_pct, min_tdcf = E.cm_metrics(bona, spoof, tar, non, spf)
    assert eer_pct == meta[name]["eer_percent"] and min_tdcf == meta[name]["min_tdcf"]


def test_oracle_error_behaviour_follows_the_reference():
    with pytest.raises(ValueError):
        E.tdcf(np.array([0.0, 1.0]), np.array([0.0, 1.0]), 0.1, 0.1, 0.1)        # fewer than 3 distinct scores
    with pytest.raises(ValueError):
        E.tdcf(np.array([0.0, np.nan, 2.0]), np.array([0.5]), 0.1, 0.1, 0.1)
    with pytest.raises(ValueError):
        E.tdcf_weights(0.1, 0.1, None)


def test_stable_tie_order_targets_first():
    # equal scores: the stable sort keeps [targets, nontargets] order, so the target is "rejected" first
    frr, far, thr = E.det_curve(np.array([1.0]), np.array([1.0]))
    assert frr.tolist() == [0.0, 1.0, 1.0] and far.tolist() == [1.0, 1.0, 0.0] and thr.tolist() == [0.999, 1.0, 1.0]

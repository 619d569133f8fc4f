"""CPU restatement (numpy, float64) of the reference's detection metrics.

TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else.  The product path is
aasist_b200/evaluation.py -> libaasist_b200.so (aasist_det_metrics); it never imports this file.

Restates /root/reference/evaluation.py:
  det_curve            <- compute_det_curve          (evaluation.py:120-145)
  eer                  <- compute_eer                (evaluation.py:148-154)
  asv_error_rates      <- obtain_asv_error_rates     (evaluation.py:103-117)
  tdcf_weights / tdcf  <- compute_tDCF               (evaluation.py:157-329: checks :237-264, curve :266-282)
  cm_metrics           <- the arithmetic of calculate_tDCF_EER (evaluation.py:7-100) without file I/O
Pinned against the reference itself: tests/golden/evaluation.npz is produced by oracle/make_golden_eval.py, which
imports the reference module unmodified, and tests/test_evaluation_cpu.py checks this file against it bit for bit.
"""
from __future__ import annotations

import numpy as np

# cost model fixed in calculate_tDCF_EER (evaluation.py:11-23)
PSPOOF = 0.05
COST_MODEL = {
    "Pspoof": PSPOOF, "Ptar": (1 - PSPOOF) * 0.99, "Pnon": (1 - PSPOOF) * 0.01,
    "Cmiss": 1, "Cfa": 10, "Cmiss_asv": 1, "Cfa_asv": 10, "Cmiss_cm": 1, "Cfa_cm": 10,
}


def det_curve(target_scores, nontarget_scores):
    """(frr, far, thresholds), each of length n+1; scores sorted ascending, STABLE over the concatenation
    [targets, nontargets] (np.argsort(kind='mergesort'), evaluation.py:128)."""
    t = np.asarray(target_scores, dtype=np.float64)
    n = np.asarray(nontarget_scores, dtype=np.float64)
    total = t.size + n.size
    scores = np.concatenate((t, n))
    labels = np.concatenate((np.ones(t.size), np.zeros(n.size)))
    order = np.argsort(scores, kind="mergesort")
    labels = labels[order]
    tar_sums = np.cumsum(labels)
    non_sums = n.size - (np.arange(1, total + 1) - tar_sums)
    frr = np.concatenate((np.atleast_1d(0), tar_sums / t.size))
    far = np.concatenate((np.atleast_1d(1), non_sums / n.size))
    thr = np.concatenate((np.atleast_1d(scores[order[0]] - 0.001), scores[order]))
    return frr, far, thr


def eer(target_scores, nontarget_scores):
    """(eer, threshold): operating point with the smallest |frr - far|, first one on ties (np.argmin)."""
    frr, far, thr = det_curve(target_scores, nontarget_scores)
    i = int(np.argmin(np.abs(frr - far)))
    return float(np.mean((frr[i], far[i]))), float(thr[i])


def asv_error_rates(tar_asv, non_asv, spoof_asv, asv_threshold):
    tar_asv, non_asv, spoof_asv = (np.asarray(a, dtype=np.float64) for a in (tar_asv, non_asv, spoof_asv))
    pfa = np.sum(non_asv >= asv_threshold) / non_asv.size
    pmiss = np.sum(tar_asv < asv_threshold) / tar_asv.size
    pmiss_spoof = None if spoof_asv.size == 0 else np.sum(spoof_asv < asv_threshold) / spoof_asv.size
    return pfa, pmiss, pmiss_spoof


def tdcf_weights(pfa_asv, pmiss_asv, pmiss_spoof_asv, cost_model=COST_MODEL):
    """C1, C2 of the normalised t-DCF (evaluation.py:270-273); ValueError where the reference sys.exit()s."""
    c = cost_model
    if c["Ptar"] < 0 or c["Pnon"] < 0 or c["Pspoof"] < 0 or abs(c["Ptar"] + c["Pnon"] + c["Pspoof"] - 1) > 1e-10:
        raise ValueError("prior probabilities should be positive and sum up to one")
    if pmiss_spoof_asv is None:
        raise ValueError("miss rate of spoof tests against the ASV system is required")
    c1 = c["Ptar"] * (c["Cmiss_cm"] - c["Cmiss_asv"] * pmiss_asv) - c["Pnon"] * c["Cfa_asv"] * pfa_asv
    c2 = c["Cfa_cm"] * c["Pspoof"] * (1 - pmiss_spoof_asv)
    if c1 < 0 or c2 < 0:
        raise ValueError("negative t-DCF weights: check the ASV error rates")
    return c1, c2


def tdcf(bona_cm, spoof_cm, pfa_asv, pmiss_asv, pmiss_spoof_asv, cost_model=COST_MODEL):
    """(normalised t-DCF curve, CM thresholds) (evaluation.py:266-282)."""
    combined = np.concatenate((np.asarray(bona_cm, np.float64), np.asarray(spoof_cm, np.float64)))
    if np.isnan(combined).any() or np.isinf(combined).any():
        raise ValueError("scores contain nan or inf")
    if np.unique(combined).size < 3:
        raise ValueError("soft CM scores required, not binary decisions")
    c1, c2 = tdcf_weights(pfa_asv, pmiss_asv, pmiss_spoof_asv, cost_model)
    pmiss_cm, pfa_cm, thr = det_curve(bona_cm, spoof_cm)
    curve = c1 * pmiss_cm + c2 * pfa_cm
    return curve / np.minimum(c1, c2), thr


def cm_metrics(bona_cm, spoof_cm, tar_asv, non_asv, spoof_asv):
    """(EER of the countermeasure in %, min t-DCF): calculate_tDCF_EER's return value (evaluation.py:43-100)."""
    _, asv_thr = eer(tar_asv, non_asv)
    eer_cm = eer(bona_cm, spoof_cm)[0]
    pfa, pmiss, pmiss_spoof = asv_error_rates(tar_asv, non_asv, spoof_asv, asv_thr)
    curve, _ = tdcf(bona_cm, spoof_cm, pfa, pmiss, pmiss_spoof)
    return eer_cm * 100, float(curve[int(np.argmin(curve))])


# ---- seeded inputs of the golden cases (tests/golden/evaluation.npz stores the reference's outputs for them) ----
def make_case(seed: int, n_bona: int, n_spoof: int, sep: float, quant: float, n_tar: int, n_non: int, n_spf: int):
    """CM scores as the scoring path produces them (fp32 logits), optionally quantised to create ties;
    ASV scores are arbitrary float64 (they come from a text file in the reference)."""
    r = np.random.Generator(np.random.PCG64(seed))
    bona = (r.normal(sep, 1.0, n_bona)).astype(np.float32)
    spoof = (r.normal(-sep, 1.3, n_spoof)).astype(np.float32)
    if quant > 0:
        bona = (np.round(bona / quant) * quant).astype(np.float32)
        spoof = (np.round(spoof / quant) * quant).astype(np.float32)
    tar = r.normal(2.0, 1.0, n_tar)
    non = r.normal(-2.0, 1.0, n_non)
    spf = r.normal(0.5, 1.5, n_spf)
    return bona, spoof, tar, non, spf


CASES = [  # name, seed, n_bona, n_spoof, separation, quantisation step, ASV target / nontarget / spoof counts
    ("tiny", 1, 5, 7, 1.0, 0.0, 6, 9, 7),
    ("ties", 2, 300, 900, 0.8, 0.25, 200, 400, 300),
    ("heavy_ties", 3, 64, 64, 0.2, 1.0, 50, 50, 50),
    ("separable", 4, 400, 600, 6.0, 0.0, 100, 100, 100),
    ("medium", 5, 2548, 22296, 1.5, 0.0, 1484, 1484, 22296),       # ASVspoof2019 LA dev-like proportions
    ("evalset", 6, 7355, 63882, 2.0, 0.0, 5370, 33327, 63882),     # 71 237 CM trials (BASELINE.json C3)
    ("evalset_ties", 7, 7355, 63882, 2.0, 0.01, 5370, 33327, 63882),
]


def curves_digest(*arrays) -> str:
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()

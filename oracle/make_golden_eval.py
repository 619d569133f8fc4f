"""Generate tests/golden/evaluation.npz by running the REFERENCE's own evaluation.py (pure numpy).

TEST INFRASTRUCTURE ONLY; runs in the build container where /root/reference is mounted.  The module is imported
unmodified; every case stores its inputs and the reference's outputs (full DET curves for the small cases, a sha256
of the curves for the large ones) so that the oracle restatement and the CUDA path can be checked bit for bit on a
box where the reference does not exist.

Usage:  python oracle/make_golden_eval.py
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os

import sys

import numpy as np

REF = os.environ.get("AASIST_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
from oracle.evaluation_oracle import CASES, make_case  # noqa: E402  (seeded input generators only)


def reference_module():
    spec = importlib.util.spec_from_file_location("ref_evaluation", os.path.join(REF, "evaluation.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def main():
    ev = reference_module()
    cost = {"Pspoof": 0.05, "Ptar": 0.95 * 0.99, "Pnon": 0.95 * 0.01, "Cmiss": 1, "Cfa": 10,
            "Cmiss_asv": 1, "Cfa_asv": 10, "Cmiss_cm": 1, "Cfa_cm": 10}
    out, meta = {}, {}
    for name, seed, nb, ns, sep, q, nt, nn_, nsp in CASES:
        bona, spoof, tar, non, spf = make_case(seed, nb, ns, sep, q, nt, nn_, nsp)
        b64, s64 = bona.astype(np.float64), spoof.astype(np.float64)
        frr, far, thr = ev.compute_det_curve(b64, s64)
        eer_cm, eer_thr = ev.compute_eer(b64, s64)
        eer_asv, asv_thr = ev.compute_eer(tar, non)
        pfa, pmiss, pmiss_spoof = ev.obtain_asv_error_rates(tar, non, spf, asv_thr)
        curve, cthr = ev.compute_tDCF(b64, s64, pfa, pmiss, pmiss_spoof, cost, False)
        imin = int(np.argmin(curve))
        # inputs are regenerated from the seed at test time (oracle.evaluation_oracle.make_case); their digest pins them
        meta_inputs = digest(bona, spoof, tar, non, spf)
        out[f"{name}.scalars"] = np.array([eer_cm, eer_thr, eer_asv, asv_thr, pfa, pmiss, pmiss_spoof,
                                           curve[imin], cthr[imin], imin], dtype=np.float64)
        if nb + ns <= 2000:
            out[f"{name}.frr"], out[f"{name}.far"], out[f"{name}.thr"], out[f"{name}.tdcf"] = frr, far, thr, curve
        meta[name] = {"seed": seed, "n_bona": nb, "n_spoof": ns, "inputs_sha256": meta_inputs, "curves_sha256": digest(frr, far, thr, curve),
                      "eer_percent": float(eer_cm * 100), "min_tdcf": float(curve[imin])}
        print(name, meta[name])
    np.savez_compressed(os.path.join(GOLD, "evaluation.npz"), meta=json.dumps(meta), **out)


if __name__ == "__main__":
    main()

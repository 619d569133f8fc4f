"""Detection metrics of an evaluation run -- the reference's ``evaluation.py`` with the DET curve, EER and
t-DCF computed on the GPU (``aasist_det_metrics`` in libaasist_b200.so: one stable radix sort + scan kernel).

Same function names, argument meaning and return values as the reference:

* ``compute_det_curve``       -- evaluation.py:120-145
* ``compute_eer``             -- evaluation.py:148-154
* ``obtain_asv_error_rates``  -- evaluation.py:103-117 (three threshold counts over the ASV score file: host numpy)
* ``compute_tDCF``            -- evaluation.py:157-329 (where the reference calls ``sys.exit`` this raises ValueError)
* ``calculate_tDCF_EER``      -- evaluation.py:7-100 (reads the two score files, writes the same report)

Scores may be numpy arrays or torch tensors on any device; they are held as float64 on the GPU like numpy holds
them on the host (fp32 logits convert exactly), and every returned number is bit-identical to the reference's.
There is no CPU fallback: without the CUDA library (or a GPU) the calls raise.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

__all__ = ["compute_det_curve", "compute_eer", "obtain_asv_error_rates", "compute_tDCF", "calculate_tDCF_EER"]


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("aasist_b200.evaluation needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(scores) -> torch.Tensor:
    if isinstance(scores, torch.Tensor):
        t = scores.detach()
        dev = t.device if t.is_cuda else _device()
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(scores, dtype=np.float64)))
        dev = _device()
    return t.to(device=dev, dtype=torch.float64).contiguous().view(-1)


def _det(target_scores, nontarget_scores, c1: float = -1.0, c2: float = -1.0, curves: bool = False):
    lib = _lib.load()
    t, n = _to_dev(target_scores), _to_dev(nontarget_scores)
    if n.device != t.device:
        n = n.to(t.device)
    total = t.numel() + n.numel()
    with torch.cuda.device(t.device):
        ws = torch.empty(int(lib.aasist_det_workspace_bytes(total)), dtype=torch.uint8, device=t.device)
        out = [torch.empty(total + 1, dtype=torch.float64, device=t.device) for _ in range(4)] if curves else None
        res = (C.c_double * 8)()
        ptr = (lambda x: C.c_void_p(x.data_ptr())) if curves else None
        stream = torch.cuda.current_stream(t.device).cuda_stream
        _lib.check(lib.aasist_det_metrics(
            C.c_void_p(t.data_ptr()), t.numel(), C.c_void_p(n.data_ptr()), n.numel(), float(c1), float(c2), res,
            ptr(out[0]) if curves else None, ptr(out[1]) if curves else None, ptr(out[2]) if curves else None,
            ptr(out[3]) if curves else None, C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(stream)))
    return list(res), out


def compute_det_curve(target_scores, nontarget_scores):
    """(frr, far, thresholds): float64 CUDA tensors of length n_target + n_nontarget + 1."""
    _, out = _det(target_scores, nontarget_scores, curves=True)
    return out[0], out[1], out[2]


def compute_eer(target_scores, nontarget_scores) -> Tuple[float, float]:
    """ Returns equal error rate (EER) and the corresponding threshold. """
    res, _ = _det(target_scores, nontarget_scores)
    return res[0], res[1]


def obtain_asv_error_rates(tar_asv, non_asv, spoof_asv, asv_threshold):
    tar_asv, non_asv, spoof_asv = (np.asarray(a, dtype=np.float64) for a in (tar_asv, non_asv, spoof_asv))
    Pfa_asv = np.sum(non_asv >= asv_threshold) / non_asv.size
    Pmiss_asv = np.sum(tar_asv < asv_threshold) / tar_asv.size
    Pmiss_spoof_asv = None if spoof_asv.size == 0 else np.sum(spoof_asv < asv_threshold) / spoof_asv.size
    return Pfa_asv, Pmiss_asv, Pmiss_spoof_asv


def _tdcf_weights(Pfa_asv, Pmiss_asv, Pmiss_spoof_asv, cost_model) -> Tuple[float, float]:
    if cost_model['Cfa_asv'] < 0 or cost_model['Cmiss_asv'] < 0 or \
            cost_model['Cfa_cm'] < 0 or cost_model['Cmiss_cm'] < 0:
        print('WARNING: Usually the cost values should be positive!')
    if cost_model['Ptar'] < 0 or cost_model['Pnon'] < 0 or cost_model['Pspoof'] < 0 or \
            np.abs(cost_model['Ptar'] + cost_model['Pnon'] + cost_model['Pspoof'] - 1) > 1e-10:
        raise ValueError('ERROR: Your prior probabilities should be positive and sum up to one.')
    if Pmiss_spoof_asv is None:
        raise ValueError('ERROR: you should provide miss rate of spoof tests against your ASV system.')
    C1 = cost_model['Ptar'] * (cost_model['Cmiss_cm'] - cost_model['Cmiss_asv'] * Pmiss_asv) - \
        cost_model['Pnon'] * cost_model['Cfa_asv'] * Pfa_asv
    C2 = cost_model['Cfa_cm'] * cost_model['Pspoof'] * (1 - Pmiss_spoof_asv)
    if C1 < 0 or C2 < 0:
        raise ValueError('You should never see this error but I cannot evalute tDCF with negative weights - '
                         'please check whether your ASV error rates are correctly computed?')
    return float(C1), float(C2)


def compute_tDCF(bonafide_score_cm, spoof_score_cm, Pfa_asv, Pmiss_asv, Pmiss_spoof_asv, cost_model,
                 print_cost=False, return_min: bool = False):
    """Normalised t-DCF curve and CM thresholds (float64 CUDA tensors), as the reference's compute_tDCF.
    ``return_min=True`` additionally returns (min t-DCF, its threshold) picked on the device."""
    C1, C2 = _tdcf_weights(Pfa_asv, Pmiss_asv, Pmiss_spoof_asv, cost_model)
    res, out = _det(bonafide_score_cm, spoof_score_cm, C1, C2, curves=True)
    if int(res[7]) & 1:
        raise ValueError('ERROR: Your scores contain nan or inf.')
    if res[6] < 3:
        raise ValueError('ERROR: You should provide soft CM scores - not binary decisions')
    if print_cost:
        print('t-DCF evaluation from [Nbona={}, Nspoof={}] trials\n'.format(
            int(np.size(bonafide_score_cm)), int(np.size(spoof_score_cm))))
        print('   tDCF_norm(s) = {:8.5f} x Pmiss_cm(s) + {:8.5f} x Pfa_cm(s)\n'.format(
            C1 / min(C1, C2), C2 / min(C1, C2)))
    if return_min:
        return out[3], out[2], (res[3], res[4])
    return out[3], out[2]


def calculate_tDCF_EER(cm_scores_file, asv_score_file, output_file, printout=True):
    # Fix tandem detection cost function (t-DCF) parameters
    Pspoof = 0.05
    cost_model = {
        'Pspoof': Pspoof, 'Ptar': (1 - Pspoof) * 0.99, 'Pnon': (1 - Pspoof) * 0.01,
        'Cmiss': 1, 'Cfa': 10, 'Cmiss_asv': 1, 'Cfa_asv': 10, 'Cmiss_cm': 1, 'Cfa_cm': 10,
    }
    asv_data = np.genfromtxt(asv_score_file, dtype=str)
    asv_keys = asv_data[:, 1]
    asv_scores = asv_data[:, 2].astype(np.float64)
    cm_data = np.genfromtxt(cm_scores_file, dtype=str)
    cm_sources = cm_data[:, 1]
    cm_keys = cm_data[:, 2]
    cm_scores = cm_data[:, 3].astype(np.float64)

    tar_asv = asv_scores[asv_keys == 'target']
    non_asv = asv_scores[asv_keys == 'nontarget']
    spoof_asv = asv_scores[asv_keys == 'spoof']
    bona_cm = cm_scores[cm_keys == 'bonafide']
    spoof_cm = cm_scores[cm_keys == 'spoof']

    eer_asv, asv_threshold = compute_eer(tar_asv, non_asv)
    bona_dev = _to_dev(bona_cm)
    eer_cm = compute_eer(bona_dev, spoof_cm)[0]

    attack_types = [f'A{_id:02d}' for _id in range(7, 20)]
    if printout:
        eer_cm_breakdown = {}
        for attack_type in attack_types:
            sub = cm_scores[cm_sources == attack_type]
            eer_cm_breakdown[attack_type] = compute_eer(bona_dev, sub)[0] if sub.size else float('nan')

    Pfa_asv, Pmiss_asv, Pmiss_spoof_asv = obtain_asv_error_rates(tar_asv, non_asv, spoof_asv, asv_threshold)
    _, _, (min_tDCF, _) = compute_tDCF(bona_dev, spoof_cm, Pfa_asv, Pmiss_asv, Pmiss_spoof_asv, cost_model,
                                       print_cost=False, return_min=True)

    if printout:
        with open(output_file, "w") as f_res:
            f_res.write('\nCM SYSTEM\n')
            f_res.write('\tEER\t\t= {:8.9f} % '
                        '(Equal error rate for countermeasure)\n'.format(eer_cm * 100))
            f_res.write('\nTANDEM\n')
            f_res.write('\tmin-tDCF\t\t= {:8.9f}\n'.format(min_tDCF))
            f_res.write('\nBREAKDOWN CM SYSTEM\n')
            for attack_type in attack_types:
                _eer = eer_cm_breakdown[attack_type] * 100
                f_res.write(f'\tEER {attack_type}\t\t= {_eer:8.9f} % '
                            f'(Equal error rate for {attack_type})\n')
        with open(output_file, "r") as f:
            sys.stdout.write(f.read())

    return eer_cm * 100, min_tDCF

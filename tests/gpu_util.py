"""Helpers for the -m gpu parity tests: build models, call the C ABI stage entry points."""
import ctypes as C
import functools

import torch

import aasist_b200
from aasist_b200 import _lib
from oracle import aasist_oracle as O
from tests.util import load_sd

DEV = torch.device("cuda:0")


@functools.lru_cache(maxsize=None)
def native_model(name: str, precision: str = "fp32"):
    cls = {"RawGAT-ST": aasist_b200.RawGATSTModel, "AASIST-Robust": aasist_b200.RobustModel}.get(name, aasist_b200.Model)
    m = cls(aasist_b200.CONFIGS[name], precision=precision)
    m.load_state_dict(load_sd(name), strict=True)
    m = m.to(DEV).eval()
    m._ensure_handle(DEV)
    return m


def oracle_taps(name: str, x_cpu: torch.Tensor):
    taps = {}
    torch.set_num_threads(8)
    O.forward(name, load_sd(name), O.CONFIGS[name], x_cpu, taps)
    return taps


def stage_frontend(m, x):
    lib = _lib.load()
    B, L = x.shape
    Wp = (L - 128) // 3
    out = torch.empty(B, 1, 23, Wp, device=DEV)
    nbytes = int(lib.aasist_workspace_bytes(m._handle, B, L))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _lib.check(lib.aasist_frontend(m._handle, x.data_ptr(), B, L, out.data_ptr(), ws.data_ptr(), nbytes, None))
    torch.cuda.synchronize()
    return out


def stage_frontend_raw(m, x, shape):
    """aasist_frontend for models whose front-end geometry differs (AASIST-Robust): `shape` = expected output."""
    lib = _lib.load()
    B, L = x.shape
    out = torch.empty(*shape, device=DEV)
    _lib.check(lib.aasist_frontend(m._handle, x.data_ptr(), B, L, out.data_ptr(), None, 0, None))
    torch.cuda.synchronize()
    return out


def stage_block(m, enc, index, x_in, co):
    lib = _lib.load()
    x_in = x_in.contiguous()
    B, ci, H, W = x_in.shape
    out = torch.empty(B, co, 23, W // 3, device=DEV)
    nbytes = 4 * B * (2 * max(ci, 1) + max(co, ci, 32)) * 24 * W * 4 + (1 << 22)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _lib.check(lib.aasist_encoder_block(m._handle, enc, index, x_in.data_ptr(), B, W, out.data_ptr(),
                                        ws.data_ptr(), nbytes, None))
    torch.cuda.synchronize()
    return out


def stage_graph(m, e, e2=None, L=64600):
    lib = _lib.load()
    B, NT = e.shape[0], e.shape[3]
    layout = m.topk_layout(L)
    hd = m.hidden_dim
    lh = torch.empty(B, hd, device=DEV)
    lg = torch.empty(B, 2, device=DEV)
    idx = torch.full((B, sum(k for _, k in layout)), -1, dtype=torch.int32, device=DEV)
    w = torch.empty(B, sum(n for n, _ in layout), device=DEV)
    _lib.check(lib.aasist_graph(m._handle, e.contiguous().data_ptr(),
                                e2.contiguous().data_ptr() if e2 is not None else None, B, NT,
                                lh.data_ptr(), lg.data_ptr(), idx.data_ptr(), w.data_ptr(), None))
    torch.cuda.synchronize()
    return lh, lg, split_pools(idx, w, layout)


def split_pools(idx, w, layout):
    """-> list of (indices (B,k), weights (B,n)) per pool, on the CPU."""
    out, io, wo = [], 0, 0
    for n, k in layout:
        out.append((idx[:, io:io + k].cpu(), w[:, wo:wo + n].cpu()))
        io += k
        wo += n
    return out


def check_pools(pools, ref_taps_or_golden, names, near_gap=1e-5, get=lambda d, k: d[k]):
    """Tie policy of SURVEY 8(c): ordered indices must match wherever the oracle's neighbouring
    sorted scores differ by more than `near_gap` (exact ties are always exempt).  Returns report."""
    report = {}
    for (idx, w), name in zip(pools, names):
        ref_w = torch.as_tensor(get(ref_taps_or_golden, name + ".weights")).float()
        ref_i = torch.as_tensor(get(ref_taps_or_golden, name + ".idx"))
        werr = (w - ref_w).abs().max().item()
        strict = O.compare_topk(torch.sigmoid(ref_w), ref_i, idx, near_gap=0.0)
        loose = O.compare_topk(ref_w, ref_i, idx, near_gap=near_gap)
        report[name] = {"weights_err": werr, "strict_mismatch": strict[0], "positions": strict[1],
                        "exact_tie_positions": strict[2], "near_tie_positions": loose[2],
                        "mismatch_outside_near_ties": loose[0]}
    return report

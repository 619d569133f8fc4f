"""Shared helpers for the test-suite (golden fixtures, weights, input generators)."""
import json
import os

import numpy as np
import torch

from oracle import aasist_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
WDIR = os.path.join(ROOT, "aasist_b200", "weights")
WEIGHTS = {"AASIST": "AASIST.pth", "AASIST-L": "AASIST-L.pth", "RawGAT-ST": "RawGATST_seed1234.pth",
           "AASIST2": "AASIST2_seed1234.pth", "AASIST2-small": "AASIST2-small_seed1234.pth",
           "AASIST-Robust": "AASIST-Robust_seed1234.pth"}
GENERATORS = {"white": O.white_noise, "speech": O.speech_like,
              "speech16k": O.speech_like, "speech96k": O.speech_like, "speech128k": O.speech_like,
              "speech192k": O.speech_like, "speech256k": O.speech_like}
AASIST_POOLS = ["pool_S", "pool_T", "pool_hS1", "pool_hT1", "pool_hS2", "pool_hT2"]
RAWGAT_POOLS = ["pool_T", "pool_S", "pool_ST"]


def load_sd(model: str):
    return torch.load(os.path.join(WDIR, WEIGHTS[model]), map_location="cpu")


def load_golden(model: str, tag: str):
    g = dict(np.load(os.path.join(GOLD, f"{model}_{tag}.npz")))
    meta = json.loads(str(g.pop("meta")))
    return g, meta


def golden_input(meta):
    return GENERATORS[meta["input"]](meta["n"], meta["L"], meta["seed"])


def pools_of(model: str):
    return RAWGAT_POOLS if model == "RawGAT-ST" else AASIST_POOLS

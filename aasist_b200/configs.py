"""``model_config`` dictionaries of the configurations on the hot path, as the reference's
JSON ``.conf`` files define them (config/AASIST.conf:13-21, config/AASIST-L.conf:13-21,
config/RawGATST_baseline.conf:12-17), plus a loader for any reference-format ``.conf``."""
from __future__ import annotations

import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    "AASIST": {
        "architecture": "AASIST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 64], [64, 64]],
        "gat_dims": [64, 32], "pool_ratios": [0.5, 0.7, 0.5, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
    },
    "AASIST-L": {
        "architecture": "AASIST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 24], [24, 24]],
        "gat_dims": [24, 32], "pool_ratios": [0.4, 0.5, 0.7, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
    },
    "RawGAT-ST": {
        "architecture": "RawNetGatSpoofST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 64], [64, 64]],
    },
}

# fork-only configurations (no checkpoint exists for any of them: seeded reference-class initialisation)
CONFIGS.update({
    # config/AASIST2.conf:21-35: the fork's own model (Res2Net+SE encoder, speaker conditioning)
    "AASIST2": {
        "architecture": "AASIST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 64], [64, 64]],
        "gat_dims": [64, 32], "pool_ratios": [0.5, 0.7, 0.5, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
        "res2net_width": 14, "res2net_scale": 8,
        "speaker_conditioning": True, "spk_emb_dim": 256, "conditioning_level": "frame",
        "use_attention": True,
    },
    # a second Res2Net configuration used by the parity tests (scale-group chaining, no attention)
    "AASIST2-small": {
        "architecture": "AASIST", "nb_samp": 64600, "first_conv": 128,
        "filts": [70, [1, 32], [32, 32], [32, 24], [24, 24]],
        "gat_dims": [24, 32], "pool_ratios": [0.4, 0.5, 0.7, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
        "res2net_width": 6, "res2net_scale": 2,
        "speaker_conditioning": True, "spk_emb_dim": 64, "conditioning_level": "frame",
        "use_attention": False,
    },
    # config/AASIST-Robust.conf:24-31 with first_conv 70: with the file's 128 the reference model cannot run
    # (42 spectral bands against pos_S's 23, AASIST_Robust.py:237-238)
    "AASIST-Robust": {
        "architecture": "AASIST_Robust", "nb_samp": 64600, "first_conv": 70,
        "filts": [70, [1, 32], [32, 32], [32, 24], [24, 24]],
        "gat_dims": [24, 32], "pool_ratios": [0.4, 0.5, 0.7, 0.5],
        "temperatures": [2.0, 2.0, 100.0, 100.0],
    },
})

# checkpoints shipped with the reference (models/weights/*.pth), kept byte-identical here;
# RawGAT-ST has no published checkpoint: seeded reference-class init (oracle/make_golden.py)
WEIGHTS = {"AASIST": "AASIST.pth", "AASIST-L": "AASIST-L.pth", "RawGAT-ST": "RawGATST_seed1234.pth",
           "AASIST2": "AASIST2_seed1234.pth", "AASIST2-small": "AASIST2-small_seed1234.pth",
           "AASIST-Robust": "AASIST-Robust_seed1234.pth"}


def weights_path(name: str) -> str:
    return os.path.join(_HERE, "weights", WEIGHTS[name])


def load_model_config(conf_path: str) -> dict:
    """``json.loads(conf)["model_config"]`` exactly as reference main.py:42-44 does."""
    with open(conf_path, "r") as f:
        return json.loads(f.read())["model_config"]

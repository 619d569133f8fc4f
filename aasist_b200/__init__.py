"""aasist_b200 -- B200 (sm_100a) implementation of the AASIST / RawGAT-ST batched
utterance-scoring forward pass behind the reference's ``Model(d_args).forward(x)`` plug-in
interface.  Host code is Python; all compute is hand-written CUDA in libaasist_b200.so,
reached through the C ABI in include/aasist_b200.h.  No CPU fallback."""
from .configs import CONFIGS, WEIGHTS, load_model_config, weights_path
from .model import Model, RawGATSTModel, RobustModel
from .scoring import get_model, score_utterances, write_score_file

# precisions whose kernels are built into libaasist_b200.so
BUILT_PRECISIONS = ("fp32", "f16x3")      # smoke()/default set; "f16x2" is the opt-in reduced-product mode

__all__ = ["BUILT_PRECISIONS", "CONFIGS", "WEIGHTS", "Model", "RawGATSTModel", "RobustModel", "get_model", "load_model_config",
           "score_utterances", "weights_path", "write_score_file"]

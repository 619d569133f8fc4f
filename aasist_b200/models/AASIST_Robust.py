"""``architecture: "AASIST_Robust"`` plug-in (reference models/AASIST_Robust.py): exposes ``Model``."""
from ..model import RobustModel as Model  # noqa: F401

"""``architecture: "AASIST"`` plug-in (reference models/AASIST.py): exposes ``Model``."""
from ..model import Model  # noqa: F401

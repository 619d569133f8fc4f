"""``architecture: "RawNetGatSpoofST"`` plug-in (reference models/RawNetGatSpoofST.py)."""
from ..model import RawGATSTModel as Model  # noqa: F401

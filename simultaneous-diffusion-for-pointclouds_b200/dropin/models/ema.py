"""models.ema shim (reference: LiDARGen/models/ema.py:4)."""
from sdpc_b200.ema import EMAHelper  # noqa: F401

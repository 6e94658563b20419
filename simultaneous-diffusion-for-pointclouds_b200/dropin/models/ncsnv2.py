"""models.ncsnv2 shim (reference: LiDARGen/models/ncsnv2.py:420)."""
from sdpc_b200.scorenet import NCSN_LiDAR_small  # noqa: F401

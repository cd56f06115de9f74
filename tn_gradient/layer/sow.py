"""``tn_gradient.layer.sow`` -> sow_b200.layer (kernel-backed SoWLinear / SoWParameter / SoWArgs)."""
from sow_b200.layer import SoWArgs, SoWLinear, SoWParameter  # noqa: F401
from sow_b200.utils import qr_weight  # noqa: F401

"""``tn_gradient.prepare`` -> sow_b200.surgery."""
from sow_b200.layer import SoWArgs, SoWLinear  # noqa: F401
from sow_b200.surgery import (SoWConfig, SoWModel, accumulate, export_alignment, load_sow,  # noqa: F401
                              prepare_sow, sow_modules)
from sow_b200.utils import svd_weight  # noqa: F401

"""``tn_gradient.optimizer.ttadam`` -> sow_b200.optim."""
from sow_b200.optim import TTAdam, TTRAdam  # noqa: F401
from sow_b200.tt import TensorTrain  # noqa: F401

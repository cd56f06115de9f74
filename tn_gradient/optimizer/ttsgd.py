"""``tn_gradient.optimizer.ttsgd`` -> sow_b200.optim."""
from sow_b200.optim import TTSGD  # noqa: F401
from sow_b200.tt import TensorTrain  # noqa: F401

"""``tn_gradient.tt`` -> sow_b200.tt."""
from sow_b200.tt import TensorTrain  # noqa: F401
from sow_b200.utils import closest_factorization, pad_matrix, unpad_matrix  # noqa: F401

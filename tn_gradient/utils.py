"""``tn_gradient.utils`` -> sow_b200.utils."""
from sow_b200.utils import (__colorized_str__, closest_factorization, generate_rank_k, left_unfolding,  # noqa: F401
                            pad_matrix, perturbe_random, qr_weight, randhaar, randuptri, right_unfolding,
                            svd_weight, unfolding, unpad_matrix)

"""``tn_gradient.layer.tensor_linear``: never instantiated by any reference script and broken there with
bias=True (tn_gradient/layer/tensor_linear.py:48); kept as an explicit out-of-scope marker."""


class TensorTrainLinear:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("TensorTrainLinear is outside the SoW hot path (SURVEY.md section 2, row 6)")


class ComposedLinear(TensorTrainLinear):  # pragma: no cover
    pass

"""gpflow.mean_functions Zero / Identity / Linear (call site reference layers.py:46; built at
experiments/build_models.py:205-209, tests/test_gp_layer.py:29,85).  Evaluated inside the fused per-point epilogue
of csrc/gp_rows_fwd.cu."""
import numpy as np

from .params import Parameter, Parameterized


class MeanFunction(Parameterized):
    kind = None


class Zero(MeanFunction):
    kind = 'Zero'

    def __init__(self, output_dim=1, name=None):
        Parameterized.__init__(self, name=name)
        self.output_dim = output_dim


class Identity(MeanFunction):
    kind = 'Identity'

    def __init__(self, input_dim=None, name=None):
        Parameterized.__init__(self, name=name)
        self.input_dim = input_dim


class Linear(MeanFunction):
    kind = 'Linear'

    def __init__(self, A=None, b=None, name=None):
        Parameterized.__init__(self, name=name)
        A = np.ones((1, 1)) if A is None else np.asarray(A, dtype=np.float64)
        b = np.zeros(A.shape[1]) if b is None else np.asarray(b, dtype=np.float64).reshape(-1)
        if b.size == 1 and A.shape[1] > 1:
            b = np.full(A.shape[1], float(b[0]))
        self.A = Parameter(A)
        self.b = Parameter(b)

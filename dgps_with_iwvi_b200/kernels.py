"""Descriptors of the GPflow 1.x stationary kernels the path evaluates (call sites: reference
temp_workaround.py:39,44,45; built at experiments/build_models.py:212,238 and tests/test_gp_layer.py:28,66).  The
arithmetic (r2 = |x/l|^2 + |z/l|^2 - 2 x.z/l^2, not clamped for RBF; r = sqrt(max(r2, 1e-40)) for the Matern
family) lives in csrc/common.cuh; these classes only hold the hyper-parameters."""
import numpy as np

from .params import Parameter, Parameterized, positive


class Stationary(Parameterized):
    kind = None

    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=None, name=None):
        Parameterized.__init__(self, name=name)
        self.input_dim = int(input_dim)
        if active_dims is not None and list(active_dims) != list(range(self.input_dim)):
            raise NotImplementedError('only the default active_dims (first input_dim columns) are on the path')
        if lengthscales is None:
            lengthscales = np.ones(self.input_dim) if ARD else 1.0
        lengthscales = np.asarray(lengthscales, dtype=np.float64)
        if ARD is None:
            ARD = lengthscales.ndim > 0 and lengthscales.size > 1
        if ARD and lengthscales.size == 1:
            lengthscales = np.full(self.input_dim, float(lengthscales.reshape(-1)[0]))
        self.ARD = bool(ARD)
        self.variance = Parameter(np.asarray(variance, dtype=np.float64).reshape(()), transform=positive)
        self.lengthscales = Parameter(lengthscales if self.ARD else lengthscales.reshape(()), transform=positive)


class RBF(Stationary):
    kind = 'RBF'


SquaredExponential = RBF


class Matern12(Stationary):
    kind = 'Matern12'


class Matern32(Stationary):
    kind = 'Matern32'


class Matern52(Stationary):
    kind = 'Matern52'


class SharedMixedMok(Parameterized):
    """Reference temp_workaround.py:107-115: latent GPs share `kernel`; outputs are mixed by W [P, L]."""

    def __init__(self, kernel, W, name=None):
        Parameterized.__init__(self, name=name)
        self.kernel = kernel
        self.W = Parameter(W)

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
        # gpflow: the kernel acts on the columns `active_dims` of its inputs (default: the first input_dim columns).  The
        # CUDA path consumes 1 / lengthscale per input column, so a restricted kernel is the full-width ARD kernel with
        # 1 / lengthscale = 0 on the inactive columns (engine / temp_workaround expand the lengthscales accordingly).
        self.active_dims = None if active_dims is None else [int(a) for a in active_dims]
        if self.active_dims is not None and len(self.active_dims) != self.input_dim:
            raise ValueError('active_dims must name input_dim = %d columns' % self.input_dim)
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


    def full_lengthscales(self, ls, D):
        """Constrained lengthscales `ls` (tensor: scalar or [input_dim]) as a width-D vector for the kernels: +inf on
        columns the kernel does not act on."""
        import torch
        if self.active_dims is None and (ls.dim() == 0 or ls.numel() == 1):
            return ls.expand(D) if self.input_dim == D else self._scatter(ls.expand(self.input_dim), list(range(self.input_dim)), D)
        if self.active_dims is None and ls.numel() == D:
            return ls
        dims = self.active_dims if self.active_dims is not None else list(range(self.input_dim))
        return self._scatter(ls.expand(len(dims)) if ls.numel() == 1 else ls, dims, D)

    @staticmethod
    def _scatter(vals, dims, D):
        import torch
        out = torch.full((D,), float('inf'), dtype=vals.dtype, device=vals.device)
        idx = torch.as_tensor(dims, dtype=torch.int64, device=vals.device)
        return out.index_copy(0, idx, vals.reshape(-1))


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

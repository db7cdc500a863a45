"""B200-native importance-weighted ELBO hot path of DGPs_with_IWVI behind the reference's layer/model API."""
from . import settings
from .features import InducingPoints, MixedKernelSharedMof
from .kernels import RBF, Matern12, Matern32, Matern52, SharedMixedMok, SquaredExponential
from .layers import Encoder, GPLayer, LatentVariableLayer, RegularizerType
from .likelihoods import Gaussian
from .mean_functions import Identity, Linear, Zero
from .models import DGP_IWVI, DGP_VI

__all__ = ['settings', 'InducingPoints', 'MixedKernelSharedMof', 'RBF', 'SquaredExponential', 'Matern12', 'Matern32',
           'Matern52', 'SharedMixedMok', 'Encoder', 'GPLayer', 'LatentVariableLayer', 'RegularizerType', 'Gaussian',
           'Identity', 'Linear', 'Zero', 'DGP_IWVI', 'DGP_VI']

"""GPLayer / LatentVariableLayer / Encoder with the reference's constructor signatures, attribute names and
`propagate(F, ...) -> (samples, mean, cov, kl)` protocol (reference dgps_with_iwvi/layers.py:14-152).  The arithmetic
runs in the CUDA library through the operator functions of temp_workaround.py (same module name as the reference's
operator file); models built from these layers train through engine.Engine, which calls the same C ABI without
autograd."""
import enum

import numpy as np

from . import settings
from .features import InducingFeature, InducingPoints
from .mean_functions import Zero
from .params import LowerTriangular, Parameter, Parameterized, ParamList


class RegularizerType(enum.Enum):   # reference layers.py:9-11
    LOCAL = 0
    GLOBAL = 1


class GPLayer(Parameterized):
    """reference layers.py:14-50.  q_mu = 0 [M, R]; q_sqrt = I [R, M, M] (lower-triangular transform)."""
    regularizer_type = RegularizerType.GLOBAL

    def __init__(self, kern, Z, num_outputs, mean_function=None, name=None):
        Parameterized.__init__(self, name=name)
        self.num_inducing = len(Z)
        self.q_mu = Parameter(np.zeros((self.num_inducing, num_outputs)))
        q_sqrt = np.tile(np.eye(self.num_inducing)[None, :, :], [num_outputs, 1, 1])
        self.q_sqrt = Parameter(q_sqrt, transform=LowerTriangular(self.num_inducing, num_matrices=num_outputs))
        self.feature = Z if isinstance(Z, InducingFeature) else InducingPoints(Z)
        self.kern = kern
        self.mean_function = mean_function or Zero()
        self.num_outputs = num_outputs
        self.jitter = settings.jitter

    def propagate(self, F, full_cov=False, eps=None, **kwargs):
        """F [..., D_in] CUDA float64 tensor -> (samples, mean, cov, kl); differentiable (torch autograd) with respect
        to F, and to the layer's parameters once they are marked as leaves (`layer.requires_grad_()`; a Parameter's
        storage does not require grad by default).  `eps` injects the N(0,1) draw of temp_workaround.py:89."""
        from . import temp_workaround as tw
        samples, mean, cov = tw.multisample_sample_conditional(
            F, self.feature, self.kern, self.q_mu, full_cov=full_cov, q_sqrt=self.q_sqrt, white=True,
            mean_function=self.mean_function, jitter=self.jitter, eps=eps)
        kl = tw.gauss_kl(self.q_mu, self.q_sqrt)
        return samples, mean, cov, kl


class Encoder(Parameterized):
    """reference layers.py:108-152: MLP [input_dim, *network_dims, 2*latent_dim] with `activation_func` between the
    layers (default tanh, :122), skip connection after the activation where widths match, sigma = softplus(raw - 3).
    Xavier-normal weights, zero biases (:129-131).  The reference takes a TensorFlow op; here the non-linearity is named:
    'tanh' | 'relu' | 'sigmoid' | 'softplus' | 'elu' | 'identity', or a callable whose __name__ is one of those
    (torch.tanh, torch.relu, torch.nn.functional.softplus, ...) -- the fused encoder kernel evaluates it on the device."""
    ACTIVATIONS = ('tanh', 'relu', 'sigmoid', 'softplus', 'elu', 'identity')

    def __init__(self, latent_dim, input_dim, network_dims, activation_func=None, name=None, seed=None):
        Parameterized.__init__(self, name=name)
        act = activation_func if activation_func is not None else 'tanh'
        if not isinstance(act, str):
            act = getattr(act, '__name__', str(act)).lower()
        if act not in self.ACTIVATIONS:
            raise NotImplementedError('Encoder(activation_func=%r): the fused encoder kernel implements %s'
                                      % (activation_func, ', '.join(self.ACTIVATIONS)))
        self.activation_func = act
        self.latent_dim = latent_dim
        self.layer_dims = [input_dim, *network_dims, latent_dim * 2]
        rng = np.random if seed is None else np.random.default_rng(seed)
        Ws, bs = [], []
        for din, dout in zip(self.layer_dims[:-1], self.layer_dims[1:]):
            xavier_std = (2. / (din + dout)) ** 0.5
            W = (rng.randn(din, dout) if seed is None else rng.standard_normal((din, dout))) * xavier_std
            Ws.append(Parameter(W))
            bs.append(Parameter(np.zeros(dout)))
        self.Ws, self.bs = ParamList(Ws), ParamList(bs)

    def named_parameters(self, prefix=''):
        # W0, b0, W1, b1, ...: the packed order the C ABI reads (include/iwvi_b200.h, iwvi_lv_fwd)
        for i, (W, b) in enumerate(zip(self.Ws, self.bs)):
            yield '%sWs.%d' % (prefix, i), W
            yield '%sbs.%d' % (prefix, i), b

    def __call__(self, Z):
        from . import temp_workaround as tw
        return tw.encoder_forward(self, Z)


class LatentVariableLayer(Parameterized):
    """reference layers.py:53-105.  `prior_mu` / `prior_sigma` stand in for the q_mu / q_sqrt placeholders with
    default (:62-64): the values used when no amortisation inputs are given."""
    regularizer_type = RegularizerType.LOCAL

    def __init__(self, latent_dim, XY_dim=None, encoder=None, name=None):
        Parameterized.__init__(self, name=name)
        self.latent_dim = latent_dim
        self.prior_mu, self.prior_sigma = 0.0, 1.0
        if encoder is None:
            assert XY_dim, 'must pass XY_dim or else an encoder'
            encoder = Encoder(latent_dim, XY_dim, [20, 20])
        self.encoder = encoder

    def propagate(self, F, inference_amorization_inputs=None, is_sampled_local_regularizer=False, eps=None, **kwargs):
        from . import temp_workaround as tw
        return tw.latent_variable_propagate(self, F, inference_amorization_inputs, is_sampled_local_regularizer, eps)

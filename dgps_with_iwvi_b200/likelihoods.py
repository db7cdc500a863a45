"""gpflow.likelihoods.Gaussian (call sites reference models.py:66,105,134).  variational_expectations is fused into
csrc/lv_elbo.cu (elbo_fwd_kernel); predict_mean_and_var is a one-line host op on the prediction path."""
import numpy as np

from .params import Parameter, Parameterized, positive


class Gaussian(Parameterized):
    def __init__(self, variance=1.0, name=None):
        Parameterized.__init__(self, name=name)
        self.variance = Parameter(np.asarray(variance, dtype=np.float64).reshape(()), transform=positive)

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + self.variance.value

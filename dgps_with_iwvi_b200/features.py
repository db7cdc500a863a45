"""gpflow.features.InducingPoints and gpflow.multioutput.MixedKernelSharedMof as used at reference layers.py:28 and
experiments/build_models.py:221-224."""
import numpy as np

from .params import Parameter, Parameterized


class InducingFeature(Parameterized):
    pass


class InducingPoints(InducingFeature):
    def __init__(self, Z, name=None):
        Parameterized.__init__(self, name=name)
        self.Z = Parameter(np.asarray(Z, dtype=np.float64))

    def __len__(self):
        return self.Z.shape[0]


class MixedKernelSharedMof(InducingFeature):
    def __init__(self, feat, name=None):
        Parameterized.__init__(self, name=name)
        self.feat = feat

    def __len__(self):
        return len(self.feat)

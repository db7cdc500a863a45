"""Minimal stand-ins for the gpflow.params / gpflow.transforms objects the reference's layer and model classes are
written against (reference layers.py:20-25,129-134; build_models.py:209,213,225-227,275-287; tests/test_gp_layer.py:46-47,91):
Parameter (constrained value <-> unconstrained storage, trainable flag, assign / read_value), ParamList,
Parameterized (attribute assignment of a value onto an existing Parameter assigns it, as GPflow does)."""
import math

import numpy as np
import torch

from . import settings


class Identity:
    name = 'identity'

    def forward(self, x):
        return x

    def backward(self, y):
        return y


class Log1pe:
    """gpflow.transforms.positive: theta = softplus(x) + 1e-6."""
    name = 'positive'
    lower = 1e-6

    def forward(self, x):
        return torch.nn.functional.softplus(x) + self.lower

    def backward(self, y):
        y = y - self.lower
        return y + torch.log(-torch.expm1(-y))


positive = Log1pe()


class LowerTriangular:
    """gpflow.transforms.LowerTriangular(M, num_matrices=R).  GPflow stores the packed lower triangle; here the
    storage is the full [R, M, M] array whose strict upper triangle is ignored (tril on read) and receives an
    exactly-zero gradient from the kernels, which is equivalent for any gradient-based update."""
    name = 'lower_triangular'

    def __init__(self, M, num_matrices=1):
        self.M, self.num_matrices = M, num_matrices

    def forward(self, x):
        return torch.tril(x)

    def backward(self, y):
        return torch.tril(y)


class Parameter:
    def __init__(self, value, transform=None, trainable=True, name=None):
        self.transform = transform or Identity()
        self.trainable = trainable
        self.name = name
        v = torch.as_tensor(np.asarray(value, dtype=np.float64), dtype=settings.float_type).to(settings.device())
        self.unconstrained = self.transform.backward(v).contiguous()

    @property
    def shape(self):
        return tuple(self.unconstrained.shape)

    @property
    def size(self):
        return self.unconstrained.numel()

    @property
    def value(self):
        """Constrained value as a torch tensor (on the parameter's device)."""
        return self.transform.forward(self.unconstrained)

    def read_value(self):
        return self.value.detach().cpu().numpy()

    def assign(self, value):
        v = torch.as_tensor(np.asarray(value, dtype=np.float64), dtype=settings.float_type).to(self.unconstrained.device)
        if tuple(v.shape) != self.shape:
            v = v.reshape(self.shape)
        self.unconstrained.copy_(self.transform.backward(v))   # in place: storage may be a view of a flat buffer

    def set_trainable(self, flag):
        self.trainable = bool(flag)

    def _rebind(self, view):
        """Moves the storage into `view` (a slice of a flat buffer), keeping the value."""
        view.copy_(self.unconstrained.reshape(view.shape))
        self.unconstrained = view


Param = Parameter


class ParamList:
    def __init__(self, items):
        self._items = list(items)

    def __iter__(self):
        return iter(self._items)

    def __len__(self):
        return len(self._items)

    def __getitem__(self, i):
        return self._items[i]

    def set_trainable(self, flag):
        for it in self._items:
            it.set_trainable(flag)


class Parameterized:
    def __init__(self, name=None):
        object.__setattr__(self, 'name', name)

    def __setattr__(self, key, value):
        cur = self.__dict__.get(key)
        if isinstance(cur, Parameter) and not isinstance(value, Parameter):
            cur.assign(value)
        else:
            object.__setattr__(self, key, value)

    def named_parameters(self, prefix=''):
        """Depth-first, in attribute definition order."""
        for k, v in self.__dict__.items():
            if isinstance(v, Parameter):
                yield prefix + k, v
            elif isinstance(v, Parameterized):
                yield from v.named_parameters(prefix + k + '.')
            elif isinstance(v, ParamList):
                for i, it in enumerate(v):
                    if isinstance(it, Parameter):
                        yield '%s%s.%d' % (prefix, k, i), it
                    elif isinstance(it, Parameterized):
                        yield from it.named_parameters('%s%s.%d.' % (prefix, k, i))

    def set_trainable(self, flag):
        for _, p in self.named_parameters():
            p.set_trainable(flag)

    def requires_grad_(self, flag=True):
        """Operator-level autograd (layer.propagate / model.propagate composed with torch): marks every parameter's
        unconstrained storage as a differentiable leaf, so that `.value` carries gradients back to it (through the
        transform) and `p.unconstrained.grad` is filled by backward().  Whole-model training does not need this: it runs
        through engine.Engine, whose hand-written backward writes the flat gradient bucket."""
        for _, p in self.named_parameters():
            p.unconstrained.requires_grad_(flag)
        return self

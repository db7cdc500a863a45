"""
TEST INFRASTRUCTURE ONLY -- numpy restatement of the Adam half of the reference's training iteration
(experiments/build_models.py:289-300):

    lr   = tf.train.exponential_decay(ARGS.lr, global_step, 1000, ARGS.lr_decay, staircase=True)      :291
    op_adam = AdamOptimizer(lr).make_optimize_tensor(model)                                           :295
    train(): s.run(op_increment); s.run(op_ng); s.run(op_adam)                                        :297-300

i.e. iteration t = 1, 2, ... runs with global_step = t (the increment comes first), so the staircase factor
rate ** (t // 1000) first changes AT iteration 1000.  tf.train.AdamOptimizer (defaults beta1 0.9, beta2 0.999,
epsilon 1e-8) minimises -ELBO over GPflow's UNCONSTRAINED variables:

    g    = d(-ELBO)/dx = -(dELBO/dtheta) * dtheta/dx;  theta = softplus(x) + 1e-6 (gpflow.transforms.positive / Log1pe)
                                                        => dtheta/dx = sigmoid(x); identity transform => 1
    m    = b1 m + (1 - b1) g;   v = b2 v + (1 - b2) g^2
    lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);   x -= lr_t m / (sqrt(v) + eps)          (TF's "epsilon hat" form)

Frozen entries (set_trainable(False), build_models.py:209,213,225-227,286-287) are not in TF's var_list: they keep
their value and have no slots.
"""
import numpy as np


def staircase_decay(base, global_step, decay_steps=1000, rate=0.98):
    """tf.train.exponential_decay(base, global_step, decay_steps, rate, staircase=True)."""
    return base * rate ** (int(global_step) // int(decay_steps))


def positive_forward(x):
    return np.logaddexp(0.0, x) + 1e-6


def adam_step(x, g_elbo_constrained, m, v, t, lr, n_pos, mask=None, b1=0.9, b2=0.999, eps=1e-8):
    """One step.  x [n] unconstrained (first n_pos entries positive-transformed); g_elbo_constrained [n] = dELBO/dtheta;
    t = 1, 2, ...; lr already decayed.  Returns (x, m, v) new arrays."""
    x, m, v = np.array(x, dtype=np.float64), np.array(m, dtype=np.float64), np.array(v, dtype=np.float64)
    g = -np.asarray(g_elbo_constrained, dtype=np.float64).copy()
    with np.errstate(over='ignore'):                        # exp(745) -> inf -> sigmoid 0, as intended
        g[:n_pos] *= 1.0 / (1.0 + np.exp(-x[:n_pos]))
    live = np.ones_like(x, dtype=bool) if mask is None else (np.asarray(mask) != 0)
    g = np.where(live, g, 0.0)
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    x = np.where(live, x - lr_t * m / (np.sqrt(v) + eps), x)
    return x, m, v

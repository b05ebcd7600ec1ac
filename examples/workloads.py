"""The five BASELINE.json workloads as plain ``autograd.numpy`` user programs.

Each function is written against ``autograd.numpy`` only, so the very same code
runs on host ndarrays (reference autograd on NumPy — the CPU oracle) and on
``autograd_b200.DeviceArray`` values (the sm_100a backend).  They restate the
model code of the reference's example scripts, which cannot be imported as-is
(module-level matplotlib / dataset downloads); the arithmetic and its order
follow the cited lines so that host results are bit-identical to the originals
(pinned by tests/golden, generated from the real example files).

  C1  mlp_*       examples/neural_net.py:15-48
  C2  tanh_*      examples/tanh.py:22-33
  C3  lstm_*      examples/lstm.py:16-64, examples/rnn.py:16-23
  C4  fluid_*     examples/fluidsim/fluidsim.py:15-82,108-119
  C5  convnet_*   examples/convnet.py:15-162
"""

from __future__ import annotations

import itertools

import numpy as onp

import autograd.numpy as np
from autograd import elementwise_grad, grad, value_and_grad  # noqa: F401  (re-exported for callers)
from autograd.misc.flatten import flatten


def _logsumexp():
    # looked up at call time: autograd_b200.install() rebinds this attribute to a dispatching primitive
    import autograd.scipy.special as special

    return special.logsumexp


def _convolve():
    import autograd.scipy.signal as signal

    return signal.convolve


# ===================================================================== C1: MLP log-posterior
def mlp_init_params(scale, layer_sizes, rs):
    """List of (W, b) per layer (neural_net.py:15-25); same RandomState call order."""
    params = []
    for m, n in itertools.pairwise(layer_sizes):
        W = scale * rs.randn(m, n)
        b = scale * rs.randn(n)
        params.append((W, b))
    return params


def mlp_predict(params, inputs):
    """Normalised class log-probabilities (neural_net.py:28-36)."""
    logsumexp = _logsumexp()
    for W, b in params:
        outputs = np.dot(inputs, W) + b
        inputs = np.tanh(outputs)
    return outputs - logsumexp(outputs, axis=1, keepdims=True)


def mlp_log_posterior(params, inputs, targets, L2_reg):
    """neural_net.py:39-48."""
    flat, _ = flatten(params)
    log_prior = -L2_reg * np.dot(flat, flat)
    log_lik = np.sum(mlp_predict(params, inputs) * targets)
    return log_prior + log_lik


def mlp_objective(params, inputs, targets, L2_reg):
    return -mlp_log_posterior(params, inputs, targets, L2_reg)


def mlp_data(batch=256, layer_sizes=(784, 200, 100, 10), seed=0):
    """Synthetic MNIST-shaped inputs of SURVEY.md §8(d) C1."""
    rs = onp.random.RandomState(seed)
    params = mlp_init_params(0.1, list(layer_sizes), rs)
    X = rs.rand(batch, layer_sizes[0])
    T = onp.eye(layer_sizes[-1])[rs.randint(0, layer_sizes[-1], batch)]
    return params, X, T


# ===================================================================== C2: nested elementwise_grad of tanh
def tanh_fn(x):
    """tanh.py:22-23."""
    return (1.0 - np.exp(-2 * x)) / (1.0 + np.exp(-(2 * x)))


def tanh_nth_derivative(order=4):
    f = tanh_fn
    for _ in range(order):
        f = elementwise_grad(f)
    return f


def tanh_data(n=2**28, lo=-7.0, hi=7.0):
    return onp.linspace(lo, hi, n)


# ===================================================================== C3: LSTM BPTT
def lstm_init_params(input_size, state_size, output_size, param_scale=0.01, rs=None, dtype=None):
    """lstm.py:16-30 (same key order → same RandomState stream)."""
    rs = rs if rs is not None else onp.random.RandomState(0)
    shapes = [
        ("init cells", (1, state_size)),
        ("init hiddens", (1, state_size)),
        ("change", (input_size + state_size + 1, state_size)),
        ("forget", (input_size + state_size + 1, state_size)),
        ("ingate", (input_size + state_size + 1, state_size)),
        ("outgate", (input_size + state_size + 1, state_size)),
        ("predict", (state_size + 1, output_size)),
    ]
    params = {}
    for name, shape in shapes:
        p = rs.randn(*shape) * param_scale
        params[name] = p if dtype is None else p.astype(dtype)
    return params


def _sigmoid(x):
    return 0.5 * (np.tanh(x) + 1.0)  # rnn.py:16-17


def _concat_and_multiply(weights, *args):
    cat_state = np.hstack(args + (np.ones((args[0].shape[0], 1)),))  # rnn.py:20-22
    return np.dot(cat_state, weights)


def lstm_predict(params, inputs):
    """lstm.py:33-55."""
    logsumexp = _logsumexp()

    def step(x_t, hiddens, cells):
        change = np.tanh(_concat_and_multiply(params["change"], x_t, hiddens))
        forget = _sigmoid(_concat_and_multiply(params["forget"], x_t, hiddens))
        ingate = _sigmoid(_concat_and_multiply(params["ingate"], x_t, hiddens))
        outgate = _sigmoid(_concat_and_multiply(params["outgate"], x_t, hiddens))
        cells = cells * forget + ingate * change
        hiddens = outgate * np.tanh(cells)
        return hiddens, cells

    def readout(hiddens):
        output = _concat_and_multiply(params["predict"], hiddens)
        return output - logsumexp(output, axis=1, keepdims=True)

    num_sequences = inputs.shape[1]
    hiddens = np.repeat(params["init hiddens"], num_sequences, axis=0)
    cells = np.repeat(params["init cells"], num_sequences, axis=0)
    outputs = [readout(hiddens)]
    for x_t in inputs:
        hiddens, cells = step(x_t, hiddens, cells)
        outputs.append(readout(hiddens))
    return outputs


def lstm_log_likelihood(params, inputs, targets):
    """lstm.py:58-64."""
    logprobs = lstm_predict(params, inputs)
    loglik = 0.0
    num_time_steps, num_examples, _ = inputs.shape
    for t in range(num_time_steps):
        loglik += np.sum(logprobs[t] * targets[t])
    return loglik / (num_time_steps * num_examples)


def lstm_objective(params, inputs, targets):
    return -lstm_log_likelihood(params, inputs, targets)


def lstm_data(seq_len=256, batch=1024, hidden=512, chars=128, seed=0, dtype=onp.float32):
    rs = onp.random.RandomState(seed)
    params = lstm_init_params(chars, hidden, chars, param_scale=0.01, rs=rs, dtype=dtype)
    inputs = onp.eye(chars, dtype=dtype)[rs.randint(0, chars, (seq_len, batch))]
    return params, inputs, inputs


# ===================================================================== C4: fluid simulation
def fluid_project(vx, vy):
    """Approximately mass-conserving projection, 10 Jacobi sweeps (fluidsim.py:15-42)."""
    p = np.zeros(vx.shape)
    h = 1.0 / vx.shape[0]
    div = (
        -0.5
        * h
        * (np.roll(vx, -1, axis=0) - np.roll(vx, 1, axis=0) + np.roll(vy, -1, axis=1) - np.roll(vy, 1, axis=1))
    )
    for _ in range(10):
        p = (div + np.roll(p, 1, axis=0) + np.roll(p, -1, axis=0) + np.roll(p, 1, axis=1) + np.roll(p, -1, axis=1)) / 4.0
    vx = vx - 0.5 * (np.roll(p, -1, axis=0) - np.roll(p, 1, axis=0)) / h
    vy = vy - 0.5 * (np.roll(p, -1, axis=1) - np.roll(p, 1, axis=1)) / h
    return vx, vy


def fluid_advect(f, vx, vy):
    """Semi-Lagrangian advection with bilinear interpolation (fluidsim.py:45-68)."""
    rows, cols = f.shape
    cell_ys, cell_xs = np.meshgrid(np.arange(rows), np.arange(cols))
    center_xs = (cell_xs - vx).ravel()
    center_ys = (cell_ys - vy).ravel()
    left_ix = np.floor(center_xs).astype(int)
    top_ix = np.floor(center_ys).astype(int)
    rw = center_xs - left_ix
    bw = center_ys - top_ix
    left_ix = np.mod(left_ix, rows)
    right_ix = np.mod(left_ix + 1, rows)
    top_ix = np.mod(top_ix, cols)
    bot_ix = np.mod(top_ix + 1, cols)
    flat_f = (1 - rw) * ((1 - bw) * f[left_ix, top_ix] + bw * f[left_ix, bot_ix]) + rw * (
        (1 - bw) * f[right_ix, top_ix] + bw * f[right_ix, bot_ix]
    )
    return np.reshape(flat_f, (rows, cols))


def fluid_simulate(vx, vy, smoke, num_time_steps):
    """fluidsim.py:71-82 without the plotting hooks."""
    for _ in range(num_time_steps):
        vx_updated = fluid_advect(vx, vx, vy)
        vy_updated = fluid_advect(vy, vx, vy)
        vx, vy = fluid_project(vx_updated, vy_updated)
        smoke = fluid_advect(smoke, vx, vy)
    return smoke


def fluid_objective(params, init_smoke, target, rows, cols, steps):
    """fluidsim.py:108-119."""
    vx = np.reshape(params[: (rows * cols)], (rows, cols))
    vy = np.reshape(params[(rows * cols):], (rows, cols))
    final_smoke = fluid_simulate(vx, vy, init_smoke, steps)
    return np.mean((target - final_smoke) ** 2)


def fluid_data(rows=2048, cols=2048, seed=0):
    rs = onp.random.RandomState(seed)
    init_smoke = rs.rand(rows, cols)
    target = rs.rand(rows, cols)
    params = rs.randn(2 * rows * cols) * 0.5
    return params, init_smoke, target


# ===================================================================== C5: LeNet-style convnet
class _Slots:
    """Name → (slice, shape) bookkeeping over one flat parameter vector (convnet.py:15-28)."""

    def __init__(self):
        self.table = {}
        self.N = 0

    def add(self, name, shape):
        start = self.N
        self.N += int(onp.prod(shape))
        self.table[name] = (slice(start, self.N), shape)

    def get(self, vect, name):
        idxs, shape = self.table[name]
        return np.reshape(vect[idxs], shape)


def _global_max_logsumexp(X, axis, keepdims=False):
    max_X = np.max(X)  # convnet.py:40-42: shifts by the GLOBAL max
    return max_X + np.log(np.sum(np.exp(X - max_X), axis=axis, keepdims=keepdims))


class ConvLayer:
    def __init__(self, kernel_shape, num_filters):
        self.kernel_shape, self.num_filters = kernel_shape, num_filters

    def build(self, input_shape):
        self.slots = _Slots()
        self.slots.add("params", (input_shape[0], self.num_filters) + self.kernel_shape)
        self.slots.add("biases", (1, self.num_filters, 1, 1))
        out = (self.num_filters,) + tuple(a - b + 1 for a, b in zip(input_shape[1:], self.kernel_shape))
        return self.slots.N, out

    def forward(self, inputs, vect):
        params = self.slots.get(vect, "params")
        biases = self.slots.get(vect, "biases")
        conv = _convolve()(inputs, params, axes=([2, 3], [2, 3]), dot_axes=([1], [0]), mode="valid")
        return conv + biases


class MaxPoolLayer:
    def __init__(self, pool_shape):
        self.pool_shape = pool_shape

    def build(self, input_shape):
        out = list(input_shape)
        for i in (0, 1):
            assert input_shape[i + 1] % self.pool_shape[i] == 0, "maxpool shape should tile input exactly"
            out[i + 1] = input_shape[i + 1] // self.pool_shape[i]
        return 0, tuple(out)

    def forward(self, inputs, vect):
        new_shape = inputs.shape[:2]
        for i in (0, 1):
            pw = self.pool_shape[i]
            new_shape += (inputs.shape[i + 2] // pw, pw)
        result = inputs.reshape(new_shape)
        return np.max(np.max(result, axis=3), axis=4)


class FullLayer:
    def __init__(self, size, nonlinearity):
        self.size, self.nonlinearity = size, nonlinearity

    def build(self, input_shape):
        input_size = int(onp.prod(input_shape, dtype=int))
        self.slots = _Slots()
        self.slots.add("params", (input_size, self.size))
        self.slots.add("biases", (self.size,))
        return self.slots.N, (self.size,)

    def forward(self, inputs, vect):
        params = self.slots.get(vect, "params")
        biases = self.slots.get(vect, "biases")
        if inputs.ndim > 2:
            inputs = inputs.reshape((inputs.shape[0], int(onp.prod(inputs.shape[1:]))))
        return self.nonlinearity(np.dot(inputs[:, :], params) + biases)


def _softmax_nl(x):
    return x - _global_max_logsumexp(x, axis=1, keepdims=True)


def convnet_make(input_shape=(1, 28, 28), L2_reg=1.0):
    """LeNet spec of convnet.py:154-162; returns (num_weights, predictions, loss)."""
    layers = [
        ConvLayer((5, 5), 6),
        MaxPoolLayer((2, 2)),
        ConvLayer((5, 5), 16),
        MaxPoolLayer((2, 2)),
        FullLayer(120, np.tanh),
        FullLayer(84, np.tanh),
        FullLayer(10, _softmax_nl),
    ]
    slots = _Slots()
    cur = input_shape
    for layer in layers:
        n, cur = layer.build(cur)
        slots.add(layer, (n,))

    def predictions(W_vect, inputs):
        cur_units = inputs
        for layer in layers:
            cur_units = layer.forward(cur_units, slots.get(W_vect, layer))
        return cur_units

    def loss(W_vect, X, T, l2=L2_reg):
        log_prior = -l2 * np.dot(W_vect, W_vect)
        log_lik = np.sum(predictions(W_vect, X) * T)
        return -log_prior - log_lik

    return slots.N, predictions, loss


def convnet_data(batch, num_weights, seed=0):
    rs = onp.random.RandomState(seed)
    W = rs.randn(num_weights) * 0.1
    X = rs.rand(batch, 1, 28, 28)
    T = onp.eye(10)[rs.randint(0, 10, batch)]
    return W, X, T


# ===================================================================== SURVEY §8(f) row 4: extra workloads
# examples/rnn.py:27-60, examples/bayesian_neural_net.py:11-52, examples/black_box_svi.py:11-36,56-61 restated the same way
# (the example modules import matplotlib at the top); pinned to the real example code by tests/golden/f4_*.npz.
def rnn_init_params(input_size, state_size, output_size, param_scale=0.01, rs=None):
    rs = rs if rs is not None else onp.random.RandomState(0)
    return {"init hiddens": rs.randn(1, state_size) * param_scale,
            "change": rs.randn(input_size + state_size + 1, state_size) * param_scale,
            "predict": rs.randn(state_size + 1, output_size) * param_scale}


def rnn_predict(params, inputs):
    logsumexp = _logsumexp()

    def update_rnn(input, hiddens):
        return np.tanh(_concat_and_multiply(params["change"], input, hiddens))

    def hiddens_to_output_probs(hiddens):
        output = _concat_and_multiply(params["predict"], hiddens)
        return output - logsumexp(output, axis=1, keepdims=True)

    num_sequences = inputs.shape[1]
    hiddens = np.repeat(params["init hiddens"], num_sequences, axis=0)
    output = [hiddens_to_output_probs(hiddens)]
    for input in inputs:
        hiddens = update_rnn(input, hiddens)
        output.append(hiddens_to_output_probs(hiddens))
    return output


def rnn_log_likelihood(params, inputs, targets):
    logprobs = rnn_predict(params, inputs)
    loglik = 0.0
    num_time_steps, num_examples, _ = inputs.shape
    for t in range(num_time_steps):
        loglik += np.sum(logprobs[t] * targets[t])
    return loglik / (num_time_steps * num_examples)


def rnn_objective(params, inputs, targets):
    return -rnn_log_likelihood(params, inputs, targets)


def rnn_data(seq_len=6, batch=8, hidden=16, chars=12, seed=0):
    rs = onp.random.RandomState(seed)
    params = rnn_init_params(chars, hidden, chars, param_scale=0.01, rs=rs)
    inputs = onp.eye(chars)[rs.randint(0, chars, (seq_len, batch))]
    return params, inputs, inputs


def bnn_make_nn_funs(layer_sizes, L2_reg, noise_variance, nonlinearity=np.tanh):
    """MLP vectorised over training examples AND weight samples (bayesian_neural_net.py:11-42)."""
    shapes = list(itertools.pairwise(layer_sizes))
    num_weights = sum((m + 1) * n for m, n in shapes)

    def unpack_layers(weights):
        num_weight_sets = len(weights)
        for m, n in shapes:
            yield (weights[:, : m * n].reshape((num_weight_sets, m, n)),
                   weights[:, m * n: m * n + n].reshape((num_weight_sets, 1, n)))
            weights = weights[:, (m + 1) * n:]

    def predictions(weights, inputs):
        inputs = np.expand_dims(inputs, 0)
        for W, b in unpack_layers(weights):
            outputs = np.einsum("mnd,mdo->mno", inputs, W) + b
            inputs = nonlinearity(outputs)
        return outputs

    def logprob(weights, inputs, targets):
        log_prior = -L2_reg * np.sum(weights**2, axis=1)
        preds = predictions(weights, inputs)
        log_lik = -np.sum((preds - targets) ** 2, axis=1)[:, 0] / noise_variance
        return log_prior + log_lik

    return num_weights, predictions, logprob


def bnn_toy_dataset(n_data=40, noise_std=0.1):
    """bayesian_neural_net.py:45-52."""
    D = 1
    rs = onp.random.RandomState(0)
    inputs = onp.concatenate([onp.linspace(0, 2, num=n_data // 2), onp.linspace(6, 8, num=n_data // 2)])
    targets = onp.cos(inputs) + rs.randn(n_data) * noise_std
    inputs = (inputs - 4.0) / 4.0
    return inputs.reshape((len(inputs), D)), targets.reshape((len(targets), D))


def svi_objective(logprob, D, num_samples, seed=0):
    """Stochastic variational lower bound of a diagonal Gaussian (black_box_svi.py:11-36).  The standard-normal draws
    come from a host RandomState exactly as in the example; they meet the device parameters in ``* np.exp(log_std)``."""
    rs = onp.random.RandomState(seed)

    def gaussian_entropy(log_std):
        return 0.5 * D * (1.0 + np.log(2 * np.pi)) + np.sum(log_std)

    def variational_objective(params, t):
        mean, log_std = params[:D], params[D:]
        samples = rs.randn(num_samples, D) * np.exp(log_std) + mean
        lower_bound = gaussian_entropy(log_std) + np.mean(logprob(samples, t))
        return -lower_bound

    return variational_objective


def svi_funnel_log_density(x, t):
    """The example's target density (black_box_svi.py:56-61)."""
    import autograd.scipy.stats.norm as norm

    mu, log_sigma = x[:, 0], x[:, 1]
    sigma_density = norm.logpdf(log_sigma, 0, 1.35)
    mu_density = norm.logpdf(mu, 0, np.exp(log_sigma))
    return sigma_density + mu_density


def fluid_step(vx, vy, smoke):
    """One time step of fluid_simulate (fluidsim.py:73-80) as a function of the whole state."""
    vx_updated = fluid_advect(vx, vx, vy)
    vy_updated = fluid_advect(vy, vx, vy)
    vx, vy = fluid_project(vx_updated, vy_updated)
    return vx, vy, fluid_advect(smoke, vx, vy)


def fluid_objective_checkpointed(params, init_smoke, target, rows, cols, steps, checkpoint=None):
    """fluid_objective with every time step wrapped in a checkpoint — ``autograd.checkpoint``
    (differential_operators.py:230-242) by default, or the wrapper passed in: the memory lever SURVEY §0-6 names for
    tapes that do not fit."""
    from autograd.builtins import tuple as ag_tuple

    if checkpoint is None:
        from autograd.differential_operators import checkpoint

    step = checkpoint(lambda state: ag_tuple(fluid_step(state[0], state[1], state[2])))
    vx = np.reshape(params[: (rows * cols)], (rows, cols))
    vy = np.reshape(params[(rows * cols):], (rows, cols))
    state = ag_tuple((vx, vy, init_smoke))
    for _ in range(steps):
        state = step(state)
    return np.mean((target - state[2]) ** 2)

"""Development harness: gradients of the REFERENCE's example models with device inputs vs the same on NumPy.

Imports the model functions of /root/reference/examples/*.py (matplotlib and the data downloaders stubbed), builds
small seeded inputs, evaluates ``grad`` of each example's objective once on host ndarrays (the reference path) and once
on DeviceArrays (fake backend here, ``--cuda`` for the real one) and compares.  Development container only: the
examples are read from the reference checkout, nothing is copied into the repo.

    python scripts/run_reference_examples_on_device.py [--cuda] [--ref /root/reference] [name ...]
"""
import argparse
import os
import sys
import traceback
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as onp  # noqa: E402


def _stub_modules():
    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Anything(self.__name__ + "." + name)

        def __call__(self, *a, **k):
            return _Anything(self.__name__ + "()")

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "matplotlib.cm", "matplotlib.colors"):
        sys.modules[name] = _Anything(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cuda", action="store_true")
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("names", nargs="*")
    args = ap.parse_args()
    ex = os.path.join(args.ref, "examples")
    if not os.path.isdir(ex):
        raise SystemExit(f"{ex} not found: this harness needs the reference checkout (development container only)")
    sys.dont_write_bytecode = True
    sys.path.insert(0, ex)
    _stub_modules()

    import autograd_b200 as b200

    if args.cuda:
        b200.backend.set_backend(b200.backend.CudaBackend())
    else:
        from fakedev import FakeBackend

        b200.backend.set_backend(FakeBackend(compile_generated=False))
    b200.install()
    import autograd.numpy as np
    from autograd import grad

    import cases

    rs = onp.random.RandomState(0)

    def compare(name, fn, host_args, tol=1e-9, argnum=0):
        with b200.host_arrays(), onp.errstate(all="ignore"):
            want = fn(*host_args)
        got = cases.to_host(fn(*cases.to_dev(list(host_args))))
        for g, w in zip(cases.leaves(got), cases.leaves(want)):
            cases.assert_close(onp.asarray(g, dtype=onp.asarray(w).dtype), onp.asarray(w), tol, name)

    probes = {}

    def probe(f):
        probes[f.__name__] = f
        return f

    @probe
    def rnn():
        import rnn as m
        with b200.host_arrays():
            params = m.create_rnn_params(input_size=16, output_size=16, state_size=12, param_scale=0.1, rs=onp.random.RandomState(0))
        inputs = onp.eye(16)[rs.randint(0, 16, (7, 5))]
        compare("rnn", grad(lambda p, x: -m.rnn_log_likelihood(p, x, x)), (params, inputs))

    @probe
    def lstm():
        import lstm as m
        with b200.host_arrays():
            params = m.init_lstm_params(12, 10, 12, param_scale=0.1, rs=onp.random.RandomState(0))
        inputs = onp.eye(12)[rs.randint(0, 12, (5, 4))]
        compare("lstm", grad(lambda p, x: -m.lstm_log_likelihood(p, x, x)), (params, inputs))

    @probe
    def neural_net():
        import neural_net as m
        with b200.host_arrays():
            params = m.init_random_params(0.1, [20, 16, 10], onp.random.RandomState(0))
        X, T = rs.rand(32, 20), onp.eye(10)[rs.randint(0, 10, 32)]
        compare("neural_net", grad(lambda p, x, t: -m.log_posterior(p, x, t, 1.0)), (params, X, T))

    @probe
    def neural_net_regression():
        import neural_net_regression as m
        with b200.host_arrays():
            params = m.init_random_params(0.1, [1, 8, 8, 1], onp.random.RandomState(0))
        X, T = rs.randn(30, 1), rs.randn(30, 1)
        compare("nn regression", grad(lambda p, x, t: -m.logprob(p, x, t)), (params, X, T))

    @probe
    def convnet():
        import convnet as m
        layers = [m.conv_layer((5, 5), 3), m.maxpool_layer((2, 2)), m.tanh_layer(12), m.softmax_layer(10)]
        L2 = 1.0
        N_weights, pred, loss, frac_err = m.make_nn_funs((1, 16, 16), layers, L2)
        W = rs.randn(N_weights) * 0.1
        X, T = rs.rand(6, 1, 16, 16), onp.eye(10)[rs.randint(0, 10, 6)]
        compare("convnet", grad(loss), (W, X, T))

    @probe
    def black_box_svi():
        import black_box_svi as m
        from autograd.scipy.stats import norm

        def log_density(x, t):
            mu, log_sigma = x[:, 0], x[:, 1]
            return norm.logpdf(log_sigma, 0, 1.35) + norm.logpdf(mu, 0, np.exp(log_sigma))

        def g(params):
            objective, gradient, _ = m.black_box_variational_inference(log_density, 2, num_samples=50)
            return gradient(params, 0)
        compare("bbsvi", g, (onp.array([0.1, -0.2, -1.0, -1.2]),))

    @probe
    def bayesian_neural_net():
        import bayesian_neural_net as m
        import black_box_svi as svi
        num_weights, predictions, logprob = m.make_nn_funs([1, 6, 6, 1], 0.1, 0.01, nonlinearity=lambda x: np.exp(-(x ** 2)))
        with b200.host_arrays():
            inputs, targets = m.build_toy_dataset()

        def g(params, inputs, targets):
            obj, gradient, _ = svi.black_box_variational_inference(lambda w, t: logprob(w, inputs, targets), num_weights, 5)
            return gradient(params, 0)
        compare("bnn", g, (rs.randn(2 * num_weights) * 0.1 - 1.0, inputs, targets))

    @probe
    def variational_autoencoder():
        import variational_autoencoder as m
        with b200.host_arrays():
            gen = m.init_net_params(0.1, [3, 8, 12], onp.random.RandomState(0))
            rec = m.init_net_params(0.1, [12, 8, 6], onp.random.RandomState(1))
        data = (rs.rand(10, 12) > 0.5).astype(onp.float64)
        compare("vae", grad(lambda gp, rp, d: m.vae_lower_bound(gp, rp, d, onp.random.RandomState(2))), (gen, rec, data))

    @probe
    def generative_adversarial_net():
        import generative_adversarial_net as m
        with b200.host_arrays():
            gp = m.init_random_params(0.1, [4, 8, 6], onp.random.RandomState(0))
            dp = m.init_random_params(0.1, [6, 8, 1], onp.random.RandomState(1))
        data = rs.randn(9, 6)
        compare("gan", grad(lambda g_, d_, x: m.gan_objective(g_, d_, x, 9, 4, onp.random.RandomState(3)), 0), (gp, dp, data))
        compare("gan dsc", grad(lambda g_, d_, x: m.gan_objective(g_, d_, x, 9, 4, onp.random.RandomState(3)), 1), (gp, dp, data))

    @probe
    def ica():
        import ica as m
        data = rs.randn(4, 25)
        compare("ica", grad(lambda W, x: m.logprob(W, x) if hasattr(m, "logprob") else 0.0), (rs.randn(4, 4), data))

    @probe
    def negative_binomial_maxlike():
        import negative_binomial_maxlike as m
        data = rs.negative_binomial(5, 0.5, size=40).astype(onp.float64)
        loglike = lambda r, x: np.sum(m.negbin_loglike(r, np.sum(x) / np.sum(r + x), x))
        compare("negbin d/dr", grad(loglike), (onp.array(4.0), data))
        compare("negbin d2/dr2", grad(grad(loglike)), (onp.array(4.0), data))

    @probe
    def mixture_variational_inference():
        import mixture_variational_inference as m
        compare("mixture vi", grad(lambda x: np.sum(m.diag_gaussian_log_density(x, np.zeros(3), np.zeros(3)))), (rs.randn(6, 3),))

    @probe
    def hmm_em():
        import hmm_em as m
        with b200.host_arrays():
            onp.random.seed(0)
            params = m.initialize_hmm_parameters(3, 5)
        seq = rs.randint(0, 5, size=12)
        compare("hmm", grad(lambda p: m.log_partition_function(p, seq)), (params,))

    @probe
    def rosenbrock():
        import rosenbrock as m
        compare("rosenbrock", grad(m.rosenbrock), (onp.array([0.3, -0.7]),))

    @probe
    def gmm():
        import gmm as m
        with b200.host_arrays():
            params = m.init_gmm_params(3, 2, 0.1, onp.random.RandomState(0))
        data = rs.randn(20, 2)
        compare("gmm", grad(lambda p, d: -m.gmm_log_likelihood(p, d)), (params, data))

    @probe
    def gaussian_process():
        import gaussian_process as m
        num_params, predict, log_marginal_likelihood = m.make_gp_funs(m.rbf_covariance, num_cov_params=2)
        X, y = rs.randn(12, 1), rs.randn(12)
        compare("gp", grad(log_marginal_likelihood), (rs.randn(num_params) * 0.1, X, y))

    @probe
    def gaussian_process_predict():
        import gaussian_process as m
        num_params, predict, lml = m.make_gp_funs(m.rbf_covariance, num_cov_params=2)
        X, y, xs = rs.randn(10, 1), rs.randn(10), rs.randn(5, 1)

        def loss(params, X, y, xs):
            mean, cov = predict(params, X, y, xs)
            return np.sum(mean ** 2) + np.sum(np.diag(cov))
        compare("gp predict", grad(loss), (rs.randn(num_params) * 0.1, X, y, xs))

    @probe
    def bayesian_optimization():
        import bayesian_optimization as m
        mean, std = rs.randn(8), rs.rand(8) + 0.5
        compare("expected_new_max", grad(lambda mu, sd: np.sum(m.expected_new_max(mu, sd, 0.3))), (mean, std))
        compare("probability_of_improvement", grad(lambda mu, sd: np.sum(m.probability_of_improvement(mu, sd, 0.3))), (mean, std))

    @probe
    def define_gradient():
        import define_gradient as m            # a user-defined primitive with its own VJP (extension API on device values)

        def example_func(y):
            return np.sum(m.logsumexp(y ** 2))
        compare("define_gradient", grad(example_func), (rs.randn(10),))
        compare("define_gradient 2nd", grad(lambda y: np.sum(grad(example_func)(y) ** 2)), (rs.randn(10),))

    @probe
    def fixed_points():
        import fixed_points as m
        compare("fixed_points sqrt", grad(m.sqrt), (onp.array(2.0),), tol=1e-6)       # fixed_point primitive, scalar iterate

    @probe
    def logistic_regression():
        # the example is a script: restate its three functions on its own toy data (logistic_regression.py:6-22)
        inputs = onp.array([[0.52, 1.12, 0.77], [0.88, -1.08, 0.15], [0.52, 0.06, -1.30], [0.74, -2.49, 1.39]])
        targets = onp.array([1.0, 1.0, 0.0, 1.0])

        def training_loss(weights, inputs, targets):
            preds = 0.5 * (np.tanh(np.dot(inputs, weights)) + 1)
            return -np.sum(np.log(preds * targets + (1 - preds) * (1 - targets)))
        compare("logistic regression", grad(training_loss), (onp.array([0.1, -0.2, 0.3]), inputs, targets))

    names = args.names or list(probes)
    ok = bad = 0
    for n in names:
        try:
            probes[n]()
            print(f"PASS  {n}")
            ok += 1
        except Exception as e:   # report, keep going
            bad += 1
            msg = str(e).splitlines()[0][:160] if str(e) else ""
            print(f"FAIL  {n}: {type(e).__name__}: {msg}")
            if os.environ.get("B200AG_TRACE"):
                traceback.print_exc()
    print(f"{ok} passed, {bad} failed")
    return 0


if __name__ == "__main__":
    sys.exit(main())

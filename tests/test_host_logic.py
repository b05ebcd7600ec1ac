"""CPU suite: host logic of the backend (graph building, dtype/shape rules, layout bookkeeping, NumPy
protocol dispatch, autograd registration) exercised on tests/fakedev.FakeBackend.  Generated CUDA
sources are still compiled with NVRTC (offline) so code-generation errors fail here."""

import numpy as onp
import pytest

import cases


@pytest.mark.parametrize("name", cases.UNARY)
def test_unary_f64(name, fake_dev):
    cases.case_unary(name)


@pytest.mark.parametrize("name", ["exp", "tanh", "log", "sqrt", "negative", "abs", "floor"])
def test_unary_f32(name, fake_dev):
    import numpy as np

    cases.case_unary(name, np.float32)


@pytest.mark.parametrize("name", cases.BINARY)
def test_binary_f64(name, fake_dev):
    cases.case_binary(name)


@pytest.mark.parametrize("name", sorted(cases.OP_CASES))
def test_ops(name, fake_dev):
    cases.OP_CASES[name]()


@pytest.mark.parametrize("name", sorted(cases.GRAD_CASES))
def test_grads(name, fake_dev):
    cases.GRAD_CASES[name]()


def test_concatenate_memo_is_invisible(fake_dev):
    """Repeated concatenation of the same pieces (the four gate hstacks of an LSTM cell) reuses the first copy, yet
    behaves like independent NumPy copies: writes to one result or to an input never leak into another result."""
    import autograd.numpy as np
    import autograd_b200 as b200

    x = b200.asarray(onp.arange(6.0).reshape(2, 3))
    h = b200.asarray(onp.ones((2, 2))) * 3.0
    a = np.hstack((x, h, np.ones((2, 1))))
    n0 = fake_dev.launch_count()
    b = np.hstack((x, h, np.ones((2, 1))))
    assert fake_dev.launch_count() == n0            # memo hit: nothing launched
    assert onp.array_equal(a.get(), b.get())
    b[0, 0] = 99.0                                  # copy-on-write: `a` keeps its value
    assert a.get()[0, 0] == 0.0 and b.get()[0, 0] == 99.0
    x[0, 0] = -5.0                                  # input written in place: the memo entry is stale
    c = np.hstack((x, h, np.ones((2, 1))))
    assert c.get()[0, 0] == -5.0 and a.get()[0, 0] == 0.0
    d = np.hstack((x, h, np.ones((2, 1)) * 2.0))    # different constant piece: different result
    assert d.get()[0, -1] == 2.0 and c.get()[0, -1] == 1.0


def test_costly_small_subexpressions_are_computed_once(fake_dev, monkeypatch):
    """lazy.materialize computes a costly pending sub-expression of a smaller shape once, at its own size, instead of
    once per element of the larger space it is broadcast into (same values either way)."""
    import autograd_b200 as b200
    from autograd_b200 import lazy

    r = onp.random.RandomState(4)
    small, big = r.randn(64, 1), r.randn(64, 2048)
    want = onp.tanh(onp.exp(small) / onp.cosh(small)) * big
    sizes = []
    orig = fake_dev.run_program

    def spy(program, leaf_info, const_values, out_ptrs, n):
        sizes.append(n)
        return orig(program, leaf_info, const_values, out_ptrs, n)

    monkeypatch.setattr(fake_dev, "run_program", spy)
    for threshold, expect in ((1.0e18, [64 * 2048]), (1.0e3, [64, 64 * 2048])):
        monkeypatch.setattr(lazy, "EXPAND_WORK", threshold)
        del sizes[:]
        s, b = b200.asarray(small), b200.asarray(big)
        got = (onp.tanh(onp.exp(s) / onp.cosh(s)) * b).get()
        assert sizes == expect
        cases.assert_close(got, want, cases.TOL_ELEMENTWISE, "broadcast sub-expression")


def test_c2_kernel_shares_reciprocals_and_prefetches(fake_dev):
    """Structure of the generated C2 kernel (egrad^4 of tanh, examples/tanh.py:22-33): after value numbering the 41
    recorded divides are ~22 distinct quotients over 3 denominators, each denominator refined once (b200_rcp), the
    quotients taken with b200_div, an exact-division twin of the body for out-of-range elements, and the heavy-body
    loop that prefetches the next iteration's operand."""
    import autograd_b200 as b200
    from autograd_b200 import codegen

    from examples import workloads as w

    x = b200.asarray(onp.linspace(-7.0, 7.0, 1 << 13))
    w.tanh_nth_derivative(4)(x).get()
    prog = max(fake_dev.programs.values(), key=lambda p: len(p.instrs))
    ops = [i[0] for i in prog.instrs]
    assert len(ops) > 300 and ops.count("divide") > 30
    rep, sgn, den_of, den_uses, _pure = codegen._value_numbers(prog.instrs)
    assert sum(1 for j, r in enumerate(rep) if r == j) < 0.6 * len(ops)          # > 40 % of the body is redundant
    shared = {d: c for d, c in den_uses.items() if c >= 2}
    assert 2 <= len(shared) <= 4 and sum(shared.values()) >= 18
    from autograd_b200 import lazy

    def variant(relax):
        q = lazy.Program(prog.instrs, prog.leaves, prog.consts, prog.outputs, prog.ndim, prog.idx64, prog.vec, prog.cost,
                         prog.tables, prog.mode, prog.bx, prog.bdims, prog.share, relax)
        src, _name, _ = codegen.generate(q)
        return src, src[src.index("__device__ __forceinline__ void body("):src.index('extern "C"')]

    # (a) the bit-exact form (B200AG_EXACT=1, and every value that is still bit-reproducible against NumPy)
    src, body = variant(())
    assert body.count("b200_rcp(") == len(shared)
    assert body.count("b200_div(") == sum(shared.values())
    assert "__fma_rn(" not in body.replace("b200_", "")
    assert "body_exact(" in body and "__noinline__ Outs body_exact" in src
    kern = src[src.index('extern "C"'):]
    assert "nx0 = *reinterpret_cast" in kern and "__launch_bounds__(256, 2)" in kern
    # (b) what actually runs: everything downstream of exp() is inexact against NumPy anyway (CUDA libm vs glibc), so
    # the recorded program carries a relax plan: multiply-add pairs become FMAs, every quotient is numerator x shared
    # reciprocal, one range check per denominator, the exact twin still catches out-of-range denominators
    inexact = codegen.taint(prog.instrs)
    assert not inexact[0] and sum(inexact) > 0.95 * len(ops)            # 2*x is exact, exp(-2x) and all after it are not
    assert prog.relax == codegen.relax_plan(prog.instrs, inexact) and len(prog.relax) > 60
    src, body = variant(prog.relax)
    assert body.count("__fma_rn(") >= 30 and "b200_div(" not in body
    # the denominators are y**2, y**4, y**8, y**16 of ONE y = 1 + exp(-2x): a single Newton-refined reciprocal, the others
    # are products of it; every denominator (and factor) is range-checked once
    assert body.count("b200_rcp(") == 1 and 4 <= body.count("b200_den_ok(") <= 6
    # factors of +-2^k (the 2.0 of d/dx exp(-2x), every `2 * y * dy`) are not multiplied out: they ride on FMAs or end up
    # in the single scale applied to the result
    rep_r, scl_r, _d, _u, pure = codegen._value_numbers(prog.instrs, prog.relax)
    assert len(pure) > 100 and sum(1 for j, c in enumerate(scl_r) if abs(c) not in (1.0,)) > 50
    assert sum(1 for line in body.splitlines() if " * " in line or "__fma_rn" in line or " + " in line or " - " in line) < 175
    assert "body_exact(" in body and "__noinline__ Outs body_exact" in src


def test_exact_values_are_never_contracted(fake_dev):
    """FMA contraction and reciprocal sharing are a property of each VALUE, not of the build: chains of + - * / on
    data that is still bit-reproducible against NumPy (uploaded arrays, C4's whole forward pass) keep one IEEE
    rounding per recorded ufunc call; the same chain downstream of a float sum (reduction order differs from NumPy's)
    may be contracted."""
    import autograd_b200 as b200
    from autograd_b200 import codegen

    r = onp.random.RandomState(0)
    a, b, c = (b200.asarray(r.randn(4096)) for _ in range(3))
    (a * b + c / a).get()
    prog = max(fake_dev.programs.values(), key=lambda p: len(p.instrs))
    assert prog.relax == ()
    src, _n, _ = codegen.generate(prog)
    assert "__fma_rn(" not in src[src.index("void body("):]
    fake_dev.programs.clear()
    s = onp.sum(a)                       # a float sum: inexact against NumPy's pairwise order
    assert s._mat().buf.inexact
    y = (a * s + c)
    y.get()
    prog = max(fake_dev.programs.values(), key=lambda p: len(p.instrs))
    assert len(prog.relax) == 2          # the product and the sum that absorbs it
    src, _n, _ = codegen.generate(prog)
    assert "__fma_rn(" in src[src.index("void body("):]
    assert y._mat().buf.inexact and not (a * b)._mat().buf.inexact


def test_scatter_bodies_never_share_reciprocals(fake_dev):
    """A body with atomics (np.add.at fused with its index/value expressions) cannot be re-run by the exact twin, so
    it keeps plain divisions even when quotients share a denominator."""
    import autograd_b200 as b200
    from autograd_b200 import codegen

    r = onp.random.RandomState(3)
    t = b200.array.zeros((8, 8)).copy()
    i, j = b200.asarray(r.randint(0, 8, 50)), b200.asarray(r.randint(0, 8, 50))
    a, d = b200.asarray(r.randn(50)), b200.asarray(r.rand(50) + 1.0)
    onp.add.at(t, (i, j), a / d + (a * a) / d)
    t.get()                                                   # np.add.at is queued: reading the target launches it
    progs = [p for p in fake_dev.programs.values() if any(ins[0] == "scatter" for ins in p.instrs)]
    assert progs
    for p in progs:
        generated = codegen.generate(p)[0].split("struct Params")[1]      # everything after the fixed prelude
        assert "body_exact" not in generated and "b200_div(" not in generated and " / " in generated


def test_lstm_gate_products_are_launched_as_groups(fake_dev):
    """np.dot on device operands is queued, not launched: the four gate products of an LSTM cell that share
    hstack(input, hiddens, 1) (examples/lstm.py:38-49) reach the GPU as ONE grouped launch, and so do the G.B^T / A^T.G
    adjoints of a backward time step (numpy_vjps.py:615-638)."""
    from autograd import grad

    from examples import workloads as w

    params, X, _ = w.lstm_data(seq_len=3, batch=8, hidden=16, chars=12)
    want = grad(w.lstm_objective)(params, X, X)
    del fake_dev.gemm_groups[:]
    got = grad(w.lstm_objective)(cases.to_dev(params), cases.to_dev(X), cases.to_dev(X))
    for g, r in zip(cases.leaves(cases.to_host(got)), cases.leaves(want)):
        cases.assert_close(g, r, cases.TOL_REDUCE, "lstm grad")
    groups = list(fake_dev.gemm_groups)
    assert groups.count(4) >= 3, groups                  # forward: the four gates of every time step
    assert max(groups) >= 8, groups                      # backward: both adjoints of the four gates together
    assert sum(groups) / len(groups) > 2.5, groups


def test_deferred_products_see_the_operands_they_were_given(fake_dev):
    """A queued product is launched before anything overwrites its operands or reads its result."""
    import autograd_b200 as b200

    r = onp.random.RandomState(5)
    A, B = r.randn(6, 5), r.randn(5, 7)
    a, b = b200.asarray(A), b200.asarray(B)
    c = onp.dot(a, b)                     # queued
    a[0, 0] = 1e6                         # in-place write to an operand: the product must run first
    d = onp.dot(a, b)
    cases.assert_close(c.get(), A @ B, cases.TOL_REDUCE, "product queued before the write")
    A2 = A.copy()
    A2[0, 0] = 1e6
    cases.assert_close(d.get(), A2 @ B, cases.TOL_REDUCE, "product queued after the write")
    e = onp.dot(onp.dot(a, b), b.T)       # chained: the inner result is needed when the outer is queued
    cases.assert_close(e.get(), A2 @ B @ B.T, cases.TOL_REDUCE, "chained products")
    f = onp.dot(a, b)
    fake_dev.synchronize()                # a synchronisation leaves nothing queued
    from autograd_b200 import lazy

    assert not lazy._gemm_queue and not f._mat().buf.pending


def test_queued_products_compute_only_the_columns_that_can_still_be_read(fake_dev, monkeypatch):
    """A queued product whose full result is unreachable by the time it is launched computes only the column range its
    slices cover (the hidden-state columns of an LSTM cell's input gradient); anything else computes everything."""
    import gc

    import autograd_b200 as b200

    r = onp.random.RandomState(9)
    A, B = r.randn(12, 7), r.randn(7, 20)
    a, b = b200.asarray(A), b200.asarray(B)
    seen = []
    orig = fake_dev.gemm_grouped
    monkeypatch.setattr(fake_dev, "gemm_grouped", lambda probs: (seen.extend(p[10] for p in probs), orig(probs))[1])

    c = onp.dot(a, b)
    s1, s2 = c[:, 3:9], c[:, 6:11]
    del c
    gc.collect()
    got1, got2 = s1.get(), s2.get()
    assert seen == [8]                                     # columns 3..10 only
    cases.assert_close(got1, (A @ B)[:, 3:9], cases.TOL_REDUCE, "slice 1 of a narrowed product")
    cases.assert_close(got2, (A @ B)[:, 6:11], cases.TOL_REDUCE, "slice 2 of a narrowed product")

    del seen[:]
    c = onp.dot(a, b)
    s = c[:, 3:9]
    got = s.get()                                          # the full result is still reachable: all 20 columns
    assert seen == [20]
    cases.assert_close(c.get(), A @ B, cases.TOL_REDUCE, "full product")

    del seen[:]
    c = onp.dot(a, b)
    s = c[:, 3:9][:, 1:3]                                  # a window on a window is not tracked: everything
    del c
    gc.collect()
    cases.assert_close(s.get(), (A @ B)[:, 4:6], cases.TOL_REDUCE, "nested slice")
    assert seen == [20]

    del seen[:]
    c = onp.dot(a, b)
    s = c[2:5, :]                                          # a row range is not a column range: everything
    t = c.T
    del c
    gc.collect()
    cases.assert_close(s.get(), (A @ B)[2:5, :], cases.TOL_REDUCE, "row slice")
    cases.assert_close(t.get(), (A @ B).T, cases.TOL_REDUCE, "transposed view")
    assert seen == [20]


def test_consecutive_add_at_into_one_target_share_a_kernel(fake_dev):
    """np.add.at calls are queued per target: the four corner scatters of a bilinear-interpolation VJP become one
    generated kernel; reading the target, or writing an operand of a queued scatter, launches the queue first."""
    import autograd_b200 as b200

    r = onp.random.RandomState(2)
    n = 5000
    i0, j0 = r.randint(0, 31, n), r.randint(0, 31, n)
    g, wgt = r.randn(n), r.rand(n)
    want = onp.zeros((32, 32))
    for di, dj, ww in ((0, 0, (1 - wgt)), (0, 1, wgt), (1, 0, wgt * 0.5), (1, 1, (1 - wgt) * 0.5)):
        onp.add.at(want, (i0 + di, j0 + dj), g * ww)
    t = b200.array.zeros((32, 32)).copy()
    di0, dj0, dg, dw = (b200.asarray(a) for a in (i0, j0, g, wgt))
    fake_dev.programs.clear()
    n0 = fake_dev.launch_count()
    for di, dj, ww in ((0, 0, (1 - dw)), (0, 1, dw), (1, 0, dw * 0.5), (1, 1, (1 - dw) * 0.5)):
        onp.add.at(t, (di0 + di, dj0 + dj), dg * ww)
    assert fake_dev.launch_count() == n0                       # nothing launched yet
    got = t.get()
    assert fake_dev.launch_count() == n0 + 1                   # one kernel for the four scatters
    prog = [p for p in fake_dev.programs.values() if any(ins[0] == "scatter" for ins in p.instrs)]
    assert len(prog) == 1 and sum(1 for ins in prog[0].instrs if ins[0] == "scatter") == 4 and len(prog[0].tables) == 1
    cases.assert_close(got, want, cases.TOL_REDUCE, "grouped np.add.at")
    # a queued scatter sees the operand values of the time it was issued
    t2 = b200.array.zeros((32, 32)).copy()
    v = b200.asarray(g.copy())
    onp.add.at(t2, (di0, dj0), v * 2.0)
    v[0] = 1e6                                                 # in-place write to an operand: the queue runs first
    want2 = onp.zeros((32, 32))
    onp.add.at(want2, (i0, j0), g * 2.0)
    cases.assert_close(t2.get(), want2, cases.TOL_REDUCE, "queued np.add.at before an operand write")


def test_pooling_window_gradient_is_one_window_per_thread_kernel(fake_dev):
    """The max-pool VJP chain (grad_chooser of the outer and the inner max, numpy_vjps.py:515-530) over a large batch is ONE
    generated kernel in which a thread owns a whole 2 x 2 window: both window dims unrolled (the innermost one through
    16-byte vector accesses), offsets in 32-bit arithmetic without terms for dims an operand does not move along, and the
    plain-division twin taking the parameter block by value (no local-memory copy of the kernel parameters)."""
    import autograd.numpy as np
    from autograd import grad

    import autograd_b200 as b200
    from autograd_b200 import codegen

    r = onp.random.RandomState(3)
    X, bias, W = r.randn(320, 6, 24, 24), r.randn(1, 6, 1, 1), r.randn(864, 3)
    X[0, 0, 0, 0] = X[0, 0, 0, 1] = X[0, 0, 1, 1] = 5.0

    def pooled(x, b, w):
        p = np.max(np.max(np.reshape(x + b, (320, 6, 12, 2, 12, 2)), axis=3), axis=4)
        return np.sum(np.dot(np.reshape(p, (320, 864)), w))     # a contraction consumes the pooled values, as in the convnet

    fake_dev.programs.clear()
    g = grad(pooled)(b200.asarray(X), b200.asarray(bias), b200.asarray(W)).get()
    cases.assert_close(g, grad(pooled)(X, bias, W), cases.TOL_REDUCE, "max-pool gradient")
    window = [p for p in fake_dev.programs.values() if p.unroll and p.ndim == 6]
    assert len(window) == 1, [(p.ndim, p.unroll) for p in fake_dev.programs.values()]
    prog = window[0]
    assert prog.unroll == (3, 5) and prog.zstr is not None and not prog.idx64
    src = codegen.generate(prog)[0]
    kernel = src[src.index('extern "C" __global__'):]
    assert kernel.count("body(") == 4 and kernel.count("ldg_pack<double, 2>") >= 3
    assert kernel.count("*reinterpret_cast<Pack<double, 2>*>(p.o0") == 2          # two 16-byte stores per window
    assert "(i64)x" not in src and "const Params p," in src
    # every operand is loaded once per window: no load is emitted twice
    loads = [ln.split("=")[1].strip() for ln in kernel.splitlines() if "__ldg(" in ln or "ldg_pack<" in ln]
    assert len(loads) == len(set(loads))

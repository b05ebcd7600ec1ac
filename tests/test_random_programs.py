"""Randomised differential test of the lazy host logic (no GPU): seeded random "user programs" - chains of elementwise
ops with broadcasting, reshapes / ravels / transposes / slices / rolls, reductions, gathers, products, and IN-PLACE writes
through arrays and through their views - run once on NumPy and once on DeviceArrays over tests/fakedev.FakeBackend; every
array alive at the end (and at random points in between) must agree.

The emulation executes the same recorded `lazy.Program`s the CUDA backend would, operand by operand with the strides,
shifts and iteration space `lazy.materialize` chose, so this exercises what the parity cases cannot enumerate: the
interplay of pending expressions with later writes (Buffer.readers, Node.consumers), of symbolic reshapes with NumPy's
view semantics (lazy._AliasGroup), of reshape nodes with the iteration-space planning (lazy._plan_reshapes), of queued
products / scatters with whatever reads or overwrites their operands.  NumPy's eager evaluation is the oracle.
"""

import numpy as onp
import pytest

import cases

ROWS, COLS = 64, 96            # 6144 elements: above the 4096 threshold of pending reshape nodes


class _Prog:
    """One random program, generated against NumPy arrays and replayed verbatim on device arrays."""

    def __init__(self, seed):
        self.r = onp.random.RandomState(seed)
        self.steps = []          # (fn(env) -> None, description)

    # ---- program construction: every step is a closure over NAMES in an environment dict
    def build(self, n_steps):
        r = self.r
        names2d, names1d = ["a", "b"], []
        counter = [0]

        def fresh(prefix):
            counter[0] += 1
            return f"{prefix}{counter[0]}"

        def pick(seq):
            return seq[r.randint(len(seq))]

        for _ in range(n_steps):
            kind = r.randint(14)
            if kind == 0:                                            # elementwise binary, same shape
                x, y, out = pick(names2d), pick(names2d), fresh("e")
                op = pick(["add", "subtract", "multiply", "maximum"])
                self.steps.append((lambda env, x=x, y=y, out=out, op=op: env.__setitem__(out, getattr(onp, op)(env[x], env[y])),
                                   f"{out} = {op}({x}, {y})"))
                names2d.append(out)
            elif kind == 1:                                          # broadcast against the row / the column
                x, out, which = pick(names2d), fresh("e"), pick(["row", "col"])
                self.steps.append((lambda env, x=x, out=out, which=which: env.__setitem__(out, env[x] * 0.5 + env[which]),
                                   f"{out} = {x} * 0.5 + {which}"))
                names2d.append(out)
            elif kind == 2:                                          # ravel / reshape(-1) of whatever the operand is
                x, out = pick(names2d), fresh("f")
                self.steps.append((lambda env, x=x, out=out: env.__setitem__(out, env[x].ravel()), f"{out} = {x}.ravel()"))
                names1d.append(out)
            elif kind == 3 and names1d:                              # back to 2-d, sometimes another 2-d shape
                x, out = pick(names1d), fresh("e")
                shape = pick([(ROWS, COLS), (ROWS, COLS), (COLS, ROWS)])
                if shape == (ROWS, COLS):
                    names2d.append(out)
                self.steps.append((lambda env, x=x, out=out, shape=shape: env.__setitem__(out, env[x].reshape(shape)),
                                   f"{out} = {x}.reshape{shape}"))
            elif kind == 4 and names1d:                              # 1-d elementwise
                x, y, out = pick(names1d), pick(names1d), fresh("f")
                self.steps.append((lambda env, x=x, y=y, out=out: env.__setitem__(out, env[x] - 2.0 * env[y]),
                                   f"{out} = {x} - 2 * {y}"))
                names1d.append(out)
            elif kind == 5:                                          # in-place write into a 2-d array (scalar / row block)
                x = pick(names2d)
                i, j, v = r.randint(ROWS), r.randint(COLS), float(r.randint(-50, 50))
                if r.randint(2):
                    self.steps.append((lambda env, x=x, i=i, j=j, v=v: env[x].__setitem__((i, j), v), f"{x}[{i}, {j}] = {v}"))
                else:
                    self.steps.append((lambda env, x=x, i=i, v=v: env[x].__setitem__(slice(i, i + 3), v), f"{x}[{i}:{i + 3}] = {v}"))
            elif kind == 6 and names1d:                              # in-place write through a flattened handle
                x = pick(names1d)
                i, v = r.randint(ROWS * COLS - 8), float(r.randint(-50, 50))
                self.steps.append((lambda env, x=x, i=i, v=v: env[x].__setitem__(slice(i, i + 5), v), f"{x}[{i}:{i + 5}] = {v}"))
            elif kind == 7:                                          # a slice view, later possibly written through
                x, out = pick(names2d), fresh("s")
                i = r.randint(ROWS - 8)
                self.steps.append((lambda env, x=x, out=out, i=i: env.__setitem__(out, env[x][i:i + 8]), f"{out} = {x}[{i}:{i + 8}]"))
                self.steps.append((lambda env, out=out: env[out].__setitem__((slice(2, 4), slice(None)), -3.0), f"{out}[2:4] = -3"))
            elif kind == 8:                                          # roll + stencil-like combination
                x, out = pick(names2d), fresh("e")
                sh, ax = pick([1, -1, 5]), pick([0, 1])
                self.steps.append((lambda env, x=x, out=out, sh=sh, ax=ax: env.__setitem__(out, (env[x] + onp.roll(env[x], sh, axis=ax)) / 4.0),
                                   f"{out} = ({x} + roll({x}, {sh}, {ax})) / 4"))
                names2d.append(out)
            elif kind == 9:                                          # reduction over an axis, broadcast back
                x, out, ax = pick(names2d), fresh("e"), pick([0, 1])
                self.steps.append((lambda env, x=x, out=out, ax=ax: env.__setitem__(out, env[x] - onp.sum(env[x], axis=ax, keepdims=True) * 0.001),
                                   f"{out} = {x} - 0.001 * sum({x}, {ax}, keepdims)"))
                names2d.append(out)
            elif kind == 10 and names1d:                             # gather through indices computed from a flattened expression
                x, t, out = pick(names1d), pick(names2d), fresh("f")
                self.steps.append((lambda env, x=x, t=t, out=out: env.__setitem__(
                    out, env[t][onp.mod(onp.floor(env[x] * 7.0).astype(int), ROWS), onp.mod(onp.floor(env[x] * 3.0).astype(int), COLS)]),
                                   f"{out} = {t}[floor(7 {x}) % R, floor(3 {x}) % C]"))
                names1d.append(out)
            elif kind == 11:                                         # a product (queued on the device)
                x, out = pick(names2d), fresh("p")
                self.steps.append((lambda env, x=x, out=out: env.__setitem__(out, onp.dot(env[x], env["w"])), f"{out} = dot({x}, w)"))
            elif kind == 12 and names1d:                             # np.add.at into a flattened handle
                x = pick(names1d)
                self.steps.append((lambda env, x=x: onp.add.at(env[x], env["idx"], 1.5), f"add.at({x}, idx, 1.5)"))
            elif kind == 13:                                         # transpose twice with work in between
                x, out = pick(names2d), fresh("e")
                self.steps.append((lambda env, x=x, out=out: env.__setitem__(out, (env[x].T * 2.0).T + 1.0), f"{out} = ({x}.T * 2).T + 1"))
                names2d.append(out)
        return self

    def run(self, make):
        r = onp.random.RandomState(12345)
        env = {"a": make(r.randn(ROWS, COLS)), "b": make(r.randn(ROWS, COLS)), "row": make(r.randn(COLS)),
               "col": make(r.randn(ROWS, 1)), "w": make(r.randn(COLS, 5)),
               "idx": make(r.randint(0, ROWS * COLS, 50))}
        for fn, _desc in self.steps:
            fn(env)
        return env


@pytest.mark.parametrize("seed", range(60))
def test_random_program_matches_numpy(seed, fake_dev):
    import autograd_b200 as b200

    prog = _Prog(seed).build(18)
    want = prog.run(lambda h: h.copy())
    got = prog.run(lambda h: b200.asarray(h))
    listing = "\n".join(d for _f, d in prog.steps)
    for name in sorted(want):
        g = cases.to_host(got[name])
        w = want[name]
        tol = cases.TOL_REDUCE if name.startswith("p") else cases.TOL_ELEMENTWISE
        cases.assert_close(g, w, tol, f"seed {seed}, array {name} after\n{listing}\n")


def _random_loss(seed):
    """A random differentiable scalar function of (x, row, col, w) built from the op families of the BASELINE configs:
    broadcasting arithmetic, tanh / exp, reshape + short-axis max (pooling), roll stencils, flattened gathers with
    computed indices (advection), products, logsumexp-style reductions."""
    import autograd.numpy as np

    r = onp.random.RandomState(seed)
    layers = [int(k) for k in r.randint(0, 9, size=6)]
    shift = int(r.choice([1, -1, 3]))

    def loss(x, row, col, w):
        h = x
        for k in layers:
            if k == 0:
                h = np.tanh(h * 0.5 + row)
            elif k == 1:
                h = h * col - np.roll(h, shift, axis=0) * 0.25
            elif k == 2:                                   # 2 x 2 max pool and broadcast back (ties are measure-zero here)
                p = np.max(np.max(np.reshape(h, (ROWS // 2, 2, COLS // 2, 2)), axis=1), axis=2)
                h = h + np.reshape(np.repeat(np.repeat(p, 2, axis=0), 2, axis=1), (ROWS, COLS)) * 0.1
            elif k == 3:                                   # flattened gather through indices computed from the values
                f = (h * 3.0 + row).ravel()
                ix = np.mod(np.floor(f).astype(int), ROWS)
                iy = np.mod(np.floor(f * 0.5).astype(int), COLS)
                frac = f - np.floor(f)
                h = np.reshape(frac * x[ix, iy] + (1.0 - frac) * h.ravel(), (ROWS, COLS))
            elif k == 4:
                h = h - np.sum(h, axis=1, keepdims=True) / COLS
            elif k == 5:
                h = np.exp(-h * h * 0.1) + col
            elif k == 6:
                h = h + np.dot(np.tanh(np.dot(h, w)), w.T) * 0.05
            elif k == 7:
                h = (h + np.roll(h, 1, axis=1) + np.roll(h, -1, axis=1) + np.roll(h, 1, axis=0)) / 4.0
            else:
                h = np.reshape(np.reshape(h, (-1,)) * 1.5, (ROWS, COLS)).T.T + row
        return np.sum(h * h) * 1e-3 + np.sum(np.log(np.sum(np.exp(h * 0.1), axis=1)))

    return loss


@pytest.mark.parametrize("seed", list(range(24)) + [1005])
def test_random_gradient_matches_reference(seed, fake_dev):
    """grad of a random composition, on device arrays, against reference autograd on NumPy (same tape rules)."""
    from autograd import grad

    import autograd_b200 as b200

    r = onp.random.RandomState(100 + seed)
    args = (r.randn(ROWS, COLS), r.randn(COLS), r.randn(ROWS, 1), r.randn(COLS, 5) * 0.3)
    loss = _random_loss(seed)
    for argnum in (0, 3):
        want = grad(loss, argnum)(*args)
        got = grad(loss, argnum)(*[b200.asarray(a) for a in args])
        cases.assert_close(cases.to_host(got), want, cases.TOL_REDUCE, f"seed {seed}, gradient wrt argument {argnum}")


# ---------------------------------------------------------------------------------------------------------------------
# Second family: library calls around the deferred-product queue - products whose results are sliced, dropped, fed to
# other products; concatenate / where / cumsum / sort / einsum / take / clip / dtype round trips / stack / tile / batched
# matmul - with in-place row writes in between.
import gc


def _build_library_program(seed, n_steps):
    r = onp.random.RandomState(seed)
    steps = []
    mats = ["A", "B"]           # (48, 40) matrices
    prods = []                  # (48, k) products
    vecs = []
    cnt = [0]
    def fresh(p):
        cnt[0] += 1; return f"{p}{cnt[0]}"
    def pick(s): return s[r.randint(len(s))]
    for _ in range(n_steps):
        k = r.randint(12)
        if k == 0:
            x, out = pick(mats), fresh("P")
            steps.append((lambda e, x=x, out=out: e.__setitem__(out, onp.dot(e[x], e["W"])), f"{out} = dot({x}, W)"))
            prods.append(out)
        elif k == 1 and prods:
            x, out = pick(prods), fresh("S")
            lo = r.randint(0, 20); hi = lo + r.randint(1, 12)
            steps.append((lambda e, x=x, out=out, lo=lo, hi=hi: e.__setitem__(out, e[x][:, lo:hi]), f"{out} = {x}[:, {lo}:{hi}]"))
            if r.randint(2):
                steps.append((lambda e, x=x: e.__setitem__(x, None), f"del {x}"))
                prods.remove(x)
        elif k == 2 and prods:
            x, out = pick(prods), fresh("M")
            steps.append((lambda e, x=x, out=out: e.__setitem__(out, onp.tanh(e[x]) @ e["W"].T), f"{out} = tanh({x}) @ W.T"))
            mats.append(out)
        elif k == 3:
            x, y, out = pick(mats), pick(mats), fresh("M")
            steps.append((lambda e, x=x, y=y, out=out: e.__setitem__(out, onp.concatenate([e[x][:24], e[y][24:]], axis=0)), f"{out} = concat({x}[:24], {y}[24:])"))
            mats.append(out)
        elif k == 4:
            x, out = pick(mats), fresh("M")
            steps.append((lambda e, x=x, out=out: e.__setitem__(out, onp.where(e[x] > 0, e[x], e[x] * 0.1)), f"{out} = leaky({x})"))
            mats.append(out)
        elif k == 5:
            x = pick(mats)
            i = r.randint(48)
            steps.append((lambda e, x=x, i=i: e[x].__setitem__(i, 0.5), f"{x}[{i}] = 0.5"))
        elif k == 6:
            x, out = pick(mats), fresh("M")
            steps.append((lambda e, x=x, out=out: e.__setitem__(out, onp.cumsum(e[x], axis=1) * 0.01 + onp.sort(e[x], axis=0)), f"{out} = cumsum({x}) + sort({x})"))
            mats.append(out)
        elif k == 7:
            x, out = pick(mats), fresh("M")
            steps.append((lambda e, x=x, out=out: e.__setitem__(out, onp.einsum("ij,kj->ik", e[x], e[x])[:, :40] * 0.01), f"{out} = einsum({x},{x})[:, :40]"))
            mats.append(out)
        elif k == 8:
            x, out = pick(mats), fresh("M")
            steps.append((lambda e, x=x, out=out: e.__setitem__(out, onp.clip(e[x], -0.5, 0.5) + onp.take(e[x], e["ti"], axis=0)), f"{out} = clip({x}) + take({x})"))
            mats.append(out)
        elif k == 9:
            x, out = pick(mats), fresh("M")
            steps.append((lambda e, x=x, out=out: e.__setitem__(out, (e[x].astype(onp.float32) * 2).astype(onp.float64) + e[x]), f"{out} = f32 roundtrip {x}"))
            mats.append(out)
        elif k == 10:
            x, out = pick(mats), fresh("M")
            steps.append((lambda e, x=x, out=out: e.__setitem__(out, onp.stack([e[x], e[x] * 2]).sum(axis=0) + onp.tile(e[x][:1], (48, 1))), f"{out} = stack/tile {x}"))
            mats.append(out)
        elif k == 11:
            x, out = pick(mats), fresh("M")
            steps.append((lambda e, x=x, out=out: e.__setitem__(out, onp.matmul(e[x][None] * onp.ones((3, 1, 1)), e["W"][None])[1] @ e["W"].T), f"{out} = batched matmul {x}"))
            mats.append(out)
    return steps

def _run_library_program(steps, make):
    r = onp.random.RandomState(777)
    e = {"A": make(r.randn(48, 40)), "B": make(r.randn(48, 40)), "W": make(r.randn(40, 24)), "ti": make(r.randint(0, 48, 48))}
    for fn, _ in steps:
        fn(e)
        gc.collect()
    return e



@pytest.mark.parametrize("seed", range(30))
def test_random_library_program_matches_numpy(seed, fake_dev):
    import autograd_b200 as b200

    steps = _build_library_program(seed, 20)
    want = _run_library_program(steps, lambda h: h.copy())
    got = _run_library_program(steps, lambda h: b200.asarray(h))
    listing = "\n".join(d for _f, d in steps)
    for name in sorted(want):
        if want[name] is not None:
            cases.assert_close(cases.to_host(got[name]), want[name], 1e-9, f"seed {seed}, array {name} after\n{listing}\n")

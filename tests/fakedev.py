"""TEST INFRASTRUCTURE ONLY — a host emulation of the device backend interface.

The CPU test-suite (``-m "not gpu"``) uses this to exercise the HOST logic of
autograd_b200 (lazy graph construction, dtype resolution, layout classification,
stride/shift bookkeeping, NumPy-protocol dispatch, autograd registration) in a
container without a GPU.  It is never imported by the package: the product path
(`autograd_b200.backend.CudaBackend`) has no CPU fallback.  "Device memory" here
is a dict of numpy byte buffers addressed by fake pointers; fused programs are
evaluated by interpreting the same `lazy.Program` the CUDA code generator
consumes, and — when libb200ag.so is present — the generated CUDA source is also
compiled with NVRTC (which needs no GPU) so syntax errors in generated kernels
surface on CPU.
"""

from __future__ import annotations

import bisect
import os

import numpy as np

_OPS1 = {
    "negative": np.negative, "positive": np.positive, "square": np.square, "reciprocal": np.reciprocal,
    "abs": np.absolute, "sign": np.sign, "exp": np.exp, "exp2": np.exp2, "expm1": np.expm1, "log": np.log,
    "log2": np.log2, "log10": np.log10, "log1p": np.log1p, "sqrt": np.sqrt, "cbrt": np.cbrt, "sin": np.sin,
    "cos": np.cos, "tan": np.tan, "arcsin": np.arcsin, "arccos": np.arccos, "arctan": np.arctan, "sinh": np.sinh,
    "cosh": np.cosh, "tanh": np.tanh, "arcsinh": np.arcsinh, "arccosh": np.arccosh, "arctanh": np.arctanh,
    "floor": np.floor, "ceil": np.ceil, "trunc": np.trunc, "rint": np.rint, "logical_not": np.logical_not,
    "isfinite": np.isfinite, "isnan": np.isnan, "isinf": np.isinf, "conjugate": np.conjugate,
    "deg2rad": np.deg2rad, "rad2deg": np.rad2deg,
}
_OPS2 = {
    "add": np.add, "subtract": np.subtract, "multiply": np.multiply, "divide": np.true_divide, "power": np.power,
    "remainder": np.remainder, "floor_divide": np.floor_divide, "maximum": np.maximum, "minimum": np.minimum,
    "fmax": np.fmax, "fmin": np.fmin, "equal": np.equal, "not_equal": np.not_equal, "less": np.less,
    "less_equal": np.less_equal, "greater": np.greater, "greater_equal": np.greater_equal,
    "logical_and": np.logical_and, "logical_or": np.logical_or, "logical_xor": np.logical_xor,
    "arctan2": np.arctan2, "hypot": np.hypot, "logaddexp": np.logaddexp, "copysign": np.copysign,
    "logaddexp2": np.logaddexp2,
}


import scipy.special as _sps

_OPS1.update({
    "sp_erf": _sps.erf, "sp_erfc": _sps.erfc, "sp_erfinv": _sps.erfinv, "sp_erfcinv": _sps.erfcinv, "sp_erfcx": _sps.erfcx,
    "sp_gammaln": _sps.gammaln, "sp_gamma": _sps.gamma, "sp_digamma": _sps.psi, "sp_ndtr": _sps.ndtr,
    "sp_ndtri": _sps.ndtri, "sp_log_ndtr": _sps.log_ndtr, "sp_expit": _sps.expit, "sp_logit": _sps.logit,
    "sp_j0": _sps.j0, "sp_j1": _sps.j1, "sp_y0": _sps.y0, "sp_y1": _sps.y1, "sp_i0": _sps.i0, "sp_i1": _sps.i1,
})


_OPS2.update({
    "sp_jn": lambda n, x: _sps.jv(n, x), "sp_yn": lambda n, x: _sps.yn(n.astype(np.int64), x),
    "sp_iv": lambda n, x: _sps.iv(n, x), "sp_ive": lambda n, x: _sps.ive(n, x),
    "sp_polygamma": lambda n, x: _sps.polygamma(n.astype(np.int64), x),
    "sp_gammainc": _sps.gammainc, "sp_gammaincc": _sps.gammaincc,
})
_OPS3 = {"sp_betainc": _sps.betainc}


class FakeBackend:
    name = "fake-host"
    sm_count = 148
    total_mem = 180 << 30

    _instances = 0

    def __init__(self, compile_generated: bool = True):
        self._bases = []     # sorted base addresses
        self._bufs = {}      # base -> np.uint8 array
        # every instance hands out addresses from its own range: a buffer of an EARLIER instance that the garbage
        # collector frees late (cycles through weak references) must not alias a live allocation of this one
        FakeBackend._instances += 1
        self._next = (FakeBackend._instances << 40) + (1 << 20)
        self._launches = 0
        self._compiled = set()
        self.programs = {}   # key -> Program (for inspection by tests)
        self.gemm_groups = []  # problems per grouped-GEMM launch (for inspection by tests)
        self.compile_generated = compile_generated and os.path.exists(
            os.path.join(os.path.dirname(__file__), "..", "autograd_b200", "libb200ag.so"))
        self.in_use = 0
        self.peak = 0
        self._host = {}

    # ------------------------------------------------------------ memory
    def malloc(self, nbytes):
        nbytes = max(int(nbytes), 1)
        base = self._next
        self._next += (nbytes + 511) // 512 * 512 + 512
        bisect.insort(self._bases, base)
        self._bufs[base] = np.zeros(nbytes, dtype=np.uint8)
        self.in_use += nbytes
        self.peak = max(self.peak, self.in_use)
        return base

    def malloc_upload(self, nbytes):
        return self.malloc(nbytes)

    def free(self, ptr):
        buf = self._bufs.pop(ptr, None)
        if buf is not None:
            self.in_use -= buf.nbytes
            self._bases.remove(ptr)

    def _resolve(self, ptr):
        i = bisect.bisect_right(self._bases, ptr) - 1
        if i < 0:
            raise RuntimeError(f"fake device: wild pointer {ptr:#x}")
        base = self._bases[i]
        buf = self._bufs[base]
        if ptr - base > buf.nbytes:
            raise RuntimeError(f"fake device: pointer {ptr:#x} past the end of its allocation")
        return buf, ptr - base

    def _elems(self, ptr, dtype):
        """(element view of the whole allocation, element index of ptr)."""
        buf, off = self._resolve(ptr)
        dtype = np.dtype(dtype)
        assert off % dtype.itemsize == 0
        n = buf.nbytes // dtype.itemsize
        return buf[: n * dtype.itemsize].view(dtype), off // dtype.itemsize

    def _strided(self, ptr, dtype, shape, strides):
        el, e0 = self._elems(ptr, dtype)
        idx = np.full(tuple(shape), e0, dtype=np.int64)
        for d, (s, st) in enumerate(zip(shape, strides)):
            shp = [1] * len(shape)
            shp[d] = s
            idx = idx + (np.arange(s, dtype=np.int64) * st).reshape(shp)
        if idx.size and (idx.min() < 0 or idx.max() >= el.size):
            raise RuntimeError("fake device: strided access out of bounds")
        return el, idx

    def h2d(self, ptr, host):
        host = np.asarray(host, order="C")
        buf, off = self._resolve(ptr)
        buf[off: off + host.nbytes] = host.reshape(-1).view(np.uint8)

    def d2h(self, host, ptr):
        buf, off = self._resolve(ptr)
        host.reshape(-1).view(np.uint8)[:] = buf[off: off + host.nbytes]

    def d2d(self, dst, src, nbytes):
        db, do = self._resolve(dst)
        sb, so = self._resolve(src)
        db[do: do + nbytes] = sb[so: so + nbytes]

    def memset(self, ptr, byte, nbytes):
        buf, off = self._resolve(ptr)
        buf[off: off + nbytes] = byte

    def host_alloc(self, nbytes):
        arr = np.zeros(max(int(nbytes), 1), dtype=np.uint8)   # "pinned" memory must be REAL host memory
        self._host[arr.ctypes.data] = arr
        return arr.ctypes.data

    def host_free(self, ptr):
        self._host.pop(ptr, None)

    def mem_stats(self):
        s = {"in_use": self.in_use, "cached": 0, "peak_in_use": self.peak}
        self.peak = self.in_use
        return s

    def empty_cache(self):
        pass

    # ------------------------------------------------------------ runtime
    def synchronize(self):
        from autograd_b200 import lazy

        lazy.flush_deferred()

    def launch_count(self):
        return self._launches

    def stream_handle(self):
        return 0

    def event_create(self):
        return [0.0]

    def event_record(self, ev):
        import time

        ev[0] = time.perf_counter()

    def event_synchronize(self, ev):
        pass

    def event_elapsed_ms(self, a, b):
        return (b[0] - a[0]) * 1e3

    def event_destroy(self, ev):
        pass

    def stream_create(self):
        return 1

    def stream_destroy(self, s):
        pass

    def stream_synchronize(self, s):
        pass

    def event_record_on(self, ev, s):
        self.event_record(ev)

    def stream_wait_event(self, s, ev):
        pass

    def h2d_on(self, ptr, host, s):
        self.h2d(ptr, host)

    def d2h_on(self, host, ptr, s):
        self.d2h(host, ptr)

    def graph_begin(self):
        raise RuntimeError("fake device: CUDA graphs need a GPU")

    # ------------------------------------------------------------ fused programs
    def run_program(self, program, leaf_info, const_values, out_ptrs, n):
        self._launches += 1
        self.programs[program.key] = program
        if self.compile_generated and program.key not in self._compiled:
            self._nvrtc_check(program)
            self._compiled.add(program.key)
        if n == 0:
            return
        leaf_ptrs, dims = leaf_info[0], leaf_info[1]
        tables = leaf_info[2] if len(leaf_info) > 2 else ()
        i = np.arange(n, dtype=np.int64)
        coords = None
        if program.ndim:
            coords, rem = [None] * program.ndim, i
            for d in range(program.ndim - 1, 0, -1):
                coords[d] = rem % dims[d]
                rem = rem // dims[d]
            coords[0] = rem
        leaves = []
        for (ch, cls, has_shift, _flag), (ptr, strides, shifts) in zip(program.leaves, leaf_ptrs):
            el, e0 = self._elems(ptr, np.dtype(ch))
            if cls == "S":
                leaves.append(el[e0])
            elif cls == "C":
                leaves.append(el[e0: e0 + n])
            else:
                off = np.full(n, e0, dtype=np.int64)
                for d in range(program.ndim):
                    x = coords[d]
                    if has_shift:
                        x = (x - shifts[d]) % dims[d]
                    off = off + x * strides[d]
                assert off.min() >= 0 and off.max() < el.size, "general-layout leaf out of bounds"
                leaves.append(el[off])
        vals = []
        with np.errstate(all="ignore"):
            for op, in_chars, out_ch, refs in program.instrs:
                args = []
                for (kind, v), ich in zip(refs, in_chars):
                    dt = np.dtype(ich)
                    if kind == "t":
                        args.append(v)
                        continue
                    if kind == "n":
                        a = vals[v]
                    elif kind == "l":
                        a = leaves[v]
                    elif kind == "c":
                        a = const_values[v]
                    else:
                        a = v
                    args.append(np.asarray(a).astype(dt))
                out_dt = np.dtype(out_ch)
                if op == "gather":
                    tptr, tdims = tables[args[0]]
                    el, e0 = self._elems(tptr, out_dt)
                    lin = np.zeros(n, dtype=np.int64)
                    for d, size in enumerate(tdims):
                        g = np.broadcast_to(args[1 + d], (n,)).astype(np.int64)
                        g = np.where(g < 0, g + size, g)
                        assert (g >= 0).all() and (g < size).all(), "gather index out of bounds"
                        lin = lin * size + g
                    vals.append(el[e0 + lin])
                    continue
                if op == "scatter":
                    tptr, tdims = tables[args[0]]
                    el, e0 = self._elems(tptr, out_dt)
                    lin = np.zeros(n, dtype=np.int64)
                    for d, size in enumerate(tdims):
                        g = np.broadcast_to(args[1 + d], (n,)).astype(np.int64)
                        g = np.where(g < 0, g + size, g)
                        assert (g >= 0).all() and (g < size).all(), "scatter index out of bounds"
                        lin = lin * size + g
                    np.add.at(el, e0 + lin, np.broadcast_to(args[1 + len(tdims)], (n,)).astype(out_dt))
                    vals.append(np.zeros((), out_dt))
                    continue
                if op == "where":
                    r = np.where(args[0], args[1], args[2])
                elif op in ("cast", "full"):
                    r = args[0]
                elif op in _OPS1:
                    r = _OPS1[op](args[0])
                elif op in _OPS2:
                    r = _OPS2[op](args[0], args[1])
                elif op in _OPS3:
                    r = _OPS3[op](args[0], args[1], args[2])
                else:
                    raise NotImplementedError(op)
                vals.append(np.asarray(r).astype(out_dt))
        self._check_value_numbers(program, vals)
        for (j, ch), optr in zip(program.outputs, out_ptrs):
            el, e0 = self._elems(optr, np.dtype(ch))
            el[e0: e0 + n] = np.broadcast_to(vals[j], (n,))

    @staticmethod
    def _check_value_numbers(program, vals):
        """The CUDA generator emits instruction j as ``[-]v{rep[j]}`` when codegen._value_numbers says so.  This
        interpreter evaluates every recorded instruction on its own, so it can check that claim on every program of
        the CPU suite: the aliased value must be the SAME BITS (signed zeros included; NaNs compare equal)."""
        from autograd_b200 import codegen

        rep, scl, _den_of, _den_uses, _pure = codegen._value_numbers(program.instrs, getattr(program, "relax", ()))
        for j, r in enumerate(rep):
            if r == j:
                continue
            a = np.asarray(vals[j])
            b = np.asarray(vals[r])
            ratio = scl[j] / scl[r]       # +-1 on exact values; +-2^k on values the generator may rescale (exact too)
            if ratio != 1.0:
                b = b * b.dtype.type(ratio) if b.dtype.kind == "f" else -b
            a, b = np.broadcast_arrays(a, b)
            if a.dtype.kind == "f":
                nan = np.isnan(a)
                assert np.array_equal(nan, np.isnan(b)), f"value numbering: NaN pattern of instr {j} vs {r} differs"
                ai = np.where(nan, 0, a).view(f"i{a.dtype.itemsize}")
                bi = np.where(nan, 0, b).view(f"i{b.dtype.itemsize}")
                assert np.array_equal(ai, bi), f"value numbering: instr {j} ({program.instrs[j][0]}) is not bit-identical to {ratio} * v{r}"
            else:
                assert np.array_equal(a, b), f"value numbering: instr {j} differs from v{r}"

    def _nvrtc_check(self, program):
        import ctypes as C

        from autograd_b200 import _lib, codegen

        source, name, _ = codegen.generate(program)
        p, sz = C.c_void_p(), C.c_size_t()
        rc = _lib.lib.b200ag_rtc_compile(source.encode(), (name + ".cu").encode(), C.byref(p), C.byref(sz))
        if rc != 0:
            raise RuntimeError("generated kernel failed to compile:\n" + _lib.lib.b200ag_last_error().decode()
                               + "\n---- source ----\n" + source)
        _lib.lib.b200ag_rtc_free(p)

    # ------------------------------------------------------------ static kernels
    def copy_strided(self, dst, dst_dt, dst_strides, src, src_dt, src_strides, shape):
        self._launches += 1
        if int(np.prod(shape)) == 0:
            return
        sel, sidx = self._strided(src, src_dt, shape, src_strides)
        del_, didx = self._strided(dst, dst_dt, shape, dst_strides)
        del_[didx] = sel[sidx].astype(np.dtype(dst_dt))

    def add_strided(self, dst, dt, dst_strides, src, src_strides, shape):
        self._launches += 1
        if int(np.prod(shape)) == 0:
            return
        sel, sidx = self._strided(src, dt, shape, src_strides)
        del_, didx = self._strided(dst, dt, shape, dst_strides)
        np.add.at(del_, didx, sel[sidx])

    def concat(self, dst, dt, src_ptrs, inner, outer):
        self._launches += 1
        total = int(sum(inner))
        del_, d0 = self._elems(dst, dt)
        d = del_[d0: d0 + outer * total].reshape(outer, total)
        pos = 0
        for ptr, k in zip(src_ptrs, inner):
            sel, s0 = self._elems(ptr, dt)
            d[:, pos: pos + k] = sel[s0: s0 + outer * k].reshape(outer, k)
            pos += k

    def concat_mixed(self, dst, dt, pieces, inner, outer):
        self._launches += 1
        total = int(sum(inner))
        del_, d0 = self._elems(dst, dt)
        d = del_[d0: d0 + outer * total].reshape(outer, total)
        pos = 0
        for (ptr, info), k in zip(pieces, inner):
            if ptr is None:
                d[:, pos: pos + k] = info
            else:
                sel, s0 = self._elems(ptr, info)
                d[:, pos: pos + k] = sel[s0: s0 + outer * k].reshape(outer, k)
            pos += k

    def _view3(self, ptr, dt, outer, red, inner):
        el, e0 = self._elems(ptr, dt)
        return el[e0: e0 + outer * red * inner].reshape(outer, red, inner)

    def reduce(self, dst, src, dt, op, outer, red, inner):
        self._launches += 1
        x = self._view3(src, dt, outer, red, inner)
        fn = {"sum": np.sum, "max": np.max, "min": np.min, "prod": np.prod}[op]
        r = fn(x, axis=1)
        el, e0 = self._elems(dst, dt)
        el[e0: e0 + outer * inner] = r.reshape(-1).astype(np.dtype(dt))

    def argreduce(self, dst, src, dt, op, outer, red, inner):
        self._launches += 1
        x = self._view3(src, dt, outer, red, inner)
        r = (np.argmax if op == "max" else np.argmin)(x, axis=1)
        el, e0 = self._elems(dst, np.int64)
        el[e0: e0 + outer * inner] = r.reshape(-1)

    def logsumexp(self, dst, src, dt, outer, red, inner):
        from scipy.special import logsumexp

        self._launches += 1
        x = self._view3(src, dt, outer, red, inner)
        el, e0 = self._elems(dst, dt)
        el[e0: e0 + outer * inner] = logsumexp(x, axis=1).reshape(-1)

    def gemm(self, c, c_dt, c_str, a, a_dt, a_str, b, b_dt, b_str, M, N, K, batch=1):
        self._launches += 1
        ael, aidx = self._strided(a, a_dt, (batch, M, K), (a_str[2], a_str[0], a_str[1]))
        bel, bidx = self._strided(b, b_dt, (batch, K, N), (b_str[2], b_str[0], b_str[1]))
        cel, cidx = self._strided(c, c_dt, (batch, M, N), (c_str[2], c_str[0], c_str[1]))
        r = np.matmul(ael[aidx].astype(np.float64), bel[bidx].astype(np.float64))
        cel[cidx] = r.astype(np.dtype(c_dt))

    def gemm_grouped(self, problems):
        self._launches += 1
        self.gemm_groups.append(len(problems))
        for (c, c_dt, c_str, a, a_dt, a_str, b, b_dt, b_str, M, N, K) in problems:
            self.gemm(c, c_dt, (c_str[0], c_str[1], 0), a, a_dt, (a_str[0], a_str[1], 0), b, b_dt, (b_str[0], b_str[1], 0),
                      M, N, K, 1)
            self._launches -= 1

    def _rows(self, idx_ptrs, dims, n_index):
        row = np.zeros(n_index, dtype=np.int64)
        for p, d in zip(idx_ptrs, dims):
            el, e0 = self._elems(p, np.int64)
            v = el[e0: e0 + n_index].copy()
            v[v < 0] += d
            assert (v >= 0).all() and (v < d).all(), "index out of bounds"
            row = row * d + v
        return row

    def gather(self, dst, src, dt, idx_ptrs, dims, n_index, inner):
        self._launches += 1
        row = self._rows(idx_ptrs, dims, n_index)
        sel, s0 = self._elems(src, dt)
        total = int(np.prod(dims)) * inner
        s = sel[s0: s0 + total].reshape(-1, inner)
        del_, d0 = self._elems(dst, dt)
        del_[d0: d0 + n_index * inner] = s[row].reshape(-1)

    def scatter_add(self, dst, src, dt, idx_ptrs, dims, n_index, inner):
        self._launches += 1
        row = self._rows(idx_ptrs, dims, n_index)
        del_, d0 = self._elems(dst, dt)
        total = int(np.prod(dims)) * inner
        d = del_[d0: d0 + total].reshape(-1, inner)
        sel, s0 = self._elems(src, dt)
        np.add.at(d, row, sel[s0: s0 + n_index * inner].reshape(n_index, inner))

    def scatter_assign(self, dst, src, dt, idx_ptrs, dims, n_index, inner):
        self._launches += 1
        row = self._rows(idx_ptrs, dims, n_index)
        del_, d0 = self._elems(dst, dt)
        total = int(np.prod(dims)) * inner
        d = del_[d0: d0 + total].reshape(-1, inner)
        sel, s0 = self._elems(src, dt)
        d[row] = sel[s0: s0 + n_index * inner].reshape(n_index, inner)

    def cumulate(self, dst, src, dt, op, outer, n, inner):
        self._launches += 1
        s = self._view3(src, dt, outer, n, inner)
        d = self._view3(dst, dt, outer, n, inner)
        d[...] = (np.cumsum if op == "sum" else np.cumprod)(s, axis=1, dtype=np.dtype(dt))

    def sort(self, keys_out, idx_out, src, dt, rows, n):
        self._launches += 1
        sel, s0 = self._elems(src, dt)
        s = sel[s0: s0 + rows * n].reshape(rows, n)
        order = np.argsort(s, axis=1, kind="stable")
        if keys_out is not None:
            kel, k0 = self._elems(keys_out, dt)
            kel[k0: k0 + rows * n] = np.take_along_axis(s, order, axis=1).reshape(-1)
        if idx_out is not None:
            iel, i0 = self._elems(idx_out, np.int64)
            iel[i0: i0 + rows * n] = order.reshape(-1)

    def conv2d(self, out, a, b, dt, N, C, H, W, F, kh, kw, mode, flip):
        self._launches += 1
        ael, a0 = self._elems(a, dt)
        bel, b0 = self._elems(b, dt)
        A = ael[a0: a0 + N * C * H * W].reshape(N, C, H, W)
        B = bel[b0: b0 + C * F * kh * kw].reshape(C, F, kh, kw)
        if flip:
            B = B[:, :, ::-1, ::-1]
        if mode == 1:
            A = np.pad(A, ((0, 0), (0, 0), (kh - 1, kh - 1), (kw - 1, kw - 1)))
        Ho, Wo = A.shape[2] - kh + 1, A.shape[3] - kw + 1
        res = np.zeros((N, F, Ho, Wo), dtype=np.dtype(dt))
        for i in range(kh):
            for j in range(kw):
                res += np.einsum("nchw,cf->nfhw", A[:, :, i: i + Ho, j: j + Wo], B[:, :, i, j])
        oel, o0 = self._elems(out, dt)
        oel[o0: o0 + res.size] = res.reshape(-1)

    def conv2d_wgrad(self, db, a, g, dt, N, C, H, W, F, kh, kw):
        self._launches += 1
        Ho, Wo = H - kh + 1, W - kw + 1
        ael, a0 = self._elems(a, dt)
        gel, g0 = self._elems(g, dt)
        A = ael[a0: a0 + N * C * H * W].reshape(N, C, H, W)
        G = gel[g0: g0 + N * F * Ho * Wo].reshape(N, F, Ho, Wo)
        res = np.zeros((C, F, kh, kw), dtype=np.dtype(dt))
        for i in range(kh):
            for j in range(kw):
                res[:, :, kh - 1 - i, kw - 1 - j] = np.einsum("nchw,nfhw->cf", A[:, :, i: i + Ho, j: j + Wo], G)
        oel, o0 = self._elems(db, dt)
        oel[o0: o0 + res.size] = res.reshape(-1)

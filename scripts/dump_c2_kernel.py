"""Capture the fused kernel of C2 (egrad^4 of tanh) on the host emulation, print its generated CUDA source and,
with --sass, compile it with NVRTC (no GPU needed) and count the SASS instructions per class.

    python scripts/dump_c2_kernel.py [--order 4] [--sass] [--out /tmp/c2]
"""
import argparse
import collections
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "examples")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--sass", action="store_true")
    ap.add_argument("--out", default="/tmp/c2")
    args = ap.parse_args()
    import autograd_b200 as b200
    from autograd_b200 import codegen
    from fakedev import FakeBackend
    import workloads

    fb = FakeBackend(compile_generated=False)
    b200.backend.set_backend(fb)
    b200.install()
    f = workloads.tanh_nth_derivative(args.order)
    x = b200.asarray(np.linspace(-7, 7, 1 << 16))
    y = f(x)
    y.get()
    progs = sorted(fb.programs.values(), key=lambda p: -len(p.instrs))
    prog = progs[0]
    ops = collections.Counter(i[0] for i in prog.instrs)
    print("instrs", len(prog.instrs), dict(ops))
    src, name, _ = codegen.generate(prog)
    os.makedirs(args.out, exist_ok=True)
    with open(os.path.join(args.out, "c2.cu"), "w") as fh:
        fh.write(src)
    print("source ->", os.path.join(args.out, "c2.cu"), "kernel", name)
    if args.sass:
        import ctypes as C
        from autograd_b200 import _lib
        lib = _lib.lib
        p, n = C.c_void_p(), C.c_size_t()
        rc = lib.b200ag_rtc_compile(src.encode(), (name + ".cu").encode(), C.byref(p), C.byref(n))
        assert rc == 0, rc
        cubin = C.string_at(p.value, n.value)
        path = os.path.join(args.out, "c2.cubin")
        open(path, "wb").write(cubin)
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        open(os.path.join(args.out, "c2.sass"), "w").write(sass)
        cnt = collections.Counter()
        for line in sass.splitlines():
            line = line.strip()
            if line.startswith("/*") and "*/" in line[2:]:
                body = line.split("*/", 1)[1].strip()
                if not body or body.startswith("/*"):
                    continue
                tok = body.split()
                if tok[0].startswith("@"):
                    tok = tok[1:]
                if tok:
                    cnt[tok[0].split(".")[0].rstrip(";")] += 1
        print("SASS total", sum(cnt.values()))
        for k, v in cnt.most_common(25):
            print(f"  {k:10s} {v}")


if __name__ == "__main__":
    main()

"""FP64 instructions PER POINT of a generated fused kernel's fast body, counted from SASS without a GPU.

The generated kernel holds the body several times (two elements per thread, a tail loop, the exact-division twin),
so instruction counts of the whole kernel do not say what one point costs.  This tool takes the generated source
(scripts/dump_c2_kernel.py --out DIR writes DIR/c2.cu), keeps the PRELUDE and the fast `body`, replaces the rare
out-of-range fallback by a store of zero and wraps the body in a one-element kernel, compiles that with NVRTC through
libb200ag.so exactly as the product does (same flags) and counts SASS mnemonics.

    python scripts/count_body_instrs.py /tmp/c2/c2.cu
"""
import collections
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def per_point_counts(src: str, workdir: str = "/tmp/count_body"):
    from autograd_b200 import _lib

    m = re.search(r"__device__ __forceinline__ void body\((.*?)\) \{", src)
    sig = m.group(1)
    start = m.start()
    end = src.index("\n}\n", start) + 3
    body = src[start:end]
    body = re.sub(r"  if \(!ok_\) \{.*?return; \}\n", "  if (!ok_) { o0 = 0; return; }\n", body)
    head = src[: src.index("struct Params {")]
    params = src[src.index("struct Params {"): src.index("};", src.index("struct Params {")) + 2]
    n_in = len([a for a in sig.split(",") if a.strip().startswith("const double l")])
    n_out = len([a for a in sig.split(",") if "& o" in a])
    ins = ", ".join(f"in[{k}]" for k in range(n_in))
    outs = ", ".join(f"o{k}" for k in range(n_out))
    kern = ['extern "C" __global__ void one(const Params p, const double* in, double* out) {']
    kern += [f"  double o{k};" for k in range(n_out)]
    kern.append(f"  body(0, p{', ' + ins if ins else ''}, {outs});")
    kern += [f"  out[{k}] = o{k};" for k in range(n_out)]
    kern.append("}")
    full = head + params + "\n" + body + "\n" + "\n".join(kern) + "\n"
    os.makedirs(workdir, exist_ok=True)
    p, n = C.c_void_p(), C.c_size_t()
    rc = _lib.lib.b200ag_rtc_compile(full.encode(), b"one.cu", C.byref(p), C.byref(n))
    if rc:
        raise RuntimeError("\n".join(l for l in _lib.lib.b200ag_last_error().decode().splitlines() if "error" in l))
    path = os.path.join(workdir, "one.cubin")
    with open(path, "wb") as f:
        f.write(C.string_at(p.value, n.value))
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    for line in sass.splitlines():
        line = line.strip()
        if line.startswith("/*") and "*/" in line[2:]:
            b = line.split("*/", 1)[1].strip()
            if not b or b.startswith("/*"):
                continue
            tok = b.split()
            if tok[0].startswith("@"):
                tok = tok[1:]
            if tok:
                cnt[tok[0].split(".")[0].rstrip(";")] += 1
    return cnt, sass


if __name__ == "__main__":
    cnt, _ = per_point_counts(open(sys.argv[1]).read())
    fp64 = sum(cnt[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print("per point: FP64 pipe", fp64, {k: cnt[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")}, "all", sum(cnt.values()))

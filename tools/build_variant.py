"""Build an alternative libvidtome_b200 with extra -D flags into build/variants/<name>/ (tuning sweeps only).
  python tools/build_variant.py poly7 -DVTM_FA_POLY_NUM=7
  python tools/build_variant.py k0pf --src=rows -DVTM_K0_PREFETCH=1      # recompile rows.cu instead of attention.cu"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vidtome_b200 import _build  # noqa: E402


def build(name, defs):
    srcs_sel = [d.split("=", 1)[1] for d in defs if d.startswith("--src=")] or ["attention"]
    defs = [d for d in defs if not d.startswith("--src=")]
    out = os.path.join(ROOT, "build", "variants", name)
    os.makedirs(out, exist_ok=True)
    nvcc = _build._nvcc()
    srcs = _build.sources()
    objs = []

    def one(src):
        base = os.path.basename(src)[:-3]
        affected = base in srcs_sel or not os.path.exists(os.path.join(_build.OBJDIR, base + ".o"))
        obj = os.path.join(out, base + ".o") if affected else os.path.join(_build.OBJDIR, base + ".o")
        if affected:
            r = subprocess.run([nvcc, *_build.NVCC_FLAGS, *defs, "-c", src, "-o", obj], capture_output=True, text=True)
            if r.returncode:
                raise RuntimeError(r.stderr)
        return obj
    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(one, srcs))
    lib = os.path.join(out, "libvidtome_b200.so")
    r = subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"],
                       capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    return lib


if __name__ == "__main__":
    print(build(sys.argv[1], sys.argv[2:]))

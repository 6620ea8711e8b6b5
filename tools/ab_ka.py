"""A/B of the KA variants selected by ops.KA_VARIANT (vtm_sim_argmax vs vtm_sim_argmax_pair):
bit-exact check of each variant against the SIMT twin, then interleaved timing at the benchmark shapes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidtome_b200 import ops  # noqa: E402


def set_variant(two_cta: bool):
    if two_cta:
        ops.KA_VARIANT = "pair"
    else:
        ops.KA_VARIANT = "cta"


def check(B, Ns, Nd, C, align):
    g = torch.Generator(device="cuda").manual_seed(Ns * 7 + Nd)
    a = torch.randn(B, Ns, C, device="cuda", generator=g).half()
    b = torch.randn(B, Nd, C, device="cuda", generator=g).half()
    a = a / a.norm(dim=-1, keepdim=True)
    b = b / b.norm(dim=-1, keepdim=True)
    ref = ops.sim_argmax(a, b, align, simt=True)
    out, keys = {}, {}
    for v in (False, True):
        set_variant(v)
        keys[v] = ops.sim_argmax(a, b, align)
        torch.cuda.synchronize()
        # random fp16 rows have near-ties that the twin's sequential fp32 FMA order resolves differently
        out["2cta_vs_simt" if v else "1cta_vs_simt"] = float((keys[v] == ref).float().mean())
    same = bool(torch.equal(keys[False], keys[True]))
    print(json.dumps({"check": [B, Ns, Nd, C, align], "2cta_equals_1cta": same, **out}), flush=True)
    return same


def timeit(B, Ns, Nd, C, align, rounds=7, inner=6):
    a = torch.randn(B, Ns, C, device="cuda").half()
    b = torch.randn(B, Nd, C, device="cuda").half()
    res = {False: [], True: []}
    for v in (False, True):
        set_variant(v)
        for _ in range(3):
            ops.sim_argmax(a, b, align)
    torch.cuda.synchronize()
    for _ in range(rounds):
        for v in (False, True):
            set_variant(v)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(inner):
                ops.sim_argmax(a, b, align)
            e1.record()
            torch.cuda.synchronize()
            res[v].append(e0.elapsed_time(e1) / inner)
    fl = 2.0 * B * Ns * Nd * C
    line = {"shape": [B, Ns, Nd, C, align]}
    for v in (False, True):
        t = sorted(res[v])[len(res[v]) // 2]
        line["2cta" if v else "1cta"] = {"ms": round(t, 4), "TF": round(fl / t / 1e9, 1)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ok = True
    for shp in [(1, 256, 256, 64, False), (2, 300, 700, 320, False), (2, 1000, 5000, 640, True),
                (3, 129, 257, 1280, False), (2, 3072, 1024, 320, False), (1, 77, 33, 64, False)]:
        ok &= check(*shp)
    if not ok:
        sys.exit(1)
    for shp in [(2, 49152, 16384, 320, False), (2, 55296, 9216, 320, False), (2, 12288, 4096, 640, False),
                (2, 13824, 2304, 640, False), (2, 49152, 16384, 320, True)]:
        timeit(*shp)

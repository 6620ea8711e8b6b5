"""Sweep the share of exponentials evaluated on the FMA pipe in the flash kernel (VTM_FA_POLY_NUM / 16).
Each variant is a separate build of the library (tools/build_variant.py), timed in its own process:
  python tools/sweep_fa_poly.py            # on the GPU box; variants must have been built beforehand
Prints one JSON line per (variant, shape): ms for qkv + flash + out (back-to-back launches over rotating inputs),
max relative error vs an fp32 torch reference, and the torch module's own time for the same shape."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(2, 10241, 320, 8), (2, 2561, 640, 8), (2, 15668, 320, 5), (2, 5325, 320, 8)]


def child():
    import torch
    sys.path.insert(0, ROOT)
    from vidtome_b200 import ops
    name = os.environ["VTM_VARIANT"]
    for (B, L, C, H) in SHAPES:
        g = torch.Generator(device="cuda").manual_seed(0)
        xs = [torch.randn((B, L, C), generator=g, device="cuda").half() for _ in range(4)]
        ws = [(torch.randn((C, C), generator=g, device="cuda") / C ** 0.5).half() for _ in range(4)]
        wqkv = torch.cat(ws[:3], 0).contiguous()
        bo = torch.zeros(C, device="cuda").half()
        d = C // H
        y = ops.attention(xs[0], wqkv, ws[3], bo, H, d ** -0.5)
        q, k, v = [(xs[0].float() @ w.float().t()).view(B, L, H, d).transpose(1, 2) for w in ws[:3]]
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, C) @ ws[3].float().t()
        err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
        for _ in range(3):
            for x in xs:
                ops.attention(x, wqkv, ws[3], bo, H, d ** -0.5)
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(12):
                ops.attention(xs[i % 4], wqkv, ws[3], bo, H, d ** -0.5)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) / 12)
        ts.sort()
        out = {"variant": name, "B": B, "L": L, "C": C, "heads": H, "head_dim": d, "ms": round(ts[2], 4), "max_rel_err": float(f"{err:.3e}")}
        if name == "default":
            def sdpa(x):
                q, k, v = [torch.nn.functional.linear(x, w).view(B, L, H, d).transpose(1, 2) for w in ws[:3]]
                o = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, C)
                return torch.nn.functional.linear(o, ws[3], bo)
            for _ in range(3):
                sdpa(xs[0])
            tt = []
            for _ in range(5):
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for i in range(12):
                    sdpa(xs[i % 4])
                e.record()
                torch.cuda.synchronize()
                tt.append(s.elapsed_time(e) / 12)
            tt.sort()
            out["torch_sdpa_path_ms"] = round(tt[2], 4)
        print(json.dumps(out), flush=True)


def main():
    vdir = os.path.join(ROOT, "build", "variants")
    names = ["default"] + sorted(n for n in os.listdir(vdir) if os.path.exists(os.path.join(vdir, n, "libvidtome_b200.so"))) if os.path.isdir(vdir) else ["default"]
    runs = [(n, {}) for n in names]
    if "--ab" in sys.argv:
        runs += [(n + "_no_embed", {"VTM_FA_EMBED": "0"}) for n in names] + [("default_groups", {"VTM_FA_GROUPS": "1"})]
    if "--pairs" in sys.argv:      # eight softmax warps per CTA (flash_attn_kernel HALVES = 2) against four
        runs += [(n + "_pairs1", {"VTM_FA_PAIRS": "1"}) for n in names] + [(n + "_pairs0", {"VTM_FA_PAIRS": "0"}) for n in names]
    if "--repeat2" in sys.argv:     # interleaved A/B: the whole list twice
        runs = runs + runs
    for n, extra in runs:
        env = dict(os.environ, VTM_VARIANT=n, **extra)
        base = n.replace("_no_embed", "").replace("_groups", "").replace("_pairs1", "").replace("_pairs0", "")
        if base != "default":
            env["VIDTOME_B200_LIB"] = os.path.join(vdir, base, "libvidtome_b200.so")
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, check=False)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()

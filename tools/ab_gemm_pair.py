"""A/B of the CTA-pair mainloop for the wide projections (VTM_GEMM_PAIR=0 / 1, one child process each):
  python tools/ab_gemm_pair.py        # on the GPU box
Prints one JSON line per (mode, op, shape): ms per call over rotating inputs, TFLOP/s, max abs difference to mode 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES_GEGLU = [(131072, 2560, 320), (32768, 5120, 640)]          # FF1 at ds1 / ds2 of the full workload (N = 8 dim)
SHAPES_LINEAR = [(20482, 1280, 320), (131072, 1280, 320)]
SHAPES_KD = [(2, 10241, 320, 8), (2, 2561, 640, 8), (2, 5325, 320, 8), (2, 15668, 320, 5)]


def timed(fn, n=10, reps=5):
    import torch
    for _ in range(3):
        fn(0)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n):
            fn(i)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / n)
    return sorted(ts)[len(ts) // 2]


def child():
    import torch
    sys.path.insert(0, ROOT)
    from vidtome_b200 import ops
    mode = os.environ.get("VTM_GEMM_PAIR", "unset")
    g = torch.Generator(device="cuda").manual_seed(0)
    out = {}
    for (M, N, K) in SHAPES_GEGLU:
        xs = [torch.randn((M, K), generator=g, device="cuda").half() for _ in range(3)]
        w = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).half()
        b = (0.1 * torch.randn((N,), generator=g, device="cuda")).half()
        w_il, b_il = ops.interleave_geglu(w, b)
        y = ops.linear_geglu(xs[0], w_il, b_il)
        ms = timed(lambda i: ops.linear_geglu(xs[i % 3], w_il, b_il))
        print(json.dumps({"mode": mode, "op": "linear_geglu", "M": M, "N": N, "K": K, "ms": round(ms, 4),
                          "tflops": round(2.0 * M * N * K / ms / 1e9, 1), "checksum": float(y.float().abs().sum())}), flush=True)
    for (M, N, K) in SHAPES_LINEAR:
        xs = [torch.randn((M, K), generator=g, device="cuda").half() for _ in range(3)]
        w = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).half()
        y = ops.linear(xs[0], w)
        ms = timed(lambda i: ops.linear(xs[i % 3], w))
        print(json.dumps({"mode": mode, "op": "linear", "M": M, "N": N, "K": K, "ms": round(ms, 4),
                          "tflops": round(2.0 * M * N * K / ms / 1e9, 1), "checksum": float(y.float().abs().sum())}), flush=True)
    for (B, L, C, H) in SHAPES_KD:
        xs = [torch.randn((B, L, C), generator=g, device="cuda").half() for _ in range(4)]
        ws = [(torch.randn((C, C), generator=g, device="cuda") / C ** 0.5).half() for _ in range(4)]
        wqkv = torch.cat(ws[:3], 0).contiguous()
        d = C // H
        y = ops.attention(xs[0], wqkv, ws[3], None, H, d ** -0.5)
        ms = timed(lambda i: ops.attention(xs[i % 4], wqkv, ws[3], None, H, d ** -0.5), n=12)
        print(json.dumps({"mode": mode, "op": "attention (KD)", "B": B, "L": L, "C": C, "heads": H, "ms": round(ms, 4),
                          "checksum": float(y.float().abs().sum())}), flush=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        for m in (sys.argv[1:] or ["0", "1"]):
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, VTM_GEMM_PAIR=m), check=False)

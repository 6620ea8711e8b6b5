"""Is the bench step GPU-bound or host-bound?  Measures, per step: host enqueue time (no sync inside the loop),
wall time with the GPU drained, and the number of host<->device syncs the step performs."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vidtome_b200
from vidtome_b200.driver import ChunkedDenoiser
from vidtome_b200.skeleton import make_skeleton

torch.manual_seed(123)
net = make_skeleton("sd15", device="cuda")
vidtome_b200.apply_patch(net, local_merge_ratio=0.9, batch_size=2)
den = ChunkedDenoiser(net, n_timesteps=50, chunk_size=16)
x = torch.randn(16, 4, 64, 64, device="cuda", dtype=torch.float16)
for i in range(3):
    x = den.step(x, i)
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for i in range(N):
    x = den.step(x, i % 50)
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue {t_enq / N * 1e3:.2f} ms/step; wall incl. drain {t_all / N * 1e3:.2f} ms/step "
      f"-> {'HOST' if t_enq > 0.9 * t_all else 'GPU'}-bound (GPU backlog at loop end {1e3 * (t_all - t_enq):.1f} ms)")

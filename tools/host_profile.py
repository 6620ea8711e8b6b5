"""cProfile of the host side of the bench step (where does the enqueue time go?)."""
import cProfile, pstats, os, sys, io
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
pr = cProfile.Profile()
pr.enable()
for i in range(10):
    x = den.step(x, i)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])

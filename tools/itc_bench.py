"""ITC head at the bench size (B=96, Q=36864): CUDA-event time per call and per kernel; target of the ncu capture."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from spmm_b200 import kernels as K
B, Q, E = 96, 36864, 256
dev = "cuda"
z = [torch.randn(B, E, device=dev) for _ in range(4)]
pq, tq = F.normalize(torch.randn(Q, E, device=dev), dim=1), F.normalize(torch.randn(Q, E, device=dev), dim=1)
temp = torch.tensor(0.07, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3):
    K.itc(z[0], z[1], z[2], z[3], pq, tq, temp, 0.4)
torch.cuda.synchronize()
for cold in (0, 1):
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2_000_000)     # let the host run ahead so the events bracket GPU work only
        s.record()
        K.itc(z[0], z[1], z[2], z[3], pq, tq, temp, 0.4)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    print("itc fwd+bwd %s: median %.1f us  min %.1f us  (algorithmic queue bytes 2 x 75.5 MB -> %.0f GB/s at median)"
          % ("cold L2" if cold else "warm L2", ts[len(ts) // 2], ts[0], 151.0e6 / ts[len(ts) // 2] / 1e3))

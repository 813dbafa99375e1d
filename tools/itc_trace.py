"""Timeline of one CTA of the ITC pass-1 kernel (%globaltimer stamps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from spmm_b200 import kernels as K, _lib
B, Q, E = 96, 36864, 256
dev = "cuda"
z = [torch.randn(B, E, device=dev) for _ in range(4)]
pq, tq = F.normalize(torch.randn(Q, E, device=dev), dim=1), F.normalize(torch.randn(Q, E, device=dev), dim=1)
temp = torch.tensor(0.07, device=dev)
for _ in range(3):
    K.itc(z[0], z[1], z[2], z[3], pq, tq, temp, 0.4)
tr = torch.zeros(64, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
_lib.lib().spmm_itc_debug_trace(tr.data_ptr())
K.itc(z[0], z[1], z[2], z[3], pq, tq, temp, 0.4)
torch.cuda.synchronize()
_lib.lib().spmm_itc_debug_trace(None)
t = tr.cpu().tolist()
names = {0: "entry", 1: "mma thread ready", 2: "q_ready seen", 3: "mma thread done", 4: "Q fill done (warp 2)", 5: "stats done (warp 2)", 6: "exit"}
for i in range(12):
    names[8 + 3 * i] = "tile%d: loop top" % i
    names[9 + 3 * i] = "tile%d: keys landed" % i
    names[10 + 3 * i] = "tile%d: S buffer free" % i
for k in sorted(names):
    if t[k]:
        print("%-26s %8.2f us" % (names[k], (t[k] - t[0]) / 1e3))

"""A few eager launches of the kernels that get an `ncu --set full` capture at the step's largest shapes:
attention backward (384 x 12 heads, Tq = Tk = 64, dropout 0.1, bias sums) and the fused LayerNorm backward (24576 and
12288 rows of 768).  Usage: ncu --set full --import-source on -k regex:"ln_bwd_fused|attn_bwd_tc" python tools/ncu_targets.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmm_b200 import kernels as K
DEV, BF = "cuda", torch.bfloat16
H, h = 768, 12
B, T = 384, 64
qkv = torch.randn(B * T, 3 * H, device=DEV).to(BF)
q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
o = torch.empty(B * T, H, device=DEV, dtype=BF)
lse = torch.empty(B * h * T, device=DEV)
K.attn_fwd(q, k, v, o, lse, B, h, T, T, None, False, 0.125, 0.1, 5)
do = torch.randn(B * T, H, device=DEV).to(BF)
dqkv = torch.empty(B * T, 3 * H, device=DEV, dtype=BF)
db = torch.zeros(3, H, device=DEV)
for _ in range(2):
    K.attn_bwd(do, q, k, v, o, lse, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], B, h, T, T, None, False, 0.125, 0.1, 5,
               dbias=(db[0], db[1], db[2]))
for rows in (24576, 12288):
    x = torch.randn(rows, H, device=DEV).to(BF)
    g, b = torch.ones(H, device=DEV), torch.zeros(H, device=DEV)
    y, mean, rstd = K.layernorm_fwd(x, g, b, 1e-12)
    dy = torch.randn(rows, H, device=DEV).to(BF)
    dg, dbt, dbias = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    for _ in range(2):
        K.layernorm_bwd(dy, x, mean, rstd, g, dg, dbt, dbias=dbias, want_branch=True, branch_dropout_p=0.1, branch_seed=3)
torch.cuda.synchronize()
print("done")

"""Debug aid: compares the intermediate buffers of spmm_itc_fwd_bwd (row LSEs, O accumulators) with torch."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from spmm_b200 import _lib
B, Q, E = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 96, 256
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(1)
z = [torch.randn(B, E, generator=g).to(dev) for _ in range(4)]
z[2] = z[0] + 0.05 * z[2]; z[3] = z[1] + 0.05 * z[3]
pq = F.normalize(torch.randn(Q, E, generator=g), dim=1).to(dev)
tq = F.normalize(torch.randn(Q, E, generator=g), dim=1).to(dev)
temp = torch.tensor(0.07, device=dev)
L = _lib.lib()
nb = L.spmm_itc_workspace_bytes(B, E, Q)
ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
f32 = dict(device=dev, dtype=torch.float32)
o = dict(loss=torch.empty((), **f32), dzp=torch.empty(B, E, **f32), dzt=torch.empty(B, E, **f32), dtemp=torch.empty((), **f32),
         s1=torch.empty(B, B, **f32), s2=torch.empty(B, B, **f32), fm1=torch.empty(B, E, **f32), fm2=torch.empty(B, E, **f32),
         nan=torch.empty((), **f32))
rc = L.spmm_itc_fwd_bwd(z[0].data_ptr(), z[1].data_ptr(), z[2].data_ptr(), z[3].data_ptr(), pq.data_ptr(), tq.data_ptr(),
                        temp.data_ptr(), 0.4, B, E, Q, o["loss"].data_ptr(), o["dzp"].data_ptr(), o["dzt"].data_ptr(),
                        o["dtemp"].data_ptr(), o["s1"].data_ptr(), o["s2"].data_ptr(), o["fm1"].data_ptr(), o["fm2"].data_ptr(),
                        o["nan"].data_ptr(), ws.data_ptr(), nb, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("rc", rc)
mt = (4 * B + 127) // 128; rows_pad = mt * 128
nh, nq = (B + 31) // 32, (Q + 31) // 32
splits = max(1, min(nh + nq, 148 // (2 * mt))); tps = -(-(nh + nq) // splits); splits = -(-(nh + nq) // tps)
base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
w = ws[base:].view(torch.float32)
off = 0
feats = w[off:off + 4 * B * E].view(4, B, E); off += 4 * B * E
off += (4 * B + 63) // 64 * 64
qm = w[off:off + 8 * B * E].view(2, 4 * B, E); off += 8 * B * E
part = w[off:off + 2 * splits * rows_pad * 2].view(2, splits, rows_pad, 2); off += 2 * splits * rows_pad * 2
lse = w[off:off + 2 * rows_pad].view(2, rows_pad); off += 2 * rows_pad
oacc = w[off:off + 8 * B * E].view(2, 4 * B, E)
fe = [F.normalize(x, dim=-1) for x in z]
print("feats err", float((feats - torch.stack(fe)).abs().max()))
Qm = [torch.cat([fe[0], fe[1], fe[2], fe[3]]), torch.cat([fe[1], fe[0], fe[3], fe[2]])]
print("qm err", float((qm - torch.stack(Qm)).abs().max()))
keys = [torch.cat([fe[3], tq]), torch.cat([fe[2], pq])]
for ks in range(2):
    S = Qm[ks] @ keys[ks].t() / temp
    l2 = torch.logsumexp(S, dim=1) / math.log(2)
    P = torch.softmax(S, dim=1)
    O = P @ keys[ks]
    print("ks", ks, "lse2 err", float((lse[ks, :4 * B] - l2).abs().max()), "O rel err", float((oacc[ks] - O).norm() / O.norm()))
    print("   lse2 got", lse[ks, :4].tolist(), "want", l2[:4].tolist())
    print("   psum row0..3 (debug)", part[ks, :, :4, 0].sum(0).tolist())
    print("   O got", oacc[ks, 0, :4].tolist(), "want", O[0, :4].tolist())
    print("   O row0 absmax", float(oacc[ks, 0].abs().max()), "nonzero count", int((oacc[ks] != 0).sum()))

"""CPU check of the algebra behind the tensor-core ITC head (spmm_b200/csrc/itc.cu): loss, d loss/d z and d loss/d temp of
the reference's contrastive block (SPMM_models.py:102-131) follow from row LSEs and O = softmax(S) K alone, and a
bit-level emulation of the kernel's TF32 operand handling (queries / P rounded to nearest, raw queue keys truncated) stays
inside the tolerances the GPU test uses (tests/test_kernels_gpu.py::test_itc_matches_oracle)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import spmm_ref


def tf32_rn(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def o_linear_itc(z, pq, tq, temp, alpha, emulate_tf32):
    B, E = z[0].shape
    feats = [F.normalize(x, dim=-1) for x in z]
    norms = [x.norm(dim=-1).clamp_min(1e-12) for x in z]
    rq = tf32_rn if emulate_tf32 else (lambda x: x)
    rk = tf32_trunc if emulate_tf32 else (lambda x: x)
    Qm = [torch.cat([feats[0], feats[1], feats[2], feats[3]]), torch.cat([feats[1], feats[0], feats[3], feats[2]])]
    keys = [torch.cat([feats[3], tq]), torch.cat([feats[2], pq])]
    inv = 1.0 / temp
    O, lse = [], []
    for ks in range(2):
        Kq = torch.cat([rq(keys[ks][:B]), rk(keys[ks][B:])])          # own-momentum head keys come from the rounded Qm
        S = (rq(Qm[ks]).double() @ Kq.double().t()).float() * inv
        l = torch.logsumexp(S, dim=1)
        P = torch.exp(S - l[:, None])
        if emulate_tf32:
            P = tf32_rn(P)
        O.append((P.double() @ Kq.double()).float())
        lse.append(l)
    gscale = inv / (2 * B)
    l_acc, dt_acc, dz = 0.0, 0.0, []
    for which in range(2):
        f, g = feats[which], torch.zeros(B, E)
        for ks in range(2):
            r0 = 0 if which == ks else B
            S_, T_, kp = O[ks][r0:r0 + B], O[ks][2 * B + r0:2 * B + r0 + B], feats[3 if ks == 0 else 2]
            ds, dt, dk = (f * S_).sum(1) * inv, (f * T_).sum(1) * inv, (f * kp).sum(1) * inv
            g = g + gscale * (S_ - alpha * T_ - (1 - alpha) * kp)
            l_acc = l_acc + (lse[ks][r0:r0 + B] - alpha * dt - (1 - alpha) * dk).sum()
            dt_acc = dt_acc + (ds - alpha * dt - (1 - alpha) * dk).sum()
        dz.append((g - f * (f * g).sum(1, keepdim=True)) / norms[which][:, None])
    return l_acc / (2 * B), dz, -dt_acc / (temp * 2 * B)


@pytest.mark.parametrize("B,Q", [(8, 96), (6, 96), (96, 4096)])
def test_itc_from_lse_and_o(B, Q):
    torch.manual_seed(0)
    E, alpha = 256, 0.4
    z = [torch.randn(B, E) for _ in range(4)]
    z[2] = z[0] + 0.05 * z[2]
    z[3] = z[1] + 0.05 * z[3]
    pq, tq = F.normalize(torch.randn(Q, E), dim=1), F.normalize(torch.randn(Q, E), dim=1)
    temp = torch.tensor(0.07)
    zp, zt, tr = z[0].clone().requires_grad_(True), z[1].clone().requires_grad_(True), temp.clone().requires_grad_(True)
    loss, _, _ = spmm_ref.itc_loss(F.normalize(zp, dim=-1), F.normalize(zt, dim=-1), F.normalize(z[2], dim=-1),
                                   F.normalize(z[3], dim=-1), pq.t().contiguous(), tq.t().contiguous(), tr, alpha)
    loss.backward()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    # exact arithmetic: the O-linear formulation IS the reference's loss and gradients
    L, dz, dtemp = o_linear_itc(z, pq, tq, temp, alpha, emulate_tf32=False)
    assert abs(float(L) - float(loss)) < 5e-6 * abs(float(loss)) + 5e-6
    assert rel(dz[0], zp.grad) < 5e-6 and rel(dz[1], zt.grad) < 5e-6
    assert abs(float(dtemp) - float(tr.grad)) < 1e-4 * abs(float(tr.grad))
    # with the kernel's TF32 operand handling: inside the GPU test's bounds (4e-3, 1e-3, 5e-3)
    L, dz, dtemp = o_linear_itc(z, pq, tq, temp, alpha, emulate_tf32=True)
    assert abs(float(L) - float(loss)) < 2e-3
    assert rel(dz[0], zp.grad) < 5e-4 and rel(dz[1], zt.grad) < 5e-4
    assert abs(float(dtemp) - float(tr.grad)) < 4e-3 * abs(float(tr.grad))

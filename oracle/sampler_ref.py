"""TEST INFRASTRUCTURE ONLY -- CPU replica of the counter-based hard-negative sampler
(spmm_b200/csrc/misc.cu `sample_neg_kernel`), which stands in for the 2B sequential
`torch.multinomial(w, 1)` draws of reference SPMM_models.py:154-178.

Distribution: index j drawn with probability w[b][j] / sum(w[b]) where w = softmax(sim[b,:B]) with the
diagonal zeroed (inverse-CDF on a Philox4x32-10 uniform).  Every operation is a separately rounded fp32
mul/add, so numpy reproduces the GPU bit for bit (no libm exp/log on either side).
"""
import numpy as np

F = np.float32
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(key, ctr):
    k0, k1 = key
    c = list(ctr)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def exact_exp_neg(x):
    x = F(x)
    if x < F(-87.0):
        return F(0.0)
    t = F(x * F(1.44269504))
    n = F(np.floor(F(t + F(0.5))))
    r = F(x - F(n * F(0.693359375)))
    r = F(r - F(n * F(-2.12194440e-4)))
    p = F(F(r * F(1.3888889e-3)) + F(8.3333333e-3))
    for c in (4.1666667e-2, 1.6666667e-1, 0.5, 1.0, 1.0):
        p = F(F(p * r) + F(c))
    e = int(n) + 127
    scale = np.array([e << 23], dtype=np.uint32).view(np.float32)[0]
    return F(p * scale)


def sample_row(sim_row, b, stream, seed, step):
    B = len(sim_row)
    sim_row = np.asarray(sim_row, dtype=np.float32)
    mx = sim_row.max()
    w = [F(0.0) if j == b else exact_exp_neg(F(sim_row[j] - mx)) for j in range(B)]
    c = philox4x32_10((seed & MASK, (seed >> 32) & MASK), (b, stream, step & MASK, (step >> 32) & MASK))
    u = F(F(c[0] >> 8) * F(5.9604644775390625e-8))
    total = F(0.0)
    for j in range(B):
        total = F(total + w[j])
    target = F(u * total)
    cum, idx, last = F(0.0), -1, (1 if (b == 0 and B > 1) else 0)
    for j in range(B):
        if w[j] > 0:
            last = j
        cum = F(cum + w[j])
        if idx < 0 and cum > target and w[j] > 0:
            idx = j
    return idx if idx >= 0 else last


def sample_negatives(sim_i2t, sim_t2i, seed, step):
    """Returns (neg_t2i, neg_i2t) like spmm_sample_negatives."""
    B = sim_i2t.shape[0]
    t2i = [sample_row(sim_t2i[b], b, 0, seed, step) for b in range(B)]
    i2t = [sample_row(sim_i2t[b], b, 1, seed, step) for b in range(B)]
    return t2i, i2t

"""TEST INFRASTRUCTURE: host restatement of the reference's beam bookkeeping (d_pv2smiles_batched.py:29-52) driven by
RECORDED per-step logits instead of a model, so the device-side `spmm_beam_step` kernel can be checked in isolation:
the same logits must produce the same beams, scores and finished list.  Never imported by spmm_b200/."""
import torch


def replay_beams(step_logits, k, cls_id, sep_id):
    """step_logits[t]: [k, V] fp32 logits of ONE molecule's k beam rows at step t (row b = beam b at that step).
    Returns (finished list [(score, tokens)], per-step live beams [list of token lists])."""
    beams = [[cls_id] for _ in range(k)]
    scores = torch.zeros(k)
    finished, history = [], []
    for t, logits in enumerate(step_logits):
        logp = torch.log_softmax(logits.float(), dim=-1)
        vals, idx = torch.topk(logp, k, dim=-1)                      # [k, k]
        if t == 0:                                                   # single [CLS] beam: only row 0 is real (:29-32)
            scores = vals[0].clone()
            beams = [[cls_id, int(idx[0, j])] for j in range(k)]
            history.append([list(b) for b in beams])
            continue
        cand = scores[:, None] + vals
        done = False
        if bool((idx == sep_id).any()):                              # :39-46
            for b, j in (idx == sep_id).nonzero(as_tuple=False).tolist():
                finished.append((float(cand[b, j]), beams[b] + [sep_id]))
                cand[b, j] = -1e5
            if len(finished) >= k:
                done = True
        if done:
            break
        scores, flat = torch.topk(cand.flatten(), k)                 # :47-49
        beams = [beams[int(f) // k] + [int(idx[int(f) // k, int(f) % k])] for f in flat]
        history.append([list(b) for b in beams])
    finished = sorted(finished, key=lambda x: x[0], reverse=True)[:k]
    return finished, history

#!/bin/bash
# In-graph cost of each kernel family: step time of `bench.py --profile` with the family's entry points turned into no-ops
# (SPMM_DEBUG_SKIP, spmm_b200/_lib.py) subtracted from the full step.  Results of the skipped runs are garbage; timing only.
# Usage (GPU box): bash tools/skip_table.sh > gpurun_out/skip_table.txt
run() { SPMM_DEBUG_SKIP="$1" python bench.py --profile --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.readline())['ms_per_step'])"; }
full=$(run "")
echo "full step: $full ms"
for fam in spmm_layernorm_fwd spmm_layernorm_bwd spmm_colsum_bf16 spmm_attn_fwd spmm_attn_bwd spmm_itc_fwd_bwd "spmm_grad_sumsq,spmm_adamw_step,spmm_adam_tick" spmm_ema_multi "spmm_embed_text_bwd,spmm_embed_inputs_bwd,spmm_pv_tokens_bwd" "spmm_itm_loss_fwd_bwd,spmm_mpm_loss_fwd_bwd,spmm_lm_loss_fwd_bwd" spmm_gemm_bf16; do
  t=$(run "$fam")
  python -c "print('%-70s without: %8.3f ms   family cost: %7.3f ms' % ('$fam', $t, $full - $t))"
done

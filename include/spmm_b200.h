/* spmm_b200 -- C ABI of the B200 (sm_100a) kernels behind the SPMM pre-training step.
 *
 * The reference (jinhojsk515/spmm) has no FFI layer: its hot path is `SPMM.forward`
 * (SPMM_models.py:79-256) executing PyTorch ops.  Each entry point below replaces one group of those
 * ops; the citation after each declaration names the reference lines it stands in for.  The host-side
 * mirror of the reference interface (spmm_b200/SPMM_models.py, spmm_b200/xbert.py) binds these through
 * ctypes (spmm_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: all pointers are DEVICE pointers owned by the caller (the library never allocates
 * persistent memory); `stream` is a cudaStream_t; bf16 = raw uint16 storage (__nv_bfloat16);
 * return value 0 = ok, <0 = argument / setup error, >0 = cudaError_t.  No global state except
 * read-only kernel attributes, so calls are thread-safe for one stream per caller.
 */
#ifndef SPMM_B200_H_
#define SPMM_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int spmm_version(void);
/* Device u64 added (times a constant) to every dropout / sampler seed inside the kernels; NULL disables.  Lets a CUDA
 * graph of the training step replay with fresh randomness: the step bumps the scalar, per-op seeds stay baked. */
int spmm_set_rng_salt_ptr(const unsigned long long* dev_ptr);

/* ------------------------------------------------------------------ GEMM (tcgen05 / TMEM / TMA)
 * C[M,N] (+)= epilogue(alpha * A[M,K] . B[N,K]^T).  Operand majors: 0 = K-major (row-major [rows][K],
 * leading dim ld), 1 = MN-major (row-major [K][rows]).  Replaces nn.Linear forward and the autograd
 * dgrad/wgrad of xbert.py:280-298,370,435,448,673,695 and SPMM_models.py:92,95,202,251. */
enum {
  SPMM_GEMM_OUT_F32 = 1,    /* C is fp32 (default bf16) */
  SPMM_GEMM_ACCUMULATE = 2, /* C += (fp32 only; wgrad accumulation into the flat gradient arena) */
  SPMM_GEMM_GELU = 4,       /* exact-erf GELU (ACT2FN["gelu"], xbert.py:430) after bias */
  SPMM_GEMM_DGELU = 8,      /* multiply by gelu'(dgelu_pre_act) (autograd of xbert.py:436) */
  SPMM_GEMM_DGELU_STORED = 16 /* with GELU + pre_act: pre_act receives gelu'(pre) (same erf evaluation) instead of pre;
                                 with DGELU: dgelu_pre_act already holds gelu'(pre), the epilogue only multiplies */
};
typedef struct spmm_gemm_epilogue {
  const float* bias;           /* [N] fp32 or NULL */
  const void* residual;        /* bf16 [M][ld_residual] added last, or NULL (xbert.py:372,450) */
  int ld_residual;
  void* pre_act;               /* bf16 [M][ld_pre_act]: value after bias, before GELU; or NULL */
  int ld_pre_act;
  const void* dgelu_pre_act;   /* bf16 [M][ld]: pre-activation for SPMM_GEMM_DGELU */
  int ld_dgelu_pre_act;
  int flags;
  float alpha;
  float dropout_p;             /* inverted dropout on (acc+bias[,gelu]) before the residual (xbert.py:371,449) */
  unsigned long long dropout_seed;
  float* colsum;               /* optional [N] fp32: += column sums of the bf16 output C (the bias gradient of the dense that
                                  produced the tensor C is the gradient of; bf16 outputs of the 2-CTA kernel only) */
} spmm_gemm_epilogue;
int spmm_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, void* C,
                   int ldc, int M, int N, int K, const spmm_gemm_epilogue* epi, void* stream);
int spmm_gemm_debug_config(int mn_lbo_bytes, int mn_sbo_bytes, int force_bn, int max_ctas);
/* debug: device buffer of 16 x u64 per CTA receiving %globaltimer phase stamps of the next GEMM launches; NULL = off */
int spmm_gemm_debug_trace(void* buf);
int spmm_gemm_debug_trace_ring(void* buf, long slots);   /* one 148 x 16 x u64 slot per launch, ring of `slots` */

/* ------------------------------------------------------------------ attention core (xbert.py:305-354)
 * softmax(Q K^T * scale + mask) V per (batch, head), head_dim 64, Tq,Tk <= 128.  q/k/v/o rows are
 * tokens (b*T + t) with leading dims ld*, head h at column 64*h.  kv_len[b] = number of valid keys
 * (pad mask, xbert.py:947 / invert_attention_mask), NULL = all; causal = is_decoder mask (xbert.py:911-931):
 * 0 = none, 1 = every batch element, 1 + n = batch elements with index >= n only (lets a bidirectional and a causal
 * pass over the same weights share one launch: rows [0, n) bidirectional, rows [n, batch) causal).
 * kv_batch_stride_rows = Tk normally, 0 broadcasts one K/V over the batch (beam decode,
 * d_pv2smiles_single.py:29-36).  lse[B*heads*Tq] fp32 is saved for backward.
 * kv_index (optional, int32[batch] on the device) with kv_batches > 0: batch element b attends to the K/V rows of
 * batch element kv_index[b] of a [kv_batches * Tk]-row K/V buffer (kv_len stays per b).  The ITM pass of
 * SPMM_models.py:137-198 pairs the SAME encoder states with several query batches (positives, hard negatives), so
 * their key/value projection is computed once per distinct state instead of once per pair.  In spmm_attn_bwd dk/dv
 * stay per batch element b ([batch * Tk] rows); spmm_segment_sum_rows_bf16 folds them onto the distinct states. */
int spmm_attn_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                  float* lse, int batch, int heads, int Tq, int Tk, const int* kv_len, int causal,
                  int kv_batch_stride_rows, float scale, float dropout_p, unsigned long long seed,
                  const int* kv_index, int kv_batches, void* stream);
int spmm_attn_debug_trace(void* buf);   /* debug: 32 x u64 %globaltimer stamps per CTA of the next forward launches */
int spmm_attn_bwd(const void* d_o, int lddo, const void* q, int ldq, const void* k, int ldk, const void* v,
                  int ldv, const void* o, int ldo, const float* lse, void* dq, int lddq, void* dk, int lddk,
                  void* dv, int lddv, int batch, int heads, int Tq, int Tk, const int* kv_len, int causal,
                  float scale, float dropout_p, unsigned long long seed, float* dbias_q, float* dbias_k, float* dbias_v,
                  const int* kv_index, int kv_batches, void* stream);
/* dbias_q/k/v (optional, all or none; [heads*64] fp32 each): += column sums of dq / dk / dv, i.e. the bias gradients of
 * the query / key / value projections (autograd of xbert.py:280-298), taken from the tiles the kernel already holds. */

/* ------------------------------------------------------------------ LayerNorm (eps 1e-12; xbert.py:184,366,444,670)
 * y = LN(x) * gamma + beta, optional inverted dropout on y (BertEmbeddings, xbert.py:219).
 * bwd: dx (bf16), atomically accumulates dgamma/dbeta (fp32).  If dx_branch != NULL also writes
 * dx_branch = dx * dropout-mask(branch_seed) (the gradient flowing into `dropout(dense(.))` of
 * xbert.py:371/449) and, if dbias != NULL, accumulates its column sums into dbias (bias grad of that dense).
 * workspace: >= (8*3*H + 8) floats, zero before first use; the kernel leaves it zeroed (reusable across calls). */
int spmm_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                       int rows, int H, float eps, float dropout_p, unsigned long long seed, void* stream);
int spmm_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                       void* dx, float* dgamma, float* dbeta, void* dx_branch, float* dbias, int rows, int H,
                       float out_dropout_p, unsigned long long out_seed, float branch_dropout_p,
                       unsigned long long branch_seed, float* workspace, void* stream);

/* ------------------------------------------------------------------ embeddings (xbert.py:193-220, SPMM_models.py:82-88)
 * text: x = word[ids] + pos[t] + type[0] (pre-LN sum, bf16).  pv: properties = cat(cls, embed(pv)*(1-m) + mask_tok*m)
 * then + pos + type with the *encoder's* tables. */
int spmm_embed_text_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0, void* x,
                        int rows, int T, int H, void* stream);
int spmm_embed_text_bwd(const void* dx, const int64_t* ids, float* dword, float* dpos, float* dtype0, int rows,
                        int T, int H, int pad_id, void* stream);
int spmm_pv_tokens_fwd(const float* pv, const float* mpm_mask, const float* w_embed, const float* b_embed,
                       const float* cls_tok, const float* mask_tok, void* properties, int batch, int n_prop, int H,
                       void* stream);
int spmm_pv_tokens_bwd(const void* dproperties, const float* pv, const float* mpm_mask, float* dw_embed,
                       float* db_embed, float* dcls, float* dmask_tok, int batch, int n_prop, int H, void* stream);
int spmm_embed_inputs_fwd(const void* inputs, const float* pos, const float* type0, void* x, int rows, int T, int H,
                          void* stream);
int spmm_embed_inputs_bwd(const void* dx, float* dpos, float* dtype0, int rows, int T, int H, void* stream);

/* ------------------------------------------------------------------ small elementwise / reductions */
int spmm_colsum_bf16(const void* x, int ld, float* out, int rows, int cols, void* stream); /* out[c] += sum_r x[r][c] */
int spmm_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
int spmm_add_bf16(void* dst, const void* src, int64_t n, void* stream); /* dst += src */
int spmm_dgelu_bf16(const void* d_act, const void* pre_act, void* d_pre, int64_t n, void* stream); /* d_pre = d_act * gelu'(pre) */
int spmm_gather_rows_bf16(const void* src, const int* idx, void* dst, int n_idx, int64_t row_elems, void* stream);
/* dst[t] = sum of src[r] over r with idx[r] == t, t < n_dst (fp32 accumulation in ascending r, written once; rows no r
 * maps to become zero); n_idx <= 4096.  Autograd of the gathers / shared K/V of SPMM_models.py:165-198. */
int spmm_segment_sum_rows_bf16(void* dst, int n_dst, const int* idx, const void* src, int n_idx, int64_t row_elems,
                               void* stream);

/* ------------------------------------------------------------------ ITC / SPC head (SPMM_models.py:92-131)
 * feats are raw projection outputs z[B,E] (fp32); the kernel L2-normalises (F.normalize, :92,95,101,105),
 * scans [own momentum feats | queue] twice on the tensor cores (tcgen05 kind::tf32 straight from the fp32 queue,
 * fp32 accumulate: row LSEs, then O = softmax(S) K, from which loss and gradients follow) without materialising
 * the 8 similarity blocks, and returns
 * loss_ita, d loss/d z_prop, d loss/d z_text, d loss/d temp, plus the in-batch student sims
 * sim_i2t[B,B], sim_t2i[B,B] (for hard negatives, :157-158) and the normalised momentum feats (for enqueue).
 * Queues are stored key-major [Q][E] (the transpose of the reference's [E][Q] buffers).
 * alpha_dev (device float, optional) overrides `alpha`: the distillation weight ramps every batch of epoch 0
 * (SPMM_models.py:355), and a device scalar lets ONE captured CUDA graph of the step serve every value. */
int spmm_itc_fwd_bwd(const float* z_prop, const float* z_text, const float* z_prop_m, const float* z_text_m,
                     const float* prop_queue, const float* text_queue, const float* temp, float alpha,
                     const float* alpha_dev, int B, int E, int Q, float* loss, float* dz_prop, float* dz_text, float* dtemp, float* sim_i2t, float* sim_t2i,
                     float* feat_prop_m, float* feat_text_m, float* nan_flag, void* workspace, int64_t workspace_bytes,
                     void* stream);
int64_t spmm_itc_workspace_bytes(int B, int E, int Q);
int spmm_itc_debug_trace(void* buf);   /* debug: 64 x u64 %globaltimer stamps of one CTA of the pass-1 kernel */

/* hard-negative sampling (SPMM_models.py:154-178): w = softmax(sim[:, :B]) with zero diagonal,
 * idx[b] = argmax_j w[b][j] / Exp(1)  (ATen multinomial n=1 formulation) with a counter-based generator:
 * u = philox4x32-10(key=seed, counter=(b, j, stream_id, step)).  CPU replica: oracle/sampler_ref.py. */
int spmm_sample_negatives(const float* sim_i2t, const float* sim_t2i, int B, unsigned long long seed,
                          unsigned long long step, int* neg_t2i, int* neg_i2t, void* stream);

/* enqueue (SPMM_models.py:271-286): feats[n][E] (all ranks, gathered) -> queue rows [ptr, ptr+n), ptr advanced on device */
int spmm_enqueue(float* prop_queue, float* text_queue, const float* prop_feats, const float* text_feats,
                 int64_t* queue_ptr, int n, int E, int Q, const float* skip_flag, void* stream);

/* ------------------------------------------------------------------ losses
 * LM: loss = (1-alpha) * CE(logits[:, :-1], ids[:, 1:]) (mean over ALL positions, PAD labels included)
 *          + alpha * mean_{label != 0} -sum softmax(teacher) * log_softmax(student)   (SPMM_models.py:233-238)
 * logits rows = b*L + t (ld elements, V valid); writes dlogits (bf16, zero for t = L-1).
 * alpha_dev (device float, optional) overrides alpha.  valid_len (device int, optional) = the batch's own padded
 * width Lv <= L when rows are stored in a longer length bucket: positions t >= Lv-1 are ignored (zero gradient, not
 * counted), so the result equals the reference's on the [B, Lv] batch (padding='longest', SPMM_models.py:352). */
int spmm_lm_loss_fwd_bwd(const void* logits, const void* teacher_logits, int ld, const int64_t* ids, int B, int L,
                         int V, float alpha, const float* alpha_dev, const int* valid_len, float* loss, void* dlogits,
                         float* workspace, void* stream);
/* ITM (SPMM_models.py:201-206): logits = x[3B,2H] . W^T + b, CE with labels [1]*B + [0]*2B.
 * dw / db are accumulated into (+=); workspace: >= 3 * n_rows floats. */
int spmm_itm_loss_fwd_bwd(const void* x, const float* w, const float* b, int n_rows, int n_pos, int D, float* loss,
                          void* dx, float* dw, float* db, float* workspace, void* stream);
/* MPM tail (SPMM_models.py:251-254): pred = t . w + b over rows (b, j<n_prop); masked MSE * 5. */
int spmm_mpm_loss_fwd_bwd(const void* t, const float* w, const float* b, const float* pv, const float* mpm_mask,
                          int batch, int n_prop, int H, float* loss, void* dt, float* dw, float* db, float* workspace,
                          void* stream);

/* ------------------------------------------------------------------ parameter arenas
 * EMA (SPMM_models.py:265-269): p_m = p_m * m + p * (1 - m), three separately rounded fp32 ops (bit-exact with
 * the reference); optionally emits the bf16 shadows of p and p_m used by the GEMMs. */
int spmm_ema_multi(const float* p, float* p_m, void* p_bf16, void* p_m_bf16, int64_t n, float momentum,
                   float one_minus_momentum, void* stream);
/* clip_grad_norm_(5.) + AdamW (SPMM_models.py:340,361-362).  sumsq_out[0] receives sum g^2, added in a fixed order
 * (bit-identical on every data-parallel replica).  workspace: >= 1024 floats, zero before first use (left zeroed). */
int spmm_grad_sumsq(const float* g, int64_t n, float* sumsq_out, float* workspace, void* stream);
/* t_dev += 1 (unless *skip_flag != 0) and hyper_dev = {*lr_dev, 1-beta1^t, sqrt(1-beta2^t)} computed on the device */
int spmm_adam_tick(long long* t_dev, const float* lr_dev, float* hyper_dev, float beta1, float beta2,
                   const float* skip_flag, void* stream);
int spmm_adamw_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, const float* sumsq, float max_norm,
                    float grad_scale, const float* skip_flag, const float* hyper_dev, void* stream);
/* hyper_dev (device float[3] = lr, 1-beta1^t, sqrt(1-beta2^t)) overrides lr/step when non-NULL (graph replay). */

/* ------------------------------------------------------------------ single-token decode (PV -> SMILES beam search)
 * Replaces the per-token full-prefix re-run of d_pv2smiles_single.py:26-36 / d_pv2smiles_batched.py:24-59: the stack is
 * causal, so keys / values of earlier positions are cached.  `t_dev` (device int) = position of the token being fed;
 * every per-step scalar is read from device memory so one CUDA graph replays for all steps.
 *   decode_embed:      x[r] = word[ids[r]] + type0 + pos[*t_dev]   (bf16 pre-LayerNorm sum, xbert.py:193-217)
 *   decode_attn_self:  appends k_new/v_new (bf16 [rows][ldkv]) to cache_k/cache_v (bf16 [tmax][rows][heads*64]) at
 *                      position t, then softmax(q K^T scale) V over positions 0..t of the row's own beam: position j < t
 *                      lives in physical row anc[r][j] (beam re-ranking never moves cache data); keys with
 *                      tokens[r][j] == 0 are masked (text_atts = where(text == 0, 0, 1), d_pv2smiles_single.py:27)
 *   decode_attn_cross: q against rows (r / group) * Tk .. + Tk of k / v (bf16 [groups*Tk][ldkv]): the property tokens
 *                      of the row's molecule, projected once and shared by its `group` beams
 *   beam_step:         per molecule (k beams = rows m*k ..): softmax + top-k of each beam's logits, candidates
 *                      score + log p; from the 2nd expansion on a candidate ending in sep_id moves to the finished
 *                      list (score, tokens) and is masked with -1e5, the molecule is done once >= k are finished;
 *                      otherwise top-k of the k*k candidates become the new beams (tokens, ancestor rows, scores,
 *                      next_ids rewritten in place); finally *t_dev += 1.  fin_cap >= k*k + k.  trace_* optional
 *                      [steps][n_mol*k][k] record of each beam's top-k (log p, token).  ticket: zeroed uint32. */
int spmm_decode_embed(const int64_t* ids, const int* t_dev, const float* word, const float* pos, const float* type0,
                      void* x, int rows, int H, void* stream);
int spmm_decode_attn_self(const void* q, int ldq, const void* k_new, const void* v_new, int ldkv, void* cache_k,
                          void* cache_v, const int* anc, const int64_t* tokens, int tmax, const int* t_dev, void* out,
                          int ldo, int rows, int heads, float scale, void* stream);
int spmm_decode_attn_cross(const void* q, int ldq, const void* k, const void* v, int ldkv, int Tk, int group,
                           const int* kv_len, void* out, int ldo, int rows, int heads, float scale, void* stream);
int spmm_beam_step(const void* logits, int ld, int V, int k, int tmax, int n_mol, int fin_cap, int cls_id, int sep_id,
                   int* t_dev, float* scores, int64_t* tokens, int* anc, int64_t* next_ids, float* fin_scores,
                   int64_t* fin_tokens, int* fin_len, int* fin_count, int* done, float* trace_logp, int* trace_tok,
                   unsigned int* ticket, void* stream);

/* ------------------------------------------------------------------ host-side WordPiece tokenizer (no CUDA)
 * Replaces BertTokenizer(do_basic_tokenize=False) + WordpieceTokenizer(max_input_chars_per_word=250) as used by
 * SPMM_pretrain.py:19-20 / SPMM_models.py:352 / d_smiles2pv.py:43,61: whitespace split, greedy longest-match-first
 * ("##" continuation pieces), whole word -> unk on any miss, [cls] ... [sep], truncation to max_length,
 * padding='longest'.  ids_out / mask_out: [n][ld] int64 host buffers (pinned for async H2D).  Returns the padded width. */
void* spmm_wordpiece_create(const char* const* tokens, int n_tokens, int unk_id, int max_input_chars_per_word);
void spmm_wordpiece_destroy(void* handle);
int spmm_wordpiece_encode_batch(void* handle, const char* const* texts, int n, int max_length, int cls_id, int sep_id,
                                int pad_id, int64_t* ids_out, int64_t* mask_out, int ld);

#ifdef __cplusplus
}
#endif
#endif /* SPMM_B200_H_ */

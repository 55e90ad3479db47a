/* libdmi_b200 -- C ABI of the B200-native adapted-projector hot path.
 *
 * The reference (ospanbatyr/sample-efficient-multimodality) is pure Python/PyTorch and has no FFI: its boundary for
 * this path is the nn.Module surface of dmi/model/ (SURVEY.md section 8b).  This header is what a ctypes binding on the
 * reference side would load; each entry point names the reference code it replaces (file:line under the reference
 * root).  Conventions:
 *   - every pointer is a DEVICE pointer unless the parameter is documented as host memory; row-major, contiguous rows;
 *   - no torch types, no allocation, no synchronisation: the caller owns all buffers and supplies the CUDA stream
 *     (cudaStream_t passed as void*); kernels are enqueued and the call returns;
 *   - return value 0 on success, negative dmi_status on failure; dmi_last_error() describes the last failure of the
 *     calling thread; nothing throws across the boundary;
 *   - "bf16" buffers are uint16_t-sized __nv_bfloat16; 16-byte alignment is required for every base pointer and every
 *     leading dimension must be a multiple of 8 elements.
 */
#ifndef DMI_B200_H
#define DMI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum dmi_status {
  DMI_STATUS_OK = 0,
  DMI_STATUS_INVALID = -1,
  DMI_STATUS_CUDA = -2,
  DMI_STATUS_UNSUPPORTED = -3
} dmi_status;

/* library introspection */
int dmi_version(void);                 /* major*10000 + minor*100 + patch */
const char* dmi_last_error(void);      /* thread-local, never NULL */
int dmi_num_sms(void);                 /* SM count of the current device (148 on B200) */
int dmi_set_option(const char* name, int value);   /* tuning switches for A/B measurements: "gemm_pair" = -1 auto | 0 off | 1 on (CTA-pair GEMM), "fused_panel" = -1 auto | 0 separate mma.sync side passes | 1 tcgen05 panel passes at any batch size, "pdl" = 1 | 0 (programmatic dependent launch of every kernel, on by default), "gemm_debug" (measurement only) */
int64_t dmi_launch_count(void);        /* number of kernels this library has launched in this process (bench.py gpu_launches) */

/* ---------------------------------------------------------------------------------------------------------------
 * Building block: C[M,N] = alpha * A[M,K] * B[N,K]^T (+ bias[N]) with a fused epilogue, tcgen05 + TMA + TMEM.
 * kind: 0 = bf16 operands, 1 = tf32 (A and B are float).  mode: 0 = store, 1 = GELU(tanh) (out0 = act, out1 = pre),
 * 2 = out0 = acc * gelu'(aux).  Replaces the F.linear / bmm call sites listed in SURVEY.md section 2a (k2, k5, k8, k9).
 * ------------------------------------------------------------------------------------------------------------- */
int dmi_gemm_tn(int kind, int mode, const void* A, int64_t lda, const void* B, int64_t ldb,
                int64_t M, int64_t N, int64_t K, float alpha, const float* bias,
                void* out0, int64_t ld0, int out0_is_f32, void* out1_bf16, int64_t ld1,
                const void* aux_bf16, int64_t ld_aux, void* stream);

/* C[M,N] (+)= alpha * A[K,M]^T B[K,N]: bf16 operands that are both contracted over their ROWS (MN-major UMMA descriptors),
 * fp32 output.  The weight-gradient GEMMs dW2 = dY^T h, dW1 = dpre^T x (autograd of projector.py:56-59), K = batch. */
int dmi_gemm_mn(const void* A_bf16, int64_t lda, const void* B_bf16, int64_t ldb, int64_t M, int64_t N, int64_t K, float alpha,
                float* out, int64_t ldo, int accumulate, void* stream);

/* out[M,R] = in[M,K] * W[R,K]^T with R = adapter rank (8/16/32/64), bf16 out.  in_is_f32 != 0: `in` is fp32 and is converted on
 * the fly, its bf16 copy written to copy_bf16 (may be NULL) -- the fused "convert + rank-r projection" pass over x and dY
 * (u = x A0, dv = dY B1^T; the per-sample bmm pair of projector.py:149-152 and its autograd). */
int dmi_skinny_rows(const void* in, int64_t ld_in, int in_is_f32, const void* W_bf16, int64_t ldw, void* out_bf16, int64_t ld_out,
                    void* copy_bf16, int64_t ld_copy, int64_t M, int64_t K, int64_t R, void* stream);

/* One fused tcgen05 pass over a bf16 activation gradient `in` [M,K] (R = 32, K = 1024 or 2048, out 16-byte aligned, ld_out % 8 == 0):
 *   out[M,R] = in W[R,K]^T,   G[R,K] += scale * L[M,R]^T in,   colsum[K] += scale * 1^T in   (colsum may be NULL)
 * = dmi_skinny_rows followed by dmi_outer_reduce over the same matrix in ONE HBM sweep: (du, dB0, dbeta0) from dpre -- the autograd
 * of the bmm pair and bias add of projector.py:146-157.  A 2-CTA cluster per 128-row panel, accumulators in TMEM, each TMA-staged
 * tile read by the tensor core as the K-major operand of the projection and as the MN-major operand of the batch reduction; the
 * column sum rides in the batch-reduction MMAs (a ones column written into the staged L panel). */
int dmi_panel_fused_tc(const void* in_bf16, int64_t ld_in, const void* W_bf16, int64_t ldw, void* out_bf16, int64_t ld_out,
                       const void* L_bf16, int64_t ldl, float* G, int64_t ldg, float* colsum, float scale, int64_t M, int64_t K,
                       int64_t R, void* stream);

/* The projection alone on the same kernel (v = h A1, u = x A0: out[M,R] = in W[R,K]^T; bf16, R = 32, K = 768 / 1024 / 2048,
 * out 16-byte aligned with ld_out % 8 == 0).  Same contract as dmi_skinny_rows for those shapes. */
int dmi_panel_tc_project(const void* in_bf16, int64_t ld_in, const void* W_bf16, int64_t ldw, void* out_bf16, int64_t ld_out, int64_t M,
                         int64_t K, int64_t R, void* stream);

/* fp32-input form of dmi_panel_fused_tc (the dY pass: (dv, dB1, dbeta1) from dY, and the bf16 copy of dY the dpre GEMM consumes;
 * copy may be NULL).  Same shapes and alignment as dmi_panel_fused_tc, copy 16-byte aligned with ld_copy % 8 == 0. */
int dmi_panel_fused_tc32(const float* in, int64_t ld_in, const void* W_bf16, int64_t ldw, void* out_bf16, int64_t ld_out, void* copy_bf16,
                         int64_t ld_copy, const void* L_bf16, int64_t ldl, float* G, int64_t ldg, float* colsum, float scale, int64_t M,
                         int64_t K, int64_t R, void* stream);

/* G[P,Q] += scale * L[B,P]^T R[B,Q] (bf16 in, fp32 atomic accumulate; optional colsum[Q] += scale * 1^T R).
 * The batch contraction behind dA/dB/dbeta of the adapter (autograd of projector.py:146-157 in the reference). */
int dmi_outer_reduce(const void* L_bf16, int64_t ldl, const void* R_bf16, int64_t ldr, int64_t B, int64_t P, int64_t Q,
                     float* G, int64_t ldg, int transpose_out, float* colsum, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Adapted MLP2 projector  (reference: dmi/model/projector.py:56-59 forward, :61-74 only_lora_forward,
 * :118-159 lora_forward, and their autograd).   y = gelu(x W1^T + b1 + (x A0) B0 + beta0) W2^T + b2 + (h A1) B1 + beta1
 * The adapter is folded into the GEMMs as r extra K columns: xext = [x | xA0], w1ext = [W1 | B0^T], ...
 * ------------------------------------------------------------------------------------------------------------- */
#define DMI_MLP_STOP_AFTER_FIRST_ACT 1   /* reproduce lora_forward exactly as written: y = gelu(pre) */
#define DMI_MLP_NO_ADAPTER 2             /* plain / merged projector: r columns absent */
#define DMI_MLP_X_PREPACKED 4            /* xext[:, :D] already holds bf16 x (written by dmi_augment) */
#define DMI_MLP_BASE_GRADS 8             /* also produce dW1,db1,dW2,db2 (train_projector / few-shot fine-tune) */
#define DMI_MLP_DROPOUT 16               /* h <- h * keep / (1-p) with a caller-provided keep mask (train_projector) */

typedef struct dmi_mlp_args {
  int64_t B, D, H, r;          /* batch rows, projector input width, LM hidden, adapter rank (0 with NO_ADAPTER) */
  int32_t flags;
  float grad_scale;            /* multiplies every gradient written by bwd (1/GA, LoRA alpha/r ...) */
  float dropout_p;
  int32_t _pad;
  /* inputs */
  const float* x;   int64_t ldx;      /* [B,D] fp32 */
  const float* dy;  int64_t lddy;     /* [B,H] fp32, gradient arriving at the projector output */
  const uint8_t* keep;                /* [B,H] dropout keep mask (bytes) or NULL */
  /* packed weights (dmi_projector_pack_base / dmi_adapter_pack) */
  const void* w1ext;                  /* bf16 [H, D+r] = [W1 | B0^T] */
  const void* w2ext;                  /* bf16 [H, H+r] = [W2 | B1^T] */
  const void* w2text;                 /* bf16 [H, H+r] = [W2^T | A1] */
  const void* a0t;                    /* bf16 [r, D]  = A0^T */
  const void* a1t;                    /* bf16 [r, H]  = A1^T */
  const void* b0;                     /* bf16 [r, H] */
  const void* b1;                     /* bf16 [r, H] */
  const float* bias0;                 /* [H] = b1 + beta0 */
  const float* bias1;                 /* [H] = b2 + beta1 */
  /* activation stash, caller allocated */
  void* xext;                         /* bf16 [B, D+r] */
  void* pre;                          /* bf16 [B, H]   */
  void* hext;                         /* bf16 [B, H+r] */
  void* dyext;                        /* bf16 [B, H+r] (bwd) */
  void* dpre;                         /* bf16 [B, H]   (bwd) */
  void* du;                           /* bf16 [B, r]   (bwd) */
  /* outputs */
  float* y;  int64_t ldy;             /* [B,H] fp32 (may be NULL if y_bf16 is given) */
  void* y_bf16; int64_t ldy_bf16;     /* optional bf16 copy of y, e.g. row 0 of inputs_embeds */
  float* dA0; float* dB0; float* dbeta0;      /* [D,r] [r,H] [H]   accumulated (+=) */
  float* dA1; float* dB1; float* dbeta1;      /* [H,r] [r,H] [H]   accumulated (+=) */
  float* dW1; float* db1; float* dW2; float* db2;   /* BASE_GRADS: [H,D] [H] [H,H] [H] accumulated (+=) */
  /* optional cudaEvent_t recorded by bwd on `stream` as soon as the layer-1 gradients (dA1,dB1,dbeta1 / dW2,db2) are
   * enqueued, so a data-parallel caller can start all-reducing that bucket while the layer-0 backward still runs */
  void* ev_layer1_grads;
} dmi_mlp_args;

/* W1 [H,ldw1>=D] , W2 [H,H] fp32 -> base columns of w1ext / w2ext / w2text (done once per frozen projector). */
int dmi_projector_pack_base(const float* W1, int64_t ldw1, const float* W2, int64_t D, int64_t H, int64_t r,
                            void* w1ext, void* w2ext, void* w2text, void* stream);
/* flat fp32 adapter (A0 [D,r], B0 [r,H], beta0 [H], A1 [H,r], B1 [r,H], beta1 [H]; as HyperNetwork.forward returns
 * them, hypernet.py:181-194) -> low-rank columns of the ext matrices, transposed bf16 copies, fused biases.
 * scale multiplies B0/B1 (LoRALayer alpha/r, lora.py:16); beta pointers may be NULL. */
int dmi_adapter_pack(const float* A0, const float* B0, const float* beta0, const float* A1, const float* B1,
                     const float* beta1, const float* b1, const float* b2, int64_t D, int64_t H, int64_t r, float scale,
                     void* w1ext, void* w2ext, void* w2text, void* a0t, void* a1t, void* b0, void* b1_bf16,
                     float* bias0, float* bias1, void* stream);

/* Adapter merge, exact fp32 (reference Projector.combine_lora, projector.py:95-103):
 * W_out[o,i] = W[o,i] + scale * sum_j A[i,j] B[j,o],  bias_out = bias + beta (beta may be NULL).  W: [H, in_dim] row stride ldw. */
int dmi_merge_adapter(const float* W, int64_t ldw, const float* bias, const float* A, const float* B, const float* beta,
                      int64_t in_dim, int64_t H, int64_t r, float scale, float* W_out, int64_t ldwo, float* bias_out, void* stream);

int dmi_adapted_mlp_fwd(const dmi_mlp_args* args, void* stream);
int dmi_adapted_mlp_bwd(const dmi_mlp_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a1: row-wise L2 normalisation x / ||x||_2  (EmbeddingManager.get_embeddings, dmi/utils/model_utils.py:47-62)
 * ------------------------------------------------------------------------------------------------------------- */
int dmi_l2_normalize(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* out, int64_t ldo, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a1-a3: augmentation + support-set assembly (HypernetTrainer._process_embeddings / _interleave_embeddings,
 * dmi/train_hypernet.py:76-108; the rotation matrix of :56-57 is an INPUT).  For each input row: optional L2
 * normalisation, optional column gather (perm; the InfFS feature selection of data/base.py:222-225 is such a gather) and
 * sign flip, then x' = x R on the tf32 tensor cores with a 3-term split (fp32-accurate).  Outputs: mm_out [B,D] (+ bf16
 * copy straight into the projector's operand buffer), and z [1+2K, Dh] = [prefix; m'_0; t_0; m'_1; t_1; ...] with the
 * rotated support rows zero-padded from D to Dh columns (pruned projector).  R == NULL: no rotation (eval / few-shot).
 * ------------------------------------------------------------------------------------------------------------- */
#define DMI_AUG_NORMALIZE 1
typedef struct dmi_augment_args {
  int64_t B, K, D, Dh, D_src;
  int32_t flags; int32_t _pad;
  const float* mm;  int64_t ld_mm;      /* [B, D_src] batch embeddings */
  const float* sup; int64_t ld_sup;     /* [K, D_src] support-set modality embeddings */
  const float* txt; int64_t ld_txt;     /* [K, Dh] support-set text embeddings (never rotated) */
  const float* prefix;                  /* [1, Dh] instruction-prefix embedding (never rotated), may be NULL */
  const float* R;                       /* [D, D] row-major orthogonal matrix or NULL */
  const int32_t* perm;                  /* [D] source column of each output column, or NULL */
  const float* sign;                    /* [D] +-1 or NULL */
  float* mm_out; int64_t ld_mm_out;     /* [B, D] fp32 (may be NULL if mm_out_bf16 is given) */
  void* mm_out_bf16; int64_t ld_mm_bf16;/* optional bf16 copy, e.g. columns [0,D) of xext */
  float* z;                             /* [1+2K, Dh] contiguous */
  void* workspace; uint64_t workspace_bytes;
} dmi_augment_args;
int64_t dmi_augment_workspace_bytes(int64_t B, int64_t K, int64_t D);
int dmi_augment(const dmi_augment_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a4-a6 + autograd: HyperNetwork.forward with hn_arch="attention", one head (dmi/model/hypernet.py:46-82, 140-196):
 * sequence [prefix_tokens; z] + positional encoding, softmax(q k^T / sqrt(D)) v for the NQ prefix rows only, then per
 * layer w_l = out_scale * (G_l e_l + c_l).  The attention is evaluated in its reduced form (no K/V projection of the S
 * tokens is materialised): q~ = Wk^T q, scores = s.q~ + q.bk, e = Wv (P s) + bv sum(P).
 * Key masking of the reference (sequence shorter than the context, hypernet.py:144-151) = passing only the valid rows.
 * ------------------------------------------------------------------------------------------------------------- */
#define DMI_MAX_GEN_LAYERS 4
typedef struct dmi_hypernet_args {
  int64_t S_z, NQ, D, n_layers;         /* rows of z, prefix tokens, hypnet_dim, generators (<= NQ) */
  float out_scale;                      /* alpha / rank */
  float dropout_p;                      /* attention dropout probability, used with keep */
  int32_t overwrite_gen_grads;          /* bwd: 1 = dgen_w/dgen_b are overwritten, 0 = accumulated into (+=) */
  int32_t _pad;
  const float* z; int64_t ldz;          /* [S_z, D] */
  const float* prefix_tokens;           /* [NQ, D] */
  const float* pe; int64_t ldpe;        /* [>= NQ+S_z, D] positional encodings (already scaled) or NULL */
  const float *wq, *bq, *wk, *bk, *wv, *bv;         /* [D,D] / [D] */
  const float* gen_w[DMI_MAX_GEN_LAYERS];           /* [gen_out[l], D] */
  const float* gen_b[DMI_MAX_GEN_LAYERS];           /* [gen_out[l]] */
  int64_t gen_out[DMI_MAX_GEN_LAYERS];
  const float* keep;                    /* [NQ, NQ+S_z] attention-dropout keep mask (0/1) or NULL (eval) */
  float* w_out[DMI_MAX_GEN_LAYERS];     /* fwd out: flat generated weights [gen_out[l]] = A_l || B_l || beta_l */
  float* stash;                         /* dmi_hypernet_stash_floats() floats, written by fwd, read by bwd */
  /* backward */
  const float* dw[DMI_MAX_GEN_LAYERS];  /* gradient wrt w_out[l], or NULL when layer l carries no gradient */
  float* scratch;                       /* dmi_hypernet_scratch_floats(NQ, S_z, D) floats */
  float *dprefix, *dwq, *dbq, *dwk, *dbk, *dwv, *dbv;      /* accumulated (+=) */
  float* dgen_w[DMI_MAX_GEN_LAYERS];    /* [gen_out[l], D]; NULL = do not materialise the rank-1 gradient (the kernel then only reads G: de = G^T dw) */
  float* dgen_b[DMI_MAX_GEN_LAYERS];    /* [gen_out[l]] (may be NULL together with dgen_w[l]) */
} dmi_hypernet_args;
int64_t dmi_hypernet_stash_floats(int64_t NQ, int64_t S_z, int64_t D);
int64_t dmi_hypernet_scratch_floats(int64_t NQ, int64_t S_z, int64_t D);
/* offset (floats) inside `stash` of the modality codes e [NQ, D] written by fwd / pool: the second factor of the rank-1 generator
 * gradient dG_l = (out_scale * dw_l) (x) e_l when bwd is called with dgen_w[l] == NULL (factors kept instead of a dense gradient) */
int64_t dmi_hypernet_stash_code_offset(int64_t NQ, int64_t S_z, int64_t D);
int dmi_hypernet_fwd(const dmi_hypernet_args* args, void* stream);
int dmi_hypernet_bwd(const dmi_hypernet_args* args, void* stream);
/* Few-shot adapter pipeline.  The generators are LINEAR in the modality codes e_l, so the element-wise mean of N
 * generated adapters (HyperNetWrapper.generate_projector_from_multiple_adapters, dmi/model/hypernet.py:234-266) equals one generator
 * pass over the mean code: the 692 MB of generator weights are streamed once instead of N times.
 *   dmi_hypernet_pool     : a4-a5 only (fill z / prefix_tokens / pe / q,k,v / stash as for dmi_hypernet_fwd); e_accum[NQ,D] += weight * e
 *   dmi_hypernet_generate : w_out[l] = out_scale * (G_l e[l,:] + c_l) for the n_layers generators; e is [>= n_layers, D] */
int dmi_hypernet_pool(const dmi_hypernet_args* args, float* e_accum, float weight, void* stream);
int dmi_hypernet_generate(const dmi_hypernet_args* args, const float* e, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Section 8f-3: Haar-random orthogonal matrix on the device, replacing the host draw of the isometry
 *   R = torch.FloatTensor(scipy.stats.ortho_group.rvs(mm_dim)).to(device)            (dmi/train_hypernet.py:56-57)
 * (LAPACK QR of a Gaussian matrix + sign fix; 72 ms at n=768 per micro-step).  `gauss` is an [n,n] fp32 matrix of i.i.d. N(0,1)
 * samples (row k supplies the n-k entries of the k-th reflector; its first k entries are ignored).  Q = H_0 ... H_{n-1} diag(d)
 * is Haar distributed (Stewart 1980 / Mezzadri 2007: the reflectors of the QR of a Gaussian matrix are independent Gaussian
 * directions); applied in compact-WY blocks of 64 reflectors, fp32 throughout.  Bit parity with LAPACK's draw is impossible by
 * construction -- this is the "statistically equivalent" mode, with its own orthogonality / distribution tests.
 * ------------------------------------------------------------------------------------------------------------- */
int64_t dmi_haar_workspace_bytes(int64_t n);
int dmi_haar_orthogonal(const float* gauss, int64_t n, float* Q_out, void* workspace, uint64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Section 8f-4: embedding-store gather.  The embedding side of the reference's collate functions and of
 * EmbeddingManager.get_embeddings in one pass over a flat device-resident table `store` [n_store_rows, d_store] (fp32 or bf16):
 *   out[b, j] = norm( store[idx[b], sel[j]] - mean[j] )
 * = torch.FloatTensor(item['emb'])[selected_features], torch.stack, `- emb_mean` (dmi/data/base.py:222-232, :238-250, :257-268), `.to(device)`
 * and `/ embs.norm(dim=1, keepdim=True)` with DMI_AUG_NORMALIZE (dmi/utils/model_utils.py:47-62).  idx == NULL: rows 0..B-1;
 * selected_features / mean may be NULL; an index outside [0, n_store_rows) sets *error_flag to 1 and fills that output row with NaN (the reference raises an IndexError on the host).
 * ------------------------------------------------------------------------------------------------------------- */
int dmi_gather_rows(const void* store, int store_is_bf16, int64_t ld_store, int64_t n_store_rows, int64_t d_store, const int64_t* idx,
                    int64_t B, int64_t d_out, const int32_t* selected_features, const float* mean, int flags, float* out, int64_t ldo,
                    void* out_bf16, int64_t ldo_bf16, int* error_flag, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a12: prefix splice (HypernetMMModel.forward, dmi/model/mmmodel.py:36-48; same block at :118-135 and :205-221):
 * out[b,0,:] = projected[b,:], out[b,1+t,:] = table[ids[b,t],:]; labels_out = [-100, labels]; mask_out = [1, mask].
 * out is fp32 (the reference's torch.cat promotion) or bf16; ids outside [0,vocab) set *error_flag to 1.
 * ------------------------------------------------------------------------------------------------------------- */
int dmi_splice(const float* proj_f32, const void* proj_bf16, int64_t ld_proj, const void* table, int table_is_bf16,
               int64_t ld_table, int64_t vocab, const int64_t* ids, int64_t B, int64_t T, int64_t H, void* out, int out_is_bf16,
               const int64_t* labels, int64_t* labels_out, const void* mask, int mask_is_i64, float* mask_out,
               int* error_flag, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * a13 / section 8f-1: optimizer step of the reference trainers -- torch.nn.utils.clip_grad_norm_(params, max_norm) followed by
 * optim.AdamW.step()  (dmi/train_hypernet.py:148-149, optimizer built at :526-532; same pair in train_projector.py / train_lora.py).
 * `tensors` is a HOST array of `count` descriptors of fp32 device tensors (parameter, gradient, exp_avg, exp_avg_sq, element count);
 * parameters without a gradient are simply not listed (AdamW skips them, weight decay included).
 *   dmi_grad_sqnorm : *sqnorm_accum += sum_i g_i^2 over all listed tensors (zero it first; all-reduce it for sharded parameters)
 *   dmi_grad_clip   : g *= min(1, max_norm / (sqrt(*sqnorm) + 1e-6))            -- clip_grad_norm_ alone
 *   dmi_adamw_step  : one pass: g' = clip(g), p *= 1 - lr*wd, m += (g'-m)(1-b1), v = b2 v + (1-b2) g'^2,
 *                     p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps);  max_grad_norm <= 0 disables clipping;
 *                     write_clipped_grads != 0 also stores g' (the in-place effect of clip_grad_norm_).  step counts from 1.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct dmi_opt_tensor {
  float* p; float* g; float* m; float* v;
  int64_t n;
} dmi_opt_tensor;
int dmi_grad_sqnorm(const dmi_opt_tensor* tensors, int count, float* sqnorm_accum, void* stream);
int dmi_grad_clip(const dmi_opt_tensor* tensors, int count, float max_norm, const float* sqnorm, void* stream);
int dmi_adamw_step(const dmi_opt_tensor* tensors, int count, double lr, double beta1, double beta2, double eps, double weight_decay,
                   int64_t step, float max_grad_norm, const float* sqnorm, int write_clipped_grads, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Data parallelism (the reference's analogue is gradient accumulation, train_hypernet.py:119-149): one-shot
 * all-reduce (SUM, then * scale) of `n` floats over the `world` <= 8 GPUs of one NVSwitch domain through peer-mapped memory, enqueued
 * IN the step's stream (no side stream, no NCCL kernel competing with the persistent GEMMs for SMs).
 *   peer_bufs  : HOST array of `world` device pointers -- every rank's input buffer (peer-mapped, e.g. torch symmetric memory), same layout
 *   peer_flags : HOST array of `world` device pointers to each rank's flag buffer, dmi_allreduce_flag_words() uint32, zeroed once
 *   multicast_ptr : NVLS multicast address of the input buffer (in-switch reduction with multimem.ld_reduce) or NULL (peer loads)
 *   out        : LOCAL fp32 output, distinct from the input;  epoch : 1, 2, 3, ... the number of this call, identical on all ranks.
 * Every rank must enqueue the call; the kernels of the ranks wait for one another (bounded spin, traps after ~2 s).
 * ------------------------------------------------------------------------------------------------------------- */
int64_t dmi_allreduce_flag_words(void);
int dmi_allreduce_oneshot(const void* const* peer_bufs, void* const* peer_flags, const void* multicast_ptr, int rank, int world, float* out,
                          int64_t n, float scale, uint32_t epoch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMI_B200_H */

/* pevit_b200 -- C ABI of the B200-native PEViT fine-tuning hot path.
 *
 * The reference (eric-ai-lab/PEViT) has no plugin / FFI layer: its hot path is the Python
 * nn.Module code in vision_benchmark/evaluation/{model,lora_model,adapter_model,
 * compacter_model}.py.  This header is therefore the boundary a maintainer binds instead
 * (ctypes stub shown in INTEGRATION.md); each entry point names the reference code it replaces.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer borrowed from the caller
 *    (torch `tensor.data_ptr()`), alive until the stream reaches the call; `stream` is a
 *    `cudaStream_t` passed as void* (torch.cuda.current_stream().cuda_stream).
 *  - stream-ordered, no implicit synchronisation, no allocation per call; scratch and saved
 *    activations are caller-allocated with the sizes the *_bytes() queries return.
 *  - return 0 on success, < 0 on error; the message is in pevit_last_error() (thread-local).
 *    Nothing throws or exits across the ABI.  There is no CPU fallback.
 *  - token rows are in the reference's (L, N, D) order: row = l * NB + n  (model.py:1042).
 *  - "bf16" buffers are raw uint16 storage (torch.bfloat16).
 */
#ifndef PEVIT_B200_H_
#define PEVIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEVIT_ABI_VERSION 2

enum pevit_method {
  PEVIT_PLAIN = 0,       /* stock block, no PEFT module (text tower / ablation) */
  PEVIT_KADAPTATION = 1, /* model.py:423-940    Kronecker delta on q, v          */
  PEVIT_LORA = 2,        /* lora_model.py:423-  rank-4 delta on q, v             */
  PEVIT_ADAPTER = 3,     /* adapter_model.py:204-336  bottleneck after the MLP   */
  PEVIT_COMPACTER = 4    /* compacter_model.py:196-503  PHM bottleneck           */
};

enum pevit_gemm_epilogue {
  PEVIT_EPI_F32 = 0,    /* out_f32 = acc + bias + resid                          */
  PEVIT_EPI_BF16 = 1,   /* out_bf16 = acc + bias                                 */
  PEVIT_EPI_QGELU = 2,  /* z = acc + bias; out = z*sigmoid(1.702 z); out2 = z    */
  PEVIT_EPI_DQGELU = 3, /* out = acc * d/dz quickgelu(aux)                       */
  PEVIT_EPI_QKV = 4     /* head-major q/8,k,v scatter + low-rank T columns        */
};

/* ------------------------------------------------------------------ library */
int pevit_abi_version(void);
const char* pevit_last_error(void);
/* 0 if the current device is sm_100 (B200); < 0 with a message otherwise. */
int pevit_check_device(void);

/* Launch accounting: every kernel launch of this library is counted; with profiling enabled each
 * launch is bracketed by CUDA events on its stream and attributed to a kernel class. */
int pevit_prof_enable(int32_t on);
int pevit_prof_reset(void);
int pevit_prof_num_classes(void);
const char* pevit_prof_class_name(int32_t cls);
int pevit_prof_read(double* ms, int64_t* launches, int32_t n); /* waits for the events, then clears */
int64_t pevit_launch_count(void);                              /* launches since the library was loaded */

/* ------------------------------------------------------------------ primitives
 * (exported so every kernel can be parity-tested on its own) */

/* C = A[M,K] * B[N,K]^T, bf16 operands, fp32 accumulation in TMEM (tcgen05).
 * Replaces F.linear at model.py:305, :816, :958-962 and their autograd dgrads. */
typedef struct pevit_gemm_args {
  const void* a; int32_t lda;        /* bf16 [M][lda]  */
  const void* b; int32_t ldb;        /* bf16 [N][ldb]  */
  int32_t m, n, k;
  int32_t epilogue;                  /* enum pevit_gemm_epilogue */
  const float* bias;                 /* [N] or NULL */
  const float* resid;                /* fp32 [M][ld_out] or NULL (EPI_F32) */
  float* out_f32;                    /* EPI_F32 */
  void* out_bf16;                    /* EPI_BF16 / QGELU / DQGELU */
  void* out2_bf16;                   /* EPI_QGELU: z (nullable) */
  const void* aux_bf16;              /* EPI_DQGELU: z */
  const void* resid_bf16;            /* EPI_BF16: bf16 [M][ld_out] added before rounding (may alias out) */
  int32_t ld_out;
  void* qkv_hm; void* t_out;         /* EPI_QKV outputs: head-major q/8|k|v, bf16 T [M][r2] */
  int32_t L, NB, H, D, r2;           /* EPI_QKV shape */
  int32_t force_bn;                  /* 0 = heuristic, else 32/64/128/256 */
} pevit_gemm_args;
int pevit_gemm_tn(const pevit_gemm_args* args, void* stream);

/* LayerNorm with fp32 statistics, eps 1e-5 (model.py:154-160). */
int pevit_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                        float* mean, float* rstd, int32_t rows, int32_t d, void* stream);
int pevit_layernorm_bwd(const float* dyn, const float* x, const float* gamma, const float* mean, const float* rstd,
                        const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta, int32_t rows,
                        int32_t d, void* stream);

/* Attention core + in-kernel low-rank delta (model.py:786-815, lora_model.py:719-733). */
typedef struct pevit_attn_args {
  int32_t L, NB, H, D, r; float alpha;
  const void *q, *k, *v;             /* bf16 head-major [NB*H][L][64], q pre-scaled */
  const void* t;                     /* bf16 [L*NB][2r] or NULL (impl 1 only: in-kernel delta) */
  const float* qmat;                 /* fp32 [2][D][r] or NULL  */
  const float* delta_bias;           /* fp32 [D] or NULL (KAdaptation attn.b) */
  void* o_tok; float* lse;           /* fwd out: bf16 [L*NB][D], fp32 [NB*H][L] */
  const void* do_tok;                /* bwd in : bf16 [L*NB][D] */
  void* dqkv; int32_t ld_dqkv;       /* bwd out: bf16 [L*NB][ld], cols dq/8 | dk | dv */
  void* ddelta;                      /* bwd out: bf16 [2][NB*H][L][64] (nullable) */
  int32_t impl;                      /* 0 = default, 1 = CUDA-core cross-check kernel, 2 = pair-streaming kernels for L > 128 */
  int32_t causal;                    /* fwd only, L <= 128: 1 = keys j > l masked out (text tower, model.py:1139-1145) */
} pevit_attn_args;
int pevit_attn_fwd(const pevit_attn_args* args, void* stream);
int pevit_attn_bwd(const pevit_attn_args* args, void* stream);

/* Factor expansion (model.py:563-580 without materialising H; lora_model.py:490-514). */
int pevit_kad_expand(const float* u1, const float* v1, const float* u2, const float* v2, const float* s,
                     const float* t, int32_t d, float alpha, void* w_ext, void* w_ext_t, float* qmat, void* qmat_t,
                     void* delta_w, void* stream);
int pevit_lora_expand(const float* aq, const float* av, const float* bq, const float* bv, int32_t d, int32_t r,
                      float alpha, void* w_ext, void* w_ext_t, float* qmat, void* qmat_t, void* delta_w, void* stream);
/* C[kc][nc] += scale * A[M][kc]^T B[M][nc] (fp32 atomics; caller zeroes C). */
int pevit_atb_accumulate(const void* a, int32_t a_is_bf16, int32_t lda, const void* b, int32_t b_is_bf16,
                         int32_t ldb, int32_t m, int32_t kc, int32_t nc, float scale, float* c, void* stream);
/* Same product on tcgen05 (bf16 operands, rows split across CTAs, fp32 red.add): b exposes nb_cols (<= 64)
 * columns, columns [n_lo, n_lo+n_cnt) of A^T B are accumulated into c[kc][0..n_cnt) (row stride ldc). */
int pevit_atb_tc(const void* a, int32_t lda, const void* b, int32_t ldb, int32_t nb_cols, int32_t m, int32_t kc,
                 int32_t n_lo, int32_t n_cnt, float scale, float* c, int32_t ldc, void* stream);
int pevit_colsum_bf16(const void* x, int32_t m, int32_t d, float* out, void* stream);
int pevit_kad_factor_grads(const float* dP, const float* dQ, const float* u1, const float* v1, const float* u2,
                           const float* v2, const float* s, const float* t, int32_t d, float* du1, float* dv1,
                           float* du2, float* dv2, float* ds, float* dt, void* stream);
/* Same contractions ADDED onto du1..dt (the caller's .grad buffers): what autograd's AccumulateGrad would do with
 * the result of pevit_kad_factor_grads, without the temporaries and the add kernels (the shared phm_rule tensors
 * receive one contribution per block, kadaptation_clip.py:352 loss.backward()). */
int pevit_kad_factor_grads_acc(const float* dP, const float* dQ, const float* u1, const float* v1, const float* u2,
                               const float* v2, const float* s, const float* t, int32_t d, float* du1, float* dv1,
                               float* du2, float* dv2, float* ds, float* dt, void* stream);
/* The same accumulation for all `count` (<= 48) layers of one backward pass in ONE launch: dP / dQ / s / t / ds / dt are
 * HOST arrays of `count` device pointers (one per layer); the rule factors u1, v1, u2, v2 and their gradient buffers
 * are the tensors all layers share (model.py:1003-1010), their per-layer contributions are added atomically. */
int pevit_kad_factor_grads_acc_batch(int32_t count, const float* const* dP, const float* const* dQ, const float* const* s,
                                     const float* const* t, float* const* ds, float* const* dt, const float* u1,
                                     const float* v1, const float* u2, const float* v2, int32_t d, float* du1, float* dv1,
                                     float* du2, float* dv2, void* stream);
/* Compacter PHM layers (compacter_model.py:196-308, :302-308 the per-call einsum): both layers of one block expanded
 * on the device into the bf16 operands of the bottleneck GEMMs -- W = H^T with H = sum_i kron(rule_i, left_i right_i).
 * rule fp32 [n][n][n]; down: left [n][d/n], right [n][bottleneck/n]; up: left [n][bottleneck/n], right [n][d/n].
 * Outputs bf16: w_down [bottleneck][d], w_down_t [d][bottleneck], w_up [d][bottleneck], w_up_t [bottleneck][d]. */
int pevit_phm_expand(const float* rule, int32_t n, const float* down_left, const float* down_right, const float* up_left,
                     const float* up_right, int32_t d, int32_t bottleneck, void* w_down, void* w_down_t, void* w_up,
                     void* w_up_t, void* stream);
/* Factor gradients (what autograd derives through the einsum) from pevit_block_bwd's dense d_w_down / d_w_up
 * ([d][bottleneck] each).  d_rule (nullable: the shared rule is frozen in the reference driver, compacter_clip.py:122)
 * is accumulated with atomics -- caller zeroes it or passes the .grad buffer; accumulate != 0: += on the others. */
int pevit_phm_factor_grads(const float* d_w_down, const float* d_w_up, const float* rule, int32_t n, const float* down_left,
                           const float* down_right, const float* up_left, const float* up_right, int32_t d,
                           int32_t bottleneck, float* d_rule, float* d_down_left, float* d_down_right, float* d_up_left,
                           float* d_up_right, int32_t accumulate, void* stream);
/* Adapter (adapter_model.py:204-295): dense fp32 down [bottleneck][d] / up [d][bottleneck] -> the same four operands. */
int pevit_bottleneck_pack(const float* w_down, const float* w_up, int32_t d, int32_t bottleneck, void* w_down_bf16,
                          void* w_down_t, void* w_up_bf16, void* w_up_t, void* stream);

/* ------------------------------------------------------------------ step tail
 * Linear head + CrossEntropyLoss (kadaptation_clip.py:176-185, :350): logits = feat W^T + b, *loss += mean over the
 * n samples of (logsumexp - logit[label]) (caller zeroes loss), dlogits = (softmax - onehot) / n.  fp32 throughout. */
int pevit_head_ce_fwd(const float* feat, const float* w, const float* b, const int64_t* labels, int32_t n, int32_t e,
                      int32_t c, float* logits, float* dlogits, float* loss, void* stream);
/* gscale: device scalar multiplying every gradient (autograd's grad_output; NULL = 1).  dfeat_bf16 [n][e] (nullable):
 * gradient w.r.t. the features, bf16 = A operand of the projection dgrad GEMM; dw [c][e], db [c] (nullable):
 * accumulate != 0 adds onto them. */
int pevit_head_ce_bwd(const float* dlogits, const float* feat, const float* w, const float* gscale, int32_t n, int32_t e,
                      int32_t c, void* dfeat_bf16, float* dw, float* db, int32_t accumulate, void* stream);
/* torch.optim.SGD(momentum, weight_decay; dampening 0, no Nesterov; optim/build.py:18-127) over flat buffers:
 * g' = grad_scale * g + wd * p; m = momentum * m + g'; p -= lr * m.  grad_scale folds 1/world_size in. */
int pevit_sgd_momentum(float* p, const float* g, float* m, size_t n, float lr, float momentum, float weight_decay,
                       float grad_scale, void* stream);
/* ------------------------------------------------------------------ data-parallel exchange fused with the update
 * (the reference's loop has optimizer.step() at kadaptation_clip.py:353 and no exchange; under
 * data parallelism it would be DDP's ncclAllReduce followed by torch.optim.SGD.)  ONE launch per rank and step:
 * a one-shot all-reduce over peer-mapped memory -- every rank loads each 16-byte chunk of all `world` gradient
 * buffers over NVLink, sums them in rank order (bit-identical sums on every rank) -- fused with the momentum-SGD
 * update of the local flat parameters.  Elements [0, n_decayed) take weight_decay, the rest 0 (optim/build.py:18-86).
 *
 * Buffers: pevit_peer_alloc() cudaMallocs n gradient floats + a control block (pevit_peer_buffer_bytes(n) in all),
 * zeroes it and returns a 64-byte CUDA IPC handle to hand to the other ranks (any transport: the caller's
 * torch.distributed all_gather); pevit_peer_open() maps a peer's buffer.  peers[r] = rank r's base pointer
 * (peers[rank] = the own allocation).  The first `n` floats of the own buffer are the flat gradient buffer the
 * backward pass fills; it keeps the LOCAL gradient (the sum goes straight into the update).
 * Synchronisation is inside the kernel (flag hand-shakes at system scope, monotonic epochs kept in device memory, so a
 * captured CUDA graph replays it); every rank must launch it once per step with the same n.  A hand-shake that waits
 * longer than 20 s (PEVIT_PEER_TIMEOUT_MS) gives up and raises the error word pevit_peer_status() reads -- never a hung device.
 * world <= 8, one node. */
size_t pevit_peer_buffer_bytes(size_t n_floats);
int pevit_peer_alloc(size_t n_floats, void** ptr, void* ipc_handle_64_bytes);
int pevit_peer_open(const void* ipc_handle_64_bytes, void** ptr);
int pevit_peer_close(void* ptr);
int pevit_peer_free(void* ptr);
int pevit_peer_status(const void* own, size_t n_floats, int32_t* timed_out, void* stream);   /* synchronises `stream` */
int pevit_allreduce_sgd(void* const* peers, int32_t world, int32_t rank, size_t n, size_t n_decayed, float* p, float* m,
                        float lr, float momentum, float weight_decay, float grad_scale, void* stream);
/* weight packing: fp32 [rows][cols] -> bf16 (same layout / transposed with leading dim ldd) */
int pevit_cast_bf16(const float* src, void* dst, size_t n, void* stream);
int pevit_transpose_bf16(const float* src, int32_t rows, int32_t cols, void* dst, int32_t ldd, void* stream);

/* ------------------------------------------------------------------ stem
 * VisionTransformer.forward up to the first block (model.py:1034-1042): stride-p conv1 as im2col + tcgen05
 * GEMM, class token, positional embedding, ln_pre, NLD -> LND.  images fp32 (N,3,R,R) -> x fp32 (L,N,D).
 * w_patch: bf16 [D][ceil8(3 p^2)] flattened conv1.weight, zero padded along K.  Frozen parameters: no backward.
 * pos_rows: rows of the positional embedding; must equal (resolution/patch)^2 + 1 (the reference fails with a
 * broadcast error otherwise, model.py:1040). */
size_t pevit_patch_embed_workspace_bytes(int32_t nb, int32_t resolution, int32_t patch, int32_t d);
int pevit_patch_embed(const float* images, const void* w_patch, const float* cls, const float* pos, const float* ln_g,
                      const float* ln_b, float* x, void* workspace, int32_t nb, int32_t resolution, int32_t patch,
                      int32_t d, int32_t pos_rows, void* stream);
/* The same stem for other pixel formats.  PEVIT_PX_BF16: the reference accepts any float dtype (encode_image casts,
 * model.py:1152) and the stem rounds pixels to bf16 anyway, so bf16 images give bit-identical patches at half the
 * host-to-device bytes.  PEVIT_PX_U8: raw [0,255] pixels, torchvision's ToTensor + Normalize -- the transform the
 * reference's data loader applies on the host -- evaluated in the kernel in fp32 with the host's operation order,
 * (u8 / 255 - mean[c]) / std[c] (IEEE division: bit-identical to the host result); mean / std are 3 HOST floats each
 * (copied at launch), ignored for the float formats. */
enum pevit_pixel_dtype { PEVIT_PX_F32 = 0, PEVIT_PX_BF16 = 1, PEVIT_PX_U8 = 2 };
int pevit_patch_embed_px(const void* images, int32_t px_dtype, const float* mean, const float* std_, const void* w_patch,
                         const float* cls, const float* pos, const float* ln_g, const float* ln_b, float* x,
                         void* workspace, int32_t nb, int32_t resolution, int32_t patch, int32_t d, int32_t pos_rows,
                         void* stream);

/* ------------------------------------------------------------------ block level
 * One ResidualAttentionBlock forward / backward (model.py:947-975, lora_model.py,
 * adapter_model.py:298-336, compacter_model.py:465-503), frozen backbone: dgrad only,
 * gradients only for the PEFT tensors. */
typedef struct pevit_block_desc {
  int32_t L, NB, D, H;
  int32_t method;      /* enum pevit_method */
  int32_t r;           /* 32 (KAdaptation), 4 (LoRA), 0 otherwise */
  float alpha;         /* 160 (KAdaptation), 32 (LoRA) */
  int32_t save;        /* 1: fill `saved` for a later pevit_block_bwd */
  int32_t attn_impl;   /* 0 default (delta GEMM + tcgen05 attention), 1 CUDA-core kernel with in-kernel delta,
                        * 2 like 0 with the pair-streaming attention kernels for L > 128 */
  int32_t need_dx;     /* bwd: 0 skips the input gradient (first layer: nothing upstream trains) */
  int32_t out_rows;    /* 0: all L*NB token rows of y are produced.  > 0: only the leading out_rows rows (LND order,
                        * a multiple of NB = whole token indices) are needed -- the last ViT block feeds only
                        * ln_post(x[0]) (model.py:1046), so its out-projection, MLP and their dgrads run on NB rows.
                        * y / dy then hold out_rows rows; attention still sees every key. */
  int32_t causal;      /* 1: causal attention mask (CLIP text tower blocks, model.py:1139-1145, 1154-1167): method
                        * PEVIT_PLAIN, save = 0 (the text tower is frozen: forward only), L <= 128 */
} pevit_block_desc;

typedef struct pevit_block_weights {
  const void* w_qkv_ext;   const void* w_qkv_ext_t;  /* bf16 [3D+2r][D], [D][3D+2r] */
  const float* b_qkv;                                /* [3D] */
  const void* w_o;         const void* w_o_t;        /* bf16 [D][D] each */
  const float* b_o;
  const void* w_fc;        const void* w_fc_t;       /* bf16 [4D][D], [D][4D] */
  const float* b_fc;
  const void* w_proj;      const void* w_proj_t;     /* bf16 [D][4D], [4D][D] */
  const float* b_proj;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const float* qmat;       const void* qmat_t;       /* fp32 [2][D][r]; bf16 [2][r][D] * alpha */
  const float* delta_bias;                           /* KAdaptation attn.b or NULL */
  const void* delta_w;                               /* bf16 [2][D][2r]: alpha*[Q_q|0], alpha*[0|Q_v] */
  /* bottleneck (Adapter / Compacter): dense down/up weights (Compacter: expanded from PHM factors) */
  const float *lna_g, *lna_b;                        /* adapter_norm_before */
  const void* w_down;      const void* w_down_t;     /* bf16 [64][D], [D][64] */
  const float* b_down;
  const void* w_up;        const void* w_up_t;       /* bf16 [D][64], [64][D] */
  const float* b_up;
} pevit_block_weights;

typedef struct pevit_block_grads {      /* fp32 outputs, all nullable when not applicable */
  float* d_pmat;   /* [D][2r]   dP = X^T dT            (KAdaptation / LoRA) */
  float* d_qmat;   /* [2][D][r] dQ = alpha dDelta^T T  (KAdaptation / LoRA) */
  float* d_bias;   /* [D]       KAdaptation attn.b */
  float *d_lna_g, *d_lna_b;            /* adapter LayerNorm */
  float *d_w_down, *d_b_down;          /* [D][64] (= dW_down^T), [64] */
  float *d_w_up, *d_b_up;              /* [D][64], [D]   */
} pevit_block_grads;

size_t pevit_block_saved_bytes(const pevit_block_desc* desc);
size_t pevit_block_workspace_bytes(const pevit_block_desc* desc);
int pevit_block_fwd(const pevit_block_desc* desc, const pevit_block_weights* w, const float* x, float* y,
                    void* saved, void* workspace, void* stream);
/* dy_bf16 (nullable in): bf16 copy of dy if the caller has one; dx_bf16 (nullable out): bf16 copy of dx, to be
 * handed to the block below as its dy_bf16 (saves one cast pass per block). */
int pevit_block_bwd(const pevit_block_desc* desc, const pevit_block_weights* w, const float* x, const float* dy,
                    const void* dy_bf16, float* dx, void* dx_bf16, const pevit_block_grads* grads, const void* saved,
                    void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PEVIT_B200_H_ */

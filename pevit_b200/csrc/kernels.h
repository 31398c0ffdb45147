// Internal C++ interface between the kernel translation units and the C-ABI layer (api.cu).
#pragma once

#include "common.cuh"

namespace pevit {

// ------------------------------------------------------------------ gemm_tcgen05.cu
enum GemmEpi { EPI_F32 = 0, EPI_BF16 = 1, EPI_ACT = 2, EPI_DACT = 3, EPI_QKV = 4 };
enum ActKind { ACT_QUICKGELU = 0, ACT_RELU = 1, ACT_GELU_NEW = 2 };

struct GemmEpilogue {
  const float* bias = nullptr;     // [N] fp32, added to the accumulator (all modes)
  const float* resid = nullptr;    // EPI_F32: fp32 [M][ld_out] added to the result
  const float* resid2 = nullptr;   // EPI_F32: second residual (bottleneck: x1 + m)
  float* out_f32 = nullptr;        // EPI_F32
  bf16* out_bf16 = nullptr;        // EPI_BF16 / EPI_ACT (act(z)) / EPI_DACT (dz)
  bf16* out2_bf16 = nullptr;       // EPI_ACT: pre-activation z (nullable)
  const bf16* aux_bf16 = nullptr;  // EPI_DACT: saved pre-activation z
  const bf16* resid_bf16 = nullptr;  // EPI_BF16: bf16 [M][ld_out] added before rounding (may alias out_bf16)
  int act = ACT_QUICKGELU;         // EPI_ACT / EPI_DACT
  int ld_out = 0;                  // row stride (elements) of out / resid / aux
  // EPI_QKV
  bf16* qkv_hm = nullptr;          // [3][NB*H][L][64] head-major q (pre-scaled 1/8), k, v
  bf16* t_out = nullptr;           // [M][r2] low-rank activations T = X P (bf16: A operand of the delta GEMM)
  int L = 0, NB = 0, H = 0, D = 0, r2 = 0;
  // batched launch (EPI_BF16 only): `batch` products of identical shape in one grid.  Batch b reads A rows
  // [b * a_batch_rows, +M) and B rows [b * b_batch_rows, +N) of the same 2-D operands, and writes either row block
  // b * c_batch_rows of the output / residual tensor (staged epilogue) or the same rows at element offset
  // b * c_batch_elems (direct epilogue: column windows of one wide matrix).
  int batch = 1, a_batch_rows = 0, b_batch_rows = 0, c_batch_rows = 0, c_batch_elems = 0;
  int a_prefetch = 0;  // k-blocks ahead of the ring at which the producer pulls A boxes into L2 (0 = off)
  int debug = 0;  // diagnostics only (tools/gemm_bench.py): 1 = no TMA loads, 2 = no MMAs, 4 = no epilogue stores
};

int gemm_tn(cudaStream_t stream, const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, int epi,
            const GemmEpilogue& ep, int force_bn = 0);

// ------------------------------------------------------------------ layernorm.cu
// y = (x - mean) * rstd * gamma + beta, statistics in fp32 (model.py:154-160).
int layernorm_fwd(cudaStream_t s, const float* x, const float* gamma, const float* beta, bf16* y_bf16, float* y_f32,
                  float* mean, float* rstd, int M, int D);
// dx = LN'(dy_n) [+ dres]; dy_n given as fp32.  gamma frozen unless dgamma/dbeta non-null
// (adapter LayerNorm: atomically accumulated, caller zeroes).  dres_rows >= 0: only the first dres_rows rows of dres
// exist (the residual gradient of the remaining rows is zero); -1: all M rows.  dyn_bf16 (frozen gamma only): dy_n
// as the bf16 output of the dgrad GEMM, used instead of dyn.
int layernorm_bwd(cudaStream_t s, const float* dyn, const float* x, const float* gamma, const float* mean,
                  const float* rstd, const float* dres, float* dx, bf16* dx_bf16, float* dgamma, float* dbeta, int M,
                  int D, int dres_rows = -1, const bf16* dyn_bf16 = nullptr);

// ------------------------------------------------------------------ attention_ref.cu
struct AttnShape {
  int L, NB, H, D, r;  // r = low-rank width per projection (32 KAdaptation, 4 LoRA, 0 none)
  float alpha;         // 160 / 32
  int causal = 0;      // forward only, L <= 128: key j of query l is masked when j > l (text tower, model.py:1139-1145)
};
// q,k,v: head-major bf16 [NB*H][L][64] (q pre-scaled).  T: bf16 [L*NB][2r] (LND rows).
// Qmat: fp32 [2][D][r] (q then v factor), bias: fp32 [D] or null.
// out: o_tok bf16 [L*NB][D] (token rows, LND), lse fp32 [NB*H][L].
int attn_delta_fwd_ref(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* T,
                       const float* Qmat, const float* bias, bf16* o_tok, float* lse);
// in: do_tok bf16 [L*NB][D].  out: dqkv bf16 [L*NB][ld] token rows, cols [0,D)=dq/8,[D,2D)=dk,[2D,3D)=dv;
// ddelta bf16 [2][NB*H][L][64] = dQ' and dV' head-major (== d(delta) viewed as [L*NB][D], F4).
int attn_delta_bwd_ref(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* T,
                       const float* Qmat, const float* bias, const bf16* o_tok, const bf16* do_tok, const float* lse,
                       bf16* dqkv, int ld_dqkv, bf16* ddelta);

// ------------------------------------------------------------------ attention_tc.cu (tcgen05 / TMEM)
// q', k, v' head-major bf16 with the low-rank delta already applied (a.r must be 0); L <= 128.
bool attn_tc_supported(const AttnShape& a);
int attn_fwd_tc(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, bf16* o_tok,
                float* lse);
int attn_bwd_tc(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* do_tok,
                const float* lse, bf16* dqkv, int ld_dqkv, bf16* ddelta);

// attention_tc_long.cu: 128 < L <= 384 (ViT-B/16, ViT-L/14), composed over 128 x 128 (query tile, key block) pairs.
// The backward needs the forward's o_tok (delta = rowsum(dO o O)).
bool attn_tc_long_supported(const AttnShape& a);
int attn_fwd_tc_long(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, bf16* o_tok,
                     float* lse);
int attn_bwd_tc_long(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* o_tok,
                     const bf16* do_tok, const float* lse, bf16* dqkv, int ld_dqkv, bf16* ddelta);

// attention_hr.cu: the same range of L with each head's operands resident in shared memory (loaded once, tail tile of
// ceil16(L mod 128) rows) and, in the forward, the probabilities kept in TMEM as the A operand of P V.
bool attn_hr_supported(const AttnShape& a);
int attn_fwd_hr(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, bf16* o_tok, float* lse);
bool attn_bwd_hr_supported(const AttnShape& a);  // additionally: Q, K, V, dO of one head + the P / dS tiles fit in shared memory
int attn_bwd_hr(cudaStream_t s, const AttnShape& a, const bf16* q, const bf16* k, const bf16* v, const bf16* o_tok,
                const bf16* do_tok, const float* lse, bf16* dqkv, int ld_dqkv, bf16* ddelta);

// ------------------------------------------------------------------ lowrank.cu
// KAdaptation factor expansion: from u1,u2 (rule*_left [32][32]), v1,v2 (rule*_right),
// s (q_proj_adapter1_left [32][D/32]), t (q_proj_adapter1_right [32][D/32]) build
//   w_ext rows [3D,3D+64)  : P_q^T | P_v^T  (bf16, K-major rows of length D)   -- forward B operand
//   w_ext_t cols [3D,3D+64): P_q | P_v      (bf16, [D][3D+64])                 -- dgrad B operand
//   qmat  fp32 [2][D][32], qmat_t bf16 [2][32][D] (B operand of dT = alpha * dDelta * Q)
//   delta_w bf16 [2][D][64]: alpha*[Q_q | 0] and alpha*[0 | Q_v], B operands of delta = T * delta_w^T (K = 2r)
int kad_expand(cudaStream_t s, const float* u1, const float* v1, const float* u2, const float* v2, const float* sfac,
               const float* tfac, int D, float alpha, bf16* w_ext, bf16* w_ext_t, float* qmat, bf16* qmat_t,
               bf16* delta_w);
// LoRA: A_q, A_v [r][D]; B_q, B_v [D][r].
int lora_expand(cudaStream_t s, const float* Aq, const float* Av, const float* Bq, const float* Bv, int D, int r,
                float alpha, bf16* w_ext, bf16* w_ext_t, float* qmat, bf16* qmat_t, bf16* delta_w);
// C[Kc][Nc] (fp32, += with atomics; caller zeroes) = scale * A[M][Kc]^T * B[M][Nc]; A bf16 or fp32, B fp32 or bf16.
int atb_accumulate(cudaStream_t s, const void* A, int a_is_bf16, int lda, const void* B, int b_is_bf16, int ldb, int M,
                   int Kc, int Nc, float scale, float* C);
// gemm_atb_tc.cu: the same product on tcgen05 (both operands MN-major, split over rows, red.add epilogue).
// B exposes nb_cols (<= 64) columns; columns [n_lo, n_lo+n_cnt) of A^T B go to C[kc][0..n_cnt) (row stride ldc).
int atb_tc(cudaStream_t s, const bf16* A, int lda, const bf16* B, int ldb, int nb_cols, int M, int Kc, int n_lo,
           int n_cnt, float scale, float* C, int ldc);
// up to three such products over the same M rows and Kc columns in ONE launch (blockIdx.z = problem)
struct AtbProblem {
  const bf16* A; int lda;
  const bf16* B; int ldb, nb_cols, n_lo, n_cnt;
  float scale;
  float* C; int ldc;
};
// optional rider of the same launch: out[c] += sum_m X0[m][c] (+ X1[m][c]) over bf16 [M][ld] matrices (first D columns);
// extra CTAs of the grid compute it while the products run (the shared bias gradient reads the same d(delta) planes)
struct AtbColsum {
  const bf16* X0; const bf16* X1; int ld, M, D;
  float* out;
};
bool atb_colsum_rider_supported(int D, int ld);
int atb_tc_batch(cudaStream_t s, const AtbProblem* probs, int count, int M, int Kc, const AtbColsum* colsum = nullptr);
// column sums of one or two (X1 nullable) bf16 [M][ld] matrices (first D columns), atomically accumulated into
// out[D] (caller zeroes).
int colsum_bf16(cudaStream_t s, const bf16* X0, const bf16* X1, int ld, int M, int D, float* out);
// KAdaptation factor gradients from dP [D][64] (q|v) and dQ [2][D][32]; accumulate: += instead of =.
int kad_factor_grads(cudaStream_t s, const float* dP, const float* dQ, const float* u1, const float* v1, const float* u2,
                     const float* v2, const float* sfac, const float* tfac, int D, float* du1, float* dv1, float* du2,
                     float* dv2, float* dsfac, float* dtfac, bool accumulate);
// ... for up to KAD_MAX_LAYERS layers in ONE launch, added onto the gradient buffers: per-layer arrays (host pointers to
// `count` device pointers each) of dP, dQ, s, t, ds, dt; the rule factors and their gradients are shared by all layers.
constexpr int KAD_MAX_LAYERS = 48;
int kad_factor_grads_batch(cudaStream_t s, int count, const float* const* dP, const float* const* dQ,
                           const float* const* sfac, const float* const* tfac, float* const* dsfac, float* const* dtfac,
                           const float* u1, const float* v1, const float* u2, const float* v2, int D, float* du1, float* dv1,
                           float* du2, float* dv2);
// dT (fp32 [M][r2 cols starting at col0]) -> bf16 into dqkv_ext[:, 3D+col0 ...]
int cast_f32_to_bf16_2d(cudaStream_t s, const float* src, int lds, bf16* dst, int ldd, int rows, int cols);

// ------------------------------------------------------------------ phm.cu
// Compacter: expand both PHM layers of a block (down: D -> B, up: B -> D; rule [n][n][n], left [n][in/n], right
// [n][out/n]) into the bf16 GEMM operands w_down [B][D], w_down_t [D][B], w_up [D][B], w_up_t [B][D].
int phm_expand(cudaStream_t s, const float* rule, int n, const float* down_left, const float* down_right,
               const float* up_left, const float* up_right, int D, int B, bf16* w_down, bf16* w_down_t, bf16* w_up,
               bf16* w_up_t);
// Factor gradients from the dense gradients of pevit_block_bwd (d_w_down [D][B] = dH_down, d_w_up [D][B] = dW_up);
// d_rule (nullable) is accumulated atomically (shared by both layers and by every block); accumulate: += for the rest.
int phm_factor_grads(cudaStream_t s, const float* d_w_down, const float* d_w_up, const float* rule, int n,
                     const float* down_left, const float* down_right, const float* up_left, const float* up_right, int D,
                     int B, float* d_rule, float* d_down_left, float* d_down_right, float* d_up_left, float* d_up_right,
                     bool accumulate);
// Adapter: fp32 dense down [B][D] / up [D][B] -> the same four bf16 operands.
int bottleneck_pack(cudaStream_t s, const float* w_down, const float* w_up, int D, int B, bf16* o_down, bf16* o_down_t,
                    bf16* o_up, bf16* o_up_t);

// ------------------------------------------------------------------ stem.cu
// Patch embedding + class token + positional embedding + ln_pre -> x (L, N, D) fp32 (model.py:1034-1042).
// w_patch: bf16 [D][Kpad], Kpad = ceil8(3 p^2), the flattened conv1 weight zero-padded along K.
size_t patch_embed_workspace_bytes(int NB, int R, int p, int D);
// px_dtype (pevit_pixel_dtype): 0 fp32, 1 bf16, 2 uint8 normalised in the kernel with mean / std (3 host floats each).
int patch_embed(cudaStream_t s, const void* img, int px_dtype, const float* mean, const float* stdv, const bf16* w_patch,
                const float* cls, const float* pos, const float* ln_g, const float* ln_b, float* x, void* workspace,
                int NB, int R, int p, int D);

// ------------------------------------------------------------------ tail.cu
// Linear head + cross-entropy (mean over N): logits [N][C], dlogits = (softmax - 1hot)/N, *loss += mean loss
// (caller zeroes loss).  labels are int64.
int head_ce_fwd(cudaStream_t s, const float* feat, const float* W, const float* b, const long long* labels, int N, int E,
                int C, float* logits, float* dlogits, float* loss);
// gscale: device scalar multiplying every gradient (autograd's grad_output; null = 1).  dfeat bf16 [N][E] (nullable),
// dW [C][E] / db [C] (nullable; accumulate: += instead of =).
int head_ce_bwd(cudaStream_t s, const float* dlogits, const float* feat, const float* W, const float* gscale, int N, int E,
                int C, bf16* dfeat, float* dW, float* db, int accumulate);
// torch.optim.SGD(momentum, weight_decay) over flat buffers; gscale folds 1/world_size of the gradient all-reduce in.
int sgd_momentum(cudaStream_t s, float* p, const float* g, float* m, size_t n, float lr, float mu, float wd, float gscale);

// ------------------------------------------------------------------ peer.cu
// One-shot all-reduce of the flat gradient buffer over peer-mapped (CUDA IPC) memory fused with the SGD update.
constexpr int PEER_MAX_WORLD = 8, PEER_MAX_CTAS = 32, PEER_HANDLE_BYTES = 64;
size_t peer_buffer_bytes(size_t n);   // n gradient floats + the control block (flags, epochs, error word)
int peer_alloc(size_t n, void** ptr, void* ipc_handle);
int peer_open(const void* ipc_handle, void** ptr);
int peer_close(void* ptr);
int peer_free(void* ptr);
int peer_status(cudaStream_t s, const void* own, size_t n, int* timed_out);
int allreduce_sgd(cudaStream_t s, void* const* peers, int world, int rank, size_t n, size_t n_decayed, float* p, float* m,
                  float lr, float mu, float wd, float gscale);

// ------------------------------------------------------------------ elementwise.cu
int cast_f32_to_bf16(cudaStream_t s, const float* src, bf16* dst, size_t n);
// dst[c][r] = src[r][c]  (fp32 -> bf16 transpose; weight prep for dgrad GEMMs)
int transpose_f32_to_bf16(cudaStream_t s, const float* src, int rows, int cols, bf16* dst, int ldd);

}  // namespace pevit

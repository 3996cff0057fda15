// Dense layer y = x W^T + b on the tcgen05 tap GEMM (MR_BF16 path of the multi-head attention projections,
// models/Modules/Attention.py:101-102,125-127: keyProject / valueProject of MHA_Encoder and MHA_User_Encoder).
//
// The A operand is either a dense fp32 activation matrix (cast to bf16 once) or -- the news encoder -- rows of the bf16
// token table gathered by token id inside the GEMM's producer warps (BERT.py:39 fused in; the [T, E] embedding tensor is
// never materialised).  Backward:
//   dense :  d_x = d_y W (tap GEMM), d_w = d_y^T x (token-reduction GEMM), d_b = column sums;
//   gather:  "sum before multiply" as for the conv (news_cnn_tc.cu): S[v, :] = sum_{t: ids[t] = v} d_y[t, :] with the sorted,
//            atomic-free segmented reduction, then d_table = S W and d_w = S^T table over V vocabulary rows instead of T tokens
//            (44x fewer tensor FLOPs at the config-5 shape), no [T, E] gradient tensor.
// Operand roundings: x / table, W and (backward) d_y / S to bf16; accumulation and outputs fp32.
#include "gemm_simt.cuh"
#include "tapgemm.cuh"
#include "tokred.cuh"

extern "C" int64_t mr_embed_grad_workspace_bytes(int64_t T, int64_t E, int64_t V);
extern "C" int mr_embed_grad_segreduce(const void* ids, int ids_i64, const void* d_emb, int d_emb_dtype, int64_t d_emb_ld,
                                       float* d_table, int64_t T, int64_t E, int64_t V, int64_t padding_idx, void* workspace,
                                       int64_t workspace_bytes, void* stream);

namespace mr {

// fp32 [M, C] with row pitch ld -> bf16 [Mp, Cp], zero padded rows / columns
__global__ void lt_cast_kernel(const float* __restrict__ src, int64_t ld, __nv_bfloat16* __restrict__ dst, int64_t M, int64_t Mp,
                               int C, int Cp) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp * Cp) return;
  const int64_t m = i / Cp;
  const int c = (int)(i - m * Cp);
  dst[i] = __float2bfloat16((m < M && c < C) ? src[m * ld + c] : 0.f);
}
// dst[m, :C] = src[m, :C] (pitches lds / ldd)
__global__ void lt_copy_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t M, int C) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const int64_t m = i / C;
  const int c = (int)(i - m * C);
  dst[m * ldd + c] = src[m * lds + c];
}
// column sums of an fp32 [R, C] matrix with row pitch ld (bias gradient), fixed order: chunks of rows, then the chunks
__global__ void lt_colsum_partial_kernel(const float* __restrict__ x, int64_t ld, float* __restrict__ partial, int64_t R, int C,
                                         int64_t rows_per_chunk) {
  __shared__ float sm[8][33];
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int c = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(R, r0 + rows_per_chunk);
  float s = 0.f;
  if (c < C)
    for (int64_t r = r0 + ry; r < r1; r += 8) s += x[r * ld + c];
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][cx];
    partial[(int64_t)blockIdx.y * C + c] = t;
  }
}
__global__ void lt_colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int64_t chunks, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int64_t i = 0; i < chunks; ++i) s += partial[i * C + c];
  out[c] = s;
}

struct LtGeom {
  int64_t M, Mp, N, Np, K, Kp, V, Vp;
  int64_t nblk_n, nbsz_n;      // output-column blocks of the forward GEMM (<= 256 columns: double-buffered accumulator)
  int64_t nblk_k, nbsz_k;      // ... of the data-gradient GEMM
};
static LtGeom lt_geom(int64_t M, int64_t N, int64_t K, int64_t V) {
  LtGeom g;
  // N is padded to 32 and cut into blocks that are multiples of 32 columns: the token-reduction GEMM then stages its Q operand
  // in 64- or 32-column swizzled blocks (the layouts the conv backward exercises), never the 16-column one
  g.M = M; g.Mp = align_up(M > 0 ? M : 1, 128); g.N = N; g.Np = align_up(N, 32); g.K = K; g.Kp = align_up(K, 16);
  g.V = V; g.Vp = align_up(V > 0 ? V : 1, 128);
  g.nblk_n = ceil_div(g.Np, 256); g.nbsz_n = align_up(ceil_div(g.Np, g.nblk_n), 32);
  g.nblk_k = ceil_div(g.Kp, 256); g.nbsz_k = align_up(ceil_div(g.Kp, g.nblk_k), 16);
  return g;
}

static int64_t lt_ws(int64_t M, int64_t N, int64_t K, int64_t V, int backward) {
  const LtGeom g = lt_geom(M, N, K, V);
  int64_t b = 256;
  if (!backward) {
    if (V <= 0) b += arena_bytes(g.Mp * g.Kp, 2);                                   // x as bf16
    b += arena_bytes(tapgemm_pack_bytes(1, (int)g.nbsz_n, (int)g.Kp), 1);
    return b;
  }
  b += arena_bytes(tapgemm_pack_bytes(1, (int)g.nbsz_k, (int)g.Np), 1);             // W for the data gradient
  b += arena_bytes(colsum_chunks(M) * N, 4);                                        // bias-gradient partials
  if (V <= 0) {
    b += arena_bytes(g.Mp * g.Kp, 2) + arena_bytes(g.Mp * g.Np, 2);                 // x, d_y as bf16
    b += arena_bytes(M * align_up(K, 4), 4);                                        // d_x with 16-byte rows (K % 4 != 0)
    b += arena_bytes(tokred_partial_bytes(g.Mp / 128, 128, 1, (int)g.Kp, (int)g.nbsz_n), 1);
  } else {
    b += arena_bytes(V * N, 4);                                                     // S fp32
    b += arena_bytes(g.Vp * g.Np, 2);                                               // S bf16
    b += arena_bytes(mr_embed_grad_workspace_bytes(M, N, V), 1);
    b += arena_bytes(tokred_partial_bytes(ceil_div(V, 32), 32, 1, (int)g.Kp, (int)g.nbsz_n), 1);
  }
  return b;
}

// out[rows, :n_cols] (fp32, pitch ldo, 16-byte rows) = A[rows, Kred] x Wt, Wt[n, k] = w[n*sn + k*sk]; A dense bf16 or gathered
static int lt_gemm(const __nv_bfloat16* a_dense, int64_t lda, const void* ids, int ids_i64, int64_t V, int64_t rows, int64_t row_tiles,
                   int tile_rows, const float* w, int64_t sn, int64_t sk, int n_cols, int k_valid, int Kred, int64_t nblk, int64_t nbsz,
                   const float* bias, float* out, int64_t ldo, uint8_t* wp, cudaStream_t st) {
  const int64_t n_pad = align_up(n_cols, 16);
  for (int64_t blk = 0; blk < nblk; ++blk) {
    const int64_t n0 = blk * nbsz;
    const int64_t nb = (n_pad - n0) < nbsz ? (n_pad - n0) : nbsz;
    const int64_t nv = (n_cols - n0) < nb ? (n_cols - n0) : nb;
    if (nv <= 0 || nb <= 0) break;
    if (int rc = tapgemm_pack(w + n0 * sn, wp, 1, (int)nb, Kred, (int)nv, k_valid, sn, sk, 0, st)) return rc;
    TapGemmArgs a{};
    TapGemmPlan plan;
    a.n_titles = row_tiles; a.L = tile_rows; a.taps = 1; a.dir = 1; a.K = Kred;
    a.n_sub = 1; a.nsz[0] = (int)nb;
    if (ids != nullptr) { a.ids = ids; a.ids_i64 = ids_i64; a.V = V; }
    a.a = a_dense; a.lda = lda;
    a.wpack = wp; a.epi = TG_EPI_BIAS_F32; a.bias = bias ? bias + n0 : nullptr; a.bias2 = nullptr; a.n_valid = (int)nv;
    a.n_rows = rows; a.out_f32 = out + n0; a.ldo = ldo; a.n_store = (int)align_up(nv, 4);
    if (int rc = tapgemm_plan(a, &plan)) return rc;
    if (int rc = tapgemm_launch(plan, st)) return rc;
  }
  return MR_OK;
}

// d_w[n, k] = sum_r Q[r, n] P[r, k]:  P (the layer input, KP = K rounded up to 16 columns) is the 128-wide-slice operand of the
// token-reduction GEMM, Q (the output gradient) its N <= 256 operand -- blocks of the output columns, one launch each
static int lt_wgrad(const __nv_bfloat16* p, int64_t ldp, const __nv_bfloat16* q, int64_t ldq, int64_t row_tiles, int tile_rows, const LtGeom& g,
                    float* partial, float* d_w, cudaStream_t st) {
  for (int64_t blk = 0; blk < g.nblk_n; ++blk) {
    const int64_t n0 = blk * g.nbsz_n;
    const int64_t nb = (g.Np - n0) < g.nbsz_n ? (g.Np - n0) : g.nbsz_n;
    const int64_t nv = (g.N - n0) < nb ? (g.N - n0) : nb;
    if (nv <= 0 || nb <= 0) break;
    TokRedArgs a{};
    TokRedPlan plan;
    a.n_titles = row_tiles; a.L = tile_rows; a.taps = 1;
    a.ids = nullptr; a.p = p; a.ldp = ldp; a.KP = (int)g.Kp;
    a.q = q + n0; a.ldq = ldq; a.NQ = (int)nb; a.partial = partial;
    if (int rc = tokred_plan(a, &plan)) return rc;
    if (int rc = tokred_launch(plan, st)) return rc;
    if (int rc = tokred_reduce(plan, d_w + n0 * g.K, (int)g.K, (int)nv, g.K, 1, 0, st)) return rc;
  }
  return MR_OK;
}

}  // namespace mr

extern "C" {
using namespace mr;

int64_t mr_linear_tc_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t V, int backward) {
  if (M < 0 || N < 1 || K < 1) return -1;
  return lt_ws(M, N, K, V, backward);
}

int mr_linear_tc_fwd(const float* x, const void* ids, int ids_i64, const void* table_bf16, int64_t table_ld, int64_t V,
                     const float* w, const float* b, float* y, int64_t ldy, int64_t M, int64_t N, int64_t K, void* workspace,
                     int64_t workspace_bytes, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(w && y && (x != nullptr || (ids != nullptr && table_bf16 != nullptr)), MR_ERR_NULL, "mr_linear_tc_fwd: null pointer");
  MR_REQUIRE(M >= 0 && N >= 1 && K >= 1 && ldy % 4 == 0 && ldy >= align_up(N, 4), MR_ERR_BAD_SHAPE,
             "mr_linear_tc_fwd: M=%lld N=%lld K=%lld ldy=%lld (ldy must be a multiple of 4 and >= N rounded up to 4)", (long long)M,
             (long long)N, (long long)K, (long long)ldy);
  MR_REQUIRE(align_up(K, 16) <= 1024, MR_ERR_UNSUPPORTED, "mr_linear_tc_fwd: K=%lld too large", (long long)K);
  if (M == 0) return MR_OK;
  const bool gather = x == nullptr;
  if (gather) MR_REQUIRE(V >= 1 && table_ld >= align_up(K, 16), MR_ERR_BAD_SHAPE, "mr_linear_tc_fwd: table pitch %lld < K rounded up to 16", (long long)table_ld);
  cudaStream_t st = as_stream(stream);
  const LtGeom g = lt_geom(M, N, K, gather ? V : 0);
  Arena ar(workspace, workspace_bytes);
  __nv_bfloat16* xb = gather ? nullptr : ar.take<__nv_bfloat16>(g.Mp * g.Kp);
  uint8_t* wp = ar.take<uint8_t>(tapgemm_pack_bytes(1, (int)g.nbsz_n, (int)g.Kp));
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_linear_tc_fwd: workspace too small (%lld given)", (long long)workspace_bytes);
  if (!gather) {
    launch_pdl(lt_cast_kernel, dim3((unsigned)ceil_div(g.Mp * g.Kp, 256)), dim3(256), 0, st, x, K, xb, M, g.Mp, (int)K, (int)g.Kp);
    MR_CHECK_LAUNCH("lt_cast_kernel");
    return lt_gemm(xb, g.Kp, nullptr, 0, 0, M, g.Mp / 128, 128, w, K, 1, (int)N, (int)K, (int)g.Kp, g.nblk_n, g.nbsz_n, b, y, ldy, wp, st);
  }
  return lt_gemm(static_cast<const __nv_bfloat16*>(table_bf16), table_ld, ids, ids_i64, V, M, ceil_div(M, 128), 128, w, K, 1, (int)N,
                 (int)K, (int)g.Kp, g.nblk_n, g.nbsz_n, b, y, ldy, wp, st);
}

int mr_linear_tc_bwd(const float* x, const void* ids, int ids_i64, const void* table_bf16, int64_t table_ld, int64_t table_rows,
                     int64_t V, int64_t padding_idx, const float* w, const float* d_y, int64_t ldy, float* d_x, float* d_table,
                     float* d_w, float* d_b, int64_t M, int64_t N, int64_t K, void* workspace, int64_t workspace_bytes,
                     void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(w && d_y && d_w && (x != nullptr || (ids != nullptr && table_bf16 != nullptr)), MR_ERR_NULL, "mr_linear_tc_bwd: null pointer");
  MR_REQUIRE(M >= 0 && N >= 1 && K >= 1 && ldy >= N, MR_ERR_BAD_SHAPE, "mr_linear_tc_bwd: bad shape");
  MR_REQUIRE(align_up(K, 16) <= 512 && align_up(N, 16) <= 1024, MR_ERR_UNSUPPORTED, "mr_linear_tc_bwd: N=%lld K=%lld too large", (long long)N,
             (long long)K);
  cudaStream_t st = as_stream(stream);
  const bool gather = x == nullptr;
  if (M == 0) {
    cudaMemsetAsync(d_w, 0, sizeof(float) * N * K, st);
    if (d_b) cudaMemsetAsync(d_b, 0, sizeof(float) * N, st);
    if (gather && d_table) cudaMemsetAsync(d_table, 0, sizeof(float) * V * K, st);
    return MR_OK;
  }
  const LtGeom g = lt_geom(M, N, K, gather ? V : 0);
  Arena ar(workspace, workspace_bytes);
  uint8_t* wp = ar.take<uint8_t>(tapgemm_pack_bytes(1, (int)g.nbsz_k, (int)g.Np));
  float* cpart = ar.take<float>(colsum_chunks(M) * N);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_linear_tc_bwd: workspace too small (%lld given)", (long long)workspace_bytes);
  if (d_b) {
    const int64_t chunks = colsum_chunks(M), rpc = ceil_div(M, chunks);
    lt_colsum_partial_kernel<<<dim3((unsigned)ceil_div(N, 32), (unsigned)chunks), dim3(32, 8), 0, st>>>(d_y, ldy, cpart, M, (int)N, rpc);
    MR_CHECK_LAUNCH("lt_colsum_partial_kernel");
    lt_colsum_final_kernel<<<(unsigned)ceil_div(N, 128), 128, 0, st>>>(cpart, d_b, chunks, (int)N);
    MR_CHECK_LAUNCH("lt_colsum_final_kernel");
  }
  if (!gather) {
    __nv_bfloat16* xb = ar.take<__nv_bfloat16>(g.Mp * g.Kp);
    __nv_bfloat16* gb = ar.take<__nv_bfloat16>(g.Mp * g.Np);
    const int64_t K4 = align_up(K, 4);
    float* dx_tmp = ar.take<float>(M * K4);
    float* partial = ar.take<float>(tokred_partial_bytes(g.Mp / 128, 128, 1, (int)g.Kp, (int)g.nbsz_n) / 4);
    MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_linear_tc_bwd: workspace too small (%lld given)", (long long)workspace_bytes);
    launch_pdl(lt_cast_kernel, dim3((unsigned)ceil_div(g.Mp * g.Kp, 256)), dim3(256), 0, st, x, K, xb, M, g.Mp, (int)K, (int)g.Kp);
    MR_CHECK_LAUNCH("lt_cast_kernel");
    launch_pdl(lt_cast_kernel, dim3((unsigned)ceil_div(g.Mp * g.Np, 256)), dim3(256), 0, st, d_y, ldy, gb, M, g.Mp, (int)N, (int)g.Np);
    MR_CHECK_LAUNCH("lt_cast_kernel");
    if (d_x) {        // d_x[m, k] = sum_n d_y[m, n] w[n, k]
      float* out = K4 == K ? d_x : dx_tmp;
      if (int rc = lt_gemm(gb, g.Np, nullptr, 0, 0, M, g.Mp / 128, 128, w, 1, K, (int)K, (int)N, (int)g.Np, g.nblk_k, g.nbsz_k, nullptr, out,
                           K4, wp, st))
        return rc;
      if (out != d_x) {
        launch_pdl(lt_copy_kernel, dim3((unsigned)ceil_div(M * K, 256)), dim3(256), 0, st, (const float*)dx_tmp, K4, d_x, K, M, (int)K);
        MR_CHECK_LAUNCH("lt_copy_kernel");
      }
    }
    // d_w[n, k] = sum_m d_y[m, n] x[m, k]
    return lt_wgrad(xb, g.Kp, gb, g.Np, g.Mp / 128, 128, g, partial, d_w, st);
  }
  // ---- gather mode: token-grouped ------------------------------------------------------------------------------------
  MR_REQUIRE(d_table != nullptr && K % 4 == 0 && table_rows >= align_up(V, 32) && table_ld >= g.Kp, MR_ERR_BAD_SHAPE,
             "mr_linear_tc_bwd: the gather mode needs d_table, K %% 4 == 0 and a bf16 table with >= %lld (zero padded) rows of pitch >= %lld",
             (long long)align_up(V, 32), (long long)g.Kp);
  float* S = ar.take<float>(V * N);
  __nv_bfloat16* Sb = ar.take<__nv_bfloat16>(g.Vp * g.Np);
  const int64_t ewb = mr_embed_grad_workspace_bytes(M, N, V);
  void* ews = ar.take<uint8_t>(ewb);
  float* partial = ar.take<float>(tokred_partial_bytes(ceil_div(V, 32), 32, 1, (int)g.Kp, (int)g.nbsz_n) / 4);
  MR_REQUIRE(ar.ok() && ewb >= 0, MR_ERR_WORKSPACE, "mr_linear_tc_bwd: workspace too small (%lld given)", (long long)workspace_bytes);
  // S[v, :] = sum over tokens with id v of d_y[t, :]  (every token counts here, the padding row too: it is an INPUT of the layer)
  if (int rc = mr_embed_grad_segreduce(ids, ids_i64, d_y, MR_F32, ldy, S, M, N, V, -1, ews, ewb, stream)) return rc;
  launch_pdl(lt_cast_kernel, dim3((unsigned)ceil_div(g.Vp * g.Np, 256)), dim3(256), 0, st, (const float*)S, N, Sb, V, g.Vp, (int)N, (int)g.Np);
  MR_CHECK_LAUNCH("lt_cast_kernel");
  // d_table[v, k] = sum_n S[v, n] w[n, k];  the padding row of the table gets no gradient (BERT.py:16-21)
  if (int rc = lt_gemm(Sb, g.Np, nullptr, 0, 0, V, ceil_div(V, 32), 32, w, 1, K, (int)K, (int)N, (int)g.Np, g.nblk_k, g.nbsz_k, nullptr, d_table, K,
                       wp, st))
    return rc;
  if (padding_idx >= 0 && padding_idx < V) cudaMemsetAsync(d_table + padding_idx * K, 0, sizeof(float) * K, st);
  // d_w[n, k] = sum_v S[v, n] table[v, k]
  return lt_wgrad(static_cast<const __nv_bfloat16*>(table_bf16), table_ld, Sb, g.Np, ceil_div(V, 32), 32, g, partial, d_w, st);
}

}  // extern "C"

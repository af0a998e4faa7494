// K3/K5/K6 entry points: bf16 tensor-core GEMMs on tcgen05 with TMEM accumulators, operands staged by TMA into 128B-swizzled
// shared memory.  Replaces nn.Linear forward / dgrad (NT) and wgrad (TN) of model2_seq.py:83-90,97-99,109,121-126.  The kernels
// live in gemm_tc2.cu; this file holds the tensor-map builders, the C ABI and its argument checks.
//
// NT:  C[M,N] = A[M,K] . B[N,K]^T ; A, B K-major (K contiguous).
// TN:  C[N',K'] += A[M,N']^T . B[M,K'] ; both operands MN-major (contraction dim M is the slow dim), split over M across CTAs,
//      fp32 vector reductions into C.
#include <algorithm>
#include <atomic>

#include "tc_common.cuh"

namespace dsf {

// ---------------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// bf16 row-major matrix [rows, cols] with leading dimension ld (elements); box = 64 cols x box_rows, 128B swizzle
int make_tmap_bf16(CUtensorMap* m, const void* base, int rows, int cols, int ld, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DSF_ELAUNCH; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box_rows=%d", (int)r, rows, cols, ld, box_rows); return DSF_ELAUNCH; }
  return DSF_OK;
}

// row-major matrix [rows, cols] of bf16 (dtype DSF_BF16) or fp32 (DSF_F32), leading dimension ld (elements);
// box = box_cols x box_rows with box_cols * elem_size == 128 B, 128B swizzle (used for the TMA-store epilogue)
int make_tmap_2d(CUtensorMap* m, const void* base, int dtype, int rows, int cols, int ld, int box_cols, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DSF_ELAUNCH; }
  const int es = dtype == DSF_F32 ? 4 : 2;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, dtype == DSF_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d) failed (%d) rows=%d cols=%d ld=%d box=%dx%d", (int)r, rows, cols, ld, box_cols, box_rows); return DSF_ELAUNCH; }
  return DSF_OK;
}

}  // namespace dsf

namespace dsf {
// persistent kernels (double-buffered TMEM accumulators; CTA pairs with cta_group::2 where the shape allows), gemm_tc2.cu
int gemm_nt_run(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int c_dtype, const float* bias, const float* residual, int M,
                int N, int K, int flags, const dsf_dropout* drop, const void* relu_src, bool pairs, cudaStream_t st);
int gemm_tn_run(const void* A, int lda, const void* B, int ldb, float* C, int ldc, int M, int Nout, int Kout, float* colsum, cudaStream_t st);
// 0 = default (NT on CTA pairs where N % 128 == 0 and M > 128, else single-CTA tiles), 2 = single-CTA tiles only.  Relaxed
// atomic: calls are re-entrant, the selector is a tuning / test knob.
static std::atomic<int> g_gemm_impl{0};
}  // namespace dsf

using namespace dsf;

extern "C" int dsf_gemm_set_impl(int32_t impl) {
  DSF_REQUIRE(impl == 0 || impl == 2, "gemm_set_impl: impl must be 0 (default: CTA pairs where the shape allows) or 2 (single-CTA tiles only)");
  g_gemm_impl.store(impl, std::memory_order_relaxed);
  return DSF_OK;
}

extern "C" int dsf_gemm_bf16_nt(const void* A, int32_t lda, const void* B, int32_t ldb, void* C, int32_t ldc, int32_t c_dtype,
                                const float* bias, const float* residual, int32_t M, int32_t N, int32_t K, int32_t epi_flags,
                                const dsf_dropout* drop, const void* relu_src, void* stream) {
  DSF_REQUIRE(A && B && C, "gemm_bf16_nt: NULL pointer");
  DSF_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16_nt: non-positive extent");
  DSF_REQUIRE(K % 64 == 0, "gemm_bf16_nt: K=%d must be a multiple of 64", K);
  DSF_REQUIRE(N % 64 == 0, "gemm_bf16_nt: N=%d must be a multiple of 64", N);
  DSF_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && lda >= K && ldb >= K && ldc >= N, "gemm_bf16_nt: bad leading dimensions");
  DSF_REQUIRE(aligned16(A) && aligned16(B) && aligned16(C) && aligned16(bias) && aligned16(residual), "gemm_bf16_nt: 16-byte alignment required");
  DSF_REQUIRE(c_dtype == DSF_F32 || c_dtype == DSF_BF16, "gemm_bf16_nt: bad c_dtype %d", c_dtype);
  DSF_REQUIRE(!(epi_flags & DSF_EPI_BIAS) || bias, "gemm_bf16_nt: bias flag without bias pointer");
  DSF_REQUIRE(!(epi_flags & DSF_EPI_RESIDUAL) || residual, "gemm_bf16_nt: residual flag without residual pointer");
  DSF_REQUIRE(!(epi_flags & DSF_EPI_ACCUM), "gemm_bf16_nt: ACCUM is not supported on the NT path");
  DSF_REQUIRE(!drop || (drop->p >= 0.f && drop->p < 1.f), "gemm_bf16_nt: dropout p must be in [0, 1)");
  DSF_REQUIRE(!relu_src || aligned16(relu_src), "gemm_bf16_nt: relu_src must be 16-byte aligned");
  return gemm_nt_run(A, lda, B, ldb, C, ldc, c_dtype, bias, residual, M, N, K, epi_flags, drop, relu_src,
                     g_gemm_impl.load(std::memory_order_relaxed) == 0, (cudaStream_t)stream);
}

extern "C" int dsf_gemm_bf16_tn(const void* A, int32_t lda, const void* B, int32_t ldb, float* C, int32_t ldc, int32_t M,
                                int32_t Nout, int32_t Kout, float* colsum_a, void* stream) {
  DSF_REQUIRE(A && B && C, "gemm_bf16_tn: NULL pointer");
  DSF_REQUIRE(M > 0 && Nout > 0 && Kout > 0, "gemm_bf16_tn: non-positive extent");
  DSF_REQUIRE(Nout % 64 == 0 && Kout % 64 == 0, "gemm_bf16_tn: output extents (%d, %d) must be multiples of 64", Nout, Kout);
  DSF_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= Nout && ldb >= Kout && ldc >= Kout, "gemm_bf16_tn: bad leading dimensions");
  DSF_REQUIRE(aligned16(A) && aligned16(B) && aligned16(C), "gemm_bf16_tn: 16-byte alignment required");
  return gemm_tn_run(A, lda, B, ldb, C, ldc, M, Nout, Kout, colsum_a, (cudaStream_t)stream);
}

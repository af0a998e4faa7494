// K3/K5/K6: bf16 tensor-core GEMMs on tcgen05 with TMEM accumulators, operands staged by TMA into
// 128B-swizzled shared memory through an mbarrier ring.  Replaces nn.Linear forward / dgrad (NT) and
// wgrad (TN) of model2_seq.py:83-90,97-99,109,121-126.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> global), one TMEM lane quarter each (warp_id % 4).
//
// NT:  C[M,N] = A[M,K] . B[N,K]^T ; A, B K-major (K contiguous).  Tile 128 x BN x 64.
// TN:  C[N',K'] += A[M,N']^T . B[M,K'] ; both operands MN-major (contraction dim M is the slow dim),
//      split over M across gridDim.z, fp32 atomics into C.
#include <algorithm>

#include "tc_common.cuh"

namespace dsf {

using namespace tc;

constexpr int GT_BM = 128, GT_BK = 64, GT_THREADS = 192;

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

// ---------------------------------------------------------------------------------- epilogue helpers
struct EpiArgs {
  void* C;
  int ldc;
  int c_dtype;
  const float* bias;
  const float* residual;
  int flags;
};

// one thread owns row `row` and 32 consecutive columns starting at n
__device__ __forceinline__ void epilogue_store32(const EpiArgs& e, int row, int n, const uint32_t (&r)[32]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (e.flags & DSF_EPI_BIAS) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (e.flags & DSF_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  const size_t off = (size_t)row * e.ldc + n;
  if (e.flags & DSF_EPI_RESIDUAL) {
    const float4* rp = reinterpret_cast<const float4*>(e.residual + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = rp[j];
      v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
  }
  if (e.c_dtype == DSF_F32) {
    float4* cp = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.C) + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) cp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint4* cp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.C) + off);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      cp[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                         pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
  }
}

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GT_BM * GT_BK * 2;
  static constexpr int B_BYTES = BN * GT_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16;
  static constexpr int DYN = TOTAL + 1024;  // slack for manual 1024-byte alignment
};

// ---------------------------------------------------------------------------------- NT kernel
template <int BN, int STAGES>
__global__ void __launch_bounds__(GT_THREADS)
gemm_bf16_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, EpiArgs epi, int M, int N, int K) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_acc = bar_empty + STAGES * 8;
  const uint32_t tmem_slot = bar_acc + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * BN;
  const int num_k = K / GT_BK;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_empty + s * 8, ph ^ 1);
        mbar_expect_tx(bar_full + s * 8, L::STAGE_BYTES);
        const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
        tma_load_2d(sa, &tmA, bar_full + s * 8, kb * GT_BK, m0);
        tma_load_2d(sb, &tmB, bar_full + s * 8, kb * GT_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GT_BM, BN, 0, 0);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_full + s * 8, ph);
        tc_fence_after();
        const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
        const uint64_t da = make_smem_desc(sa, 16, 1024, SWZ_128B);
        const uint64_t db = make_smem_desc(sb, 16, 1024, SWZ_128B);
#pragma unroll
        for (int k = 0; k < GT_BK / 16; ++k)
          tc_mma_bf16(tmem_base, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (kb | k) != 0);
        tc_commit(bar_empty + s * 8);  // frees the smem slot once these MMAs retire
      }
      tc_commit(bar_acc);  // accumulator complete
    }
  } else {
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, r);
      tmem_wait_ld();
      if (row < M) epilogue_store32(epi, row, n0 + c, r);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------- TN kernel (wgrad)
// A = dY [M, N'] (tile 64 rows x 128 cols as two 64-col TMA boxes), B = X [M, K'] (BN/64 boxes).
template <int BN, int STAGES>
__global__ void __launch_bounds__(GT_THREADS)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* C, int ldc, int M,
                    int Nout, int Kout, int m_chunk) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_acc = bar_empty + STAGES * 8;
  const uint32_t tmem_slot = bar_acc + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * GT_BM;  // rows of C (N' index)
  const int k0 = blockIdx.x * BN;     // cols of C (K' index)
  const int m_lo = blockIdx.z * m_chunk, m_hi = min(M, m_lo + m_chunk);
  const int num_it = (m_hi - m_lo + GT_BK - 1) / GT_BK;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr int BOX_BYTES = GT_BK * 128;  // 64 rows x 128 B

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < num_it; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(bar_empty + s * 8, ph ^ 1);
        mbar_expect_tx(bar_full + s * 8, L::STAGE_BYTES);
        const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
        const int m = m_lo + it * GT_BK;
#pragma unroll
        for (int j = 0; j < GT_BM / 64; ++j) tma_load_2d(sa + j * BOX_BYTES, &tmA, bar_full + s * 8, n0 + j * 64, m);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * BOX_BYTES, &tmB, bar_full + s * 8, k0 + j * 64, m);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GT_BM, BN, 1, 1);
      for (int it = 0; it < num_it; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(bar_full + s * 8, ph);
        tc_fence_after();
        const uint32_t sa = base + s * L::STAGE_BYTES, sb = sa + L::A_BYTES;
        // MN-major, 128B swizzle: LBO = next 64-element MN group (one TMA box), SBO = next 8 K rows
        const uint64_t da = make_smem_desc(sa, BOX_BYTES, 1024, SWZ_128B);
        const uint64_t db = make_smem_desc(sb, BOX_BYTES, 1024, SWZ_128B);
#pragma unroll
        for (int k = 0; k < GT_BK / 16; ++k)
          tc_mma_bf16(tmem_base, desc_advance(da, k * 2048), desc_advance(db, k * 2048), idesc, (it | k) != 0);
        tc_commit(bar_empty + s * 8);
      }
      tc_commit(bar_acc);
    }
  } else {
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int row = n0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c, r);
      tmem_wait_ld();
      if (row < Nout) {
        float* cp = C + (size_t)row * ldc + k0 + c;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (k0 + c + j < Kout) atomicAdd(cp + j, __uint_as_float(r[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES>
int launch_nt(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiArgs& epi, int M, int N, int K, cudaStream_t st) {
  using L = GemmSmem<BN, STAGES>;
  static bool configured_on[64] = {};
  bool& configured = per_device_flag(configured_on);
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_bf16_nt_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN) != cudaSuccess)
      return check_launch("gemm_bf16_nt/attr");
    configured = true;
  }
  dim3 grid(N / BN, cdiv(M, GT_BM));
  gemm_bf16_nt_kernel<BN, STAGES><<<grid, GT_THREADS, L::DYN, st>>>(tmA, tmB, epi, M, N, K);
  return check_launch("gemm_bf16_nt");
}

template <int BN, int STAGES>
int launch_tn(const CUtensorMap& tmA, const CUtensorMap& tmB, float* C, int ldc, int M, int Nout, int Kout, cudaStream_t st) {
  using L = GemmSmem<BN, STAGES>;
  static bool configured_on[64] = {};
  bool& configured = per_device_flag(configured_on);
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN) != cudaSuccess)
      return check_launch("gemm_bf16_tn/attr");
    configured = true;
  }
  const int tiles = cdiv(Nout, GT_BM) * cdiv(Kout, BN);
  // split the contraction so that ~2 waves of CTAs exist; chunks are multiples of 64 rows
  int splits = std::max(1, std::min(cdiv(M, 4 * GT_BK), (2 * num_sms()) / std::max(1, tiles)));
  int m_chunk = cdiv(cdiv(M, splits), GT_BK) * GT_BK;
  splits = cdiv(M, m_chunk);
  dim3 grid(cdiv(Kout, BN), cdiv(Nout, GT_BM), splits);
  gemm_bf16_tn_kernel<BN, STAGES><<<grid, GT_THREADS, L::DYN, st>>>(tmA, tmB, C, ldc, M, Nout, Kout, m_chunk);
  return check_launch("gemm_bf16_tn");
}

}  // namespace dsf

namespace dsf {
// v2 (persistent, double-buffered TMEM accumulators, 128 x 256 tiles), gemm_tc2.cu
int gemm_nt_v2(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int c_dtype, const float* bias, const float* residual, int M,
               int N, int K, int flags, const dsf_dropout* drop, const void* relu_src, cudaStream_t st);
int gemm_tn_v2(const void* A, int lda, const void* B, int ldb, float* C, int ldc, int M, int Nout, int Kout, cudaStream_t st);
extern bool g_nt_pairs;      // gemm_tc2.cu: NT GEMMs on CTA pairs (cta_group::2) where the shape allows
static int g_gemm_impl = 0;  // 0 = default (v3), 1 = v1 (one CTA per 128 x 128 tile), 2 = v2 (persistent single CTA), 3 = v3 (CTA pairs)
int gemm_nt_ln(const void* A, int lda, const void* B, int ldb, float* C, int ldc, const float* bias, const float* residual, void* H, int ldh,
               const float* gamma, const float* beta, float* mean, float* rstd, float eps, int M, int K, const dsf_dropout* drop,
               cudaStream_t st);  // gemm_tc2.cu
}  // namespace dsf

using namespace dsf;

extern "C" int dsf_gemm_set_impl(int32_t impl) {
  DSF_REQUIRE(impl >= 0 && impl <= 3, "gemm_set_impl: impl must be 0 (default), 1, 2 or 3");
  g_gemm_impl = impl;
  g_nt_pairs = (impl == 0 || impl == 3);
  return DSF_OK;
}

extern "C" int dsf_gemm_bf16_nt(const void* A, int32_t lda, const void* B, int32_t ldb, void* C, int32_t ldc, int32_t c_dtype,
                                const float* bias, const float* residual, int32_t M, int32_t N, int32_t K, int32_t epi_flags,
                                const dsf_dropout* drop, const void* relu_src, void* stream) {
  DSF_REQUIRE(A && B && C, "gemm_bf16_nt: NULL pointer");
  DSF_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16_nt: non-positive extent");
  DSF_REQUIRE(K % GT_BK == 0, "gemm_bf16_nt: K=%d must be a multiple of 64", K);
  DSF_REQUIRE(N % 64 == 0, "gemm_bf16_nt: N=%d must be a multiple of 64", N);
  DSF_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && lda >= K && ldb >= K && ldc >= N, "gemm_bf16_nt: bad leading dimensions");
  DSF_REQUIRE(aligned16(A) && aligned16(B) && aligned16(C) && aligned16(bias) && aligned16(residual), "gemm_bf16_nt: 16-byte alignment required");
  DSF_REQUIRE(c_dtype == DSF_F32 || c_dtype == DSF_BF16, "gemm_bf16_nt: bad c_dtype %d", c_dtype);
  DSF_REQUIRE(!(epi_flags & DSF_EPI_BIAS) || bias, "gemm_bf16_nt: bias flag without bias pointer");
  DSF_REQUIRE(!(epi_flags & DSF_EPI_RESIDUAL) || residual, "gemm_bf16_nt: residual flag without residual pointer");
  DSF_REQUIRE(!(epi_flags & DSF_EPI_ACCUM), "gemm_bf16_nt: ACCUM is not supported on the NT path");
  DSF_REQUIRE(!drop || (drop->p >= 0.f && drop->p < 1.f), "gemm_bf16_nt: dropout p must be in [0, 1)");
  DSF_REQUIRE(!relu_src || aligned16(relu_src), "gemm_bf16_nt: relu_src must be 16-byte aligned");
  if (g_gemm_impl != 1) return gemm_nt_v2(A, lda, B, ldb, C, ldc, c_dtype, bias, residual, M, N, K, epi_flags, drop, relu_src, (cudaStream_t)stream);
  if ((drop && drop->p > 0.f) || relu_src) { set_error("gemm_bf16_nt: dropout / relu-mask epilogues are only implemented in the v2 kernels"); return DSF_EUNSUPPORTED; }
  const int BN = (N % 128 == 0) ? 128 : 64;
  CUtensorMap tmA, tmB;
  if (int e = make_tmap_bf16(&tmA, A, M, K, lda, GT_BM)) return e;
  if (int e = make_tmap_bf16(&tmB, B, N, K, ldb, BN)) return e;
  EpiArgs epi{C, ldc, c_dtype, bias, residual, epi_flags};
  cudaStream_t st = (cudaStream_t)stream;
  if (BN == 128) return launch_nt<128, 3>(tmA, tmB, epi, M, N, K, st);
  return launch_nt<64, 4>(tmA, tmB, epi, M, N, K, st);
}

extern "C" int dsf_gemm_bf16_nt_ln(const void* A, int32_t lda, const void* B, int32_t ldb, float* C, int32_t ldc, const float* bias,
                                   const float* residual, void* H, int32_t ldh, const float* gamma, const float* beta, float* mean,
                                   float* rstd, float eps, int32_t M, int32_t N, int32_t K, const dsf_dropout* drop, void* stream) {
  DSF_REQUIRE(A && B && C && H && gamma && beta && mean && rstd, "gemm_bf16_nt_ln: NULL pointer");
  DSF_REQUIRE(M > 0 && K > 0, "gemm_bf16_nt_ln: non-positive extent");
  DSF_REQUIRE(N == 512, "gemm_bf16_nt_ln: the fused LayerNorm epilogue needs N = 512 (one CTA pair owns full rows); got N=%d", N);
  DSF_REQUIRE(K % GT_BK == 0, "gemm_bf16_nt_ln: K=%d must be a multiple of 64", K);
  DSF_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && ldh % 8 == 0 && lda >= K && ldb >= K && ldc >= N && ldh >= N,
              "gemm_bf16_nt_ln: bad leading dimensions");
  DSF_REQUIRE(aligned16(A) && aligned16(B) && aligned16(C) && aligned16(H) && aligned16(bias) && aligned16(residual) && aligned16(gamma) &&
                  aligned16(beta),
              "gemm_bf16_nt_ln: 16-byte alignment required");
  DSF_REQUIRE(!drop || (drop->p >= 0.f && drop->p < 1.f), "gemm_bf16_nt_ln: dropout p must be in [0, 1)");
  return gemm_nt_ln(A, lda, B, ldb, C, ldc, bias, residual, H, ldh, gamma, beta, mean, rstd, eps, M, K, drop, (cudaStream_t)stream);
}

extern "C" int dsf_gemm_bf16_tn(const void* A, int32_t lda, const void* B, int32_t ldb, float* C, int32_t ldc, int32_t M,
                                int32_t Nout, int32_t Kout, void* stream) {
  DSF_REQUIRE(A && B && C, "gemm_bf16_tn: NULL pointer");
  DSF_REQUIRE(M > 0 && Nout > 0 && Kout > 0, "gemm_bf16_tn: non-positive extent");
  DSF_REQUIRE(Nout % 64 == 0 && Kout % 64 == 0, "gemm_bf16_tn: output extents (%d, %d) must be multiples of 64", Nout, Kout);
  DSF_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= Nout && ldb >= Kout && ldc >= Kout, "gemm_bf16_tn: bad leading dimensions");
  DSF_REQUIRE(aligned16(A) && aligned16(B) && aligned16(C), "gemm_bf16_tn: 16-byte alignment required");
  if (g_gemm_impl != 1) return gemm_tn_v2(A, lda, B, ldb, C, ldc, M, Nout, Kout, (cudaStream_t)stream);
  const int BN = (Kout % 128 == 0) ? 128 : 64;
  CUtensorMap tmA, tmB;
  if (int e = make_tmap_bf16(&tmA, A, M, Nout, lda, GT_BK)) return e;
  if (int e = make_tmap_bf16(&tmB, B, M, Kout, ldb, GT_BK)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (BN == 128) return launch_tn<128, 3>(tmA, tmB, C, ldc, M, Nout, Kout, st);
  return launch_tn<64, 4>(tmA, tmB, C, ldc, M, Nout, Kout, st);
}

// fp32 parity path and small helpers (plain SIMT kernels, no tensor cores):
//   - dsf_gemm_f32: generic strided, two-level batched FFMA GEMM (every nn.Linear and both attention
//     contractions of model2_seq.py:97-109,121-126 in the fp32 mode; fp32 is the <=1e-3 parity mode,
//     bf16/tcgen05 is the performance mode)
//   - dsf_softmax_fwd / bwd over materialised scores (model2_seq.py:103)
//   - dsf_colsum (bias gradients), dsf_relu_bwd, dsf_cast_f32_bf16
#include <algorithm>

#include "common.cuh"

namespace dsf {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

__global__ void __launch_bounds__(SG_THREADS)
gemm_f32_kernel(dsf_gemm_f32_desc d, const float* __restrict__ A, const float* __restrict__ B, float* C,
                const float* __restrict__ bias, const float* residual) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
  const int t = threadIdx.x;
  const int b1 = blockIdx.z / d.nb2, b2 = blockIdx.z % d.nb2;
  A += b1 * d.a_b1 + b2 * d.a_b2;
  B += b1 * d.b_b1 + b2 * d.b_b2;
  const int64_t c_off = b1 * d.c_b1 + b2 * d.c_b2;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const bool a_kc = (d.a_k == 1), b_kc = (d.b_k == 1);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ty = t / 16, tx = t % 16;
  for (int k0 = 0; k0 < d.K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (a_kc) { k = t % SG_BK; m = t / SG_BK + 16 * i; }
      else { m = t % SG_BM; k = t / SG_BM + 4 * i; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < d.M && gk < d.K) ? A[gm * d.a_m + gk * d.a_k] : 0.f;
      int n, kb;
      if (b_kc) { kb = t % SG_BK; n = t / SG_BK + 16 * i; }
      else { n = t % SG_BN; kb = t / SG_BN + 4 * i; }
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < d.N && gkb < d.K) ? B[gn * d.b_n + gkb * d.b_k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= d.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= d.N) continue;
      float v = d.alpha * acc[i][j];
      if (d.epi_flags & DSF_EPI_BIAS) v += bias[gn];
      if (d.epi_flags & DSF_EPI_RELU) v = fmaxf(v, 0.f);
      const int64_t off = c_off + gm * d.c_m + gn * d.c_n;
      if (d.epi_flags & DSF_EPI_RESIDUAL) v += residual[off];
      if (d.epi_flags & DSF_EPI_ACCUM) v += C[off];
      C[off] = v;
    }
  }
}

// one warp per row; in place
__global__ void __launch_bounds__(256) softmax_fwd_kernel(float* __restrict__ s, int64_t rows, int T) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * 8) {
    float* r = s + row * T;
    float mx = -INFINITY;
    for (int i = lane; i < T; i += 32) mx = fmaxf(mx, r[i]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < T; i += 32) { const float e = expf(r[i] - mx); r[i] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int i = lane; i < T; i += 32) r[i] *= inv;
  }
}

__global__ void __launch_bounds__(256) softmax_bwd_kernel(float* __restrict__ dp, const float* __restrict__ p, int64_t rows, int T) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * 8) {
    float* d = dp + row * T;
    const float* pr = p + row * T;
    float dot = 0.f;
    for (int i = lane; i < T; i += 32) dot += d[i] * pr[i];
    dot = warp_sum(dot);
    for (int i = lane; i < T; i += 32) d[i] = pr[i] * (d[i] - dot);
  }
}

// out[n] += sum_m X[m, n]; block = 32 lanes x 8 row phases, each lane owns 2 adjacent columns
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, int ldx, float* __restrict__ out, int M, int N, int rows_per_block) {
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5;
  const int n = blockIdx.x * 64 + lane * 2;
  const int m_lo = blockIdx.y * rows_per_block, m_hi = min(M, m_lo + rows_per_block);
  float a0 = 0.f, a1 = 0.f;
  if (n < N) {
    for (int m = m_lo + ph; m < m_hi; m += 8) {
      const T* p = X + (size_t)m * ldx + n;
      a0 += to_f<T>(p[0]);
      if (n + 1 < N) a1 += to_f<T>(p[1]);
    }
  }
  red[ph][lane * 2] = a0;
  red[ph][lane * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    const int nn = blockIdx.x * 64 + threadIdx.x;
    if (nn < N) atomicAdd(out + nn, s);
  }
}

// Vectorised variant (N and ldx multiples of 16 B / sizeof(T), 16-byte aligned base): each lane owns one 16-byte column
// group, a warp covers 32 of them per row, the 8 warps of a CTA interleave rows with 4 independent loads in flight.
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ X, int ldx, float* __restrict__ out, int M, int N, int rows_per_block) {
  constexpr int VEC = 16 / (int)sizeof(T);
  __shared__ float red[8][32 * VEC + 1];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + lane) * VEC;
  const int m_lo = blockIdx.y * rows_per_block, m_hi = min(M, m_lo + rows_per_block);
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  if (n < N) {
    const T* base = X + n;
    int m = m_lo + w;
    for (; m + 24 < m_hi; m += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(m + 8 * u) * ldx));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const T* e = reinterpret_cast<const T*>(&v[u]);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] += to_f<T>(e[k]);
      }
    }
    for (; m < m_hi; m += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)m * ldx));
      const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] += to_f<T>(e[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) red[w][lane * VEC + k] = acc[k];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += red[k][c];
    const int nn = blockIdx.x * 32 * VEC + c;
    if (nn < N) atomicAdd(out + nn, sum);
  }
}

template <typename T>
static void launch_colsum(const T* X, int ldx, float* out, int M, int N, cudaStream_t st) {
  constexpr int VEC = 16 / (int)sizeof(T);
  if (N % VEC == 0 && ldx % VEC == 0 && aligned16(X)) {
    const int col_blocks = cdiv(N, 32 * VEC);
    const int row_blocks = max(1, min(cdiv(M, 64), (num_sms() * 6) / col_blocks));
    const int rpb = cdiv(M, row_blocks);
    launch_pdl(colsum_vec_kernel<T>, dim3(col_blocks, cdiv(M, rpb)), dim3(256), 0, st, X, ldx, out, M, N, rpb);
    return;
  }
  const int row_blocks = max(1, min(cdiv(M, 64), (num_sms() * 4) / max(1, cdiv(N, 64))));
  const int rpb = cdiv(M, row_blocks);
  colsum_kernel<T><<<dim3(cdiv(N, 64), cdiv(M, rpb)), 256, 0, st>>>(X, ldx, out, M, N, rpb);
}

template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(T* __restrict__ dy, const T* __restrict__ h, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float d[4], hv[4];
    Vec4<T>::load(dy + i * 4, d);
    Vec4<T>::load(h + i * 4, hv);
#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = hv[k] > 0.f ? d[k] : 0.f;
    Vec4<T>::store(dy + i * 4, d);
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4];
    Vec4<float>::load(src + i * 4, v);
    Vec4<__nv_bfloat16>::store(dst + i * 4, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[n4 * 4 + threadIdx.x] = __float2bfloat16_rn(src[n4 * 4 + threadIdx.x]);
}

}  // namespace dsf

using namespace dsf;

extern "C" int dsf_gemm_f32(const dsf_gemm_f32_desc* d, const float* A, const float* B, float* C, const float* bias,
                            const float* residual, void* stream) {
  DSF_REQUIRE(d && A && B && C, "gemm_f32: NULL pointer");
  DSF_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0 && d->nb1 > 0 && d->nb2 > 0, "gemm_f32: non-positive extent");
  DSF_REQUIRE(!(d->epi_flags & DSF_EPI_BIAS) || bias, "gemm_f32: bias flag without bias pointer");
  DSF_REQUIRE(!(d->epi_flags & DSF_EPI_RESIDUAL) || residual, "gemm_f32: residual flag without residual pointer");
  DSF_REQUIRE((int64_t)d->nb1 * d->nb2 <= 65535, "gemm_f32: too many batches");
  dim3 grid(cdiv(d->N, SG_BN), cdiv(d->M, SG_BM), d->nb1 * d->nb2);
  gemm_f32_kernel<<<grid, SG_THREADS, 0, (cudaStream_t)stream>>>(*d, A, B, C, bias, residual);
  return check_launch("gemm_f32");
}

extern "C" int dsf_softmax_fwd(float* s, int64_t rows, int32_t T, void* stream) {
  DSF_REQUIRE(s && rows > 0 && T > 0, "softmax_fwd: bad arguments");
  const int blocks = (int)std::min<int64_t>(cdiv64(rows, 8), (int64_t)num_sms() * 16);
  softmax_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(s, rows, T);
  return check_launch("softmax_fwd");
}

extern "C" int dsf_softmax_bwd(float* dp, const float* p, int64_t rows, int32_t T, void* stream) {
  DSF_REQUIRE(dp && p && rows > 0 && T > 0, "softmax_bwd: bad arguments");
  const int blocks = (int)std::min<int64_t>(cdiv64(rows, 8), (int64_t)num_sms() * 16);
  softmax_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dp, p, rows, T);
  return check_launch("softmax_bwd");
}

extern "C" int dsf_colsum(const void* X, int32_t x_dtype, int32_t ldx, float* out, int32_t M, int32_t N, void* stream) {
  DSF_REQUIRE(X && out && M > 0 && N > 0 && ldx >= N, "colsum: bad arguments");
  DSF_REQUIRE(x_dtype == DSF_F32 || x_dtype == DSF_BF16, "colsum: bad dtype %d", x_dtype);
  if (x_dtype == DSF_F32) launch_colsum<float>((const float*)X, ldx, out, M, N, (cudaStream_t)stream);
  else launch_colsum<__nv_bfloat16>((const __nv_bfloat16*)X, ldx, out, M, N, (cudaStream_t)stream);
  return check_launch("colsum");
}

extern "C" int dsf_relu_bwd(void* dy, const void* h, int32_t dtype, int64_t n, void* stream) {
  DSF_REQUIRE(dy && h && n > 0 && n % 4 == 0, "relu_bwd: bad arguments (n must be a multiple of 4)");
  DSF_REQUIRE(aligned16(dy) && aligned16(h), "relu_bwd: 16-byte alignment required");
  DSF_REQUIRE(dtype == DSF_F32 || dtype == DSF_BF16, "relu_bwd: bad dtype %d", dtype);
  const int64_t n4 = n / 4;
  const int blocks = (int)std::min<int64_t>(cdiv64(n4, 256), (int64_t)num_sms() * 16);
  if (dtype == DSF_F32) relu_bwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)dy, (const float*)h, n4);
  else relu_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dy, (const __nv_bfloat16*)h, n4);
  return check_launch("relu_bwd");
}

extern "C" int dsf_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  DSF_REQUIRE(src && dst && n > 0, "cast: bad arguments");
  DSF_REQUIRE(aligned16(src) && aligned16(dst), "cast: 16-byte alignment required");
  const int blocks = (int)std::min<int64_t>(cdiv64(std::max<int64_t>(n / 4, 1), 256), (int64_t)num_sms() * 16);
  cast_f32_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  return check_launch("cast_f32_bf16");
}

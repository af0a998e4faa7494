// Small fused helpers around the tensor-core GEMMs (all HBM/L2-bound, one pass each):
//   - dsf_pack_block_weights: per transformer block, fp32 master weights -> bf16 shadows in the two
//     layouts the tcgen05 GEMMs want (plain [N,K] for forward, transposed [K,N] for the data gradient),
//     with query/key/value fused into one [3C, C] matrix (model2_seq.py:83-85) and their biases concatenated.
//   - dsf_relu_bwd_colsum: ReLU mask (model2_seq.py:123) applied to the fc2 data gradient in place and the
//     fc1 bias gradient (column sums of the masked tensor) in the same pass.
#include <algorithm>

#include "common.cuh"

namespace dsf {

struct PackJob {
  const float* src;       // [rows, cols] fp32 row-major
  __nv_bfloat16* dst;     // plain copy: element (r, c) at dst[(row_off + r) * cols + c]
  __nv_bfloat16* dst_t;   // transposed copy: element (r, c) at dst_t[c * ld_t + row_off + r]
  int rows, cols, row_off, ld_t;
  int tile0;              // first 32x32 tile index of this job in the launch
};
struct PackJobs {
  PackJob j[6];
  int n_tiles;
  const float* bq; const float* bk; const float* bv;  // biases, concatenated into bqkv
  float* bqkv;
  int C;
};

__global__ void __launch_bounds__(256) pack_block_weights_kernel(PackJobs p) {
  __shared__ float tile[32][33];
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x;
  if (t >= p.n_tiles) {  // last CTA: bias concat [query | key | value]
    for (int i = threadIdx.x; i < 3 * p.C; i += 256) {
      const int which = i / p.C, c = i % p.C;
      p.bqkv[i] = which == 0 ? p.bq[c] : (which == 1 ? p.bk[c] : p.bv[c]);
    }
    return;
  }
  int ji = 0;
#pragma unroll
  for (int k = 1; k < 6; ++k)
    if (t >= p.j[k].tile0) ji = k;
  const PackJob& J = p.j[ji];
  const int lt = t - J.tile0;
  const int tiles_c = J.cols / 32;
  const int r0 = (lt / tiles_c) * 32, c0 = (lt % tiles_c) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k;
    const float v = J.src[(size_t)(r0 + r) * J.cols + c0 + tx];
    tile[r][tx] = v;
    J.dst[(size_t)(J.row_off + r0 + r) * J.cols + c0 + tx] = __float2bfloat16_rn(v);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = ty + 8 * k;
    J.dst_t[(size_t)(c0 + c) * J.ld_t + J.row_off + r0 + tx] = __float2bfloat16_rn(tile[tx][c]);
  }
}

// dy <- dy * (h > 0); out[n] += sum_m dy[m, n].  Block = 32 lanes x 8 row phases, lane owns 2 adjacent columns.
__global__ void __launch_bounds__(256)
relu_bwd_colsum_kernel(__nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ h, float* __restrict__ out, int M, int N,
                       int rows_per_block) {
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5;
  const int n = blockIdx.x * 64 + lane * 2;
  const int m_lo = blockIdx.y * rows_per_block, m_hi = min(M, m_lo + rows_per_block);
  float a0 = 0.f, a1 = 0.f;
  if (n < N) {
#pragma unroll 4
    for (int m = m_lo + ph; m < m_hi; m += 8) {
      const size_t off = (size_t)m * N + n;
      const __nv_bfloat162 d = *reinterpret_cast<const __nv_bfloat162*>(dy + off);
      const __nv_bfloat162 hv = *reinterpret_cast<const __nv_bfloat162*>(h + off);
      const float d0 = __low2float(hv) > 0.f ? __low2float(d) : 0.f;
      const float d1 = __high2float(hv) > 0.f ? __high2float(d) : 0.f;
      *reinterpret_cast<__nv_bfloat162*>(dy + off) = __floats2bfloat162_rn(d0, d1);
      a0 += d0;
      a1 += d1;
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

}  // namespace dsf

using namespace dsf;

extern "C" int dsf_pack_block_weights(const float* wq, const float* wk, const float* wv, const float* wp, const float* w1,
                                      const float* w2, const float* bq, const float* bk, const float* bv, int32_t C, int32_t F,
                                      void* wqkv, void* wqkv_t, void* wp_b, void* wp_t, void* w1_b, void* w1_t, void* w2_b, void* w2_t,
                                      float* bqkv, void* stream) {
  DSF_REQUIRE(wq && wk && wv && wp && w1 && w2 && bq && bk && bv && wqkv && wqkv_t && wp_b && wp_t && w1_b && w1_t && w2_b && w2_t && bqkv,
              "pack_block_weights: NULL pointer");
  DSF_REQUIRE(C > 0 && F > 0 && C % 32 == 0 && F % 32 == 0, "pack_block_weights: C=%d and F=%d must be multiples of 32", C, F);
  PackJobs p;
  __nv_bfloat16* qkv = (__nv_bfloat16*)wqkv;
  __nv_bfloat16* qkv_t = (__nv_bfloat16*)wqkv_t;
  // fused [3C, C] in the order [query | key | value]; transposed shadow is [C, 3C]
  const float* srcs[6] = {wq, wk, wv, wp, w1, w2};
  __nv_bfloat16* dsts[6] = {qkv, qkv, qkv, (__nv_bfloat16*)wp_b, (__nv_bfloat16*)w1_b, (__nv_bfloat16*)w2_b};
  __nv_bfloat16* dsts_t[6] = {qkv_t, qkv_t, qkv_t, (__nv_bfloat16*)wp_t, (__nv_bfloat16*)w1_t, (__nv_bfloat16*)w2_t};
  const int rows[6] = {C, C, C, C, F, C}, cols[6] = {C, C, C, C, C, F};
  const int row_off[6] = {0, C, 2 * C, 0, 0, 0};
  const int ld_t[6] = {3 * C, 3 * C, 3 * C, C, F, C};
  int tiles = 0;
  for (int i = 0; i < 6; ++i) {
    p.j[i] = PackJob{srcs[i], dsts[i], dsts_t[i], rows[i], cols[i], row_off[i], ld_t[i], tiles};
    tiles += (rows[i] / 32) * (cols[i] / 32);
  }
  p.n_tiles = tiles;
  p.bq = bq; p.bk = bk; p.bv = bv; p.bqkv = bqkv; p.C = C;
  dsf::launch_pdl(pack_block_weights_kernel, dim3(tiles + 1), dim3(256), 0, (cudaStream_t)stream, p);
  return check_launch("pack_block_weights");
}

extern "C" int dsf_relu_bwd_colsum(void* dy, const void* h, float* out, int32_t M, int32_t N, void* stream) {
  DSF_REQUIRE(dy && h && out && M > 0 && N > 0 && N % 2 == 0, "relu_bwd_colsum: bad arguments (N must be even)");
  DSF_REQUIRE(aligned16(dy) && aligned16(h), "relu_bwd_colsum: 16-byte alignment required");
  const int row_blocks = std::max(1, std::min(cdiv(M, 64), (num_sms() * 8) / std::max(1, cdiv(N, 64))));
  const int rpb = cdiv(M, row_blocks);
  dim3 grid(cdiv(N, 64), cdiv(M, rpb));
  relu_bwd_colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dy, (const __nv_bfloat16*)h, out, M, N, rpb);
  return check_launch("relu_bwd_colsum");
}

// ---------------------------------------------------------------- embedding dropout (model2_seq.py:272) and its backward
namespace dsf {
__global__ void __launch_bounds__(256) dropout_inplace_kernel(float* __restrict__ x, int64_t n8, DropArgs a) {
  pdl_trigger();
  pdl_wait();
  a = resolve_drop(a);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v0[4], v1[4], m[8];
    Vec4<float>::load(x + i * 8, v0);
    Vec4<float>::load(x + i * 8 + 4, v1);
    drop_scale8(a, (uint64_t)i, m);
#pragma unroll
    for (int k = 0; k < 4; ++k) { v0[k] *= m[k]; v1[k] *= m[4 + k]; }
    Vec4<float>::store(x + i * 8, v0);
    Vec4<float>::store(x + i * 8 + 4, v1);
  }
}
}  // namespace dsf

extern "C" int dsf_dropout_inplace(float* x, int64_t n, const dsf_dropout* d, void* stream) {
  DSF_REQUIRE(x && n > 0 && n % 8 == 0, "dropout_inplace: bad arguments (n must be a positive multiple of 8)");
  DSF_REQUIRE(aligned16(x), "dropout_inplace: 16-byte alignment required");
  DSF_REQUIRE(!d || (d->p >= 0.f && d->p < 1.f), "dropout_inplace: p must be in [0, 1)");
  const dsf::DropArgs a = dsf::make_drop(d);
  if (a.thresh == 0) return DSF_OK;
  const int64_t n8 = n / 8;
  const int blocks = (int)std::min<int64_t>(dsf::cdiv64(n8, 256), (int64_t)dsf::num_sms() * 16);
  dsf::launch_pdl(dsf::dropout_inplace_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, n8, a);
  return dsf::check_launch("dropout_inplace");
}

// K2: warp-shuffle LayerNorm forward / backward (nn.LayerNorm(C), eps 1e-5; model2_seq.py:118-119,199,
// used :131-132,274).  One warp owns one token row held entirely in registers (C <= 1024), statistics
// with two-pass mean / centred variance in fp32; HBM-bound (reads x, writes y).
//
// Backward: dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) [+ dx_add], and per-CTA partial
// dgamma / dbeta column sums that are folded into the fp32 outputs with one atomicAdd per column per CTA.
// Because dx is the gradient of the residual stream, it is also (a) the dY of the preceding Linear's
// bias (model2_seq.py:109,124) and (b) a tensor-core GEMM operand: the kernel optionally emits the column
// sums of dx (bias gradient) and a bf16 copy in the same pass, saving two more passes over the tensor.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace dsf {

constexpr int LN_WARPS = 8;

// VPT = float4 vectors per lane; covers C <= 128*VPT
template <typename TY, int VPT>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     TY* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int M, int C, float eps) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 2;
  const float invC = 1.0f / (float)C;
  for (int row = blockIdx.x * LN_WARPS + warp; row < M; row += gridDim.x * LN_WARPS) {
    const float* xr = x + (size_t)row * C;
    float v[VPT][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        Vec4<float>::load(xr + vi * 4, v[i]);
        s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
      } else {
        v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
      }
    }
    const float mu = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float d = v[i][k] - mu; q += d * d; }
      }
    }
    const float rs = rsqrtf(warp_sum(q) * invC + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    TY* yr = y + (size_t)row * C;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float g4[4], b4[4], o[4];
        Vec4<float>::load(gamma + vi * 4, g4);
        Vec4<float>::load(beta + vi * 4, b4);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = (v[i][k] - mu) * rs * g4[k] + b4[k];
        Vec4<TY>::store(yr + vi * 4, o);
      }
    }
  }
}

// DROP is a template parameter: the Philox code costs ~13 registers, which would take the common p = 0 instantiation
// from 2 to 1 resident CTAs per SM.
template <typename TY, int VPT, bool EXTRA, bool DROP>
__global__ void __launch_bounds__(LN_WARPS * 32, (VPT <= 4 && !DROP) ? 2 : 1)
layernorm_bwd_kernel(const TY* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* dx_add, float* dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dx_bf16,
                     float* __restrict__ dx_colsum, DropArgs drop, int M, int C) {
  // column partials of the CTA's warps: [quantity (dgamma, dbeta, by-product sums)][warp][column]; quantities are
  // reduced one after another for C > 512 (static shared memory is capped at 48 KB), all at once otherwise
  constexpr int NQ = EXTRA ? 3 : 2;
  constexpr int QPP = VPT <= 4 ? NQ : 1;  // quantities per pass
  __shared__ float red[QPP][LN_WARPS][VPT * 128];
  pdl_trigger();
  pdl_wait();
  if (DROP) drop = resolve_drop(drop);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 2;
  const float invC = 1.0f / (float)C;
  float g4[VPT][4], dg[VPT][4], db[VPT][4], dc[EXTRA ? VPT : 1][4];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = lane + 32 * i;
#pragma unroll
    for (int k = 0; k < 4; ++k) { dg[i][k] = 0.f; db[i][k] = 0.f; g4[i][k] = 0.f; }
    if (EXTRA) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dc[EXTRA ? i : 0][k] = 0.f;
    }
    if (vi < nvec) Vec4<float>::load(gamma + vi * 4, g4[i]);
  }
  for (int row = blockIdx.x * LN_WARPS + warp; row < M; row += gridDim.x * LN_WARPS) {
    const float mu = mean[row], rs = rstd[row];
    const float* xr = x + (size_t)row * C;
    const TY* dyr = dy + (size_t)row * C;
    float xh[VPT][4], gy[VPT][4], ad[VPT][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float xv[4], dv[4];
        Vec4<float>::load(xr + vi * 4, xv);
        Vec4<TY>::load(dyr + vi * 4, dv);
        if (dx_add) Vec4<float>::load(dx_add + (size_t)row * C + vi * 4, ad[i]);  // issued early: independent of the reductions
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          xh[i][k] = (xv[k] - mu) * rs;
          gy[i][k] = dv[k] * g4[i][k];
          s1 += gy[i][k];
          s2 += gy[i][k] * xh[i][k];
          dg[i][k] += dv[k] * xh[i][k];
          db[i][k] += dv[k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { xh[i][k] = 0.f; gy[i][k] = 0.f; }
      }
    }
    const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
    float* dxr = dx + (size_t)row * C;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = rs * (gy[i][k] - m1 - xh[i][k] * m2);
        if (dx_add) {
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] += ad[i][k];
        }
        Vec4<float>::store(dxr + vi * 4, o);
        if (EXTRA) {
          if (DROP) {  // by-products are the gradient of the preceding Linear's PRE-dropout output
            float m[4];
            drop_scale4(drop, (uint64_t)row * nvec + vi, m);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] *= m[k];
          }
          if (dx_bf16) Vec4<__nv_bfloat16>::store(dx_bf16 + (size_t)row * C + vi * 4, o);
#pragma unroll
          for (int k = 0; k < 4; ++k) dc[EXTRA ? i : 0][k] += o[k];
        }
      }
    }
  }
  // CTA reduction of the column partials, then one atomic per column and quantity
#pragma unroll 1
  for (int pass = 0; pass < NQ / QPP; ++pass) {
    if (pass > 0) __syncthreads();
#pragma unroll
    for (int qq = 0; qq < QPP; ++qq) {
      const int qn = pass * QPP + qq;
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int vi = lane + 32 * i;
        const float* src = qn == 0 ? dg[i] : (qn == 1 ? db[i] : dc[EXTRA ? i : 0]);
        *reinterpret_cast<float4*>(&red[qq][warp][vi * 4]) = make_float4(src[0], src[1], src[2], src[3]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int qq = 0; qq < QPP; ++qq) {
      const int qn = pass * QPP + qq;
      float* out = qn == 0 ? dgamma : (qn == 1 ? dbeta : dx_colsum);
      if (out == nullptr) continue;
      for (int c = threadIdx.x; c < C; c += LN_WARPS * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) s += red[qq][w][c];
        atomicAdd(out + c, s);
      }
    }
  }
}

template <typename TY>
int launch_ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int C,
                  float eps, cudaStream_t st) {
  const int blocks = std::min(cdiv(M, LN_WARPS), num_sms() * 8);
  TY* yy = reinterpret_cast<TY*>(y);
#define DSF_LN_FWD(V) launch_pdl(layernorm_fwd_kernel<TY, V>, dim3(blocks), dim3(LN_WARPS * 32), 0, st, x, gamma, beta, yy, mean, rstd, M, C, eps)
  if (C <= 128) DSF_LN_FWD(1);
  else if (C <= 256) DSF_LN_FWD(2);
  else if (C <= 512) DSF_LN_FWD(4);
  else DSF_LN_FWD(8);
#undef DSF_LN_FWD
  return check_launch("layernorm_fwd");
}

template <typename TY, bool EXTRA>
int launch_ln_bwd(const void* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                  const float* dx_add, float* dx, float* dgamma, float* dbeta, void* dx_bf16, float* dx_colsum, DropArgs drop, int M, int C,
                  cudaStream_t st) {
  // few, fat CTAs: every CTA ends with 2-3 x C atomics
  static const int mult = getenv("DSF_LN_BWD_MULT") ? std::max(1, atoi(getenv("DSF_LN_BWD_MULT"))) : 4;
  const int blocks = std::min(cdiv(M, LN_WARPS * 2), num_sms() * mult);
  const TY* d = reinterpret_cast<const TY*>(dy);
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
#define DSF_LN_BWD1(V, D) launch_pdl(layernorm_bwd_kernel<TY, V, EXTRA, D>, dim3(blocks), dim3(LN_WARPS * 32), 0, st, d, x, gamma, mean, rstd, dx_add, dx, dgamma, dbeta, xb, dx_colsum, drop, M, C)
#define DSF_LN_BWD(V)                                   \
  do {                                                  \
    if (EXTRA && drop.thresh != 0) DSF_LN_BWD1(V, EXTRA); \
    else DSF_LN_BWD1(V, false);                         \
  } while (0)
  if (C <= 128) DSF_LN_BWD(1);
  else if (C <= 256) DSF_LN_BWD(2);
  else if (C <= 512) DSF_LN_BWD(4);
  else DSF_LN_BWD(8);
#undef DSF_LN_BWD
#undef DSF_LN_BWD1
  return check_launch("layernorm_bwd");
}

}  // namespace dsf

using namespace dsf;

extern "C" int dsf_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int32_t y_dtype,
                                 float* mean, float* rstd, int32_t M, int32_t C, float eps, void* stream) {
  DSF_REQUIRE(x && gamma && beta && y && mean && rstd, "layernorm_fwd: NULL pointer");
  DSF_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && C <= 1024, "layernorm_fwd: need M>0 and C multiple of 4 up to 1024 (got M=%d C=%d)", M, C);
  DSF_REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(beta) && aligned16(y), "layernorm_fwd: 16-byte alignment required");
  DSF_REQUIRE(y_dtype == DSF_F32 || y_dtype == DSF_BF16, "layernorm_fwd: bad y_dtype %d", y_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (y_dtype == DSF_F32) return launch_ln_fwd<float>(x, gamma, beta, y, mean, rstd, M, C, eps, st);
  return launch_ln_fwd<__nv_bfloat16>(x, gamma, beta, y, mean, rstd, M, C, eps, st);
}

extern "C" int dsf_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* gamma, const float* mean,
                                 const float* rstd, const float* dx_add, float* dx_out, float* dgamma, float* dbeta,
                                 void* dx_bf16, float* dx_colsum, const dsf_dropout* byprod_drop, int32_t M, int32_t C, void* stream) {
  DSF_REQUIRE(dy && x && gamma && mean && rstd && dx_out && dgamma && dbeta, "layernorm_bwd: NULL pointer");
  DSF_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && C <= 1024, "layernorm_bwd: need M>0 and C multiple of 4 up to 1024 (got M=%d C=%d)", M, C);
  DSF_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(gamma) && aligned16(dx_add) && aligned16(dx_out) && aligned16(dx_bf16),
              "layernorm_bwd: 16-byte alignment required");
  DSF_REQUIRE(dy_dtype == DSF_F32 || dy_dtype == DSF_BF16, "layernorm_bwd: bad dy_dtype %d", dy_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const bool extra = dx_bf16 != nullptr || dx_colsum != nullptr;
  DSF_REQUIRE(!byprod_drop || (byprod_drop->p >= 0.f && byprod_drop->p < 1.f), "layernorm_bwd: dropout p must be in [0, 1)");
  const DropArgs dr = make_drop(byprod_drop);
  if (dy_dtype == DSF_F32) {
    if (extra) return launch_ln_bwd<float, true>(dy, x, gamma, mean, rstd, dx_add, dx_out, dgamma, dbeta, dx_bf16, dx_colsum, dr, M, C, st);
    return launch_ln_bwd<float, false>(dy, x, gamma, mean, rstd, dx_add, dx_out, dgamma, dbeta, nullptr, nullptr, dr, M, C, st);
  }
  if (extra) return launch_ln_bwd<__nv_bfloat16, true>(dy, x, gamma, mean, rstd, dx_add, dx_out, dgamma, dbeta, dx_bf16, dx_colsum, dr, M, C, st);
  return launch_ln_bwd<__nv_bfloat16, false>(dy, x, gamma, mean, rstd, dx_add, dx_out, dgamma, dbeta, nullptr, nullptr, dr, M, C, st);
}

// Optimizer step of the training loop as ONE multi-tensor kernel:
//   torch.optim.AdamW.step()  (train2_seq.py:131, 539; decoupled weight decay, bias-corrected moments)
// + EMA.update()              (train2_seq.py:133-134, 315-320: shadow = (1 - decay) * param + decay * shadow)
// + the fp32 -> bf16 repack of the GPT weights into the layouts the tcgen05 GEMMs read (plain [N,K] and transposed [K,N],
//   query/key/value fused, biases concatenated) — what dsf_pack_block_weights otherwise does at the start of every forward.
// Every parameter element is read once (p, g, m, v, ema) and written once (p, m, v, ema, shadows): the repack costs no extra
// pass over the weights.  HBM-bound: 4 B x (4 reads + 4 writes) + 4 B of bf16 shadows per GPT weight element.
#include <algorithm>

#include "common.cuh"

namespace dsf {

// 2-D tensors are processed in 32 x 32 tiles (so that the transposed bf16 shadow is written with full 64-byte rows through
// shared memory), everything else in chunks of 1024 contiguous elements.
__global__ void __launch_bounds__(256)
adamw_ema_pack_kernel(const dsf_opt_tensor* __restrict__ tab, const int32_t* __restrict__ tile0, int n_tensors, float lr, float beta1,
                      float beta2, float eps, float ema_decay, const int64_t* __restrict__ step_dev, float grad_scale, float om_beta1,
                      float om_beta2, float om_decay) {   // om_* = 1 - x rounded from double on the host, as torch computes them
  __shared__ float tile[32][33];
  __shared__ int s_ti;
  __shared__ float s_bc[2];
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x;
  if (threadIdx.x == 0) {  // largest ti with tile0[ti] <= t
    int lo = 0, hi = n_tensors - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile0[mid] <= t) lo = mid; else hi = mid - 1;
    }
    s_ti = lo;
    // bias corrections in double: 1 - beta2^t cancels badly in fp32 at small t (torch computes them in double on the host)
    const double step = (double)(*step_dev);
    s_bc[0] = (float)(1.0 - pow((double)beta1, step));
    s_bc[1] = (float)(1.0 - pow((double)beta2, step));
  }
  __syncthreads();
  const dsf_opt_tensor T = tab[s_ti];
  const int lt = t - tile0[s_ti];
  const float bc1 = s_bc[0], bc2 = s_bc[1];
  const float step_size = lr / bc1, sqrt_bc2 = sqrtf(bc2), decay_mul = 1.f - lr * T.weight_decay;

  auto update = [&](int64_t i) -> float {
    const float g = T.g[i] * grad_scale;
    float p = T.p[i] * decay_mul;
    const float m = beta1 * T.m[i] + om_beta1 * g;
    const float v = beta2 * T.v[i] + om_beta2 * g * g;
    p -= step_size * (m / (sqrtf(v) / sqrt_bc2 + eps));   // torch: addcdiv_(exp_avg, sqrt(v) / sqrt(bc2) + eps, -step_size)
    T.p[i] = p;
    T.m[i] = m;
    T.v[i] = v;
    if (T.ema) T.ema[i] = ema_decay * T.ema[i] + om_decay * p;
    return p;
  };

  if (T.shadow_t == nullptr) {  // flat chunk of 1024 elements (1-D parameters, tensors without a transposed shadow)
    const int64_t n = (int64_t)T.rows * T.cols;
    const int64_t base = (int64_t)lt * 1024;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = base + threadIdx.x + 256 * k;
      if (i < n) {
        const float p = update(i);
        if (T.shadow) reinterpret_cast<__nv_bfloat16*>(T.shadow)[(int64_t)T.row_off * T.cols + i] = __float2bfloat16_rn(p);
        if (T.copy_f32) T.copy_f32[i] = p;
      }
    }
    return;
  }
  const int tiles_c = T.cols / 32;
  const int r0 = (lt / tiles_c) * 32, c0 = (lt % tiles_c) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(T.shadow);
  __nv_bfloat16* sh_t = reinterpret_cast<__nv_bfloat16*>(T.shadow_t);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k;
    const float p = update((int64_t)(r0 + r) * T.cols + c0 + tx);
    tile[r][tx] = p;
    sh[(size_t)(T.row_off + r0 + r) * T.cols + c0 + tx] = __float2bfloat16_rn(p);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = ty + 8 * k;
    sh_t[(size_t)(c0 + c) * T.ld_t + T.row_off + r0 + tx] = __float2bfloat16_rn(tile[tx][c]);
  }
}

// 16-byte copy of the tensor table from PINNED HOST memory (device-accessible under unified addressing) into device memory by a
// kernel: unlike a host-to-device memcpy node it does not queue behind whatever large input prefetch occupies the copy engine.
__global__ void __launch_bounds__(256) opt_table_upload_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src_host, int n16) {
  pdl_trigger();
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src_host[i];
}

}  // namespace dsf

extern "C" int dsf_opt_upload_table(void* dst_dev, const void* src_pinned_host, int64_t nbytes, void* stream) {
  DSF_REQUIRE(dst_dev && src_pinned_host && nbytes > 0 && nbytes % 16 == 0, "opt_upload_table: bad arguments (nbytes must be a positive multiple of 16)");
  DSF_REQUIRE(dsf::aligned16(dst_dev) && dsf::aligned16(src_pinned_host), "opt_upload_table: 16-byte alignment required");
  const int n16 = (int)(nbytes / 16);
  dsf::launch_pdl(dsf::opt_table_upload_kernel, dim3(std::min(dsf::cdiv(n16, 256), 32)), dim3(256), 0, (cudaStream_t)stream, (uint4*)dst_dev,
                  (const uint4*)src_pinned_host, n16);
  return dsf::check_launch("opt_upload_table");
}

extern "C" int32_t dsf_opt_tiles(int32_t rows, int32_t cols, int32_t transposed_shadow) {
  if (rows <= 0 || cols <= 0) return 0;
  if (transposed_shadow) return (rows % 32 == 0 && cols % 32 == 0) ? (rows / 32) * (cols / 32) : -1;
  return (int32_t)(((int64_t)rows * cols + 1023) / 1024);
}

extern "C" int dsf_adamw_ema_pack(const dsf_opt_tensor* tensors_dev, const int32_t* tile0_dev, int32_t n_tensors, int32_t n_tiles, double lr,
                                  double beta1, double beta2, double eps, double ema_decay, const int64_t* step_dev, double grad_scale,
                                  void* stream) {
  DSF_REQUIRE(tensors_dev && tile0_dev && step_dev, "adamw_ema_pack: NULL table / step pointer");
  DSF_REQUIRE(n_tensors > 0 && n_tiles > 0, "adamw_ema_pack: empty launch (n_tensors=%d n_tiles=%d)", n_tensors, n_tiles);
  DSF_REQUIRE(lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps > 0. && ema_decay >= 0. && ema_decay <= 1.,
              "adamw_ema_pack: bad hyper-parameters");
  dsf::launch_pdl(dsf::adamw_ema_pack_kernel, dim3(n_tiles), dim3(256), 0, (cudaStream_t)stream, tensors_dev, tile0_dev, (int)n_tensors, (float)lr,
                  (float)beta1, (float)beta2, (float)eps, (float)ema_decay, step_dev, (float)grad_scale, (float)(1.0 - beta1), (float)(1.0 - beta2),
                  (float)(1.0 - ema_decay));
  return dsf::check_launch("adamw_ema_pack");
}

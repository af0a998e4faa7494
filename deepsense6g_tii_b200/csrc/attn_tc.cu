// K4: fused flash-style multi-head self-attention (no mask) on tcgen05 with S and O accumulators in TMEM.
// Replaces model2_seq.py:102-106 (q@k^T * 1/sqrt(hs), softmax, att@v, transpose+contiguous) and its
// autograd; the (B, nh, T, T) score tensor is never materialised.
//
// Layout: qkv (B, T, 3C) bf16 = [q | k | v] per token (the fused QKV GEMM output), head h at columns
// h*hs of each third.  Tiles are staged by the CTA's threads with 16-byte loads into the canonical
// no-swizzle UMMA core-matrix layout [row/8][col/8][row%8][col%8] (8 rows x 16 B = one 128-byte core
// matrix), which lets one shared-memory image serve both as a K-major operand (rows = M/N) and as an
// MN-major operand (rows = K): only LBO/SBO in the descriptor change.
//
// Forward, per CTA = (128 queries, head, sample):  for each KV tile:  S = Q K^T (TMEM) -> online
// softmax in registers (one thread per query row, exp2 with lazy rescale) -> P (bf16, smem) ->
// O += P V (TMEM).  LSE is saved for the backward.
#include <algorithm>

#include "tc_common.cuh"

namespace dsf {

using namespace tc;

constexpr int AT_BQ = 128;       // query rows per CTA (= UMMA M)
constexpr int AT_THREADS = 128;  // one thread per query row / TMEM lane

// 16-byte chunk i of a [R x CC] bf16 tile lives at byte offset i*16; it holds row (i/(8*nc8))*8 + i%8,
// columns 8*((i/8) % nc8) .. +7  (nc8 = CC/8).
template <int CC>
__device__ __forceinline__ void chunk_coord(int i, int& row, int& col8) {
  constexpr int nc8 = CC / 8;
  row = (i / (8 * nc8)) * 8 + (i & 7);
  col8 = (i >> 3) % nc8;
}

// stage a [ROWS x HS] tile (rows r0.. of a (T, ld)-strided matrix) into smem; rows >= T are zero-filled
template <int ROWS, int HS>
__device__ __forceinline__ void stage_tile(uint8_t* smem, const __nv_bfloat16* __restrict__ g, int ld, int r0, int T) {
  constexpr int NCH = ROWS * HS / 8;
  for (int i = threadIdx.x; i < NCH; i += AT_THREADS) {
    int row, c8;
    chunk_coord<HS>(i, row, c8);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + row < T) v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)(r0 + row) * ld + c8 * 8));
    *reinterpret_cast<uint4*>(smem + (size_t)i * 16) = v;
  }
}

// K-major operand over a [R x CC] tile: core matrices along K are 128 B apart, 8-row groups CC*16 B apart
template <int CC>
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) { return make_smem_desc(saddr, 128, CC * 16, SWZ_NONE); }
// MN-major operand over a [Krows x CC] tile: MN-adjacent core matrices 128 B apart (SBO), 8-row K groups CC*16 B apart (LBO)
template <int CC>
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) { return make_smem_desc(saddr, CC * 16, 128, SWZ_NONE); }

template <int HS, int BKV>
struct AttnFwdSmem {
  static constexpr int Q_BYTES = AT_BQ * HS * 2;
  static constexpr int K_BYTES = BKV * HS * 2;
  static constexpr int P_BYTES = AT_BQ * BKV * 2;
  static constexpr int Q_OFF = 0, K_OFF = Q_BYTES, V_OFF = K_OFF + K_BYTES, P_OFF = V_OFF + K_BYTES;
  static constexpr int BAR_OFF = P_OFF + P_BYTES;
  static constexpr int TOTAL = BAR_OFF + 32;
  static constexpr int DYN = TOTAL + 128;
  static constexpr uint32_t TMEM_COLS = (BKV + HS) <= 32 ? 32 : (BKV + HS) <= 64 ? 64 : (BKV + HS) <= 128 ? 128 : (BKV + HS) <= 256 ? 256 : 512;
};

template <int HS, int BKV>
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ y, float* __restrict__ lse, int T, int C, int nh,
                float scale_log2) {
  using L = AttnFwdSmem<HS, BKV>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_s = sbase + L::BAR_OFF, bar_o = bar_s + 8, tmem_slot = bar_s + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int ld = 3 * C;
  const __nv_bfloat16* qg = qkv + (size_t)b * T * ld + h * HS;
  const __nv_bfloat16* kg = qg + C;
  const __nv_bfloat16* vg = qg + 2 * C;

  if (threadIdx.x == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, L::TMEM_COLS); tmem_relinquish(); }
  stage_tile<AT_BQ, HS>(smem + L::Q_OFF, qg, ld, q0, T);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tm_s = tmem_base, tm_o = tmem_base + BKV;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  constexpr uint32_t idesc_s = make_idesc_bf16(AT_BQ, BKV, 0, 0);  // S = Q K^T, both K-major
  constexpr uint32_t idesc_o = make_idesc_bf16(AT_BQ, HS, 0, 1);   // O = P V, V is MN-major
  const uint64_t dq = desc_kmajor<HS>(sbase + L::Q_OFF);
  const uint64_t dk = desc_kmajor<HS>(sbase + L::K_OFF);
  const uint64_t dp = desc_kmajor<BKV>(sbase + L::P_OFF);
  const uint64_t dv = desc_mnmajor<HS>(sbase + L::V_OFF);

  float m_run = -INFINITY;  // running (stale-tolerant) max in log2 units
  float l_run = 0.f;
  const int n_kv = (T + BKV - 1) / BKV;
  const int my_row = warp * 32 + lane;

  for (int j = 0; j < n_kv; ++j) {
    const int kv0 = j * BKV;
    if (j > 0) mbar_wait(bar_o, (j - 1) & 1);  // previous P.V retired: K/V/P smem and O are free
    stage_tile<BKV, HS>(smem + L::K_OFF, kg, ld, kv0, T);
    stage_tile<BKV, HS>(smem + L::V_OFF, vg, ld, kv0, T);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HS / 16; ++k) tc_mma_bf16(tm_s, desc_advance(dq, k * 256), desc_advance(dk, k * 256), idesc_s, k != 0);
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, j & 1);
    tc_fence_after();

    // ---- softmax over this tile's BKV scores of my query row
    float p_max = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < BKV; c += 32) {
      uint32_t r[32];
      tmem_ld32(tm_s + lane_off + c, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (kv0 + c + i < T) p_max = fmaxf(p_max, __uint_as_float(r[i]));
    }
    p_max *= scale_log2;
    // lazy rescale: keep the stale max unless the new one exceeds it by more than 2^8
    const bool need = p_max > m_run + 8.f;
    if (__any_sync(0xffffffffu, need)) {
      const float m_new = need ? p_max : m_run;
      const float alpha = exp2f(m_run - m_new);  // m_run = -inf on the first tile -> alpha = 0 (O not yet valid)
      if (j > 0) {
#pragma unroll 1
        for (int c = 0; c < HS; c += 16) {
          uint32_t o[16];
          tmem_ld16(tm_o + lane_off + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tm_o + lane_off + c, o);
        }
        tmem_wait_st();
      }
      l_run *= alpha;
      m_run = m_new;
    }
    float l_add = 0.f;
#pragma unroll 1
    for (int c = 0; c < BKV; c += 32) {
      uint32_t r[32];
      tmem_ld32(tm_s + lane_off + c, r);
      tmem_wait_ld();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = (kv0 + c + i < T) ? exp2f(__uint_as_float(r[i]) * scale_log2 - m_run) : 0.f;
        float p1 = (kv0 + c + i + 1 < T) ? exp2f(__uint_as_float(r[i + 1]) * scale_log2 - m_run) : 0.f;
        // the row sum uses the bf16-rounded probabilities that the P.V MMA will see
        __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        l_add += __low2float(pb) + __high2float(pb);
        pk[i / 2] = *reinterpret_cast<uint32_t*>(&pb);
      }
      // P tile [128 x BKV] K-major core-matrix layout: chunk (row, c8) at ((row/8)*(BKV/8) + c8)*128 + (row%8)*16
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c8 = c / 8 + cc;
        uint8_t* dst = smem + L::P_OFF + ((size_t)(my_row >> 3) * (BKV / 8) + c8) * 128 + (my_row & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * cc], pk[4 * cc + 1], pk[4 * cc + 2], pk[4 * cc + 3]);
      }
    }
    l_run += l_add;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k)
        tc_mma_bf16(tm_o, desc_advance(dp, k * 256), desc_advance(dv, k * 2 * HS * 16), idesc_o, (j | k) != 0);
      tc_commit(bar_o);
    }
  }
  mbar_wait(bar_o, (n_kv - 1) & 1);
  tc_fence_after();
  const int q = q0 + my_row;
  const float inv_l = 1.0f / l_run;
  __nv_bfloat16* yrow = y + ((size_t)b * T + q) * C + h * HS;
#pragma unroll 1
  for (int c = 0; c < HS; c += 16) {
    uint32_t o[16];
    tmem_ld16(tm_o + lane_off + c, o);
    tmem_wait_ld();
    if (q < T) {
      uint4 w0 = make_uint4(pack_bf16x2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l));
      uint4 w1 = make_uint4(pack_bf16x2(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l));
      *reinterpret_cast<uint4*>(yrow + c) = w0;
      *reinterpret_cast<uint4*>(yrow + c + 8) = w1;
    }
  }
  if (q < T) lse[((size_t)b * nh + h) * T + q] = (m_run + log2f(l_run)) * 0.6931471805599453f;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, L::TMEM_COLS);
}

template <int HS, int BKV>
int launch_attn_fwd(const void* qkv, void* y, float* lse, int B, int T, int C, int nh, cudaStream_t st) {
  using L = AttnFwdSmem<HS, BKV>;
  static bool configured_on[64] = {};
  bool& configured = per_device_flag(configured_on);
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fwd_kernel<HS, BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN) != cudaSuccess)
      return check_launch("attn_fwd/attr");
    configured = true;
  }
  const float scale_log2 = (1.0f / sqrtf((float)HS)) * 1.4426950408889634f;
  dim3 grid(cdiv(T, AT_BQ), nh, B);
  attn_fwd_kernel<HS, BKV><<<grid, AT_THREADS, L::DYN, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)y, lse, T, C, nh, scale_log2);
  return check_launch("attn_fwd");
}


// ============================================================================================ backward
// delta[b,h,t] = sum_d dy[b,t,h*hs+d] * y[b,t,h*hs+d]   (rowsum(dO * O)); one warp per token row.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy, float* __restrict__ delta, int B, int T,
                  int C, int nh) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int hs = C / nh, gl = hs / 8;  // lanes per head (power of two, 2..16)
  const int nchunk = C / 8;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < (int64_t)B * T; row += (int64_t)gridDim.x * 8) {
    const int b = (int)(row / T), t = (int)(row % T);
    for (int c0 = 0; c0 < nchunk; c0 += 32) {
      const int ch = c0 + lane;
      float acc = 0.f;
      if (ch < nchunk) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(y + row * C + ch * 8));
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(dy + row * C + ch * 8));
        const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* dp = reinterpret_cast<const __nv_bfloat162*>(&d);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 af = __bfloat1622float2(ap[k]), df = __bfloat1622float2(dp[k]);
          acc += af.x * df.x + af.y * df.y;
        }
      }
      for (int o = gl >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (ch < nchunk && (lane % gl) == 0) delta[((size_t)b * nh + (ch * 8) / hs) * T + t] = acc;
    }
  }
}

// ---- kernel A: dK, dV.  CTA = (128 keys, head, sample); loop over query tiles of BQ rows.
//   S^T = K Q^T, dP^T = V dO^T (TMEM) -> P^T, dS^T (bf16, smem; thread = key row) -> dV += P^T dO, dK += dS^T Q.
template <int HS, int BQ>
struct AttnBwdKVSmem {
  static constexpr int KV_BYTES = 128 * HS * 2;
  static constexpr int Q_BYTES = BQ * HS * 2;
  static constexpr int P_BYTES = 128 * BQ * 2;
  static constexpr int K_OFF = 0, V_OFF = KV_BYTES, Q_OFF = 2 * KV_BYTES, DO_OFF = Q_OFF + Q_BYTES;
  static constexpr int P_OFF = DO_OFF + Q_BYTES, DS_OFF = P_OFF + P_BYTES;
  static constexpr int LSE_OFF = DS_OFF + P_BYTES, DELTA_OFF = LSE_OFF + BQ * 4;
  static constexpr int BAR_OFF = DELTA_OFF + BQ * 4;
  static constexpr int TOTAL = BAR_OFF + 32;
  static constexpr int DYN = TOTAL + 128;
  static constexpr int COLS = 2 * BQ + 2 * HS;
  static constexpr uint32_t TMEM_COLS = COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
};

template <int HS, int BQ>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_kv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dy, const float* __restrict__ lse,
                   const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T, int C, int nh, float scale) {
  using L = AttnBwdKVSmem<HS, BQ>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_s = sbase + L::BAR_OFF, bar_acc = bar_s + 8, tmem_slot = bar_s + 16;
  float* s_lse = reinterpret_cast<float*>(smem + L::LSE_OFF);
  float* s_delta = reinterpret_cast<float*>(smem + L::DELTA_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int ld = 3 * C;
  const __nv_bfloat16* qg = qkv + (size_t)b * T * ld + h * HS;
  const __nv_bfloat16* kg = qg + C;
  const __nv_bfloat16* vg = qg + 2 * C;
  const __nv_bfloat16* dog = dy + (size_t)b * T * C + h * HS;
  const float* lse_g = lse + ((size_t)b * nh + h) * T;
  const float* delta_g = delta + ((size_t)b * nh + h) * T;
  const float scale_log2 = scale * 1.4426950408889634f;

  if (threadIdx.x == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, L::TMEM_COLS); tmem_relinquish(); }
  stage_tile<128, HS>(smem + L::K_OFF, kg, ld, kv0, T);
  stage_tile<128, HS>(smem + L::V_OFF, vg, ld, kv0, T);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + BQ, tm_dv = tmem_base + 2 * BQ, tm_dk = tmem_base + 2 * BQ + HS;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  constexpr uint32_t idesc_s = make_idesc_bf16(128, BQ, 0, 0);   // S^T = K Q^T ; dP^T = V dO^T
  constexpr uint32_t idesc_g = make_idesc_bf16(128, HS, 0, 1);   // dV = P^T dO ; dK = dS^T Q  (B MN-major)
  const uint64_t dk_ = desc_kmajor<HS>(sbase + L::K_OFF);
  const uint64_t dv_ = desc_kmajor<HS>(sbase + L::V_OFF);
  const uint64_t dq_k = desc_kmajor<HS>(sbase + L::Q_OFF);
  const uint64_t ddo_k = desc_kmajor<HS>(sbase + L::DO_OFF);
  const uint64_t dq_mn = desc_mnmajor<HS>(sbase + L::Q_OFF);
  const uint64_t ddo_mn = desc_mnmajor<HS>(sbase + L::DO_OFF);
  const uint64_t dp_ = desc_kmajor<BQ>(sbase + L::P_OFF);
  const uint64_t dds_ = desc_kmajor<BQ>(sbase + L::DS_OFF);

  const int my_row = warp * 32 + lane;
  const bool key_ok = (kv0 + my_row) < T;
  const int n_q = (T + BQ - 1) / BQ;
  for (int i = 0; i < n_q; ++i) {
    const int q0 = i * BQ;
    if (i > 0) mbar_wait(bar_acc, (i - 1) & 1);  // previous dV/dK MMAs retired: Q/dO/P/dS smem free
    stage_tile<BQ, HS>(smem + L::Q_OFF, qg, ld, q0, T);
    stage_tile<BQ, HS>(smem + L::DO_OFF, dog, C, q0, T);
    for (int c = threadIdx.x; c < BQ; c += AT_THREADS) {
      const bool ok = (q0 + c) < T;
      s_lse[c] = ok ? lse_g[q0 + c] * 1.4426950408889634f : INFINITY;
      s_delta[c] = ok ? delta_g[q0 + c] : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HS / 16; ++k) tc_mma_bf16(tm_s, desc_advance(dk_, k * 256), desc_advance(dq_k, k * 256), idesc_s, k != 0);
#pragma unroll
      for (int k = 0; k < HS / 16; ++k) tc_mma_bf16(tm_dp, desc_advance(dv_, k * 256), desc_advance(ddo_k, k * 256), idesc_s, k != 0);
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, i & 1);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BQ; c += 32) {
      uint32_t rs[32], rp[32];
      tmem_ld32(tm_s + lane_off + c, rs);
      tmem_ld32(tm_dp + lane_off + c, rp);
      tmem_wait_ld();
      uint32_t pk[16], dk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        float p0 = key_ok ? exp2f(__uint_as_float(rs[e]) * scale_log2 - s_lse[c + e]) : 0.f;
        float p1 = key_ok ? exp2f(__uint_as_float(rs[e + 1]) * scale_log2 - s_lse[c + e + 1]) : 0.f;
        float d0 = p0 * (__uint_as_float(rp[e]) - s_delta[c + e]) * scale;
        float d1 = p1 * (__uint_as_float(rp[e + 1]) - s_delta[c + e + 1]) * scale;
        pk[e / 2] = pack_bf16x2(p0, p1);
        dk[e / 2] = pack_bf16x2(d0, d1);
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const size_t off = ((size_t)(my_row >> 3) * (BQ / 8) + c / 8 + cc) * 128 + (my_row & 7) * 16;
        *reinterpret_cast<uint4*>(smem + L::P_OFF + off) = make_uint4(pk[4 * cc], pk[4 * cc + 1], pk[4 * cc + 2], pk[4 * cc + 3]);
        *reinterpret_cast<uint4*>(smem + L::DS_OFF + off) = make_uint4(dk[4 * cc], dk[4 * cc + 1], dk[4 * cc + 2], dk[4 * cc + 3]);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < BQ / 16; ++k)
        tc_mma_bf16(tm_dv, desc_advance(dp_, k * 256), desc_advance(ddo_mn, k * 2 * HS * 16), idesc_g, (i | k) != 0);
#pragma unroll
      for (int k = 0; k < BQ / 16; ++k)
        tc_mma_bf16(tm_dk, desc_advance(dds_, k * 256), desc_advance(dq_mn, k * 2 * HS * 16), idesc_g, (i | k) != 0);
      tc_commit(bar_acc);
    }
  }
  mbar_wait(bar_acc, (n_q - 1) & 1);
  tc_fence_after();
  __nv_bfloat16* dk_row = dqkv + ((size_t)b * T + kv0 + my_row) * ld + C + h * HS;
  __nv_bfloat16* dv_row = dk_row + C;
#pragma unroll 1
  for (int c = 0; c < HS; c += 16) {
    uint32_t a[16], v[16];
    tmem_ld16(tm_dk + lane_off + c, a);
    tmem_ld16(tm_dv + lane_off + c, v);
    tmem_wait_ld();
    if (key_ok) {
#pragma unroll
      for (int e = 0; e < 16; e += 8) {
        *reinterpret_cast<uint4*>(dk_row + c + e) =
            make_uint4(pack_bf16x2(__uint_as_float(a[e]), __uint_as_float(a[e + 1])), pack_bf16x2(__uint_as_float(a[e + 2]), __uint_as_float(a[e + 3])),
                       pack_bf16x2(__uint_as_float(a[e + 4]), __uint_as_float(a[e + 5])), pack_bf16x2(__uint_as_float(a[e + 6]), __uint_as_float(a[e + 7])));
        *reinterpret_cast<uint4*>(dv_row + c + e) =
            make_uint4(pack_bf16x2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), pack_bf16x2(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])),
                       pack_bf16x2(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])), pack_bf16x2(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, L::TMEM_COLS);
}

// ---- kernel B: dQ.  CTA = (128 queries, head, sample); loop over KV tiles of 64 keys.
//   S = Q K^T, dP = dO V^T (TMEM) -> dS (bf16, smem; thread = query row) -> dQ += dS K.
constexpr int ATB_BKV = 64;
template <int HS>
struct AttnBwdQSmem {
  static constexpr int Q_BYTES = AT_BQ * HS * 2;
  static constexpr int KV_BYTES = ATB_BKV * HS * 2;
  static constexpr int DS_BYTES = AT_BQ * ATB_BKV * 2;
  static constexpr int Q_OFF = 0, DO_OFF = Q_BYTES, K_OFF = 2 * Q_BYTES, V_OFF = K_OFF + KV_BYTES, DS_OFF = V_OFF + KV_BYTES;
  static constexpr int BAR_OFF = DS_OFF + DS_BYTES;
  static constexpr int TOTAL = BAR_OFF + 32;
  static constexpr int DYN = TOTAL + 128;
  static constexpr int COLS = 2 * ATB_BKV + HS;
  static constexpr uint32_t TMEM_COLS = COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
};

template <int HS>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_q_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dy, const float* __restrict__ lse,
                  const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T, int C, int nh, float scale) {
  using L = AttnBwdQSmem<HS>;
  constexpr int BKV = ATB_BKV;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_s = sbase + L::BAR_OFF, bar_acc = bar_s + 8, tmem_slot = bar_s + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int ld = 3 * C;
  const __nv_bfloat16* qg = qkv + (size_t)b * T * ld + h * HS;
  const __nv_bfloat16* kg = qg + C;
  const __nv_bfloat16* vg = qg + 2 * C;
  const __nv_bfloat16* dog = dy + (size_t)b * T * C + h * HS;
  const float scale_log2 = scale * 1.4426950408889634f;

  if (threadIdx.x == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, L::TMEM_COLS); tmem_relinquish(); }
  stage_tile<AT_BQ, HS>(smem + L::Q_OFF, qg, ld, q0, T);
  stage_tile<AT_BQ, HS>(smem + L::DO_OFF, dog, C, q0, T);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + BKV, tm_dq = tmem_base + 2 * BKV;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  constexpr uint32_t idesc_s = make_idesc_bf16(AT_BQ, BKV, 0, 0);
  constexpr uint32_t idesc_q = make_idesc_bf16(AT_BQ, HS, 0, 1);
  const uint64_t dq_ = desc_kmajor<HS>(sbase + L::Q_OFF);
  const uint64_t ddo_ = desc_kmajor<HS>(sbase + L::DO_OFF);
  const uint64_t dk_k = desc_kmajor<HS>(sbase + L::K_OFF);
  const uint64_t dv_k = desc_kmajor<HS>(sbase + L::V_OFF);
  const uint64_t dk_mn = desc_mnmajor<HS>(sbase + L::K_OFF);
  const uint64_t dds_ = desc_kmajor<BKV>(sbase + L::DS_OFF);

  const int my_row = warp * 32 + lane;
  const int q = q0 + my_row;
  const bool q_ok = q < T;
  const float my_lse = q_ok ? lse[((size_t)b * nh + h) * T + q] * 1.4426950408889634f : INFINITY;
  const float my_delta = q_ok ? delta[((size_t)b * nh + h) * T + q] : 0.f;
  const int n_kv = (T + BKV - 1) / BKV;
  for (int j = 0; j < n_kv; ++j) {
    const int kv0 = j * BKV;
    if (j > 0) mbar_wait(bar_acc, (j - 1) & 1);
    stage_tile<BKV, HS>(smem + L::K_OFF, kg, ld, kv0, T);
    stage_tile<BKV, HS>(smem + L::V_OFF, vg, ld, kv0, T);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HS / 16; ++k) tc_mma_bf16(tm_s, desc_advance(dq_, k * 256), desc_advance(dk_k, k * 256), idesc_s, k != 0);
#pragma unroll
      for (int k = 0; k < HS / 16; ++k) tc_mma_bf16(tm_dp, desc_advance(ddo_, k * 256), desc_advance(dv_k, k * 256), idesc_s, k != 0);
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, j & 1);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BKV; c += 32) {
      uint32_t rs[32], rp[32];
      tmem_ld32(tm_s + lane_off + c, rs);
      tmem_ld32(tm_dp + lane_off + c, rp);
      tmem_wait_ld();
      uint32_t dk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        float p0 = (kv0 + c + e < T) ? exp2f(__uint_as_float(rs[e]) * scale_log2 - my_lse) : 0.f;
        float p1 = (kv0 + c + e + 1 < T) ? exp2f(__uint_as_float(rs[e + 1]) * scale_log2 - my_lse) : 0.f;
        dk[e / 2] = pack_bf16x2(p0 * (__uint_as_float(rp[e]) - my_delta) * scale, p1 * (__uint_as_float(rp[e + 1]) - my_delta) * scale);
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const size_t off = ((size_t)(my_row >> 3) * (BKV / 8) + c / 8 + cc) * 128 + (my_row & 7) * 16;
        *reinterpret_cast<uint4*>(smem + L::DS_OFF + off) = make_uint4(dk[4 * cc], dk[4 * cc + 1], dk[4 * cc + 2], dk[4 * cc + 3]);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k)
        tc_mma_bf16(tm_dq, desc_advance(dds_, k * 256), desc_advance(dk_mn, k * 2 * HS * 16), idesc_q, (j | k) != 0);
      tc_commit(bar_acc);
    }
  }
  mbar_wait(bar_acc, (n_kv - 1) & 1);
  tc_fence_after();
  __nv_bfloat16* dq_row = dqkv + ((size_t)b * T + q) * ld + h * HS;
#pragma unroll 1
  for (int c = 0; c < HS; c += 16) {
    uint32_t a[16];
    tmem_ld16(tm_dq + lane_off + c, a);
    tmem_wait_ld();
    if (q_ok) {
#pragma unroll
      for (int e = 0; e < 16; e += 8)
        *reinterpret_cast<uint4*>(dq_row + c + e) =
            make_uint4(pack_bf16x2(__uint_as_float(a[e]), __uint_as_float(a[e + 1])), pack_bf16x2(__uint_as_float(a[e + 2]), __uint_as_float(a[e + 3])),
                       pack_bf16x2(__uint_as_float(a[e + 4]), __uint_as_float(a[e + 5])), pack_bf16x2(__uint_as_float(a[e + 6]), __uint_as_float(a[e + 7])));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, L::TMEM_COLS);
}

template <int HS, int BQ>
int launch_attn_bwd(const void* qkv, const void* y, const void* dy, const float* lse, float* delta, void* dqkv, int B, int T, int C,
                    int nh, cudaStream_t st) {
  using LA = AttnBwdKVSmem<HS, BQ>;
  using LB = AttnBwdQSmem<HS>;
  static bool configured_on[64] = {};
  bool& configured = per_device_flag(configured_on);
  if (!configured) {
    if (cudaFuncSetAttribute(attn_bwd_kv_kernel<HS, BQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, LA::DYN) != cudaSuccess ||
        cudaFuncSetAttribute(attn_bwd_q_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, LB::DYN) != cudaSuccess)
      return check_launch("attn_bwd/attr");
    configured = true;
  }
  const float scale = 1.0f / sqrtf((float)HS);
  const int rows = B * T;
  attn_delta_kernel<<<std::min(cdiv(rows, 8), num_sms() * 8), 256, 0, st>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)dy, delta, B, T, C, nh);
  if (int e = check_launch("attn_bwd/delta")) return e;
  dim3 gridA(cdiv(T, 128), nh, B);
  attn_bwd_kv_kernel<HS, BQ><<<gridA, AT_THREADS, LA::DYN, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dy, lse, delta,
                                                                  (__nv_bfloat16*)dqkv, T, C, nh, scale);
  if (int e = check_launch("attn_bwd/kv")) return e;
  dim3 gridB(cdiv(T, AT_BQ), nh, B);
  attn_bwd_q_kernel<HS><<<gridB, AT_THREADS, LB::DYN, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dy, lse, delta,
                                                            (__nv_bfloat16*)dqkv, T, C, nh, scale);
  return check_launch("attn_bwd/q");
}

// v2 (warp-specialised, TMA-fed) implementations, attn_tc2.cu
int attn_fwd_v2(const void* qkv, void* y, float* lse, int B, int T, int C, int nh, cudaStream_t st);
int attn_fwd_v3(const void* qkv, void* y, float* lse, int B, int T, int C, int nh, const dsf_dropout* drop, uint32_t* bits, bool p_in_tmem,
                int force_nwg, cudaStream_t st);
int attn_fwd_v5(const void* qkv, void* y, float* lse, int B, int T, int C, int nh, const dsf_dropout* drop, uint32_t* bits, cudaStream_t st);
int attn_bwd_v2(const void* qkv, const void* y, const void* dy, const float* lse, float* delta, void* dqkv, int B, int T, int C, int nh,
                const dsf_dropout* drop, uint32_t* bits, bool p_in_tmem, int parts, cudaStream_t st);

int run_attn_delta(const void* y, const void* dy, float* delta, int B, int T, int C, int nh, cudaStream_t st) {
  launch_pdl(attn_delta_kernel, dim3(std::min(cdiv(B * T, 8), num_sms() * 8)), dim3(256), 0, st, (const __nv_bfloat16*)y, (const __nv_bfloat16*)dy,
             delta, B, T, C, nh);
  return check_launch("attn_bwd/delta");
}

// 0 = default (= 4), 1 = v1 (simple synchronous), 2 = v2 (pipelined), 3 = v3 (fwd: double-buffered S/P, P through shared memory),
// 4 = v4 (v3 schedule with P / dS kept in tensor memory: TS-mode tcgen05.mma, no shared-memory round trip),
// 5 = v4 with the four-warpgroup forward, 6 / 7 = v4 with the forward forced to 128-row CTAs (two per SM) / 256-row CTAs
static int g_attn_impl = 0;

}  // namespace dsf

using namespace dsf;

extern "C" int dsf_attn_set_impl(int32_t impl) {
  DSF_REQUIRE(impl >= 0 && impl <= 7, "attn_set_impl: impl must be 0 (default) or 1 .. 7");
  g_attn_impl = impl;
  return DSF_OK;
}

// attn_drop is implemented by the default (v3 forward / v2 backward) kernels only
static int check_attn_drop(const char* who, const dsf_dropout* drop, const uint32_t* bits, bool& on) {
  on = drop && drop->p > 0.f;
  if (!on) return DSF_OK;
  if (!(drop->p < 1.f)) { set_error("%s: dropout p must be in [0, 1)", who); return DSF_EINVAL; }
  if (!bits) { set_error("%s: attention dropout needs the drop_bits buffer", who); return DSF_EINVAL; }
  if (g_attn_impl == 1 || g_attn_impl == 2) { set_error("%s: attention dropout is only implemented in the v3 / v4 kernels", who); return DSF_EUNSUPPORTED; }
  return DSF_OK;
}

extern "C" int64_t dsf_attn_drop_words(int32_t B, int32_t T, int32_t nh) {
  if (B <= 0 || T <= 0 || nh <= 0) return 0;
  return (int64_t)B * nh * T * (2 * cdiv(T, 64));
}

extern "C" int dsf_attn_fwd(const void* qkv, void* y, float* lse, int32_t B, int32_t T, int32_t C, int32_t nh, const dsf_dropout* drop,
                            uint32_t* drop_bits, void* stream) {
  DSF_REQUIRE(qkv && y && lse, "attn_fwd: NULL pointer");
  bool drop_on;
  if (int e = check_attn_drop("attn_fwd", drop, drop_bits, drop_on)) return e;
  DSF_REQUIRE(B > 0 && T > 0 && C > 0 && nh > 0 && C % nh == 0, "attn_fwd: bad shape B=%d T=%d C=%d nh=%d", B, T, C, nh);
  DSF_REQUIRE(aligned16(qkv) && aligned16(y), "attn_fwd: 16-byte alignment required");
  DSF_REQUIRE(B <= 65535 && nh <= 65535, "attn_fwd: grid too large");
  const int hs = C / nh;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_attn_impl == 2) return attn_fwd_v2(qkv, y, lse, B, T, C, nh, st);
  if (g_attn_impl == 5) return attn_fwd_v5(qkv, y, lse, B, T, C, nh, drop_on ? drop : nullptr, drop_bits, st);
  if (g_attn_impl != 1)
    return attn_fwd_v3(qkv, y, lse, B, T, C, nh, drop_on ? drop : nullptr, drop_bits, g_attn_impl != 3, g_attn_impl == 6 ? 1 : (g_attn_impl == 7 ? 2 : 0), st);
  switch (hs) {
    case 16: return launch_attn_fwd<16, 128>(qkv, y, lse, B, T, C, nh, st);
    case 32: return launch_attn_fwd<32, 128>(qkv, y, lse, B, T, C, nh, st);
    case 64: return launch_attn_fwd<64, 128>(qkv, y, lse, B, T, C, nh, st);
    case 128: return launch_attn_fwd<128, 64>(qkv, y, lse, B, T, C, nh, st);
    default:
      set_error("attn_fwd: head size %d not supported (16, 32, 64, 128)", hs);
      return DSF_EUNSUPPORTED;
  }
}

static int attn_bwd_impl(const void* qkv, const void* y, const void* dy, const float* lse, float* delta, void* dqkv, int32_t B, int32_t T,
                         int32_t C, int32_t nh, const dsf_dropout* drop, const uint32_t* drop_bits, int parts, void* stream);

extern "C" int dsf_attn_bwd(const void* qkv, const void* y, const void* dy, const float* lse, float* delta, void* dqkv, int32_t B,
                            int32_t T, int32_t C, int32_t nh, const dsf_dropout* drop, const uint32_t* drop_bits, void* stream) {
  return attn_bwd_impl(qkv, y, dy, lse, delta, dqkv, B, T, C, nh, drop, drop_bits, 7, stream);
}

extern "C" int dsf_attn_bwd_parts(const void* qkv, const void* y, const void* dy, const float* lse, float* delta, void* dqkv, int32_t B,
                                  int32_t T, int32_t C, int32_t nh, const dsf_dropout* drop, const uint32_t* drop_bits, int32_t parts,
                                  void* stream) {
  DSF_REQUIRE(parts > 0 && parts <= 7, "attn_bwd_parts: parts must be a non-empty subset of {1 = delta, 2 = dK/dV, 4 = dQ}");
  DSF_REQUIRE(parts == 7 || g_attn_impl != 1, "attn_bwd_parts: the v1 kernels only run all three parts together");
  return attn_bwd_impl(qkv, y, dy, lse, delta, dqkv, B, T, C, nh, drop, drop_bits, parts, stream);
}

static int attn_bwd_impl(const void* qkv, const void* y, const void* dy, const float* lse, float* delta, void* dqkv, int32_t B, int32_t T,
                         int32_t C, int32_t nh, const dsf_dropout* drop, const uint32_t* drop_bits, int parts, void* stream) {
  DSF_REQUIRE(qkv && y && dy && lse && delta && dqkv, "attn_bwd: NULL pointer");
  bool drop_on;
  if (int e = check_attn_drop("attn_bwd", drop, drop_bits, drop_on)) return e;
  DSF_REQUIRE(B > 0 && T > 0 && C > 0 && nh > 0 && C % nh == 0, "attn_bwd: bad shape B=%d T=%d C=%d nh=%d", B, T, C, nh);
  DSF_REQUIRE(aligned16(qkv) && aligned16(y) && aligned16(dy) && aligned16(dqkv), "attn_bwd: 16-byte alignment required");
  DSF_REQUIRE(B <= 65535 && nh <= 65535, "attn_bwd: grid too large");
  const int hs = C / nh;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_attn_impl != 1)
    return attn_bwd_v2(qkv, y, dy, lse, delta, dqkv, B, T, C, nh, drop_on ? drop : nullptr, const_cast<uint32_t*>(drop_bits), g_attn_impl != 3 && g_attn_impl != 2, parts, st);
  switch (hs) {
    case 16: return launch_attn_bwd<16, 128>(qkv, y, dy, lse, delta, dqkv, B, T, C, nh, st);
    case 32: return launch_attn_bwd<32, 128>(qkv, y, dy, lse, delta, dqkv, B, T, C, nh, st);
    case 64: return launch_attn_bwd<64, 128>(qkv, y, dy, lse, delta, dqkv, B, T, C, nh, st);
    case 128: return launch_attn_bwd<128, 64>(qkv, y, dy, lse, delta, dqkv, B, T, C, nh, st);
    default:
      set_error("attn_bwd: head size %d not supported (16, 32, 64, 128)", hs);
      return DSF_EUNSUPPORTED;
  }
}

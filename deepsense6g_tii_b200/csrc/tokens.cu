// K1: fused adaptive-avgpool + token concat + pos-emb gather (forward) and its backward.
// Replaces model2_seq.py:515-517 (AdaptiveAvgPool2d x3) and :256-272 (view/cat/permute/contiguous/
// cat(gps)/+pos_emb) with one pass over the feature maps; HBM-bound (reads E_f, writes E_t).
//
// NCHW: one CTA owns (frame, CT-channel tile), CT = 32 for small planes and 8 for planes of >= 1024 pixels (stages 1-2:
// four times as many CTAs, i.e. enough 16-byte loads in flight to cover the HBM latency).  Lanes run along the contiguous
// W axis when reading the pooling windows: consecutive lanes load consecutive 16-byte chunks of a row (fully coalesced),
// the kw/4 lanes of a window combine their partial sums with shuffles; the pooled (cell, channel) tile is transposed
// through shared memory and written channel-contiguous with pos_emb added on the way out.
// NHWC: pure streaming; one thread owns 4 channels of one anchor cell.
#include <algorithm>

#include "common.cuh"

namespace dsf {

constexpr int TOK_CT = 32;       // largest channel tile (NCHW kernels)
constexpr int TOK_THREADS = 256;


struct FrameRef {
  const void* base;
  int n;  // frame index inside that tensor
};

__device__ __forceinline__ void frame_of(const dsf_geom& g, int b, int sl, const void* img, const void* lidar,
                                         const void* radar, const void*& base, int& n) {
  const int vs = g.V * g.S;
  if (sl < vs) { base = img; n = b * vs + sl; }
  else if (sl < vs + g.S) { base = lidar; n = b * g.S + (sl - vs); }
  else { base = radar; n = b * g.S + (sl - vs - g.S); }
}

// ------------------------------------------------------------------------------------ forward NCHW
template <typename FT, int CT>
__global__ void __launch_bounds__(TOK_THREADS)
tokens_fwd_nchw_kernel(dsf_geom g, const void* __restrict__ img, const void* __restrict__ lidar,
                       const void* __restrict__ radar, const float* __restrict__ gps,
                       const float* __restrict__ pos_emb, float* __restrict__ x) {
  extern __shared__ float sm[];  // [cells][CT + 1]
  const int cells = g.A_h * g.A_w;
  const int slots = (g.V + 2) * g.S;
  const int Tm = slots * cells, T = Tm + 2;
  const int F = g.B * slots;
  const int f = blockIdx.x;
  const int c0 = blockIdx.y * CT;
  const int nct = min(CT, g.C - c0);
  const int tid = threadIdx.x;
  if (f >= F) {  // GPS tokens of sample b (model2_seq.py:270)
    const int b = f - F;
    for (int i = tid; i < 2 * nct; i += TOK_THREADS) {
      const int j = i / nct, c = c0 + i % nct;
      x[((size_t)b * T + Tm + j) * g.C + c] = gps[((size_t)b * 2 + j) * g.C + c] + pos_emb[(size_t)(Tm + j) * g.C + c];
    }
    return;
  }
  const int b = f / slots, sl = f % slots;
  const void* basev; int n;
  frame_of(g, b, sl, img, lidar, radar, basev, n);
  const int HW = g.H * g.W;
  const FT* plane0 = reinterpret_cast<const FT*>(basev) + ((size_t)n * g.C + c0) * HW;
  const int kh = g.H / g.A_h, kw = g.W / g.A_w;
  const float inv = 1.0f / (float)(kh * kw);
  const bool vec = (kw % 4 == 0) && (g.W % 4 == 0);
  if (kh == 1 && kw == 1 && cells % 4 == 0) {
    // stage 4 (feature map == anchor grid): the (channel tile x cells) block is contiguous -> 16-byte loads, transposed
    // into shared memory
    for (int i = tid; i < nct * cells / 4; i += TOK_THREADS) {
      float v[4];
      Vec4<FT>::load(plane0 + 4 * i, v);
      const int cl = (4 * i) / cells, cell = (4 * i) % cells;
#pragma unroll
      for (int k = 0; k < 4; ++k) sm[(cell + k) * (CT + 1) + cl] = v[k];
    }
  } else if (vec && (kw / 4 == 1 || kw / 4 == 2 || kw / 4 == 4 || kw / 4 == 8) && ((nct * g.A_h * (g.W / 4)) % 32 == 0)) {
    // lanes along W: item = (channel, anchor row, 16-byte chunk of the row); the LW = kw/4 lanes of one window are adjacent
    const int W4 = g.W / 4, LW = kw / 4;
    const int items = nct * g.A_h * W4;
    for (int o = tid; o < items; o += TOK_THREADS) {
      const int j = o % W4, cy = (o / W4) % g.A_h, cl = o / (W4 * g.A_h);
      const FT* p = plane0 + (size_t)cl * HW + (size_t)(cy * kh) * g.W + 4 * j;
      float acc = 0.f;
#pragma unroll 4
      for (int r = 0; r < kh; ++r) {
        float v[4];
        Vec4<FT>::load(p + (size_t)r * g.W, v);
        acc += (v[0] + v[1]) + (v[2] + v[3]);
      }
      for (int d = 1; d < LW; d <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
      if (j % LW == 0) sm[(cy * g.A_w + j / LW) * (CT + 1) + cl] = acc * inv;
    }
  } else
  for (int o = tid; o < nct * cells; o += TOK_THREADS) {
    const int cl = o / cells, cell = o % cells;
    const int cy = cell / g.A_w, cx = cell % g.A_w;
    const FT* p = plane0 + (size_t)cl * HW + (size_t)(cy * kh) * g.W + cx * kw;
    float acc = 0.f;
    if (vec) {
      for (int r = 0; r < kh; ++r) {
        for (int q = 0; q < kw; q += 4) {
          float v[4];
          Vec4<FT>::load(p + (size_t)r * g.W + q, v);
          acc += v[0]; acc += v[1]; acc += v[2]; acc += v[3];
        }
      }
    } else {
      for (int r = 0; r < kh; ++r)
        for (int q = 0; q < kw; ++q) acc += to_f<FT>(p[(size_t)r * g.W + q]);
    }
    sm[cell * (CT + 1) + cl] = acc * inv;
  }
  __syncthreads();
  const int tok0 = sl * cells;
  if (nct == CT && g.C % 4 == 0) {  // 16-byte stores: 8 lanes cover the 32 channels of one token
    for (int o = tid; o < cells * (CT / 4); o += TOK_THREADS) {
      const int cell = o / (CT / 4), cl = (o % (CT / 4)) * 4;
      const int tok = tok0 + cell;
      float pe[4], v[4];
      Vec4<float>::load(pos_emb + (size_t)tok * g.C + c0 + cl, pe);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = sm[cell * (CT + 1) + cl + k] + pe[k];
      Vec4<float>::store(x + ((size_t)b * T + tok) * g.C + c0 + cl, v);
    }
    return;
  }
  for (int o = tid; o < cells * nct; o += TOK_THREADS) {
    const int cell = o / nct, cl = o % nct;
    const int tok = tok0 + cell;
    x[((size_t)b * T + tok) * g.C + c0 + cl] = sm[cell * (CT + 1) + cl] + pos_emb[(size_t)tok * g.C + c0 + cl];
  }
}

// ------------------------------------------------------------------------------------ forward NHWC
template <typename FT>
__global__ void __launch_bounds__(256)
tokens_fwd_nhwc_kernel(dsf_geom g, const void* __restrict__ img, const void* __restrict__ lidar,
                       const void* __restrict__ radar, const float* __restrict__ gps,
                       const float* __restrict__ pos_emb, float* __restrict__ x) {
  const int cells = g.A_h * g.A_w;
  const int slots = (g.V + 2) * g.S;
  const int Tm = slots * cells, T = Tm + 2;
  const int c4n = g.C / 4;
  const int64_t total = (int64_t)g.B * T * c4n;
  const int kh = g.H / g.A_h, kw = g.W / g.A_w;
  const float inv = 1.0f / (float)(kh * kw);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int64_t bt = i / c4n;
    const int tok = (int)(bt % T), b = (int)(bt / T);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (tok >= Tm) {
      Vec4<float>::load(gps + ((size_t)b * 2 + (tok - Tm)) * g.C + c, acc);
    } else {
      const int sl = tok / cells, cell = tok % cells;
      const int cy = cell / g.A_w, cx = cell % g.A_w;
      const void* basev; int n;
      frame_of(g, b, sl, img, lidar, radar, basev, n);
      const FT* p = reinterpret_cast<const FT*>(basev) + (((size_t)n * g.H + cy * kh) * g.W + cx * kw) * g.C + c;
      for (int r = 0; r < kh; ++r)
        for (int q = 0; q < kw; ++q) {
          float v[4];
          Vec4<FT>::load(p + ((size_t)r * g.W + q) * g.C, v);
          acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
        }
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] *= inv;
    }
    float pe[4];
    Vec4<float>::load(pos_emb + (size_t)tok * g.C + c, pe);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += pe[k];
    Vec4<float>::store(x + ((size_t)b * T + tok) * g.C + c, acc);
  }
}

// ------------------------------------------------------------------------------------ backward NCHW
template <typename FT, int CT>
__global__ void __launch_bounds__(TOK_THREADS)
tokens_bwd_nchw_kernel(dsf_geom g, const float* __restrict__ dx, const void* __restrict__ dres_img,
                       const void* __restrict__ dres_lidar, const void* __restrict__ dres_radar,
                       void* __restrict__ dimg, void* __restrict__ dlidar, void* __restrict__ dradar,
                       float* __restrict__ dgps) {
  extern __shared__ float sm[];  // [cells][CT + 1]
  const int cells = g.A_h * g.A_w;
  const int slots = (g.V + 2) * g.S;
  const int Tm = slots * cells, T = Tm + 2;
  const int F = g.B * slots;
  const int f = blockIdx.x;
  const int c0 = blockIdx.y * CT;
  const int nct = min(CT, g.C - c0);
  const int tid = threadIdx.x;
  if (f >= F) {
    const int b = f - F;
    for (int i = tid; i < 2 * nct; i += TOK_THREADS) {
      const int j = i / nct, c = c0 + i % nct;
      dgps[((size_t)b * 2 + j) * g.C + c] = dx[((size_t)b * T + Tm + j) * g.C + c];
    }
    return;
  }
  const int b = f / slots, sl = f % slots;
  const int kh = g.H / g.A_h, kw = g.W / g.A_w;
  const float inv = 1.0f / (float)(kh * kw);
  const int tok0 = sl * cells;
  if (nct == CT && g.C % 4 == 0) {  // 16-byte loads: 8 lanes cover the 32 channels of one token
    for (int o = tid; o < cells * (CT / 4); o += TOK_THREADS) {
      const int cell = o / (CT / 4), cl = (o % (CT / 4)) * 4;
      float v[4];
      Vec4<float>::load(dx + ((size_t)b * T + tok0 + cell) * g.C + c0 + cl, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) sm[cell * (CT + 1) + cl + k] = v[k] * inv;
    }
  } else {
    for (int o = tid; o < cells * nct; o += TOK_THREADS) {
      const int cell = o / nct, cl = o % nct;
      sm[cell * (CT + 1) + cl] = dx[((size_t)b * T + tok0 + cell) * g.C + c0 + cl] * inv;
    }
  }
  __syncthreads();
  const void* resv; int n;
  frame_of(g, b, sl, dres_img, dres_lidar, dres_radar, resv, n);
  void* outv; { const void* t; int n2; frame_of(g, b, sl, dimg, dlidar, dradar, t, n2); outv = const_cast<void*>(t); }
  const int HW = g.H * g.W;
  const size_t off0 = ((size_t)n * g.C + c0) * HW;
  const FT* res = resv ? reinterpret_cast<const FT*>(resv) + off0 : nullptr;
  FT* out = reinterpret_cast<FT*>(outv) + off0;
  if (g.W % 4 == 0) {
    const int W4 = g.W / 4;
    const int per = g.H * W4;
    const int total = nct * per;
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    for (int o0 = tid; o0 < total; o0 += U * TOK_THREADS) {
      float rv[U][4];
      size_t offs[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int o = o0 + u * TOK_THREADS;
        const int oc = o < total ? o : o0;
        const int cl = oc / per, r = oc % per;
        offs[u] = (size_t)cl * HW + (size_t)(r / W4) * g.W + (r % W4) * 4;
        if (res) Vec4<FT>::load(res + offs[u], rv[u]);
        else { rv[u][0] = rv[u][1] = rv[u][2] = rv[u][3] = 0.f; }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int o = o0 + u * TOK_THREADS;
        if (o >= total) break;
        const int cl = o / per, r = o % per;
        const int h = r / W4, w = (r % W4) * 4;
        const int cy = h / kh;
#pragma unroll
        for (int k = 0; k < 4; ++k) rv[u][k] += sm[(cy * g.A_w + (w + k) / kw) * (CT + 1) + cl];
        Vec4<FT>::store(out + offs[u], rv[u]);
      }
    }
  } else {
    for (int o = tid; o < nct * HW; o += TOK_THREADS) {
      const int cl = o / HW, r = o % HW;
      const int h = r / g.W, w = r % g.W;
      float v = sm[((h / kh) * g.A_w + w / kw) * (CT + 1) + cl];
      const size_t off = (size_t)cl * HW + r;
      if (res) v += to_f<FT>(res[off]);
      out[off] = from_f<FT>(v);
    }
  }
}

// ------------------------------------------------------------------------------------ backward NHWC
template <typename FT>
__global__ void __launch_bounds__(256)
tokens_bwd_nhwc_kernel(dsf_geom g, const float* __restrict__ dx, const void* __restrict__ dres_img,
                       const void* __restrict__ dres_lidar, const void* __restrict__ dres_radar,
                       void* __restrict__ dimg, void* __restrict__ dlidar, void* __restrict__ dradar,
                       float* __restrict__ dgps) {
  const int cells = g.A_h * g.A_w;
  const int slots = (g.V + 2) * g.S;
  const int Tm = slots * cells, T = Tm + 2;
  const int c4n = g.C / 4;
  const int kh = g.H / g.A_h, kw = g.W / g.A_w;
  const float inv = 1.0f / (float)(kh * kw);
  const int64_t n_map = (int64_t)g.B * slots * g.H * g.W * c4n;
  const int64_t n_gps = (int64_t)g.B * 2 * c4n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_map + n_gps; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= n_map) {
      const int64_t j = i - n_map;
      const int c = (int)(j % c4n) * 4;
      const int bj = (int)(j / c4n);
      const int b = bj / 2, jj = bj % 2;
      float v[4];
      Vec4<float>::load(dx + ((size_t)b * T + Tm + jj) * g.C + c, v);
      Vec4<float>::store(dgps + (size_t)bj * g.C + c, v);
      continue;
    }
    const int c = (int)(i % c4n) * 4;
    int64_t r = i / c4n;
    const int w = (int)(r % g.W); r /= g.W;
    const int h = (int)(r % g.H); r /= g.H;
    const int sl = (int)(r % slots), b = (int)(r / slots);
    const int tok = sl * cells + (h / kh) * g.A_w + w / kw;
    float v[4];
    Vec4<float>::load(dx + ((size_t)b * T + tok) * g.C + c, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] *= inv;
    const void* resv; int n;
    frame_of(g, b, sl, dres_img, dres_lidar, dres_radar, resv, n);
    const void* t; int n2;
    frame_of(g, b, sl, dimg, dlidar, dradar, t, n2);
    const size_t off = (((size_t)n * g.H + h) * g.W + w) * g.C + c;
    if (resv) {
      float rv[4];
      Vec4<FT>::load(reinterpret_cast<const FT*>(resv) + off, rv);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] += rv[k];
    }
    Vec4<FT>::store(reinterpret_cast<FT*>(const_cast<void*>(t)) + off, v);
  }
}

// dpos_emb[t,c] = sum_b dx[b,t,c]   (model2_seq.py:272 broadcast add)
__global__ void __launch_bounds__(256)
batch_sum_kernel(const float* __restrict__ dx, float* __restrict__ out, int B, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < B; ++b) {
      float v[4];
      Vec4<float>::load(dx + ((size_t)b * n4 + i) * 4, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += v[k];
    }
    Vec4<float>::store(out + i * 4, acc);
  }
}

int check_geom(const dsf_geom* g) {
  DSF_REQUIRE(g != nullptr, "geom is NULL");
  DSF_REQUIRE(g->B > 0 && g->S > 0 && g->V > 0 && g->A_h > 0 && g->A_w > 0, "geom: non-positive extent");
  DSF_REQUIRE(g->C > 0 && g->C % 4 == 0, "geom: C=%d must be a positive multiple of 4", g->C);
  DSF_REQUIRE(g->H > 0 && g->W > 0 && g->H % g->A_h == 0 && g->W % g->A_w == 0,
              "geom: feature map %dx%d is not a multiple of the anchor grid %dx%d", g->H, g->W, g->A_h, g->A_w);
  DSF_REQUIRE(g->feat_dtype == DSF_F32 || g->feat_dtype == DSF_BF16, "geom: bad feat_dtype %d", g->feat_dtype);
  DSF_REQUIRE(g->layout == DSF_NCHW || g->layout == DSF_NHWC, "geom: bad layout %d", g->layout);
  DSF_REQUIRE((size_t)g->A_h * g->A_w * (TOK_CT + 1) * 4 <= 160 * 1024, "geom: anchor grid too large");
  return DSF_OK;
}

}  // namespace dsf

using namespace dsf;

extern "C" int dsf_tokens_fwd(const dsf_geom* g, const void* img, const void* lidar, const void* radar,
                              const float* gps, const float* pos_emb, float* x, void* stream) {
  if (int e = check_geom(g)) return e;
  DSF_REQUIRE(img && lidar && radar && gps && pos_emb && x, "tokens_fwd: NULL pointer");
  DSF_REQUIRE(aligned16(img) && aligned16(lidar) && aligned16(radar) && aligned16(gps) && aligned16(pos_emb) && aligned16(x),
              "tokens_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int cells = g->A_h * g->A_w;
  const int slots = (g->V + 2) * g->S;
  if (g->layout == DSF_NCHW) {
    const int ct = nchw_channel_tile(g);
    dim3 grid(g->B * slots + g->B, cdiv(g->C, ct));
    size_t smem = (size_t)cells * (ct + 1) * sizeof(float);
#define DSF_TOK_FWD(FT, CT_)                                                                                                       \
  do {                                                                                                                             \
    if (smem > 48 * 1024) cudaFuncSetAttribute(tokens_fwd_nchw_kernel<FT, CT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    tokens_fwd_nchw_kernel<FT, CT_><<<grid, TOK_THREADS, smem, st>>>(*g, img, lidar, radar, gps, pos_emb, x);                       \
  } while (0)
    if (g->feat_dtype == DSF_F32) { if (ct == 8) DSF_TOK_FWD(float, 8); else DSF_TOK_FWD(float, 32); }
    else { if (ct == 8) DSF_TOK_FWD(__nv_bfloat16, 8); else DSF_TOK_FWD(__nv_bfloat16, 32); }
#undef DSF_TOK_FWD
  } else {
    const int64_t total = (int64_t)g->B * (slots * cells + 2) * (g->C / 4);
    int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)num_sms() * 16);
    if (g->feat_dtype == DSF_F32)
      tokens_fwd_nhwc_kernel<float><<<blocks, 256, 0, st>>>(*g, img, lidar, radar, gps, pos_emb, x);
    else
      tokens_fwd_nhwc_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(*g, img, lidar, radar, gps, pos_emb, x);
  }
  return check_launch("tokens_fwd");
}

extern "C" int dsf_tokens_bwd(const dsf_geom* g, const float* dx, const void* dres_img, const void* dres_lidar,
                              const void* dres_radar, void* dimg, void* dlidar, void* dradar, float* dgps,
                              float* dpos_emb, void* stream) {
  if (int e = check_geom(g)) return e;
  DSF_REQUIRE(dx && dimg && dlidar && dradar && dgps && dpos_emb, "tokens_bwd: NULL pointer");
  const bool any = dres_img || dres_lidar || dres_radar;
  DSF_REQUIRE(!any || (dres_img && dres_lidar && dres_radar), "tokens_bwd: dres_* must be all NULL or all set");
  DSF_REQUIRE(aligned16(dx) && aligned16(dimg) && aligned16(dlidar) && aligned16(dradar) && aligned16(dgps) && aligned16(dpos_emb) &&
              aligned16(dres_img) && aligned16(dres_lidar) && aligned16(dres_radar), "tokens_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int cells = g->A_h * g->A_w;
  const int slots = (g->V + 2) * g->S;
  const int T = slots * cells + 2;
  if (g->layout == DSF_NCHW) {
    const int ct = nchw_channel_tile(g);
    dim3 grid(g->B * slots + g->B, cdiv(g->C, ct));
    size_t smem = (size_t)cells * (ct + 1) * sizeof(float);
#define DSF_TOK_BWD(FT, CT_)                                                                                                       \
  do {                                                                                                                             \
    if (smem > 48 * 1024) cudaFuncSetAttribute(tokens_bwd_nchw_kernel<FT, CT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    tokens_bwd_nchw_kernel<FT, CT_><<<grid, TOK_THREADS, smem, st>>>(*g, dx, dres_img, dres_lidar, dres_radar, dimg, dlidar, dradar, dgps); \
  } while (0)
    if (g->feat_dtype == DSF_F32) { if (ct == 8) DSF_TOK_BWD(float, 8); else DSF_TOK_BWD(float, 32); }
    else { if (ct == 8) DSF_TOK_BWD(__nv_bfloat16, 8); else DSF_TOK_BWD(__nv_bfloat16, 32); }
#undef DSF_TOK_BWD
  } else {
    const int64_t total = (int64_t)g->B * slots * g->H * g->W * (g->C / 4) + (int64_t)g->B * 2 * (g->C / 4);
    int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)num_sms() * 16);
    if (g->feat_dtype == DSF_F32)
      tokens_bwd_nhwc_kernel<float><<<blocks, 256, 0, st>>>(*g, dx, dres_img, dres_lidar, dres_radar, dimg, dlidar, dradar, dgps);
    else
      tokens_bwd_nhwc_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(*g, dx, dres_img, dres_lidar, dres_radar, dimg, dlidar, dradar, dgps);
  }
  if (int e = check_launch("tokens_bwd")) return e;
  const int64_t n4 = (int64_t)T * g->C / 4;
  int blocks = (int)std::min<int64_t>(cdiv64(n4, 256), (int64_t)num_sms() * 8);
  batch_sum_kernel<<<blocks, 256, 0, st>>>(dx, dpos_emb, g->B, n4);
  return check_launch("tokens_bwd/pos_emb");
}

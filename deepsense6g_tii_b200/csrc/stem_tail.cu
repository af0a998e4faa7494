// Either side of the four fusion stages: the input stem and the pooled tail of Encoder.forward.
//
// dsf_stem_pack   normalize_imagenet (model2_seq.py:36-45; per-channel (x/255 - mean)/std, applied at :481-482) +
//                 torch.stack(frames, dim=1).view(B*S, C, H, W) (:491-493) + the dtype / layout change conv1 wants under
//                 autocast with channels_last weights (fp32 NCHW frames -> bf16 NHWC), as ONE pass: the stacked input of a
//                 trunk is written exactly once, in the form cuDNN consumes.  HBM-bound: reads 4 B, writes 2 (or 4) B per element.
// dsf_tail_fwd    AdaptiveAvgPool2d((1,1)) of the three stage-4 maps + flatten + view + cat with the GPS tokens + sum over the
//                 17 rows (:581-595): fused[b, c] = sum_{maps, frames} mean_px f[(b, t), c, :, :] + sum_j gps[b, j, c].
// dsf_tail_bwd    its autograd: every pixel of every frame of batch b receives dfused[b, c] / (H*W), the GPS tokens dfused[b, c].
#include <algorithm>

#include "common.cuh"

namespace dsf {

constexpr int STEM_MAX_FRAMES = 16;
struct StemFrames { const float* p[STEM_MAX_FRAMES]; };
struct StemAffine { float a[4], b[4]; };

// One thread = 4 consecutive pixels of one (batch, frame): float4 loads per channel plane (coalesced across the warp),
// output either NHWC (the 4 * CIN values of those pixels are contiguous) or NCHW (4 values per channel plane).
template <int CIN, typename OT, bool NHWC>
__global__ void __launch_bounds__(256)
stem_pack_kernel(StemFrames fr, StemAffine af, OT* __restrict__ out, int B, int S, int HW) {
  const int q = HW / 4;  // pixel quads per plane
  const int64_t total = (int64_t)B * S * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pq = (int)(i % q);
    const int64_t n = i / q;            // stacked frame index b * S + t  (torch.stack(dim=1).view(B*S, ...))
    const int t = (int)(n % S), b = (int)(n / S);
    const float* src = fr.p[t] + (size_t)b * CIN * HW + (size_t)pq * 4;
    float v[CIN][4];
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float4 x = __ldcs(reinterpret_cast<const float4*>(src + (size_t)c * HW));  // read once: streaming
      v[c][0] = fmaf(x.x, af.a[c], af.b[c]);
      v[c][1] = fmaf(x.y, af.a[c], af.b[c]);
      v[c][2] = fmaf(x.z, af.a[c], af.b[c]);
      v[c][3] = fmaf(x.w, af.a[c], af.b[c]);
    }
    if (NHWC) {
      OT* dst = out + ((size_t)n * HW + (size_t)pq * 4) * CIN;
      float w[4 * CIN];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < CIN; ++c) w[k * CIN + c] = v[c][k];
#pragma unroll
      for (int e = 0; e < 4 * CIN; e += 4) {  // 4 * CIN values: a multiple of 4 for every CIN
        const float t4[4] = {w[e], w[e + 1], w[e + 2], w[e + 3]};
        Vec4<OT>::store(dst + e, t4);
      }
    } else {
      OT* dst = out + (size_t)n * CIN * HW + (size_t)pq * 4;
#pragma unroll
      for (int c = 0; c < CIN; ++c) Vec4<OT>::store(dst + (size_t)c * HW, v[c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ tail
struct Tail3 { const void* p[3]; int frames[3]; };  // stage-4 maps of the image views / lidar / radar: (B * frames, C, H, W)
struct TailOut3 { void* p[3]; int frames[3]; };

// CTA = (batch b, 32-channel group); NCHW: warp w sums whole planes (c = c0 + lane is strided by HW -> one plane per warp
// iteration, lanes over pixels); NHWC: lanes over channels, warps over (frame, pixel).  fp32 accumulation, fixed order.
template <typename FT, bool NHWC>
__global__ void __launch_bounds__(256)
tail_fwd_kernel(Tail3 in, const float* __restrict__ gps, float* __restrict__ fused, int C, int HW) {
  __shared__ float part[8][33];
  const int b = blockIdx.x, c0 = blockIdx.y * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv = 1.0f / (float)HW;
  float acc = 0.f;  // NHWC: this lane's channel c0 + lane;  NCHW: partial of plane (c0 + cl) held by lane, reduced per plane
  if (NHWC) {
    for (int m = 0; m < 3; ++m) {
      const FT* base = reinterpret_cast<const FT*>(in.p[m]) + (size_t)b * in.frames[m] * HW * C;
      const int rows = in.frames[m] * HW;
      if (c0 + lane < C)
        for (int r = warp; r < rows; r += 8) acc += to_f<FT>(base[(size_t)r * C + c0 + lane]);
    }
    part[warp][lane] = acc * inv;
  } else {
    float mine = 0.f;  // lane cl keeps the finished sum of channel c0 + cl for the planes this warp visited
    for (int m = 0; m < 3; ++m) {
      const FT* base = reinterpret_cast<const FT*>(in.p[m]) + (size_t)b * in.frames[m] * C * HW;
      const int planes = in.frames[m] * 32;  // (frame, channel of this group)
      for (int pi = warp; pi < planes; pi += 8) {
        const int f = pi / 32, cl = pi % 32;
        if (c0 + cl >= C) continue;
        const FT* pl = base + ((size_t)f * C + c0 + cl) * HW;
        float s = 0.f;
        for (int i = lane; i < HW; i += 32) s += to_f<FT>(pl[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == cl) mine += s;
      }
    }
    part[warp][lane] = mine * inv;
  }
  __syncthreads();
  if (warp == 0 && c0 + lane < C) {
    float s = gps[((size_t)b * 2) * C + c0 + lane] + gps[((size_t)b * 2 + 1) * C + c0 + lane];
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][lane];
    fused[(size_t)b * C + c0 + lane] = s;
  }
}

template <typename FT, bool NHWC>
__global__ void __launch_bounds__(256)
tail_bwd_kernel(const float* __restrict__ dfused, TailOut3 out, float* __restrict__ dgps, int B, int C, int HW) {
  // one thread = 4 consecutive elements of one map (C*HW and HW are multiples of 4 in either layout)
  int64_t n4[3], tot = 0;
  for (int m = 0; m < 3; ++m) { n4[m] = (int64_t)B * out.frames[m] * C * HW / 4; tot += n4[m]; }
  const float inv = 1.0f / (float)HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot + (int64_t)B * 2 * C / 4; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= tot) {  // GPS tokens: dgps[b, j, c] = dfused[b, c]
      const int64_t e = (i - tot) * 4;
      const int c = (int)(e % C), b = (int)(e / ((int64_t)2 * C));
      float v[4];
      Vec4<float>::load(dfused + (size_t)b * C + c, v);
      Vec4<float>::store(dgps + e, v);
      continue;
    }
    int m = 0;
    int64_t j = i;
    while (j >= n4[m]) { j -= n4[m]; ++m; }
    const int64_t e = j * 4;
    const int64_t per_b = (int64_t)out.frames[m] * C * HW;
    const int b = (int)(e / per_b);
    const int64_t r = e % per_b;
    float v[4];
    if (NHWC) {
      const int c = (int)(r % C);  // 4 consecutive channels
      Vec4<float>::load(dfused + (size_t)b * C + c, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] *= inv;
    } else {
      const int c = (int)((r / HW) % C);  // 4 consecutive pixels of one channel plane
      const float g = dfused[(size_t)b * C + c] * inv;
      v[0] = v[1] = v[2] = v[3] = g;
    }
    Vec4<FT>::store(reinterpret_cast<FT*>(out.p[m]) + e, v);
  }
}

}  // namespace dsf

using namespace dsf;

extern "C" int dsf_stem_pack(const void* const* frames, int32_t n_frames, int32_t B, int32_t C_in, int32_t H, int32_t W,
                             const float* scale, const float* shift, void* out, int32_t out_dtype, int32_t out_layout, void* stream) {
  DSF_REQUIRE(frames && out, "stem_pack: NULL pointer");
  DSF_REQUIRE(n_frames > 0 && n_frames <= STEM_MAX_FRAMES, "stem_pack: 1..%d frames per call, got %d", STEM_MAX_FRAMES, n_frames);
  DSF_REQUIRE(B > 0 && H > 0 && W > 0 && C_in >= 1 && C_in <= 3, "stem_pack: bad shape B=%d C=%d H=%d W=%d (1..3 input channels)", B, C_in, H, W);
  DSF_REQUIRE(((int64_t)H * W) % 4 == 0, "stem_pack: H*W must be a multiple of 4");
  DSF_REQUIRE(out_dtype == DSF_F32 || out_dtype == DSF_BF16, "stem_pack: bad out_dtype %d", out_dtype);
  DSF_REQUIRE(out_layout == DSF_NCHW || out_layout == DSF_NHWC, "stem_pack: bad out_layout %d", out_layout);
  DSF_REQUIRE(aligned16(out), "stem_pack: 16-byte alignment required");
  StemFrames fr{};
  for (int i = 0; i < n_frames; ++i) {
    DSF_REQUIRE(frames[i] && aligned16(frames[i]), "stem_pack: frame %d is NULL or not 16-byte aligned", i);
    fr.p[i] = static_cast<const float*>(frames[i]);
  }
  StemAffine af{};
  for (int c = 0; c < C_in; ++c) { af.a[c] = scale ? scale[c] : 1.f; af.b[c] = shift ? shift[c] : 0.f; }  // HOST arrays of C_in floats
  const int HW = H * W;
  const int64_t total = (int64_t)B * n_frames * (HW / 4);
  const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
#define DSF_STEM(CIN, OT, NHWC_) stem_pack_kernel<CIN, OT, NHWC_><<<blocks, 256, 0, st>>>(fr, af, (OT*)out, B, n_frames, HW)
#define DSF_STEM_C(OT, NHWC_) do { if (C_in == 1) DSF_STEM(1, OT, NHWC_); else if (C_in == 2) DSF_STEM(2, OT, NHWC_); else DSF_STEM(3, OT, NHWC_); } while (0)
  if (out_dtype == DSF_BF16) { if (out_layout == DSF_NHWC) DSF_STEM_C(__nv_bfloat16, true); else DSF_STEM_C(__nv_bfloat16, false); }
  else { if (out_layout == DSF_NHWC) DSF_STEM_C(float, true); else DSF_STEM_C(float, false); }
#undef DSF_STEM_C
#undef DSF_STEM
  return check_launch("stem_pack");
}

static int check_tail(const char* who, int B, int C, int H, int W, int f0, int f1, int f2, int dtype, int layout) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || f0 <= 0 || f1 <= 0 || f2 <= 0) { set_error("%s: bad shape", who); return DSF_EINVAL; }
  if (C % 4 != 0 || ((int64_t)H * W) % 4 != 0) { set_error("%s: C and H*W must be multiples of 4", who); return DSF_EINVAL; }
  if ((dtype != DSF_F32 && dtype != DSF_BF16) || (layout != DSF_NCHW && layout != DSF_NHWC)) { set_error("%s: bad dtype / layout", who); return DSF_EINVAL; }
  return DSF_OK;
}

extern "C" int dsf_tail_fwd(const void* img, const void* lidar, const void* radar, const float* gps, float* fused, int32_t B,
                            int32_t frames_img, int32_t frames_lidar, int32_t frames_radar, int32_t C, int32_t H, int32_t W,
                            int32_t feat_dtype, int32_t layout, void* stream) {
  DSF_REQUIRE(img && lidar && radar && gps && fused, "tail_fwd: NULL pointer");
  if (int e = check_tail("tail_fwd", B, C, H, W, frames_img, frames_lidar, frames_radar, feat_dtype, layout)) return e;
  Tail3 in{{img, lidar, radar}, {frames_img, frames_lidar, frames_radar}};
  dim3 grid(B, cdiv(C, 32));
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  if (feat_dtype == DSF_F32) {
    if (layout == DSF_NHWC) tail_fwd_kernel<float, true><<<grid, 256, 0, st>>>(in, gps, fused, C, HW);
    else tail_fwd_kernel<float, false><<<grid, 256, 0, st>>>(in, gps, fused, C, HW);
  } else {
    if (layout == DSF_NHWC) tail_fwd_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(in, gps, fused, C, HW);
    else tail_fwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(in, gps, fused, C, HW);
  }
  return check_launch("tail_fwd");
}

extern "C" int dsf_tail_bwd(const float* dfused, void* dimg, void* dlidar, void* dradar, float* dgps, int32_t B, int32_t frames_img,
                            int32_t frames_lidar, int32_t frames_radar, int32_t C, int32_t H, int32_t W, int32_t feat_dtype,
                            int32_t layout, void* stream) {
  DSF_REQUIRE(dfused && dimg && dlidar && dradar && dgps, "tail_bwd: NULL pointer");
  if (int e = check_tail("tail_bwd", B, C, H, W, frames_img, frames_lidar, frames_radar, feat_dtype, layout)) return e;
  DSF_REQUIRE(aligned16(dfused) && aligned16(dimg) && aligned16(dlidar) && aligned16(dradar) && aligned16(dgps), "tail_bwd: 16-byte alignment required");
  TailOut3 out{{dimg, dlidar, dradar}, {frames_img, frames_lidar, frames_radar}};
  const int HW = H * W;
  const int64_t tot = (int64_t)B * (frames_img + frames_lidar + frames_radar) * C * HW / 4 + (int64_t)B * 2 * C / 4;
  const int blocks = (int)std::min<int64_t>(cdiv64(tot, 256), (int64_t)num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (feat_dtype == DSF_F32) {
    if (layout == DSF_NHWC) tail_bwd_kernel<float, true><<<blocks, 256, 0, st>>>(dfused, out, dgps, B, C, HW);
    else tail_bwd_kernel<float, false><<<blocks, 256, 0, st>>>(dfused, out, dgps, B, C, HW);
  } else {
    if (layout == DSF_NHWC) tail_bwd_kernel<__nv_bfloat16, true><<<blocks, 256, 0, st>>>(dfused, out, dgps, B, C, HW);
    else tail_bwd_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>(dfused, out, dgps, B, C, HW);
  }
  return check_launch("tail_bwd");
}

// K7: fused un-tokenise + bilinear upsample (align_corners=False) + residual add, forward and backward.
// Replaces model2_seq.py:275-286 (slice/view/permute/contiguous), :521-523 / :539-541 / :558-560
// (F.interpolate bilinear, scale 8/4/2; none at stage 4) and :524-526 etc. (feat + up).
// HBM-bound: forward reads E_t + E_f and writes E_f; backward reads E_f and writes E_t.
//
// Sampling rule (ATen area_pixel_compute_source_index, align_corners=False):
//   src = max((i + 0.5) / scale - 0.5, 0); i0 = floor(src); i1 = min(i0 + 1, A - 1); lam = src - i0.
#include <algorithm>

#include "common.cuh"

namespace dsf {

constexpr int UP_THREADS = 256;
constexpr int UP_WARPS = UP_THREADS / 32;

struct Tap { int i0, i1; float lam; };
__device__ __forceinline__ Tap tap_of(int i, float rscale, int A) {
  float s = fmaxf(((float)i + 0.5f) * rscale - 0.5f, 0.f);
  Tap t;
  t.i0 = min((int)s, A - 1);
  t.i1 = min(t.i0 + 1, A - 1);
  t.lam = s - (float)t.i0;
  return t;
}

__device__ __forceinline__ void slot_frame(const dsf_geom& g, int b, int sl, int& which, int& n) {
  const int vs = g.V * g.S;
  if (sl < vs) { which = 0; n = b * vs + sl; }
  else if (sl < vs + g.S) { which = 1; n = b * g.S + (sl - vs); }
  else { which = 2; n = b * g.S + (sl - vs - g.S); }
}

struct Ptr3 { const void* p[3]; };
struct MPtr3 { void* p[3]; };

// ------------------------------------------------------------------------------------ forward NCHW
template <typename FT, int UP_CT>
__global__ void __launch_bounds__(UP_THREADS)
upsample_add_fwd_nchw_kernel(dsf_geom g, const float* __restrict__ y, Ptr3 feat, MPtr3 out) {
  extern __shared__ float sm[];  // [cells][UP_CT + 1]
  const int cells = g.A_h * g.A_w;
  const int slots = (g.V + 2) * g.S;
  const int T = slots * cells + 2;
  const int f = blockIdx.x, c0 = blockIdx.y * UP_CT;
  const int nct = min(UP_CT, g.C - c0);
  const int tid = threadIdx.x;
  const int b = f / slots, sl = f % slots;
  if (nct == UP_CT && g.C % 4 == 0) {  // 16-byte loads: 8 lanes cover the 32 channels of one token
    for (int o = tid; o < cells * (UP_CT / 4); o += UP_THREADS) {
      const int cell = o / (UP_CT / 4), cl = (o % (UP_CT / 4)) * 4;
      float v[4];
      Vec4<float>::load(y + ((size_t)b * T + sl * cells + cell) * g.C + c0 + cl, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) sm[cell * (UP_CT + 1) + cl + k] = v[k];
    }
  } else {
    for (int o = tid; o < cells * nct; o += UP_THREADS) {
      const int cell = o / nct, cl = o % nct;
      sm[cell * (UP_CT + 1) + cl] = y[((size_t)b * T + sl * cells + cell) * g.C + c0 + cl];
    }
  }
  __syncthreads();
  int which, n;
  slot_frame(g, b, sl, which, n);
  const int HW = g.H * g.W;
  const size_t off0 = ((size_t)n * g.C + c0) * HW;
  const FT* fin = reinterpret_cast<const FT*>(feat.p[which]) + off0;
  FT* fout = reinterpret_cast<FT*>(out.p[which]) + off0;
  const float rsh = (float)g.A_h / (float)g.H, rsw = (float)g.A_w / (float)g.W;
  if (g.W % 4 == 0) {
    const int W4 = g.W / 4, per = g.H * W4, total = nct * per;
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    for (int o0 = tid; o0 < total; o0 += U * UP_THREADS) {
      float v[U][4];
      size_t offs[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int o = o0 + u * UP_THREADS;
        const int oc = o < total ? o : o0;
        const int cl = oc / per, r = oc % per;
        offs[u] = (size_t)cl * HW + (size_t)(r / W4) * g.W + (r % W4) * 4;
        Vec4<FT>::load(fin + offs[u], v[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int o = o0 + u * UP_THREADS;
        if (o >= total) break;
        const int cl = o / per, r = o % per;
        const int h = r / W4, w = (r % W4) * 4;
        const Tap ty = tap_of(h, rsh, g.A_h);
        const float* r0 = sm + (ty.i0 * g.A_w) * (UP_CT + 1) + cl;
        const float* r1 = sm + (ty.i1 * g.A_w) * (UP_CT + 1) + cl;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Tap tx = tap_of(w + k, rsw, g.A_w);
          const float top = (1.f - tx.lam) * r0[tx.i0 * (UP_CT + 1)] + tx.lam * r0[tx.i1 * (UP_CT + 1)];
          const float bot = (1.f - tx.lam) * r1[tx.i0 * (UP_CT + 1)] + tx.lam * r1[tx.i1 * (UP_CT + 1)];
          v[u][k] += (1.f - ty.lam) * top + ty.lam * bot;
        }
        Vec4<FT>::store(fout + offs[u], v[u]);
      }
    }
  } else {
    for (int o = tid; o < nct * HW; o += UP_THREADS) {
      const int cl = o / HW, r = o % HW;
      const int h = r / g.W, w = r % g.W;
      const Tap ty = tap_of(h, rsh, g.A_h), tx = tap_of(w, rsw, g.A_w);
      const float* r0 = sm + (ty.i0 * g.A_w) * (UP_CT + 1) + cl;
      const float* r1 = sm + (ty.i1 * g.A_w) * (UP_CT + 1) + cl;
      const float top = (1.f - tx.lam) * r0[tx.i0 * (UP_CT + 1)] + tx.lam * r0[tx.i1 * (UP_CT + 1)];
      const float bot = (1.f - tx.lam) * r1[tx.i0 * (UP_CT + 1)] + tx.lam * r1[tx.i1 * (UP_CT + 1)];
      const size_t off = (size_t)cl * HW + r;
      fout[off] = from_f<FT>(to_f<FT>(fin[off]) + (1.f - ty.lam) * top + ty.lam * bot);
    }
  }
}

// ------------------------------------------------------------------------------------ forward NHWC
template <typename FT>
__global__ void __launch_bounds__(256)
upsample_add_fwd_nhwc_kernel(dsf_geom g, const float* __restrict__ y, Ptr3 feat, MPtr3 out) {
  const int cells = g.A_h * g.A_w;
  const int slots = (g.V + 2) * g.S;
  const int T = slots * cells + 2;
  const int c4n = g.C / 4;
  const float rsh = (float)g.A_h / (float)g.H, rsw = (float)g.A_w / (float)g.W;
  const int64_t total = (int64_t)g.B * slots * g.H * g.W * c4n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    int64_t r = i / c4n;
    const int w = (int)(r % g.W); r /= g.W;
    const int h = (int)(r % g.H); r /= g.H;
    const int sl = (int)(r % slots), b = (int)(r / slots);
    const Tap ty = tap_of(h, rsh, g.A_h), tx = tap_of(w, rsw, g.A_w);
    const float* base = y + ((size_t)b * T + sl * cells) * g.C + c;
    float a00[4], a01[4], a10[4], a11[4];
    Vec4<float>::load(base + (size_t)(ty.i0 * g.A_w + tx.i0) * g.C, a00);
    Vec4<float>::load(base + (size_t)(ty.i0 * g.A_w + tx.i1) * g.C, a01);
    Vec4<float>::load(base + (size_t)(ty.i1 * g.A_w + tx.i0) * g.C, a10);
    Vec4<float>::load(base + (size_t)(ty.i1 * g.A_w + tx.i1) * g.C, a11);
    int which, n;
    slot_frame(g, b, sl, which, n);
    const size_t off = (((size_t)n * g.H + h) * g.W + w) * g.C + c;
    float v[4];
    Vec4<FT>::load(reinterpret_cast<const FT*>(feat.p[which]) + off, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float top = (1.f - tx.lam) * a00[k] + tx.lam * a01[k];
      const float bot = (1.f - tx.lam) * a10[k] + tx.lam * a11[k];
      v[k] += (1.f - ty.lam) * top + ty.lam * bot;
    }
    Vec4<FT>::store(reinterpret_cast<FT*>(out.p[which]) + off, v);
  }
}

// ------------------------------------------------------------------------------------ backward NCHW
// Separable adjoint: (1) each warp walks one channel plane with lanes along W and scatters the
// row-weighted values into its private tmp[A_h][W] tile, (2) the W axis is folded per anchor cell,
// (3) the (cell, channel) tile is transposed through shared memory and written channel-contiguous.
template <typename FT, int UP_CT>
__global__ void __launch_bounds__(UP_THREADS)
upsample_add_bwd_nchw_kernel(dsf_geom g, Ptr3 dout, const float* __restrict__ dgps_out, float* __restrict__ dy) {
  extern __shared__ float sm[];
  const int cells = g.A_h * g.A_w;
  const int slots = (g.V + 2) * g.S;
  const int Tm = slots * cells, T = Tm + 2;
  const int F = g.B * slots;
  const int f = blockIdx.x, c0 = blockIdx.y * UP_CT;
  const int nct = min(UP_CT, g.C - c0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (f >= F) {
    const int b = f - F;
    for (int i = tid; i < 2 * nct; i += UP_THREADS) {
      const int j = i / nct, c = c0 + i % nct;
      dy[((size_t)b * T + Tm + j) * g.C + c] = dgps_out ? dgps_out[((size_t)b * 2 + j) * g.C + c] : 0.f;
    }
    return;
  }
  float* outT = sm;                                   // [cells][UP_CT + 1]
  float* tmp = sm + cells * (UP_CT + 1) + warp * (g.A_h * g.W);  // [A_h][W] per warp
  const int tab_off = (cells * (UP_CT + 1) + UP_WARPS * g.A_h * g.W + 3) & ~3;  // 16-byte aligned tap tables behind them
  const int b = f / slots, sl = f % slots;
  int which, n;
  slot_frame(g, b, sl, which, n);
  const int HW = g.H * g.W;
  const FT* src = reinterpret_cast<const FT*>(dout.p[which]) + ((size_t)n * g.C + c0) * HW;
  const float rsh = (float)g.A_h / (float)g.H, rsw = (float)g.A_w / (float)g.W;
  const int sw = g.W / g.A_w;
  if (g.H == g.A_h && g.W == g.A_w && cells % 4 == 0) {
    // stage 4 (no upsampling): plain transpose of the contiguous (channel tile x cells) block, 16-byte loads
    for (int i = tid; i < nct * cells / 4; i += UP_THREADS) {
      float v[4];
      Vec4<FT>::load(src + 4 * i, v);
      const int cl = (4 * i) / cells, cell = (4 * i) % cells;
#pragma unroll
      for (int k = 0; k < 4; ++k) outT[(cell + k) * (UP_CT + 1) + cl] = v[k];
    }
  } else {
  const int W4 = g.W / 4, sh = g.H / g.A_h;
  // Per-row / per-column adjoint taps, computed once per CTA (they are the same for every plane): entry i = {weight into anchor
  // k-1, weight into anchor k, weight into anchor k+1, k} with k = i / scale.  A source row (column) of anchor block k only
  // touches anchors k-1, k, k+1 (ty.i0 is k-1 or k, ty.i1 is k or k+1; both k at a clamped border).  Looking the taps up costs
  // one broadcast 16-byte load per row where recomputing them (float index arithmetic + an integer division) made these
  // kernels instruction-bound (ncu: 3000 warp instructions per 32 x 32 plane, 1.1 TB/s).
  float4* rowtab = reinterpret_cast<float4*>(sm + tab_off);
  float4* coltab = rowtab + g.H;
  const bool tables = g.H % g.A_h == 0 && g.W % g.A_w == 0;
  if (tables) {
    for (int i = tid; i < g.H + g.W; i += UP_THREADS) {
      const bool is_row = i < g.H;
      const int idx = is_row ? i : i - g.H, scale = is_row ? sh : sw, A = is_row ? g.A_h : g.A_w;
      const Tap t = tap_of(idx, is_row ? rsh : rsw, A);
      const int k = idx / scale;
      float4 e = make_float4(0.f, 0.f, 0.f, __int_as_float(k));
      if (t.i0 == k) e.y += 1.f - t.lam; else e.x += 1.f - t.lam;
      if (t.i1 == k) e.y += t.lam; else e.z += t.lam;
      (is_row ? rowtab : coltab)[idx] = e;
    }
    __syncthreads();
  }
  // stage-1-like planes (column scale a multiple of 8, >= 2 rows per lane and anchor block): fold both axes in registers
  const bool direct = tables && g.W % 4 == 0 && (W4 == 8 || W4 == 16 || W4 == 32) && sw % 8 == 0 && sh >= 2 * (32 / W4);
  // narrow planes (W = 8 / 16 / 32, stages 2-3): rows streamed with coalesced 4-byte loads, 32 / W rows per load instruction,
  // folded along H in registers (same three-row argument as the direct path), then along W from shared memory
  const bool narrow = !direct && tables && (g.W == 8 || g.W == 16 || g.W == 32) && sh % (32 / g.W) == 0;
  if (direct) {
    for (int i = tid; i < cells * (UP_CT + 1); i += UP_THREADS) outT[i] = 0.f;
    __syncthreads();
  }
  if (narrow) {
    const int rpl = 32 / g.W, par = lane / g.W, w = lane % g.W;
    float a_prev = 0.f, a_cur = 0.f, a_next = 0.f;
    int blk = -1;
    // The row phases of one load instruction lie in the same anchor block (sh is a multiple of 32 / W), so a block change is
    // warp-uniform: the phases are folded with shuffles and lane w alone updates column w (fixed summation order, no atomics).
    auto flush = [&]() {
      if (blk < 0) return;
      for (int o = g.W; o < 32; o <<= 1) {
        a_prev += __shfl_xor_sync(0xffffffffu, a_prev, o);
        a_cur += __shfl_xor_sync(0xffffffffu, a_cur, o);
        a_next += __shfl_xor_sync(0xffffffffu, a_next, o);
      }
      if (par == 0) {
        if (blk > 0) tmp[(blk - 1) * g.W + w] += a_prev;
        tmp[blk * g.W + w] += a_cur;
        if (blk + 1 < g.A_h) tmp[(blk + 1) * g.W + w] += a_next;
      }
      a_prev = a_cur = a_next = 0.f;
    };
    constexpr int U = 8;  // independent loads in flight per lane and plane
    auto begin_plane = [&]() {
      for (int i = lane; i < g.A_h * g.W; i += 32) tmp[i] = 0.f;
      __syncwarp();
      blk = -1;
    };
    auto fold_rows = [&](int h0, const float (&v)[U]) {  // rows h0, h0 + rpl, ... of this lane's row phase
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int h = h0 + u * rpl;
        if (h >= g.H) break;
        const float4 t = rowtab[h];
        const int k = __float_as_int(t.w);
        if (k != blk) { flush(); blk = k; }
        a_prev += t.x * v[u];
        a_cur += t.y * v[u];
        a_next += t.z * v[u];
      }
    };
    auto end_plane = [&](int cl) {
      flush();
      __syncwarp();
      // fold along W: cell (cy, cx) collects the columns of anchor blocks cx-1, cx, cx+1 with their tabulated weights
      for (int cell = lane; cell < cells; cell += 32) {
        const int cy = cell / g.A_w, cx = cell % g.A_w;
        const float* trow = tmp + cy * g.W;
        const int wc = cx * sw;
        float acc = 0.f;
        for (int j = 0; j < sw; ++j) acc += coltab[wc + j].y * trow[wc + j];
        if (cx > 0)
          for (int j = 0; j < sw; ++j) acc += coltab[wc - sw + j].z * trow[wc - sw + j];
        if (cx + 1 < g.A_w)
          for (int j = 0; j < sw; ++j) acc += coltab[wc + sw + j].x * trow[wc + sw + j];
        outT[cell * (UP_CT + 1) + cl] = acc;
      }
      __syncwarp();
    };
    if (g.H <= U * rpl) {
      // tiny planes (<= 256 pixels, one load round): the rows of FOUR of this warp's planes are requested before the first one
      // is folded — with one 1 KB plane in flight per warp the kernel sat at 0.7 TB/s on DRAM latency
      constexpr int PL = 4;
      for (int cl = warp; cl < nct; cl += UP_WARPS * PL) {
        float v[PL][U];
#pragma unroll
        for (int p = 0; p < PL; ++p) {
          const int c = cl + p * UP_WARPS;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int h = par + u * rpl;
            v[p][u] = (c < nct && h < g.H) ? to_f<FT>(src[(size_t)c * HW + (size_t)h * g.W + w]) : 0.f;
          }
        }
#pragma unroll
        for (int p = 0; p < PL; ++p) {
          const int c = cl + p * UP_WARPS;
          if (c >= nct) break;
          begin_plane();
          fold_rows(par, v[p]);
          end_plane(c);
        }
      }
    } else {
      for (int cl = warp; cl < nct; cl += UP_WARPS) {
        const FT* pl = src + (size_t)cl * HW;
        begin_plane();
        for (int h0 = par; h0 < g.H; h0 += U * rpl) {
          float v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int h = h0 + u * rpl;
            v[u] = h < g.H ? to_f<FT>(pl[(size_t)h * g.W + w]) : 0.f;
          }
          fold_rows(h0, v);
        }
        end_plane(cl);
      }
    }
  } else
  for (int cl = warp; cl < nct; cl += UP_WARPS) {
    const FT* pl = src + (size_t)cl * HW;
    if (!direct) {
      for (int i = lane; i < g.A_h * g.W; i += 32) tmp[i] = 0.f;
      __syncwarp();
    }
    if (direct) {
      // Rows of the plane are streamed with 16-byte loads: lane = (row phase, 16-byte chunk), npar = 32 / W4 rows per load
      // instruction.  With a column scale that is a multiple of 8 the four columns of a chunk share one pair of anchor columns
      // (ix0, ix1), so a row's chunk folds to two values; source rows of anchor block k (k*sh .. k*sh+sh-1) touch anchor rows
      // k-1, k, k+1 only, so a lane keeps 3 x 2 register accumulators and adds them to the (cell, channel) tile in shared memory
      // (atomics: lanes of other row phases / chunks hit the same cells) whenever its next row belongs to another block.
      const int npar = 32 / W4, par = lane / W4, w0 = (lane % W4) * 4;
      const Tap tx0 = tap_of(w0, rsw, g.A_w);
      float lx[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) lx[k] = tap_of(w0 + k, rsw, g.A_w).lam;
      float a_prev[2] = {0.f, 0.f}, a_cur[2] = {0.f, 0.f}, a_next[2] = {0.f, 0.f};
      int blk = -1;
      auto flush = [&]() {
        if (blk < 0) return;
        float* o0 = outT + (size_t)(blk * g.A_w + tx0.i0) * (UP_CT + 1) + cl;
        float* o1 = outT + (size_t)(blk * g.A_w + tx0.i1) * (UP_CT + 1) + cl;
        const int up = g.A_w * (UP_CT + 1);
        if (blk > 0) { atomicAdd(o0 - up, a_prev[0]); atomicAdd(o1 - up, a_prev[1]); }
        atomicAdd(o0, a_cur[0]); atomicAdd(o1, a_cur[1]);
        if (blk + 1 < g.A_h) { atomicAdd(o0 + up, a_next[0]); atomicAdd(o1 + up, a_next[1]); }
        a_prev[0] = a_prev[1] = a_cur[0] = a_cur[1] = a_next[0] = a_next[1] = 0.f;
      };
      constexpr int U = 4;
      for (int h0 = par; h0 < g.H; h0 += U * npar) {
        float v[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int h = h0 + u * npar;
          if (h < g.H) Vec4<FT>::load(pl + (size_t)h * g.W + w0, v[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int h = h0 + u * npar;
          if (h >= g.H) break;
          const float4 t = rowtab[h];
          const int k = __float_as_int(t.w);
          if (k != blk) { flush(); blk = k; }
          float c0 = 0.f, c1 = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) { c0 += (1.f - lx[q]) * v[u][q]; c1 += lx[q] * v[u][q]; }
          a_prev[0] += t.x * c0; a_prev[1] += t.x * c1;
          a_cur[0] += t.y * c0; a_cur[1] += t.y * c1;
          a_next[0] += t.z * c0; a_next[1] += t.z * c1;
        }
      }
      flush();
      continue;   // the (cell, channel) tile is complete for this plane: no separate fold along W
    } else
    for (int w = lane; w < g.W; w += 32) {
#pragma unroll 4
      for (int h = 0; h < g.H; ++h) {
        const float v = to_f<FT>(pl[(size_t)h * g.W + w]);
        const Tap ty = tap_of(h, rsh, g.A_h);
        tmp[ty.i0 * g.W + w] += (1.f - ty.lam) * v;
        tmp[ty.i1 * g.W + w] += ty.lam * v;
      }
    }
    __syncwarp();
    for (int cell = lane; cell < cells; cell += 32) {
      const int cy = cell / g.A_w, cx = cell % g.A_w;
      const int w_lo = max(0, cx * sw - sw), w_hi = min(g.W, cx * sw + 2 * sw);
      float acc = 0.f;
      for (int w = w_lo; w < w_hi; ++w) {
        const Tap tx = tap_of(w, rsw, g.A_w);
        float wt = 0.f;
        if (tx.i0 == cx) wt += 1.f - tx.lam;
        if (tx.i1 == cx) wt += tx.lam;
        acc += wt * tmp[cy * g.W + w];
      }
      outT[cell * (UP_CT + 1) + cl] = acc;
    }
    __syncwarp();
  }
  }
  __syncthreads();
  if (nct == UP_CT && g.C % 4 == 0) {  // 16-byte stores
    for (int o = tid; o < cells * (UP_CT / 4); o += UP_THREADS) {
      const int cell = o / (UP_CT / 4), cl = (o % (UP_CT / 4)) * 4;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = outT[cell * (UP_CT + 1) + cl + k];
      Vec4<float>::store(dy + ((size_t)b * T + sl * cells + cell) * g.C + c0 + cl, v);
    }
    return;
  }
  for (int o = tid; o < cells * nct; o += UP_THREADS) {
    const int cell = o / nct, cl = o % nct;
    dy[((size_t)b * T + sl * cells + cell) * g.C + c0 + cl] = outT[cell * (UP_CT + 1) + cl];
  }
}

// ------------------------------------------------------------------------------------ backward NHWC
template <typename FT>
__global__ void __launch_bounds__(256)
upsample_add_bwd_nhwc_kernel(dsf_geom g, Ptr3 dout, const float* __restrict__ dgps_out, float* __restrict__ dy) {
  const int cells = g.A_h * g.A_w;
  const int slots = (g.V + 2) * g.S;
  const int Tm = slots * cells, T = Tm + 2;
  const int c4n = g.C / 4;
  const float rsh = (float)g.A_h / (float)g.H, rsw = (float)g.A_w / (float)g.W;
  const int sh = g.H / g.A_h, sw = g.W / g.A_w;
  const int64_t total = (int64_t)g.B * T * c4n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int64_t bt = i / c4n;
    const int tok = (int)(bt % T), b = (int)(bt / T);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (tok >= Tm) {
      if (dgps_out) Vec4<float>::load(dgps_out + ((size_t)b * 2 + (tok - Tm)) * g.C + c, acc);
    } else {
      const int sl = tok / cells, cell = tok % cells;
      const int cy = cell / g.A_w, cx = cell % g.A_w;
      int which, n;
      slot_frame(g, b, sl, which, n);
      const FT* src = reinterpret_cast<const FT*>(dout.p[which]) + (size_t)n * g.H * g.W * g.C + c;
      const int h_lo = max(0, cy * sh - sh), h_hi = min(g.H, cy * sh + 2 * sh);
      const int w_lo = max(0, cx * sw - sw), w_hi = min(g.W, cx * sw + 2 * sw);
      for (int h = h_lo; h < h_hi; ++h) {
        const Tap ty = tap_of(h, rsh, g.A_h);
        float wy = 0.f;
        if (ty.i0 == cy) wy += 1.f - ty.lam;
        if (ty.i1 == cy) wy += ty.lam;
        if (wy == 0.f) continue;
        for (int w = w_lo; w < w_hi; ++w) {
          const Tap tx = tap_of(w, rsw, g.A_w);
          float wx = 0.f;
          if (tx.i0 == cx) wx += 1.f - tx.lam;
          if (tx.i1 == cx) wx += tx.lam;
          if (wx == 0.f) continue;
          float v[4];
          Vec4<FT>::load(src + ((size_t)h * g.W + w) * g.C, v);
          const float wt = wy * wx;
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[k] += wt * v[k];
        }
      }
    }
    Vec4<float>::store(dy + ((size_t)b * T + tok) * g.C + c, acc);
  }
}

int check_geom(const dsf_geom* g);

}  // namespace dsf

using namespace dsf;

extern "C" int dsf_upsample_add_fwd(const dsf_geom* g, const float* y, const void* img, const void* lidar,
                                    const void* radar, void* out_img, void* out_lidar, void* out_radar, void* stream) {
  if (int e = check_geom(g)) return e;
  DSF_REQUIRE(y && img && lidar && radar && out_img && out_lidar && out_radar, "upsample_add_fwd: NULL pointer");
  DSF_REQUIRE(aligned16(y) && aligned16(img) && aligned16(lidar) && aligned16(radar) && aligned16(out_img) &&
              aligned16(out_lidar) && aligned16(out_radar), "upsample_add_fwd: 16-byte alignment required");
  cudaStream_t st = (cudaStream_t)stream;
  const int cells = g->A_h * g->A_w, slots = (g->V + 2) * g->S;
  Ptr3 fin{{img, lidar, radar}};
  MPtr3 fout{{out_img, out_lidar, out_radar}};
  if (g->layout == DSF_NCHW) {
    const int ct = nchw_channel_tile(g);
    dim3 grid(g->B * slots, cdiv(g->C, ct));
    const size_t smem = (size_t)cells * (ct + 1) * sizeof(float);
#define DSF_UP_FWD(FT, CT_)                                                                                                              \
  do {                                                                                                                                   \
    if (smem > 48 * 1024) cudaFuncSetAttribute(upsample_add_fwd_nchw_kernel<FT, CT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    upsample_add_fwd_nchw_kernel<FT, CT_><<<grid, UP_THREADS, smem, st>>>(*g, y, fin, fout);                                              \
  } while (0)
    if (g->feat_dtype == DSF_F32) { if (ct == 8) DSF_UP_FWD(float, 8); else DSF_UP_FWD(float, 32); }
    else { if (ct == 8) DSF_UP_FWD(__nv_bfloat16, 8); else DSF_UP_FWD(__nv_bfloat16, 32); }
#undef DSF_UP_FWD
  } else {
    const int64_t total = (int64_t)g->B * slots * g->H * g->W * (g->C / 4);
    const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)num_sms() * 16);
    if (g->feat_dtype == DSF_F32) upsample_add_fwd_nhwc_kernel<float><<<blocks, 256, 0, st>>>(*g, y, fin, fout);
    else upsample_add_fwd_nhwc_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(*g, y, fin, fout);
  }
  return check_launch("upsample_add_fwd");
}

extern "C" int dsf_upsample_add_bwd(const dsf_geom* g, const void* dout_img, const void* dout_lidar,
                                    const void* dout_radar, const float* dgps_out, float* dy, void* stream) {
  if (int e = check_geom(g)) return e;
  DSF_REQUIRE(dout_img && dout_lidar && dout_radar && dy, "upsample_add_bwd: NULL pointer");
  DSF_REQUIRE(aligned16(dout_img) && aligned16(dout_lidar) && aligned16(dout_radar) && aligned16(dgps_out) && aligned16(dy),
              "upsample_add_bwd: 16-byte alignment required");
  cudaStream_t st = (cudaStream_t)stream;
  const int cells = g->A_h * g->A_w, slots = (g->V + 2) * g->S;
  Ptr3 din{{dout_img, dout_lidar, dout_radar}};
  if (g->layout == DSF_NCHW) {
    const int ct = nchw_channel_tile(g);
    dim3 grid(g->B * slots + g->B, cdiv(g->C, ct));
    const size_t smem = ((((size_t)cells * (ct + 1) + (size_t)UP_WARPS * g->A_h * g->W + 3) & ~(size_t)3) + 4 * (size_t)(g->H + g->W)) * sizeof(float);
    DSF_REQUIRE(smem <= 227 * 1024, "upsample_add_bwd: tile does not fit shared memory (%zu B)", smem);
#define DSF_UP_BWD(FT, CT_)                                                                                                              \
  do {                                                                                                                                   \
    if (smem > 48 * 1024) cudaFuncSetAttribute(upsample_add_bwd_nchw_kernel<FT, CT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    upsample_add_bwd_nchw_kernel<FT, CT_><<<grid, UP_THREADS, smem, st>>>(*g, din, dgps_out, dy);                                         \
  } while (0)
    if (g->feat_dtype == DSF_F32) { if (ct == 8) DSF_UP_BWD(float, 8); else DSF_UP_BWD(float, 32); }
    else { if (ct == 8) DSF_UP_BWD(__nv_bfloat16, 8); else DSF_UP_BWD(__nv_bfloat16, 32); }
#undef DSF_UP_BWD
  } else {
    const int64_t total = (int64_t)g->B * (slots * cells + 2) * (g->C / 4);
    const int blocks = (int)std::min<int64_t>(cdiv64(total, 256), (int64_t)num_sms() * 16);
    if (g->feat_dtype == DSF_F32) upsample_add_bwd_nhwc_kernel<float><<<blocks, 256, 0, st>>>(*g, din, dgps_out, dy);
    else upsample_add_bwd_nhwc_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(*g, din, dgps_out, dy);
  }
  return check_launch("upsample_add_bwd");
}

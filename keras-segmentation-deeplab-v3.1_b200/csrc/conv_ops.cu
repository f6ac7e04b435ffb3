// Direct (im2col-free) convolutions of the DeepLabV3+ graph that are not GEMMs:
//   * stem 3x3 stride-2 conv on the 3-channel image with the x/127.5-1 preprocessing fused
//     (deeplabv3p.py:270, :317-321 / :283-284)
//   * depthwise / atrous depthwise 3x3, stride 1|2, any dilation, TF-SAME or explicit padding
//     (deeplabv3p.py:73-74, :186-188, :61-69), forward, backward-data and backward-weight.
// NHWC, 8-channel (16 B) vector accesses, fp32 math, consumer-side BatchNorm+activation prologue so the
// normalised tensor of the previous layer never round-trips through HBM, BatchNorm batch statistics of the
// output accumulated in the same pass (fp32 per thread -> smem -> one fp64 atomic per CTA and channel).
//
// These kernels are HBM/L2-bandwidth bound: per output element 2 B read + 2 B write algorithmic traffic at
// 16 bit for 18 flop.  Grids are sized as a multiple of the SM count and loop (persistent style).
#include <atomic>
#include <type_traits>

#include "common.cuh"

namespace dlb {

extern std::atomic<long long> g_launches;

// ---------------------------------------------------------------------------------------------
// depthwise forward
// ---------------------------------------------------------------------------------------------
struct DwArgs {
  int B, H, W, C, Ho, Wo, stride, dil, pad_t, pad_l;
  const void* x; void* y; const float* w;
  const float* in_scale; const float* in_shift; int in_act;
  const float* out_scale; const float* out_shift; int out_act;
  double* stat_sum; double* stat_sqs;
  int cv;        // C / 8
  int ppb;       // pixels per block
  long long npix;
};

template <typename T>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const DwArgs a) {
  pdl_prologue();
  extern __shared__ float s_stats[];   // [2*C] when stats are requested
  const int tid = threadIdx.x;
  const bool active = tid < a.ppb * a.cv;
  const int p_in_blk = tid / a.cv;
  const int cvi = tid - p_in_blk * a.cv;
  const int c0 = cvi * 8;
  const bool stats = a.stat_sum != nullptr;
  if (stats) {
    for (int i = tid; i < 2 * a.C; i += blockDim.x) s_stats[i] = 0.f;
    __syncthreads();
  }
  float wreg[9][8];
  float isc[8], ish[8], osc[8], osh[8];
  if (active) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) wreg[t][i] = a.w[t * a.C + c0 + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      isc[i] = a.in_scale ? a.in_scale[c0 + i] : 1.f;
      ish[i] = a.in_scale ? a.in_shift[c0 + i] : 0.f;
      osc[i] = a.out_scale ? a.out_scale[c0 + i] : 1.f;
      osh[i] = a.out_scale ? a.out_shift[c0 + i] : 0.f;
    }
  }
  float ssum[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ssqs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const T* x = reinterpret_cast<const T*>(a.x);
  T* y = reinterpret_cast<T*>(a.y);
  if (active) {
    for (long long pix = static_cast<long long>(blockIdx.x) * a.ppb + p_in_blk; pix < a.npix;
         pix += static_cast<long long>(gridDim.x) * a.ppb) {
      const int wo = static_cast<int>(pix % a.Wo);
      const long long t1 = pix / a.Wo;
      const int ho = static_cast<int>(t1 % a.Ho);
      const int b = static_cast<int>(t1 / a.Ho);
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int h = ho * a.stride - a.pad_t + ky * a.dil;
        if (h < 0 || h >= a.H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int w = wo * a.stride - a.pad_l + kx * a.dil;
          if (w < 0 || w >= a.W) continue;
          float v[8];
          Vec8<T>::ld(x + ((static_cast<size_t>(b) * a.H + h) * a.W + w) * a.C + c0, v);
          if (a.in_scale) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = apply_act(fmaf(v[i], isc[i], ish[i]), a.in_act);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], wreg[ky * 3 + kx][i], acc[i]);
        }
      }
      if (a.out_scale) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = apply_act(fmaf(acc[i], osc[i], osh[i]), a.out_act);
      }
      if (stats) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float q = Act<T>::rnd(acc[i]); ssum[i] += q; ssqs[i] += q * q; }
      }
      Vec8<T>::st(y + static_cast<size_t>(pix) * a.C + c0, acc);
    }
  }
  if (stats) {
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { atomicAdd(&s_stats[c0 + i], ssum[i]); atomicAdd(&s_stats[a.C + c0 + i], ssqs[i]); }
    }
    __syncthreads();
    for (int c = tid; c < a.C; c += blockDim.x) {
      atomicAdd(&a.stat_sum[c], static_cast<double>(s_stats[c]));
      atomicAdd(&a.stat_sqs[c], static_cast<double>(s_stats[a.C + c]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise backward-data: da[b,h,w,c] = sum_taps w[ky,kx,c] * dy[b,ho,wo,c]
// ---------------------------------------------------------------------------------------------
struct DwBwdArgs {
  int B, H, W, C, Ho, Wo, stride, dil, pad_t, pad_l;
  const void* x; const void* dy; void* dx; const float* w; float* dw;
  const float* in_scale; const float* in_shift; int in_act;
  int cv, ppb; long long npix;
};

template <typename T>
__global__ void __launch_bounds__(256) dw_bwd_data_kernel(const DwBwdArgs a) {
  pdl_prologue();
  const int tid = threadIdx.x;
  if (tid >= a.ppb * a.cv) return;
  const int p_in_blk = tid / a.cv;
  const int c0 = (tid - p_in_blk * a.cv) * 8;
  float wreg[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) wreg[t][i] = a.w[t * a.C + c0 + i];
  const T* dy = reinterpret_cast<const T*>(a.dy);
  T* dx = reinterpret_cast<T*>(a.dx);
  for (long long pix = static_cast<long long>(blockIdx.x) * a.ppb + p_in_blk; pix < a.npix;
       pix += static_cast<long long>(gridDim.x) * a.ppb) {
    const int w = static_cast<int>(pix % a.W);
    const long long t1 = pix / a.W;
    const int h = static_cast<int>(t1 % a.H);
    const int b = static_cast<int>(t1 / a.H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hn = h + a.pad_t - ky * a.dil;
      if (hn < 0 || (hn % a.stride) != 0) continue;
      const int ho = hn / a.stride;
      if (ho >= a.Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int wn = w + a.pad_l - kx * a.dil;
        if (wn < 0 || (wn % a.stride) != 0) continue;
        const int wo = wn / a.stride;
        if (wo >= a.Wo) continue;
        float v[8];
        Vec8<T>::ld(dy + ((static_cast<size_t>(b) * a.Ho + ho) * a.Wo + wo) * a.C + c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(v[i], wreg[ky * 3 + kx][i], acc[i]);
      }
    }
    Vec8<T>::st(dx + static_cast<size_t>(pix) * a.C + c0, acc);
  }
}

// depthwise backward-weight: dw[ky,kx,c] += sum a[b, ho*s - pad + ky*d, wo*s - pad + kx*d, c] * dy[b,ho,wo,c]
template <typename T>
__global__ void __launch_bounds__(256) dw_bwd_weight_kernel(const DwBwdArgs a) {
  pdl_prologue();
  extern __shared__ float s_dw[];   // [9*C]
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * a.C; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  const bool active = tid < a.ppb * a.cv;
  if (active) {
    const int p_in_blk = tid / a.cv;
    const int c0 = (tid - p_in_blk * a.cv) * 8;
    float isc[8], ish[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      isc[i] = a.in_scale ? a.in_scale[c0 + i] : 1.f;
      ish[i] = a.in_scale ? a.in_shift[c0 + i] : 0.f;
    }
    float acc[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
    const T* x = reinterpret_cast<const T*>(a.x);
    const T* dy = reinterpret_cast<const T*>(a.dy);
    for (long long pix = static_cast<long long>(blockIdx.x) * a.ppb + p_in_blk; pix < a.npix;
         pix += static_cast<long long>(gridDim.x) * a.ppb) {
      const int wo = static_cast<int>(pix % a.Wo);
      const long long t1 = pix / a.Wo;
      const int ho = static_cast<int>(t1 % a.Ho);
      const int b = static_cast<int>(t1 / a.Ho);
      float g[8];
      Vec8<T>::ld(dy + static_cast<size_t>(pix) * a.C + c0, g);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int h = ho * a.stride - a.pad_t + ky * a.dil;
        if (h < 0 || h >= a.H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int w = wo * a.stride - a.pad_l + kx * a.dil;
          if (w < 0 || w >= a.W) continue;
          float v[8];
          Vec8<T>::ld(x + ((static_cast<size_t>(b) * a.H + h) * a.W + w) * a.C + c0, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float av = a.in_scale ? apply_act(fmaf(v[i], isc[i], ish[i]), a.in_act) : v[i];
            acc[ky * 3 + kx][i] = fmaf(av, g[i], acc[ky * 3 + kx][i]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s_dw[t * a.C + c0 + i], acc[t][i]);
  }
  __syncthreads();
  for (int i = tid; i < 9 * a.C; i += blockDim.x) atomicAdd(&a.dw[i], s_dw[i]);
}

// ---------------------------------------------------------------------------------------------
// shared-memory tiled depthwise kernels (the fast path).
//   CTA = TH x TW output pixels x 64 channels.  The input tile (with halo) is staged once in shared memory with the
//   consumer-side prologue (BN + activation of the producing layer) and the zero padding already applied, so every
//   input element is read from HBM/L2 once per tile instead of 9 times, and transformed once instead of 9 times.
//   A thread owns one 8-channel vector (its 72 filter taps live in registers) and 4 output pixels.
//   flip = 1 turns the same kernel into the backward-data pass of a stride-1 SAME conv (correlation with the
//   180-degree rotated filter on dy).
// ---------------------------------------------------------------------------------------------
constexpr int kTH = 8, kTW = 16, kCV = 8;     // tile rows, cols, channel vectors (64 channels)

struct DwTileArgs {
  int B, H, W, C, Ho, Wo, stride, dil, pad_t, pad_l;
  const void* x; void* y; const float* w;
  const float* in_scale; const float* in_shift; int in_act;
  int has_fin; dlb_bn_fin fin;   // dw_fwd_tma_h_kernel: the prologue affine is finalised here from the batch statistics
  const float* out_scale; const float* out_shift; int out_act;
  double* stat_sum; double* stat_sqs;
  const void* dy; float* dw;     // backward-weight mode
  int flip;
  int tiles_x, tiles_y, chunks, num_tiles, ih, iw;
  // TMA kernels only: dilation-phase decomposition (sub = d: a tile is one phase of the d x d sub-sampled grids, read
  // with TMA element strides, so the dilated conv is an undilated one in shared memory), smem dilation, NaN padding
  int sub, sdil, nan_fill;
  int n_buf;       // TMA ring depth of the *_tma_h kernels (2..4 tiles in flight per CTA)
};

template <typename T>
__device__ __forceinline__ void dw_stage_input(const DwTileArgs& a, T* s_in, int b, int oy0, int ox0, int c0, bool cv_ok) {
  // cooperative load of the (ih x iw) x 64-channel input tile, prologue + zero padding applied
  const int tid = threadIdx.x, v = tid & 7;
  const int cc = c0 + v * 8;
  float isc[8], ish[8];
  if (a.in_scale && cv_ok) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { isc[i] = a.in_scale[cc + i]; ish[i] = a.in_shift[cc + i]; }
  }
  const T* x = reinterpret_cast<const T*>(a.x);
  const int gy0 = oy0 * a.stride - a.pad_t, gx0 = ox0 * a.stride - a.pad_l;
  const int npos = a.ih * a.iw;
  // 4 positions per thread in flight (loads first, then transform + st.shared): the staging is latency bound otherwise
  constexpr int U = 4;
  for (int p0 = tid >> 3; p0 < npos; p0 += 32 * U) {
    Raw8<T> raw[U];
    bool inb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + 32 * u;
      const int py = p / a.iw, px = p - py * a.iw;
      const int gy = gy0 + py, gx = gx0 + px;
      inb[u] = p < npos && cv_ok && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
      if (inb[u]) raw_ld<T>(x + ((static_cast<size_t>(b) * a.H + gy) * a.W + gx) * a.C + cc, raw[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + 32 * u;
      if (p >= npos) continue;
      float vals[8];
      if (inb[u]) {
        raw_unpack(raw[u], vals);
        if (a.in_scale) {
#pragma unroll
          for (int i = 0; i < 8; ++i) vals[i] = apply_act(fmaf(vals[i], isc[i], ish[i]), a.in_act);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) vals[i] = 0.f;
      }
      Vec8<T>::st(s_in + (static_cast<size_t>(p) * kCV + v) * 8, vals);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 2) dw_fwd_tiled_kernel(const DwTileArgs a) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t s_raw[];
  T* s_in = reinterpret_cast<T*>(s_raw);
  float* s_stats = reinterpret_cast<float*>(s_raw + static_cast<size_t>(a.ih) * a.iw * kCV * 8 * sizeof(T));   // [2][64]
  const int tid = threadIdx.x, v = tid & 7;
  const bool stats = a.stat_sum != nullptr;
  float wreg[9][8];
  float ssum[8], ssqs[8];
  int cur_chunk = -1;
  T* y = reinterpret_cast<T*>(a.y);
  const int per_chunk = a.B * a.tiles_y * a.tiles_x;
  // contiguous tile range per CTA (chunk-major order): the channel chunk -- filter registers, statistics /
  // gradient accumulators -- changes at most a couple of times per CTA instead of on almost every tile
  const int tile_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * a.num_tiles / gridDim.x);
  const int tile_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * a.num_tiles / gridDim.x);
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int chunk = tile / per_chunk;
    int r = tile - chunk * per_chunk;
    const int b = r / (a.tiles_y * a.tiles_x); r -= b * a.tiles_y * a.tiles_x;
    const int ty = r / a.tiles_x, tx = r - ty * a.tiles_x;
    const int c0 = chunk * 64, cc = c0 + v * 8;
    const bool cv_ok = cc < a.C;
    if (chunk != cur_chunk) {
      if (stats && cur_chunk >= 0) {
        // flush the statistics of the finished channel chunk
        __syncthreads();
        if (tid < 128) s_stats[tid] = 0.f;
        __syncthreads();
        const int pc = cur_chunk * 64 + v * 8;
        if (pc < a.C) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { atomicAdd(&s_stats[v * 8 + i], ssum[i]); atomicAdd(&s_stats[64 + v * 8 + i], ssqs[i]); }
        }
        __syncthreads();
        if (tid < 64 && cur_chunk * 64 + tid < a.C) {
          atomicAdd(&a.stat_sum[cur_chunk * 64 + tid], static_cast<double>(s_stats[tid]));
          atomicAdd(&a.stat_sqs[cur_chunk * 64 + tid], static_cast<double>(s_stats[64 + tid]));
        }
      }
      cur_chunk = chunk;
#pragma unroll
      for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssqs[i] = 0.f; }
      if (cv_ok) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ts = a.flip ? 8 - t : t;
#pragma unroll
          for (int i = 0; i < 8; ++i) wreg[t][i] = a.w[ts * a.C + cc + i];
        }
      }
    }
    const int oy0 = ty * kTH, ox0 = tx * kTW;
    __syncthreads();                       // previous tile's compute is done with s_in
    dw_stage_input<T>(a, s_in, b, oy0, ox0, c0, cv_ok);
    __syncthreads();
    if (cv_ok) {
      float osc[8], osh[8];
      if (a.out_scale) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { osc[i] = a.out_scale[cc + i]; osh[i] = a.out_shift[cc + i]; }
      }
#pragma unroll 1
      for (int j = 0; j < (kTH * kTW) / 32; ++j) {
        const int q = (tid >> 3) + 32 * j;
        const int oy = q / kTW, ox = q - oy * kTW;
        if (oy0 + oy >= a.Ho || ox0 + ox >= a.Wo) continue;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int p = (oy * a.stride + ky * a.dil) * a.iw + ox * a.stride + kx * a.dil;
            float xv[8];
            Vec8<T>::ld(s_in + (static_cast<size_t>(p) * kCV + v) * 8, xv);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(xv[i], wreg[ky * 3 + kx][i], acc[i]);
          }
        if (a.out_scale) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = apply_act(fmaf(acc[i], osc[i], osh[i]), a.out_act);
        }
        if (stats) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float qv = Act<T>::rnd(acc[i]); ssum[i] += qv; ssqs[i] += qv * qv; }
        }
        Vec8<T>::st(y + ((static_cast<size_t>(b) * a.Ho + oy0 + oy) * a.Wo + ox0 + ox) * a.C + cc, acc);
      }
    }
  }
  if (stats && cur_chunk >= 0) {
    __syncthreads();
    if (tid < 128) s_stats[tid] = 0.f;
    __syncthreads();
    const int pc = cur_chunk * 64 + v * 8;
    if (pc < a.C) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { atomicAdd(&s_stats[v * 8 + i], ssum[i]); atomicAdd(&s_stats[64 + v * 8 + i], ssqs[i]); }
    }
    __syncthreads();
    if (tid < 64 && cur_chunk * 64 + tid < a.C) {
      atomicAdd(&a.stat_sum[cur_chunk * 64 + tid], static_cast<double>(s_stats[tid]));
      atomicAdd(&a.stat_sqs[cur_chunk * 64 + tid], static_cast<double>(s_stats[64 + tid]));
    }
  }
}

// backward-weight, tiled: dw[tap, c] += sum_pixels a[pix + tap] * dy[pix]; 72 accumulators per thread
template <typename T>
__global__ void __launch_bounds__(256, 1) dw_wgrad_tiled_kernel(const DwTileArgs a) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t s_raw[];
  T* s_in = reinterpret_cast<T*>(s_raw);
  float* s_dw = reinterpret_cast<float*>(s_raw + static_cast<size_t>(a.ih) * a.iw * kCV * 8 * sizeof(T));   // [9][64]
  const int tid = threadIdx.x, v = tid & 7, lane = tid & 31;
  float acc[9][8];
  int cur_chunk = -1;
  const T* dy = reinterpret_cast<const T*>(a.dy);
  const int per_chunk = a.B * a.tiles_y * a.tiles_x;

  auto flush = [&](int chunk) {
    // reduce over the 32 threads that share channel vector v: lanes v, v+8, v+16, v+24 of each of the 8 warps
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = acc[t][i];
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        acc[t][i] = x;
      }
    __syncthreads();
    for (int i = tid; i < 9 * 64; i += 256) s_dw[i] = 0.f;
    __syncthreads();
    if (lane < 8) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(&s_dw[t * 64 + v * 8 + i], acc[t][i]);
    }
    __syncthreads();
    for (int i = tid; i < 9 * 64; i += 256) {
      const int t = i >> 6, c = chunk * 64 + (i & 63);
      if (c < a.C) atomicAdd(&a.dw[t * a.C + c], s_dw[i]);
    }
  };

  // contiguous tile range per CTA (chunk-major order): the channel chunk -- filter registers, statistics /
  // gradient accumulators -- changes at most a couple of times per CTA instead of on almost every tile
  const int tile_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * a.num_tiles / gridDim.x);
  const int tile_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * a.num_tiles / gridDim.x);
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int chunk = tile / per_chunk;
    int r = tile - chunk * per_chunk;
    const int b = r / (a.tiles_y * a.tiles_x); r -= b * a.tiles_y * a.tiles_x;
    const int ty = r / a.tiles_x, tx = r - ty * a.tiles_x;
    const int c0 = chunk * 64, cc = c0 + v * 8;
    const bool cv_ok = cc < a.C;
    if (chunk != cur_chunk) {
      if (cur_chunk >= 0) flush(cur_chunk);
      cur_chunk = chunk;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
    }
    const int oy0 = ty * kTH, ox0 = tx * kTW;
    __syncthreads();
    dw_stage_input<T>(a, s_in, b, oy0, ox0, c0, cv_ok);
    __syncthreads();
    if (cv_ok) {
      Raw8<T> gall[(kTH * kTW) / 32];
#pragma unroll
      for (int j = 0; j < (kTH * kTW) / 32; ++j) {
        const int q = (tid >> 3) + 32 * j;
        const int oy = q / kTW, ox = q - oy * kTW;
        if (oy0 + oy < a.Ho && ox0 + ox < a.Wo)
          raw_ld<T>(dy + ((static_cast<size_t>(b) * a.Ho + oy0 + oy) * a.Wo + ox0 + ox) * a.C + cc, gall[j]);
      }
#pragma unroll
      for (int j = 0; j < (kTH * kTW) / 32; ++j) {
        const int q = (tid >> 3) + 32 * j;
        const int oy = q / kTW, ox = q - oy * kTW;
        if (oy0 + oy >= a.Ho || ox0 + ox >= a.Wo) continue;
        float g[8];
        raw_unpack(gall[j], g);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int p = (oy * a.stride + ky * a.dil) * a.iw + ox * a.stride + kx * a.dil;
            float xv[8];
            Vec8<T>::ld(s_in + (static_cast<size_t>(p) * kCV + v) * 8, xv);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[ky * 3 + kx][i] = fmaf(xv[i], g[i], acc[ky * 3 + kx][i]);
          }
      }
    }
  }
  if (cur_chunk >= 0) flush(cur_chunk);
}

// ---------------------------------------------------------------------------------------------
// fp16 specialisations of the tiled kernels.  The generic versions above are instruction-issue bound (ncu: ~750
// warp instructions per 8-channel output vector, 40 % issue utilisation at 0.5 TB/s).  Here the prologue and the
// 3-tap row sums run on packed half2 (HFMA2 / HMNMX2), rows are added in fp32, filter taps are 36 half2
// registers instead of 72 floats, and tap offsets are precomputed: ~4x fewer instructions per output.
// ---------------------------------------------------------------------------------------------
// 16-bit pair traits: the TMA depthwise kernels are written once for __half and __nv_bfloat16 (packed HFMA2 /
// HFMA2.BF16; the arithmetic intrinsics are overloaded for both pair types)
template <typename T> struct P2;
template <> struct P2<__half> {
  using t = __half2;
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2half2_rn(a, b); }
  static __device__ __forceinline__ t bcast(float a) { return __float2half2_rn(a); }
  static __device__ __forceinline__ float2 unpack(t v) { return __half22float2(v); }
};
template <> struct P2<__nv_bfloat16> {
  using t = __nv_bfloat162;
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static __device__ __forceinline__ t bcast(float a) { return __float2bfloat162_rn(a); }
  static __device__ __forceinline__ float2 unpack(t v) { return __bfloat1622float2(v); }
};
template <typename T> struct __align__(16) V8 { typename P2<T>::t h[4]; };
// 16-byte global accesses go through uint4: nvcc scalarises a copy of the half2[4] struct into four 32-bit LDG/STG
template <typename T> __device__ __forceinline__ V8<T> ldg_h8(const void* p) { uint4 u = *reinterpret_cast<const uint4*>(p); return *reinterpret_cast<V8<T>*>(&u); }
template <typename T> __device__ __forceinline__ void stg_h8(void* p, const V8<T>& o) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&o); }

// ---------------------------------------------------------------------------------------------
// TMA-staged, double-buffered fp16 depthwise kernels.  One cp.async.bulk.tensor (4D box [ih, iw, 64 ch], zero fill
// outside the image = the convolution's padding) brings in the halo tile of the NEXT output tile while the current
// one is being computed; the consumer-side BatchNorm+ReLU6 prologue is an in-place pass over the landed tile.
// (The register-staged version above spends most of a tile's time waiting on its own loads: 2 CTAs/SM, no overlap.)
// ---------------------------------------------------------------------------------------------
int make_tmap_nhwc(CUtensorMap* map, int dtype, const void* ptr, int B, int H, int W, int C, int box_c, int box_w, int box_h,
                   int elem_stride, int nan_fill);

__device__ __forceinline__ void dw_decode_tile(const DwTileArgs& a, int r, int& b, int& oy0, int& ox0) {
  // r = ((b * tiles_y + ty) * tiles_x + tx) * sub^2 + phase : the phases of one region are neighbours in the schedule,
  // so CTAs running at the same time read one dense region between them.  Output pixel (i, j) of the tile is
  // (oy0 + sub * i, ox0 + sub * j).
  const int nph = a.sub * a.sub;
  const int t = r / nph, ph = r - t * nph;
  const int ry = ph / a.sub, rx = ph - ry * a.sub;
  b = t / (a.tiles_y * a.tiles_x);
  const int q = t - b * a.tiles_y * a.tiles_x;
  const int ty = q / a.tiles_x, tx = q - ty * a.tiles_x;
  oy0 = ry + a.sub * ty * kTH; ox0 = rx + a.sub * tx * kTW;
}

// Explicit shared-window accesses: the tile buffers are selected at run time (double buffer), which makes nvcc fall
// back to generic 32-bit LD.E/ST.E (4 instructions and 4-way bank conflicts per 16-byte vector) -- ld/st.shared.v4 on a
// 32-bit shared address keeps every tile access one LDS.128 / STS.128.
template <typename T> __device__ __forceinline__ V8<T> lds_h8(uint32_t saddr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(saddr));
  return *reinterpret_cast<V8<T>*>(&u);
}
template <typename T> __device__ __forceinline__ void sts_h8(uint32_t saddr, const V8<T>& o) {
  const uint4 u = *reinterpret_cast<const uint4*>(&o);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

template <typename T>
__device__ __forceinline__ void dw_transform_tile_h(const DwTileArgs& a, uint32_t s_in, int oy0, int ox0, int c0, bool cv_ok,
                                                    const uint4* s_aff = nullptr) {
  // in place: a = act(x * scale + shift) inside the image, exact zeros in the padding
  const int tid = threadIdx.x, v = tid & 7;
  const int cc = c0 + v * 8;
  if (!cv_ok) return;
  typename P2<T>::t sc2[4], sh2[4];
  if (s_aff) {       // packed (scale, shift) pairs of the CTA's 64 channels, staged once per CTA
    const uint4 us = s_aff[v], uh = s_aff[8 + v];
    *reinterpret_cast<uint4*>(sc2) = us;
    *reinterpret_cast<uint4*>(sh2) = uh;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sc2[i] = P2<T>::pack(a.in_scale[cc + 2 * i], a.in_scale[cc + 2 * i + 1]);
      sh2[i] = P2<T>::pack(a.in_shift[cc + 2 * i], a.in_shift[cc + 2 * i + 1]);
    }
  }
  const typename P2<T>::t zero2 = P2<T>::bcast(0.f), six2 = P2<T>::bcast(6.f);
  const int npos = a.ih * a.iw;
  int p = tid >> 3;
  uint32_t addr = s_in + static_cast<uint32_t>(p * kCV + v) * 16u;
  if (a.nan_fill) {
    // The tensor map fills out-of-image elements with NaN: fma keeps the NaN and max(NaN, 0) = 0 (PTX max returns the
    // non-NaN operand), so ReLU / ReLU6 produce the exact zero padding without any coordinate test.
    if (a.in_act == DLB_ACT_RELU6) {
      for (; p < npos; p += 32, addr += 32u * kCV * 16u) {
        V8<T> o = lds_h8<T>(addr);
#pragma unroll
        for (int i = 0; i < 4; ++i) o.h[i] = __hmin2(__hmax2(__hfma2(o.h[i], sc2[i], sh2[i]), zero2), six2);
        sts_h8(addr, o);
      }
    } else {
      for (; p < npos; p += 32, addr += 32u * kCV * 16u) {
        V8<T> o = lds_h8<T>(addr);
#pragma unroll
        for (int i = 0; i < 4; ++i) o.h[i] = __hmax2(__hfma2(o.h[i], sc2[i], sh2[i]), zero2);
        sts_h8(addr, o);
      }
    }
    return;
  }
  const int gy0 = oy0 * a.stride - a.pad_t, gx0 = ox0 * a.stride - a.pad_l;
  // (py, px) advance by 32 positions per step without a division
  const int step_y = 32 / a.iw, step_x = 32 - step_y * a.iw;
  int py = p / a.iw, px = p - py * a.iw;
  for (; p < npos; p += 32, addr += 32u * kCV * 16u) {
    const int gy = gy0 + py * a.sub, gx = gx0 + px * a.sub;
    if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
      V8<T> o = lds_h8<T>(addr);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        typename P2<T>::t z = __hfma2(o.h[i], sc2[i], sh2[i]);
        if (a.in_act == DLB_ACT_RELU6) z = __hmin2(__hmax2(z, zero2), six2);
        else if (a.in_act == DLB_ACT_RELU) z = __hmax2(z, zero2);
        o.h[i] = z;
      }
      sts_h8(addr, o);
    } else {
      V8<T> o;
#pragma unroll
      for (int i = 0; i < 4; ++i) o.h[i] = zero2;
      sts_h8(addr, o);
    }
    py += step_y; px += step_x;
    if (px >= a.iw) { px -= a.iw; ++py; }
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 2) dw_fwd_tma_h_kernel(const __grid_constant__ CUtensorMap tmap, const DwTileArgs a) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  const uint32_t tile_bytes = static_cast<uint32_t>(a.ih) * a.iw * kCV * sizeof(V8<T>);
  const uint32_t buf_stride = (tile_bytes + 127u) & ~127u;
  const uint32_t s_base = smem_u32(s_raw);
  const int nb = a.n_buf;                                                     // ring of nb tiles (up to 4)
  uint64_t* full = reinterpret_cast<uint64_t*>(s_raw + nb * buf_stride);
  float* s_stats = reinterpret_cast<float*>(s_raw + nb * buf_stride + 32);   // [2][64]
  uint4* s_aff = reinterpret_cast<uint4*>(s_raw + nb * buf_stride + 32 + 512);          // packed prologue scale / shift [2][8]
  int* s_info = reinterpret_cast<int*>(s_raw + nb * buf_stride + 32 + 512 + 256);       // per ring slot: b, oy0, ox0
  const int tid = threadIdx.x, v = tid & 7;
  const bool stats = a.stat_sum != nullptr;
  const bool pro = a.in_scale != nullptr;
  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < nb; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  typename P2<T>::t w2[9][4];
  float ssum[8], ssqs[8];
  unsigned long long ssum2[4] = {0ull, 0ull, 0ull, 0ull}, ssqs2[4] = {0ull, 0ull, 0ull, 0ull};   // packed running statistics
  int off[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) off[ky * 3 + kx] = (ky * a.sdil * a.iw + kx * a.sdil) * kCV * 16;
  int cur_chunk = -1;
  T* y = reinterpret_cast<T*>(a.y);

  auto flush_stats = [&](int chunk) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 s = f32x2_unpack(ssum2[i]), q = f32x2_unpack(ssqs2[i]);
      ssum[2 * i] = s.x; ssum[2 * i + 1] = s.y; ssqs[2 * i] = q.x; ssqs[2 * i + 1] = q.y;
    }
    __syncthreads();
    if (tid < 128) s_stats[tid] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {        // the 4 lanes of a warp that share a channel vector are folded first
      ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 8);  ssqs[i] += __shfl_xor_sync(0xffffffffu, ssqs[i], 8);
      ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 16); ssqs[i] += __shfl_xor_sync(0xffffffffu, ssqs[i], 16);
    }
    if ((tid & 31) < 8 && chunk * 64 + v * 8 < a.C) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { atomicAdd(&s_stats[v * 8 + i], ssum[i]); atomicAdd(&s_stats[64 + v * 8 + i], ssqs[i]); }
    }
    __syncthreads();
    if (tid < 64 && chunk * 64 + tid < a.C) {
      atomicAdd(&a.stat_sum[chunk * 64 + tid], static_cast<double>(s_stats[tid]));
      atomicAdd(&a.stat_sqs[chunk * 64 + tid], static_cast<double>(s_stats[64 + tid]));
    }
  };
  // CTA -> (channel chunk, spatial group): neighbouring CTAs work on the different 64-channel chunks of the SAME
  // spatial tiles at the same time, so the 128-byte pieces of one pixel row are requested together (DRAM page
  // locality), while a CTA keeps its chunk -- filter taps, statistics / gradient accumulators -- for its lifetime.
  const int chunk = blockIdx.x % a.chunks, grp = blockIdx.x / a.chunks, ngrp = gridDim.x / a.chunks;
  const int n_sp = a.B * a.tiles_y * a.tiles_x * a.sub * a.sub;
  auto issue = [&](int sp, int slot) {
    // the tile coordinates are decoded ONCE, by the issuing thread (5 integer divisions: 14 % of the kernel's
    // instructions when all 256 threads did it), and travel with the slot; the mbarrier orders them for the readers
    int b, oy0, ox0;
    dw_decode_tile(a, sp, b, oy0, ox0);
    s_info[slot * 4 + 0] = b; s_info[slot * 4 + 1] = oy0; s_info[slot * 4 + 2] = ox0;
    mbar_expect_tx(&full[slot], tile_bytes);
    tma_load_4d(s_raw + slot * buf_stride, &tmap, &full[slot], chunk * 64, ox0 * a.stride - a.pad_l, oy0 * a.stride - a.pad_t, b);
  };

  pdl_wait();     // index math and barrier init above overlap the preceding kernel's tail; no global access before this
  if (grp >= ngrp) return;
  // nb - 1 tiles are requested ahead: one 23 KB tile in flight per CTA left the kernel latency bound (Little: 2 CTAs x
  // 23 KB x 148 SMs / ~2 us = 3.4 TB/s of input at best)
  if (tid == 0) {
    for (int i = 0; i < nb - 1; ++i)
      if (grp + i * ngrp < n_sp) issue(grp + i * ngrp, i);
  }
  int slot = 0; uint32_t par = 0;
  for (int sp = grp; sp < n_sp; sp += ngrp) {
    const int c0 = chunk * 64, cc = c0 + v * 8;
    const bool cv_ok = cc < a.C;
    if (chunk != cur_chunk) {
      if (stats && cur_chunk >= 0) flush_stats(cur_chunk);
      cur_chunk = chunk;
#pragma unroll
      for (int i = 0; i < 4; ++i) { ssum2[i] = 0ull; ssqs2[i] = 0ull; }
      if (cv_ok) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ts = a.flip ? 8 - t : t;
#pragma unroll
          for (int i = 0; i < 4; ++i) w2[t][i] = P2<T>::pack(a.w[ts * a.C + cc + 2 * i], a.w[ts * a.C + cc + 2 * i + 1]);
        }
      }
      if (pro && a.has_fin) {
        // consumer-side BatchNorm finalisation: 32 threads derive the chunk's 64 channels from the fp64 sums; the CTA
        // of spatial group 0 publishes them (the backward pass reads scale / shift / mean / rstd)
        if (tid < 32) {
          typename P2<T>::t* af = reinterpret_cast<typename P2<T>::t*>(s_aff);
          const int ch = c0 + 2 * tid;
          float sc0 = 0.f, sh0 = 0.f, sc1 = 0.f, sh1 = 0.f;
          if (ch + 1 < a.C) {
            bn_fin_channel(a.fin, ch, grp == 0, sc0, sh0);
            bn_fin_channel(a.fin, ch + 1, grp == 0, sc1, sh1);
          }
          af[tid] = P2<T>::pack(sc0, sc1);
          af[32 + tid] = P2<T>::pack(sh0, sh1);
        }
      } else if (pro && tid < 64) {
        typename P2<T>::t* af = reinterpret_cast<typename P2<T>::t*>(s_aff);
        const int ch = c0 + 2 * (tid & 31);
        const float* src = tid < 32 ? a.in_scale : a.in_shift;
        af[tid] = ch + 1 < a.C ? P2<T>::pack(src[ch], src[ch + 1]) : P2<T>::bcast(0.f);
      }
      __syncthreads();
    }
    // prefetch the next tile into the other buffer (free since the __syncthreads that ended the previous iteration)
    if (tid == 0 && sp + (nb - 1) * ngrp < n_sp) issue(sp + (nb - 1) * ngrp, slot == 0 ? nb - 1 : slot - 1);
    mbar_wait(&full[slot], par);
    const int b = s_info[slot * 4 + 0], oy0 = s_info[slot * 4 + 1], ox0 = s_info[slot * 4 + 2];
    const uint32_t s_in = s_base + slot * buf_stride;
    if (pro) {
      dw_transform_tile_h<T>(a, s_in, oy0, ox0, c0, cv_ok, s_aff);
      __syncthreads();
    }
    if (cv_ok && a.stride == 1 && a.sdil == 1) {
      // Column walk: a thread owns 4 consecutive output rows of one column and rolls over the 6 input rows they need,
      // so every input vector is read from shared memory 3 times (once per horizontal neighbour) instead of 9 -- the
      // tile loads, not the math, were at 62 % of the shared-memory bandwidth (ncu l1tex throughput).
      const int slot = tid >> 3;
      const int ox = slot & (kTW - 1), oyb = (slot / kTW) * 4;
      const int gx = ox0 + ox * a.sub;
      if (gx < a.Wo && oy0 + oyb * a.sub < a.Ho) {
        unsigned long long acc2[4][4];
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc2[o][i] = 0ull;
        const uint32_t base = s_in + static_cast<uint32_t>((oyb * a.iw + ox) * kCV + v) * 16u;
        const uint32_t row_b = static_cast<uint32_t>(a.iw) * kCV * 16u;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const V8<T> x0 = lds_h8<T>(base + r * row_b), x1 = lds_h8<T>(base + r * row_b + kCV * 16u),
                      x2 = lds_h8<T>(base + r * row_b + 2u * kCV * 16u);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int ky = r - o;
            if (ky < 0 || ky > 2) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const typename P2<T>::t racc =
                  __hfma2(x2.h[i], w2[ky * 3 + 2][i], __hfma2(x1.h[i], w2[ky * 3 + 1][i], __hmul2(x0.h[i], w2[ky * 3][i])));
              const float2 f = P2<T>::unpack(racc);
              acc2[o][i] = f32x2_add(acc2[o][i], f32x2_pack(f.x, f.y));
            }
          }
        }
        T* yp = y + ((static_cast<size_t>(b) * a.Ho + oy0 + oyb * a.sub) * a.Wo + gx) * a.C + cc;
        const size_t ystep = static_cast<size_t>(a.sub) * a.Wo * a.C;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int gy = oy0 + (oyb + o) * a.sub;
          if (gy >= a.Ho) break;
          if (a.out_scale) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = f32x2_unpack(acc2[o][i]);
              acc2[o][i] = f32x2_pack(apply_act(fmaf(f.x, a.out_scale[cc + 2 * i], a.out_shift[cc + 2 * i]), a.out_act),
                                      apply_act(fmaf(f.y, a.out_scale[cc + 2 * i + 1], a.out_shift[cc + 2 * i + 1]), a.out_act));
            }
          }
          if (stats) {
#pragma unroll
            for (int i = 0; i < 4; ++i) f32x2_acc_v(ssum2[i], ssqs2[i], acc2[o][i]);
          }
          V8<T> ov;
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float2 f = f32x2_unpack(acc2[o][i]); ov.h[i] = P2<T>::pack(f.x, f.y); }
          stg_h8(yp + o * ystep, ov);
        }
      }
    } else if (cv_ok) {
#pragma unroll 2
      for (int j = 0; j < (kTH * kTW) / 32; ++j) {
        const int q = (tid >> 3) + 32 * j;
        const int oy = q / kTW, ox = q - oy * kTW;
        const int gy = oy0 + oy * a.sub, gx = ox0 + ox * a.sub;
        if (gy >= a.Ho || gx >= a.Wo) continue;
        const uint32_t base = s_in + static_cast<uint32_t>((oy * a.stride * a.iw + ox * a.stride) * kCV + v) * 16u;
        unsigned long long acc2[4] = {0ull, 0ull, 0ull, 0ull};     // fp32 pairs: rows are folded with FADD2
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          typename P2<T>::t racc[4];
          {
            const V8<T> xv = lds_h8<T>(base + off[ky * 3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) racc[i] = __hmul2(xv.h[i], w2[ky * 3][i]);
          }
#pragma unroll
          for (int kx = 1; kx < 3; ++kx) {
            const V8<T> xv = lds_h8<T>(base + off[ky * 3 + kx]);
#pragma unroll
            for (int i = 0; i < 4; ++i) racc[i] = __hfma2(xv.h[i], w2[ky * 3 + kx][i], racc[i]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float2 f = P2<T>::unpack(racc[i]); acc2[i] = f32x2_add(acc2[i], f32x2_pack(f.x, f.y)); }
        }
        if (a.out_scale) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = f32x2_unpack(acc2[i]);
            acc2[i] = f32x2_pack(apply_act(fmaf(f.x, a.out_scale[cc + 2 * i], a.out_shift[cc + 2 * i]), a.out_act),
                                 apply_act(fmaf(f.y, a.out_scale[cc + 2 * i + 1], a.out_shift[cc + 2 * i + 1]), a.out_act));
          }
        }
        if (stats) {
#pragma unroll
          for (int i = 0; i < 4; ++i) f32x2_acc_v(ssum2[i], ssqs2[i], acc2[i]);
        }
        V8<T> o;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = f32x2_unpack(acc2[i]); o.h[i] = P2<T>::pack(f.x, f.y); }
        stg_h8(y + ((static_cast<size_t>(b) * a.Ho + gy) * a.Wo + gx) * a.C + cc, o);
      }
    }
    if (pro) fence_proxy_async();     // generic-proxy writes of the transform vs the TMA that will refill this buffer
    __syncthreads();
    if (++slot == nb) { slot = 0; par ^= 1; }
  }
  if (stats && cur_chunk >= 0) flush_stats(cur_chunk);
}

template <typename T>
__global__ void __launch_bounds__(256, 1) dw_wgrad_tma_h_kernel(const __grid_constant__ CUtensorMap tmap, const DwTileArgs a) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  const uint32_t tile_bytes = static_cast<uint32_t>(a.ih) * a.iw * kCV * sizeof(V8<T>);
  const uint32_t buf_stride = (tile_bytes + 127u) & ~127u;
  const uint32_t s_base = smem_u32(s_raw);
  const int nb = a.n_buf;                                                  // ring of nb tiles (up to 8: one CTA per SM)
  uint64_t* full = reinterpret_cast<uint64_t*>(s_raw + nb * buf_stride);
  float* s_dw = reinterpret_cast<float*>(s_raw + nb * buf_stride + 64);   // [9][64]
  uint4* s_aff = reinterpret_cast<uint4*>(s_raw + nb * buf_stride + 64 + 9 * 64 * 4);        // packed prologue scale / shift [2][8]
  int* s_info = reinterpret_cast<int*>(s_raw + nb * buf_stride + 64 + 9 * 64 * 4 + 256);     // per ring slot: b, oy0, ox0
  const int tid = threadIdx.x, v = tid & 7, lane = tid & 31;
  const bool pro = a.in_scale != nullptr;
  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < nb; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  float acc[9][8];
  int off[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) off[ky * 3 + kx] = (ky * a.sdil * a.iw + kx * a.sdil) * kCV * 16;
  int cur_chunk = -1;
  const T* dy = reinterpret_cast<const T*>(a.dy);

  auto flush = [&](int chunk) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = acc[t][i];
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        acc[t][i] = x;
      }
    __syncthreads();
    for (int i = tid; i < 9 * 64; i += 256) s_dw[i] = 0.f;
    __syncthreads();
    if (lane < 8) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(&s_dw[t * 64 + v * 8 + i], acc[t][i]);
    }
    __syncthreads();
    for (int i = tid; i < 9 * 64; i += 256) {
      const int t = i >> 6, c = chunk * 64 + (i & 63);
      if (c < a.C) atomicAdd(&a.dw[t * a.C + c], s_dw[i]);
    }
  };
  // CTA -> (channel chunk, spatial group): neighbouring CTAs work on the different 64-channel chunks of the SAME
  // spatial tiles at the same time, so the 128-byte pieces of one pixel row are requested together (DRAM page
  // locality), while a CTA keeps its chunk -- filter taps, statistics / gradient accumulators -- for its lifetime.
  const int chunk = blockIdx.x % a.chunks, grp = blockIdx.x / a.chunks, ngrp = gridDim.x / a.chunks;
  const int n_sp = a.B * a.tiles_y * a.tiles_x * a.sub * a.sub;
  auto issue = [&](int sp, int slot) {
    int b, oy0, ox0;           // decoded once, by the issuing thread; travels with the ring slot
    dw_decode_tile(a, sp, b, oy0, ox0);
    s_info[slot * 4 + 0] = b; s_info[slot * 4 + 1] = oy0; s_info[slot * 4 + 2] = ox0;
    mbar_expect_tx(&full[slot], tile_bytes);
    tma_load_4d(s_raw + slot * buf_stride, &tmap, &full[slot], chunk * 64, ox0 * a.stride - a.pad_l, oy0 * a.stride - a.pad_t, b);
  };

  pdl_wait();     // index math and barrier init above overlap the preceding kernel's tail; no global access before this
  if (grp >= ngrp) return;
  if (tid == 0) {
    for (int i = 0; i < nb - 1; ++i)
      if (grp + i * ngrp < n_sp) issue(grp + i * ngrp, i);
  }
  __syncthreads();           // slot info of the first tile is read before its mbarrier wait (dy prefetch below)
  int slot = 0; uint32_t par = 0;
  for (int sp = grp; sp < n_sp; sp += ngrp) {
    const int c0 = chunk * 64, cc = c0 + v * 8;
    const bool cv_ok = cc < a.C;
    if (chunk != cur_chunk) {
      if (cur_chunk >= 0) flush(cur_chunk);
      cur_chunk = chunk;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
      if (pro && tid < 64) {
        typename P2<T>::t* af = reinterpret_cast<typename P2<T>::t*>(s_aff);
        const int ch = c0 + 2 * (tid & 31);
        const float* src = tid < 32 ? a.in_scale : a.in_shift;
        af[tid] = ch + 1 < a.C ? P2<T>::pack(src[ch], src[ch + 1]) : P2<T>::bcast(0.f);
      }
      __syncthreads();
    }
    // slot info of THIS tile was written at least one __syncthreads ago (prologue or an earlier iteration's issue)
    const int b = s_info[slot * 4 + 0], oy0 = s_info[slot * 4 + 1], ox0 = s_info[slot * 4 + 2];
    if (tid == 0 && sp + (nb - 1) * ngrp < n_sp) issue(sp + (nb - 1) * ngrp, slot == 0 ? nb - 1 : slot - 1);
    // the thread's four dy vectors are requested before waiting for the input tile
    constexpr int NP = (kTH * kTW) / 32;
    V8<T> g[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int q = (tid >> 3) + 32 * j;
      const int oy = q / kTW, ox = q - oy * kTW;
      const int gy = oy0 + oy * a.sub, gx = ox0 + ox * a.sub;
      if (cv_ok && gy < a.Ho && gx < a.Wo)
        g[j] = ldg_h8<T>(dy + ((static_cast<size_t>(b) * a.Ho + gy) * a.Wo + gx) * a.C + cc);
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i) g[j].h[i] = P2<T>::bcast(0.f);
      }
    }
    mbar_wait(&full[slot], par);
    const uint32_t s_in = s_base + slot * buf_stride;
    if (pro) {
      dw_transform_tile_h<T>(a, s_in, oy0, ox0, c0, cv_ok, s_aff);
      __syncthreads();
    }
    if (cv_ok) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        typename P2<T>::t p2[4];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const int q = (tid >> 3) + 32 * j;
          const int oy = q / kTW, ox = q - oy * kTW;
          const V8<T> xv = lds_h8<T>(s_in + static_cast<uint32_t>((oy * a.stride * a.iw + ox * a.stride) * kCV + v) * 16u + off[t]);
#pragma unroll
          for (int i = 0; i < 4; ++i) p2[i] = j == 0 ? __hmul2(xv.h[i], g[j].h[i]) : __hfma2(xv.h[i], g[j].h[i], p2[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = P2<T>::unpack(p2[i]); acc[t][2 * i] += f.x; acc[t][2 * i + 1] += f.y; }
      }
    }
    if (pro) fence_proxy_async();
    __syncthreads();
    if (++slot == nb) { slot = 0; par ^= 1; }
  }
  if (cur_chunk >= 0) flush(cur_chunk);
}

// ---------------------------------------------------------------------------------------------
// Backward-data of the stride-2 3x3 depthwise conv (fp16): with u = iy + pad_t, v = ix + pad_l the transposed conv splits
// by parity -- the 2x2 block (u, v) in {2m, 2m+1} x {2n, 2n+1} only needs dy[m-1..m, n-1..n]:
//   dx[2m  ,2n  ] = dy[m,n] w00 + dy[m-1,n] w20 + dy[m,n-1] w02 + dy[m-1,n-1] w22
//   dx[2m  ,2n+1] = dy[m,n] w01 + dy[m-1,n] w21
//   dx[2m+1,2n  ] = dy[m,n] w10 + dy[m,n-1] w12
//   dx[2m+1,2n+1] = dy[m,n] w11
// One TMA box (kTH+1) x (kTW+1) x 64 channels of dy (zero fill outside) per tile of kTH x kTW cells, double-buffered.
// ---------------------------------------------------------------------------------------------
struct DwS2Args {
  int B, H, W, C, Ho, Wo, pad_t, pad_l;
  const float* w; void* dx;
  int tiles_x, tiles_y, chunks;
};

template <typename T>
__global__ void __launch_bounds__(256, 2) dw_bwd_data_s2_tma_h_kernel(const __grid_constant__ CUtensorMap tmap, const DwS2Args a) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  constexpr int IH = kTH + 1, IW = kTW + 1;
  constexpr uint32_t tile_bytes = IH * IW * kCV * 16;
  constexpr uint32_t buf_stride = (tile_bytes + 127u) & ~127u;
  const uint32_t s_base = smem_u32(s_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_raw + 2 * buf_stride);
  const int tid = threadIdx.x, v = tid & 7;
  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int chunk = blockIdx.x % a.chunks, grp = blockIdx.x / a.chunks, ngrp = gridDim.x / a.chunks;
  const int n_sp = a.B * a.tiles_y * a.tiles_x;
  const int cc = chunk * 64 + v * 8;
  const bool cv_ok = cc < a.C;
  typename P2<T>::t w2[9][4];
  if (cv_ok) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) w2[t][i] = P2<T>::pack(a.w[t * a.C + cc + 2 * i], a.w[t * a.C + cc + 2 * i + 1]);
  }
  auto decode = [&](int sp, int& b, int& m0, int& n0) {
    b = sp / (a.tiles_y * a.tiles_x);
    const int q = sp - b * a.tiles_y * a.tiles_x;
    const int ty = q / a.tiles_x;
    m0 = ty * kTH; n0 = (q - ty * a.tiles_x) * kTW;
  };
  auto issue = [&](int sp, int slot) {
    int b, m0, n0;
    decode(sp, b, m0, n0);
    mbar_expect_tx(&full[slot], tile_bytes);
    tma_load_4d(s_raw + slot * buf_stride, &tmap, &full[slot], chunk * 64, n0 - 1, m0 - 1, b);
  };
  pdl_wait();     // index math and barrier init above overlap the preceding kernel's tail; no global access before this
  if (grp >= ngrp) return;
  if (tid == 0 && grp < n_sp) issue(grp, 0);
  T* dx = reinterpret_cast<T*>(a.dx);
  for (int sp = grp, it = 0; sp < n_sp; sp += ngrp, ++it) {
    const int slot = it & 1;
    int b, m0, n0;
    decode(sp, b, m0, n0);
    if (tid == 0 && sp + ngrp < n_sp) issue(sp + ngrp, slot ^ 1);
    mbar_wait(&full[slot], (it >> 1) & 1);
    const uint32_t s_in = s_base + slot * buf_stride;
    if (cv_ok) {
#pragma unroll 2
      for (int j = 0; j < (kTH * kTW) / 32; ++j) {
        const int q = (tid >> 3) + 32 * j;
        const int cm = q / kTW, cn = q - cm * kTW;          // cell (m0 + cm, n0 + cn); smem position (cm + 1, cn + 1)
        const uint32_t base = s_in + static_cast<uint32_t>(((cm + 1) * IW + cn + 1) * kCV + v) * 16u;
        const V8<T> d11 = lds_h8<T>(base);                                     // dy[m, n]
        const V8<T> d01 = lds_h8<T>(base - IW * kCV * 16);                     // dy[m-1, n]
        const V8<T> d10 = lds_h8<T>(base - kCV * 16);                          // dy[m, n-1]
        const V8<T> d00 = lds_h8<T>(base - (IW + 1) * kCV * 16);               // dy[m-1, n-1]
        V8<T> o00, o01, o10, o11;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o00.h[i] = __hfma2(d00.h[i], w2[8][i], __hfma2(d10.h[i], w2[2][i], __hfma2(d01.h[i], w2[6][i], __hmul2(d11.h[i], w2[0][i]))));
          o01.h[i] = __hfma2(d01.h[i], w2[7][i], __hmul2(d11.h[i], w2[1][i]));
          o10.h[i] = __hfma2(d10.h[i], w2[5][i], __hmul2(d11.h[i], w2[3][i]));
          o11.h[i] = __hmul2(d11.h[i], w2[4][i]);
        }
        const int iy0 = 2 * (m0 + cm) - a.pad_t, ix0 = 2 * (n0 + cn) - a.pad_l;
        const bool y0 = iy0 >= 0 && iy0 < a.H, y1 = iy0 + 1 >= 0 && iy0 + 1 < a.H;
        const bool x0 = ix0 >= 0 && ix0 < a.W, x1 = ix0 + 1 >= 0 && ix0 + 1 < a.W;
        T* p = dx + ((static_cast<long long>(b) * a.H + iy0) * a.W + ix0) * a.C + cc;
        const long long rs = static_cast<long long>(a.W) * a.C;
        if (y0 && x0) stg_h8(p, o00);
        if (y0 && x1) stg_h8(p + a.C, o01);
        if (y1 && x0) stg_h8(p + rs, o10);
        if (y1 && x1) stg_h8(p + rs + a.C, o11);
      }
    }
    __syncthreads();
  }
}

// Geometry of the TMA kernels: stride-1 dilated layers (rate 2..8) are decomposed into their d x d phases.
static void fill_tma_geometry(DwTileArgs& a) {
  const bool decomp = a.stride == 1 && a.dil > 1 && a.dil <= 8;
  a.sub = decomp ? a.dil : 1;
  a.sdil = decomp ? 1 : a.dil;
  a.ih = (kTH - 1) * a.stride + 2 * a.sdil + 1;
  a.iw = (kTW - 1) * a.stride + 2 * a.sdil + 1;
  const int hs = (a.Ho + a.sub - 1) / a.sub, ws = (a.Wo + a.sub - 1) / a.sub;
  a.tiles_y = (hs + kTH - 1) / kTH;
  a.tiles_x = (ws + kTW - 1) / kTW;
  a.chunks = (a.C + 63) / 64;
  a.nan_fill = (a.in_scale != nullptr && (a.in_act == DLB_ACT_RELU6 || a.in_act == DLB_ACT_RELU)) ? 1 : 0;
}

static void fill_tile_geometry(DwTileArgs& a) {
  a.sub = 1; a.sdil = a.dil; a.nan_fill = 0;
  a.tiles_y = (a.Ho + kTH - 1) / kTH;
  a.tiles_x = (a.Wo + kTW - 1) / kTW;
  a.chunks = (a.C + 63) / 64;
  a.num_tiles = a.chunks * a.B * a.tiles_y * a.tiles_x;
  a.ih = (kTH - 1) * a.stride + 2 * a.dil + 1;
  a.iw = (kTW - 1) * a.stride + 2 * a.dil + 1;
}

// ---------------------------------------------------------------------------------------------
// stem conv: 3x3 stride 2, Cin = 3, TF-SAME ((0,1) padding for even sizes), preprocessing fused
// ---------------------------------------------------------------------------------------------
struct StemArgs {
  int B, H, W, Cout, Ho, Wo, pad_t, pad_l;
  const float* x; void* y; const float* w;
  const float* out_scale; const float* out_shift; int out_act;
  double* stat_sum; double* stat_sqs;
  long long npix;
};

template <typename T>
__global__ void __launch_bounds__(128) stem_fwd_kernel(const StemArgs a) {
  pdl_prologue();
  __shared__ float s_w[27 * 32];
  __shared__ float s_stat[64];
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * 32; i += blockDim.x) s_w[i] = a.w[i];
  if (tid < 64) s_stat[tid] = 0.f;
  __syncthreads();
  const bool stats = a.stat_sum != nullptr;
  float ssum[32], ssqs[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { ssum[i] = 0.f; ssqs[i] = 0.f; }
  T* y = reinterpret_cast<T*>(a.y);
  for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + tid; pix < a.npix;
       pix += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int hw_ = a.Ho * a.Wo;
    const int b = static_cast<int>(pix / hw_);
    const int r_ = static_cast<int>(pix - static_cast<long long>(b) * hw_);
    const int ho = r_ / a.Wo, wo = r_ - ho * a.Wo;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int h = ho * 2 - a.pad_t + ky;
      if (h < 0 || h >= a.H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int w = wo * 2 - a.pad_l + kx;
        if (w < 0 || w >= a.W) continue;
        const float* px = a.x + ((static_cast<size_t>(b) * a.H + h) * a.W + w) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float v = px[ci] / 127.5f - 1.f;
          const float* wr = &s_w[((ky * 3 + kx) * 3 + ci) * 32];
#pragma unroll
          for (int co = 0; co < 32; ++co) acc[co] = fmaf(v, wr[co], acc[co]);
        }
      }
    }
    if (a.out_scale) {
#pragma unroll
      for (int co = 0; co < 32; ++co) acc[co] = apply_act(fmaf(acc[co], a.out_scale[co], a.out_shift[co]), a.out_act);
    }
    if (stats) {
#pragma unroll
      for (int co = 0; co < 32; ++co) { const float q = Act<T>::rnd(acc[co]); ssum[co] += q; ssqs[co] += q * q; }
    }
#pragma unroll
    for (int v8 = 0; v8 < 4; ++v8) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = acc[v8 * 8 + i];
      Vec8<T>::st(y + static_cast<size_t>(pix) * 32 + v8 * 8, o);
    }
  }
  if (stats) {
#pragma unroll
    for (int co = 0; co < 32; ++co) {
      float s = ssum[co], q = ssqs[co];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if ((tid & 31) == 0) { atomicAdd(&s_stat[co], s); atomicAdd(&s_stat[32 + co], q); }
    }
    __syncthreads();
    if (tid < 32) {
      atomicAdd(&a.stat_sum[tid], static_cast<double>(s_stat[tid]));
      atomicAdd(&a.stat_sqs[tid], static_cast<double>(s_stat[32 + tid]));
    }
  }
}

// stem weight gradient dW[27, 32] = sum_pixels xcol[pix, 27] * dy[pix, 32], register blocked:
// 224 threads = 4 pixel groups x (7 tap groups of 4) x (8 output-channel groups of 4), 16 accumulators each;
// a CTA stages 64-pixel slabs of the (preprocessed, padded) 27-tap patch matrix and of dy in shared memory.
template <typename T>
__global__ void __launch_bounds__(224) stem_wgrad_kernel(int B, int H, int W, int Ho, int Wo, int pad_t, int pad_l,
                                                         const float* __restrict__ x, const T* __restrict__ dy,
                                                         float* __restrict__ dw, long long npix) {
  pdl_prologue();
  constexpr int PIX = 64;
  __shared__ __align__(16) float s_x[PIX][28];
  __shared__ __align__(16) float s_dy[PIX][32];
  const int tid = threadIdx.x;
  const int pg = tid / 56, rem = tid - pg * 56;
  const int tg = rem >> 3, cg = rem & 7;       // taps tg*4.., output channels cg*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long p0 = static_cast<long long>(blockIdx.x) * PIX; p0 < npix; p0 += static_cast<long long>(gridDim.x) * PIX) {
    __syncthreads();
    // patch matrix: thread -> (pixel p = i / 28, tap t); the slab's pixel coordinates come from ONE 32-bit decode of
    // its first pixel (64-bit div/mod per element cost a third of the kernel)
    {
      const int hw = Ho * Wo;
      const int b0 = static_cast<int>(p0 / hw);
      const int r0 = static_cast<int>(p0 - static_cast<long long>(b0) * hw);
      for (int i = tid; i < PIX * 28; i += 224) {
        const int p = i / 28, t = i - p * 28;
        float v = 0.f;
        if (p0 + p < npix && t < 27) {
          int r = r0 + p, b = b0;
          if (r >= hw) { r -= hw; ++b; }           // PIX <= Ho*Wo: at most one image boundary inside a slab
          const int ho = r / Wo, wo = r - ho * Wo;
          const int kk = t / 3, ci = t - kk * 3, ky = kk / 3, kx = kk - ky * 3;
          const int h = ho * 2 - pad_t + ky, w = wo * 2 - pad_l + kx;
          if (h >= 0 && h < H && w >= 0 && w < W) v = x[((static_cast<size_t>(b) * H + h) * W + w) * 3 + ci] / 127.5f - 1.f;
        }
        s_x[p][t] = v;
      }
    }
    for (int i = tid; i < PIX * 4; i += 224) {
      const int p = i >> 2, c8 = (i & 3) * 8;
      const long long pix = p0 + p;
      float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (pix < npix) Vec8<T>::ld(dy + static_cast<size_t>(pix) * 32 + c8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) s_dy[p][c8 + k] = v[k];
    }
    __syncthreads();
#pragma unroll 4
    for (int pp = 0; pp < PIX / 4; ++pp) {
      const int p = pg * (PIX / 4) + pp;
      const float4 xv = *reinterpret_cast<const float4*>(&s_x[p][tg * 4]);
      const float4 gv = *reinterpret_cast<const float4*>(&s_dy[p][cg * 4]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ga[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tap = tg * 4 + i;
    if (tap >= 27) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&dw[tap * 32 + cg * 4 + j], acc[i][j]);
  }
}

// ---------------------------------------------------------------------------------------------
// dense 3x3 conv, stride 1, SAME (Xception entry_flow_conv1_2, deeplabv3p.py:287-291 via _conv2d_same :87-103)
// direct conv: 256 threads = 64 pixels x 4 groups of 16 output channels; weights [3,3,Cin,Cout] staged in smem
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv3x3_fwd_kernel(int B, int H, int W, int Cin, int Cout, const T* __restrict__ x,
                                                          const float* __restrict__ w, T* __restrict__ y,
                                                          const float* __restrict__ out_scale,
                                                          const float* __restrict__ out_shift, int out_act) {
  pdl_prologue();
  extern __shared__ float s_w[];   // [9*Cin][Cout]
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * Cin * Cout; i += 256) s_w[i] = w[i];
  __syncthreads();
  const int cog = tid >> 6;                 // output-channel group of 16 within a 64-channel slab
  const int pl = tid & 63;
  const long long npix = static_cast<long long>(B) * H * W;
  for (int co_base = 0; co_base < Cout; co_base += 64) {
    const int co0 = co_base + cog * 16;
    if (co0 >= Cout) continue;
    for (long long pix = static_cast<long long>(blockIdx.x) * 64 + pl; pix < npix; pix += static_cast<long long>(gridDim.x) * 64) {
      const int wx = static_cast<int>(pix % W);
      const long long t1 = pix / W;
      const int hy = static_cast<int>(t1 % H);
      const int b = static_cast<int>(t1 / H);
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = hy - 1 + ky;
        if (yy < 0 || yy >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = wx - 1 + kx;
          if (xx < 0 || xx >= W) continue;
          const T* px = x + ((static_cast<size_t>(b) * H + yy) * W + xx) * Cin;
          const float* wt = s_w + static_cast<size_t>((ky * 3 + kx) * Cin) * Cout + co0;
          for (int c8 = 0; c8 < Cin; c8 += 8) {
            float v[8];
            Vec8<T>::ld(px + c8, v);
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
              const float* wr = wt + static_cast<size_t>(c8 + ci) * Cout;
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] = fmaf(v[ci], wr[i], acc[i]);
            }
          }
        }
      }
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float vv = acc[h8 * 8 + i];
          if (out_scale) vv = fmaf(vv, out_scale[co0 + h8 * 8 + i], out_shift[co0 + h8 * 8 + i]);
          o[i] = apply_act(vv, out_act);
        }
        Vec8<T>::st(y + static_cast<size_t>(pix) * Cout + co0 + h8 * 8, o);
      }
    }
  }
}

// every `step`-th pixel of x in both dimensions (input of a 1x1 stride-2 'valid' conv, deeplabv3p.py:106-116 with k=1)
template <typename T>
__global__ void __launch_bounds__(256) subsample_kernel(int B, int H, int W, int C, int step, int Ho, int Wo,
                                                        const T* __restrict__ x, T* __restrict__ y) {
  pdl_prologue();
  const int cv = C / 8;
  const long long total = static_cast<long long>(B) * Ho * Wo * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int v = static_cast<int>(i % cv);
    long long t = i / cv;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const long long b = t / Ho;
    float vals[8];
    Vec8<T>::ld(x + ((b * H + static_cast<long long>(ho) * step) * W + static_cast<long long>(wo) * step) * C + v * 8, vals);
    Vec8<T>::st(y + i * 8, vals);
  }
}

// legacy TF1 bilinear resize of an NHWC feature map (deeplabv3p.py:418), output channel pitch ldo (concat slices)
template <typename T>
__global__ void __launch_bounds__(256) resize_feat_kernel(int B, int h, int w, int C, int H, int W, int ldo,
                                                          const T* __restrict__ x, T* __restrict__ y) {
  pdl_prologue();
  const int cv = C / 8;
  const long long total = static_cast<long long>(B) * H * W * cv;
  const float sy = static_cast<float>(h) / static_cast<float>(H), sx = static_cast<float>(w) / static_cast<float>(W);
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int v = static_cast<int>(i % cv);
    long long t = i / cv;
    const int X = static_cast<int>(t % W); t /= W;
    const int Y = static_cast<int>(t % H);
    const long long b = t / H;
    const float fy_ = Y * sy, fx_ = X * sx;
    const int y0 = static_cast<int>(floorf(fy_)), x0 = static_cast<int>(floorf(fx_));
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float fy = fy_ - y0, fx = fx_ - x0;
    float tl[8], tr[8], bl[8], br[8], o[8];
    const T* base = x + b * h * w * C + v * 8;
    Vec8<T>::ld(base + (static_cast<size_t>(y0) * w + x0) * C, tl);
    Vec8<T>::ld(base + (static_cast<size_t>(y0) * w + x1) * C, tr);
    Vec8<T>::ld(base + (static_cast<size_t>(y1) * w + x0) * C, bl);
    Vec8<T>::ld(base + (static_cast<size_t>(y1) * w + x1) * C, br);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float top = tl[k] + (tr[k] - tl[k]) * fx;
      const float bot = bl[k] + (br[k] - bl[k]) * fx;
      o[k] = top + (bot - top) * fy;
    }
    Vec8<T>::st(y + ((b * H + Y) * W + X) * ldo + v * 8, o);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused ASPP atrous depthwise stage (deeplabv3p.py:392-399, the three SepConv_BN depthwise halves with rates
// 6/12/18 or 12/24/36 + BN + ReLU): ONE pass over the backbone feature map produces all three branch inputs.
// With rate 36 on a 64x64 map the halo is the map, so a CTA stages the whole HxW plane of a 32-byte channel slice
// (16 fp16 / 8 fp32 channels, <= 200 KB) in shared memory and computes 3 rates x 9 taps from it.
// Algorithmic traffic: read x once + write three outputs (4 x |x|) instead of 3 reads + 3 writes.
// ---------------------------------------------------------------------------------------------
struct AsppArgs {
  int B, H, W, C;
  const void* x;
  const float* w[3]; int rate[3];
  const float* scale[3]; const float* shift[3];
  void* y[3];
};

template <typename T>
__global__ void __launch_bounds__(256, 1) aspp_dw3_kernel(const AsppArgs a) {
  pdl_prologue();
  constexpr int VP = 32 / (8 * sizeof(T));          // 8-channel vectors per pixel in the slice (2 for 16-bit, 1 for fp32)
  constexpr int CC = VP * 8;
  extern __shared__ __align__(16) uint8_t s_raw[];
  T* plane = reinterpret_cast<T*>(s_raw);
  const int tid = threadIdx.x;
  const int npix = a.H * a.W;
  const int slices = a.C / CC;
  const T* x = reinterpret_cast<const T*>(a.x);
  for (int item = blockIdx.x; item < a.B * slices; item += gridDim.x) {
    const int b = item / slices, c0 = (item - b * slices) * CC;
    __syncthreads();
    constexpr int U = 4;
    for (int i0 = tid; i0 < npix * VP; i0 += 256 * U) {
      Raw8<T> raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + 256 * u;
        if (i < npix * VP) raw_ld<T>(x + (static_cast<size_t>(b) * npix + i / VP) * a.C + c0 + (i % VP) * 8, raw[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + 256 * u;
        if (i < npix * VP) {
          float v[8];
          raw_unpack(raw[u], v);
          Vec8<T>::st(plane + static_cast<size_t>(i) * 8, v);
        }
      }
    }
    __syncthreads();
    const int v = tid % VP;
    const int cc = c0 + v * 8;
#pragma unroll 1
    for (int br = 0; br < 3; ++br) {
      float wreg[9][8], sc[8], sh[8];
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) wreg[t][i] = a.w[br][t * a.C + cc + i];
#pragma unroll
      for (int i = 0; i < 8; ++i) { sc[i] = a.scale[br][cc + i]; sh[i] = a.shift[br][cc + i]; }
      const int d = a.rate[br];
      T* y = reinterpret_cast<T*>(a.y[br]);
      for (int p = tid / VP; p < npix; p += 256 / VP) {
        const int py = p / a.W, px = p - py * a.W;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = py + (ky - 1) * d;
          if (yy < 0 || yy >= a.H) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = px + (kx - 1) * d;
            if (xx < 0 || xx >= a.W) continue;
            float xv[8];
            Vec8<T>::ld(plane + (static_cast<size_t>(yy * a.W + xx) * VP + v) * 8, xv);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(xv[i], wreg[ky * 3 + kx][i], acc[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaxf(fmaf(acc[i], sc[i], sh[i]), 0.f);
        Vec8<T>::st(y + (static_cast<size_t>(b) * npix + p) * a.C + cc, acc);
      }
    }
  }
}

static int pick_grid(long long work_blocks, int per_sm) {
  long long cap = static_cast<long long>(num_sms()) * per_sm;
  return static_cast<int>(work_blocks < cap ? (work_blocks > 0 ? work_blocks : 1) : cap);
}

}  // namespace dlb

using namespace dlb;

// the halo tile must fit in shared memory: large ASPP dilations (12/24/36) on small maps take the gather kernels
static bool dw_tile_fits(int stride, int dil, int elem) {
  const size_t ih = (kTH - 1) * stride + 2 * dil + 1, iw = (kTW - 1) * stride + 2 * dil + 1;
  return ih * iw * kCV * 8 * elem + 9 * 64 * sizeof(float) <= 96 * 1024;
}

template <typename T>
static int launch_dw_gather_fwd(const dlb_dw_conv_params* p, cudaStream_t st) {
  DLB_REQUIRE(p->C / 8 <= 256, "dw_conv_fwd: C <= 2048 for the gather kernel");
  DwArgs a{};
  a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.Ho = p->Ho; a.Wo = p->Wo;
  a.stride = p->stride; a.dil = p->dilation; a.pad_t = p->pad_top; a.pad_l = p->pad_left;
  a.x = p->x; a.y = p->y; a.w = p->w;
  a.in_scale = p->in_scale; a.in_shift = p->in_shift; a.in_act = p->in_act;
  a.out_scale = p->out_scale; a.out_shift = p->out_shift; a.out_act = p->out_act;
  a.stat_sum = p->stat_sum; a.stat_sqs = p->stat_sqs;
  a.cv = p->C / 8; a.ppb = 256 / a.cv; a.npix = static_cast<long long>(p->B) * p->Ho * p->Wo;
  const int grid = pick_grid((a.npix + a.ppb - 1) / a.ppb, 16);
  const size_t smem = p->stat_sum ? 2 * p->C * sizeof(float) : 0;
  launch_k(dw_fwd_kernel<T>, grid, 256, smem, st, a);
  g_launches++;
  return check_launch("dw_fwd_kernel");
}

template <typename T>
static int launch_dw_tiled(DwTileArgs& a, cudaStream_t st) {
  fill_tile_geometry(a);
  const size_t smem = static_cast<size_t>(a.ih) * a.iw * kCV * 8 * sizeof(T) + 128 * sizeof(float);
  DLB_REQUIRE(smem <= 200 * 1024, "dw_conv: dilation %d needs a %zu-byte tile (> 200 KB)", a.dil, smem);
  if (smem > 48 * 1024)
    DLB_CUDA(cudaFuncSetAttribute(dw_fwd_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if constexpr (sizeof(T) == 2) {     // fp16 and bf16 take the TMA kernels
    constexpr int kDt = std::is_same<T, __half>::value ? DLB_F16 : DLB_BF16;
    fill_tma_geometry(a);
    const size_t tile_b = (static_cast<size_t>(a.ih) * a.iw * kCV * 16 + 127) & ~size_t(127);
    // ring depth: as many tiles as fit in half an SM's shared memory (2 CTAs / SM), 2..4
    a.n_buf = static_cast<int>((110 * 1024 - 32 - 128 * sizeof(float) - 256 - 64) / tile_b);
    a.n_buf = a.n_buf > 4 ? 4 : (a.n_buf < 2 ? 2 : a.n_buf);
    const size_t smem_t = a.n_buf * tile_b + 32 + 128 * sizeof(float) + 256 + 64;   // ring + barriers + stats + affine + slot info
    CUtensorMap tm;
    int rc = make_tmap_nhwc(&tm, kDt, a.x, a.B, a.H, a.W, a.C, 64, a.iw * a.sub, a.ih * a.sub, a.sub, a.nan_fill);
    if (rc) return rc;
    DLB_CUDA(cudaFuncSetAttribute(dw_fwd_tma_h_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    const int per_sm_h = smem_t > 110 * 1024 ? 1 : 2;
    const int n_sp = a.B * a.tiles_y * a.tiles_x * a.sub * a.sub;
    int ngrp = (num_sms() * per_sm_h) / a.chunks;
    if (ngrp < 1) ngrp = 1;
    if (ngrp > n_sp) ngrp = n_sp;
    launch_k(dw_fwd_tma_h_kernel<T>, ngrp * a.chunks, 256, smem_t, st, tm, a);
    g_launches++;
    return check_launch("dw_fwd_tma_h_kernel");
  } else {
    const int per_sm = smem > 100 * 1024 ? 1 : 2;
    const int cap = num_sms() * per_sm;
    const int grid = a.num_tiles < cap ? a.num_tiles : cap;
    launch_k(dw_fwd_tiled_kernel<T>, grid, 256, smem, st, a);
    g_launches++;
    return check_launch("dw_fwd_tiled_kernel");
  }
}

template <typename T>
static int launch_dw_wgrad_tiled(DwTileArgs& a, cudaStream_t st) {
  fill_tile_geometry(a);
  const size_t smem = static_cast<size_t>(a.ih) * a.iw * kCV * 8 * sizeof(T) + 9 * 64 * sizeof(float);
  DLB_REQUIRE(smem <= 200 * 1024, "dw_conv_bwd: dilation %d needs a %zu-byte tile (> 200 KB)", a.dil, smem);
  if (smem > 48 * 1024)
    DLB_CUDA(cudaFuncSetAttribute(dw_wgrad_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int cap = num_sms();           // 1 CTA/SM (72 accumulators + staging need ~200 registers)
  const int grid = a.num_tiles < cap ? a.num_tiles : cap;
  if constexpr (sizeof(T) == 2) {
    constexpr int kDt = std::is_same<T, __half>::value ? DLB_F16 : DLB_BF16;
    fill_tma_geometry(a);
    const size_t tile_b = (static_cast<size_t>(a.ih) * a.iw * kCV * 16 + 127) & ~size_t(127);
    const size_t fixed = 64 + 9 * 64 * sizeof(float) + 256 + 128;     // barriers, dw staging, affine, slot info
    a.n_buf = static_cast<int>((200 * 1024 - fixed) / tile_b);       // one CTA per SM: up to 8 tiles in flight
    a.n_buf = a.n_buf > 8 ? 8 : (a.n_buf < 2 ? 2 : a.n_buf);
    const size_t smem_t = a.n_buf * tile_b + fixed;
    CUtensorMap tm;
    int rc = make_tmap_nhwc(&tm, kDt, a.x, a.B, a.H, a.W, a.C, 64, a.iw * a.sub, a.ih * a.sub, a.sub, a.nan_fill);
    if (rc) return rc;
    DLB_CUDA(cudaFuncSetAttribute(dw_wgrad_tma_h_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    const int n_sp = a.B * a.tiles_y * a.tiles_x * a.sub * a.sub;
    int ngrp = num_sms() / a.chunks;
    if (ngrp < 1) ngrp = 1;
    if (ngrp > n_sp) ngrp = n_sp;
    launch_k(dw_wgrad_tma_h_kernel<T>, ngrp * a.chunks, 256, smem_t, st, tm, a);
    g_launches++;
    return check_launch("dw_wgrad_tma_h_kernel");
  } else {
    launch_k(dw_wgrad_tiled_kernel<T>, grid, 256, smem, st, a);
    g_launches++;
    return check_launch("dw_wgrad_tiled_kernel");
  }
}

extern "C" int dlb_dw_conv_fwd(const dlb_dw_conv_params* p, void* stream) {
  DLB_REQUIRE(p && p->x && p->y && p->w, "dw_conv_fwd: null pointer");
  DLB_REQUIRE(p->C % 8 == 0, "dw_conv_fwd: C must be a multiple of 8 (C=%d)", p->C);
  DLB_REQUIRE(p->stride == 1 || p->stride == 2, "dw_conv_fwd: stride %d", p->stride);
  DwTileArgs a{};
  a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.Ho = p->Ho; a.Wo = p->Wo;
  a.stride = p->stride; a.dil = p->dilation; a.pad_t = p->pad_top; a.pad_l = p->pad_left;
  a.x = p->x; a.y = p->y; a.w = p->w;
  a.in_scale = p->in_scale; a.in_shift = p->in_shift; a.in_act = p->in_act;
  a.out_scale = p->out_scale; a.out_shift = p->out_shift; a.out_act = p->out_act;
  a.stat_sum = p->stat_sum; a.stat_sqs = p->stat_sqs;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tiled = dw_tile_fits(p->stride, p->dilation, dtype_size(p->dtype));
  dlb_dw_conv_params q = *p;
  if (p->in_fin) {
    DLB_REQUIRE(p->in_scale == nullptr, "dw_conv_fwd: give in_fin or in_scale / in_shift, not both");
    const int rc = check_bn_fin(p->in_fin, "dw_conv_fwd");
    if (rc) return rc;
    a.in_scale = q.in_scale = p->in_fin->scale; a.in_shift = q.in_shift = p->in_fin->shift;
    if (tiled && p->dtype != DLB_F32 && p->C % 2 == 0) {
      a.has_fin = 1; a.fin = *p->in_fin;        // the TMA kernel finalises in its prologue
    } else {
      const int rf = bn_fin_standalone(p->C, p->in_fin, stream);   // gather / fp32 kernels read finished tables
      if (rf) return rf;
    }
    q.in_fin = nullptr;
  }
  if (!tiled) {
    if (p->dtype == DLB_F16) return launch_dw_gather_fwd<__half>(&q, st);
    if (p->dtype == DLB_BF16) return launch_dw_gather_fwd<__nv_bfloat16>(&q, st);
    return launch_dw_gather_fwd<float>(&q, st);
  }
  if (p->dtype == DLB_F16) return launch_dw_tiled<__half>(a, st);
  if (p->dtype == DLB_BF16) return launch_dw_tiled<__nv_bfloat16>(a, st);
  return launch_dw_tiled<float>(a, st);
}

extern "C" int dlb_dw_conv_bwd(const dlb_dw_conv_bwd_params* p, void* stream) {
  DLB_REQUIRE(p && p->dy && p->w, "dw_conv_bwd: null pointer");
  DLB_REQUIRE(p->C % 8 == 0, "dw_conv_bwd: C must be a multiple of 8 (C=%d)", p->C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dx) {
    const bool tiled = dw_tile_fits(p->stride, p->dilation, dtype_size(p->dtype));
    if (p->stride == 1 && tiled) {
      // backward-data of a stride-1 conv = correlation of dy with the rotated filter, padding 2*d - pad
      DwTileArgs a{};
      a.B = p->B; a.H = p->Ho; a.W = p->Wo; a.C = p->C; a.Ho = p->H; a.Wo = p->W;
      a.stride = 1; a.dil = p->dilation; a.pad_t = 2 * p->dilation - p->pad_top; a.pad_l = 2 * p->dilation - p->pad_left;
      a.x = p->dy; a.y = p->dx; a.w = p->w; a.flip = 1;
      int rc;
      if (p->dtype == DLB_F16) rc = launch_dw_tiled<__half>(a, st);
      else if (p->dtype == DLB_BF16) rc = launch_dw_tiled<__nv_bfloat16>(a, st);
      else rc = launch_dw_tiled<float>(a, st);
      if (rc) return rc;
    } else if (p->stride == 2 && p->dilation == 1 && (p->dtype == DLB_F16 || p->dtype == DLB_BF16)) {
      DwS2Args a{};
      a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.Ho = p->Ho; a.Wo = p->Wo;
      a.pad_t = p->pad_top; a.pad_l = p->pad_left; a.w = p->w; a.dx = p->dx;
      // cells m = 0 .. floor((H - 1 + pad_t) / 2)
      const int cells_y = (p->H - 1 + p->pad_top) / 2 + 1, cells_x = (p->W - 1 + p->pad_left) / 2 + 1;
      a.tiles_y = (cells_y + kTH - 1) / kTH; a.tiles_x = (cells_x + kTW - 1) / kTW;
      a.chunks = (p->C + 63) / 64;
      CUtensorMap tm;
      int rc = make_tmap_nhwc(&tm, p->dtype, p->dy, p->B, p->Ho, p->Wo, p->C, 64, kTW + 1, kTH + 1, 1, 0);
      if (rc) return rc;
      const size_t smem_t = 2 * ((static_cast<size_t>(kTH + 1) * (kTW + 1) * kCV * 16 + 127) & ~size_t(127)) + 16;
      DLB_CUDA(cudaFuncSetAttribute(dw_bwd_data_s2_tma_h_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      DLB_CUDA(cudaFuncSetAttribute(dw_bwd_data_s2_tma_h_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      const int n_sp = a.B * a.tiles_y * a.tiles_x;
      int ngrp = (num_sms() * 2) / a.chunks;
      if (ngrp < 1) ngrp = 1;
      if (ngrp > n_sp) ngrp = n_sp;
      if (p->dtype == DLB_F16) launch_k(dw_bwd_data_s2_tma_h_kernel<__half>, ngrp * a.chunks, 256, smem_t, st, tm, a);
      else launch_k(dw_bwd_data_s2_tma_h_kernel<__nv_bfloat16>, ngrp * a.chunks, 256, smem_t, st, tm, a);
      g_launches++;
      rc = check_launch("dw_bwd_data_s2_tma_h_kernel");
      if (rc) return rc;
    } else {
      DLB_REQUIRE(p->C / 8 <= 256, "dw_conv_bwd: C <= 2048 for the strided backward-data kernel");
      DwBwdArgs a{};
      a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.Ho = p->Ho; a.Wo = p->Wo;
      a.stride = p->stride; a.dil = p->dilation; a.pad_t = p->pad_top; a.pad_l = p->pad_left;
      a.x = p->x; a.dy = p->dy; a.dx = p->dx; a.w = p->w; a.dw = p->dw;
      a.cv = p->C / 8; a.ppb = 256 / a.cv;
      a.npix = static_cast<long long>(p->B) * p->H * p->W;
      const int grid = pick_grid((a.npix + a.ppb - 1) / a.ppb, 16);
      if (p->dtype == DLB_F16) launch_k(dw_bwd_data_kernel<__half>, grid, 256, 0, st, a);
      else if (p->dtype == DLB_BF16) launch_k(dw_bwd_data_kernel<__nv_bfloat16>, grid, 256, 0, st, a);
      else launch_k(dw_bwd_data_kernel<float>, grid, 256, 0, st, a);
      g_launches++;
      int rc = check_launch("dw_bwd_data_kernel");
      if (rc) return rc;
    }
  }
  if (p->dw) {
    DLB_REQUIRE(p->x, "dw_conv_bwd: x required for the weight gradient");
    if (!dw_tile_fits(p->stride, p->dilation, dtype_size(p->dtype))) {
      DLB_REQUIRE(p->C / 8 <= 256, "dw_conv_bwd: C <= 2048 for the gather kernel");
      DwBwdArgs g{};
      g.B = p->B; g.H = p->H; g.W = p->W; g.C = p->C; g.Ho = p->Ho; g.Wo = p->Wo;
      g.stride = p->stride; g.dil = p->dilation; g.pad_t = p->pad_top; g.pad_l = p->pad_left;
      g.x = p->x; g.dy = p->dy; g.dx = p->dx; g.w = p->w; g.dw = p->dw;
      g.in_scale = p->in_scale; g.in_shift = p->in_shift; g.in_act = p->in_act;
      g.cv = p->C / 8; g.ppb = 256 / g.cv;
      g.npix = static_cast<long long>(p->B) * p->Ho * p->Wo;
      const int grid = pick_grid((g.npix + g.ppb - 1) / g.ppb, 2);
      const size_t smem = 9 * p->C * sizeof(float);
#define LG(TT)                                                                                                 \
  do {                                                                                                         \
    if (smem > 48 * 1024)                                                                                      \
      DLB_CUDA(cudaFuncSetAttribute(dw_bwd_weight_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_k(dw_bwd_weight_kernel<TT>, grid, 256, smem, st, g);                                                      \
  } while (0)
      if (p->dtype == DLB_F16) LG(__half);
      else if (p->dtype == DLB_BF16) LG(__nv_bfloat16);
      else LG(float);
#undef LG
      g_launches++;
      return check_launch("dw_bwd_weight_kernel");
    }
    DwTileArgs a{};
    a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.Ho = p->Ho; a.Wo = p->Wo;
    a.stride = p->stride; a.dil = p->dilation; a.pad_t = p->pad_top; a.pad_l = p->pad_left;
    a.x = p->x; a.dy = p->dy; a.dw = p->dw;
    a.in_scale = p->in_scale; a.in_shift = p->in_shift; a.in_act = p->in_act;
    if (p->dtype == DLB_F16) return launch_dw_wgrad_tiled<__half>(a, st);
    if (p->dtype == DLB_BF16) return launch_dw_wgrad_tiled<__nv_bfloat16>(a, st);
    return launch_dw_wgrad_tiled<float>(a, st);
  }
  return DLB_OK;
}

namespace dlb {
// stem_mma.cu: warp-MMA stem kernels for 16-bit activations
struct StemMmaArgs {
  int B, H, W, Ho, Wo, pad_t, pad_l;
  const float* x; void* y; const float* w;
  const float* out_scale; const float* out_shift; int out_act;
  double* stat_sum; double* stat_sqs;
  const void* dy; float* dw;
  long long npix;
  int n_tiles;
};
int stem_fwd_mma(int dtype, const StemMmaArgs& a, cudaStream_t st);
int stem_wgrad_mma(int dtype, const StemMmaArgs& a, cudaStream_t st);
}  // namespace dlb

extern "C" int dlb_stem_conv_fwd(const dlb_stem_conv_params* p, void* stream) {
  DLB_REQUIRE(p && p->x && p->y && p->w, "stem_conv_fwd: null pointer");
  DLB_REQUIRE(p->Cout == 32, "stem_conv_fwd: Cout must be 32 (got %d)", p->Cout);
  StemArgs a{};
  a.B = p->B; a.H = p->H; a.W = p->W; a.Cout = p->Cout; a.Ho = p->Ho; a.Wo = p->Wo;
  // TF SAME for k=3, s=2: pad_total = max((Ho-1)*2 + 3 - H, 0); before = total // 2
  const int pt = (p->Ho - 1) * 2 + 3 - p->H, pl = (p->Wo - 1) * 2 + 3 - p->W;
  a.pad_t = (pt > 0 ? pt : 0) / 2; a.pad_l = (pl > 0 ? pl : 0) / 2;
  a.x = p->x; a.y = p->y; a.w = p->w;
  a.out_scale = p->out_scale; a.out_shift = p->out_shift; a.out_act = p->out_act;
  a.stat_sum = p->stat_sum; a.stat_sqs = p->stat_sqs;
  a.npix = static_cast<long long>(p->B) * p->Ho * p->Wo;
  const int grid = pick_grid((a.npix + 127) / 128, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dtype != DLB_F32 && a.npix < (1ll << 31)) {
    StemMmaArgs m{};
    m.B = a.B; m.H = a.H; m.W = a.W; m.Ho = a.Ho; m.Wo = a.Wo; m.pad_t = a.pad_t; m.pad_l = a.pad_l;
    m.x = a.x; m.y = a.y; m.w = a.w; m.out_scale = a.out_scale; m.out_shift = a.out_shift; m.out_act = a.out_act;
    m.stat_sum = a.stat_sum; m.stat_sqs = a.stat_sqs; m.npix = a.npix;
    return stem_fwd_mma(p->dtype, m, st);
  }
  if (p->dtype == DLB_F16) launch_k(stem_fwd_kernel<__half>, grid, 128, 0, st, a);
  else if (p->dtype == DLB_BF16) launch_k(stem_fwd_kernel<__nv_bfloat16>, grid, 128, 0, st, a);
  else launch_k(stem_fwd_kernel<float>, grid, 128, 0, st, a);
  g_launches++;
  return check_launch("stem_fwd_kernel");
}

extern "C" int dlb_stem_conv_wgrad(int B, int H, int W, int Cout, int dtype, const float* x, const void* dy,
                                   float* dw, void* stream) {
  DLB_REQUIRE(x && dy && dw, "stem_conv_wgrad: null pointer");
  DLB_REQUIRE(Cout == 32, "stem_conv_wgrad: Cout must be 32");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int pt = (Ho - 1) * 2 + 3 - H, pl = (Wo - 1) * 2 + 3 - W;
  const int pad_t = (pt > 0 ? pt : 0) / 2, pad_l = (pl > 0 ? pl : 0) / 2;
  const long long npix = static_cast<long long>(B) * Ho * Wo;
  DLB_REQUIRE(Ho * Wo >= 64, "stem_conv_wgrad: output map %dx%d smaller than one 64-pixel slab", Ho, Wo);
  const int grid = pick_grid((npix + 63) / 64, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype != DLB_F32 && npix < (1ll << 31)) {
    StemMmaArgs m{};
    m.B = B; m.H = H; m.W = W; m.Ho = Ho; m.Wo = Wo; m.pad_t = pad_t; m.pad_l = pad_l;
    m.x = x; m.dy = dy; m.dw = dw; m.npix = npix;
    return stem_wgrad_mma(dtype, m, st);
  }
  if (dtype == DLB_F16)
    launch_k(stem_wgrad_kernel<__half>, grid, 224, 0, st, B, H, W, Ho, Wo, pad_t, pad_l, x, (const __half*)dy, dw, npix);
  else if (dtype == DLB_BF16)
    launch_k(stem_wgrad_kernel<__nv_bfloat16>, grid, 224, 0, st, B, H, W, Ho, Wo, pad_t, pad_l, x, (const __nv_bfloat16*)dy, dw, npix);
  else
    launch_k(stem_wgrad_kernel<float>, grid, 224, 0, st, B, H, W, Ho, Wo, pad_t, pad_l, x, (const float*)dy, dw, npix);
  g_launches++;
  return check_launch("stem_wgrad_kernel");
}

extern "C" int dlb_conv3x3_fwd(int B, int H, int W, int Cin, int Cout, int dtype, const void* x, const float* w, void* y,
                               const float* out_scale, const float* out_shift, int out_act, void* stream) {
  DLB_REQUIRE(x && w && y, "conv3x3_fwd: null pointer");
  DLB_REQUIRE(Cin % 8 == 0 && Cout % 16 == 0, "conv3x3_fwd: Cin %% 8 == 0 and Cout %% 16 == 0 required");
  const size_t smem = static_cast<size_t>(9) * Cin * Cout * sizeof(float);
  DLB_REQUIRE(smem <= 200 * 1024, "conv3x3_fwd: weights (%zu B) do not fit in shared memory", smem);
  const long long npix = static_cast<long long>(B) * H * W;
  const int grid = pick_grid((npix + 63) / 64, 2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define L3(TT)                                                                                                   \
  do {                                                                                                           \
    if (smem > 48 * 1024)                                                                                        \
      DLB_CUDA(cudaFuncSetAttribute(conv3x3_fwd_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_k(conv3x3_fwd_kernel<TT>, grid, 256, smem, st, B, H, W, Cin, Cout, (const TT*)x, w, (TT*)y, out_scale, out_shift, out_act); \
  } while (0)
  if (dtype == DLB_F16) L3(__half);
  else if (dtype == DLB_BF16) L3(__nv_bfloat16);
  else L3(float);
#undef L3
  g_launches++;
  return check_launch("conv3x3_fwd_kernel");
}

extern "C" int dlb_subsample(int B, int H, int W, int C, int step, int dtype, const void* x, void* y, void* stream) {
  DLB_REQUIRE(x && y && C % 8 == 0 && step >= 1, "subsample: bad arguments");
  const int Ho = (H + step - 1) / step, Wo = (W + step - 1) / step;
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 8);
  const int grid = pick_grid((total + 255) / 256, 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == DLB_F32) launch_k(subsample_kernel<float>, grid, 256, 0, st, B, H, W, C, step, Ho, Wo, (const float*)x, (float*)y);
  else if (dtype == DLB_F16) launch_k(subsample_kernel<__half>, grid, 256, 0, st, B, H, W, C, step, Ho, Wo, (const __half*)x, (__half*)y);
  else launch_k(subsample_kernel<__nv_bfloat16>, grid, 256, 0, st, B, H, W, C, step, Ho, Wo, (const __nv_bfloat16*)x, (__nv_bfloat16*)y);
  g_launches++;
  return check_launch("subsample_kernel");
}

extern "C" int dlb_resize_bilinear(int B, int h, int w, int C, int H, int W, int ldo, int dtype, const void* x, void* y,
                                   void* stream) {
  DLB_REQUIRE(x && y && C % 8 == 0 && ldo >= C && ldo % 8 == 0, "resize_bilinear: bad arguments");
  const long long total = static_cast<long long>(B) * H * W * (C / 8);
  const int grid = pick_grid((total + 255) / 256, 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == DLB_F32) launch_k(resize_feat_kernel<float>, grid, 256, 0, st, B, h, w, C, H, W, ldo, (const float*)x, (float*)y);
  else if (dtype == DLB_F16) launch_k(resize_feat_kernel<__half>, grid, 256, 0, st, B, h, w, C, H, W, ldo, (const __half*)x, (__half*)y);
  else launch_k(resize_feat_kernel<__nv_bfloat16>, grid, 256, 0, st, B, h, w, C, H, W, ldo, (const __nv_bfloat16*)x, (__nv_bfloat16*)y);
  g_launches++;
  return check_launch("resize_feat_kernel");
}

extern "C" int dlb_aspp_dw3_fwd(int B, int H, int W, int C, int dtype, const void* x, const float* const* w,
                                const int* rates, const float* const* scale, const float* const* shift,
                                void* const* y, void* stream) {
  DLB_REQUIRE(x && w && rates && scale && shift && y, "aspp_dw3_fwd: null pointer");
  const int cc = dtype == DLB_F32 ? 8 : 16;
  DLB_REQUIRE(C % cc == 0, "aspp_dw3_fwd: C must be a multiple of %d", cc);
  const size_t smem = static_cast<size_t>(H) * W * 32;
  if (smem > 200 * 1024) { set_last_error("aspp_dw3_fwd: %dx%d plane does not fit in shared memory", H, W); return DLB_ERR_UNSUPPORTED; }
  AsppArgs a{};
  a.B = B; a.H = H; a.W = W; a.C = C; a.x = x;
  for (int i = 0; i < 3; ++i) { a.w[i] = w[i]; a.rate[i] = rates[i]; a.scale[i] = scale[i]; a.shift[i] = shift[i]; a.y[i] = y[i]; }
  const int items = B * (C / cc);
  const int grid = items < num_sms() ? items : num_sms();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LA(TT)                                                                                              \
  do {                                                                                                      \
    DLB_CUDA(cudaFuncSetAttribute(aspp_dw3_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_k(aspp_dw3_kernel<TT>, grid, 256, smem, st, a);                                                        \
  } while (0)
  if (dtype == DLB_F16) LA(__half);
  else if (dtype == DLB_BF16) LA(__nv_bfloat16);
  else LA(float);
#undef LA
  g_launches++;
  return check_launch("aspp_dw3_kernel");
}

// Stem convolution (3x3, stride 2, Cin = 3 -> 32, TF-SAME, preprocessing x/127.5 - 1 fused; deeplabv3p.py:270,
// :317-321 and the Xception entry_flow_conv1_1 :283-286) on warp-level tensor-core MMAs for 16-bit activations.
//
// The layer is a [pixels, 27] x [27, 32] contraction -- far too thin for a tcgen05 tile pipeline (one 128 x 32 x 32 MMA
// per 7 KB of input) and 864 FMAs per pixel on the CUDA cores is what bounded the SIMT kernels at 5-7 % of the HBM
// roofline.  Here a warp owns 16 consecutive output pixels per step and issues mma.sync.m16n8k16 with fp32
// accumulation; operands are built in registers straight from global memory, no shared-memory staging:
//   * A = (x - 127.5): x is an integer 0..255, so x - 127.5 is EXACT in fp16 and in bf16; zero padding is the value 0
//     (the reference pads after preprocessing), and the 1/127.5 is applied to the fp32 accumulator.  The 9 floats of
//     one kernel row (3 pixels x 3 channels) are contiguous in NHWC, so tap k = ky*9 + j reads row_base(ky) + j.
//   * forward: B = the weights split hi + lo into two 16-bit values (two MMAs) so no weight rounding enters; the
//     N columns are permuted so lane (g, t) ends up with channels 8t..8t+7 of rows g and g+8: one 16-byte store each.
//   * weight gradient: D[k, ch] = sum_pix A[pix, k] * dy[pix, ch], pixels are the MMA K dimension; dy is exact in its
//     storage type, so the result equals the fp32 kernel's up to summation order.  Channels are permuted so a lane
//     reads 4 contiguous channels (8 bytes) per pixel.
// The fp32 (parity mode) path stays on the exact SIMT kernels in conv_ops.cu.
#include "common.cuh"

namespace dlb {

extern std::atomic<long long> g_launches;

struct StemMmaArgs {
  int B, H, W, Ho, Wo, pad_t, pad_l;
  const float* x; void* y; const float* w;          // w: [27, 32] fp32 (HWIO flattened)
  const float* out_scale; const float* out_shift; int out_act;
  double* stat_sum; double* stat_sqs;
  const void* dy; float* dw;
  long long npix;
  int n_tiles;                                      // 16-pixel tiles
};

template <typename T> struct Mma16;
template <> struct Mma16<__half> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
};
template <> struct Mma16<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

// one pixel of the output map: its patch base coordinates
struct StemPix {
  const float* img;    // image base
  int h0, w0;          // top-left input coordinate of the 3x3 window (may be -1)
  bool ok;
};
__device__ __forceinline__ StemPix stem_decode(const StemMmaArgs& a, long long pix) {
  StemPix p;
  p.ok = pix < a.npix;
  const int hw = a.Ho * a.Wo;
  const int pp = p.ok ? static_cast<int>(pix) : 0;          // host guarantees npix < 2^31: 32-bit divisions
  const int b = pp / hw;
  const int r = pp - b * hw;
  const int ho = r / a.Wo, wo = r - ho * a.Wo;
  p.img = a.x + static_cast<size_t>(b) * a.H * a.W * 3;
  p.h0 = ho * 2 - a.pad_t;
  p.w0 = wo * 2 - a.pad_l;
  return p;
}
// (x - 127.5) of tap k = ky*9 + kx*3 + ci, or 0 for padding / k >= 27 / pixels past the end
__device__ __forceinline__ float stem_tap(const StemMmaArgs& a, const StemPix& p, int ky, int j) {
  const int h = p.h0 + ky, w = p.w0 + j / 3;
  if (!p.ok || ky >= 3 || h < 0 || h >= a.H || w < 0 || w >= a.W) return 0.f;
  return __ldg(p.img + (static_cast<long long>(h) * a.W + p.w0) * 3 + j) - 127.5f;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256, 2) stem_fwd_mma_kernel(const StemMmaArgs a) {
  __shared__ float s_stat[64];
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  if (tid < 64) s_stat[tid] = 0.f;
  // the lane's four k values per 16-wide k step s: 16s + 2t + {0, 1, 8, 9}  ->  (ky, j)
  int kky[8], kj[8];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 16 * s + 2 * t + (q & 1) + (q >> 1) * 8;
      kky[s * 4 + q] = k / 9;
      kj[s * 4 + q] = k - (k / 9) * 9;
    }
  pdl_wait();
  // B fragments (weights are parameters, written by the optimizer long before the preceding kernel): n-tile nt,
  // column g  <->  channel 8*(g>>1) + 2*nt + (g&1); hi + lo split
  uint32_t bh[2][4][2], bl[2][4][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int ch = 8 * (g >> 1) + 2 * nt + (g & 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int k0 = 16 * s + 2 * t + half * 8;
        const float w0 = k0 < 27 ? a.w[k0 * 32 + ch] : 0.f;
        const float w1 = k0 + 1 < 27 ? a.w[(k0 + 1) * 32 + ch] : 0.f;
        const float h0 = Mma16<T>::rnd(w0), h1 = Mma16<T>::rnd(w1);
        bh[s][nt][half] = Mma16<T>::pack(h0, h1);
        bl[s][nt][half] = Mma16<T>::pack(w0 - h0, w1 - h1);
      }
    }
  // epilogue constants of the lane's 8 channels 8t..8t+7
  float osc[8], osh[8];
  const bool affine = a.out_scale != nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    osc[i] = (affine ? a.out_scale[8 * t + i] : 1.f) * (1.f / 127.5f);
    osh[i] = affine ? a.out_shift[8 * t + i] : 0.f;
  }
  const bool stats = a.stat_sum != nullptr;
  float ssum[8], ssqs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssqs[i] = 0.f; }
  __syncthreads();
  T* y = reinterpret_cast<T*>(a.y);
  const int warp_g = blockIdx.x * (blockDim.x >> 5) + (tid >> 5), n_warps = gridDim.x * (blockDim.x >> 5);
  // (prefetching the next tile's taps was measured slower here -- 157 vs 117 us -- the register cap serialises the
  // loads; the weight-gradient kernel below does profit from it)
  for (int tile = warp_g; tile < a.n_tiles; tile += n_warps) {
    const long long p0 = static_cast<long long>(tile) * 16;
    const StemPix pa = stem_decode(a, p0 + g), pb = stem_decode(a, p0 + g + 8);
    float va[8], vb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { va[i] = stem_tap(a, pa, kky[i], kj[i]); vb[i] = stem_tap(a, pb, kky[i], kj[i]); }
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      // a0,a1: row g, k 2t,2t+1 ; a2,a3: row g+8 ; a4,a5: row g, k 2t+8,+9 ; a6,a7: row g+8
      const uint32_t af[4] = {Mma16<T>::pack(va[s * 4 + 0], va[s * 4 + 1]), Mma16<T>::pack(vb[s * 4 + 0], vb[s * 4 + 1]),
                              Mma16<T>::pack(va[s * 4 + 2], va[s * 4 + 3]), Mma16<T>::pack(vb[s * 4 + 2], vb[s * 4 + 3])};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        Mma16<T>::mma(acc[nt], af, bh[s][nt][0], bh[s][nt][1]);
        Mma16<T>::mma(acc[nt], af, bl[s][nt][0], bl[s][nt][1]);
      }
    }
    // lane holds channels 8t + 2nt + {0,1} of rows g (acc[nt][0..1]) and g+8 (acc[nt][2..3])
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const bool ok = p0 + g + 8 * r < a.npix;
      float o[8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        o[2 * nt] = fmaf(acc[nt][2 * r], osc[2 * nt], osh[2 * nt]);
        o[2 * nt + 1] = fmaf(acc[nt][2 * r + 1], osc[2 * nt + 1], osh[2 * nt + 1]);
      }
      if (affine) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = apply_act(o[i], a.out_act);
      }
      if (ok) {
        if (stats) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float q = Mma16<T>::rnd(o[i]); ssum[i] += q; ssqs[i] += q * q; }
        }
        Vec8<T>::st(y + static_cast<size_t>(p0 + g + 8 * r) * 32 + 8 * t, o);
      }
    }
  }
  if (stats) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = ssum[i], q = ssqs[i];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if (g == 0) { atomicAdd(&s_stat[8 * t + i], s); atomicAdd(&s_stat[32 + 8 * t + i], q); }
    }
    __syncthreads();
    if (tid < 32) {
      atomicAdd(&a.stat_sum[tid], static_cast<double>(s_stat[tid]));
      atomicAdd(&a.stat_sqs[tid], static_cast<double>(s_stat[32 + tid]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dw[k, ch] += (1/127.5) * sum_pix (x - 127.5)[pix, k] * dy[pix, ch]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256, 2) stem_wgrad_mma_kernel(const StemMmaArgs a) {
  __shared__ float s_dw[27 * 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 27 * 32; i += blockDim.x) s_dw[i] = 0.f;
  // the lane's A rows: k = g, g+8, g+16, g+24  ->  (ky, j)
  int kky[4], kj[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int k = g + 8 * q;
    kky[q] = k / 9;
    kj[q] = k - (k / 9) * 9;
  }
  float acc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  pdl_wait();
  __syncthreads();
  const T* dy = reinterpret_cast<const T*>(a.dy);
  const int warp_g = blockIdx.x * (blockDim.x >> 5) + (tid >> 5), n_warps = gridDim.x * (blockDim.x >> 5);
  auto load_tile = [&](int tile, float (&v)[4][4], uint2 (&d)[4]) {
    const long long p0 = static_cast<long long>(tile) * 16;
    // the lane's 4 pixels (MMA K indices 2t, 2t+1, 2t+8, 2t+9)
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      const long long pix = p0 + 2 * t + (ps & 1) + (ps >> 1) * 8;
      const StemPix p = stem_decode(a, pix);
#pragma unroll
      for (int q = 0; q < 4; ++q) v[ps][q] = stem_tap(a, p, kky[q], kj[q]);
      d[ps] = p.ok ? __ldg(reinterpret_cast<const uint2*>(dy + static_cast<size_t>(pix) * 32 + 4 * g)) : make_uint2(0u, 0u);
    }
  };
  float v[4][4];           // [pixel slot][k slot]
  uint2 d[4];              // dy channels 4g..4g+3 of each pixel
  if (warp_g < a.n_tiles) load_tile(warp_g, v, d);
  for (int tile = warp_g; tile < a.n_tiles; tile += n_warps) {
    float nv[4][4];
    uint2 nd[4];
    const bool more = tile + n_warps < a.n_tiles;
    if (more) load_tile(tile + n_warps, nv, nd);        // next tile in flight during this tile's MMAs
    // B fragments: n-tile nt, column g <-> channel 4g + nt ; b0,b1 = pixels 2t,2t+1 ; b2,b3 = pixels 2t+8,2t+9
    uint32_t bf[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const uint32_t sel = (nt & 1) ? 0x7632u : 0x5410u;        // high / low halves of the two words
      const uint32_t w00 = nt < 2 ? d[0].x : d[0].y, w01 = nt < 2 ? d[1].x : d[1].y;
      const uint32_t w10 = nt < 2 ? d[2].x : d[2].y, w11 = nt < 2 ? d[3].x : d[3].y;
      bf[nt][0] = __byte_perm(w00, w01, sel);
      bf[nt][1] = __byte_perm(w10, w11, sel);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      // a0,a1: row g (k = 16mt+g), pixels 2t,2t+1 ; a2,a3: row g+8 ; a4,a5: row g, pixels 2t+8,+9 ; a6,a7: row g+8
      const uint32_t af[4] = {Mma16<T>::pack(v[0][2 * mt], v[1][2 * mt]), Mma16<T>::pack(v[0][2 * mt + 1], v[1][2 * mt + 1]),
                              Mma16<T>::pack(v[2][2 * mt], v[3][2 * mt]), Mma16<T>::pack(v[2][2 * mt + 1], v[3][2 * mt + 1])};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) Mma16<T>::mma(acc[mt][nt], af, bf[nt][0], bf[nt][1]);
    }
    if (more) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        d[ps] = nd[ps];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[ps][q] = nv[ps][q];
      }
    }
  }
  // acc[mt][nt]: c0,c1 = row k = 16mt+g, columns 2t,2t+1 of n-tile nt (channels 4*(2t)+nt, 4*(2t+1)+nt); c2,c3 = row +8
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = 16 * mt + g + 8 * (i >> 1);
        const int ch = 4 * (2 * t + (i & 1)) + nt;
        if (k < 27) atomicAdd(&s_dw[k * 32 + ch], acc[mt][nt][i]);
      }
  __syncthreads();
  for (int i = tid; i < 27 * 32; i += blockDim.x) atomicAdd(&a.dw[i], s_dw[i] * (1.f / 127.5f));
}

static int stem_grid(int n_tiles) {
  // 8 warps per CTA, 2 CTAs per SM, at least one tile per warp
  const long long want = (static_cast<long long>(n_tiles) + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 2;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

int stem_fwd_mma(int dtype, const StemMmaArgs& a0, cudaStream_t st) {
  StemMmaArgs a = a0;
  a.n_tiles = static_cast<int>((a.npix + 15) / 16);
  const int grid = stem_grid(a.n_tiles);
  if (dtype == DLB_F16) launch_k(stem_fwd_mma_kernel<__half>, grid, 256, 0, st, a);
  else launch_k(stem_fwd_mma_kernel<__nv_bfloat16>, grid, 256, 0, st, a);
  g_launches++;
  return check_launch("stem_fwd_mma_kernel");
}

int stem_wgrad_mma(int dtype, const StemMmaArgs& a0, cudaStream_t st) {
  StemMmaArgs a = a0;
  a.n_tiles = static_cast<int>((a.npix + 15) / 16);
  const int grid = stem_grid(a.n_tiles);
  if (dtype == DLB_F16) launch_k(stem_wgrad_mma_kernel<__half>, grid, 256, 0, st, a);
  else launch_k(stem_wgrad_mma_kernel<__nv_bfloat16>, grid, 256, 0, st, a);
  g_launches++;
  return check_launch("stem_wgrad_mma_kernel");
}

}  // namespace dlb

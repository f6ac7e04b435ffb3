// Fused ASPP / decoder separable-conv branches: atrous depthwise 3x3 + BN + ReLU + pointwise 1x1 + BN + ReLU in ONE
// kernel, the depthwise result never leaves the SM.
//
// Replaces, per launch, the parallel spatial ASPP branches `aspp0` (plain 1x1) and `aspp1..3` = SepConv_BN(x, 256,
// rate = 12/24/36 | 6/12/18, depth_activation=True) of deeplabv3p.py:385-399 (SepConv_BN :47-84), and with one branch
// the decoder's `decoder_conv0/1` SepConv_BN (:426-429).  Unfused, a branch writes the [M, C] depthwise result to HBM
// and the pointwise GEMM reads it back (x: 1 read, dw: 3 writes + 3 reads, aspp0: 1 more read of x = 8 passes over
// a 64x64x2048 map); here x is the only activation read and the [M, 256] branch outputs the only writes.
//
//   work item = (image, tile of TH full rows with TH*W <= 128 pixels, branch); persistent CTAs stride over items.
//   K loop over 64-channel chunks:
//     warp 0      TMA producer: up to three 4D boxes [TH rows, W, 64 ch] (rows y-d, y, y+d; out-of-image rows are
//                 zero-filled by the tensor map = the conv's padding; fully outside groups are not loaded at all),
//                 128-byte swizzle; one bulk copy of the chunk's packed depthwise taps + folded BN; the [N, 64]
//                 pointwise weight tile (2D box, 128-byte swizzle)
//     warps 6-13  depthwise producers: thread = (pixel, 8-channel vector); 9 taps on packed HFMA2 from conflict-free
//                 LDS.128 (8 lanes cover one pixel's 128 B), rows folded in fp32, BN affine + ReLU, 16-bit result
//                 written straight into the K-major SWIZZLE_128B layout tcgen05 expects for the A operand
//                 (fence.proxy.async -> mbarrier).  rate 0 (aspp0): the centre box IS the A tile, copied through.
//     warp 1      one thread issues tcgen05.mma (M=128, N<=256, K=16 x4 per chunk), fp32 accumulators in TMEM,
//                 double-buffered (2 x 256 columns) so the epilogue of one item overlaps the next item's K loop
//     warps 2-5   epilogue: tcgen05.ld -> pointwise BN affine + ReLU -> swizzled staging tile -> 128-byte row stores
//                 into the branch's channel slice of the concat buffer (ldc)
//
// Bound: L2->SM traffic (each input row is needed by three output rows that are d >= TH apart, so it is re-read
// from L2, not from HBM) and the depthwise issue rate; HBM sees x once.  See DESIGN.md section 3.
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace dlb {

extern std::atomic<long long> g_launches;
int make_tmap_2d(CUtensorMap* map, int dtype, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols);
int make_tmap_nhwc_sw128(CUtensorMap* map, int dtype, const void* ptr, int B, int H, int W, int C, int box_w, int box_h);

constexpr int kFMaxBr = 4;
constexpr int kFThreads = 448;          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, warps 6-13 depthwise
constexpr int kFDwWarps = 8;
constexpr int kFDwThread0 = 192;
constexpr int kFGroupBytes = 16384;     // 128 pixels x 64 channels x 2 B
constexpr int kFPackBytes = 1664;       // 9 taps x 64 ch x 2 B + 64 scale + 64 shift (fp32)
constexpr int kFInStage = 3 * kFGroupBytes + 2048;
constexpr int kFStages = 2;
constexpr int kFABytes = 16384;
constexpr int kFBBytes = 32768;         // N <= 256 rows x 128 B
constexpr int kFStageTile = 4096;       // per epilogue warp: 32 rows x 128 B
constexpr int kFOffA = kFStages * kFInStage;
constexpr int kFOffB = kFOffA + kFStages * kFABytes;
constexpr int kFOffStg = kFOffB + kFStages * kFBBytes;
constexpr int kFOffPw = kFOffStg + 4 * kFStageTile;
constexpr int kFOffBar = kFOffPw + kFMaxBr * 512 * 4;
constexpr int kFSmem = 1024 + kFOffBar + 256;

struct WMaps { CUtensorMap m[kFMaxBr]; };

struct FusedArgs {
  int B, H, W, C, N;
  int TH, tiles_y, n_items, nk, n_br;
  int rate[kFMaxBr];
  int pack_idx[kFMaxBr];
  const uint8_t* pack;
  const float* pw_scale[kFMaxBr];
  const float* pw_shift[kFMaxBr];
  void* out[kFMaxBr];
  int ldc, dw_act, pw_act;
  uint32_t idesc;
};

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(saddr));
  return u;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

template <typename T> struct H2;
template <> struct H2<__half> {
  using t = __half2;
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2half2_rn(a, b); }
  static __device__ __forceinline__ float2 unpack(t v) { return __half22float2(v); }
  static __device__ __forceinline__ t zero() { return __float2half2_rn(0.f); }
};
template <> struct H2<__nv_bfloat16> {
  using t = __nv_bfloat162;
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static __device__ __forceinline__ float2 unpack(t v) { return __bfloat1622float2(v); }
  static __device__ __forceinline__ t zero() { return __float2bfloat162_rn(0.f); }
};

struct ItemGeom {
  int br, b, y0, d;
  bool g0, g2;      // row groups y0-d.. / y0+d.. intersect the image
};
__device__ __forceinline__ ItemGeom decode_item(const FusedArgs& a, int item) {
  ItemGeom g;
  g.br = item % a.n_br;
  const int t = item / a.n_br;
  const int ty = t % a.tiles_y;
  g.b = t / a.tiles_y;
  g.y0 = ty * a.TH;
  g.d = a.rate[g.br];
  g.g0 = g.d > 0 && g.y0 - g.d + a.TH - 1 >= 0;
  g.g2 = g.d > 0 && g.y0 + g.d < a.H;
  return g;
}

template <typename T>
__global__ void __launch_bounds__(kFThreads, 1)
sepconv_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ WMaps wmaps, const FusedArgs a) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t s0 = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFOffBar);
  uint64_t* in_full = bars;            // [2]
  uint64_t* in_empty = bars + 2;       // [2]
  uint64_t* a_full = bars + 4;         // [2]
  uint64_t* a_empty = bars + 6;        // [2]
  uint64_t* b_full = bars + 8;         // [2]
  uint64_t* b_empty = bars + 10;       // [2]
  uint64_t* t_full = bars + 12;        // [2]
  uint64_t* t_empty = bars + 14;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* s_pw = reinterpret_cast<float*>(smem + kFOffPw);     // [branch][scale 256 | shift 256]

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < a.n_br; ++i) tma_prefetch_desc(&wmaps.m[i]);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_full[i], 1);  mbar_init(&in_empty[i], kFDwWarps);
      mbar_init(&a_full[i], kFDwWarps); mbar_init(&a_empty[i], 1);
      mbar_init(&b_full[i], 1);   mbar_init(&b_empty[i], 1);
      mbar_init(&t_full[i], 1);   mbar_init(&t_empty[i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < a.n_br * 512; i += kFThreads) {
    const int br = i >> 9, j = i & 511;
    float v;
    if (j < 256) v = (j < a.N && a.pw_scale[br]) ? a.pw_scale[br][j] : 1.f;
    else v = (j - 256 < a.N && a.pw_shift[br]) ? a.pw_shift[br][j - 256] : 0.f;
    s_pw[i] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t box_bytes = static_cast<uint32_t>(a.TH) * a.W * 128u;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int st = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const ItemGeom g = decode_item(a, item);
        const uint32_t in_bytes = box_bytes * (1u + (g.g0 ? 1u : 0u) + (g.g2 ? 1u : 0u)) + (g.d > 0 ? kFPackBytes : 0u);
        const uint8_t* pk = a.pack + static_cast<size_t>(g.d > 0 ? a.pack_idx[g.br] : 0) * a.nk * kFPackBytes;
        for (int kc = 0; kc < a.nk; ++kc) {
          mbar_wait(&in_empty[st], ph ^ 1);
          uint8_t* si = smem + st * kFInStage;
          mbar_expect_tx(&in_full[st], in_bytes);
          if (g.g0) tma_load_4d(si, &tmap_x, &in_full[st], kc * 64, 0, g.y0 - g.d, g.b);
          tma_load_4d(si + kFGroupBytes, &tmap_x, &in_full[st], kc * 64, 0, g.y0, g.b);
          if (g.g2) tma_load_4d(si + 2 * kFGroupBytes, &tmap_x, &in_full[st], kc * 64, 0, g.y0 + g.d, g.b);
          if (g.d > 0) bulk_load_1d(s0 + st * kFInStage + 3 * kFGroupBytes, pk + static_cast<size_t>(kc) * kFPackBytes, kFPackBytes, &in_full[st]);
          mbar_wait(&b_empty[st], ph ^ 1);
          mbar_expect_tx(&b_full[st], static_cast<uint32_t>(a.N) * 128u);
          tma_load_2d(smem + kFOffB + st * kFBBytes, &wmaps.m[g.br], &b_full[st], kc * 64, 0);
          if (++st == kFStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      int st = 0; uint32_t ph = 0;
      int it = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&t_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kc = 0; kc < a.nk; ++kc) {
          mbar_wait(&a_full[st], ph);
          mbar_wait(&b_full[st], ph);
          tc_fence_after();
          const uint32_t sa = s0 + kFOffA + st * kFABytes;
          const uint32_t sb = s0 + kFOffB + st * kFBBytes;
          const int k_left = a.C - kc * 64;
          const int ksteps = min(4, (k_left + 15) / 16);
          for (int k = 0; k < ksteps; ++k)
            umma_f16(d_tmem, make_sw128_desc(sa + k * 32, 16, 1024), make_sw128_desc(sb + k * 32, 16, 1024), a.idesc,
                     (kc > 0 || k > 0) ? 1u : 0u);
          umma_commit(&a_empty[st]);
          umma_commit(&b_empty[st]);
          if (++st == kFStages) { st = 0; ph ^= 1; }
        }
        umma_commit(&t_full[as]);
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue warps (2..5) =====================
    const int quad = warp & 3;
    uint4* stg = reinterpret_cast<uint4*>(smem + kFOffStg) + static_cast<size_t>(warp - 2) * 32 * 8;
    int it = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
      const ItemGeom g = decode_item(a, item);
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const long long m0 = (static_cast<long long>(g.b) * a.H + g.y0) * a.W;
      const int rows_valid = min(a.TH, a.H - g.y0) * a.W;
      const float* sc = s_pw + g.br * 512;
      const float* sh = sc + 256;
      T* out = reinterpret_cast<T*>(a.out[g.br]);
      mbar_wait(&t_full[as], aph);
      tc_fence_after();
      for (int j64 = 0; j64 < a.N; j64 += 64) {
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int j = j64 + sub * 32;
          uint32_t r[2][16];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * 256 + j;
          tmem_ld16(taddr, r[0]);
          tmem_ld16(taddr + 16, r[1]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int col = j + c * 8 + i;
              o[i] = apply_act(fmaf(__uint_as_float(r[c >> 1][(c & 1) * 8 + i]), sc[col], sh[col]), a.pw_act);
            }
            uint4 pk;
            Vec8<T>::st(reinterpret_cast<T*>(&pk), o);
            stg[lane * 8 + ((sub * 4 + c) ^ (lane & 7))] = pk;
          }
        }
        __syncwarp();
        const int cchunk = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + (lane >> 3);
          const int p = quad * 32 + row;
          if (p < rows_valid) {
            const uint4 pk = stg[row * 8 + (cchunk ^ (row & 7))];
            *reinterpret_cast<uint4*>(out + (m0 + p) * a.ldc + j64 + cchunk * 8) = pk;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[as]);
    }
  } else {
    // ===================== depthwise producers (warps 6..13) =====================
    using P = typename H2<T>::t;
    const int td = threadIdx.x - kFDwThread0;
    const int cv = td & 7;          // 8-channel vector inside the 64-channel chunk
    const int slot = td >> 3;       // 0..31
    const int n_px = a.TH * a.W;
    int st = 0; uint32_t ph = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      const ItemGeom g = decode_item(a, item);
      int prow[4], pcol[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = slot + 32 * i;
        prow[i] = p / a.W;
        pcol[i] = p - prow[i] * a.W;
      }
      for (int kc = 0; kc < a.nk; ++kc) {
        mbar_wait(&in_full[st], ph);
        const uint32_t sin = s0 + st * kFInStage;
        const uint32_t sa = s0 + kFOffA + st * kFABytes;
        if (g.d == 0) {
          mbar_wait(&a_empty[st], ph ^ 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int p = slot + 32 * i;
            if (p < n_px) {
              const uint32_t off = static_cast<uint32_t>(p) * 128u + (static_cast<uint32_t>(cv ^ (p & 7)) << 4);
              sts128(sa + off, lds128(sin + kFGroupBytes + off));
            }
          }
        } else {
          P w2[9][4];
          float scl[8], shf[8];
          {
            const uint32_t sp = sin + 3 * kFGroupBytes;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint4 u = lds128(sp + t * 128 + cv * 16);
              const P* h = reinterpret_cast<const P*>(&u);
#pragma unroll
              for (int i = 0; i < 4; ++i) w2[t][i] = h[i];
            }
            const uint4 s_a = lds128(sp + 1152 + cv * 32), s_b = lds128(sp + 1152 + cv * 32 + 16);
            const uint4 h_a = lds128(sp + 1408 + cv * 32), h_b = lds128(sp + 1408 + cv * 32 + 16);
            scl[0] = __uint_as_float(s_a.x); scl[1] = __uint_as_float(s_a.y); scl[2] = __uint_as_float(s_a.z); scl[3] = __uint_as_float(s_a.w);
            scl[4] = __uint_as_float(s_b.x); scl[5] = __uint_as_float(s_b.y); scl[6] = __uint_as_float(s_b.z); scl[7] = __uint_as_float(s_b.w);
            shf[0] = __uint_as_float(h_a.x); shf[1] = __uint_as_float(h_a.y); shf[2] = __uint_as_float(h_a.z); shf[3] = __uint_as_float(h_a.w);
            shf[4] = __uint_as_float(h_b.x); shf[5] = __uint_as_float(h_b.y); shf[6] = __uint_as_float(h_b.z); shf[7] = __uint_as_float(h_b.w);
          }
          mbar_wait(&a_empty[st], ph ^ 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int p = slot + 32 * i;
            if (p >= n_px) continue;
            const int x = pcol[i];
            const int qrow = prow[i] * a.W;
            const bool xl = x - g.d >= 0, xr = x + g.d < a.W;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              if ((ky == 0 && !g.g0) || (ky == 2 && !g.g2)) continue;
              const uint32_t gb = sin + ky * kFGroupBytes;
              P racc[4];
              {
                const int q = qrow + x;
                const uint4 u = lds128(gb + static_cast<uint32_t>(q) * 128u + (static_cast<uint32_t>(cv ^ (q & 7)) << 4));
                const P* h = reinterpret_cast<const P*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) racc[j] = __hmul2(h[j], w2[ky * 3 + 1][j]);
              }
              if (xl) {
                const int q = qrow + x - g.d;
                const uint4 u = lds128(gb + static_cast<uint32_t>(q) * 128u + (static_cast<uint32_t>(cv ^ (q & 7)) << 4));
                const P* h = reinterpret_cast<const P*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) racc[j] = __hfma2(h[j], w2[ky * 3][j], racc[j]);
              }
              if (xr) {
                const int q = qrow + x + g.d;
                const uint4 u = lds128(gb + static_cast<uint32_t>(q) * 128u + (static_cast<uint32_t>(cv ^ (q & 7)) << 4));
                const P* h = reinterpret_cast<const P*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) racc[j] = __hfma2(h[j], w2[ky * 3 + 2][j], racc[j]);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = H2<T>::unpack(racc[j]);
                acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
              }
            }
            uint4 o;
            P* oh = reinterpret_cast<P*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              oh[j] = H2<T>::pack(apply_act(fmaf(acc[2 * j], scl[2 * j], shf[2 * j]), a.dw_act),
                                  apply_act(fmaf(acc[2 * j + 1], scl[2 * j + 1], shf[2 * j + 1]), a.dw_act));
            sts128(sa + static_cast<uint32_t>(p) * 128u + (static_cast<uint32_t>(cv ^ (p & 7)) << 4), o);
          }
        }
        fence_proxy_async();      // generic-proxy writes of the A tile -> visible to tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) { mbar_arrive(&a_full[st]); mbar_arrive(&in_empty[st]); }
        if (++st == kFStages) { st = 0; ph ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// packed per-(branch, 64-channel chunk) depthwise parameters: [9 taps][64 ch] 16-bit | scale[64] f32 | shift[64] f32
template <typename T>
__global__ void aspp_pack_kernel(int C, int nk, const float* w, const float* scale, const float* shift, uint8_t* pack) {
  const int kc = blockIdx.x, ch = threadIdx.x;     // 64 threads
  const int c = kc * 64 + ch;
  uint8_t* dst = pack + static_cast<size_t>(kc) * kFPackBytes;
  T* wt = reinterpret_cast<T*>(dst);
  float* sc = reinterpret_cast<float*>(dst + 1152);
  float* sh = reinterpret_cast<float*>(dst + 1408);
  for (int t = 0; t < 9; ++t) Act<T>::st(&wt[t * 64 + ch], c < C ? w[t * C + c] : 0.f);
  sc[ch] = (c < C) ? (scale ? scale[c] : 1.f) : 0.f;
  sh[ch] = (c < C && shift) ? shift[c] : 0.f;
}

}  // namespace dlb

extern "C" int64_t dlb_sepconv_pack_bytes(int C, int n_branches) {
  return static_cast<int64_t>(n_branches) * ((C + 63) / 64) * dlb::kFPackBytes;
}

extern "C" int dlb_sepconv_pack_dw(int C, int dtype, int n_branches, const float* const* w_dw, const float* const* scale,
                                   const float* const* shift, void* pack, void* stream) {
  using namespace dlb;
  DLB_REQUIRE(w_dw && pack && n_branches > 0 && n_branches <= kFMaxBr, "sepconv_pack_dw: bad arguments");
  DLB_REQUIRE(dtype == DLB_F16 || dtype == DLB_BF16, "sepconv_pack_dw: 16-bit storage types only");
  DLB_REQUIRE((reinterpret_cast<uintptr_t>(pack) & 15) == 0, "sepconv_pack_dw: pack must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nk = (C + 63) / 64;
  for (int i = 0; i < n_branches; ++i) {
    DLB_REQUIRE(w_dw[i], "sepconv_pack_dw: null depthwise kernel");
    uint8_t* dst = static_cast<uint8_t*>(pack) + static_cast<size_t>(i) * nk * kFPackBytes;
    const float* sc = scale ? scale[i] : nullptr;
    const float* sh = shift ? shift[i] : nullptr;
    if (dtype == DLB_F16) aspp_pack_kernel<__half><<<nk, 64, 0, st>>>(C, nk, w_dw[i], sc, sh, dst);
    else aspp_pack_kernel<__nv_bfloat16><<<nk, 64, 0, st>>>(C, nk, w_dw[i], sc, sh, dst);
    g_launches++;
  }
  return check_launch("aspp_pack_kernel");
}

extern "C" int dlb_sepconv_fused_fwd(const dlb_sepconv_fused_params* p, void* stream) {
  using namespace dlb;
  DLB_REQUIRE(p && p->x, "sepconv_fused_fwd: null pointer");
  DLB_REQUIRE(p->dtype == DLB_F16 || p->dtype == DLB_BF16, "sepconv_fused_fwd: 16-bit storage types only (f32 takes the unfused exact path)");
  DLB_REQUIRE(p->n_branches >= 1 && p->n_branches <= kFMaxBr, "sepconv_fused_fwd: 1..4 branches");
  DLB_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0 && p->W <= 128, "sepconv_fused_fwd: W must be <= 128 (got %d)", p->W);
  DLB_REQUIRE(p->C % 8 == 0 && p->C >= 16, "sepconv_fused_fwd: C must be a multiple of 8");
  DLB_REQUIRE(p->N % 64 == 0 && p->N >= 64 && p->N <= 256, "sepconv_fused_fwd: N must be 64, 128, 192 or 256");
  DLB_REQUIRE(p->ldc % 8 == 0 && p->ldc >= p->N, "sepconv_fused_fwd: ldc must be a multiple of 8 and >= N");
  DLB_REQUIRE((reinterpret_cast<uintptr_t>(p->x) & 15) == 0, "sepconv_fused_fwd: x must be 16-byte aligned");
  FusedArgs a{};
  a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.N = p->N;
  a.TH = 128 / p->W; if (a.TH > p->H) a.TH = p->H;
  a.tiles_y = (p->H + a.TH - 1) / a.TH;
  a.n_br = p->n_branches;
  a.n_items = p->B * a.tiles_y * a.n_br;
  a.nk = (p->C + 63) / 64;
  a.pack = static_cast<const uint8_t*>(p->dw_pack);
  a.ldc = p->ldc; a.dw_act = p->dw_act; a.pw_act = p->pw_act;
  a.idesc = make_idesc(p->dtype == DLB_BF16 ? 1 : 0, 128, p->N, 0, 0);
  WMaps wm;
  std::memset(&wm, 0, sizeof(wm));
  int n_dw = 0;
  for (int i = 0; i < a.n_br; ++i) {
    DLB_REQUIRE(p->w_pw[i] && p->out[i], "sepconv_fused_fwd: branch %d: null pointer", i);
    DLB_REQUIRE(p->rates[i] >= 0, "sepconv_fused_fwd: negative rate");
    DLB_REQUIRE((reinterpret_cast<uintptr_t>(p->out[i]) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w_pw[i]) & 15) == 0,
                "sepconv_fused_fwd: branch %d: pointers must be 16-byte aligned", i);
    a.rate[i] = p->rates[i];
    a.pack_idx[i] = p->rates[i] > 0 ? n_dw++ : 0;
    a.pw_scale[i] = p->pw_scale[i]; a.pw_shift[i] = p->pw_shift[i];
    a.out[i] = p->out[i];
    int rc = make_tmap_2d(&wm.m[i], p->dtype, p->w_pw[i], p->N, p->C, p->C, p->N, 64);
    if (rc) return rc;
  }
  DLB_REQUIRE(n_dw == 0 || p->dw_pack, "sepconv_fused_fwd: dw_pack missing");
  DLB_REQUIRE((reinterpret_cast<uintptr_t>(p->dw_pack) & 15) == 0, "sepconv_fused_fwd: dw_pack must be 16-byte aligned");
  CUtensorMap tx;
  int rc = make_tmap_nhwc_sw128(&tx, p->dtype, p->x, p->B, p->H, p->W, p->C, p->W, a.TH);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = a.n_items < num_sms() ? a.n_items : num_sms();
  if (p->dtype == DLB_F16) {
    DLB_CUDA(cudaFuncSetAttribute(sepconv_fused_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    sepconv_fused_kernel<__half><<<grid, kFThreads, kFSmem, st>>>(tx, wm, a);
  } else {
    DLB_CUDA(cudaFuncSetAttribute(sepconv_fused_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    sepconv_fused_kernel<__nv_bfloat16><<<grid, kFThreads, kFSmem, st>>>(tx, wm, a);
  }
  g_launches++;
  return check_launch("sepconv_fused_kernel");
}

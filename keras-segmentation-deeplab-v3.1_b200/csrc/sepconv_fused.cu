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
constexpr int kFDwWarps = 8;
constexpr int kFThreads = 192 + 32 * kFDwWarps;   // warp 0 TMA + scheduler, warp 1 MMA, warps 2-5 epilogue, then the depthwise warps
constexpr int kFDwThread0 = 192;
constexpr int kFDwSlots = kFDwWarps * 4;          // pixels per pass (8 lanes = the 8 channel vectors of one pixel)
constexpr int kFDwIter = 128 / kFDwSlots;
constexpr int kFGroupBytes = 16384;     // 128 pixels x 64 channels x 2 B
constexpr int kFPackBytes = 1408;       // (9 taps + BN scale + BN shift) x 64 ch x 2 B
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
constexpr int kFSmem = 1024 + kFOffBar + 512;
constexpr int kFRing = 4;               // scheduler ring depth
constexpr int kFConsumers = 1 + 4 + kFDwWarps;   // MMA thread + epilogue warps + depthwise warps read every ticket

struct WMaps { CUtensorMap m[kFMaxBr]; };

struct FusedArgs {
  int B, H, W, C, N;
  int TH, tiles_y, n_items, nk, n_br;
  int rate[kFMaxBr];
  int pack_idx[kFMaxBr];
  int br_order[kFMaxBr];          // branches by descending cost (tickets are handed out longest-first inside an image)
  const uint8_t* pack;
  const float* pw_scale[kFMaxBr];
  const float* pw_shift[kFMaxBr];
  void* out[kFMaxBr];
  const void* res[kFMaxBr];       // optional residual added after the pointwise BN / activation (MobileNetV2 skip)
  int ldc, ldr, dw_act, pw_act;
  int n_valid;                    // real output channels (N is padded to a multiple of 64: zero-filled weight rows)
  uint32_t idesc;
  unsigned int* ticket;           // zeroed by the host before every launch
};

__device__ unsigned int g_sepconv_ticket;

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <typename T> struct H2;
template <> struct H2<__half> {
  using t = __half2;
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2half2_rn(a, b); }
  static __device__ __forceinline__ t zero() { return __float2half2_rn(0.f); }
};
template <> struct H2<__nv_bfloat16> {
  using t = __nv_bfloat162;
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static __device__ __forceinline__ t zero() { return __float2bfloat162_rn(0.f); }
};

// work item (ticket) = (image, tile of TH rows, branch); inside an image the expensive branches come first
struct ItemGeom {
  int br, b, y0, d;
  bool g0, g2;      // row groups y0-d.. / y0+d.. intersect the image
};
__device__ __forceinline__ ItemGeom decode_item(const FusedArgs& a, int item) {
  ItemGeom g;
  const int per_img = a.tiles_y * a.n_br;
  g.b = item / per_img;
  const int idx = item - g.b * per_img;
  const int rank = idx / a.tiles_y;
  g.br = a.br_order[rank];
  g.y0 = (idx - rank * a.tiles_y) * a.TH;
  g.d = a.rate[g.br];
  g.g0 = g.d > 0 && g.y0 - g.d + a.TH - 1 >= 0;
  g.g2 = g.d > 0 && g.y0 + g.d < a.H;
  return g;
}

template <typename T>
__global__ void __launch_bounds__(kFThreads, 1)
sepconv_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ WMaps wmaps, const FusedArgs a) {
  pdl_prologue();
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t s0 = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFOffBar);
  uint64_t* in_full = bars;            // [2]  TMA -> depthwise warps
  uint64_t* in_empty = bars + 2;       // [2]
  uint64_t* a_full = bars + 4;         // [2]  depthwise warps -> MMA
  uint64_t* a_empty = bars + 6;        // [2]
  uint64_t* b_full = bars + 8;         // [2]  TMA -> MMA (pointwise weights)
  uint64_t* b_empty = bars + 10;       // [2]
  uint64_t* t_full = bars + 12;        // [2]  accumulator stage complete -> epilogue
  uint64_t* t_empty = bars + 14;       // [2]  accumulator stage drained -> MMA
  uint64_t* q_full = bars + 16;        // [4]  scheduler ring
  uint64_t* q_empty = bars + 20;       // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  volatile int* q_item = reinterpret_cast<volatile int*>(bars + 25);   // [4]
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
    for (int i = 0; i < kFRing; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], kFConsumers); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < a.n_br * 512; i += kFThreads) {
    const int br = i >> 9, j = i & 511;
    float v;
    if (j < 256) v = (j < a.n_valid && a.pw_scale[br]) ? a.pw_scale[br][j] : 1.f;
    else v = (j - 256 < a.n_valid && a.pw_shift[br]) ? a.pw_shift[br][j - 256] : 0.f;
    s_pw[i] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t box_bytes = static_cast<uint32_t>(a.TH) * a.W * 128u;

  // Dynamic tile scheduler: the producer thread draws tickets from a global counter and publishes them in a 4-deep
  // shared-memory ring; every consumer role reads ticket n from slot n % 4.  -1 ends the kernel.  (A static stride
  // gave CTA c the same branch every round -- 148 % 4 == 0 -- and a 40 % idle tail.)
  auto next_item = [&](int n) -> int {
    const int slot = n & (kFRing - 1);
    mbar_wait(&q_full[slot], (n / kFRing) & 1);
    const int item = q_item[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(&q_empty[slot]);
    return item;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== scheduler + TMA producer =====================
      int st = 0; uint32_t ph = 0;
      for (int n = 0;; ++n) {
        const int slot = n & (kFRing - 1);
        mbar_wait(&q_empty[slot], ((n / kFRing) & 1) ^ 1);
        const unsigned int t = atomicAdd(a.ticket, 1u);
        const int item = t < static_cast<unsigned int>(a.n_items) ? static_cast<int>(t) : -1;
        q_item[slot] = item;
        mbar_arrive(&q_full[slot]);
        if (item < 0) break;
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
      for (int n = 0;; ++n) {
        const int slot = n & (kFRing - 1);
        mbar_wait(&q_full[slot], (n / kFRing) & 1);
        const int item = q_item[slot];
        mbar_arrive(&q_empty[slot]);
        if (item < 0) break;
        const int as = n & 1;
        mbar_wait(&t_empty[as], ((n >> 1) & 1) ^ 1);
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
    for (int n = 0;; ++n) {
      const int item = next_item(n);
      if (item < 0) break;
      const ItemGeom g = decode_item(a, item);
      const int as = n & 1;
      const long long m0 = (static_cast<long long>(g.b) * a.H + g.y0) * a.W;
      const int rows_valid = min(a.TH, a.H - g.y0) * a.W;
      const float* sc = s_pw + g.br * 512;
      const float* sh = sc + 256;
      T* out = reinterpret_cast<T*>(a.out[g.br]);
      const T* res = reinterpret_cast<const T*>(a.res[g.br]);
      mbar_wait(&t_full[as], (n >> 1) & 1);
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
          if (p < rows_valid && j64 + cchunk * 8 < a.n_valid) {
            uint4 pk = stg[row * 8 + (cchunk ^ (row & 7))];
            if (res) {
              float o[8], rr[8];
              Vec8<T>::ld(reinterpret_cast<const T*>(&pk), o);
              Vec8<T>::ld(res + (m0 + p) * a.ldr + j64 + cchunk * 8, rr);
#pragma unroll
              for (int q = 0; q < 8; ++q) o[q] += rr[q];
              Vec8<T>::st(reinterpret_cast<T*>(&pk), o);
            }
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
    // ===================== depthwise producers =====================
    // Everything that depends only on (pixel, rate) -- the three swizzled tap-column offsets and their validity -- is
    // computed once per item; the chunk loop is loads + packed HFMA2 only.  The 9 taps are three independent 16-bit
    // row chains (HFMA2), summed, then BN affine + activation on packed 16-bit (the result is rounded to 16 bit for
    // the tensor core anyway).
    using P = typename H2<T>::t;
    const int td = threadIdx.x - kFDwThread0;
    const int cv = td & 7;          // 8-channel vector inside the 64-channel chunk
    const int slot = td >> 3;       // 0..kFDwSlots-1
    const int n_px = a.TH * a.W;
    const P zero2 = H2<T>::zero();
    const P hi2 = H2<T>::pack(a.dw_act == DLB_ACT_RELU6 ? 6.f : 65000.f, a.dw_act == DLB_ACT_RELU6 ? 6.f : 65000.f);
    int st = 0; uint32_t ph = 0;
    for (int n = 0;; ++n) {
      const int item = next_item(n);
      if (item < 0) break;
      const ItemGeom g = decode_item(a, item);
      uint32_t offC[kFDwIter], offL[kFDwIter], offR[kFDwIter];
      bool vp[kFDwIter], vl[kFDwIter], vr[kFDwIter];
#pragma unroll
      for (int i = 0; i < kFDwIter; ++i) {
        const int p = slot + kFDwSlots * i;
        const int row = p / a.W, x = p - row * a.W;
        vp[i] = p < n_px;
        vl[i] = vp[i] && x - g.d >= 0;
        vr[i] = vp[i] && x + g.d < a.W;
        const int ql = vl[i] ? p - g.d : p, qr = vr[i] ? p + g.d : p;
        offC[i] = static_cast<uint32_t>(p) * 128u + (static_cast<uint32_t>(cv ^ (p & 7)) << 4);
        offL[i] = static_cast<uint32_t>(ql) * 128u + (static_cast<uint32_t>(cv ^ (ql & 7)) << 4);
        offR[i] = static_cast<uint32_t>(qr) * 128u + (static_cast<uint32_t>(cv ^ (qr & 7)) << 4);
      }
      for (int kc = 0; kc < a.nk; ++kc) {
        mbar_wait(&in_full[st], ph);
        const uint8_t* sin = smem + st * kFInStage;
        uint8_t* sa = smem + kFOffA + st * kFABytes;
        if (g.d == 0) {
          mbar_wait(&a_empty[st], ph ^ 1);
#pragma unroll
          for (int i = 0; i < kFDwIter; ++i)
            if (vp[i]) *reinterpret_cast<uint4*>(sa + offC[i]) = *reinterpret_cast<const uint4*>(sin + kFGroupBytes + offC[i]);
        } else {
          P w2[9][4], scl[4], shf[4];
          {
            const uint8_t* sp = sin + 3 * kFGroupBytes + cv * 16;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint4 u = *reinterpret_cast<const uint4*>(sp + t * 128);
              const P* h = reinterpret_cast<const P*>(&u);
#pragma unroll
              for (int j = 0; j < 4; ++j) w2[t][j] = h[j];
            }
            const uint4 us = *reinterpret_cast<const uint4*>(sp + 9 * 128), uh = *reinterpret_cast<const uint4*>(sp + 10 * 128);
            const P* hs = reinterpret_cast<const P*>(&us);
            const P* hh = reinterpret_cast<const P*>(&uh);
#pragma unroll
            for (int j = 0; j < 4; ++j) { scl[j] = hs[j]; shf[j] = hh[j]; }
          }
          mbar_wait(&a_empty[st], ph ^ 1);
#pragma unroll
          for (int i = 0; i < kFDwIter; ++i) {
            if (!vp[i]) continue;
            P racc[3][4];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
              for (int j = 0; j < 4; ++j) racc[ky][j] = zero2;
              if ((ky == 0 && !g.g0) || (ky == 2 && !g.g2)) continue;      // item-uniform
              const uint8_t* gb = sin + ky * kFGroupBytes;
              const uint4 uc = *reinterpret_cast<const uint4*>(gb + offC[i]);
              uint4 ul = make_uint4(0u, 0u, 0u, 0u), ur = make_uint4(0u, 0u, 0u, 0u);
              if (vl[i]) ul = *reinterpret_cast<const uint4*>(gb + offL[i]);
              if (vr[i]) ur = *reinterpret_cast<const uint4*>(gb + offR[i]);
              const P* hc = reinterpret_cast<const P*>(&uc);
              const P* hl = reinterpret_cast<const P*>(&ul);
              const P* hr = reinterpret_cast<const P*>(&ur);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                racc[ky][j] = __hfma2(hr[j], w2[ky * 3 + 2][j], __hfma2(hl[j], w2[ky * 3][j], __hmul2(hc[j], w2[ky * 3 + 1][j])));
            }
            uint4 o;
            P* oh = reinterpret_cast<P*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              P v = __hfma2(__hadd2(__hadd2(racc[0][j], racc[2][j]), racc[1][j]), scl[j], shf[j]);
              if (a.dw_act != DLB_ACT_NONE) v = __hmin2(__hmax2(v, zero2), hi2);
              oh[j] = v;
            }
            *reinterpret_cast<uint4*>(sa + offC[i]) = o;
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

// packed per-(branch, 64-channel chunk) depthwise parameters, all 16-bit: [9 taps][64 ch] | scale[64] | shift[64]
template <typename T>
__global__ void aspp_pack_kernel(int C, int nk, const float* w, const float* scale, const float* shift, uint8_t* pack) {
  pdl_prologue();
  const int kc = blockIdx.x, ch = threadIdx.x;     // 64 threads
  const int c = kc * 64 + ch;
  uint8_t* dst = pack + static_cast<size_t>(kc) * kFPackBytes;
  T* wt = reinterpret_cast<T*>(dst);
  for (int t = 0; t < 9; ++t) Act<T>::st(&wt[t * 64 + ch], c < C ? w[t * C + c] : 0.f);
  Act<T>::st(&wt[9 * 64 + ch], (c < C) ? (scale ? scale[c] : 1.f) : 0.f);
  Act<T>::st(&wt[10 * 64 + ch], (c < C && shift) ? shift[c] : 0.f);
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
    if (dtype == DLB_F16) launch_k(aspp_pack_kernel<__half>, nk, 64, 0, st, C, nk, w_dw[i], sc, sh, dst);
    else launch_k(aspp_pack_kernel<__nv_bfloat16>, nk, 64, 0, st, C, nk, w_dw[i], sc, sh, dst);
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
  DLB_REQUIRE(p->N % 8 == 0 && p->N >= 8 && p->N <= 256, "sepconv_fused_fwd: N must be a multiple of 8, 8..256 (got %d)", p->N);
  DLB_REQUIRE(p->ldc % 8 == 0 && p->ldc >= p->N, "sepconv_fused_fwd: ldc must be a multiple of 8 and >= N");
  DLB_REQUIRE((reinterpret_cast<uintptr_t>(p->x) & 15) == 0, "sepconv_fused_fwd: x must be 16-byte aligned");
  FusedArgs a{};
  // accumulator / weight-tile width: N rounded up to a 64-column epilogue block; the weight rows past N are zero-filled
  // by the tensor map (out-of-bounds box rows), the store masks them
  const int n_pad = (p->N + 63) / 64 * 64;
  a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.N = n_pad; a.n_valid = p->N;
  a.TH = 128 / p->W; if (a.TH > p->H) a.TH = p->H;
  a.tiles_y = (p->H + a.TH - 1) / a.TH;
  a.n_br = p->n_branches;
  a.n_items = p->B * a.tiles_y * a.n_br;
  a.nk = (p->C + 63) / 64;
  a.pack = static_cast<const uint8_t*>(p->dw_pack);
  a.ldc = p->ldc; a.ldr = p->ldr; a.dw_act = p->dw_act; a.pw_act = p->pw_act;
  a.idesc = make_idesc(p->dtype == DLB_BF16 ? 1 : 0, 128, n_pad, 0, 0);
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
    a.res[i] = p->res[i];
    DLB_REQUIRE(p->res[i] == nullptr || (p->ldr % 8 == 0 && p->ldr >= p->N && (reinterpret_cast<uintptr_t>(p->res[i]) & 15) == 0),
                "sepconv_fused_fwd: branch %d: residual needs ldr %% 8 == 0, ldr >= N and 16-byte alignment", i);
    int rc = make_tmap_2d(&wm.m[i], p->dtype, p->w_pw[i], p->N, p->C, p->C, n_pad, 64);
    if (rc) return rc;
  }
  DLB_REQUIRE(n_dw == 0 || p->dw_pack, "sepconv_fused_fwd: dw_pack missing");
  // longest-first ticket order inside an image: small positive rates keep the most taps inside the map, rate 0 has none
  for (int i = 0; i < a.n_br; ++i) a.br_order[i] = i;
  for (int i = 1; i < a.n_br; ++i)
    for (int j = i; j > 0; --j) {
      const int ra = a.rate[a.br_order[j - 1]], rb = a.rate[a.br_order[j]];
      const bool swap = (ra == 0 && rb != 0) || (ra != 0 && rb != 0 && rb < ra);
      if (!swap) break;
      const int t = a.br_order[j]; a.br_order[j] = a.br_order[j - 1]; a.br_order[j - 1] = t;
    }
  DLB_REQUIRE((reinterpret_cast<uintptr_t>(p->dw_pack) & 15) == 0, "sepconv_fused_fwd: dw_pack must be 16-byte aligned");
  CUtensorMap tx;
  int rc = make_tmap_nhwc_sw128(&tx, p->dtype, p->x, p->B, p->H, p->W, p->C, p->W, a.TH);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = a.n_items < num_sms() ? a.n_items : num_sms();
  // dynamic tile scheduler: one device counter, zeroed in stream order before every launch (a memset node under graph
  // capture); launches of this kernel must therefore not overlap on different streams
  DLB_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&a.ticket), g_sepconv_ticket));
  DLB_CUDA(cudaMemsetAsync(a.ticket, 0, sizeof(unsigned int), st));
  if (p->dtype == DLB_F16) {
    DLB_CUDA(cudaFuncSetAttribute(sepconv_fused_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    launch_k(sepconv_fused_kernel<__half>, grid, kFThreads, kFSmem, st, tx, wm, a);
  } else {
    DLB_CUDA(cudaFuncSetAttribute(sepconv_fused_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    launch_k(sepconv_fused_kernel<__nv_bfloat16>, grid, kFThreads, kFSmem, st, tx, wm, a);
  }
  g_launches++;
  return check_launch("sepconv_fused_kernel");
}

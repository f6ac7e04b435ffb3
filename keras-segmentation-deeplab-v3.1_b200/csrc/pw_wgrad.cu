// Weight gradient of the pointwise convolutions:  dW[K, N] = A[M, K]^T * dY[M, N]   (reduction over M = B*H*W)
//
// This is the backward counterpart of pw_gemm.cu (TF autodiff of Conv2D 1x1 in the reference's fit_generator
// step, utils.py:231-241).  The reduction dimension (pixels) is the *non*-contiguous one of both operands, so
// the natural NHWC tiles [rows, 64 channels] are fed to tcgen05.mma as MN-major operands (same TMA box, same
// 128-byte swizzle as the forward GEMM, only the descriptor's major bit and LBO change).
//
//   * output tile  : 128 (K channels) x up to 512 (N channels) fp32 accumulators in TMEM
//   * split-M      : each output tile's pixel range is split over S CTAs so that tiles*S ~ #SMs; every CTA adds its
//                    partial tile into dW with 16-byte vector reductions (red.global.add.v4.f32, resolved in L2);
//                    dW is zeroed first when beta = 0.  (A [S, K, N] workspace + reduce kernel was measured at
//                    2.5 % of the training step and doubled the HBM traffic of the 6C-wide layers.)
//   * pipeline     : warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 TMEM -> global epilogue
#include <atomic>

#include "common.cuh"

namespace dlb {

extern std::atomic<long long> g_launches;
int make_tmap_2d(CUtensorMap* map, int dtype, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols);

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 8;

struct WgArgs {
  int M, N, K;
  int rows_per_stage;          // R: reduction rows per pipeline stage (32 or 64)
  int n_boxes;                 // 64-wide dY boxes per stage (N group / 64)
  int n_chunks, chunk_n;       // UMMA N chunks per group: chunks 0..n_chunks-2 are chunk_n wide, the last one chunk_last
  int chunk_last, group_cols;
  uint32_t idesc_last;
  int k_tiles, n_groups, splits;
  int rows_per_split;          // multiple of R
  int num_stages;
  uint32_t stage_bytes, a_bytes;
  uint32_t idesc;
  float* dW;                   // [K, N] fp32, accumulated with vector reductions (red.global.add.v4.f32)
  const float* a_scale;        // A-operand transform: A := act(A * a_scale[k] + a_shift[k]) on the landed tile (or NULL)
  const float* a_shift;
  int a_act;
};

// kXform: the (otherwise idle until the end) epilogue warps apply BatchNorm-affine + activation to every landed A tile
// in place -- the weight gradient of the project conv reads the RAW depthwise output (see xform_row_sw128).
template <typename T, bool kXform>
__global__ void __launch_bounds__(kWgThreads, 1)
pw_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_y,
                   const WgArgs g) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset, not a uintptr_t round trip: keeps LDS/STS
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* tail = smem + static_cast<size_t>(g.num_stages) * g.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kWgMaxStages;
  uint64_t* done_bar = empty_bar + kWgMaxStages;
  uint64_t* xf_bar = done_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xf_bar + kWgMaxStages);
  float* s_asc = reinterpret_cast<float*>(tmem_slot + 4);       // [128] scales, [128] shifts of this CTA's K tile
  float* s_ash = s_asc + 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_y);
    for (int i = 0; i < g.num_stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(done_bar, 1);
    if (kXform) for (int i = 0; i < g.num_stages; ++i) mbar_init(&xf_bar[i], 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // barrier init / TMEM allocation above overlap the preceding kernel's tail

  // work item of this CTA
  int wi = blockIdx.x;
  const int split = wi % g.splits; wi /= g.splits;
  const int ng = wi % g.n_groups;
  const int kt = wi / g.n_groups;
  const int k0 = kt * 128;
  const int n0 = ng * g.n_boxes * 64;
  const int m_begin = split * g.rows_per_split;
  const int m_end = min(g.M, m_begin + g.rows_per_split);
  const int R = g.rows_per_stage;
  const int iters = m_end > m_begin ? (m_end - m_begin + R - 1) / R : 0;
  if (kXform) {
    if (threadIdx.x < 128) {
      const int k = k0 + threadIdx.x;
      s_asc[threadIdx.x] = k < g.K ? g.a_scale[k] : 0.f;       // channels past K are TMA zero fill and stay zero
      s_ash[threadIdx.x] = k < g.K ? g.a_shift[k] : 0.f;
    }
    __syncthreads();
  }

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + static_cast<size_t>(stage) * g.stage_bytes;
        uint8_t* sy = sa + g.a_bytes;
        const int m = m_begin + it * R;
        mbar_expect_tx(&full_bar[stage], g.a_bytes + g.n_boxes * R * 128);
        tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m);
        tma_load_2d(sa + R * 128, &tmap_a, &full_bar[stage], k0 + 64, m);
        for (int bx = 0; bx < g.n_boxes; ++bx) tma_load_2d(sy + bx * R * 128, &tmap_y, &full_bar[stage], n0 + bx * 64, m);
        if (++stage == g.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(kXform ? &xf_bar[stage] : &full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * g.stage_bytes);
        const uint32_t sy = sa + g.a_bytes;
        for (int ks = 0; ks < R / 16; ++ks) {
          // A operand: 128 channels = two 64-wide MN blocks, LBO = one box (R*128 B); 16 reduction rows per MMA
          const uint64_t adesc = make_sw128_desc(sa + ks * 2048, R * 128, 1024);
          for (int c = 0; c < g.n_chunks; ++c) {
            const uint64_t bdesc = make_sw128_desc(sy + c * (g.chunk_n / 64) * R * 128 + ks * 2048, R * 128, 1024);
            umma_f16(tmem_base + c * g.chunk_n, adesc, bdesc, c + 1 == g.n_chunks ? g.idesc_last : g.idesc,
                     (it > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == g.num_stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else {
    if constexpr (kXform) {
      // the two landed [R x 64 channel] boxes: a thread owns rows r and r + R/2 of one box and half (R = 64) or all
      // (R = 32) of their 16-byte chunks
      const int t = threadIdx.x - 64;
      const int box = t >> 6, tr = t & 63;
      const uint32_t asc = smem_u32(s_asc), ash = smem_u32(s_ash);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * g.stage_bytes);
        if (R == 64) xform_rows2<T, 4>(sa + box * R * 128, tr & 31, 32, (tr >> 5) * 4, asc + box * 256, ash + box * 256, g.a_act);
        else if (tr < 16 || (tr >= 32 && tr < 48)) xform_rows2<T, 4>(sa + box * R * 128, tr & 15, 16, (tr >> 5) * 4, asc + box * 256, ash + box * 256, g.a_act);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_bar[stage]);
        if (++stage == g.num_stages) { stage = 0; phase ^= 1; }
      }
    }
    const int quad = warp & 3;
    const int k = k0 + quad * 32 + lane;
    float* dst_row = g.dW + static_cast<size_t>(k) * g.N;
    const int ncols = g.group_cols;
    if (iters > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    for (int j = 0; iters > 0 && j < ncols; j += 16) {
      const int n = n0 + j;
      if (n >= g.N) break;
      uint32_t r[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + j, r);
      tmem_ld_wait();
      if (k < g.K) {
        // split-M partial sums go straight into dW: one 16-byte reduction per 4 columns (the [splits, K, N] workspace
        // round trip + reduce kernel cost as much HBM traffic as the GEMM operands on the 6C-wide layers)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int nn = n + q * 4;
          if (nn + 4 <= g.N && (g.N & 3) == 0) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + nn), "f"(__uint_as_float(r[q * 4])),
                         "f"(__uint_as_float(r[q * 4 + 1])), "f"(__uint_as_float(r[q * 4 + 2])),
                         "f"(__uint_as_float(r[q * 4 + 3]))
                         : "memory");
          } else {
            for (int i = 0; i < 4; ++i)
              if (nn + i < g.N) atomicAdd(dst_row + nn + i, __uint_as_float(r[q * 4 + i]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// exact fp32 SIMT path: blocks (k tile, n tile, split), fp32 atomics into dW
__global__ void __launch_bounds__(256) pw_wgrad_simt_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                            const float* __restrict__ dY, int ldy, float* __restrict__ dW,
                                                            int ldw, int rows_per_split, const float* __restrict__ a_scale,
                                                            const float* __restrict__ a_shift, int a_act) {
  pdl_prologue();
  constexpr int T = 64, TR = 16;
  __shared__ float sA[TR][T + 4];
  __shared__ float sB[TR][T + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.x * T, n0 = blockIdx.y * T;
  const int m_begin = blockIdx.z * rows_per_split, m_end = min(M, m_begin + rows_per_split);
  float acc[4][4] = {};
  const int lr = tid >> 4, lc = (tid & 15) * 4;
  for (int m0 = m_begin; m0 < m_end; m0 += TR) {
    const int m = m0 + lr;
    float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
    if (m < m_end) {
      const float* ap = A + static_cast<size_t>(m) * lda + k0 + lc;
      const float* bp = dY + static_cast<size_t>(m) * ldy + n0 + lc;
      if (k0 + lc + 3 < K) a = *reinterpret_cast<const float4*>(ap);
      else { float t[4] = {0, 0, 0, 0}; for (int i = 0; i < 4; ++i) if (k0 + lc + i < K) t[i] = ap[i]; a = make_float4(t[0], t[1], t[2], t[3]); }
      if (a_scale) {       // A-operand transform (BatchNorm affine + activation of the producing layer)
        float t[4] = {a.x, a.y, a.z, a.w};
        for (int i = 0; i < 4; ++i)
          t[i] = k0 + lc + i < K ? apply_act(fmaf(t[i], a_scale[k0 + lc + i], a_shift[k0 + lc + i]), a_act) : 0.f;
        a = make_float4(t[0], t[1], t[2], t[3]);
      }
      if (n0 + lc + 3 < N) b = *reinterpret_cast<const float4*>(bp);
      else { float t[4] = {0, 0, 0, 0}; for (int i = 0; i < 4; ++i) if (n0 + lc + i < N) t[i] = bp[i]; b = make_float4(t[0], t[1], t[2], t[3]); }
    }
    *reinterpret_cast<float4*>(&sA[lr][lc]) = a;
    *reinterpret_cast<float4*>(&sB[lr][lc]) = b;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[r][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sB[r][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= K) continue;
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) atomicAdd(&dW[static_cast<size_t>(k) * ldw + n], acc[i][j]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(int M, int N, const T* __restrict__ dY, int ldy, float* __restrict__ out) {
  pdl_prologue();
  // one block per 32 columns-slab; threads stride rows
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  __shared__ float s[8][33];
  float acc = 0.f;
  if (n < N)
    for (int m = blockIdx.y * 8 + rl; m < M; m += gridDim.y * 8) acc += Act<T>::ld(dY + static_cast<size_t>(m) * ldy + n);
  s[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && n < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x & 31];
    atomicAdd(&out[n], t);
  }
}

}  // namespace dlb

using namespace dlb;

extern "C" int64_t dlb_pw_wgrad_workspace_bytes(int M, int N, int K) {
  // no scratch is needed any more (partials are reduced in L2); kept in the ABI, callers may pass workspace = NULL
  (void)M; (void)N; (void)K;
  return 0;
}

extern "C" int dlb_pw_wgrad(const dlb_pw_wgrad_params* p, void* stream) {
  DLB_REQUIRE(p && p->A && p->dY && p->dW, "pw_wgrad: null pointer");
  DLB_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0 && p->ldw >= p->N, "pw_wgrad: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dbias) {
    DLB_CUDA(cudaMemsetAsync(p->dbias, 0, sizeof(float) * p->N, st));
    dim3 grid((p->N + 31) / 32, 128);
    if (p->dtype == DLB_F16) launch_k(colsum_kernel<__half>, grid, 256, 0, st, p->M, p->N, (const __half*)p->dY, p->ldy, p->dbias);
    else if (p->dtype == DLB_BF16) launch_k(colsum_kernel<__nv_bfloat16>, grid, 256, 0, st, p->M, p->N, (const __nv_bfloat16*)p->dY, p->ldy, p->dbias);
    else launch_k(colsum_kernel<float>, grid, 256, 0, st, p->M, p->N, (const float*)p->dY, p->ldy, p->dbias);
    g_launches++;
  }
  if (p->dtype == DLB_F32) {
    DLB_REQUIRE(p->lda % 4 == 0 && p->ldy % 4 == 0, "pw_wgrad(f32): lda/ldy must be multiples of 4");
    if (p->beta == 0.f) DLB_CUDA(cudaMemsetAsync(p->dW, 0, sizeof(float) * p->K * p->ldw, st));
    const int kt = (p->K + 63) / 64, nt = (p->N + 63) / 64;
    int splits = (2 * num_sms() + kt * nt - 1) / (kt * nt);
    int rps = ((p->M + splits - 1) / splits + 15) / 16 * 16;
    splits = (p->M + rps - 1) / rps;
    dim3 grid(kt, nt, splits);
    launch_k(pw_wgrad_simt_kernel, grid, 256, 0, st, p->M, p->N, p->K, (const float*)p->A, p->lda, (const float*)p->dY,
                                               p->ldy, p->dW, p->ldw, rps, p->a_scale, p->a_shift, p->a_act);
    g_launches++;
    return check_launch("pw_wgrad_simt_kernel");
  }
  DLB_REQUIRE(p->ldw == p->N, "pw_wgrad: ldw must equal N on the tensor-core path");
  DLB_REQUIRE(p->beta == 0.f || p->beta == 1.f, "pw_wgrad: beta must be 0 (overwrite) or 1 (accumulate), got %f", p->beta);
  DLB_REQUIRE((reinterpret_cast<uintptr_t>(p->dW) & 15) == 0, "pw_wgrad: dW must be 16-byte aligned");
  DLB_REQUIRE((p->lda * 2) % 16 == 0 && (p->ldy * 2) % 16 == 0, "pw_wgrad: lda/ldy must be 16-byte multiples");
  WgArgs g{};
  g.M = p->M; g.N = p->N; g.K = p->K;
  const int npad64 = (p->N + 63) / 64 * 64;
  g.n_groups = (npad64 + 511) / 512;
  const int group_cols = ((npad64 / 64 + g.n_groups - 1) / g.n_groups) * 64;   // <= 512, multiple of 64
  g.n_boxes = group_cols / 64;
  // UMMA N chunks: multiples of 64 (a chunk starts on a 64-column box), <= 256; uneven splits take a narrower last
  // chunk (320 = 192 + 128) instead of falling back to five 64-wide MMAs
  g.group_cols = group_cols;
  g.n_chunks = (group_cols + 255) / 256;
  g.chunk_n = ((g.n_boxes + g.n_chunks - 1) / g.n_chunks) * 64;
  g.chunk_last = group_cols - (g.n_chunks - 1) * g.chunk_n;
  // 64 reduction rows per stage whenever >= 3 stages of that size fit
  g.rows_per_stage = (2 + g.n_boxes) * 64 * 128 * 3 <= 200 * 1024 ? 64 : 32;
  g.k_tiles = (p->K + 127) / 128;
  const int tiles = g.k_tiles * g.n_groups;
  // one CTA per SM and ONE wave: tiles * splits must not exceed the SM count (ceil() gave 152 CTAs for 8 K tiles --
  // four stragglers in a second wave doubled the kernel time: 30 % tensor-pipe activity at 1.9 TB/s in the ncu capture)
  int splits = num_sms() / tiles;
  if (splits < 1) splits = 1;
  const int R = g.rows_per_stage;
  int rps = ((p->M + splits - 1) / splits + R - 1) / R * R;
  splits = (p->M + rps - 1) / rps;
  g.splits = splits; g.rows_per_split = rps;
  g.dW = p->dW;
  g.a_scale = p->a_scale; g.a_shift = p->a_shift; g.a_act = p->a_act;
  DLB_REQUIRE(!p->a_scale || p->a_shift, "pw_wgrad: a_scale without a_shift");
  if (p->beta == 0.f) DLB_CUDA(cudaMemsetAsync(p->dW, 0, sizeof(float) * p->K * p->N, st));
  g.a_bytes = 2 * R * 128;
  g.stage_bytes = g.a_bytes + g.n_boxes * R * 128;
  g.num_stages = (200 * 1024) / g.stage_bytes;
  if (g.num_stages > kWgMaxStages) g.num_stages = kWgMaxStages;
  g.idesc = make_idesc(p->dtype == DLB_BF16 ? 1 : 0, 128, g.chunk_n, 1, 1);
  g.idesc_last = make_idesc(p->dtype == DLB_BF16 ? 1 : 0, 128, g.chunk_last, 1, 1);
  CUtensorMap ta, ty;
  int rc = make_tmap_2d(&ta, p->dtype, p->A, p->M, p->K, p->lda, R, 64);
  if (rc) return rc;
  rc = make_tmap_2d(&ty, p->dtype, p->dY, p->M, p->N, p->ldy, R, 64);
  if (rc) return rc;
  const size_t smem_bytes = 1024 + static_cast<size_t>(g.num_stages) * g.stage_bytes + (3 * kWgMaxStages + 2) * 8 + 16 + 256 * 4;
#define LAUNCH(T, X)                                                                                                   \
  do {                                                                                                                 \
    DLB_CUDA(cudaFuncSetAttribute(pw_wgrad_tc_kernel<T, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)); \
    launch_k(pw_wgrad_tc_kernel<T, X>, tiles * splits, kWgThreads, smem_bytes, st, ta, ty, g);                          \
  } while (0)
  if (p->dtype == DLB_BF16) { if (p->a_scale) LAUNCH(__nv_bfloat16, true); else LAUNCH(__nv_bfloat16, false); }
  else { if (p->a_scale) LAUNCH(__half, true); else LAUNCH(__half, false); }
#undef LAUNCH
  g_launches++;
  return check_launch("pw_wgrad_tc_kernel");
}

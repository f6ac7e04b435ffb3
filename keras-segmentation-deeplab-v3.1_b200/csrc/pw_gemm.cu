// Pointwise (1x1) convolution as a GEMM on the Blackwell tensor cores.
//
//   C[M, N] = epilogue(A[M, K] * Bt[N, K]^T)       M = B*H*W (large), K, N = channel counts (16..1344)
//
// Replaces every Conv2D(filters, (1,1)) of the reference graph (deeplabv3p.py:78-82, :175-177, :194-196,
// :385, :406, :420, :438; utils.py:189; subpixel.py:90-91) together with the BatchNorm / ReLU6 / Add /
// bias / phase-shift that follow it, which live in the epilogue.
//
// All of these GEMMs are HBM-bound (arithmetic intensity K*N/(K+N) <= 137 flop/B at 16 bit vs a ridge of
// ~250 flop/B), so the design goal is: stream A exactly once from HBM, keep the (tiny) weights L2-resident,
// and never let the tensor pipe or the epilogue stall the stream.
//
//   * persistent CTAs (one per SM), static tile striding over M tiles of 128 rows
//   * warp 0   : TMA producer  (cp.async.bulk.tensor 2D, 128-byte swizzle, OOB zero fill pads K and N)
//   * warp 1   : tcgen05.mma issuer (one thread), accumulators in TMEM (up to 512 fp32 columns,
//                double-buffered when a column group is <= 256 wide), tcgen05.commit -> mbarriers
//   * warps 2-9: epilogue (two warps per TMEM lane quarter, 32 columns per iteration), tcgen05.ld 32x32b ->
//                registers -> fused BN-affine / bias / per-image bias / activation / residual / BatchNorm batch
//                statistics / Subpixel phase-shift store
//   * N > 512 (expand convs, Subpixel) is processed in column groups; A tiles are re-fetched through L2.
//
// DLB_F32 inputs take an exact-fp32 SIMT path (the 1e-3 parity mode of BASELINE.json; tf32 would not hold it).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "common.cuh"

namespace dlb {

extern std::atomic<long long> g_launches;

constexpr int kBlockM = 128;
constexpr int kSwzBytes = 128;
constexpr int kABytes = kBlockM * kSwzBytes;   // 16 KB per A stage
constexpr int kMaxStages = 8;
constexpr int kMaxAccStages = 4;
constexpr int kStageTileBytes = 4096;          // per epilogue warp: 32 rows x 128 B (64 x 16-bit columns)
constexpr int kMaxSmem = 227 * 1024;

struct GemmArgs {
  int M, N, K;
  int n_store, ldc, ldr;
  void* C;
  const void* R;
  const float* col_scale;
  const float* col_shift;
  const float* row_bias;
  int rows_per_img, ld_row_bias;
  int act;
  double* stat_sum;
  double* stat_sqs;
  int shuffle_r, shuffle_h, shuffle_w, shuffle_cs;
  int chunk_n, chunks_per_group, n_groups, n_chunks;
  int num_k_blocks, num_m_tiles, num_stages, acc_stages, acc_cols;
  int k_elems_per_block, umma_k;   // 64/16 for 16-bit inputs
  uint32_t idesc;
  uint32_t stage_bytes;
  int alt_tiles;   // narrow outputs: epilogue warp set s owns accumulator stage s and drains whole M tiles alone
  const float* a_scale;   // A-operand transform (kXform): A := act(A * a_scale[k] + a_shift[k]) applied to the landed tile
  const float* a_shift;
  int a_act;
  int has_afin; dlb_bn_fin afin;   // kXform: a_scale / a_shift are finalised in the prologue from the batch statistics
  int b_lo_given;  // tf32x3: the caller supplies the weights already split (Bt = hi part, Bt_lo = remainder)
  int tf32x3;      // fp32 operands on the tensor cores: A and B tiles are split in shared memory into a tf32-exact high part
                   // and the remainder, three kind::tf32 MMAs per k-step (hi*lo + lo*hi + hi*hi) ~ fp32 accuracy
  uint32_t stg_bytes;     // epilogue staging tiles (0 for outputs that are stored straight from registers)
  int b_resident;  // the CTA's weight slice [acc_cols, K] stays in shared memory for the whole kernel (b_res_bytes), the
  uint32_t b_res_bytes;   // pipeline stages then carry the A tile only
  int tma_store;   // 16-bit row-major output without residual: staged 32x64 tiles leave through cp.async.bulk.tensor
  // CTA -> column group: group gi is served by CTAs [grp_cta0[gi], grp_cta0[gi] + grp_ctas[gi]) which stride over the M
  // tiles.  A CTA never changes its group, so an epilogue warp always owns the same output columns and keeps their
  // BatchNorm statistics in registers for the whole kernel.
  int grp_cta0[8], grp_ctas[8];
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n)); }

// explicit shared-space accesses on 32-bit addresses (run-time selected tile pointers otherwise decay to generic LD/ST)
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(saddr));
  return u;
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t u;
  asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(u) : "r"(saddr));
  return u;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_saddr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}\n" ::"r"(bar_saddr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
template <typename OutT> __device__ __forceinline__ float2 word_to_float2(uint32_t w);
template <> __device__ __forceinline__ float2 word_to_float2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <> __device__ __forceinline__ float2 word_to_float2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 word_to_float2<float>(uint32_t w) { return make_float2(0.f, 0.f); }

// kSets epilogue warp sets of 4 warps (one warp per TMEM lane quarter).  kSets == 4 is the 16-bit staged-output
// instance (host guarantees: 16-bit OutT, no phase-shift store); kSets == 2 also carries the fp32 / phase-shift stores.
// kLean: no BN-affine and no per-image bias (every training-step GEMM but concat_projection) -- the drain is then
// tcgen05.ld -> pack -> st.shared with no option tests.
// kTma: the staged tile leaves through one cp.async.bulk.tensor store per 32 x 64 block (host guarantees: kSets == 4, no
// residual, no post-statistics activation); the manual coalesced write-back and its registers are compiled out.
// kXform: the last warp set does not drain accumulators but applies BatchNorm-affine + activation to every landed A tile
// (xform_row_sw128) before the MMA issuer may read it: the project conv consumes the raw depthwise output.
template <typename OutT, int kSets, bool kLean, bool kTma, bool kXform>
__global__ void __launch_bounds__(64 + 128 * kSets, 1)
pw_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_bl,
                  const GemmArgs g) {
  constexpr int kThreads = 64 + 128 * kSets;
  constexpr int kEpiSets = kXform ? kSets - 1 : kSets;
  constexpr int kEpiThreads = 128 * kEpiSets;
  constexpr bool kStagedOnly = kSets == 4;
  constexpr bool kCanStage = sizeof(OutT) == 2;
  extern __shared__ uint8_t smem_dyn[];
  // round up to the 1024-byte swizzle-atom alignment by OFFSETTING the shared array (a uintptr_t round trip makes
  // every later access a generic LD/ST instead of LDS/STS)
  uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tail: per-warp staging tiles first (stage_bytes is a multiple of 2 KB, so they stay 1024-byte aligned: the swizzle of
  // the TMA store is a function of the shared-memory address), then barriers and per-column tables
  uint8_t* s_bres = smem + static_cast<size_t>(g.num_stages) * g.stage_bytes;      // resident weights (may be empty)
  uint8_t* s_stage = s_bres + g.b_res_bytes;
  uint8_t* tail = s_stage + g.stg_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + kMaxAccStages;
  uint64_t* bres_bar = tempty_bar + kMaxAccStages;
  uint64_t* xf_bar = bres_bar + 2;          // (+2: the float4 tables below stay 16-byte aligned)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xf_bar + kMaxStages);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);
  const int npad = g.n_chunks * g.chunk_n;
  float* s_shift = s_scale + npad;
  float* s_sum = s_shift + npad;
  float* s_sqs = s_sum + npad;
  float* s_asc = s_sqs + npad;                              // kXform: per-input-channel scale / shift, padded to k-blocks
  float* s_ash = s_asc + g.num_k_blocks * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (kTma) tma_prefetch_desc(&tmap_c);
    if (g.b_lo_given) tma_prefetch_desc(&tmap_bl);
    for (int i = 0; i < g.num_stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < kMaxAccStages; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], g.alt_tiles ? 4 : 4 * kEpiSets); }
    mbar_init(bres_bar, 1);
    if (kXform) for (int i = 0; i < g.num_stages; ++i) mbar_init(&xf_bar[i], 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  pdl_wait();     // everything above is independent of the preceding kernel's output
  for (int i = threadIdx.x; i < npad; i += kThreads) {
    s_scale[i] = (g.col_scale && i < g.N) ? g.col_scale[i] : 1.f;
    s_shift[i] = (g.col_shift && i < g.N) ? g.col_shift[i] : 0.f;
    s_sum[i] = 0.f;
    s_sqs[i] = 0.f;
  }
  if (kXform && sizeof(OutT) == 2)
    for (int i = threadIdx.x; i < g.num_k_blocks * 64; i += kThreads) {
      float sc = 0.f, sh = 0.f;                            // channels past K are TMA zero fill and must stay zero
      if (i < g.K) {
        if (g.has_afin) bn_fin_channel(g.afin, i, blockIdx.x == 0, sc, sh);    // CTA 0 publishes for the backward pass
        else { sc = g.a_scale[i]; sh = g.a_shift[i]; }
      }
      s_asc[i] = sc;
      s_ash[i] = sh;
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int group_rows = g.chunks_per_group * g.chunk_n;
  // this CTA's column group and its share of the M tiles (constant indices: the table lives in the parameter space)
  int grp = 0, cta0 = 0, tstep = g.grp_ctas[0];
#pragma unroll
  for (int gi = 1; gi < 8; ++gi)
    if (gi < g.n_groups && static_cast<int>(blockIdx.x) >= g.grp_cta0[gi]) { grp = gi; cta0 = g.grp_cta0[gi]; tstep = g.grp_ctas[gi]; }
  const int tile0 = static_cast<int>(blockIdx.x) - cta0;
  const int chunks = min(g.chunks_per_group, g.n_chunks - grp * g.chunks_per_group);

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0; uint32_t phase = 0;
      if (g.b_resident && tile0 < g.num_m_tiles) {
        // weight residency: this CTA's [acc_cols, K] slice is fetched once (k-block major, like a pipeline stage)
        mbar_expect_tx(bres_bar, g.num_k_blocks * chunks * g.chunk_n * kSwzBytes);
        for (int kb = 0; kb < g.num_k_blocks; ++kb)
          for (int c = 0; c < chunks; ++c)
            tma_load_2d(s_bres + (kb * g.acc_cols + c * g.chunk_n) * kSwzBytes, &tmap_b, bres_bar,
                        kb * g.k_elems_per_block, (grp * g.chunks_per_group + c) * g.chunk_n);
      }
      for (int tile = tile0; tile < g.num_m_tiles; tile += tstep) {
        const int m0 = tile * kBlockM;
        for (int kb = 0; kb < g.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + static_cast<size_t>(stage) * g.stage_bytes;
          uint8_t* sb = sa + (g.tf32x3 ? 2 * kABytes : kABytes);
          const uint32_t b_bytes = g.b_resident ? 0 : chunks * g.chunk_n * kSwzBytes;
          mbar_expect_tx(&full_bar[stage], kABytes + (g.b_lo_given ? 2 * b_bytes : b_bytes));
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * g.k_elems_per_block, m0);
          if (!g.b_resident)
            for (int c = 0; c < chunks; ++c) {
              tma_load_2d(sb + c * g.chunk_n * kSwzBytes, &tmap_b, &full_bar[stage], kb * g.k_elems_per_block,
                          (grp * g.chunks_per_group + c) * g.chunk_n);
              if (g.b_lo_given)     // pre-split weights: the remainder goes straight into the "B lo" half of the stage
                tma_load_2d(sb + (g.acc_cols + c * g.chunk_n) * kSwzBytes, &tmap_bl, &full_bar[stage],
                            kb * g.k_elems_per_block, (grp * g.chunks_per_group + c) * g.chunk_n);
            }
          if (++stage == g.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      if (g.b_resident && tile0 < g.num_m_tiles) mbar_wait(bres_bar, 0);
      for (int tile = tile0; tile < g.num_m_tiles; tile += tstep) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < g.num_k_blocks; ++kb) {
          mbar_wait(kXform ? &xf_bar[stage] : &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * g.stage_bytes);
          const uint32_t sb = g.b_resident ? smem_u32(s_bres) + kb * g.acc_cols * kSwzBytes : sa + kABytes;
          const int k_left = g.K - kb * g.k_elems_per_block;
          const int ksteps = min(g.k_elems_per_block / g.umma_k, (k_left + g.umma_k - 1) / g.umma_k);
          if constexpr (kXform && sizeof(OutT) == 4) {
            // 3xTF32: stage = [A hi 16 KB][A lo 16 KB][B hi acc_cols x 128 B][B lo ...]; small terms first
            const uint32_t sal = sa + kABytes, sbh = sa + 2 * kABytes, sbl = sbh + g.acc_cols * kSwzBytes;
            for (int c = 0; c < chunks; ++c) {
              const uint32_t d_tmem = tmem_base + as * g.acc_cols + c * g.chunk_n;
              for (int k = 0; k < ksteps; ++k) {
                const uint32_t bo = c * g.chunk_n * kSwzBytes + k * 32;
                const uint64_t ah = make_sw128_desc(sa + k * 32, 16, 1024), al = make_sw128_desc(sal + k * 32, 16, 1024);
                const uint64_t bh = make_sw128_desc(sbh + bo, 16, 1024), bl = make_sw128_desc(sbl + bo, 16, 1024);
                umma_tf32(d_tmem, ah, bl, g.idesc, (kb > 0 || k > 0) ? 1u : 0u);
                umma_tf32(d_tmem, al, bh, g.idesc, 1u);
                umma_tf32(d_tmem, ah, bh, g.idesc, 1u);
              }
            }
          } else {
          for (int c = 0; c < chunks; ++c) {
            const uint32_t d_tmem = tmem_base + as * g.acc_cols + c * g.chunk_n;
            for (int k = 0; k < ksteps; ++k) {
              // +32 B per UMMA_K step inside the 128-byte swizzle atom
              const uint64_t adesc = make_sw128_desc(sa + k * 32, 16, 1024);
              const uint64_t bdesc = make_sw128_desc(sb + c * g.chunk_n * kSwzBytes + k * 32, 16, 1024);
              umma_f16(d_tmem, adesc, bdesc, g.idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          }
          umma_commit(&empty_bar[stage]);     // frees this smem stage when the MMAs retire
          if (++stage == g.num_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);          // accumulator group complete -> epilogue
        if (++as == g.acc_stages) { as = 0; aphase ^= 1; }
      }
    }
  } else if (kXform && warp >= 2 + 4 * kEpiSets) {
    // ===================== A-operand transform warps (the last warp set) =====================
    // thread = rows r and r + 64 of the landed 128-row A tile, one half of their 16-byte chunks; waits for the TMA bytes,
    // rewrites in place, makes the writes visible to the tensor-core (async) proxy, lane 0 arrives for the MMA issuer
    const int t = threadIdx.x - (64 + kEpiThreads);
    const int r = t & 63, c0 = (t >> 6) * 4;
    const uint32_t asc = smem_u32(s_asc), ash = smem_u32(s_ash);
    int stage = 0; uint32_t phase = 0;
    for (int tile = tile0; tile < g.num_m_tiles; tile += tstep) {
      for (int kb = 0; kb < g.num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * g.stage_bytes);
        if constexpr (sizeof(OutT) == 4) {
          // 3xTF32 operand split: x = hi + lo with hi = x with the 13 low mantissa bits cleared (exact in tf32, so the
          // tensor core's own fp32 -> tf32 conversion cannot change it) and lo = x - hi (exact in fp32).  Rows of the
          // A tile and of the weight tile are 128 B = 8 chunks; the 16-byte pieces keep their (swizzled) position.
          const uint32_t sbh = sa + 2 * kABytes;
          const int b_rows = g.b_lo_given ? 0 : chunks * g.chunk_n;
          for (int row = t; row < kBlockM + b_rows; row += 128) {
            const uint32_t hi = row < kBlockM ? sa + row * 128 : sbh + (row - kBlockM) * 128;
            const uint32_t lo = row < kBlockM ? hi + kABytes : hi + g.acc_cols * kSwzBytes;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t off = static_cast<uint32_t>(i ^ (row & 7)) << 4;
              uint4 u = lds128_u32(hi + off), l;
              const uint4 h = make_uint4(u.x & 0xffffe000u, u.y & 0xffffe000u, u.z & 0xffffe000u, u.w & 0xffffe000u);
              l.x = __float_as_uint(__uint_as_float(u.x) - __uint_as_float(h.x));
              l.y = __float_as_uint(__uint_as_float(u.y) - __uint_as_float(h.y));
              l.z = __float_as_uint(__uint_as_float(u.z) - __uint_as_float(h.z));
              l.w = __float_as_uint(__uint_as_float(u.w) - __uint_as_float(h.w));
              sts128_u32(hi + off, h);
              sts128_u32(lo + off, l);
            }
          }
        } else {
          xform_rows2<OutT, 4>(sa, r, 64, c0, asc + kb * 256, ash + kb * 256, g.a_act);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_bar[stage]);
        if (++stage == g.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2 .. 2 + 4*kEpiSets) =====================
    // A warp set = 4 warps, one per TMEM lane quarter.  Wide outputs: all sets drain the same accumulator group and
    // take every kSets-th 64-column block.  Narrow outputs (alt_tiles): set s owns accumulator stage s, i.e. every
    // acc_stages-th M tile, and drains all of its blocks.
    //  * thread = accumulator row: tcgen05.ld (2 x 16 columns in flight) -> BN-affine / per-image bias
    //  * 16-bit outputs are staged through a swizzled 32 x 64 shared-memory tile per warp; BatchNorm statistics are
    //    taken from the staged (rounded) tile with a lane = 2 adjacent columns walk on packed fp32 pairs, and the tile
    //    is written back with 4 rows x 128 B per instruction (a thread-per-row store costs 32 L1 wavefronts per 512 B);
    //    residual and activation are applied in that coalesced phase.
    const int quad = warp & 3;                  // TMEM lane quarter this warp may access
    const int set = (warp - 2) >> 2;
    const bool do_stats = g.stat_sum != nullptr;
    const bool affine = !kLean && (g.col_scale != nullptr || g.col_shift != nullptr);
    const bool staged = kStagedOnly || (kCanStage && g.shuffle_r == 0);
    const uint32_t stg = smem_u32(s_stage) + static_cast<uint32_t>(warp - 2) * kStageTileBytes;   // [32 rows][8 chunks of 16 B]
    // drain: chunk q of row `lane` -> slot q ^ (lane & 7)   (= the SWIZZLE_128B pattern of the output tensor map)
    const uint32_t stg_row = stg + lane * 128;
    const uint32_t lane7 = lane & 7;
    // statistics walk: lane owns word `lane` (columns 2*lane, 2*lane+1) of every row; rows r and r + 8k share a swizzle
    const uint32_t st_chunk = lane >> 2, st_word = (lane & 3) * 4;
    // write-back: 8 lanes cover the 128 B of one row, 4 rows per instruction; rows i*4 + rsub alternate two swizzles
    const int cchunk = lane & 7, rsub = lane >> 3;
    const uint32_t wb0 = stg + rsub * 128 + ((cchunk ^ rsub) << 4);
    const uint32_t wb1 = stg + (rsub + 4) * 128 + ((cchunk ^ (rsub + 4)) << 4);
    // without statistics the activation is applied to the accumulators in registers, so the staged tile is final
    const bool act_in_drain = staged && !do_stats && g.act != DLB_ACT_NONE;
    const bool wb_plain = g.R == nullptr && (g.act == DLB_ACT_NONE || act_in_drain);
    constexpr bool use_tma = kTma;
    const int red_col = reduce16_col_of_lane(lane);
    const uint32_t tfull_a = smem_u32(tfull_bar);
    const int j_first = g.alt_tiles ? 0 : set * 64;
    const int j_step = g.alt_tiles ? 64 : 64 * kEpiSets;
    const int col_base = grp * group_rows;
    // valid accumulator columns of this group (the last chunk of a 64-aligned split may be partly padding)
    const int gcols = min(chunks * g.chunk_n, ((g.N + 15) & ~15) - col_base);
    // BatchNorm statistics of this warp's (up to three) 64-column blocks, two packed fp32 columns per lane, kept in
    // registers across all M tiles of the CTA
    unsigned long long S0 = 0ull, Q0 = 0ull, S1 = 0ull, Q1 = 0ull, S2 = 0ull, Q2 = 0ull;
    int as = 0; uint32_t aphase = 0;
    for (int tile = tile0; tile < g.num_m_tiles; tile += tstep) {
      const int m_base = tile * kBlockM + quad * 32;
      const int m = m_base + lane;
      const bool row_ok = m < g.M;
      const int rows_left = g.M - m_base;
      const float* rb = nullptr;
      if (!kLean && g.row_bias && row_ok) rb = g.row_bias + static_cast<size_t>(m / g.rows_per_img) * g.ld_row_bias;
      size_t shuf_row_base = 0;
      if (!kStagedOnly && g.shuffle_r > 0 && row_ok) {
        const int hw = g.shuffle_h * g.shuffle_w;
        const int b = m / hw, rem = m - b * hw;
        const int a = rem / g.shuffle_w, bb = rem - a * g.shuffle_w;
        shuf_row_base = ((static_cast<size_t>(b) * g.shuffle_h * g.shuffle_r + static_cast<size_t>(a) * g.shuffle_r) *
                             (static_cast<size_t>(g.shuffle_w) * g.shuffle_r) +
                         static_cast<size_t>(bb) * g.shuffle_r) *
                        g.shuffle_cs;
      }
      const int as_cur = as;
      const uint32_t aphase_cur = aphase;
      if (++as == g.acc_stages) { as = 0; aphase ^= 1; }
      if (g.alt_tiles && as_cur != set) continue;     // narrow outputs: this warp set owns accumulator stage `set`
      mbar_wait_s(tfull_a + as_cur * 8, aphase_cur);
      tc_fence_after();
      int bi = 0;
      for (int j64 = j_first; j64 < gcols; j64 += j_step, ++bi) {
        if (use_tma) {
          // the previous block's bulk store must have finished READING the staging tile before it is refilled
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int j = j64 + sub * 32;
          if (j >= gcols) break;
          const bool two = j + 16 < gcols;
          uint32_t r[2][16];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as_cur * g.acc_cols + j;
          tmem_ld16(taddr, r[0]);
          if (two) tmem_ld16(taddr + 16, r[1]);
          tmem_ld_wait();
          float v[2][16];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[h][i] = __uint_as_float(r[h][i]);
          // kernel-uniform options are tested once per 32 columns, not per element
          if (!kLean && affine) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4* sc4 = reinterpret_cast<const float4*>(s_scale + col_base + j + h * 16);
              const float4* sh4 = reinterpret_cast<const float4*>(s_shift + col_base + j + h * 16);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 a4 = sc4[q], b4 = sh4[q];
                v[h][4 * q + 0] = fmaf(v[h][4 * q + 0], a4.x, b4.x);
                v[h][4 * q + 1] = fmaf(v[h][4 * q + 1], a4.y, b4.y);
                v[h][4 * q + 2] = fmaf(v[h][4 * q + 2], a4.z, b4.z);
                v[h][4 * q + 3] = fmaf(v[h][4 * q + 3], a4.w, b4.w);
              }
            }
          }
          if (!kLean && rb) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int n0 = col_base + j + h * 16;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (n0 + i < g.N) v[h][i] += rb[n0 + i];
            }
          }
          if (!two || (do_stats && !row_ok)) {
            const bool all = do_stats && !row_ok;       // rows past M must not reach the statistics
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[1][i] = 0.f; if (all) v[0][i] = 0.f; }
          }
          if (staged) {
            if (act_in_drain) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 16; ++i) v[h][i] = apply_act(v[h][i], g.act);
            }
            // pack to 16 bit and park in the swizzled staging tile
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = v[c >> 1][(c & 1) * 8 + i];
              uint4 pk;
              Vec8<OutT>::st(reinterpret_cast<OutT*>(&pk), o);
              sts128(stg_row + (((sub * 4 + c) ^ lane7) << 4), pk);
            }
          }
          if constexpr (!kStagedOnly) {
            if (!staged) {
              if (do_stats) {
                float s1[2][16], s2[2][16];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const float q = v[h][i];
                    s1[h][i] = q; s2[h][i] = q * q;
                  }
                const float t1a = warp_reduce16(s1[0], lane), t2a = warp_reduce16(s2[0], lane);
                const float t1b = warp_reduce16(s1[1], lane), t2b = warp_reduce16(s2[1], lane);
                if ((lane & 1) == 0) {
                  const int n0 = col_base + j;
                  atomicAdd(&s_sum[n0 + red_col], t1a);
                  atomicAdd(&s_sqs[n0 + red_col], t2a);
                  if (two) {
                    atomicAdd(&s_sum[n0 + 16 + red_col], t1b);
                    atomicAdd(&s_sqs[n0 + 16 + red_col], t2b);
                  }
                }
              }
              if (row_ok) {
#pragma unroll
                for (int h8 = 0; h8 < 4; ++h8) {
                  const int n = col_base + j + h8 * 8;
                  if (n >= g.n_store || (h8 >= 2 && !two)) continue;
                  float o[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) o[i] = apply_act(v[h8 >> 1][(h8 & 1) * 8 + i], g.act);
                  if (g.R) {
                    float rr[8];
                    Vec8<OutT>::ld(reinterpret_cast<const OutT*>(g.R) + static_cast<size_t>(m) * g.ldr + n, rr);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] += rr[i];
                  }
                  OutT* dst;
                  if (g.shuffle_r > 0) {
                    const int rowlen = g.shuffle_r * g.shuffle_cs;       // (i, k) run contiguous in the output
                    const int jj = n / rowlen, rem = n - jj * rowlen;
                    dst = reinterpret_cast<OutT*>(g.C) + shuf_row_base +
                          static_cast<size_t>(jj) * (static_cast<size_t>(g.shuffle_w) * g.shuffle_r) * g.shuffle_cs + rem;
                  } else {
                    dst = reinterpret_cast<OutT*>(g.C) + static_cast<size_t>(m) * g.ldc + n;
                  }
                  if (sizeof(OutT) == 4 && n + 8 > g.n_store) {
                    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);   // fp32 rows may end on a multiple of 4
                  } else {
                    Vec8<OutT>::st(dst, o);
                  }
                }
              }
            }
          }
        }
        if (staged) {
          if constexpr (kCanStage) {
            if (use_tma) fence_proxy_async();       // this lane's st.shared -> visible to the bulk-copy engine
            __syncwarp();
            if (use_tma && lane == 0) {
              // one instruction writes the 32 x 64 tile; rows past M and columns past n_store are clipped by the map
              tma_store_2d(&tmap_c, stg, col_base + j64, m_base);
              bulk_commit();
            }
            if (do_stats) {
              // BatchNorm statistics of the (rounded, pre-activation) tile: conflict-free, a warp reads one
              // 128-byte row per step
              unsigned long long acc_s = 0ull, acc_q = 0ull, acc_s2 = 0ull, acc_q2 = 0ull;   // two chains per statistic
#pragma unroll
              for (int r8 = 0; r8 < 8; ++r8) {
                const uint32_t a0 = stg + r8 * 128 + (((st_chunk ^ r8) << 4) | st_word);
#pragma unroll
                for (int k = 0; k < 4; k += 2) {
                  const float2 f = word_to_float2<OutT>(lds32(a0 + k * 1024));
                  const float2 h = word_to_float2<OutT>(lds32(a0 + (k + 1) * 1024));
                  f32x2_acc(acc_s, acc_q, f.x, f.y);
                  f32x2_acc(acc_s2, acc_q2, h.x, h.y);
                }
              }
              acc_s = f32x2_add(acc_s, acc_s2);
              acc_q = f32x2_add(acc_q, acc_q2);
              if constexpr (kTma) {
                if (bi == 0) { S0 = f32x2_add(S0, acc_s); Q0 = f32x2_add(Q0, acc_q); }
                else if (bi == 1) { S1 = f32x2_add(S1, acc_s); Q1 = f32x2_add(Q1, acc_q); }
                else { S2 = f32x2_add(S2, acc_s); Q2 = f32x2_add(Q2, acc_q); }
              } else {
                // fallback instance (unaligned output / DLB_GEMM_TMA_STORE=0): per-block shared-memory atomics
                const int jc = j64 + 2 * lane;
                if (jc < gcols) {
                  const float2 sv = f32x2_unpack(acc_s), qv = f32x2_unpack(acc_q);
                  atomicAdd(&s_sum[col_base + jc], sv.x); atomicAdd(&s_sum[col_base + jc + 1], sv.y);
                  atomicAdd(&s_sqs[col_base + jc], qv.x); atomicAdd(&s_sqs[col_base + jc + 1], qv.y);
                }
              }
            }
            if constexpr (!use_tma) {
              // coalesced write-back
              const int n = col_base + j64 + cchunk * 8;
              if (n < g.n_store && j64 + cchunk * 8 < gcols) {
                OutT* dst = reinterpret_cast<OutT*>(g.C) + static_cast<size_t>(m_base + rsub) * g.ldc + n;
                const size_t step = static_cast<size_t>(4) * g.ldc;
                if (wb_plain) {
#pragma unroll
                  for (int i0 = 0; i0 < 8; i0 += 4) {
                    uint4 pk[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) pk[i] = lds128((((i0 + i) & 1) ? wb1 : wb0) + ((i0 + i) >> 1) * 1024);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                      if ((i0 + i) * 4 + rsub < rows_left) *reinterpret_cast<uint4*>(dst + (i0 + i) * step) = pk[i];
                  }
                } else {
                  // activation, then the residual (loaded 4 rows ahead: the thread-per-vector loads are latency bound)
                  const OutT* res = g.R ? reinterpret_cast<const OutT*>(g.R) + static_cast<size_t>(m_base + rsub) * g.ldr + n : nullptr;
                  const size_t rstep = static_cast<size_t>(4) * g.ldr;
#pragma unroll
                  for (int i0 = 0; i0 < 8; i0 += 4) {
                    uint4 rr[4];
                    if (res) {
#pragma unroll
                      for (int i = 0; i < 4; ++i)
                        rr[i] = ((i0 + i) * 4 + rsub < rows_left) ? *reinterpret_cast<const uint4*>(res + (i0 + i) * rstep)
                                                                  : make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      const int ii = i0 + i;
                      const uint4 pk = lds128(((ii & 1) ? wb1 : wb0) + (ii >> 1) * 1024);
                      float o[8];
                      Vec8<OutT>::ld(reinterpret_cast<const OutT*>(&pk), o);
                      if (!act_in_drain) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) o[q] = apply_act(o[q], g.act);
                      }
                      if (res) {
                        float rf[8];
                        Vec8<OutT>::ld(reinterpret_cast<const OutT*>(&rr[i]), rf);
#pragma unroll
                        for (int q = 0; q < 8; ++q) o[q] += rf[q];
                      }
                      if (ii * 4 + rsub < rows_left) Vec8<OutT>::st(dst + ii * step, o);
                    }
                  }
                }
              }
              __syncwarp();
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as_cur]);
    }
    if (use_tma && lane == 0) bulk_wait_read0();     // the staging tile must outlive the last bulk store's read
    if (do_stats) {
      if constexpr (kTma) {
        // register-resident partial sums -> the CTA's per-column table (once per kernel), then fp64 to global
#pragma unroll
        for (int b3 = 0; b3 < 3; ++b3) {
          const int jc = j_first + b3 * j_step + 2 * lane;
          if (jc < gcols) {
            const float2 sv = f32x2_unpack(b3 == 0 ? S0 : (b3 == 1 ? S1 : S2));
            const float2 qv = f32x2_unpack(b3 == 0 ? Q0 : (b3 == 1 ? Q1 : Q2));
            atomicAdd(&s_sum[col_base + jc], sv.x); atomicAdd(&s_sum[col_base + jc + 1], sv.y);
            atomicAdd(&s_sqs[col_base + jc], qv.x); atomicAdd(&s_sqs[col_base + jc + 1], qv.y);
          }
        }
      }
      named_bar_sync(1, kEpiThreads);
      for (int n = col_base + threadIdx.x - 64; n < min(col_base + gcols, g.N); n += kEpiThreads) {
        atomicAdd(&g.stat_sum[n], static_cast<double>(s_sum[n]));
        atomicAdd(&g.stat_sqs[n], static_cast<double>(s_sqs[n]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// exact fp32 SIMT path (parity mode)
// ---------------------------------------------------------------------------------------------
struct SimtArgs {
  int M, N, K, lda, ldb, ldc, ldr, n_store;
  const float* A; const float* Bt; float* C; const float* R;
  const float* col_scale; const float* col_shift; const float* row_bias;
  int rows_per_img, ld_row_bias, act;
  double* stat_sum; double* stat_sqs;
  int shuffle_r, shuffle_h, shuffle_w, shuffle_cs;
  const float* a_scale; const float* a_shift; int a_act;
};

__global__ void __launch_bounds__(256) pw_gemm_simt_kernel(const SimtArgs g) {
  pdl_prologue();
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float sA[TK][TM + 4];
  __shared__ float sB[TK][TN + 4];
  __shared__ float s_sum[TN], s_sqs[TN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int tx = tid & 15, ty = tid >> 4;   // thread computes rows ty*4.., cols tx*4..
  if (tid < TN) { s_sum[tid] = 0.f; s_sqs[tid] = 0.f; }
  float acc[4][4] = {};
  const int lr = tid >> 2, lk = (tid & 3) * 4;   // loader: row lr (0..63), k offset lk
  for (int k0 = 0; k0 < g.K; k0 += TK) {
    {
      float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
      const int m = m0 + lr, n = n0 + lr, k = k0 + lk;
      if (m < g.M && k < g.K) {
        a = *reinterpret_cast<const float4*>(g.A + static_cast<size_t>(m) * g.lda + k);
        if (g.a_scale) {       // A-operand transform: BatchNorm affine + activation of the producing layer
          const float4 sc = *reinterpret_cast<const float4*>(g.a_scale + k), sh = *reinterpret_cast<const float4*>(g.a_shift + k);
          a.x = apply_act(fmaf(a.x, sc.x, sh.x), g.a_act); a.y = apply_act(fmaf(a.y, sc.y, sh.y), g.a_act);
          a.z = apply_act(fmaf(a.z, sc.z, sh.z), g.a_act); a.w = apply_act(fmaf(a.w, sc.w, sh.w), g.a_act);
        }
      }
      if (n < g.N && k < g.K) b = *reinterpret_cast<const float4*>(g.Bt + static_cast<size_t>(n) * g.ldb + k);
      sA[lk + 0][lr] = a.x; sA[lk + 1][lr] = a.y; sA[lk + 2][lr] = a.z; sA[lk + 3][lr] = a.w;
      sB[lk + 0][lr] = b.x; sB[lk + 1][lr] = b.y; sB[lk + 2][lr] = b.z; sB[lk + 3][lr] = b.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float cs[4] = {0, 0, 0, 0}, cq[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N && n >= g.n_store) continue;
      float v = acc[i][j];
      if (n < g.N) {
        if (g.col_scale) v *= g.col_scale[n];
        if (g.col_shift) v += g.col_shift[n];
        if (g.row_bias) v += g.row_bias[static_cast<size_t>(m / g.rows_per_img) * g.ld_row_bias + n];
        cs[j] += v; cq[j] += v * v;
      } else {
        v = 0.f;
      }
      v = apply_act(v, g.act);
      if (n < g.n_store) {
        if (g.R) v += g.R[static_cast<size_t>(m) * g.ldr + n];
        size_t off;
        if (g.shuffle_r > 0) {
          const int hw = g.shuffle_h * g.shuffle_w;
          const int b = m / hw, rem = m - b * hw;
          const int a = rem / g.shuffle_w, bb = rem - a * g.shuffle_w;
          const int rowlen = g.shuffle_r * g.shuffle_cs;
          const int jj = n / rowlen, r2 = n - jj * rowlen;
          off = ((static_cast<size_t>(b) * g.shuffle_h * g.shuffle_r + static_cast<size_t>(a) * g.shuffle_r + jj) *
                     (static_cast<size_t>(g.shuffle_w) * g.shuffle_r) +
                 static_cast<size_t>(bb) * g.shuffle_r) * g.shuffle_cs + r2;
        } else {
          off = static_cast<size_t>(m) * g.ldc + n;
        }
        g.C[off] = v;
      }
    }
  }
  if (g.stat_sum) {
    for (int j = 0; j < 4; ++j) { atomicAdd(&s_sum[tx * 4 + j], cs[j]); atomicAdd(&s_sqs[tx * 4 + j], cq[j]); }
    __syncthreads();
    if (tid < TN && n0 + tid < g.N) {
      atomicAdd(&g.stat_sum[n0 + tid], static_cast<double>(s_sum[tid]));
      atomicAdd(&g.stat_sqs[n0 + tid], static_cast<double>(s_sqs[tid]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// 2D row-major tensor [rows, cols] with row pitch `ld` elements; box = [box_rows, box_cols], 128 B swizzle
int make_tmap_2d(CUtensorMap* map, int dtype, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled driver entry point unavailable"); return DLB_ERR_CUDA; }
  const int es = dtype_size(dtype);
  CUtensorMapDataType dt = dtype == DLB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                         : dtype == DLB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                             : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * es};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u ptr=%p", (int)r,
                   (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols, ptr);
    return DLB_ERR_CUDA;
  }
  return DLB_OK;
}

// NHWC activation tensor [B, H, W, C] as a 4D tensor map (dims innermost first: C, W, H, B), un-swizzled boxes
// [1, box_h, box_w, box_c]; out-of-bounds coordinates (negative or past the edge) are zero-filled = conv padding
int make_tmap_nhwc(CUtensorMap* map, int dtype, const void* ptr, int B, int H, int W, int C, int box_c, int box_w,
                   int box_h, int elem_stride, int nan_fill) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled driver entry point unavailable"); return DLB_ERR_CUDA; }
  const cuuint64_t es = dtype_size(dtype);
  CUtensorMapDataType dt = dtype == DLB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                         : dtype == DLB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                             : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  // spatial element strides > 1 read every elem_stride-th pixel (box = pixels * stride, see cuTensorMapEncodeTiled)
  cuuint32_t estr[4] = {1, (cuuint32_t)elem_stride, (cuuint32_t)elem_stride, 1};
  CUresult r = fn(map, dt, 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(4D) failed (%d): B=%d H=%d W=%d C=%d box=%dx%dx%d", (int)r, B, H, W, C, box_h,
                   box_w, box_c);
    return DLB_ERR_CUDA;
  }
  return DLB_OK;
}

// NHWC activation tensor as a 4D tensor map with 128-byte swizzled boxes [1, box_h, box_w, 64 channels]: a landed box
// is pixel-major rows of 128 B = directly a K-major SWIZZLE_128B tcgen05 operand tile (sepconv_fused.cu); coordinates
// outside the image (and channels past C) are zero-filled.
int make_tmap_nhwc_sw128(CUtensorMap* map, int dtype, const void* ptr, int B, int H, int W, int C, int box_w, int box_h) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled driver entry point unavailable"); return DLB_ERR_CUDA; }
  const cuuint64_t es = 2;
  CUtensorMapDataType dt = dtype == DLB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, dt, 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(4D sw128) failed (%d): B=%d H=%d W=%d C=%d box=%dx%d", (int)r, B, H, W, C, box_h, box_w);
    return DLB_ERR_CUDA;
  }
  return DLB_OK;
}

// Tiling plan of the tensor-core kernel (pure host arithmetic; exported as dlb_pw_gemm_plan for the CPU tests).
// Returns the number of epilogue warp sets (4 or 2), or 0 if no pipeline fits.
static int plan_tc(int N, int K, int M, int out_dtype, int shuffle_r, int xform, int tf32, GemmArgs* gp, size_t* tail_out) {
  GemmArgs& g = *gp;
  // 16-bit row-major outputs take the 4-set (16 epilogue warps) staged instance; fp32 / phase-shift stores the 2-set one
  // fp32 operands (3xTF32): 2 draining sets + the set that splits the landed tiles into hi / lo parts
  const int sets = tf32 ? 3 : ((out_dtype != DLB_F32 && shuffle_r == 0) ? 4 : 2);
  const int esets = sets - ((xform || tf32) ? 1 : 0);      // warp sets that drain accumulators (the last one transforms tiles)
  const int npad = (N + 15) / 16 * 16;
  g.tf32x3 = tf32 ? 1 : 0;
  g.k_elems_per_block = tf32 ? 32 : 64; g.umma_k = tf32 ? 8 : 16;
  g.num_k_blocks = (K + g.k_elems_per_block - 1) / g.k_elems_per_block;
  g.stg_bytes = (out_dtype != DLB_F32 && shuffle_r == 0) ? static_cast<uint32_t>(4 * sets) * kStageTileBytes : 0u;
  g.num_m_tiles = (M + kBlockM - 1) / kBlockM;
  static const int tune_cpg = [] { const char* e = getenv("DLB_GEMM_CPG"); return e ? atoi(e) : 0; }();   // tuning aid
  // Two chunks per accumulator group share one A fetch (matters when K is large); for small K prefer one chunk per
  // group so the group fits twice in TMEM and the epilogue of one group overlaps the MMAs of the next.
  for (int pass = 0; pass < 2; ++pass) {
    g.n_chunks = (npad + 255) / 256;
    g.chunk_n = ((npad + g.n_chunks - 1) / g.n_chunks + 15) / 16 * 16;
    int cpg = (pass == 0 && !tf32 && g.n_chunks >= 2 && K > 256 && 2 * g.chunk_n <= 512 && tune_cpg != 1) ? 2 : 1;
    int n_groups = (g.n_chunks + cpg - 1) / cpg;
    if (n_groups > 1) {
      // several column groups (each served by its own CTAs): group boundaries on multiples of 64 columns, so every
      // 64-column epilogue block lies inside one group (960 = 256 + 256 + 256 + 192; the padding rows of the last
      // weight chunk are zero-filled by TMA and never drained)
      g.chunk_n = (g.chunk_n + 63) / 64 * 64;
      g.n_chunks = (npad + g.chunk_n - 1) / g.chunk_n;
      if (cpg == 2 && 2 * g.chunk_n > 512) cpg = 1;
      n_groups = (g.n_chunks + cpg - 1) / cpg;
    }
    if (n_groups > 8) return 0;
    g.chunks_per_group = cpg;
    g.n_groups = n_groups;
    g.acc_cols = cpg * g.chunk_n;
    g.acc_stages = g.acc_cols <= 256 ? 2 : 1;
    g.alt_tiles = 0;
    if (g.n_groups == 1) {
      // per-tile critical path in 64-column block units: all sets on one group vs one set per accumulator stage
      int alt_stages = 512 / g.acc_cols;
      if (alt_stages > esets) alt_stages = esets;
      const int blocks = (g.acc_cols + 63) / 64;
      const int cost_split = (blocks + esets - 1) / esets * 64;
      if (alt_stages >= 2 && g.acc_cols < cost_split * alt_stages) { g.alt_tiles = 1; g.acc_stages = alt_stages; }
    }
    const size_t tail = g.stg_bytes + (3 * kMaxStages + 2 * kMaxAccStages + 2) * 8 + 16 +
                        4 * static_cast<size_t>(g.n_chunks * g.chunk_n) * 4 + 16 +
                        (xform ? 2 * static_cast<size_t>(g.num_k_blocks) * 64 * 4 : 0);
    *tail_out = tail;
    // Weight residency: a CTA serves one column group for the whole kernel, so its weight slice (k-blocks x acc_cols
    // x 128 B) can be fetched once instead of once per M tile (ncu, round 2: the epilogue warps wait on the
    // accumulator barrier -- the kernel is bound by shared-memory fill traffic, two thirds of it weights coming back
    // from L2 for every tile).  Taken when it leaves >= 3 A-only stages (narrow inputs: A is small and shared through
    // L2 by the CTAs of the other column groups) or >= 5 (K > 256: A streams from HBM and needs bytes in flight).
    static const bool bres_on = [] { const char* e = getenv("DLB_GEMM_BRES"); return !(e && e[0] == '0'); }();
    const size_t bres = static_cast<size_t>(g.num_k_blocks) * g.acc_cols * kSwzBytes;
    g.b_resident = 0; g.b_res_bytes = 0;
    const size_t need_stages = K > 256 ? 5 : 3;
    if (tf32) {
      // stage = A hi + A lo + weight hi + weight lo; the MMA issue rate (3 per k-step) bounds this kernel, 2 stages do
      g.stage_bytes = 2 * (kABytes + g.acc_cols * kSwzBytes);
      if (static_cast<size_t>(kMaxSmem) < 1024 + tail + 2 * static_cast<size_t>(g.stage_bytes)) return 0;
      g.num_stages = static_cast<int>((kMaxSmem - 1024 - tail) / g.stage_bytes);
      if (g.num_stages > kMaxStages) g.num_stages = kMaxStages;
      return sets;
    }
    if (bres_on && static_cast<size_t>(kMaxSmem) >= 1024 + tail + bres + need_stages * static_cast<size_t>(kABytes)) {
      g.b_resident = 1; g.b_res_bytes = static_cast<uint32_t>(bres);
      g.stage_bytes = kABytes;
      g.num_stages = static_cast<int>((kMaxSmem - 1024 - tail - bres) / kABytes);
      if (g.num_stages > kMaxStages) g.num_stages = kMaxStages;
      return sets;
    }
    g.stage_bytes = kABytes + g.acc_cols * kSwzBytes;
    if (static_cast<size_t>(kMaxSmem) < 1024 + tail + 2 * static_cast<size_t>(g.stage_bytes)) {
      if (cpg == 1) return 0;
      continue;          // retry with one chunk per group
    }
    g.num_stages = static_cast<int>((kMaxSmem - 1024 - tail) / g.stage_bytes);
    if (g.num_stages > kMaxStages) g.num_stages = kMaxStages;
    return sets;
  }
  return 0;
}

// CTAs per column group, proportional to the group's valid width (a narrower last group gets fewer CTAs, each of
// which then walks more M tiles: equal epilogue work per CTA).  Returns the grid size.
static int split_ctas(GemmArgs* gp, int N) {
  GemmArgs& g = *gp;
  const int group_cols = g.chunks_per_group * g.chunk_n;
  const int n16 = (N + 15) / 16 * 16;
  int width[8], total_w = 0;
  for (int gi = 0; gi < g.n_groups; ++gi) {
    width[gi] = n16 - gi * group_cols;
    if (width[gi] > group_cols) width[gi] = group_cols;
    total_w += width[gi];
  }
  const int sms = num_sms();
  const long long n_units = static_cast<long long>(g.num_m_tiles) * g.n_groups;
  int grid = n_units < sms ? static_cast<int>(n_units) : sms;
  if (grid < g.n_groups) grid = g.n_groups;
  int used = 0;
  for (int gi = 0; gi < g.n_groups; ++gi) {
    int c = static_cast<int>(static_cast<long long>(grid) * width[gi] / total_w);
    if (c < 1) c = 1;
    if (c > g.num_m_tiles) c = g.num_m_tiles;
    g.grp_ctas[gi] = c;
    used += c;
  }
  // hand the rounding remainder to the groups with the most work per CTA
  while (used < grid) {
    int best = -1; double worst = 0.0;
    for (int gi = 0; gi < g.n_groups; ++gi) {
      if (g.grp_ctas[gi] >= g.num_m_tiles) continue;
      const double load = static_cast<double>(width[gi]) * ((g.num_m_tiles + g.grp_ctas[gi] - 1) / g.grp_ctas[gi]);
      if (load > worst) { worst = load; best = gi; }
    }
    if (best < 0) break;
    g.grp_ctas[best]++; used++;
  }
  while (used > grid) {        // (only when a group was clamped up to 1)
    int best = 0;
    for (int gi = 1; gi < g.n_groups; ++gi) if (g.grp_ctas[gi] > g.grp_ctas[best]) best = gi;
    if (g.grp_ctas[best] <= 1) break;
    g.grp_ctas[best]--; used--;
  }
  int c0 = 0;
  for (int gi = 0; gi < 8; ++gi) {
    if (gi >= g.n_groups) { g.grp_cta0[gi] = 1 << 30; g.grp_ctas[gi] = 1; continue; }
    g.grp_cta0[gi] = c0; c0 += g.grp_ctas[gi];
  }
  return c0;
}

static int launch_tc(const dlb_pw_gemm_params* p, cudaStream_t st) {
  GemmArgs g{};
  g.M = p->M; g.N = p->N; g.K = p->K;
  g.n_store = p->n_store; g.ldc = p->ldc; g.ldr = p->ldr;
  g.C = p->C; g.R = p->R;
  g.col_scale = p->col_scale; g.col_shift = p->col_shift; g.row_bias = p->row_bias;
  g.rows_per_img = p->rows_per_img > 0 ? p->rows_per_img : 1; g.ld_row_bias = p->ld_row_bias;
  g.act = p->act; g.stat_sum = p->stat_sum; g.stat_sqs = p->stat_sqs;
  g.shuffle_r = p->shuffle_r; g.shuffle_h = p->shuffle_h; g.shuffle_w = p->shuffle_w;
  g.shuffle_cs = p->shuffle_r > 0 ? p->N / (p->shuffle_r * p->shuffle_r) : 0;

  const int tf32 = p->dtype == DLB_F32;
  const int xform = !tf32 && (p->a_scale != nullptr || p->a_fin != nullptr);
  g.a_scale = p->a_scale; g.a_shift = p->a_shift; g.a_act = p->a_act;
  if (p->a_fin) { g.has_afin = 1; g.afin = *p->a_fin; }
  size_t tail = 0;
  const int sets = plan_tc(p->N, p->K, p->M, p->out_dtype, p->shuffle_r, xform, tf32, &g, &tail);
  DLB_REQUIRE(sets > 0, "pw_gemm: N=%d K=%d leaves no room for a 2-stage pipeline", p->N, p->K);
  g.idesc = make_idesc(tf32 ? 2 : (p->dtype == DLB_BF16 ? 1 : 0), 128, g.chunk_n, 0, 0);

  CUtensorMap ta, tb, tc, tbl;
  int rc = make_tmap_2d(&ta, p->dtype, p->A, p->M, p->K, p->lda, kBlockM, g.k_elems_per_block);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, p->dtype, p->Bt, p->N, p->K, p->ldb, g.chunk_n, g.k_elems_per_block);
  if (rc) return rc;
  tbl = tb;
  g.b_lo_given = 0;
  if (tf32 && p->Bt_lo) {
    DLB_REQUIRE((reinterpret_cast<uintptr_t>(p->Bt_lo) & 15) == 0, "pw_gemm: Bt_lo must be 16-byte aligned");
    rc = make_tmap_2d(&tbl, p->dtype, p->Bt_lo, p->N, p->K, p->ldb, g.chunk_n, g.k_elems_per_block);
    if (rc) return rc;
    g.b_lo_given = 1;
  }
  // output map for the bulk tensor store of staged 32-row x 64-column tiles (16-bit row-major outputs without residual)
  static const bool tma_store_on = [] { const char* e = getenv("DLB_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
  g.tma_store = 0;
  tc = ta;
  if (tma_store_on && sets == 4 && p->R == nullptr && (p->act == DLB_ACT_NONE || p->stat_sum == nullptr) &&
      (reinterpret_cast<uintptr_t>(p->C) & 15) == 0 && p->ldc % 8 == 0) {
    rc = make_tmap_2d(&tc, p->out_dtype, p->C, p->M, p->n_store, p->ldc, 32, 64);
    if (rc) return rc;
    g.tma_store = 1;
  }

  const size_t smem_bytes = 1024 + static_cast<size_t>(g.num_stages) * g.stage_bytes + g.b_res_bytes + tail;
  const int grid = split_ctas(&g, p->N);

  const bool lean = !p->col_scale && !p->col_shift && !p->row_bias;
  if (tf32) {
#define LAUNCHF(LEAN)                                                                                                       \
  do {                                                                                                                      \
    DLB_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<float, 3, LEAN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem_bytes));                                                                        \
    launch_k(pw_gemm_tc_kernel<float, 3, LEAN, false, true>, grid, 64 + 128 * 3, smem_bytes, st, ta, tb, tc, tbl, g);                  \
  } while (0)
    if (lean) LAUNCHF(true); else LAUNCHF(false);
#undef LAUNCHF
    g_launches++;
    return check_launch("pw_gemm_tc_kernel(tf32x3)");
  }
  if (xform) {
    // the training-forward project conv: raw depthwise output in, raw project output + statistics out
    DLB_REQUIRE(p->a_fin != nullptr || p->a_shift != nullptr, "pw_gemm: a_scale without a_shift");
    DLB_REQUIRE(sets == 4 && lean && g.tma_store && p->dtype == p->out_dtype,
                "pw_gemm: the A-operand transform needs a 16-bit row-major output of the input dtype, no output "
                "affine / bias / residual, and a 16-byte aligned C");
#define LAUNCHX(OT)                                                                                                       \
  do {                                                                                                                    \
    DLB_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<OT, 4, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem_bytes));                                                                      \
    launch_k(pw_gemm_tc_kernel<OT, 4, true, true, true>, grid, 64 + 128 * 4, smem_bytes, st, ta, tb, tc, tbl, g);                    \
  } while (0)
    if (p->out_dtype == DLB_F16) LAUNCHX(__half); else LAUNCHX(__nv_bfloat16);
#undef LAUNCHX
    g_launches++;
    return check_launch("pw_gemm_tc_kernel");
  }
#define LAUNCH(OT, SETS, LEAN, TMA)                                                                                   \
  do {                                                                                                                \
    DLB_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<OT, SETS, LEAN, TMA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem_bytes));                                                                  \
    launch_k(pw_gemm_tc_kernel<OT, SETS, LEAN, TMA, false>, grid, 64 + 128 * SETS, smem_bytes, st, ta, tb, tc, tbl, g);                 \
  } while (0)
#define LAUNCH2(OT, SETS, TMA) do { if (lean) LAUNCH(OT, SETS, true, TMA); else LAUNCH(OT, SETS, false, TMA); } while (0)
#define LAUNCH3(OT) do { if (sets == 4) { if (g.tma_store) LAUNCH2(OT, 4, true); else LAUNCH2(OT, 4, false); } else LAUNCH2(OT, 2, false); } while (0)
  if (p->out_dtype == DLB_F16) LAUNCH3(__half);
  else if (p->out_dtype == DLB_BF16) LAUNCH3(__nv_bfloat16);
  else LAUNCH2(float, 2, false);
#undef LAUNCH3
#undef LAUNCH2
#undef LAUNCH
  g_launches++;
  return check_launch("pw_gemm_tc_kernel");
}

static int launch_simt(const dlb_pw_gemm_params* p, cudaStream_t st) {
  DLB_REQUIRE(p->out_dtype == DLB_F32, "pw_gemm: f32 inputs require f32 output");
  DLB_REQUIRE(p->K % 4 == 0 && p->lda % 4 == 0 && p->ldb % 4 == 0, "pw_gemm(f32): K, lda, ldb must be multiples of 4");
  SimtArgs g{};
  g.M = p->M; g.N = p->N; g.K = p->K; g.lda = p->lda; g.ldb = p->ldb; g.ldc = p->ldc; g.ldr = p->ldr;
  g.n_store = p->n_store;
  g.A = (const float*)p->A; g.Bt = (const float*)p->Bt; g.C = (float*)p->C; g.R = (const float*)p->R;
  g.col_scale = p->col_scale; g.col_shift = p->col_shift; g.row_bias = p->row_bias;
  g.rows_per_img = p->rows_per_img > 0 ? p->rows_per_img : 1; g.ld_row_bias = p->ld_row_bias; g.act = p->act;
  g.stat_sum = p->stat_sum; g.stat_sqs = p->stat_sqs;
  g.shuffle_r = p->shuffle_r; g.shuffle_h = p->shuffle_h; g.shuffle_w = p->shuffle_w;
  g.shuffle_cs = p->shuffle_r > 0 ? p->N / (p->shuffle_r * p->shuffle_r) : 0;
  g.a_scale = p->a_scale; g.a_shift = p->a_shift; g.a_act = p->a_act;
  DLB_REQUIRE(!p->a_scale || (p->a_shift && (reinterpret_cast<uintptr_t>(p->a_scale) & 15) == 0 &&
                              (reinterpret_cast<uintptr_t>(p->a_shift) & 15) == 0),
              "pw_gemm(f32): a_scale / a_shift must both be given and 16-byte aligned");
  const int ncols = p->n_store > p->N ? p->n_store : p->N;
  dim3 grid((p->M + 63) / 64, (ncols + 63) / 64);
  launch_k(pw_gemm_simt_kernel, grid, 256, 0, st, g);
  g_launches++;
  return check_launch("pw_gemm_simt_kernel");
}

}  // namespace dlb

extern "C" int dlb_pw_gemm_plan(int M, int N, int K, int out_dtype, int shuffle_r, int* plan) {
  using namespace dlb;
  GemmArgs g{};
  size_t tail = 0;
  const int sets = plan_tc(N, K, M, out_dtype, shuffle_r & 0xffff, (shuffle_r >> 16) & 1, (shuffle_r >> 17) & 1, &g, &tail);
  if (sets == 0) return DLB_ERR_INVALID;
  plan[0] = sets; plan[1] = g.chunk_n; plan[2] = g.n_chunks; plan[3] = g.chunks_per_group; plan[4] = g.n_groups;
  plan[5] = g.acc_cols; plan[6] = g.acc_stages; plan[7] = g.alt_tiles; plan[8] = g.num_stages;
  plan[9] = static_cast<int>(1024 + static_cast<size_t>(g.num_stages) * g.stage_bytes + g.b_res_bytes + tail);
  plan[10] = split_ctas(&g, N);                                  // grid size
  for (int i = 0; i < 8; ++i) plan[11 + i] = i < g.n_groups ? g.grp_ctas[i] : 0;   // CTAs per column group
  return DLB_OK;
}

extern "C" int dlb_pw_gemm(const dlb_pw_gemm_params* p, void* stream) {
  using namespace dlb;
  DLB_REQUIRE(p && p->A && p->Bt && p->C, "pw_gemm: null pointer");
  DLB_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0, "pw_gemm: bad shape M=%d N=%d K=%d", p->M, p->N, p->K);
  DLB_REQUIRE(p->n_store > 0 && (p->shuffle_r > 0 || p->n_store <= p->ldc), "pw_gemm: n_store %d > ldc %d",
              p->n_store, p->ldc);
  if (p->shuffle_r > 0) {
    const int r = p->shuffle_r;
    DLB_REQUIRE(p->N % (r * r) == 0 && p->n_store == p->N, "pw_gemm: shuffle needs N %% r^2 == 0 and n_store == N");
    DLB_REQUIRE((r * (p->N / (r * r))) % 8 == 0, "pw_gemm: shuffle needs r*Cs %% 8 == 0");
    DLB_REQUIRE(p->M % (p->shuffle_h * p->shuffle_w) == 0, "pw_gemm: shuffle geometry does not divide M");
    DLB_REQUIRE(p->R == nullptr, "pw_gemm: residual with shuffle store unsupported");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->a_fin) {
    DLB_REQUIRE(p->a_scale == nullptr, "pw_gemm: give a_fin or a_scale / a_shift, not both");
    const int rc = check_bn_fin(p->a_fin, "pw_gemm");
    if (rc) return rc;
    if (p->dtype == DLB_F32) {      // the fp32 kernels read finished tables: stand-alone finalize, then a_scale / a_shift
      const int rf = bn_fin_standalone(p->K, p->a_fin, stream);
      if (rf) return rf;
      dlb_pw_gemm_params q = *p;
      q.a_scale = p->a_fin->scale; q.a_shift = p->a_fin->shift; q.a_fin = nullptr;
      return dlb_pw_gemm(&q, stream);
    }
  }
  if (p->dtype == DLB_F32) {
    // fp32 operands: 3xTF32 on the tensor cores (error ~1e-6 relative, the fp32 parity mode of BASELINE config 3) unless
    // the shape is tiny / unaligned / carries an A-operand transform, or DLB_F32_SIMT=1 asks for the exact FMA kernel
    static const bool force_simt = [] { const char* e = getenv("DLB_F32_SIMT"); return e && e[0] == '1'; }();
    const bool tc_ok = !force_simt && p->out_dtype == DLB_F32 && p->M >= 64 && p->K % 4 == 0 && p->lda % 4 == 0 &&
                       p->ldb % 4 == 0 && p->a_scale == nullptr && p->N <= 2048 &&
                       (reinterpret_cast<uintptr_t>(p->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->Bt) & 15) == 0 &&
                       p->n_store % 4 == 0 && (p->shuffle_r > 0 || p->ldc % 4 == 0) && (p->R == nullptr || p->ldr % 8 == 0);
    if (!tc_ok) return launch_simt(p, st);
    return launch_tc(p, st);
  }
  const int es = 2;
  DLB_REQUIRE((p->lda * es) % 16 == 0 && (p->ldb * es) % 16 == 0, "pw_gemm: lda/ldb must be 16-byte multiples");
  DLB_REQUIRE((reinterpret_cast<uintptr_t>(p->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->Bt) & 15) == 0,
              "pw_gemm: A/Bt must be 16-byte aligned");
  const int vec = p->out_dtype == DLB_F32 ? 4 : 8;
  DLB_REQUIRE(p->n_store % vec == 0 && (p->shuffle_r > 0 || p->ldc % vec == 0), "pw_gemm: n_store/ldc must be multiples of %d", vec);
  DLB_REQUIRE(p->R == nullptr || p->ldr % 8 == 0, "pw_gemm: ldr must be a multiple of 8");
  DLB_REQUIRE(p->N <= 2048, "pw_gemm: N > 2048 unsupported");
  return launch_tc(p, st);
}

// Shared device/host helpers for the sm_100a DeepLabV3+ kernels.
// PTX wrappers (mbarrier, TMA, tcgen05, TMEM), dtype traits, error plumbing.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <atomic>
#include <type_traits>

#include "../../include/deeplab_b200.h"

namespace dlb {

// ---------------------------------------------------------------------------------------------
// error plumbing (never throws across the C ABI)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int  check_launch(const char* what);   // cudaGetLastError -> DLB_ERR_CUDA

#define DLB_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      ::dlb::set_last_error(__VA_ARGS__);                       \
      return DLB_ERR_INVALID;                                   \
    }                                                           \
  } while (0)

#define DLB_CUDA(call)                                                           \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      ::dlb::set_last_error("%s failed: %s", #call, cudaGetErrorString(e__));    \
      return DLB_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

inline int dtype_size(int dt) { return dt == DLB_F32 ? 4 : 2; }
int num_sms();

// ---------------------------------------------------------------------------------------------
// dtype traits: activations are stored as half / bf16 / float, math is always fp32
// ---------------------------------------------------------------------------------------------
template <typename T> struct Act;
template <> struct Act<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float rnd(float v) { return v; }
};
template <> struct Act<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
};
template <> struct Act<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

// 8-wide vector access (16 B for 16-bit types, 2x16 B for float)
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__half> {
  static __device__ __forceinline__ void ld(const __half* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void st(__half* p, const float (&v)[8]) {
    uint4 u; __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

// raw (unconverted) 8-element vectors: keep many loads in flight at 4 registers each, convert at use
template <typename T> struct Raw8 { uint4 q; };
template <> struct Raw8<float> { uint4 q, r; };
template <typename T> __device__ __forceinline__ void raw_ld(const T* p, Raw8<T>& o) { o.q = *reinterpret_cast<const uint4*>(p); }
template <> __device__ __forceinline__ void raw_ld<float>(const float* p, Raw8<float>& o) {
  o.q = *reinterpret_cast<const uint4*>(p); o.r = *reinterpret_cast<const uint4*>(p + 4);
}
__device__ __forceinline__ void raw_unpack(const Raw8<__half>& r, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r.q);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void raw_unpack(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.q);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void raw_unpack(const Raw8<float>& r, float (&v)[8]) {
  v[0] = __uint_as_float(r.q.x); v[1] = __uint_as_float(r.q.y); v[2] = __uint_as_float(r.q.z); v[3] = __uint_as_float(r.q.w);
  v[4] = __uint_as_float(r.r.x); v[5] = __uint_as_float(r.r.y); v[6] = __uint_as_float(r.r.z); v[7] = __uint_as_float(r.r.w);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DLB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DLB_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}
// derivative mask of the activation evaluated at pre-activation z
__device__ __forceinline__ float act_mask(float z, int act) {
  if (act == DLB_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == DLB_ACT_RELU6) return (z > 0.f && z < 6.f) ? 1.f : 0.f;
  return 1.f;
}

// ---------------------------------------------------------------------------------------------
// PTX: mbarrier / TMA / tcgen05
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// --- TMEM ---
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)),
               "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; one elected thread issues for the whole CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit: arrives on the mbarrier once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread = lane = accumulator row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// UMMA shared-memory descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor layout, version 1).
//   K-major : rows of 128 B (64 x 16-bit or 32 x fp32), 8-row swizzle atoms, SBO = 1024 B, LBO unused.
//   MN-major: the same physical tile read transposed: MN runs along the 128-B row, LBO = byte distance
//             between 64-element (128 B) MN blocks, SBO = 1024 B between groups of 8 K rows.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, dense
//   fmt: 0 = f16, 1 = bf16, 2 = tf32 ; major: 0 = K-major, 1 = MN-major
__host__ __device__ inline uint32_t make_idesc(int fmt, int m, int n, int a_major, int b_major) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= static_cast<uint32_t>(fmt) << 7;
  d |= static_cast<uint32_t>(fmt) << 10;
  d |= static_cast<uint32_t>(a_major) << 15;
  d |= static_cast<uint32_t>(b_major) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

// packed fp32 pairs (sm_100 FADD2 / FFMA2): one instruction per two columns in the statistics walk
__device__ __forceinline__ void f32x2_acc(unsigned long long& s, unsigned long long& q, float a, float b) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s) : "l"(v));
  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q) : "l"(v));
}
__device__ __forceinline__ float2 f32x2_unpack(unsigned long long v) {
  float2 f;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(v));
  return f;
}
__device__ __forceinline__ unsigned long long f32x2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f32x2_pack(float a, float b) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
  return v;
}
// s += v ; q += v * v   on a packed pair
__device__ __forceinline__ void f32x2_acc_v(unsigned long long& s, unsigned long long& q, unsigned long long v) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s) : "l"(v));
  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q) : "l"(v));
}

// Sum 16 per-thread values (one per column) over the 32 lanes of a warp with 15+1 shuffles instead of 80:
// recursive halving -- after the call, lane l holds the full 32-lane sum of column (l >> 1) & 15 ... see below.
// Returns the total for column `col_of_lane(lane)`; both lanes 2c and 2c+1 return the same value.
__device__ __forceinline__ int reduce16_col_of_lane(int lane) {
  // bit4 of lane picks half of 16, bit3 half of 8, bit2 half of 4, bit1 half of 2
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
  // step 1: partner lane^16, keep 8
  float a8[8];
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float send = hi ? v[i] : v[i + 8];
      float keep = hi ? v[i + 8] : v[i];
      a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  float a4[4];
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float send = hi ? a8[i] : a8[i + 4];
      float keep = hi ? a8[i + 4] : a8[i];
      a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  float a2[2];
  {
    const bool hi = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = hi ? a4[i] : a4[i + 2];
      float keep = hi ? a4[i + 2] : a4[i];
      a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  float a1;
  {
    const bool hi = (lane & 2) != 0;
    float send = hi ? a2[0] : a2[1];
    float keep = hi ? a2[1] : a2[0];
    a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}

// ---------------------------------------------------------------------------------------------
// A-operand transform: BatchNorm affine + activation applied IN PLACE to a landed SWIZZLE_128B tile (rows of 64
// 16-bit channels = 128 B, 16-byte chunk c of row r stored at chunk c ^ (r & 7)) between the TMA load and tcgen05.mma.
// Lets a GEMM consume the raw (pre-BatchNorm) output of the previous convolution, so the normalised activation is
// never written to HBM (deeplabv3p.py:189-196: depthwise_BN + relu6 feeding the project conv).  The caller then
// executes fence.proxy.async and signals the MMA issuer.  sc / sh: shared-memory addresses of this box's 64 scales /
// shifts (fp32).  Lanes of a warp own consecutive rows, so chunk index i ^ (r & 7) is conflict-free.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 lds128_u32(uint32_t saddr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(saddr));
  return u;
}
__device__ __forceinline__ void sts128_u32(uint32_t saddr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
// one 32-bit word (two 16-bit channels): fp32 affine on a packed pair (FFMA2), rounded back, clamped on the packed
// 16-bit pair -- 6 instructions per two channels
template <int kAct> __device__ __forceinline__ uint32_t xform_word_h(uint32_t w, unsigned long long sc2, unsigned long long sh2) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
  unsigned long long v = f32x2_pack(f.x, f.y);
  asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(sc2), "l"(sh2));
  const float2 o = f32x2_unpack(v);
  __half2 h = __floats2half2_rn(o.x, o.y);
  if (kAct != DLB_ACT_NONE) h = __hmax2(h, __float2half2_rn(0.f));
  if (kAct == DLB_ACT_RELU6) h = __hmin2(h, __float2half2_rn(6.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <int kAct> __device__ __forceinline__ uint32_t xform_word_b(uint32_t w, unsigned long long sc2, unsigned long long sh2) {
  unsigned long long v = f32x2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
  asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(sc2), "l"(sh2));
  const float2 o = f32x2_unpack(v);
  __nv_bfloat162 h = __floats2bfloat162_rn(o.x, o.y);
  if (kAct != DLB_ACT_NONE) h = __hmax2(h, __float2bfloat162_rn(0.f));
  if (kAct == DLB_ACT_RELU6) h = __hmin2(h, __float2bfloat162_rn(6.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
// Rows r and r + dr (same r & 7) of one box, logical chunks c0 .. c0 + kChunks - 1: the per-channel tables are read once
// for both rows (broadcast LDS are half of the transform's shared-memory instructions otherwise).
template <typename T, int kAct, int kChunks>
__device__ __forceinline__ void xform_rows2_act(uint32_t box_saddr, int r, int dr, int c0, uint32_t sc, uint32_t sh) {
  auto pair = [](uint32_t lo, uint32_t hi) { return (static_cast<unsigned long long>(hi) << 32) | lo; };
#pragma unroll
  for (int i = c0; i < c0 + kChunks; ++i) {
    const uint32_t a0 = box_saddr + r * 128 + (static_cast<uint32_t>(i ^ (r & 7)) << 4);
    const uint32_t a1 = a0 + dr * 128;
    uint4 u = lds128_u32(a0), w = lds128_u32(a1);
    const uint4 s0 = lds128_u32(sc + i * 32), s1 = lds128_u32(sc + i * 32 + 16);
    const uint4 h0 = lds128_u32(sh + i * 32), h1 = lds128_u32(sh + i * 32 + 16);
    const unsigned long long sA = pair(s0.x, s0.y), sB = pair(s0.z, s0.w), sC = pair(s1.x, s1.y), sD = pair(s1.z, s1.w);
    const unsigned long long hA = pair(h0.x, h0.y), hB = pair(h0.z, h0.w), hC = pair(h1.x, h1.y), hD = pair(h1.z, h1.w);
    if constexpr (std::is_same<T, __half>::value) {
      u.x = xform_word_h<kAct>(u.x, sA, hA); u.y = xform_word_h<kAct>(u.y, sB, hB);
      u.z = xform_word_h<kAct>(u.z, sC, hC); u.w = xform_word_h<kAct>(u.w, sD, hD);
      w.x = xform_word_h<kAct>(w.x, sA, hA); w.y = xform_word_h<kAct>(w.y, sB, hB);
      w.z = xform_word_h<kAct>(w.z, sC, hC); w.w = xform_word_h<kAct>(w.w, sD, hD);
    } else {
      u.x = xform_word_b<kAct>(u.x, sA, hA); u.y = xform_word_b<kAct>(u.y, sB, hB);
      u.z = xform_word_b<kAct>(u.z, sC, hC); u.w = xform_word_b<kAct>(u.w, sD, hD);
      w.x = xform_word_b<kAct>(w.x, sA, hA); w.y = xform_word_b<kAct>(w.y, sB, hB);
      w.z = xform_word_b<kAct>(w.z, sC, hC); w.w = xform_word_b<kAct>(w.w, sD, hD);
    }
    sts128_u32(a0, u);
    sts128_u32(a1, w);
  }
}
template <typename T, int kChunks>
__device__ __forceinline__ void xform_rows2(uint32_t box_saddr, int r, int dr, int c0, uint32_t sc, uint32_t sh, int act) {
  if (act == DLB_ACT_RELU6) xform_rows2_act<T, DLB_ACT_RELU6, kChunks>(box_saddr, r, dr, c0, sc, sh);
  else if (act == DLB_ACT_RELU) xform_rows2_act<T, DLB_ACT_RELU, kChunks>(box_saddr, r, dr, c0, sc, sh);
  else xform_rows2_act<T, DLB_ACT_NONE, kChunks>(box_saddr, r, dr, c0, sc, sh);
}

// ---------------------------------------------------------------------------------------------
// Training-mode BatchNorm finalisation of one channel from the fp64 sums (dlb_bn_fin, deeplab_b200.h): the one place
// this arithmetic lives -- the stand-alone kernel and the consumer-side prologues (bn_stream, the A-operand transform
// of pw_gemm, the depthwise prologue) all call it, so a layer's scale / shift do not depend on which kernel finalised it.
// Keras 2.2.4 / TF backend: the moving average is fed the Bessel-corrected variance.
// ---------------------------------------------------------------------------------------------
// Every multiply-add is spelled as an explicit fma / mul intrinsic: left to the compiler, contraction differed between the
// kernels this is inlined into and the results were one ulp apart.
__device__ __forceinline__ void bn_fin_channel(const dlb_bn_fin& f, int c, bool publish, float& sc, float& sh) {
  const double mean = f.sum[c] / f.count;
  double var = __fma_rn(-mean, mean, f.sqs[c] / f.count);
  if (var < 0.0) var = 0.0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(f.eps)));
  const float mean_f = static_cast<float>(mean);
  sc = __fmul_rn(f.gamma[c], rstd);
  sh = __fmaf_rn(-mean_f, sc, f.beta[c]);
  if (publish) {
    f.scale[c] = sc;
    f.shift[c] = sh;
    if (f.mean) f.mean[c] = mean_f;
    if (f.rstd) f.rstd[c] = rstd;
    if (f.moving_mean) {
      const double unbiased = f.count > 1.0 ? var * f.count / (f.count - 1.0) : var;
      const float keep = 1.f - f.momentum;
      f.moving_mean[c] = __fmaf_rn(f.moving_mean[c], f.momentum, __fmul_rn(keep, mean_f));
      f.moving_var[c] = __fmaf_rn(f.moving_var[c], f.momentum, __fmul_rn(keep, static_cast<float>(unbiased)));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_prologue(), so the CTAs of kernel i+1 are
// scheduled (and run their launch / index-math prologue) while kernel i drains; griddepcontrol.wait returns only
// after the preceding grid has completed and flushed, so no kernel touches global memory early.  A step is ~430
// dependent launches; this hides most of the kernel-boundary latency inside the captured CUDA graph.
// DLB_PDL=0 in the environment disables the attribute (plain stream order).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
#ifdef DLB_PDL_EARLY_TRIGGER
  pdl_trigger();
#endif
  pdl_wait();
}
bool pdl_enabled();
// dlb_bn_fin given to a kernel's host entry: argument check, and the stand-alone finalize launch for the shapes whose
// kernel variant has no consumer-side prologue (bn_ops.cu)
int check_bn_fin(const dlb_bn_fin* f, const char* who);
int bn_fin_standalone(int C, const dlb_bn_fin* f, void* stream);
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through check_launch()
}

}  // namespace dlb

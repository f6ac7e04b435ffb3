// BatchNormalization (training-mode batch statistics + inference fold), activation/residual/dropout apply,
// BN backward (two passes), ASPP global average pooling, and a tiny fp32 GEMM for the [B, C] pooled branch.
// Reference sites: BatchNormalization deeplabv3p.py:76,80,178,189,197,322,379,386,408 ; relu6 Lambda :181,192,325 ;
// Add :202 ; Dropout :410 ; AveragePooling2D :375.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace dlb {

extern std::atomic<long long> g_launches;
// DLB_BN_STREAM=0 selects the register-staged BatchNorm streaming kernels (A/B measurements)
static const bool g_bn_stream = [] { const char* e = getenv("DLB_BN_STREAM"); return !(e && e[0] == '0'); }();

__global__ void bn_finalize_kernel(int C, const dlb_bn_fin f, double* sum, double* sqs, int reset) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sc, sh;
  bn_fin_channel(f, c, true, sc, sh);
  if (reset) { sum[c] = 0.0; sqs[c] = 0.0; }
}

__global__ void bn_fold_kernel(int C, const float* gamma, const float* beta, const float* mm, const float* mv,
                               float eps, float* scale, float* shift) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(mv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - mm[c] * sc;
}

// counter-based uniform in [0,1): splitmix64 finaliser on (seed, index)
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}

struct ApplyArgs {
  long long nvec; int C; const void* x; void* y; const void* res;
  const float* scale; const float* shift; int act; float drop_rate; uint64_t seed; const long long* seed_dev;
};

// thread = fixed 8-channel vector (scale/shift in registers); a CTA iteration covers U * rpb CONSECUTIVE rows (one
// contiguous block of memory), 4 rows in flight per thread
template <typename T>
__global__ void __launch_bounds__(256, 2) bn_apply_kernel(const ApplyArgs a) {
  pdl_prologue();
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* res = reinterpret_cast<const T*>(a.res);
  T* y = reinterpret_cast<T*>(a.y);
  const int cv = a.C >> 3;
  const int rpb = 256 / cv;
  if (static_cast<int>(threadIdx.x) >= rpb * cv) return;
  const int r_in = threadIdx.x / cv;
  const int c0 = (threadIdx.x - r_in * cv) * 8;
  const long long M = a.nvec / cv;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = a.scale ? a.scale[c0 + k] : 1.f; sh[k] = a.scale ? a.shift[c0 + k] : 0.f; }
  const float keep_inv = a.drop_rate > 0.f ? 1.f / (1.f - a.drop_rate) : 1.f;
  const uint64_t seed = a.seed + (a.seed_dev ? static_cast<uint64_t>(*a.seed_dev) * 0x632BE59BD9B4E019ull : 0ull);
  const long long rstride = static_cast<long long>(gridDim.x) * rpb;
  constexpr int U = 4;
  for (long long r0 = static_cast<long long>(blockIdx.x) * rpb * U + r_in; r0 < M; r0 += rstride * U) {
    float v[U][8], rr[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * rpb;
      if (r < M) {
        Vec8<T>::ld(x + r * a.C + c0, v[u]);
        if (res) Vec8<T>::ld(res + r * a.C + c0, rr[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * rpb;
      if (r >= M) continue;
      const long long e0 = r * a.C + c0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float z = apply_act(fmaf(v[u][k], sc[k], sh[k]), a.act);
        if (a.drop_rate > 0.f) z = hash_uniform(seed, static_cast<uint64_t>(e0 + k)) >= a.drop_rate ? z * keep_inv : 0.f;
        if (res) z += rr[u][k];
        v[u][k] = z;
      }
      Vec8<T>::st(y + e0, v[u]);
    }
  }
}

struct BwdArgs {
  long long M; int C; const void* x; const void* da; void* dx;
  const float* scale; const float* shift; const float* mean; const float* rstd; int act;
  double* red; float* dgamma; float* dbeta; float drop_rate; uint64_t seed; const long long* seed_dev; int frozen;
  int cv, rpb;   // channel vectors, rows per block
};

template <typename T>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(const BwdArgs a) {
  pdl_prologue();
  extern __shared__ float s_red[];   // [2*C]
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * a.C; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  if (tid < a.rpb * a.cv) {
    const int r_in = tid / a.cv;
    const int c0 = (tid - r_in * a.cv) * 8;
    float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = a.scale[c0 + k]; sh[k] = a.shift[c0 + k]; mu[k] = a.mean[c0 + k]; rs[k] = a.rstd[c0 + k]; }
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const T* x = reinterpret_cast<const T*>(a.x);
    const T* da = reinterpret_cast<const T*>(a.da);
    const float keep_inv = a.drop_rate > 0.f ? 1.f / (1.f - a.drop_rate) : 1.f;
    const uint64_t seed = a.seed + (a.seed_dev ? static_cast<uint64_t>(*a.seed_dev) * 0x632BE59BD9B4E019ull : 0ull);
    const long long rstride = static_cast<long long>(gridDim.x) * a.rpb;
    constexpr int U = 4;
    for (long long r0 = static_cast<long long>(blockIdx.x) * a.rpb * U + r_in; r0 < a.M; r0 += rstride * U) {
      float xv[U][8], gv[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * a.rpb;
        if (r < a.M) { Vec8<T>::ld(x + r * a.C + c0, xv[u]); Vec8<T>::ld(da + r * a.C + c0, gv[u]); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * a.rpb;
        if (r >= a.M) continue;
        const long long e0 = r * a.C + c0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float z = fmaf(xv[u][k], sc[k], sh[k]);
          float dz = gv[u][k] * act_mask(z, a.act);
          if (a.drop_rate > 0.f) dz = hash_uniform(seed, static_cast<uint64_t>(e0 + k)) >= a.drop_rate ? dz * keep_inv : 0.f;
          s1[k] += dz;
          s2[k] += dz * (xv[u][k] - mu[k]) * rs[k];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s_red[c0 + k], s1[k]); atomicAdd(&s_red[a.C + c0 + k], s2[k]); }
  }
  __syncthreads();
  for (int i = tid; i < 2 * a.C; i += blockDim.x) atomicAdd(&a.red[i], static_cast<double>(s_red[i]));
}

// fp16 specialisation of the reduce pass: packed half2 arithmetic, 4-row partial sums on half2 folded into fp32
// accumulators (the generic kernel is instruction-issue bound: ~13 instructions per element, 1.7 TB/s)
template <typename T> struct BP2;
template <> struct BP2<__half> {
  using t = __half2;
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2half2_rn(a, b); }
  static __device__ __forceinline__ t bcast(float a) { return __float2half2_rn(a); }
  static __device__ __forceinline__ float2 unpack(t v) { return __half22float2(v); }
};
template <> struct BP2<__nv_bfloat16> {
  using t = __nv_bfloat162;
  static __device__ __forceinline__ t pack(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static __device__ __forceinline__ t bcast(float a) { return __float2bfloat162_rn(a); }
  static __device__ __forceinline__ float2 unpack(t v) { return __bfloat1622float2(v); }
};
template <typename T> struct __align__(16) BV8 { typename BP2<T>::t h[4]; };
// 16-byte global accesses go through uint4: nvcc scalarises a copy of the half2[4] struct into four 32-bit LDG/STG
template <typename T> __device__ __forceinline__ BV8<T> ldg_bh8(const void* p) { uint4 u = *reinterpret_cast<const uint4*>(p); return *reinterpret_cast<BV8<T>*>(&u); }
template <typename T> __device__ __forceinline__ void stg_bh8(void* p, const BV8<T>& o) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&o); }

template <typename T>
__global__ void __launch_bounds__(256, 3) bn_bwd_reduce_h_kernel(const BwdArgs a) {
  pdl_prologue();
  extern __shared__ float s_red[];   // [rpb][2*C] per-row-group partial sums (no shared atomics: they were a 32-way
                                     // CAS contention, a third of the kernel on the narrow project BNs)
  const int tid = threadIdx.x;
  if (tid < a.rpb * a.cv) {
    const int r_in = tid / a.cv;
    const int c0 = (tid - r_in * a.cv) * 8;
    typename BP2<T>::t sc2[4], sh2[4], mu2[4], rs2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sc2[k] = BP2<T>::pack(a.scale[c0 + 2 * k], a.scale[c0 + 2 * k + 1]);
      sh2[k] = BP2<T>::pack(a.shift[c0 + 2 * k], a.shift[c0 + 2 * k + 1]);
      mu2[k] = BP2<T>::pack(a.mean[c0 + 2 * k], a.mean[c0 + 2 * k + 1]);
      rs2[k] = BP2<T>::pack(a.rstd[c0 + 2 * k], a.rstd[c0 + 2 * k + 1]);
    }
    const typename BP2<T>::t zero2 = BP2<T>::bcast(0.f), six2 = BP2<T>::bcast(6.f);
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const T* x = reinterpret_cast<const T*>(a.x);
    const T* da = reinterpret_cast<const T*>(a.da);
    const long long rstride = static_cast<long long>(gridDim.x) * a.rpb;
    constexpr int U = 4;
    for (long long r0 = static_cast<long long>(blockIdx.x) * a.rpb * U + r_in; r0 < a.M; r0 += rstride * U) {
      BV8<T> xv[U], gv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * a.rpb;
        if (r < a.M) {
          xv[u] = ldg_bh8<T>(x + r * a.C + c0);
          gv[u] = ldg_bh8<T>(da + r * a.C + c0);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) { xv[u].h[k] = zero2; gv[u].h[k] = zero2; }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        typename BP2<T>::t p1 = zero2, p2 = zero2;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          typename BP2<T>::t dz = gv[u].h[k];
          if (a.act != DLB_ACT_NONE) {
            const typename BP2<T>::t z = __hfma2(xv[u].h[k], sc2[k], sh2[k]);
            typename BP2<T>::t m = __hgt2(z, zero2);
            if (a.act == DLB_ACT_RELU6) m = __hmul2(m, __hlt2(z, six2));
            dz = __hmul2(dz, m);
          }
          p1 = __hadd2(p1, dz);
          p2 = __hfma2(dz, __hmul2(__hsub2(xv[u].h[k], mu2[k]), rs2[k]), p2);
        }
        const float2 f1 = BP2<T>::unpack(p1), f2 = BP2<T>::unpack(p2);
        s1[2 * k] += f1.x; s1[2 * k + 1] += f1.y; s2[2 * k] += f2.x; s2[2 * k + 1] += f2.y;
      }
    }
    float* row = s_red + static_cast<size_t>(r_in) * 2 * a.C;
    *reinterpret_cast<float4*>(row + c0) = make_float4(s1[0], s1[1], s1[2], s1[3]);
    *reinterpret_cast<float4*>(row + c0 + 4) = make_float4(s1[4], s1[5], s1[6], s1[7]);
    *reinterpret_cast<float4*>(row + a.C + c0) = make_float4(s2[0], s2[1], s2[2], s2[3]);
    *reinterpret_cast<float4*>(row + a.C + c0 + 4) = make_float4(s2[4], s2[5], s2[6], s2[7]);
  }
  __syncthreads();
  for (int i = tid; i < 2 * a.C; i += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < a.rpb; ++r) t += s_red[static_cast<size_t>(r) * 2 * a.C + i];
    atomicAdd(&a.red[i], static_cast<double>(t));
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(const BwdArgs a) {
  pdl_prologue();
  const int tid = threadIdx.x;
  if (blockIdx.x == 0) {
    for (int c = tid; c < a.C; c += blockDim.x) {
      if (a.dbeta) a.dbeta[c] = static_cast<float>(a.red[c]);
      // d gamma = sum dz * xhat
      if (a.dgamma) a.dgamma[c] = static_cast<float>(a.red[a.C + c]);
    }
  }
  if (tid >= a.rpb * a.cv) return;
  const int r_in = tid / a.cv;
  const int c0 = (tid - r_in * a.cv) * 8;
  float sc[8], sh[8], mu[8], rs[8], k1[8], k2[8];
  const float invM = 1.f / static_cast<float>(a.M);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = a.scale[c0 + k]; sh[k] = a.shift[c0 + k]; mu[k] = a.mean[c0 + k]; rs[k] = a.rstd[c0 + k];
    k1[k] = a.frozen ? 0.f : static_cast<float>(a.red[c0 + k]) * invM;
    k2[k] = a.frozen ? 0.f : static_cast<float>(a.red[a.C + c0 + k]) * invM;
  }
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* da = reinterpret_cast<const T*>(a.da);
  T* dx = reinterpret_cast<T*>(a.dx);
  const float keep_inv = a.drop_rate > 0.f ? 1.f / (1.f - a.drop_rate) : 1.f;
  const uint64_t seed = a.seed + (a.seed_dev ? static_cast<uint64_t>(*a.seed_dev) * 0x632BE59BD9B4E019ull : 0ull);
  const long long rstride = static_cast<long long>(gridDim.x) * a.rpb;
  constexpr int U = 4;
  for (long long r0 = static_cast<long long>(blockIdx.x) * a.rpb * U + r_in; r0 < a.M; r0 += rstride * U) {
    float xv[U][8], gv[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * a.rpb;
      if (r < a.M) { Vec8<T>::ld(x + r * a.C + c0, xv[u]); Vec8<T>::ld(da + r * a.C + c0, gv[u]); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * a.rpb;
      if (r >= a.M) continue;
      const long long e0 = r * a.C + c0;
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float z = fmaf(xv[u][k], sc[k], sh[k]);
        float dz = gv[u][k] * act_mask(z, a.act);
        if (a.drop_rate > 0.f) dz = hash_uniform(seed, static_cast<uint64_t>(e0 + k)) >= a.drop_rate ? dz * keep_inv : 0.f;
        const float xhat = (xv[u][k] - mu[k]) * rs[k];
        o[k] = sc[k] * (dz - k1[k] - xhat * k2[k]);
      }
      Vec8<T>::st(dx + e0, o);
    }
  }
}

// fp16 specialisation of the apply pass: mask and x-hat on packed half2, the projection
// dz - mean(dz) - xhat * mean(dz * xhat) in fp32 (the mean terms are far below one fp16 ulp of dz)
template <typename T>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply_h_kernel(const BwdArgs a) {
  pdl_prologue();
  const int tid = threadIdx.x;
  if (blockIdx.x == 0) {
    for (int c = tid; c < a.C; c += blockDim.x) {
      if (a.dbeta) a.dbeta[c] = static_cast<float>(a.red[c]);
      if (a.dgamma) a.dgamma[c] = static_cast<float>(a.red[a.C + c]);
    }
  }
  if (tid >= a.rpb * a.cv) return;
  const int r_in = tid / a.cv;
  const int c0 = (tid - r_in * a.cv) * 8;
  typename BP2<T>::t sc2[4], sh2[4], mu2[4], rs2[4];
  float sc[8], k1[8], k2[8];
  const float invM = 1.f / static_cast<float>(a.M);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sc2[k] = BP2<T>::pack(a.scale[c0 + 2 * k], a.scale[c0 + 2 * k + 1]);
    sh2[k] = BP2<T>::pack(a.shift[c0 + 2 * k], a.shift[c0 + 2 * k + 1]);
    mu2[k] = BP2<T>::pack(a.mean[c0 + 2 * k], a.mean[c0 + 2 * k + 1]);
    rs2[k] = BP2<T>::pack(a.rstd[c0 + 2 * k], a.rstd[c0 + 2 * k + 1]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = a.scale[c0 + k];
    k1[k] = a.frozen ? 0.f : static_cast<float>(a.red[c0 + k]) * invM;
    k2[k] = a.frozen ? 0.f : static_cast<float>(a.red[a.C + c0 + k]) * invM;
  }
  const typename BP2<T>::t zero2 = BP2<T>::bcast(0.f), six2 = BP2<T>::bcast(6.f);
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* da = reinterpret_cast<const T*>(a.da);
  T* dx = reinterpret_cast<T*>(a.dx);
  const long long rstride = static_cast<long long>(gridDim.x) * a.rpb;
  constexpr int U = 4;
  for (long long r0 = static_cast<long long>(blockIdx.x) * a.rpb * U + r_in; r0 < a.M; r0 += rstride * U) {
    BV8<T> xv[U], gv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * a.rpb;
      if (r < a.M) {
        xv[u] = ldg_bh8<T>(x + r * a.C + c0);
        gv[u] = ldg_bh8<T>(da + r * a.C + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * a.rpb;
      if (r >= a.M) continue;
      BV8<T> o;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        typename BP2<T>::t dz = gv[u].h[k];
        if (a.act != DLB_ACT_NONE) {
          const typename BP2<T>::t z = __hfma2(xv[u].h[k], sc2[k], sh2[k]);
          typename BP2<T>::t m = __hgt2(z, zero2);
          if (a.act == DLB_ACT_RELU6) m = __hmul2(m, __hlt2(z, six2));
          dz = __hmul2(dz, m);
        }
        const float2 dzf = BP2<T>::unpack(dz);
        const float2 xh = BP2<T>::unpack(__hmul2(__hsub2(xv[u].h[k], mu2[k]), rs2[k]));
        const float o0 = sc[2 * k] * (dzf.x - k1[2 * k] - xh.x * k2[2 * k]);
        const float o1 = sc[2 * k + 1] * (dzf.y - k1[2 * k + 1] - xh.y * k2[2 * k + 1]);
        o.h[k] = BP2<T>::pack(o0, o1);
      }
      stg_bh8(dx + r * a.C + c0, o);
    }
  }
}

// global average pool over HW of act(x*scale+shift): out[b, c] (+)= partial means (out pre-zeroed)
template <typename T>
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(int HW, int C, int cv, int rpb, int splits, const T* x,
                                                          const float* in_scale, const float* in_shift, int in_act,
                                                          float* out) {
  pdl_prologue();
  extern __shared__ float s_acc[];   // [C]
  const int tid = threadIdx.x;
  const int b = blockIdx.x / splits, sp = blockIdx.x - b * splits;
  for (int i = tid; i < C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  if (tid < rpb * cv) {
    const int r_in = tid / cv;
    const int c0 = (tid - r_in * cv) * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = sp * rpb + r_in; r < HW; r += splits * rpb) {
      float v[8];
      Vec8<T>::ld(x + (static_cast<size_t>(b) * HW + r) * C + c0, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float z = in_scale ? fmaf(v[k], in_scale[c0 + k], in_shift[c0 + k]) : v[k];
        acc[k] += apply_act(z, in_act);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_acc[c0 + k], acc[k]);
  }
  __syncthreads();
  const float inv = 1.f / static_cast<float>(HW);
  for (int i = tid; i < C; i += blockDim.x) atomicAdd(&out[static_cast<size_t>(b) * C + i], s_acc[i] * inv);
}

template <typename T>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(long long nvec, int HW, int C, const float* dout, T* dx,
                                                          int accumulate) {
  pdl_prologue();
  const float inv = 1.f / static_cast<float>(HW);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e0 = i * 8;
    const int c0 = static_cast<int>(e0 % C);
    const long long b = e0 / (static_cast<long long>(HW) * C);
    float v[8];
    if (accumulate) Vec8<T>::ld(dx + e0, v);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += dout[b * C + c0 + k] * inv;
    Vec8<T>::st(dx + e0, v);
  }
}

// tiny fp32 GEMM, one thread per output element (M, N <= a few hundred)
__global__ void small_gemm_kernel(int M, int N, int K, const float* A, int lda, int tA, const float* B, int ldb,
                                  int tB, float* C, int ldc, float alpha, float beta) {
  pdl_prologue();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = tA ? A[static_cast<size_t>(k) * lda + m] : A[static_cast<size_t>(m) * lda + k];
    const float b = tB ? B[static_cast<size_t>(n) * ldb + k] : B[static_cast<size_t>(k) * ldb + n];
    acc = fmaf(a, b, acc);
  }
  float* c = C + static_cast<size_t>(m) * ldc + n;
  *c = alpha * acc + (beta != 0.f ? beta * *c : 0.f);
}

// C[m, n] = alpha * sum_k A[m, k] * B[n, k] (+ beta * C): both operands contiguous along k (the dgrad GEMMs of the image
// pooling branch, M = batch).  One thread per output walked B rows with a stride of ldb between lanes and a serial
// K-long FMA chain (25 us for 16 x 256 x 256); here a warp owns column n, lanes stride over k (coalesced), the B row is
// read once into registers and reused for all M rows, one shuffle reduction per output.
constexpr int kNtKP = 16;     // K <= 32 * kNtKP
__global__ void __launch_bounds__(256) small_gemm_nt_kernel(int M, int N, int K, const float* A, int lda, const float* B,
                                                            int ldb, float* C, int ldc, float alpha, float beta) {
  pdl_prologue();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  float b[kNtKP];
#pragma unroll
  for (int i = 0; i < kNtKP; ++i) {
    const int k = lane + 32 * i;
    b[i] = k < K ? B[static_cast<size_t>(n) * ldb + k] : 0.f;
  }
  for (int m = 0; m < M; ++m) {
    const float* a = A + static_cast<size_t>(m) * lda;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < kNtKP; ++i) {
      const int k = lane + 32 * i;
      if (k < K) acc = fmaf(a[k], b[i], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float* c = C + static_cast<size_t>(m) * ldc + n;
      *c = alpha * acc + (beta != 0.f ? beta * *c : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Bulk-copy pipelined versions of the three BatchNorm streaming kernels (16-bit storage, no dropout).
//
// The register-staged kernels above top out at 50-60 % of the HBM copy bandwidth: ncu shows every warp parked on the
// long scoreboard with a third of the warp slots filled (80 registers x 256 threads) -- too few bytes in flight.
// Here one producer thread streams whole row blocks (rows are contiguous in NHWC: R rows = R*C*2 bytes in ONE
// cp.async.bulk) into a ring of shared-memory stages, ~190 KB in flight per SM at no register cost, and 16 consumer
// warps read them back with conflict-free LDS.128 (thread = fixed 8-channel vector, so the per-channel constants stay
// in registers); results go straight out with coalesced 16-byte stores.
//   mode 0: y = act(x*scale+shift) (+res)                  (bn_act_apply;   in: x [, res]   out: y)
//   mode 1: red += sum dz, sum dz*xhat                     (bn_bwd_reduce;  in: x, da)
//   mode 2: dx = scale*(dz - mean(dz) - xhat*mean(dz*xhat)) (bn_bwd_apply;   in: x, da      out: dx)
// ---------------------------------------------------------------------------------------------
constexpr int kSConsumers = 480;      // 15 consumer warps + 1 producer warp = 512 threads -> 128 registers each
constexpr int kSThreads = kSConsumers + 32;
constexpr int kSMaxStages = 8;

struct StreamArgs {
  long long M; int C, cv, rpb;          // rows, channels, channel vectors, rows per consumer pass
  int U, rows_per_stage, n_stages, n_in;
  uint32_t slot_bytes;                  // bytes reserved per tensor per stage (128-byte multiple)
  const void* in0; const void* in1; void* out;
  const float* scale; const float* shift; const float* mean; const float* rstd;
  int act, frozen;
  double* red; float* dgamma; float* dbeta;
  float drop_rate; uint64_t seed; const long long* seed_dev;      // Dropout after the activation (deeplabv3p.py:410)
  int has_fin; dlb_bn_fin fin;          // kMode 0: scale / shift are finalised here from the batch statistics
};

__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// kDrop: the Dropout instance (one layer of the graph) -- a template parameter so that the hash and its registers
// stay out of the other ~120 launches per step (as a run-time flag it cost them 4 % each)
template <typename T, int kMode, bool kDrop>
__global__ void __launch_bounds__(kSThreads, 1) bn_stream_kernel(const StreamArgs a) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  using P = typename BP2<T>::t;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t stage_bytes = a.slot_bytes * a.n_in;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_raw + static_cast<size_t>(a.n_stages) * stage_bytes);
  uint64_t* empty = full + kSMaxStages;
  if (tid == 0) {
    for (int i = 0; i < a.n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kSConsumers / 32); }
    mbar_fence_init();
  }
  pdl_wait();     // barrier init above overlaps the preceding kernel's tail
  if (kMode == 2 && blockIdx.x == 0) {
    for (int c = tid; c < a.C; c += kSThreads) {
      if (a.dbeta) a.dbeta[c] = static_cast<float>(a.red[c]);
      if (a.dgamma) a.dgamma[c] = static_cast<float>(a.red[a.C + c]);
    }
  }
  __syncthreads();
  const long long n_blk = (a.M + a.rows_per_stage - 1) / a.rows_per_stage;
  const size_t row_bytes = static_cast<size_t>(a.C) * sizeof(T);

  if (warp == kSConsumers / 32) {
    if (lane == 0) {
      // ===================== producer =====================
      int st = 0; uint32_t ph = 0;
      for (long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
        const long long r0 = blk * a.rows_per_stage;
        const long long rows = a.M - r0 < a.rows_per_stage ? a.M - r0 : a.rows_per_stage;
        const uint32_t bytes = static_cast<uint32_t>(rows * row_bytes);
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], bytes * a.n_in);
        const uint32_t dst = smem_u32(s_raw) + st * stage_bytes;
        bulk_g2s(dst, static_cast<const uint8_t*>(a.in0) + r0 * row_bytes, bytes, &full[st]);
        if (a.n_in > 1) bulk_g2s(dst + a.slot_bytes, static_cast<const uint8_t*>(a.in1) + r0 * row_bytes, bytes, &full[st]);
        if (++st == a.n_stages) { st = 0; ph ^= 1; }
      }
    }
    return;
  }

  // ===================== consumers =====================
  const bool active = tid < a.rpb * a.cv;
  const int r_in = active ? tid / a.cv : 0;
  const int c0 = active ? (tid - r_in * a.cv) * 8 : 0;
  // per-channel constants of this thread's 8 channels
  float sc[8], sh[8], k1[8], k2[8];
  P sc2[4], sh2[4], mu2[4], rs2[4];
  const float invM = 1.f / static_cast<float>(a.M);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (kMode == 0 && a.has_fin) {
      // consumer-side finalisation: every thread derives its 8 channels; the first row of CTA 0 publishes them
      bn_fin_channel(a.fin, c0 + k, blockIdx.x == 0 && active && r_in == 0, sc[k], sh[k]);
    } else {
      sc[k] = a.scale ? a.scale[c0 + k] : 1.f;
      sh[k] = a.scale ? a.shift[c0 + k] : 0.f;
    }
    k1[k] = (kMode == 2 && !a.frozen) ? static_cast<float>(a.red[c0 + k]) * invM : 0.f;
    k2[k] = (kMode == 2 && !a.frozen) ? static_cast<float>(a.red[a.C + c0 + k]) * invM : 0.f;
  }
  if (kMode != 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sc2[k] = BP2<T>::pack(sc[2 * k], sc[2 * k + 1]);
      sh2[k] = BP2<T>::pack(sh[2 * k], sh[2 * k + 1]);
      mu2[k] = BP2<T>::pack(a.mean[c0 + 2 * k], a.mean[c0 + 2 * k + 1]);
      rs2[k] = BP2<T>::pack(a.rstd[c0 + 2 * k], a.rstd[c0 + 2 * k + 1]);
    }
  }
  const P zero2 = BP2<T>::bcast(0.f), six2 = BP2<T>::bcast(6.f);
  // inverted dropout: the keep mask is a counter-based hash of (seed, device iteration counter, element index), so the
  // forward pass, both backward passes and a replayed CUDA graph regenerate the same mask without storing it
  constexpr bool drop = kDrop;
  const float keep_inv = drop ? 1.f / (1.f - a.drop_rate) : 1.f;
  const uint64_t dseed = a.seed + (a.seed_dev ? static_cast<uint64_t>(*a.seed_dev) * 0x632BE59BD9B4E019ull : 0ull);
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  T* out = reinterpret_cast<T*>(a.out);
  const uint32_t t_off = static_cast<uint32_t>((r_in * a.C + c0) * sizeof(T));      // this thread inside one pass
  const uint32_t pass_bytes = static_cast<uint32_t>(a.rpb * row_bytes);

  int st = 0; uint32_t ph = 0;
  for (long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
    mbar_wait(&full[st], ph);
    if (active) {
      const uint8_t* s_in0 = s_raw + st * stage_bytes + t_off;
      const uint8_t* s_in1 = s_in0 + a.slot_bytes;
      const long long r_base = blk * a.rows_per_stage + r_in;
      constexpr int UB = 4;
      for (int u0 = 0; u0 < a.U; u0 += UB) {
        BV8<T> xv[UB], gv[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          if (u0 + u < a.U) {
            const uint4 q = *reinterpret_cast<const uint4*>(s_in0 + (u0 + u) * pass_bytes);
            xv[u] = *reinterpret_cast<const BV8<T>*>(&q);
            if (kMode != 0 || a.n_in > 1) {
              const uint4 q1 = *reinterpret_cast<const uint4*>(s_in1 + (u0 + u) * pass_bytes);
              gv[u] = *reinterpret_cast<const BV8<T>*>(&q1);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const long long r = r_base + static_cast<long long>(u0 + u) * a.rpb;
          if (u0 + u >= a.U || r >= a.M) continue;
          if (kMode == 0) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = BP2<T>::unpack(xv[u].h[k]);
              o[2 * k] = apply_act(fmaf(f.x, sc[2 * k], sh[2 * k]), a.act);
              o[2 * k + 1] = apply_act(fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]), a.act);
              if (drop) {
                const uint64_t e = static_cast<uint64_t>(r * a.C + c0 + 2 * k);
                o[2 * k] = hash_uniform(dseed, e) >= a.drop_rate ? o[2 * k] * keep_inv : 0.f;
                o[2 * k + 1] = hash_uniform(dseed, e + 1) >= a.drop_rate ? o[2 * k + 1] * keep_inv : 0.f;
              }
              if (a.n_in > 1) {
                const float2 g = BP2<T>::unpack(gv[u].h[k]);
                o[2 * k] += g.x; o[2 * k + 1] += g.y;
              }
            }
            Vec8<T>::st(out + r * a.C + c0, o);
          } else {
            BV8<T> ov;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              P dz = gv[u].h[k];
              if (a.act != DLB_ACT_NONE) {
                const P z = __hfma2(xv[u].h[k], sc2[k], sh2[k]);
                P m = __hgt2(z, zero2);
                if (a.act == DLB_ACT_RELU6) m = __hmul2(m, __hlt2(z, six2));
                dz = __hmul2(dz, m);
              }
              if (drop) {
                const uint64_t e = static_cast<uint64_t>(r * a.C + c0 + 2 * k);
                const float2 d = BP2<T>::unpack(dz);
                dz = BP2<T>::pack(hash_uniform(dseed, e) >= a.drop_rate ? d.x * keep_inv : 0.f,
                                  hash_uniform(dseed, e + 1) >= a.drop_rate ? d.y * keep_inv : 0.f);
              }
              const P xh2 = __hmul2(__hsub2(xv[u].h[k], mu2[k]), rs2[k]);
              if (kMode == 1) {
                const float2 f1 = BP2<T>::unpack(dz), f2 = BP2<T>::unpack(__hmul2(dz, xh2));
                s1[2 * k] += f1.x; s1[2 * k + 1] += f1.y; s2[2 * k] += f2.x; s2[2 * k + 1] += f2.y;
              } else {
                const float2 dzf = BP2<T>::unpack(dz), xh = BP2<T>::unpack(xh2);
                ov.h[k] = BP2<T>::pack(sc[2 * k] * (dzf.x - k1[2 * k] - xh.x * k2[2 * k]),
                                       sc[2 * k + 1] * (dzf.y - k1[2 * k + 1] - xh.y * k2[2 * k + 1]));
              }
            }
            if (kMode == 2) stg_bh8(out + r * a.C + c0, ov);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (++st == a.n_stages) { st = 0; ph ^= 1; }
  }
  if (kMode == 1) {
    // per-row-group partials -> stage memory (all stages are drained) -> one fp64 atomic per channel and CTA
    asm volatile("bar.sync 1, %0;" ::"n"(kSConsumers) : "memory");
    float* s_red = reinterpret_cast<float*>(s_raw);       // [rpb][2C] = 32 KB
    if (active) {
      float* row = s_red + static_cast<size_t>(r_in) * 2 * a.C;
      *reinterpret_cast<float4*>(row + c0) = make_float4(s1[0], s1[1], s1[2], s1[3]);
      *reinterpret_cast<float4*>(row + c0 + 4) = make_float4(s1[4], s1[5], s1[6], s1[7]);
      *reinterpret_cast<float4*>(row + a.C + c0) = make_float4(s2[0], s2[1], s2[2], s2[3]);
      *reinterpret_cast<float4*>(row + a.C + c0 + 4) = make_float4(s2[4], s2[5], s2[6], s2[7]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kSConsumers) : "memory");
    for (int i = tid; i < 2 * a.C; i += kSConsumers) {
      float t = 0.f;
      for (int r = 0; r < a.rpb; ++r) t += s_red[static_cast<size_t>(r) * 2 * a.C + i];
      atomicAdd(&a.red[i], static_cast<double>(t));
    }
  }
}

// returns 0 if the streaming kernel was launched, 1 if the shape is not eligible (caller falls back), <0 on error
template <int kMode>
static int launch_bn_stream(StreamArgs a, int dtype, cudaStream_t st) {
  if (dtype != DLB_F16 && dtype != DLB_BF16) return 1;
  if (a.C % 8 != 0 || a.C / 8 > kSConsumers / 2 || a.M < 1024) return 1;
  const uintptr_t al = reinterpret_cast<uintptr_t>(a.in0) | reinterpret_cast<uintptr_t>(a.in1) | reinterpret_cast<uintptr_t>(a.out);
  if (al & 15) return 1;
  a.cv = a.C / 8;
  a.rpb = kSConsumers / a.cv;
  const long long pass_bytes = static_cast<long long>(a.rpb) * a.C * 2;
  static const long long stage_target = [] { const char* e = getenv("DLB_BN_STAGE_KB"); return (e ? atoll(e) : 32) * 1024; }();
  a.U = static_cast<int>(stage_target / pass_bytes);
  if (a.U < 1) a.U = 1;
  a.rows_per_stage = a.rpb * a.U;
  a.slot_bytes = static_cast<uint32_t>((static_cast<long long>(a.rows_per_stage) * a.C * 2 + 127) / 128 * 128);
  const long long budget = 200 * 1024;
  a.n_stages = static_cast<int>(budget / (static_cast<long long>(a.slot_bytes) * a.n_in));
  if (a.n_stages > kSMaxStages) a.n_stages = kSMaxStages;
  if (a.n_stages < 2) return 1;
  size_t smem = static_cast<size_t>(a.n_stages) * a.slot_bytes * a.n_in + 2 * kSMaxStages * 8;
  if (kMode == 1 && smem < 32768 + 256) smem = 32768 + 256;
  const long long n_blk = (a.M + a.rows_per_stage - 1) / a.rows_per_stage;
  const int grid = static_cast<int>(n_blk < num_sms() ? n_blk : num_sms());
  if (dtype == DLB_F16) {
    if (a.drop_rate > 0.f) {
      DLB_CUDA(cudaFuncSetAttribute(bn_stream_kernel<__half, kMode, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_k(bn_stream_kernel<__half, kMode, true>, grid, kSThreads, smem, st, a);
    } else {
      DLB_CUDA(cudaFuncSetAttribute(bn_stream_kernel<__half, kMode, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_k(bn_stream_kernel<__half, kMode, false>, grid, kSThreads, smem, st, a);
    }
  } else {
    if (a.drop_rate > 0.f) {
      DLB_CUDA(cudaFuncSetAttribute(bn_stream_kernel<__nv_bfloat16, kMode, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_k(bn_stream_kernel<__nv_bfloat16, kMode, true>, grid, kSThreads, smem, st, a);
    } else {
      DLB_CUDA(cudaFuncSetAttribute(bn_stream_kernel<__nv_bfloat16, kMode, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_k(bn_stream_kernel<__nv_bfloat16, kMode, false>, grid, kSThreads, smem, st, a);
    }
  }
  g_launches++;
  return check_launch("bn_stream_kernel");
}

static int grid_for(long long n, int threads, int per_sm) {
  long long blocks = (n + threads - 1) / threads;
  long long cap = static_cast<long long>(num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace dlb

using namespace dlb;

extern "C" int dlb_bn_finalize(int C, double count, double* sum, double* sqs, const float* gamma, const float* beta,
                               float eps, float momentum, float* moving_mean, float* moving_var, float* scale,
                               float* shift, float* mean, float* rstd, int reset, void* stream) {
  DLB_REQUIRE(C > 0 && sum && sqs && gamma && beta && scale && shift, "bn_finalize: null pointer");
  DLB_REQUIRE(count > 0, "bn_finalize: count must be positive");
  const dlb_bn_fin f{sum, sqs, gamma, beta, eps, momentum, count, moving_mean, moving_var, scale, shift, mean, rstd};
  launch_k(bn_finalize_kernel, (C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), C, f, sum, sqs, reset);
  g_launches++;
  return check_launch("bn_finalize_kernel");
}

extern "C" int dlb_bn_fold(int C, const float* gamma, const float* beta, const float* moving_mean,
                           const float* moving_var, float eps, float* scale, float* shift, void* stream) {
  DLB_REQUIRE(C > 0 && gamma && beta && moving_mean && moving_var && scale && shift, "bn_fold: null pointer");
  launch_k(bn_fold_kernel, (C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), C, gamma, beta, moving_mean,
                                                                                 moving_var, eps, scale, shift);
  g_launches++;
  return check_launch("bn_fold_kernel");
}

int dlb::check_bn_fin(const dlb_bn_fin* f, const char* who) {
  DLB_REQUIRE(f->sum && f->sqs && f->gamma && f->beta && f->scale && f->shift, "%s: dlb_bn_fin with a null pointer", who);
  DLB_REQUIRE(f->count > 0, "%s: dlb_bn_fin.count must be positive", who);
  DLB_REQUIRE((f->moving_mean == nullptr) == (f->moving_var == nullptr), "%s: dlb_bn_fin needs both moving statistics or none", who);
  return DLB_OK;
}

int dlb::bn_fin_standalone(int C, const dlb_bn_fin* f, void* stream) {
  return dlb_bn_finalize(C, f->count, const_cast<double*>(f->sum), const_cast<double*>(f->sqs), f->gamma, f->beta, f->eps,
                         f->momentum, f->moving_mean, f->moving_var, f->scale, f->shift, f->mean, f->rstd, 0, stream);
}

extern "C" int dlb_bn_act_apply(const dlb_bn_apply_params* p, void* stream) {
  DLB_REQUIRE(p && p->x && p->y, "bn_act_apply: null pointer");
  DLB_REQUIRE(p->C % 8 == 0, "bn_act_apply: C must be a multiple of 8 (C=%d)", p->C);
  DLB_REQUIRE(p->C / 8 <= 256, "bn_act_apply: C <= 2048");
  const float* scale = p->scale; const float* shift = p->shift;
  if (p->fin) {
    const int rc = check_bn_fin(p->fin, "bn_act_apply");
    if (rc) return rc;
    scale = p->fin->scale; shift = p->fin->shift;
  }
  cudaStream_t st0 = static_cast<cudaStream_t>(stream);
  if (g_bn_stream) {
    StreamArgs sa{};
    sa.M = p->M; sa.C = p->C; sa.in0 = p->x; sa.in1 = p->res; sa.out = p->y; sa.n_in = p->res ? 2 : 1;
    sa.scale = scale; sa.shift = shift; sa.act = p->act;
    sa.drop_rate = p->drop_rate; sa.seed = p->drop_seed; sa.seed_dev = reinterpret_cast<const long long*>(p->drop_seed_dev);
    if (p->fin) { sa.has_fin = 1; sa.fin = *p->fin; }
    const int rc = launch_bn_stream<0>(sa, p->dtype, st0);
    if (rc <= 0) return rc;
  }
  if (p->fin) {       // shapes the streaming kernel does not take: finalise with the stand-alone kernel first
    const int rc = bn_fin_standalone(p->C, p->fin, stream);
    if (rc) return rc;
  }
  ApplyArgs a{p->M * p->C / 8, p->C, p->x, p->y, p->res, scale, shift, p->act, p->drop_rate, p->drop_seed,
              reinterpret_cast<const long long*>(p->drop_seed_dev)};
  const int rpb_ = 256 / (p->C / 8);
  long long blocks_ = (p->M + static_cast<long long>(rpb_) * 4 - 1) / (static_cast<long long>(rpb_) * 4);
  const long long cap_ = static_cast<long long>(num_sms()) * 8;
  const int grid = static_cast<int>(blocks_ < cap_ ? (blocks_ > 0 ? blocks_ : 1) : cap_);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dtype == DLB_F16) launch_k(bn_apply_kernel<__half>, grid, 256, 0, st, a);
  else if (p->dtype == DLB_BF16) launch_k(bn_apply_kernel<__nv_bfloat16>, grid, 256, 0, st, a);
  else launch_k(bn_apply_kernel<float>, grid, 256, 0, st, a);
  g_launches++;
  return check_launch("bn_apply_kernel");
}

static int fill_bwd(const dlb_bn_bwd_params* p, BwdArgs* a) {
  DLB_REQUIRE(p && p->x && p->da && p->scale && p->shift && p->mean && p->rstd && p->red, "bn_bwd: null pointer");
  DLB_REQUIRE(p->C % 8 == 0 && p->C / 8 <= 256, "bn_bwd: C must be a multiple of 8 and <= 2048 (C=%d)", p->C);
  a->M = p->M; a->C = p->C; a->x = p->x; a->da = p->da; a->dx = p->dx;
  a->scale = p->scale; a->shift = p->shift; a->mean = p->mean; a->rstd = p->rstd; a->act = p->act;
  a->red = p->red; a->dgamma = p->dgamma; a->dbeta = p->dbeta; a->drop_rate = p->drop_rate; a->seed = p->drop_seed;
  a->seed_dev = reinterpret_cast<const long long*>(p->drop_seed_dev);
  a->frozen = p->frozen_stats; a->cv = p->C / 8; a->rpb = 256 / a->cv;
  return DLB_OK;
}

extern "C" int dlb_bn_bwd_reduce(const dlb_bn_bwd_params* p, void* stream) {
  BwdArgs a{};
  int rc = fill_bwd(p, &a);
  if (rc) return rc;
  if (g_bn_stream) {
    StreamArgs sa{};
    sa.M = p->M; sa.C = p->C; sa.in0 = p->x; sa.in1 = p->da; sa.out = nullptr; sa.n_in = 2;
    sa.scale = p->scale; sa.shift = p->shift; sa.mean = p->mean; sa.rstd = p->rstd; sa.act = p->act; sa.red = p->red;
    sa.drop_rate = p->drop_rate; sa.seed = p->drop_seed; sa.seed_dev = reinterpret_cast<const long long*>(p->drop_seed_dev);
    rc = launch_bn_stream<1>(sa, p->dtype, static_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
  long long blocks = (a.M + a.rpb * 4LL - 1) / (a.rpb * 4LL);
  long long cap = static_cast<long long>(num_sms()) * 6;
  const int grid = static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
  const size_t smem = 2 * p->C * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dtype == DLB_F16 && p->drop_rate <= 0.f)
    launch_k(bn_bwd_reduce_h_kernel<__half>, grid, 256, static_cast<size_t>(a.rpb) * smem, st, a);       // [rpb][2C] <= 16 KB
  else if (p->dtype == DLB_BF16 && p->drop_rate <= 0.f)
    launch_k(bn_bwd_reduce_h_kernel<__nv_bfloat16>, grid, 256, static_cast<size_t>(a.rpb) * smem, st, a);
  else if (p->dtype == DLB_F16) launch_k(bn_bwd_reduce_kernel<__half>, grid, 256, smem, st, a);
  else if (p->dtype == DLB_BF16) launch_k(bn_bwd_reduce_kernel<__nv_bfloat16>, grid, 256, smem, st, a);
  else launch_k(bn_bwd_reduce_kernel<float>, grid, 256, smem, st, a);
  g_launches++;
  return check_launch("bn_bwd_reduce_kernel");
}

extern "C" int dlb_bn_bwd_apply(const dlb_bn_bwd_params* p, void* stream) {
  BwdArgs a{};
  int rc = fill_bwd(p, &a);
  if (rc) return rc;
  DLB_REQUIRE(p->dx, "bn_bwd_apply: dx is null");
  if (g_bn_stream) {
    StreamArgs sa{};
    sa.M = p->M; sa.C = p->C; sa.in0 = p->x; sa.in1 = p->da; sa.out = p->dx; sa.n_in = 2;
    sa.scale = p->scale; sa.shift = p->shift; sa.mean = p->mean; sa.rstd = p->rstd; sa.act = p->act; sa.red = p->red;
    sa.frozen = p->frozen_stats; sa.dgamma = p->dgamma; sa.dbeta = p->dbeta;
    sa.drop_rate = p->drop_rate; sa.seed = p->drop_seed; sa.seed_dev = reinterpret_cast<const long long*>(p->drop_seed_dev);
    rc = launch_bn_stream<2>(sa, p->dtype, static_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
  long long blocks = (a.M + a.rpb * 4LL - 1) / (a.rpb * 4LL);
  long long cap = static_cast<long long>(num_sms()) * 6;
  const int grid = static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dtype == DLB_F16 && p->drop_rate <= 0.f) launch_k(bn_bwd_apply_h_kernel<__half>, grid, 256, 0, st, a);
  else if (p->dtype == DLB_BF16 && p->drop_rate <= 0.f) launch_k(bn_bwd_apply_h_kernel<__nv_bfloat16>, grid, 256, 0, st, a);
  else if (p->dtype == DLB_F16) launch_k(bn_bwd_apply_kernel<__half>, grid, 256, 0, st, a);
  else if (p->dtype == DLB_BF16) launch_k(bn_bwd_apply_kernel<__nv_bfloat16>, grid, 256, 0, st, a);
  else launch_k(bn_bwd_apply_kernel<float>, grid, 256, 0, st, a);
  g_launches++;
  return check_launch("bn_bwd_apply_kernel");
}

extern "C" int dlb_global_avgpool_fwd(int B, int HW, int C, int dtype, const void* x, const float* in_scale,
                                      const float* in_shift, int in_act, float* out, void* stream) {
  DLB_REQUIRE(x && out && B > 0 && HW > 0, "global_avgpool_fwd: bad arguments");
  DLB_REQUIRE(C % 8 == 0 && C / 8 <= 256, "global_avgpool_fwd: C must be a multiple of 8 and <= 2048");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DLB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * B * C, st));
  const int cv = C / 8, rpb = 256 / cv;
  int splits = (2 * num_sms() + B - 1) / B;
  const int max_splits = (HW + rpb - 1) / rpb;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const size_t smem = C * sizeof(float);
  if (dtype == DLB_F16) launch_k(avgpool_fwd_kernel<__half>, B * splits, 256, smem, st, HW, C, cv, rpb, splits, (const __half*)x, in_scale, in_shift, in_act, out);
  else if (dtype == DLB_BF16) launch_k(avgpool_fwd_kernel<__nv_bfloat16>, B * splits, 256, smem, st, HW, C, cv, rpb, splits, (const __nv_bfloat16*)x, in_scale, in_shift, in_act, out);
  else launch_k(avgpool_fwd_kernel<float>, B * splits, 256, smem, st, HW, C, cv, rpb, splits, (const float*)x, in_scale, in_shift, in_act, out);
  g_launches++;
  return check_launch("avgpool_fwd_kernel");
}

extern "C" int dlb_global_avgpool_bwd(int B, int HW, int C, int dtype, const float* dout, void* dx, int accumulate,
                                      void* stream) {
  DLB_REQUIRE(dout && dx && C % 8 == 0, "global_avgpool_bwd: bad arguments");
  const long long nvec = static_cast<long long>(B) * HW * C / 8;
  const int grid = grid_for(nvec, 256, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == DLB_F16) launch_k(avgpool_bwd_kernel<__half>, grid, 256, 0, st, nvec, HW, C, dout, (__half*)dx, accumulate);
  else if (dtype == DLB_BF16) launch_k(avgpool_bwd_kernel<__nv_bfloat16>, grid, 256, 0, st, nvec, HW, C, dout, (__nv_bfloat16*)dx, accumulate);
  else launch_k(avgpool_bwd_kernel<float>, grid, 256, 0, st, nvec, HW, C, dout, (float*)dx, accumulate);
  g_launches++;
  return check_launch("avgpool_bwd_kernel");
}

extern "C" int dlb_small_gemm(int M, int N, int K, const float* A, int lda, int transA, const float* B, int ldb,
                              int transB, float* C, int ldc, float alpha, float beta, void* stream) {
  DLB_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "small_gemm: bad arguments");
  if (!transA && transB && K <= 32 * kNtKP && M <= 64) {
    launch_k(small_gemm_nt_kernel, (N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream), M, N, K, A, lda, B, ldb, C, ldc,
             alpha, beta);
    g_launches++;
    return check_launch("small_gemm_nt_kernel");
  }
  dim3 grid((N + 127) / 128, M);
  launch_k(small_gemm_kernel, grid, 128, 0, static_cast<cudaStream_t>(stream), M, N, K, A, lda, transA, B, ldb, transB, C,
                                                                        ldc, alpha, beta);
  g_launches++;
  return check_launch("small_gemm_kernel");
}

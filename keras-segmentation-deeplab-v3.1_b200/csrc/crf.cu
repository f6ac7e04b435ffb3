// Dense-CRF mean-field inference on the permutohedral lattice, sm_100a.
//
// Replaces the pydensecrf call chain of the reference's post-process (utils.py:74-91):
//   DenseCRF2D(W, H, M) ; setUnaryEnergy(U) ; addPairwiseGaussian(sxy, compat) ;
//   addPairwiseBilateral(sxy, srgb, rgbim, compat) ; inference(iters)
// and reproduces densecrf's lattice itself (same embedding, rounding, rank, barycentric weights, blur order
// 0..d, alpha, DIAG_KERNEL + NORMALIZE_SYMMETRIC, Potts compatibility) -- an exact Gaussian would *not* match the
// lattice approximation within the 1e-2 budget.  Differences to the CPU restatement are float
// summation order only.
//
// GPU formulation (all HBM/L2-bandwidth-bound integer/float gather work, no tensor cores):
//   build (once per image and kernel):
//     embed      : one thread per pixel -> d+1 lattice keys packed into 64 bits -> lock-free hash insert (atomicCAS)
//     compact    : hash slots -> dense lattice indices; per-vertex incidence counts -> exclusive scan -> CSR
//     neighbours : +-1 keys along each of the d+1 axes looked up once
//   per mean-field iteration and kernel: splat = CSR *gather* (no atomics, coalesced over the label dimension),
//     d+1 blur passes (ping-pong), slice fused with normalisation, Potts weight, unary add and the softmax.
//   Values are pixel-/vertex-major with the label dimension padded to a multiple of 4 floats (16-byte vectors).
#include <atomic>

#include "common.cuh"

namespace dlb {

extern std::atomic<long long> g_launches;

constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

// 31 bits per coordinate for d = 2, 12 for d = 5: the top bits stay clear so no key equals the empty marker
template <int D> struct KeyBits { static constexpr int bits = D == 2 ? 31 : 64 / D; };

template <int D>
__device__ __forceinline__ unsigned long long pack_key(const int* k) {
  constexpr int B = KeyBits<D>::bits;
  constexpr unsigned long long mask = (B == 64) ? ~0ull : ((1ull << B) - 1ull);
  unsigned long long r = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) r |= (static_cast<unsigned long long>(static_cast<long long>(k[i])) & mask) << (i * B);
  return r;
}
template <int D>
__device__ __forceinline__ void unpack_key(unsigned long long key, int* k) {
  constexpr int B = KeyBits<D>::bits;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    long long v = static_cast<long long>(key << (64 - (i + 1) * B)) >> (64 - B);   // sign-extend field i
    k[i] = static_cast<int>(v);
  }
}
__device__ __forceinline__ unsigned int hash_key64(unsigned long long k, unsigned int cap) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return static_cast<unsigned int>(k % cap);
}
__device__ __forceinline__ int hash_insert(unsigned long long* table, unsigned int cap, unsigned long long key) {
  unsigned int h = hash_key64(key, cap);
  for (;;) {
    const unsigned long long old = atomicCAS(&table[h], kEmptyKey, key);
    if (old == kEmptyKey || old == key) return static_cast<int>(h);
    if (++h == cap) h = 0;
  }
}
__device__ __forceinline__ int hash_lookup(const unsigned long long* table, unsigned int cap, unsigned long long key) {
  unsigned int h = hash_key64(key, cap);
  for (;;) {
    const unsigned long long cur = table[h];
    if (cur == key) return static_cast<int>(h);
    if (cur == kEmptyKey) return -1;
    if (++h == cap) h = 0;
  }
}

struct Lattice {
  int d;
  unsigned int cap;
  unsigned long long* table;   // [cap]
  int* slot_idx;               // [cap] dense index of a slot
  unsigned long long* keys;    // [nv_max] key of dense vertex
  int* offset;                 // [N*(d+1)] slot (after embed) -> dense index (after relabel)
  float* bary;                 // [N*(d+1)]
  int* count;                  // device scalar: number of lattice vertices
  int* deg;                    // [nv_max+1] incidence counts -> exclusive starts (CSR row pointers)
  int* cursor;                 // [nv_max]
  int* ent_pix;                // [N*(d+1)] CSR: pixel of each incidence
  float* ent_w;                // [N*(d+1)] CSR: barycentric weight
  int* nbr;                    // [(d+1), nv_max, 2]
  float* norm;                 // [N]  1/sqrt(K1 + 1e-20)
  int* block_sums;             // scan scratch
  int* overflow;               // key-range overflow flag
};

// ------------------------------------------------------------------------------------------- build kernels
template <int D>
__global__ void __launch_bounds__(256) crf_embed_kernel(int H, int W, const uint8_t* __restrict__ im, float sx, float sy,
                                                        float sr, Lattice L) {
  pdl_prologue();
  const int N = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  float f[D];
  const int x = p % W, y = p / W;
  f[0] = x / sx;
  f[1] = y / sy;
  if (D == 5) {
    f[2] = im[p * 3 + 0] / sr;
    f[3] = im[p * 3 + 1] / sr;
    f[4] = im[p * 3 + 2] / sr;
  }
  // Permutohedral::init of densecrf, float arithmetic without FMA contraction (matches the CPU restatement)
  float elevated[D + 1];
  const float inv_std_dev = sqrtf(2.0f / 3.0f) * (D + 1);
  float sm = 0.f;
#pragma unroll
  for (int j = D; j > 0; --j) {
    const float scale = static_cast<float>(1.0 / sqrt(static_cast<double>((j + 1) * j)) * inv_std_dev);
    const float cf = __fmul_rn(f[j - 1], scale);
    elevated[j] = __fsub_rn(sm, __fmul_rn(static_cast<float>(j), cf));
    sm = __fadd_rn(sm, cf);
  }
  elevated[0] = sm;
  const float down_factor = 1.0f / (D + 1);
  const float up_factor = static_cast<float>(D + 1);
  int rem0[D + 1], rank[D + 1];
  int sum = 0;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const float v = __fmul_rn(down_factor, elevated[i]);
    const float up = __fmul_rn(ceilf(v), up_factor);
    const float down = __fmul_rn(floorf(v), up_factor);
    int rd2;
    if (__fsub_rn(up, elevated[i]) < __fsub_rn(elevated[i], down)) rd2 = static_cast<short>(up);
    else rd2 = static_cast<short>(down);
    rem0[i] = rd2;
    sum += static_cast<int>(__fmul_rn(static_cast<float>(rd2), down_factor));
  }
#pragma unroll
  for (int i = 0; i <= D; ++i) rank[i] = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const float di = __fsub_rn(elevated[i], static_cast<float>(rem0[i]));
#pragma unroll
    for (int j = i + 1; j <= D; ++j) {
      if (di < __fsub_rn(elevated[j], static_cast<float>(rem0[j]))) rank[i]++;
      else rank[j]++;
    }
  }
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    rank[i] += sum;
    if (rank[i] < 0) { rank[i] += D + 1; rem0[i] += D + 1; }
    else if (rank[i] > D) { rank[i] -= D + 1; rem0[i] -= D + 1; }
  }
  float bary[D + 2];
#pragma unroll
  for (int i = 0; i <= D + 1; ++i) bary[i] = 0.f;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const float v = __fmul_rn(__fsub_rn(elevated[i], static_cast<float>(rem0[i])), down_factor);
    // dynamic register indexing avoided: scatter with compares
#pragma unroll
    for (int s = 0; s <= D + 1; ++s) {
      if (s == D - rank[i]) bary[s] = __fadd_rn(bary[s], v);
      if (s == D - rank[i] + 1) bary[s] = __fsub_rn(bary[s], v);
    }
  }
  bary[0] = __fadd_rn(bary[0], __fadd_rn(1.0f, bary[D + 1]));
  constexpr int B = KeyBits<D>::bits;
  const int lim = (B >= 32) ? 0x7FFFFFFF : ((1 << (B - 1)) - 1);
#pragma unroll
  for (int r = 0; r <= D; ++r) {
    int key[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      // canonical[r*(D+1) + rank[i]] = r if rank[i] <= D - r else r - (D+1)
      const int can = (rank[i] <= D - r) ? r : r - (D + 1);
      key[i] = rem0[i] + can;
      if (key[i] > lim || key[i] < -lim - 1) *L.overflow = 1;
    }
    const int slot = hash_insert(L.table, L.cap, pack_key<D>(key));
    L.offset[static_cast<size_t>(p) * (D + 1) + r] = slot;
    L.bary[static_cast<size_t>(p) * (D + 1) + r] = bary[r];
  }
}

__global__ void __launch_bounds__(256) crf_compact_kernel(Lattice L) {
  pdl_prologue();
  const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= L.cap) return;
  const unsigned long long k = L.table[s];
  if (k == kEmptyKey) { L.slot_idx[s] = -1; return; }
  const int idx = atomicAdd(L.count, 1);
  L.slot_idx[s] = idx;
  L.keys[idx] = k;
}

__global__ void __launch_bounds__(256) crf_relabel_kernel(int n_inc, Lattice L) {
  pdl_prologue();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_inc) return;
  const int idx = L.slot_idx[L.offset[e]];
  L.offset[e] = idx;
  atomicAdd(&L.deg[idx], 1);
}

// exclusive scan of deg[0..n) in place, n read from the device (n = *count, padded grid); 3 phases
constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(kScanBlock) scan_phase1(int* data, const int* n_dev, int* block_sums) {
  pdl_prologue();
  __shared__ int s[kScanBlock];
  const int n = *n_dev + 1;
  const int i = blockIdx.x * kScanBlock + threadIdx.x;
  if (blockIdx.x * kScanBlock >= n) return;
  int v = i < n ? data[i] : 0;
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < kScanBlock; o <<= 1) {
    int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < n) data[i] = s[threadIdx.x] - v;   // exclusive
  if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void __launch_bounds__(kScanBlock) scan_phase2(int* block_sums, const int* n_dev) {
  pdl_prologue();
  __shared__ int s[kScanBlock];
  __shared__ int carry;
  const int nb = (*n_dev + 1 + kScanBlock - 1) / kScanBlock;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += kScanBlock) {
    const int i = base + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {
      int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) block_sums[i] = s[threadIdx.x] - v + carry;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry += s[threadIdx.x];
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kScanBlock) scan_phase3(int* data, const int* n_dev, const int* block_sums) {
  pdl_prologue();
  const int n = *n_dev + 1;
  const int i = blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n) data[i] += block_sums[blockIdx.x];
}

__global__ void __launch_bounds__(256) crf_fill_kernel(int n_inc, int dp1, Lattice L) {
  pdl_prologue();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_inc) return;
  const int idx = L.offset[e];
  const int pos = L.deg[idx] + atomicAdd(&L.cursor[idx], 1);
  L.ent_pix[pos] = e / dp1;
  L.ent_w[pos] = L.bary[e];
}

template <int D>
__global__ void __launch_bounds__(256) crf_neighbors_kernel(int nv_max, Lattice L) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *L.count) return;
  int key[D];
  unpack_key<D>(L.keys[i], key);
#pragma unroll
  for (int j = 0; j <= D; ++j) {
    int n1[D], n2[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { n1[k] = key[k] - 1; n2[k] = key[k] + 1; }
    if (j < D) { n1[j] = key[j] + D; n2[j] = key[j] - D; }
    const int s1 = hash_lookup(L.table, L.cap, pack_key<D>(n1));
    const int s2 = hash_lookup(L.table, L.cap, pack_key<D>(n2));
    int* out = L.nbr + (static_cast<size_t>(j) * nv_max + i) * 2;
    out[0] = s1 >= 0 ? L.slot_idx[s1] : -1;
    out[1] = s2 >= 0 ? L.slot_idx[s2] : -1;
  }
}

// ------------------------------------------------------------------------------------------- filter kernels
// values layout: [vertex][VS] floats, VS multiple of 4.  `src` is pixel-major [N][VS]; optional per-pixel scale.
__global__ void __launch_bounds__(256) crf_splat_kernel(int vs4, const float4* __restrict__ src,
                                                        const float* __restrict__ pix_scale, float4* __restrict__ values,
                                                        Lattice L) {
  pdl_prologue();
  const long long total = static_cast<long long>(*L.count) * vs4;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(t / vs4), q = static_cast<int>(t - static_cast<long long>(i) * vs4);
    const int beg = L.deg[i], end = L.deg[i + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = beg; e < end; ++e) {
      const int p = L.ent_pix[e];
      float w = L.ent_w[e];
      if (pix_scale) w *= pix_scale[p];
      const float4 v = src[static_cast<size_t>(p) * vs4 + q];
      acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
    }
    values[t] = acc;
  }
}

__global__ void __launch_bounds__(256) crf_blur_kernel(int vs4, int axis, int nv_max, const float4* __restrict__ in,
                                                       float4* __restrict__ out, Lattice L) {
  pdl_prologue();
  const long long total = static_cast<long long>(*L.count) * vs4;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(t / vs4), q = static_cast<int>(t - static_cast<long long>(i) * vs4);
    const int2 nb = *reinterpret_cast<const int2*>(L.nbr + (static_cast<size_t>(axis) * nv_max + i) * 2);
    float4 v = in[t];
    float4 a = nb.x >= 0 ? in[static_cast<size_t>(nb.x) * vs4 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 b = nb.y >= 0 ? in[static_cast<size_t>(nb.y) * vs4 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
    v.x += 0.5f * (a.x + b.x); v.y += 0.5f * (a.y + b.y); v.z += 0.5f * (a.z + b.z); v.w += 0.5f * (a.w + b.w);
    out[t] = v;
  }
}

// slice of a 4-wide value (norm computation): out[p] = 1/sqrt(alpha * sum_j w_j values[off_j].x + 1e-20)
__global__ void __launch_bounds__(256) crf_slice_norm_kernel(int N, int dp1, float alpha, const float4* __restrict__ values,
                                                             Lattice L) {
  pdl_prologue();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  float acc = 0.f;
  for (int j = 0; j < dp1; ++j) {
    const int o = L.offset[static_cast<size_t>(p) * dp1 + j];
    acc += L.bary[static_cast<size_t>(p) * dp1 + j] * values[o].x * alpha;
  }
  L.norm[p] = 1.0f / sqrtf(acc + 1e-20f);
}

// slice + Potts message: dst[p][k] = base[p][k] + compat * norm[p] * alpha * sum_j w_j values[off_j][k]
// (base = -U for the first kernel term, the running sum for later ones).  If `softmax_out` the result is
// column-normalised exp (expAndNormalize of densecrf) and written to Q, optionally with the arg-max label.
template <int VS>
__global__ void __launch_bounds__(128) crf_slice_kernel(int N, int M, int dp1, float alpha, float compat,
                                                        const float* __restrict__ values, const float* __restrict__ base,
                                                        float* __restrict__ dst, int softmax_out, uint8_t* __restrict__ map_out,
                                                        Lattice L) {
  pdl_prologue();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  float acc[VS];
#pragma unroll
  for (int k = 0; k < VS; ++k) acc[k] = 0.f;
  for (int j = 0; j < dp1; ++j) {
    const int o = L.offset[static_cast<size_t>(p) * dp1 + j];
    const float w = L.bary[static_cast<size_t>(p) * dp1 + j];
    const float4* v = reinterpret_cast<const float4*>(values + static_cast<size_t>(o) * VS);
#pragma unroll
    for (int q = 0; q < VS / 4; ++q) {
      const float4 t = v[q];
      acc[q * 4 + 0] += w * t.x * alpha; acc[q * 4 + 1] += w * t.y * alpha;
      acc[q * 4 + 2] += w * t.z * alpha; acc[q * 4 + 3] += w * t.w * alpha;
    }
  }
  const float s = compat * L.norm[p];
  const float4* b4 = reinterpret_cast<const float4*>(base + static_cast<size_t>(p) * VS);
#pragma unroll
  for (int q = 0; q < VS / 4; ++q) {
    const float4 t = b4[q];
    acc[q * 4 + 0] = t.x + s * acc[q * 4 + 0]; acc[q * 4 + 1] = t.y + s * acc[q * 4 + 1];
    acc[q * 4 + 2] = t.z + s * acc[q * 4 + 2]; acc[q * 4 + 3] = t.w + s * acc[q * 4 + 3];
  }
  if (softmax_out) {
    float mx = acc[0]; int am = 0;
#pragma unroll
    for (int k = 1; k < VS; ++k) if (k < M && acc[k] > mx) { mx = acc[k]; am = k; }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < VS; ++k) { acc[k] = k < M ? expf(acc[k] - mx) : 0.f; sum += acc[k]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int k = 0; k < VS; ++k) acc[k] *= inv;
    if (map_out) map_out[p] = static_cast<uint8_t>(am);
  }
  float4* d4 = reinterpret_cast<float4*>(dst + static_cast<size_t>(p) * VS);
#pragma unroll
  for (int q = 0; q < VS / 4; ++q) d4[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
}

// label-major [M, N] <-> pixel-major [N, VS] transposes through shared memory
__global__ void __launch_bounds__(256) crf_unary_in_kernel(int N, int M, int VS, const float* __restrict__ unary,
                                                           float* __restrict__ negU, float* __restrict__ Q) {
  pdl_prologue();
  // negU[p][k] = -unary[k][p] ; Q = softmax_k(negU)
  extern __shared__ float s[];    // [256][VS+1]
  const int p0 = blockIdx.x * 256;
  for (int k = 0; k < M; ++k) {
    const int p = p0 + threadIdx.x;
    s[threadIdx.x * (VS + 1) + k] = p < N ? -unary[static_cast<size_t>(k) * N + p] : 0.f;
  }
  __syncthreads();
  {
    float* row = &s[threadIdx.x * (VS + 1)];
    float mx = row[0];
    for (int k = 1; k < M; ++k) mx = fmaxf(mx, row[k]);
    float sum = 0.f;
    for (int k = 0; k < M; ++k) sum += expf(row[k] - mx);
    const float inv = 1.f / sum;
    const int p = p0 + threadIdx.x;
    if (p < N) {
      for (int k = 0; k < VS; ++k) {
        negU[static_cast<size_t>(p) * VS + k] = k < M ? row[k] : 0.f;
        Q[static_cast<size_t>(p) * VS + k] = k < M ? expf(row[k] - mx) * inv : 0.f;
      }
    }
  }
}
__global__ void __launch_bounds__(256) crf_q_out_kernel(int N, int M, int VS, const float* __restrict__ Q, float* __restrict__ out) {
  pdl_prologue();
  extern __shared__ float s[];    // [256][VS+1]
  const int p0 = blockIdx.x * 256;
  const int n_here = min(256, N - p0);
  for (int i = threadIdx.x; i < n_here * VS; i += 256) {
    const int pl = i / VS, k = i - pl * VS;
    s[pl * (VS + 1) + k] = Q[static_cast<size_t>(p0) * VS + i];
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < n_here)
    for (int k = 0; k < M; ++k) out[static_cast<size_t>(k) * N + p0 + threadIdx.x] = s[threadIdx.x * (VS + 1) + k];
}
__global__ void fill_ones4_kernel(int N, float4* v) {
  pdl_prologue();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < N) v[p] = make_float4(1.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------- host side
struct WsLayout {
  size_t total = 0;
  size_t take(size_t bytes) { size_t o = total; total += (bytes + 255) / 256 * 256; return o; }
};

struct LatticeOffsets {
  size_t table, slot_idx, keys, offset, bary, count, deg, cursor, ent_pix, ent_w, nbr, norm, block_sums, overflow;
  unsigned int cap; int nv_max;
};

static LatticeOffsets plan_lattice(WsLayout& w, int N, int d) {
  LatticeOffsets o{};
  const size_t inc = static_cast<size_t>(N) * (d + 1);
  o.nv_max = static_cast<int>(inc);
  o.cap = static_cast<unsigned int>(2 * inc + 1);
  o.table = w.take(sizeof(unsigned long long) * o.cap);
  o.slot_idx = w.take(sizeof(int) * o.cap);
  o.keys = w.take(sizeof(unsigned long long) * inc);
  o.offset = w.take(sizeof(int) * inc);
  o.bary = w.take(sizeof(float) * inc);
  o.count = w.take(256);
  o.deg = w.take(sizeof(int) * (inc + 2));
  o.cursor = w.take(sizeof(int) * inc);
  o.ent_pix = w.take(sizeof(int) * inc);
  o.ent_w = w.take(sizeof(float) * inc);
  o.nbr = w.take(sizeof(int) * 2 * inc * (d + 1));
  o.norm = w.take(sizeof(float) * N);
  o.block_sums = w.take(sizeof(int) * ((inc + 2) / kScanBlock + 2));
  o.overflow = w.take(256);
  return o;
}

struct CrfPlan {
  LatticeOffsets g, b;
  size_t negU, Q, tmp, val0, val1;
  int VS;
  size_t total;
};

static CrfPlan plan_crf(const dlb_crf_config* c) {
  CrfPlan p{};
  WsLayout w;
  const int N = c->H * c->W;
  p.VS = (c->M + 3) / 4 * 4;
  p.g = plan_lattice(w, N, 2);
  p.b = plan_lattice(w, N, 5);
  p.negU = w.take(sizeof(float) * N * p.VS);
  p.Q = w.take(sizeof(float) * N * p.VS);
  p.tmp = w.take(sizeof(float) * N * p.VS);
  const size_t nv = static_cast<size_t>(N) * 6;      // worst case: every incidence its own vertex (bilateral)
  p.val0 = w.take(sizeof(float) * nv * p.VS);
  p.val1 = w.take(sizeof(float) * nv * p.VS);
  p.total = w.total;
  return p;
}

static Lattice bind(const LatticeOffsets& o, uint8_t* base, int d) {
  Lattice L{};
  L.d = d; L.cap = o.cap;
  L.table = reinterpret_cast<unsigned long long*>(base + o.table);
  L.slot_idx = reinterpret_cast<int*>(base + o.slot_idx);
  L.keys = reinterpret_cast<unsigned long long*>(base + o.keys);
  L.offset = reinterpret_cast<int*>(base + o.offset);
  L.bary = reinterpret_cast<float*>(base + o.bary);
  L.count = reinterpret_cast<int*>(base + o.count);
  L.deg = reinterpret_cast<int*>(base + o.deg);
  L.cursor = reinterpret_cast<int*>(base + o.cursor);
  L.ent_pix = reinterpret_cast<int*>(base + o.ent_pix);
  L.ent_w = reinterpret_cast<float*>(base + o.ent_w);
  L.nbr = reinterpret_cast<int*>(base + o.nbr);
  L.norm = reinterpret_cast<float*>(base + o.norm);
  L.block_sums = reinterpret_cast<int*>(base + o.block_sums);
  L.overflow = reinterpret_cast<int*>(base + o.overflow);
  return L;
}

static int grid1d(long long n, int threads) { return static_cast<int>((n + threads - 1) / threads); }
static int grid_cap(long long n, int threads) {
  long long b = (n + threads - 1) / threads, cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(b < cap ? (b > 0 ? b : 1) : cap);
}

template <int D>
static int build_lattice(const dlb_crf_config* c, const LatticeOffsets& o, Lattice& L, const uint8_t* image, float sx,
                         float sr, float* val0, float* val1, float* scratch_n4, cudaStream_t st) {
  const int N = c->H * c->W;
  const int n_inc = N * (D + 1);
  DLB_CUDA(cudaMemsetAsync(L.table, 0xFF, sizeof(unsigned long long) * L.cap, st));
  DLB_CUDA(cudaMemsetAsync(L.count, 0, 256, st));
  DLB_CUDA(cudaMemsetAsync(L.overflow, 0, 256, st));
  DLB_CUDA(cudaMemsetAsync(L.deg, 0, sizeof(int) * (static_cast<size_t>(n_inc) + 2), st));
  DLB_CUDA(cudaMemsetAsync(L.cursor, 0, sizeof(int) * static_cast<size_t>(n_inc), st));
  launch_k(crf_embed_kernel<D>, grid1d(N, 256), 256, 0, st, c->H, c->W, image, sx, sx, sr, L);
  launch_k(crf_compact_kernel, grid1d(L.cap, 256), 256, 0, st, L);
  launch_k(crf_relabel_kernel, grid1d(n_inc, 256), 256, 0, st, n_inc, L);
  const int nblk = (n_inc + 1 + kScanBlock - 1) / kScanBlock;
  launch_k(scan_phase1, nblk, kScanBlock, 0, st, L.deg, L.count, L.block_sums);
  launch_k(scan_phase2, 1, kScanBlock, 0, st, L.block_sums, L.count);
  launch_k(scan_phase3, nblk, kScanBlock, 0, st, L.deg, L.count, L.block_sums);
  launch_k(crf_fill_kernel, grid1d(n_inc, 256), 256, 0, st, n_inc, D + 1, L);
  launch_k(crf_neighbors_kernel<D>, grid1d(o.nv_max, 256), 256, 0, st, o.nv_max, L);
  g_launches += 8;
  // norm = 1/sqrt(K 1 + 1e-20): filter a ones vector (value width 4, channel 0)
  float4* ones = reinterpret_cast<float4*>(scratch_n4);   // [N] float4 scratch (the mean-field tmp buffer)
  launch_k(fill_ones4_kernel, grid1d(N, 256), 256, 0, st, N, ones);
  float4* v0 = reinterpret_cast<float4*>(val0);
  float4* v1 = reinterpret_cast<float4*>(val1);
  launch_k(crf_splat_kernel, grid_cap(static_cast<long long>(o.nv_max), 256), 256, 0, st, 1, ones, nullptr, v0, L);
  for (int j = 0; j <= D; ++j) {
    launch_k(crf_blur_kernel, grid_cap(static_cast<long long>(o.nv_max), 256), 256, 0, st, 1, j, o.nv_max, v0, v1, L);
    float4* t = v0; v0 = v1; v1 = t;
  }
  const float alpha = 1.0f / (1.0f + powf(2.0f, -static_cast<float>(D)));
  launch_k(crf_slice_norm_kernel, grid1d(N, 256), 256, 0, st, N, D + 1, alpha, v0, L);
  g_launches += 3 + D + 1;
  return check_launch("crf build");
}

template <int VS>
static int run_meanfield(const dlb_crf_config* c, const CrfPlan& P, uint8_t* base, Lattice& Lg, Lattice& Lb,
                         uint8_t* map_out, cudaStream_t st) {
  const int N = c->H * c->W, M = c->M;
  float* negU = reinterpret_cast<float*>(base + P.negU);
  float* Q = reinterpret_cast<float*>(base + P.Q);
  float* tmp = reinterpret_cast<float*>(base + P.tmp);
  float* val0 = reinterpret_cast<float*>(base + P.val0);
  float* val1 = reinterpret_cast<float*>(base + P.val1);
  const bool use_g = c->compat_gauss != 0.f, use_b = c->compat_bilat != 0.f;
  const int vs4 = VS / 4;
  for (int it = 0; it < c->iters; ++it) {
    const float* basep = negU;
    for (int term = 0; term < 2; ++term) {
      const bool is_g = term == 0;
      if ((is_g && !use_g) || (!is_g && !use_b)) continue;
      Lattice& L = is_g ? Lg : Lb;
      const LatticeOffsets& o = is_g ? P.g : P.b;
      const int D = is_g ? 2 : 5;
      float4* v0 = reinterpret_cast<float4*>(val0);
      float4* v1 = reinterpret_cast<float4*>(val1);
      const long long work = static_cast<long long>(o.nv_max) * vs4;
      launch_k(crf_splat_kernel, grid_cap(work, 256), 256, 0, st, vs4, reinterpret_cast<const float4*>(Q), L.norm, v0, L);
      for (int j = 0; j <= D; ++j) {
        launch_k(crf_blur_kernel, grid_cap(work, 256), 256, 0, st, vs4, j, o.nv_max, v0, v1, L);
        float4* t = v0; v0 = v1; v1 = t;
      }
      const bool last = (!is_g) || !use_b;
      const float alpha = 1.0f / (1.0f + powf(2.0f, -static_cast<float>(D)));
      const bool want_map = last && map_out && it == c->iters - 1;
      launch_k(crf_slice_kernel<VS>, grid1d(N, 128), 128, 0, st, N, M, D + 1, alpha, is_g ? c->compat_gauss : c->compat_bilat,
                                                           reinterpret_cast<const float*>(v0), basep, last ? Q : tmp,
                                                           last ? 1 : 0, want_map ? map_out : nullptr, L);
      basep = tmp;
      g_launches += 3 + D;
    }
  }
  return check_launch("crf meanfield");
}

}  // namespace dlb

using namespace dlb;

extern "C" int64_t dlb_crf_workspace_bytes(const dlb_crf_config* cfg) {
  if (!cfg || cfg->H <= 0 || cfg->W <= 0 || cfg->M <= 0) return 0;
  return static_cast<int64_t>(plan_crf(cfg).total);
}

extern "C" int dlb_crf_inference(const dlb_crf_config* cfg, const float* unary, const uint8_t* image, float* Q_out,
                                 uint8_t* map_out, void* workspace, int64_t workspace_bytes, void* stream) {
  DLB_REQUIRE(cfg && unary && image && Q_out && workspace, "crf_inference: null pointer");
  DLB_REQUIRE(cfg->M >= 1 && cfg->M <= 32, "crf_inference: 1 <= labels <= 32 supported (got %d)", cfg->M);
  DLB_REQUIRE(cfg->compat_gauss != 0.f || cfg->compat_bilat != 0.f, "crf_inference: no pairwise term enabled");
  const CrfPlan P = plan_crf(cfg);
  DLB_REQUIRE(workspace_bytes >= static_cast<int64_t>(P.total), "crf_inference: workspace too small (%lld < %lld)",
              (long long)workspace_bytes, (long long)P.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  const int N = cfg->H * cfg->W;
  Lattice Lg = bind(P.g, base, 2), Lb = bind(P.b, base, 5);
  float* val0 = reinterpret_cast<float*>(base + P.val0);
  float* val1 = reinterpret_cast<float*>(base + P.val1);
  int rc;
  if (cfg->compat_gauss != 0.f) {
    rc = build_lattice<2>(cfg, P.g, Lg, image, cfg->sxy_gauss, 1.f, val0, val1, reinterpret_cast<float*>(base + P.tmp), st);
    if (rc) return rc;
  }
  if (cfg->compat_bilat != 0.f) {
    rc = build_lattice<5>(cfg, P.b, Lb, image, cfg->sxy_bilat, cfg->srgb_bilat, val0, val1, reinterpret_cast<float*>(base + P.tmp), st);
    if (rc) return rc;
  }
  float* negU = reinterpret_cast<float*>(base + P.negU);
  float* Q = reinterpret_cast<float*>(base + P.Q);
  const size_t smem = sizeof(float) * 256 * (P.VS + 1);
  launch_k(crf_unary_in_kernel, (N + 255) / 256, 256, smem, st, N, cfg->M, P.VS, unary, negU, Q);
  g_launches++;
  switch (P.VS) {
#define VSCASE(V) case V: rc = run_meanfield<V>(cfg, P, base, Lg, Lb, map_out, st); break;
    VSCASE(4) VSCASE(8) VSCASE(12) VSCASE(16) VSCASE(20) VSCASE(24) VSCASE(28) VSCASE(32)
#undef VSCASE
    default: set_last_error("crf_inference: bad value stride"); return DLB_ERR_INVALID;
  }
  if (rc) return rc;
  launch_k(crf_q_out_kernel, (N + 255) / 256, 256, smem, st, N, cfg->M, P.VS, Q, Q_out);
  g_launches++;
  return check_launch("crf_q_out_kernel");
}

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
  unsigned int cap;            // hash slots per image (2 * inc + 1)
  int inc;                     // incidences per image = N * (d + 1) (also the vertex capacity)
  int shared;                  // 1: one lattice for the whole batch (Gaussian: features depend on H, W, sxy only)
  unsigned long long* table;   // [B][cap]
  int* slot_idx;               // [B][cap]  build: smallest pixel touching the slot; then dense vertex index of the slot
  unsigned long long* keys;    // [B][inc]  key of dense vertex
  int* offset;                 // [B][inc]  slot (after embed) -> dense index (after relabel)
  float* bary;                 // [B][inc]
  int* count;                  // [B][64]   number of lattice vertices
  int* deg;                    // [B][inc+2] incidence counts -> exclusive starts (CSR row pointers)
  int* cursor;                 // [B][inc+2] build scratch: owner flags / their scan, then CSR fill cursors
  int* ent_pix;                // [B][inc]  CSR: pixel of each incidence
  float* ent_w;                // [B][inc]  CSR: barycentric weight
  int* nbr;                    // [B][d+1][inc][2]
  float* norm;                 // [B][N]    1/sqrt(K1 + 1e-20)
  int* block_sums;             // [B][nblk] scan scratch
  int* overflow;               // [B][64]   key-range overflow flag
  int n_pix, nblk;
};

// the arrays of image b (the batch index is blockIdx.y everywhere)
__device__ __forceinline__ Lattice lattice_at(Lattice L, int b) {
  if (L.shared || b == 0) return L;
  const size_t cap = L.cap, inc = L.inc;
  L.table += b * cap; L.slot_idx += b * cap;
  L.keys += b * inc; L.offset += b * inc; L.bary += b * inc; L.ent_pix += b * inc; L.ent_w += b * inc;
  L.count += b * 64; L.overflow += b * 64;
  L.deg += b * (inc + 2); L.cursor += b * (inc + 2);
  L.nbr += b * (2 * inc * (L.d + 1));
  L.norm += static_cast<size_t>(b) * L.n_pix;
  L.block_sums += static_cast<size_t>(b) * L.nblk;
  return L;
}

// ------------------------------------------------------------------------------------------- build kernels
template <int D>
__global__ void __launch_bounds__(256) crf_embed_kernel(int H, int W, const uint8_t* __restrict__ im_all, float sx, float sy,
                                                        float sr, Lattice L0) {
  pdl_prologue();
  const Lattice L = lattice_at(L0, blockIdx.y);
  const int N = H * W;
  const uint8_t* im = im_all + static_cast<size_t>(blockIdx.y) * N * 3;
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
    atomicMin(&L.slot_idx[slot], p);      // the first pixel (raster order) that touches the vertex will number it
    L.offset[static_cast<size_t>(p) * (D + 1) + r] = slot;
    L.bary[static_cast<size_t>(p) * (D + 1) + r] = bary[r];
  }
}

// Dense vertex numbering in RASTER order of the first incident pixel (round 1 numbered vertices in hash-slot order =
// random: the per-vertex gathers of the splat then had no locality at all).  flag[e] = 1 iff incidence e = (p, r)
// belongs to the pixel that owns the vertex; an exclusive scan over e (pixel-major) gives the owner's dense index.
__global__ void __launch_bounds__(256) crf_owner_flag_kernel(int dp1, Lattice L0) {
  pdl_prologue();
  const Lattice L = lattice_at(L0, blockIdx.y);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > L.inc) return;
  L.cursor[e] = (e < L.inc && L.slot_idx[L.offset[e]] == e / dp1) ? 1 : 0;     // [inc] is the scan sentinel
}
__global__ void __launch_bounds__(256) crf_assign_kernel(Lattice L0) {
  pdl_prologue();
  const Lattice L = lattice_at(L0, blockIdx.y);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L.inc) return;
  const int idx = L.cursor[e];
  if (L.cursor[e + 1] != idx) {            // owner incidence
    const int slot = L.offset[e];
    L.keys[idx] = L.table[slot];
    L.slot_idx[slot] = idx;                // (pixel index -> dense index; every ownership test is finished)
  }
  if (e == L.inc - 1) *L.count = L.cursor[L.inc];
}

__global__ void __launch_bounds__(256) crf_relabel_kernel(Lattice L0) {
  pdl_prologue();
  const Lattice L = lattice_at(L0, blockIdx.y);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L.inc) return;
  const int idx = L.slot_idx[L.offset[e]];
  L.offset[e] = idx;
  atomicAdd(&L.deg[idx], 1);
}

// exclusive scan in place, 3 phases, batch = blockIdx.y.  n = n_const if n_dev == nullptr else *n_dev + 1.
constexpr int kScanBlock = 1024;
__device__ __forceinline__ int scan_n(int n_const, const int* n_dev, int b) { return n_dev ? n_dev[b * 64] + 1 : n_const; }
__global__ void __launch_bounds__(kScanBlock) scan_phase1(int* data_all, size_t stride, int n_const, const int* n_dev,
                                                          int* block_sums_all, int nblk) {
  pdl_prologue();
  __shared__ int s[kScanBlock];
  const int b = blockIdx.y;
  int* data = data_all + b * stride;
  int* block_sums = block_sums_all + static_cast<size_t>(b) * nblk;
  const int n = scan_n(n_const, n_dev, b);
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
__global__ void __launch_bounds__(kScanBlock) scan_phase2(int* block_sums_all, int nblk_stride, int n_const, const int* n_dev) {
  pdl_prologue();
  __shared__ int s[kScanBlock];
  __shared__ int carry;
  const int b = blockIdx.y;
  int* block_sums = block_sums_all + static_cast<size_t>(b) * nblk_stride;
  const int nb = (scan_n(n_const, n_dev, b) + kScanBlock - 1) / kScanBlock;
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
__global__ void __launch_bounds__(kScanBlock) scan_phase3(int* data_all, size_t stride, int n_const, const int* n_dev,
                                                          const int* block_sums_all, int nblk) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int n = scan_n(n_const, n_dev, b);
  const int i = blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n) data_all[b * stride + i] += block_sums_all[static_cast<size_t>(b) * nblk + blockIdx.x];
}

__global__ void __launch_bounds__(256) crf_fill_kernel(int dp1, Lattice L0) {
  pdl_prologue();
  const Lattice L = lattice_at(L0, blockIdx.y);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L.inc) return;
  const int idx = L.offset[e];
  const int pos = L.deg[idx] + atomicAdd(&L.cursor[idx], 1);
  L.ent_pix[pos] = e / dp1;
  L.ent_w[pos] = L.bary[e];
}

template <int D>
__global__ void __launch_bounds__(256) crf_neighbors_kernel(Lattice L0) {
  pdl_prologue();
  const Lattice L = lattice_at(L0, blockIdx.y);
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
    int* out = L.nbr + (static_cast<size_t>(j) * L.inc + i) * 2;
    out[0] = s1 >= 0 ? L.slot_idx[s1] : -1;
    out[1] = s2 >= 0 ? L.slot_idx[s2] : -1;
  }
}

// ------------------------------------------------------------------------------------------- filter kernels
// values layout: [image][vertex][VS] floats, VS multiple of 4.  `src` is pixel-major [image][N][VS]; optional per-pixel
// scale.  Vertices are numbered in raster order, so neighbouring threads gather neighbouring pixels.
// One WARP per vertex (grid-stride).  The warp is cut into G = 32 / vs4 groups of vs4 lanes; a group owns one incident
// pixel at a time (lane = float4 column of its row, so a row is read as one contiguous 16 * vs4-byte piece) and the
// groups walk the vertex's incidence list in parallel, four entries per group in flight (the ncu capture of the
// thread-per-column version showed 18-20 warps stalled on the dependent ent_pix -> row chain per issued instruction).
// Weights arrive pre-multiplied by the pixel's symmetric-normalisation factor (crf_scale_weights_kernel).
template <int VS4>
__global__ void __launch_bounds__(256) crf_splat_kernel(const float4* __restrict__ src_all, size_t src_stride,
                                                        float4* __restrict__ values_all, size_t val_stride, Lattice L0) {
  pdl_prologue();
  constexpr int G = 32 / VS4;                      // entry groups per warp
  constexpr int kUnroll = 4;
  const Lattice L = lattice_at(L0, blockIdx.y);
  const float4* src = src_all + blockIdx.y * src_stride;
  float4* values = values_all + blockIdx.y * val_stride;
  const int nv = *L.count;
  const int lane = threadIdx.x & 31;
  const int grp = lane / VS4, q = lane - grp * VS4;
  const bool live = grp < G;                       // (32 % VS4 lanes at the end of the warp have no group)
  const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < nv; i += warps_per_grid) {
    const int beg = L.deg[i], end = L.deg[i + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      for (int e0 = beg + grp; e0 < end; e0 += G * kUnroll) {
        int p[kUnroll]; float w[kUnroll]; float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int e = e0 + u * G;
          p[u] = e < end ? L.ent_pix[e] : -1;
          w[u] = e < end ? L.ent_w[e] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          v[u] = p[u] >= 0 ? src[static_cast<size_t>(p[u]) * VS4 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          acc.x = fmaf(w[u], v[u].x, acc.x); acc.y = fmaf(w[u], v[u].y, acc.y);
          acc.z = fmaf(w[u], v[u].z, acc.z); acc.w = fmaf(w[u], v[u].w, acc.w);
        }
      }
    }
    // fold the groups into group 0
    if constexpr ((VS4 & (VS4 - 1)) == 0) {
#pragma unroll
      for (int o = VS4; o < 32; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
      }
    } else {
      float4 t = acc;
#pragma unroll
      for (int k = 1; k < G; ++k) {
        const int srcl = (lane + k * VS4) & 31;
        t.x += __shfl_sync(0xffffffffu, acc.x, srcl); t.y += __shfl_sync(0xffffffffu, acc.y, srcl);
        t.z += __shfl_sync(0xffffffffu, acc.z, srcl); t.w += __shfl_sync(0xffffffffu, acc.w, srcl);
      }
      acc = t;
    }
    if (lane < VS4) values[static_cast<size_t>(i) * VS4 + lane] = acc;
  }
}

// ent_w[e] *= norm[ent_pix[e]]: the symmetric normalisation of the splat input is a property of the lattice, not of Q
__global__ void __launch_bounds__(256) crf_scale_weights_kernel(Lattice L0) {
  pdl_prologue();
  const Lattice L = lattice_at(L0, blockIdx.y);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L.inc) return;
  L.ent_w[e] *= L.norm[L.ent_pix[e]];
}

__global__ void __launch_bounds__(256) crf_blur_kernel(int vs4, int axis, const float4* __restrict__ in_all,
                                                       float4* __restrict__ out_all, size_t val_stride, Lattice L0) {
  pdl_prologue();
  const Lattice L = lattice_at(L0, blockIdx.y);
  const float4* in = in_all + blockIdx.y * val_stride;
  float4* out = out_all + blockIdx.y * val_stride;
  const long long total = static_cast<long long>(*L.count) * vs4;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(t / vs4), q = static_cast<int>(t - static_cast<long long>(i) * vs4);
    const int2 nb = *reinterpret_cast<const int2*>(L.nbr + (static_cast<size_t>(axis) * L.inc + i) * 2);
    float4 v = in[t];
    float4 a = nb.x >= 0 ? in[static_cast<size_t>(nb.x) * vs4 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 b = nb.y >= 0 ? in[static_cast<size_t>(nb.y) * vs4 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
    v.x += 0.5f * (a.x + b.x); v.y += 0.5f * (a.y + b.y); v.z += 0.5f * (a.z + b.z); v.w += 0.5f * (a.w + b.w);
    out[t] = v;
  }
}

// slice of a 4-wide value (norm computation): out[p] = 1/sqrt(alpha * sum_j w_j values[off_j].x + 1e-20)
__global__ void __launch_bounds__(256) crf_slice_norm_kernel(int dp1, float alpha, const float4* __restrict__ values_all,
                                                             size_t val_stride, Lattice L0) {
  pdl_prologue();
  const Lattice L = lattice_at(L0, blockIdx.y);
  const float4* values = values_all + blockIdx.y * val_stride;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L.n_pix) return;
  float acc = 0.f;
  for (int j = 0; j < dp1; ++j) {
    const int o = L.offset[static_cast<size_t>(p) * dp1 + j];
    acc += L.bary[static_cast<size_t>(p) * dp1 + j] * values[o].x * alpha;
  }
  L.norm[p] = 1.0f / sqrtf(acc + 1e-20f);
}

// Slice of BOTH lattices + Potts messages + unary + softmax in one pass (round 1: one slice kernel per term with the
// running sum round-tripping through HBM):
//   Q[p][k] = softmax_k( -U[p][k] + cg * normG[p] * aG * sum_j wG_j VG[offG_j][k] + cb * normB[p] * aB * sum_j wB_j VB[offB_j][k] )
// Two lanes own one pixel, each VH = vs4 / 2 (rounded up) float4 columns of the VS-wide rows: a vertex row is read as
// two adjacent pieces, there are no idle lanes at 21 labels (vs4 = 6), and the softmax folds over one shuffle.  (An
// eight-lanes-per-pixel version issued 121 warp instructions per 4 pixels and ran at 82 % issue utilisation.)
template <int VH>
__device__ __forceinline__ void slice_term(const Lattice& L, const float4* __restrict__ values, int p, int c0, int vs4,
                                           int dp1, float alpha, float compat, float4 (&v)[VH]) {
  float4 m[VH];
#pragma unroll
  for (int c = 0; c < VH; ++c) m[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int* off = L.offset + static_cast<size_t>(p) * dp1;
  const float* bw = L.bary + static_cast<size_t>(p) * dp1;
  for (int j = 0; j < dp1; ++j) {
    const float4* rowp = values + static_cast<size_t>(off[j]) * vs4 + c0;
    const float wa = bw[j] * alpha;
#pragma unroll
    for (int c = 0; c < VH; ++c) {
      if (c0 + c < vs4) {
        const float4 t = rowp[c];
        m[c].x = fmaf(wa, t.x, m[c].x); m[c].y = fmaf(wa, t.y, m[c].y);
        m[c].z = fmaf(wa, t.z, m[c].z); m[c].w = fmaf(wa, t.w, m[c].w);
      }
    }
  }
  const float s = compat * L.norm[p];
#pragma unroll
  for (int c = 0; c < VH; ++c) {
    v[c].x = fmaf(s, m[c].x, v[c].x); v[c].y = fmaf(s, m[c].y, v[c].y);
    v[c].z = fmaf(s, m[c].z, v[c].z); v[c].w = fmaf(s, m[c].w, v[c].w);
  }
}

template <int VH>
__global__ void __launch_bounds__(256) crf_slice2_kernel(int N, int M, int vs4, int use_g, float alpha_g, float compat_g,
                                                         const float4* __restrict__ val_g_all, size_t val_g_stride, Lattice Lg0,
                                                         int use_b, float alpha_b, float compat_b,
                                                         const float4* __restrict__ val_b_all, size_t val_b_stride, Lattice Lb0,
                                                         const float4* __restrict__ negU_all, float4* __restrict__ Q_all,
                                                         uint8_t* __restrict__ map_all) {
  pdl_prologue();
  const int b = blockIdx.y;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = gid >> 1, half = gid & 1;
  const int c0 = half * VH;                        // first float4 column of this lane
  const bool valid = p < N;
  const size_t row = static_cast<size_t>(b) * N * vs4 + static_cast<size_t>(valid ? p : 0) * vs4 + c0;
  float4 v[VH];
#pragma unroll
  for (int c = 0; c < VH; ++c) v[c] = (valid && c0 + c < vs4) ? negU_all[row + c] : make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    if (use_g) slice_term<VH>(lattice_at(Lg0, b), val_g_all + b * val_g_stride, p, c0, vs4, 3, alpha_g, compat_g, v);
    if (use_b) slice_term<VH>(lattice_at(Lb0, b), val_b_all + b * val_b_stride, p, c0, vs4, 6, alpha_b, compat_b, v);
  }
  // expAndNormalize of densecrf over the M valid labels; first maximum wins the arg-max (sequential `>` scan)
  float mx = -INFINITY; int am = 0;
#pragma unroll
  for (int c = 0; c < VH; ++c) {
    const float cc[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = (c0 + c) * 4 + i;
      if (c0 + c < vs4 && k < M && cc[i] > mx) { mx = cc[i]; am = k; }
    }
  }
  {
    const float mo = __shfl_xor_sync(0xffffffffu, mx, 1);
    const int ao = __shfl_xor_sync(0xffffffffu, am, 1);
    if (mo > mx || (mo == mx && ao < am)) { mx = mo; am = ao; }
  }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < VH; ++c) {
    float cc[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = (c0 + c) * 4 + i;
      cc[i] = (c0 + c < vs4 && k < M) ? expf(cc[i] - mx) : 0.f;
      sum += cc[i];
    }
    v[c] = make_float4(cc[0], cc[1], cc[2], cc[3]);
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = 1.f / sum;
  if (valid) {
#pragma unroll
    for (int c = 0; c < VH; ++c)
      if (c0 + c < vs4) Q_all[row + c] = make_float4(v[c].x * inv, v[c].y * inv, v[c].z * inv, v[c].w * inv);
    if (map_all && half == 0) map_all[static_cast<size_t>(b) * N + p] = static_cast<uint8_t>(am);
  }
}

// label-major [M, N] <-> pixel-major [N, VS] transposes through shared memory (batch = blockIdx.y)
__global__ void __launch_bounds__(256) crf_unary_in_kernel(int N, int M, int VS, const float* __restrict__ unary_all,
                                                           float* __restrict__ negU_all, float* __restrict__ Q_all,
                                                           int from_probs) {
  pdl_prologue();
  // negU[p][k] = -unary[k][p] ; Q = softmax_k(negU).  Label-major reads are coalesced over pixels; the pixel-major
  // rows of the block's 256 pixels form ONE contiguous range, written back from shared memory in order (a thread per
  // row wrote 32 scattered 4-byte pieces per store instruction: 2.0 ms for 8 images at 1024^2 x 21).
  extern __shared__ float s[];    // [256][VS+1] values, then [256] max, [256] 1/sum
  float* s_mx = s + 256 * (VS + 1);
  float* s_inv = s_mx + 256;
  const float* unary = unary_all + static_cast<size_t>(blockIdx.y) * M * N;
  float* negU = negU_all + static_cast<size_t>(blockIdx.y) * N * VS;
  float* Q = Q_all + static_cast<size_t>(blockIdx.y) * N * VS;
  const int p0 = blockIdx.x * 256;
  const int n_here = min(256, N - p0);
  if (from_probs) {
    // pixel-major probabilities [N, M] straight from the network's softmax: -U = log p
    for (int i = threadIdx.x; i < n_here * M; i += 256) {
      const int pl = i / M, k = i - pl * M;
      s[pl * (VS + 1) + k] = logf(fmaxf(unary[static_cast<size_t>(p0) * M + i], 1e-30f));
    }
    __syncthreads();
  } else {
    for (int k = 0; k < M; ++k) {
      const int p = p0 + threadIdx.x;
      s[threadIdx.x * (VS + 1) + k] = p < N ? -unary[static_cast<size_t>(k) * N + p] : 0.f;
    }
  }
  {
    const float* row = &s[threadIdx.x * (VS + 1)];
    float mx = row[0];
    for (int k = 1; k < M; ++k) mx = fmaxf(mx, row[k]);
    float sum = 0.f;
    for (int k = 0; k < M; ++k) sum += expf(row[k] - mx);
    s_mx[threadIdx.x] = mx;
    s_inv[threadIdx.x] = 1.f / sum;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_here * VS; i += 256) {
    const int pl = i / VS, k = i - pl * VS;
    const float v = s[pl * (VS + 1) + k];
    negU[static_cast<size_t>(p0) * VS + i] = k < M ? v : 0.f;
    Q[static_cast<size_t>(p0) * VS + i] = k < M ? expf(v - s_mx[pl]) * s_inv[pl] : 0.f;
  }
}
__global__ void __launch_bounds__(256) crf_q_out_kernel(int N, int M, int VS, const float* __restrict__ Q_all,
                                                        float* __restrict__ out_all) {
  pdl_prologue();
  extern __shared__ float s[];    // [256][VS+1]
  const float* Q = Q_all + static_cast<size_t>(blockIdx.y) * N * VS;
  float* out = out_all + static_cast<size_t>(blockIdx.y) * M * N;
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
  size_t val0, val1;            // ping-pong vertex values [nb][inc * VS] floats (per image even for the shared lattice)
  unsigned int cap; int inc, nblk, nb;   // nb = images that own a lattice (1 when shared)
};

static LatticeOffsets plan_lattice(WsLayout& w, int N, int d, int nb, int batch, int VS) {
  LatticeOffsets o{};
  const size_t inc = static_cast<size_t>(N) * (d + 1);
  o.inc = static_cast<int>(inc);
  o.nb = nb;
  o.cap = static_cast<unsigned int>(2 * inc + 1);
  o.nblk = static_cast<int>((inc + 2 + kScanBlock - 1) / kScanBlock + 1);
  o.table = w.take(sizeof(unsigned long long) * o.cap * nb);
  o.slot_idx = w.take(sizeof(int) * static_cast<size_t>(o.cap) * nb);
  o.keys = w.take(sizeof(unsigned long long) * inc * nb);
  o.offset = w.take(sizeof(int) * inc * nb);
  o.bary = w.take(sizeof(float) * inc * nb);
  o.count = w.take(256 * nb);
  o.deg = w.take(sizeof(int) * (inc + 2) * nb);
  o.cursor = w.take(sizeof(int) * (inc + 2) * nb);
  o.ent_pix = w.take(sizeof(int) * inc * nb);
  o.ent_w = w.take(sizeof(float) * inc * nb);
  o.nbr = w.take(sizeof(int) * 2 * inc * (d + 1) * nb);
  o.norm = w.take(sizeof(float) * N * nb);
  o.block_sums = w.take(sizeof(int) * static_cast<size_t>(o.nblk) * nb);
  o.overflow = w.take(256 * nb);
  // worst case: every incidence its own vertex; both lattices' blurred values are alive in the fused slice
  o.val0 = w.take(sizeof(float) * inc * VS * batch);
  o.val1 = w.take(sizeof(float) * inc * VS * batch);
  return o;
}

struct CrfPlan {
  LatticeOffsets g, b;
  size_t negU, Q;
  int VS;
  size_t total;
};

static CrfPlan plan_crf(const dlb_crf_config* c, int batch) {
  CrfPlan p{};
  WsLayout w;
  const int N = c->H * c->W;
  p.VS = (c->M + 3) / 4 * 4;
  p.g = plan_lattice(w, N, 2, 1, batch, p.VS);          // Gaussian: one lattice for the whole batch
  p.b = plan_lattice(w, N, 5, batch, batch, p.VS);
  p.negU = w.take(sizeof(float) * N * p.VS * batch);
  p.Q = w.take(sizeof(float) * N * p.VS * batch);
  p.total = w.total;
  return p;
}

static Lattice bind(const LatticeOffsets& o, uint8_t* base, int d, int N) {
  Lattice L{};
  L.d = d; L.cap = o.cap; L.inc = o.inc; L.shared = o.nb == 1 ? 1 : 0; L.n_pix = N; L.nblk = o.nblk;
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

static dim3 grid2(long long n, int threads, int nb) { return dim3(static_cast<unsigned>((n + threads - 1) / threads), nb); }
static dim3 grid_cap2(long long n, int threads, int nb) {
  long long b = (n + threads - 1) / threads, cap = static_cast<long long>(num_sms()) * 16;
  return dim3(static_cast<unsigned>(b < cap ? (b > 0 ? b : 1) : cap), nb);
}

template <int D>
static int build_lattice(const dlb_crf_config* c, const LatticeOffsets& o, Lattice& L, const uint8_t* image, float sx,
                         float sr, uint8_t* base, float4* ones, cudaStream_t st) {
  const int N = c->H * c->W;
  const int nb = o.nb;
  const size_t inc = o.inc;
  DLB_CUDA(cudaMemsetAsync(L.table, 0xFF, sizeof(unsigned long long) * o.cap * nb, st));
  DLB_CUDA(cudaMemsetAsync(L.slot_idx, 0x7F, sizeof(int) * static_cast<size_t>(o.cap) * nb, st));   // "no pixel yet"
  DLB_CUDA(cudaMemsetAsync(L.count, 0, 256 * nb, st));
  DLB_CUDA(cudaMemsetAsync(L.overflow, 0, 256 * nb, st));
  DLB_CUDA(cudaMemsetAsync(L.deg, 0, sizeof(int) * (inc + 2) * nb, st));
  launch_k(crf_embed_kernel<D>, grid2(N, 256, nb), 256, 0, st, c->H, c->W, image, sx, sx, sr, L);
  // raster-order vertex numbering: owner flags -> exclusive scan -> dense indices
  launch_k(crf_owner_flag_kernel, grid2(inc + 1, 256, nb), 256, 0, st, D + 1, L);
  const int n1 = static_cast<int>(inc) + 1;
  const int nblk1 = (n1 + kScanBlock - 1) / kScanBlock;
  launch_k(scan_phase1, dim3(nblk1, nb), kScanBlock, 0, st, L.cursor, inc + 2, n1, (const int*)nullptr, L.block_sums, o.nblk);
  launch_k(scan_phase2, dim3(1, nb), kScanBlock, 0, st, L.block_sums, o.nblk, n1, (const int*)nullptr);
  launch_k(scan_phase3, dim3(nblk1, nb), kScanBlock, 0, st, L.cursor, inc + 2, n1, (const int*)nullptr, (const int*)L.block_sums, o.nblk);
  launch_k(crf_assign_kernel, grid2(inc, 256, nb), 256, 0, st, L);
  launch_k(crf_relabel_kernel, grid2(inc, 256, nb), 256, 0, st, L);
  DLB_CUDA(cudaMemsetAsync(L.cursor, 0, sizeof(int) * (inc + 2) * nb, st));
  // CSR row starts: exclusive scan of the incidence counts over count + 1 entries (count read on the device)
  const int nblk2 = (static_cast<int>(inc) + 1 + kScanBlock - 1) / kScanBlock;
  launch_k(scan_phase1, dim3(nblk2, nb), kScanBlock, 0, st, L.deg, inc + 2, 0, (const int*)L.count, L.block_sums, o.nblk);
  launch_k(scan_phase2, dim3(1, nb), kScanBlock, 0, st, L.block_sums, o.nblk, 0, (const int*)L.count);
  launch_k(scan_phase3, dim3(nblk2, nb), kScanBlock, 0, st, L.deg, inc + 2, 0, (const int*)L.count, (const int*)L.block_sums, o.nblk);
  launch_k(crf_fill_kernel, grid2(inc, 256, nb), 256, 0, st, D + 1, L);
  launch_k(crf_neighbors_kernel<D>, grid2(inc, 256, nb), 256, 0, st, L);
  g_launches += 13;
  // norm = 1/sqrt(K 1 + 1e-20): filter a ones vector (value width 4, channel 0).  `ones` = [N] float4 scratch (the Q
  // buffer, not yet in use); the vertex values ping-pong between val0 and val1 with inc float4 per image.
  const size_t vstride = inc;
  float4* a = reinterpret_cast<float4*>(base + o.val0);
  float4* b = reinterpret_cast<float4*>(base + o.val1);
  launch_k(fill_ones4_kernel, dim3((N + 255) / 256), 256, 0, st, N, ones);
  launch_k(crf_splat_kernel<1>, grid_cap2(static_cast<long long>(inc) * 32, 256, nb), 256, 0, st, (const float4*)ones,
           static_cast<size_t>(0), a, vstride, L);
  for (int j = 0; j <= D; ++j) {
    launch_k(crf_blur_kernel, grid_cap2(static_cast<long long>(inc), 256, nb), 256, 0, st, 1, j, (const float4*)a, b, vstride, L);
    float4* t = a; a = b; b = t;
  }
  const float alpha = 1.0f / (1.0f + powf(2.0f, -static_cast<float>(D)));
  launch_k(crf_slice_norm_kernel, grid2(N, 256, nb), 256, 0, st, D + 1, alpha, (const float4*)a, vstride, L);
  launch_k(crf_scale_weights_kernel, grid2(inc, 256, nb), 256, 0, st, L);
  g_launches += 4 + D + 1;
  return check_launch("crf build");
}

static int run_meanfield(const dlb_crf_config* c, int batch, const CrfPlan& P, uint8_t* base, Lattice& Lg, Lattice& Lb,
                         uint8_t* map_out, cudaStream_t st) {
  const int N = c->H * c->W, M = c->M, VS = P.VS, vs4 = VS / 4;
  const float4* negU = reinterpret_cast<const float4*>(base + P.negU);
  float4* Q = reinterpret_cast<float4*>(base + P.Q);
  const bool use_g = c->compat_gauss != 0.f, use_b = c->compat_bilat != 0.f;
  const size_t q_stride = static_cast<size_t>(N) * vs4;
  const float alpha_g = 1.0f / (1.0f + powf(2.0f, -2.0f)), alpha_b = 1.0f / (1.0f + powf(2.0f, -5.0f));
  for (int it = 0; it < c->iters; ++it) {
    const float4* res[2] = {nullptr, nullptr};
    size_t res_stride[2] = {0, 0};
    for (int term = 0; term < 2; ++term) {
      const bool is_g = term == 0;
      if ((is_g && !use_g) || (!is_g && !use_b)) continue;
      Lattice& L = is_g ? Lg : Lb;
      const LatticeOffsets& o = is_g ? P.g : P.b;
      const int D = is_g ? 2 : 5;
      const size_t vstride = static_cast<size_t>(o.inc) * vs4;
      float4* v0 = reinterpret_cast<float4*>(base + o.val0);
      float4* v1 = reinterpret_cast<float4*>(base + o.val1);
      const long long work = static_cast<long long>(o.inc) * vs4;
      const dim3 sg = grid_cap2(static_cast<long long>(o.inc) * 32, 256, batch);
      switch (vs4) {
#define SPLAT(V) case V: launch_k(crf_splat_kernel<V>, sg, 256, 0, st, (const float4*)Q, q_stride, v0, vstride, L); break;
        SPLAT(1) SPLAT(2) SPLAT(3) SPLAT(4) SPLAT(5) SPLAT(6) SPLAT(7) SPLAT(8)
#undef SPLAT
      }
      for (int j = 0; j <= D; ++j) {
        launch_k(crf_blur_kernel, grid_cap2(work, 256, batch), 256, 0, st, vs4, j, (const float4*)v0, v1, vstride, L);
        float4* t = v0; v0 = v1; v1 = t;
      }
      res[term] = v0; res_stride[term] = vstride;
      g_launches += 2 + D;
    }
    const bool want_map = map_out && it == c->iters - 1;
#define SLICE(VH)                                                                                              \
    launch_k(crf_slice2_kernel<VH>, grid2(static_cast<long long>(N) * 2, 256, batch), 256, 0, st, N, M, vs4,     \
             use_g ? 1 : 0, alpha_g, c->compat_gauss, res[0], res_stride[0], Lg,                                 \
             use_b ? 1 : 0, alpha_b, c->compat_bilat, res[1], res_stride[1], Lb, negU, Q,                        \
             want_map ? map_out : static_cast<uint8_t*>(nullptr))
    switch ((vs4 + 1) / 2) { case 1: SLICE(1); break; case 2: SLICE(2); break; case 3: SLICE(3); break; default: SLICE(4); break; }
#undef SLICE
    g_launches++;
  }
  return check_launch("crf meanfield");
}

}  // namespace dlb

using namespace dlb;

extern "C" int64_t dlb_crf_workspace_bytes_batched(const dlb_crf_config* cfg, int batch) {
  if (!cfg || cfg->H <= 0 || cfg->W <= 0 || cfg->M <= 0 || batch <= 0) return 0;
  return static_cast<int64_t>(plan_crf(cfg, batch).total);
}
extern "C" int64_t dlb_crf_workspace_bytes(const dlb_crf_config* cfg) { return dlb_crf_workspace_bytes_batched(cfg, 1); }

extern "C" int dlb_crf_inference_batched(const dlb_crf_config* cfg, int batch, const float* unary, const uint8_t* image,
                                         float* Q_out, uint8_t* map_out, void* workspace, int64_t workspace_bytes,
                                         void* stream) {
  DLB_REQUIRE(cfg && unary && image && Q_out && workspace, "crf_inference: null pointer");
  DLB_REQUIRE(batch >= 1 && batch <= 65535, "crf_inference: 1 <= batch <= 65535 (got %d)", batch);
  DLB_REQUIRE(cfg->M >= 1 && cfg->M <= 32, "crf_inference: 1 <= labels <= 32 supported (got %d)", cfg->M);
  DLB_REQUIRE(cfg->compat_gauss != 0.f || cfg->compat_bilat != 0.f, "crf_inference: no pairwise term enabled");
  DLB_REQUIRE(cfg->unary_layout == 0 || cfg->unary_layout == 1, "crf_inference: unary_layout must be 0 or 1");
  DLB_REQUIRE(static_cast<long long>(cfg->H) * cfg->W * 6 < (1ll << 30), "crf_inference: image too large");
  const CrfPlan P = plan_crf(cfg, batch);
  DLB_REQUIRE(workspace_bytes >= static_cast<int64_t>(P.total), "crf_inference: workspace too small (%lld < %lld)",
              (long long)workspace_bytes, (long long)P.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  const int N = cfg->H * cfg->W;
  Lattice Lg = bind(P.g, base, 2, N), Lb = bind(P.b, base, 5, N);
  int rc;
  float* negU = reinterpret_cast<float*>(base + P.negU);
  float* Q = reinterpret_cast<float*>(base + P.Q);
  if (cfg->compat_gauss != 0.f) {
    rc = build_lattice<2>(cfg, P.g, Lg, image, cfg->sxy_gauss, 1.f, base, reinterpret_cast<float4*>(Q), st);
    if (rc) return rc;
  }
  if (cfg->compat_bilat != 0.f) {
    rc = build_lattice<5>(cfg, P.b, Lb, image, cfg->sxy_bilat, cfg->srgb_bilat, base, reinterpret_cast<float4*>(Q), st);
    if (rc) return rc;
  }
  const size_t smem = sizeof(float) * 256 * (P.VS + 1);
  launch_k(crf_unary_in_kernel, dim3((N + 255) / 256, batch), 256, smem + 2 * 256 * sizeof(float), st, N, cfg->M, P.VS,
           unary, negU, Q, cfg->unary_layout == 1 ? 1 : 0);
  g_launches++;
  rc = run_meanfield(cfg, batch, P, base, Lg, Lb, map_out, st);
  if (rc) return rc;
  launch_k(crf_q_out_kernel, dim3((N + 255) / 256, batch), 256, smem, st, N, cfg->M, P.VS, (const float*)Q, Q_out);
  g_launches++;
  return check_launch("crf_q_out_kernel");
}

extern "C" int dlb_crf_inference(const dlb_crf_config* cfg, const float* unary, const uint8_t* image, float* Q_out,
                                 uint8_t* map_out, void* workspace, int64_t workspace_bytes, void* stream) {
  return dlb_crf_inference_batched(cfg, 1, unary, image, Q_out, map_out, workspace, workspace_bytes, stream);
}

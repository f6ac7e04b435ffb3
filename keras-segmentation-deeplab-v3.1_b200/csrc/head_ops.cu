// Segmentation head: legacy TF1 bilinear resize + softmax (+ void-ignoring sparse cross-entropy and its
// gradient), Subpixel phase shift, confusion counts.
//   resize   : K.tf.image.resize_bilinear(x, size)   deeplabv3p.py:382,:418,:439 ; utils.py:190
//              (align_corners=False, no half-pixel centres: src = dst * in/out, hi = min(lo+1, in-1))
//   softmax  : Activation('softmax')                 deeplabv3p.py:440-444 ; utils.py:192,197
//   loss     : sparse_crossentropy_ignoring_last_label utils.py:127-130 (+ Keras categorical_crossentropy clip
//              1e-7 and the temporal sample-weight normalisation of weighted_masked_objective)
//   shuffle  : Subpixel._phase_shift                 subpixel.py:77-88
//   metrics  : utils.py:132-157 work from an argmax map + confusion counts
#include <atomic>

#include "common.cuh"

namespace dlb {

extern std::atomic<long long> g_launches;

constexpr int kMaxC = 32;   // classes held in registers

__device__ __forceinline__ void bilinear_coeffs(int dst, int in_size, float scale, int& lo, int& hi, float& f) {
  const float src = static_cast<float>(dst) * scale;
  lo = static_cast<int>(floorf(src));
  hi = min(lo + 1, in_size - 1);
  f = src - static_cast<float>(lo);
}

// ---------------------------------------------------------------------------------------------
// inference: probs [B, H*W, C] fp32 and/or argmax
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) resize_softmax_fwd_kernel(int B, int h, int w, int ldl, int H, int W,
                                                                 const float* __restrict__ logits,
                                                                 float* __restrict__ probs,
                                                                 uint8_t* __restrict__ argmax_out) {
  pdl_prologue();
  __shared__ float s_out[128 * C];
  const long long npix = static_cast<long long>(B) * H * W;
  const float sy = static_cast<float>(h) / static_cast<float>(H), sx = static_cast<float>(w) / static_cast<float>(W);
  for (long long p0 = static_cast<long long>(blockIdx.x) * 128; p0 < npix; p0 += static_cast<long long>(gridDim.x) * 128) {
    const long long pix = p0 + threadIdx.x;
    if (pix < npix) {
      const int X = static_cast<int>(pix % W);
      const long long t = pix / W;
      const int Y = static_cast<int>(t % H);
      const int b = static_cast<int>(t / H);
      int y0, y1, x0, x1; float fy, fx;
      bilinear_coeffs(Y, h, sy, y0, y1, fy);
      bilinear_coeffs(X, w, sx, x0, x1, fx);
      const float* base = logits + static_cast<size_t>(b) * h * w * ldl;
      const float* tl = base + (static_cast<size_t>(y0) * w + x0) * ldl;
      const float* tr = base + (static_cast<size_t>(y0) * w + x1) * ldl;
      const float* bl = base + (static_cast<size_t>(y1) * w + x0) * ldl;
      const float* br = base + (static_cast<size_t>(y1) * w + x1) * ldl;
      float v[C];
      float mx = -INFINITY; int am = 0;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float top = tl[c] + (tr[c] - tl[c]) * fx;
        const float bot = bl[c] + (br[c] - bl[c]) * fx;
        v[c] = top + (bot - top) * fy;
        if (v[c] > mx) { mx = v[c]; am = c; }
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - mx); sum += v[c]; }
      const float inv = 1.f / sum;
#pragma unroll
      for (int c = 0; c < C; ++c) s_out[threadIdx.x * C + c] = v[c] * inv;
      if (argmax_out) argmax_out[pix] = static_cast<uint8_t>(am);
    }
    if (probs) {
      __syncthreads();
      const long long n_here = min(static_cast<long long>(128), npix - p0) * C;
      float* dst = probs + p0 * C;
      for (int i = threadIdx.x; i < n_here; i += 128) dst[i] = s_out[i];
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// training: fused resize + softmax + CE loss + gradient w.r.t. the low-resolution logits.
// One CTA owns a TILE x TILE block of low-res cells = (TILE*S)^2 output pixels; the transposed resize is
// accumulated in shared memory (the S pixels that share a pair of source columns are first reduced with
// warp shuffles), then flushed with one global atomic per low-res logit of the (TILE+1)^2 halo tile.
// ---------------------------------------------------------------------------------------------
struct CeArgs {
  int B, h, w, ldl, H, W, S;
  const float* logits; const float* labels; const float* sample_w; const float* grad_scale;
  float* dlogits; double* loss_sum; uint8_t* argmax;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int C, int S>
__global__ void __launch_bounds__(256, 2) resize_softmax_ce_kernel(const CeArgs a) {
  pdl_prologue();
  constexpr int TILE = 64 / S >= 8 ? 8 : 64 / S;      // low-res cells per tile side (8 for S=8, 8 for S=4 -> 32px)
  constexpr int TP = TILE + 1;
  constexpr int PX = TILE * S;                        // output pixels per tile side (64 or 32)
  constexpr int RG = 256 / PX;                        // thread = (pixel column, row group); a row group owns cell rows
  static_assert(TILE % RG == 0 && S <= 32 && (S & (S - 1)) == 0, "tile / thread mapping");
  // gradient tiles are warp-private: shared-memory atomicAdd(float) is a compare-and-swap loop (37 % of this kernel's
  // stall samples sat on it); a warp's leader lanes instead read-modify-write the warp's own tile, ordered by
  // __syncwarp, and the eight tiles are summed once at the end
  constexpr int CW = 32 / S;                          // low-res cells a warp spans
  constexpr int GC = CW + 1;                          // corner columns of a warp's tile
  constexpr int NIT = TILE / RG;                      // cell rows a warp visits; its tile holds their 2 corner rows each
  constexpr int GSZ = 2 * NIT * GC * C;
  __shared__ float s_log[TP * TP * C];
  __shared__ float s_grad[8][GSZ];
  __shared__ float s_loss[8];
  const int tiles_x = (a.w + TILE - 1) / TILE, tiles_y = (a.h + TILE - 1) / TILE;
  const int tile = blockIdx.x % (tiles_x * tiles_y);
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int ty0 = (tile / tiles_x) * TILE, tx0 = (tile % tiles_x) * TILE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < TP * TP * C; i += 256) {
    const int c = i % C, cell = i / C;
    const int cy = min(ty0 + cell / TP, a.h - 1), cx = min(tx0 + cell % TP, a.w - 1);
    s_log[i] = a.logits[((static_cast<size_t>(b) * a.h + cy) * a.w + cx) * a.ldl + c];
  }
  for (int i = tid; i < 8 * GSZ; i += 256) (&s_grad[0][0])[i] = 0.f;
  __syncthreads();
  const float gs = *a.grad_scale;
  float loss_acc = 0.f;
  // A thread walks the S pixels of one pixel column inside one cell row.  The column's two source columns and its fx
  // are fixed, so the x-lerped logits (top / bottom rows) are computed once, the transposed resize is accumulated
  // in registers over the S rows (weights 1-fy / fy), and only then folded over the S lanes that share the source
  // columns -- 4 shared-memory atomics per class and cell instead of 4 per class and pixel group.
  const int lx = tid % PX, rg = tid / PX;
  const int X = tx0 * S + lx;
  const int cx0 = lx / S; const float fx = static_cast<float>(lx % S) / static_cast<float>(S);
  const int cx1 = (tx0 + cx0 + 1 <= a.w - 1) ? cx0 + 1 : cx0;
  for (int cy0 = rg; cy0 < TILE; cy0 += RG) {
    const int cy1 = (ty0 + cy0 + 1 <= a.h - 1) ? cy0 + 1 : cy0;
    const bool col_ok = X < a.W && ty0 + cy0 < a.h;
    float top[C], dlt[C], acc0[C], acc1[C];
    {
      const float* tl = &s_log[(cy0 * TP + cx0) * C];
      const float* tr = &s_log[(cy0 * TP + cx1) * C];
      const float* bl = &s_log[(cy1 * TP + cx0) * C];
      const float* br = &s_log[(cy1 * TP + cx1) * C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        top[c] = tl[c] + (tr[c] - tl[c]) * fx;
        dlt[c] = (bl[c] + (br[c] - bl[c]) * fx) - top[c];
        acc0[c] = 0.f; acc1[c] = 0.f;
      }
    }
    if (col_ok) {
#pragma unroll 1
      for (int r = 0; r < S; ++r) {
        const int Y = (ty0 + cy0) * S + r;
        const float fy = static_cast<float>(r) / static_cast<float>(S);
        float g[C];
        float mx = -INFINITY; int am = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          g[c] = top[c] + dlt[c] * fy;
          if (g[c] > mx) { mx = g[c]; am = c; }
        }
        // exp(g - mx) as ex2.approx(g * log2(e) - mx * log2(e)): one FFMA + one MUFU per class (the accurate expf is 9
        // instructions and was 22 % of this kernel); relative error of a probability <= 2e-6, far inside the 1e-4 the
        // loss / gradient parity tests ask for.  The inference kernel (probabilities returned to the user) keeps expf.
        const float nmx = -mx * 1.4426950408889634f;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) { g[c] = ex2_approx(fmaf(g[c], 1.4426950408889634f, nmx)); sum += g[c]; }
        const float inv = 1.f / sum;
        const size_t pix = (static_cast<size_t>(b) * a.H + Y) * a.W + X;
        if (a.argmax) a.argmax[pix] = static_cast<uint8_t>(am);
        const int label = static_cast<int>(a.labels[pix]);
        const float sw = a.sample_w ? a.sample_w[pix] : 1.f;
        float py = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) { g[c] *= inv; if (c == label) py = g[c]; }
        const bool valid = label >= 0 && label < C;
        // Keras: p = clip(p, 1e-7, 1 - 1e-7); loss = -log p[y]; the clip has zero gradient outside its range
        const bool in_range = valid && py >= 1e-7f && py <= 1.f - 1e-7f;
        if (valid) loss_acc += sw * -logf(fminf(fmaxf(py, 1e-7f), 1.f - 1e-7f));
        const float k = in_range ? sw * gs : 0.f;
        const float k0 = k * (1.f - fy), k1 = k * fy;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float d = g[c] - (c == label ? 1.f : 0.f);
          acc0[c] = fmaf(k0, d, acc0[c]);
          acc1[c] = fmaf(k1, d, acc1[c]);
        }
      }
    }
    // fold over the S lanes of the cell (same cx0 / cx1), split by the x weights; the leaders of a warp own distinct
    // cells, so their left-corner updates never collide, nor do their right-corner updates -- only a cell's right
    // corner with its neighbour's left one, which the __syncwarp between the two halves orders
    const bool leader = (lane % S) == 0;
    const int gx0 = cx0 - (((warp * 32) % PX) / S);    // corner column inside the warp's tile
    // on the right image edge cx1 is clamped to cx0: that cell's right-corner share is folded into its left corner, or
    // its right-corner update would land on the slot the neighbouring leader updates in the same instruction
    const bool xclamp = cx1 == cx0;
    const int gx1 = gx0 + 1;
    const int r0 = 2 * ((cy0 - rg) / RG), r1 = r0 + (cy1 - cy0);      // corner rows inside the warp's tile
    float* g00 = &s_grad[warp][(r0 * GC + gx0) * C];
    float* g10 = &s_grad[warp][(r1 * GC + gx0) * C];
    float* g01 = &s_grad[warp][(r0 * GC + gx1) * C];
    float* g11 = &s_grad[warp][(r1 * GC + gx1) * C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float v00 = acc0[c] * (1.f - fx), v01 = acc0[c] * fx, v10 = acc1[c] * (1.f - fx), v11 = acc1[c] * fx;
#pragma unroll
      for (int o = S / 2; o > 0; o >>= 1) {
        v00 += __shfl_xor_sync(0xffffffffu, v00, o);
        v01 += __shfl_xor_sync(0xffffffffu, v01, o);
        v10 += __shfl_xor_sync(0xffffffffu, v10, o);
        v11 += __shfl_xor_sync(0xffffffffu, v11, o);
      }
      if (xclamp) { v00 += v01; v10 += v11; v01 = 0.f; v11 = 0.f; }
      if (leader) { g00[c] += v00; g10[c] += v10; }      // (r1 == r0 on the bottom edge: same thread, program order)
      __syncwarp();
      if (leader) { g01[c] += v01; g11[c] += v11; }
      __syncwarp();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
  if (lane == 0) s_loss[warp] = loss_acc;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s_loss[i];
    atomicAdd(a.loss_sum, static_cast<double>(t));
  }
  // sum the warp tiles: warp w covers corner columns [w0, w0 + CW] of the CTA tile (w0 = ((w * 32) % PX) / S) and the
  // corner rows rg_w + it * RG + {0, 1} (rg_w = (w * 32) / PX)
  for (int i = tid; i < TP * TP * C; i += 256) {
    const int c = i % C, cell = i / C;
    const int ry = cell / TP, rx = cell % TP;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int gx = rx - ((w * 32) % PX) / S;
      if (gx < 0 || gx >= GC) continue;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int dy = ry - ((w * 32) / PX + it * RG);
        if (dy == 0 || dy == 1) v += s_grad[w][((2 * it + dy) * GC + gx) * C + c];
      }
    }
    if (v == 0.f) continue;
    const int cy = ty0 + ry, cx = tx0 + rx;
    if (cy < a.h && cx < a.w) atomicAdd(&a.dlogits[((static_cast<size_t>(b) * a.h + cy) * a.w + cx) * a.ldl + c], v);
  }
}

// S == 1 (Subpixel head: logits already at full resolution)
template <int C>
__global__ void __launch_bounds__(256) softmax_ce_full_kernel(const CeArgs a) {
  pdl_prologue();
  __shared__ float s_loss[8];
  const long long npix = static_cast<long long>(a.B) * a.H * a.W;
  const float gs = *a.grad_scale;
  float loss_acc = 0.f;
  for (long long pix = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; pix < npix;
       pix += static_cast<long long>(gridDim.x) * 256) {
    const float* l = a.logits + pix * a.ldl;
    float g[C]; float mx = -INFINITY; int am = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) { g[c] = l[c]; if (g[c] > mx) { mx = g[c]; am = c; } }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { g[c] = expf(g[c] - mx); sum += g[c]; }
    const float inv = 1.f / sum;
    if (a.argmax) a.argmax[pix] = static_cast<uint8_t>(am);
    const int label = static_cast<int>(a.labels[pix]);
    const float sw = a.sample_w ? a.sample_w[pix] : 1.f;
    float py = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { g[c] *= inv; if (c == label) py = g[c]; }
    const bool valid = label >= 0 && label < C;
    const bool in_range = valid && py >= 1e-7f && py <= 1.f - 1e-7f;
    if (valid) loss_acc += sw * -logf(fminf(fmaxf(py, 1e-7f), 1.f - 1e-7f));
    const float k = in_range ? sw * gs : 0.f;
    float* d = a.dlogits + pix * a.ldl;
#pragma unroll
    for (int c = 0; c < C; ++c) d[c] = k * (g[c] - (c == label ? 1.f : 0.f));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
  if ((threadIdx.x & 31) == 0) s_loss[threadIdx.x >> 5] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s_loss[i];
    atomicAdd(a.loss_sum, static_cast<double>(t));
  }
}

__global__ void count_nonzero_kernel(long long n, const float* sw, double* count) {
  pdl_prologue();
  __shared__ unsigned int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  unsigned int c = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    c += sw[i] != 0.f ? 1u : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(count, static_cast<double>(s_cnt));
}
__global__ void grad_scale_kernel(long long n, int has_sw, double* count, float* grad_scale, float loss_scale,
                                  const float* ls_state) {
  pdl_prologue();
  if (!has_sw) *count = static_cast<double>(n);
  const double c = *count;
  const double ls = static_cast<double>(loss_scale) * (ls_state ? static_cast<double>(ls_state[0]) : 1.0);
  *grad_scale = c > 0.0 ? static_cast<float>(ls / c) : 0.f;
}

// ---------------------------------------------------------------------------------------------
// Subpixel phase shift, standalone:  out[n, a*r+j, b*r+i, k] = in[n, a, b, k*r*r + i*r + j]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void phase_shift_kernel(long long total, int h, int w, int Cs, int r, const T* in, T* out, int inverse) {
  pdl_prologue();
  for (long long o = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; o < total;
       o += static_cast<long long>(gridDim.x) * blockDim.x) {
    // o indexes the high-res tensor [n, h*r, w*r, Cs]
    const int k = static_cast<int>(o % Cs);
    long long t = o / Cs;
    const int X = static_cast<int>(t % (w * r)); t /= (w * r);
    const int Y = static_cast<int>(t % (h * r));
    const long long n = t / (h * r);
    const int a = Y / r, j = Y % r, b = X / r, i = X % r;
    const long long src = ((n * h + a) * w + b) * (static_cast<long long>(Cs) * r * r) + static_cast<long long>(k) * r * r + i * r + j;
    if (inverse) out[src] = in[o];
    else out[o] = in[src];
  }
}

// backward of the FUSED Subpixel store of dlb_pw_gemm: the GEMM's columns are ordered (jj, i, k), so
//   dst[n, a, b, (jj*r + i)*Cs + k] = src[n, a*r + jj, b*r + i, k]      (fp32 gradient -> storage dtype)
// reads and writes are both contiguous runs of r*Cs elements.
template <typename T>
__global__ void __launch_bounds__(256) subpixel_grad_gather_kernel(long long total, int h, int w, int Cs, int r,
                                                                   const float* __restrict__ src, T* __restrict__ dst) {
  pdl_prologue();
  const int run = r * Cs;
  for (long long o = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; o < total;
       o += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(o % run);          // (i, k)
    long long t = o / run;
    const int jj = static_cast<int>(t % r); t /= r;
    const int b = static_cast<int>(t % w); t /= w;
    const int a = static_cast<int>(t % h);
    const long long n = t / h;
    const long long s = ((n * h * r + static_cast<long long>(a) * r + jj) * (static_cast<long long>(w) * r) + static_cast<long long>(b) * r) * Cs + e;
    Act<T>::st(&dst[o], src[s]);
  }
}

__global__ void __launch_bounds__(256) confusion_kernel(long long npix, int C, const float* labels,
                                                        const uint8_t* argmax, unsigned long long* conf) {
  pdl_prologue();
  extern __shared__ unsigned int s_conf[];   // [(C+1)*C]
  const int b = blockIdx.y;
  const int nb = (C + 1) * C;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s_conf[i] = 0;
  __syncthreads();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int l = static_cast<int>(labels[b * npix + i]);
    const int p = argmax[b * npix + i];
    if (l >= 0 && l <= C && p < C) atomicAdd(&s_conf[l * C + p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    if (s_conf[i]) atomicAdd(&conf[static_cast<size_t>(b) * nb + i], static_cast<unsigned long long>(s_conf[i]));
}


// ---------------------------------------------------------------------------------------------
// SegmentationGenerator label contract (reference utils.py:360-399): void remap + per-image balanced class weights.
//   y  = label, every value outside 0..n_classes-1 -> n_classes (void)            (utils.py:360-365)
//   sw = n_valid / (n_present * count[y]) for valid pixels, 0 for void             (utils.py:388-399, sklearn
//        compute_class_weight('balanced'): n_samples / (n_classes_present * bincount), float64 then stored as float32)
// Two passes over the labels: per-image histogram (shared-memory bins), then weights; bit-exact with the numpy path.
// ---------------------------------------------------------------------------------------------
template <typename LT>
__device__ __forceinline__ int label_remap(LT v, int n_classes) {
  const long long l = static_cast<long long>(v);
  return (l < 0 || l >= n_classes || static_cast<LT>(l) != v) ? n_classes : static_cast<int>(l);
}

template <typename LT>
__global__ void __launch_bounds__(256) label_hist_kernel(long long npix, int n_classes, const LT* __restrict__ labels,
                                                         unsigned long long* __restrict__ counts) {
  pdl_prologue();
  extern __shared__ unsigned int s_hist[];   // [n_classes + 1]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i <= n_classes; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const LT* lb = labels + static_cast<size_t>(b) * npix;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    atomicAdd(&s_hist[label_remap(lb[i], n_classes)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i <= n_classes; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&counts[static_cast<size_t>(b) * (n_classes + 1) + i], static_cast<unsigned long long>(s_hist[i]));
}

template <typename LT>
__global__ void __launch_bounds__(256) label_weight_kernel(long long npix, int n_classes, const LT* __restrict__ labels,
                                                           const unsigned long long* __restrict__ counts,
                                                           float* __restrict__ y, float* __restrict__ sw) {
  pdl_prologue();
  extern __shared__ float s_w[];             // [n_classes + 1]
  __shared__ double s_nvalid;
  __shared__ int s_present;
  const int b = blockIdx.y;
  const unsigned long long* cnt = counts + static_cast<size_t>(b) * (n_classes + 1);
  if (threadIdx.x == 0) {
    unsigned long long nv = 0; int k = 0;
    for (int c = 0; c < n_classes; ++c) { nv += cnt[c]; k += cnt[c] > 0; }
    s_nvalid = static_cast<double>(nv); s_present = k;
  }
  __syncthreads();
  for (int c = threadIdx.x; c <= n_classes; c += blockDim.x) {
    float w = 0.f;
    if (c < n_classes && cnt[c] > 0)
      w = static_cast<float>(s_nvalid / (static_cast<double>(s_present) * static_cast<double>(cnt[c])));
    s_w[c] = w;
  }
  __syncthreads();
  const LT* lb = labels + static_cast<size_t>(b) * npix;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int l = label_remap(lb[i], n_classes);
    if (y) y[static_cast<size_t>(b) * npix + i] = static_cast<float>(l);
    if (sw) sw[static_cast<size_t>(b) * npix + i] = s_w[l];
  }
}

}  // namespace dlb

using namespace dlb;

extern "C" int dlb_label_weights(int B, int64_t npix, int n_classes, int label_type, const void* labels,
                                 unsigned long long* counts, float* y, float* sw, void* stream) {
  DLB_REQUIRE(labels && counts && (y || sw), "label_weights: null pointer");
  DLB_REQUIRE(B > 0 && npix > 0 && n_classes > 0 && n_classes <= 4096, "label_weights: bad shape B=%d npix=%lld classes=%d", B,
              (long long)npix, n_classes);
  DLB_REQUIRE(label_type >= 0 && label_type <= 2, "label_weights: label_type must be 0 (u8), 1 (i32) or 2 (f32)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DLB_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * B * (n_classes + 1), st));
  long long blocks = (npix + 256 * 8 - 1) / (256 * 8);
  const long long cap = (static_cast<long long>(num_sms()) * 8 + B - 1) / B;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), B);
  const size_t smem = (n_classes + 1) * sizeof(unsigned int);
#define LW(LT)                                                                                              \
  do {                                                                                                      \
    launch_k(label_hist_kernel<LT>, grid, 256, smem, st, (long long)npix, n_classes, (const LT*)labels, counts);   \
    launch_k(label_weight_kernel<LT>, grid, 256, smem, st, (long long)npix, n_classes, (const LT*)labels,          \
             (const unsigned long long*)counts, y, sw);                                                     \
  } while (0)
  if (label_type == 0) LW(uint8_t);
  else if (label_type == 1) LW(int);
  else LW(float);
#undef LW
  g_launches += 2;
  return check_launch("label_weight_kernel");
}

extern "C" int dlb_resize_softmax_fwd(int B, int h, int w, int C, int ldl, int H, int W, const float* logits,
                                      float* probs, uint8_t* argmax, void* stream) {
  DLB_REQUIRE(logits && (probs || argmax), "resize_softmax_fwd: null pointer");
  DLB_REQUIRE(C >= 1 && C <= ldl, "resize_softmax_fwd: C=%d ldl=%d", C, ldl);
  const long long npix = static_cast<long long>(B) * H * W;
  long long blocks = (npix + 127) / 128;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (C) {
#define CASE(CC) case CC: launch_k(resize_softmax_fwd_kernel<CC>, grid, 128, 0, st, B, h, w, ldl, H, W, logits, probs, argmax); break;
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
    CASE(15) CASE(16) CASE(17) CASE(18) CASE(19) CASE(20) CASE(21) CASE(22) CASE(23) CASE(24)
#undef CASE
    default:
      set_last_error("resize_softmax_fwd: classes=%d unsupported (2..24)", C);
      return DLB_ERR_UNSUPPORTED;
  }
  g_launches++;
  return check_launch("resize_softmax_fwd_kernel");
}

template <int C>
static int launch_ce(const CeArgs& a, cudaStream_t st) {
  if (a.S == 1) {
    const long long npix = static_cast<long long>(a.B) * a.H * a.W;
    long long blocks = (npix + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 8;
    launch_k(softmax_ce_full_kernel<C>, static_cast<int>(blocks < cap ? blocks : cap), 256, 0, st, a);
  } else if (a.S == 8) {
    const int tiles = ((a.w + 7) / 8) * ((a.h + 7) / 8);
    launch_k(resize_softmax_ce_kernel<C, 8>, a.B * tiles, 256, 0, st, a);
  } else if (a.S == 4) {
    const int tiles = ((a.w + 7) / 8) * ((a.h + 7) / 8);
    launch_k(resize_softmax_ce_kernel<C, 4>, a.B * tiles, 256, 0, st, a);
  } else {
    set_last_error("resize_softmax_ce: scale %d unsupported (1, 4, 8)", a.S);
    return DLB_ERR_UNSUPPORTED;
  }
  g_launches++;
  return check_launch("resize_softmax_ce_kernel");
}

extern "C" int dlb_resize_softmax_ce(const dlb_softmax_ce_params* p, void* stream) {
  DLB_REQUIRE(p && p->logits && p->labels && p->grad_scale_dev && p->dlogits && p->loss_sum,
              "resize_softmax_ce: null pointer");
  DLB_REQUIRE(p->H % p->h == 0 && p->W % p->w == 0 && p->H / p->h == p->W / p->w,
              "resize_softmax_ce: integer isotropic scale required (h=%d H=%d w=%d W=%d)", p->h, p->H, p->w, p->W);
  CeArgs a{p->B, p->h, p->w, p->ldl, p->H, p->W, p->H / p->h, p->logits, p->labels, p->sample_w,
           p->grad_scale_dev, p->dlogits, p->loss_sum, p->argmax};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (p->C) {     // the same 2..24 range as dlb_resize_softmax_fwd: create_seg_model(n=...) accepts any class count
#define CASE(CC) case CC: return launch_ce<CC>(a, st);
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
    CASE(15) CASE(16) CASE(17) CASE(18) CASE(19) CASE(20) CASE(21) CASE(22) CASE(23) CASE(24)
#undef CASE
    default:
      set_last_error("resize_softmax_ce: classes=%d unsupported (2..24)", p->C);
      return DLB_ERR_UNSUPPORTED;
  }
}

extern "C" int dlb_ce_grad_scale(int64_t n, const float* sample_w, float* grad_scale_dev, double* wcount,
                                 float loss_scale, const float* loss_scale_state, void* stream) {
  DLB_REQUIRE(grad_scale_dev && wcount && n > 0, "ce_grad_scale: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DLB_CUDA(cudaMemsetAsync(wcount, 0, sizeof(double), st));
  if (sample_w) {
    long long blocks = (n + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 8;
    launch_k(count_nonzero_kernel, static_cast<int>(blocks < cap ? blocks : cap), 256, 0, st, n, sample_w, wcount);
    g_launches++;
  }
  launch_k(grad_scale_kernel, 1, 1, 0, st, n, sample_w != nullptr, wcount, grad_scale_dev, loss_scale, loss_scale_state);
  g_launches++;
  return check_launch("grad_scale_kernel");
}

extern "C" int dlb_phase_shift(int B, int h, int w, int Cs, int r, int dtype, const void* in, void* out, int inverse,
                               void* stream) {
  DLB_REQUIRE(in && out && r >= 1, "phase_shift: bad arguments");
  const long long total = static_cast<long long>(B) * h * r * w * r * Cs;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == DLB_F32) launch_k(phase_shift_kernel<float>, grid, 256, 0, st, total, h, w, Cs, r, (const float*)in, (float*)out, inverse);
  else launch_k(phase_shift_kernel<uint16_t>, grid, 256, 0, st, total, h, w, Cs, r, (const uint16_t*)in, (uint16_t*)out, inverse);
  g_launches++;
  return check_launch("phase_shift_kernel");
}

extern "C" int dlb_subpixel_grad_gather(int B, int h, int w, int Cs, int r, const float* dlogits, int dst_dtype, void* dst,
                                        void* stream) {
  DLB_REQUIRE(dlogits && dst && B > 0 && h > 0 && w > 0 && Cs > 0 && r >= 1, "subpixel_grad_gather: bad arguments");
  const long long total = static_cast<long long>(B) * h * w * r * r * Cs;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dst_dtype == DLB_F16) launch_k(subpixel_grad_gather_kernel<__half>, grid, 256, 0, st, total, h, w, Cs, r, dlogits, (__half*)dst);
  else if (dst_dtype == DLB_BF16) launch_k(subpixel_grad_gather_kernel<__nv_bfloat16>, grid, 256, 0, st, total, h, w, Cs, r, dlogits, (__nv_bfloat16*)dst);
  else launch_k(subpixel_grad_gather_kernel<float>, grid, 256, 0, st, total, h, w, Cs, r, dlogits, (float*)dst);
  g_launches++;
  return check_launch("subpixel_grad_gather_kernel");
}

extern "C" int dlb_confusion(int B, int64_t npix, int C, const float* labels, const uint8_t* argmax,
                             unsigned long long* conf, void* stream) {
  DLB_REQUIRE(labels && argmax && conf && C >= 1 && C <= 255, "confusion: bad arguments");
  long long blocks = (npix + 255) / 256;
  if (blocks > 64) blocks = 64;
  dim3 grid(static_cast<unsigned>(blocks), B);
  launch_k(confusion_kernel, grid, 256, (C + 1) * C * sizeof(unsigned int), static_cast<cudaStream_t>(stream), 
      npix, C, labels, argmax, conf);
  g_launches++;
  return check_launch("confusion_kernel");
}

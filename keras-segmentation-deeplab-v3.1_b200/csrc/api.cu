// C-ABI plumbing (version, thread-local error, launch counter) + optimizer / cast / fill kernels.
//   Adam : Keras 2.2.4 optimizers.Adam as configured at segmentation.ipynb cell "compile" (ipynb:107):
//          Adam(lr=7e-4, epsilon=1e-8, decay=1e-6)
#include <cstdlib>
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace dlb {

std::atomic<long long> g_launches{0};
static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return DLB_ERR_CUDA;
  }
  return DLB_OK;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("DLB_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// loss_scale_state (device float[4], optional): {scale, good_steps, found_inf, growth_interval} -- dynamic fp16 loss
// scaling.  found_inf is set by grad_check_kernel; a step that saw a non-finite gradient leaves p / m / v and the
// iteration counter untouched and halves the scale (scale_update_kernel).
__global__ void adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, const long long* __restrict__ step, float lr, float b1, float b2,
                            float eps, float decay, float gmult, const float* __restrict__ mask,
                            const float* __restrict__ ls_state) {
  pdl_prologue();
  if (ls_state) {
    if (ls_state[2] != 0.f) return;           // overflow somewhere in this step's gradients: skip the update
    gmult = gmult / ls_state[0];
  }
  const long long it = *step;                 // iterations before this update
  const float t = static_cast<float>(it) + 1.f;
  float lr_t = lr;
  if (decay > 0.f) lr_t = lr_t * (1.f / (1.f + decay * static_cast<float>(it)));
  lr_t = lr_t * (sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t)));
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // frozen (trainable=False) weights are not part of the update at all in Keras: no m / v decay, no step
    if (mask && mask[i] == 0.f) continue;
    const float gi = g[i] * gmult;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}
__global__ void step_inc_kernel(long long* step, float* ls_state) {
  pdl_prologue();
  if (!ls_state) { *step += 1; return; }
  if (ls_state[2] != 0.f) {                   // overflow: halve, retry with the next batch
    ls_state[0] = fmaxf(ls_state[0] * 0.5f, 1.f);
    ls_state[1] = 0.f; ls_state[2] = 0.f;
  } else {
    *step += 1;
    ls_state[1] += 1.f;
    if (ls_state[3] > 0.f && ls_state[1] >= ls_state[3]) { ls_state[0] = fminf(ls_state[0] * 2.f, 65536.f); ls_state[1] = 0.f; }
  }
}
__global__ void __launch_bounds__(256) grad_check_kernel(long long n4, const float4* __restrict__ g, float* ls_state) {
  pdl_prologue();
  bool bad = false;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 q = g[i];
    // x - x is 0 for finite x and NaN for +-inf / NaN
    const float z = (q.x - q.x) + (q.y - q.y) + (q.z - q.z) + (q.w - q.w);
    bad |= !(z == 0.f);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) ls_state[2] = 1.f;
}

template <typename T>
__global__ void cast_weight_kernel(int K, int N, const float* __restrict__ w, T* __restrict__ w_kn, T* __restrict__ w_nk) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * N) return;
  const int k = i / N, n = i - k * N;
  const float v = w[i];
  if (w_kn) Act<T>::st(&w_kn[i], v);
  if (w_nk) Act<T>::st(&w_nk[static_cast<size_t>(n) * K + k], v);
}

// one launch for all layers: entry found by binary search over the prefix of flat element counts
__global__ void __launch_bounds__(256) cast_weights_batched_kernel(int n_entries, const long long* __restrict__ tab, long long total) {
  pdl_prologue();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int lo = 0, hi = n_entries - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tab[mid * 8 + 7] <= i) lo = mid; else hi = mid - 1;
    }
    const long long* e = tab + lo * 8;
    const float* w = reinterpret_cast<const float*>(e[0]);
    const int K = static_cast<int>(e[3]), N = static_cast<int>(e[4]), ld = static_cast<int>(e[5]), dt = static_cast<int>(e[6]);
    const int j = static_cast<int>(i - e[7]);
    const int k = j / N, n = j - k * N;
    const float v = w[j];
    const size_t o_kn = static_cast<size_t>(k) * ld + n, o_nk = static_cast<size_t>(n) * K + k;
    if (dt == DLB_F16) {
      if (e[1]) reinterpret_cast<__half*>(e[1])[o_kn] = __float2half_rn(v);
      if (e[2]) reinterpret_cast<__half*>(e[2])[o_nk] = __float2half_rn(v);
    } else if (dt == DLB_BF16) {
      if (e[1]) reinterpret_cast<__nv_bfloat16*>(e[1])[o_kn] = __float2bfloat16_rn(v);
      if (e[2]) reinterpret_cast<__nv_bfloat16*>(e[2])[o_nk] = __float2bfloat16_rn(v);
    } else {
      if (e[1]) reinterpret_cast<float*>(e[1])[o_kn] = v;
      if (e[2]) reinterpret_cast<float*>(e[2])[o_nk] = v;
    }
  }
}

template <typename S, typename D>
__global__ void cast_kernel(long long n, const S* __restrict__ s, D* __restrict__ d) {
  pdl_prologue();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    Act<D>::st(&d[i], Act<S>::ld(&s[i]));
}

}  // namespace dlb

using namespace dlb;

extern "C" int dlb_version(void) { return DLB_ABI_VERSION; }
extern "C" const char* dlb_last_error(void) { return g_err; }
extern "C" int64_t dlb_launch_count(void) { return g_launches.load(); }

extern "C" int dlb_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return major == 10 ? 1 : 0;
}

extern "C" int dlb_adam_step(int64_t n, float* param, const float* grad, float* m, float* v, int64_t* step_dev,
                             float lr, float beta1, float beta2, float eps, float decay, float grad_mult,
                             const float* train_mask, float* loss_scale_state, void* stream) {
  DLB_REQUIRE(n > 0 && param && grad && m && v && step_dev, "adam_step: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  launch_k(adam_kernel, static_cast<int>(blocks < cap ? blocks : cap), 256, 0, st,
      n, param, grad, m, v, reinterpret_cast<const long long*>(step_dev), lr, beta1, beta2, eps, decay, grad_mult,
      train_mask, static_cast<const float*>(loss_scale_state));
  launch_k(step_inc_kernel, 1, 1, 0, st, reinterpret_cast<long long*>(step_dev), loss_scale_state);
  g_launches += 2;
  return check_launch("adam_kernel");
}

extern "C" int dlb_grad_finite_check(int64_t n, const float* grad, float* loss_scale_state, void* stream) {
  DLB_REQUIRE(n > 0 && n % 4 == 0 && grad && loss_scale_state, "grad_finite_check: bad arguments (n must be a multiple of 4)");
  DLB_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "grad_finite_check: grad must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  launch_k(grad_check_kernel, static_cast<int>(blocks < cap ? blocks : cap), 256, 0, st, static_cast<long long>(n / 4),
           reinterpret_cast<const float4*>(grad), loss_scale_state);
  g_launches++;
  return check_launch("grad_check_kernel");
}

extern "C" int dlb_cast_weight(int K, int N, const float* w, int dtype, void* w_kn, void* w_nk, void* stream) {
  DLB_REQUIRE(w && (w_kn || w_nk) && K > 0 && N > 0, "cast_weight: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (K * N + 255) / 256;
  if (dtype == DLB_F16) launch_k(cast_weight_kernel<__half>, grid, 256, 0, st, K, N, w, (__half*)w_kn, (__half*)w_nk);
  else if (dtype == DLB_BF16) launch_k(cast_weight_kernel<__nv_bfloat16>, grid, 256, 0, st, K, N, w, (__nv_bfloat16*)w_kn, (__nv_bfloat16*)w_nk);
  else launch_k(cast_weight_kernel<float>, grid, 256, 0, st, K, N, w, (float*)w_kn, (float*)w_nk);
  g_launches++;
  return check_launch("cast_weight_kernel");
}

extern "C" int dlb_cast_weights_batched(int n_entries, const int64_t* table, int64_t total, void* stream) {
  DLB_REQUIRE(n_entries > 0 && table && total > 0, "cast_weights_batched: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  launch_k(cast_weights_batched_kernel, static_cast<int>(blocks < cap ? blocks : cap), 256, 0, st, 
      n_entries, reinterpret_cast<const long long*>(table), total);
  g_launches++;
  return check_launch("cast_weights_batched_kernel");
}

extern "C" int dlb_cast(int64_t n, int src_dtype, const void* src, int dst_dtype, void* dst, void* stream) {
  DLB_REQUIRE(n > 0 && src && dst, "cast: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
#define GO(S, D) launch_k(cast_kernel<S, D>, grid, 256, 0, st, n, (const S*)src, (D*)dst)
  if (src_dtype == DLB_F32 && dst_dtype == DLB_F16) GO(float, __half);
  else if (src_dtype == DLB_F32 && dst_dtype == DLB_BF16) GO(float, __nv_bfloat16);
  else if (src_dtype == DLB_F16 && dst_dtype == DLB_F32) GO(__half, float);
  else if (src_dtype == DLB_BF16 && dst_dtype == DLB_F32) GO(__nv_bfloat16, float);
  else if (src_dtype == DLB_F32 && dst_dtype == DLB_F32) GO(float, float);
  else { set_last_error("cast: unsupported dtype pair %d -> %d", src_dtype, dst_dtype); return DLB_ERR_UNSUPPORTED; }
#undef GO
  g_launches++;
  return check_launch("cast_kernel");
}

extern "C" int dlb_fill_zero(void* p, int64_t bytes, void* stream) {
  DLB_REQUIRE(p && bytes >= 0, "fill_zero: bad arguments");
  DLB_CUDA(cudaMemsetAsync(p, 0, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream)));
  return DLB_OK;
}

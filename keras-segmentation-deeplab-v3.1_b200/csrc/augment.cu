// SegmentationGenerator's per-image augmentations on the device (reference utils.py:319-365):
//
//   cv2.GaussianBlur(image, (k, k), 0)            :323-324   k in {3, 5, 7}
//   cv2.flip(image / label, 1 | 0)                :333-339
//   cv2.LUT(image, gamma table)                   :340-345   (the 256-entry table is built by the caller, as the reference does)
//   cv2.warpAffine(image / label, M, (W, H))      :346-357   INTER_LINEAR for BOTH image and label, constant border 0
//   labels that did not exist in the decoded label image -> void (n_classes)   :360-365
//
// then the float32 image X [B, H, W, 3] and the label map that dlb_label_weights turns into Y / SW.
//
// Bit-exact with OpenCV's uint8 paths (tests/test_augment.py checks against cv2 itself -- the one third-party
// dependency of the reference that exists in this image):
//   * GaussianBlur with sigma 0 and k <= 7 uses the fixed small kernels {1,2,1}/4, {1,4,6,4,1}/16, {2,7,14,18,14,7,2}/64
//     in 8.8 fixed point, BORDER_REFLECT_101, one rounding at the end: (sum_ij a_i a_j x_ij + half) >> shift
//   * warpAffine: destination -> source coordinates in 10-bit fixed point from the inverted matrix (doubles, rounded to
//     int per term), 1/32-pixel interpolation grid, 2x2 int16 weights that sum to 32768, (acc + 16384) >> 15
// Pure HBM-bound byte work: two passes over 4 bytes per pixel, thread per output pixel, coalesced along x.
#include "common.cuh"

namespace dlb {

extern std::atomic<long long> g_launches;

__constant__ int c_gauss[3][7] = {{1, 2, 1, 0, 0, 0, 0}, {1, 4, 6, 4, 1, 0, 0}, {2, 7, 14, 18, 14, 7, 2}};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i < 0 ? 0 : (i >= n ? n - 1 : i);       // (n = 1 .. k/2: degenerate sizes)
}

// labels present in the decoded label image (before any augmentation): 256-bit set per image
__global__ void __launch_bounds__(256) aug_present_kernel(long long npix, const uint8_t* __restrict__ label, unsigned int* __restrict__ present) {
  pdl_prologue();
  __shared__ unsigned int s[8];
  if (threadIdx.x < 8) s[threadIdx.x] = 0u;
  __syncthreads();
  const uint8_t* lb = label + static_cast<size_t>(blockIdx.y) * npix;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npix; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned int v = lb[i];
    atomicOr(&s[v >> 5], 1u << (v & 31));
  }
  __syncthreads();
  if (threadIdx.x < 8 && s[threadIdx.x]) atomicOr(&present[blockIdx.y * 8 + threadIdx.x], s[threadIdx.x]);
}

// pass 1: [blur] -> flips -> [LUT] on the image, flips on the label; uint8 in, uint8 out
__global__ void __launch_bounds__(256) aug_pass1_kernel(int H, int W, const uint8_t* __restrict__ img, const uint8_t* __restrict__ label,
                                                        const dlb_aug_params* __restrict__ prm, const uint8_t* __restrict__ luts,
                                                        uint8_t* __restrict__ img_out, uint8_t* __restrict__ label_out) {
  pdl_prologue();
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const dlb_aug_params p = prm[b];
  const size_t ioff = static_cast<size_t>(b) * H * W;
  // destination (x, y) of the flips <- source (sx, sy)
  const int sx = p.hflip ? W - 1 - x : x, sy = p.vflip ? H - 1 - y : y;
  int v[3];
  if (p.blur_ksize >= 3) {
    const int k = p.blur_ksize, r = k >> 1, t = r - 1;       // table row: k = 3, 5, 7 -> 0, 1, 2
    const int shift = 4 * r;                                   // (4, 16, 64)^2 = 2^(4r)
    int acc[3] = {0, 0, 0};
    for (int i = 0; i < k; ++i) {
      const int yy = reflect101(sy + i - r, H);
      for (int j = 0; j < k; ++j) {
        const int xx = reflect101(sx + j - r, W);
        const int w = c_gauss[t][i] * c_gauss[t][j];
        const uint8_t* q = img + (ioff + static_cast<size_t>(yy) * W + xx) * 3;
        acc[0] += w * q[0]; acc[1] += w * q[1]; acc[2] += w * q[2];
      }
    }
    const int half = 1 << (shift - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (acc[c] + half) >> shift;
  } else {
    const uint8_t* q = img + (ioff + static_cast<size_t>(sy) * W + sx) * 3;
    v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
  }
  if (luts) {
    const uint8_t* lut = luts + b * 256;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = lut[v[c]];
  }
  uint8_t* o = img_out + (ioff + static_cast<size_t>(y) * W + x) * 3;
  o[0] = static_cast<uint8_t>(v[0]); o[1] = static_cast<uint8_t>(v[1]); o[2] = static_cast<uint8_t>(v[2]);
  label_out[ioff + static_cast<size_t>(y) * W + x] = label[ioff + static_cast<size_t>(sy) * W + sx];
}

// cv::initInterTab2D(INTER_LINEAR, fixed point): the 2x2 int16 weights of sub-pixel position (ay, ax) / 32
__device__ __forceinline__ void bilinear_itab(int ay, int ax, int (&w)[4]) {
  const float fy = static_cast<float>(ay) * (1.f / 32.f), fx = static_cast<float>(ax) * (1.f / 32.f);
  const float ty[2] = {1.f - fy, fy}, tx[2] = {1.f - fx, fx};
  int sum = 0, imax = 0, imin = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float f = __fmul_rn(__fmul_rn(ty[k >> 1], tx[k & 1]), 32768.f);
    int iv = __float2int_rn(f);                       // saturate_cast<short>(cvRound)
    iv = iv > 32767 ? 32767 : (iv < -32768 ? -32768 : iv);
    w[k] = iv; sum += iv;
  }
  if (sum != 32768) {
    // the first maximum / first minimum (row-major) absorbs the rounding difference
#pragma unroll
    for (int k = 1; k < 4; ++k) { if (w[k] > w[imax]) imax = k; if (w[k] < w[imin]) imin = k; }
    const int d = 32768 - sum;
    if (d < 0) w[imax] += d; else w[imin] += d;
  }
}

// pass 2: [warpAffine of image and label] -> float32 image, label with "new values -> void"
__global__ void __launch_bounds__(256) aug_pass2_kernel(int H, int W, int n_classes, const uint8_t* __restrict__ img,
                                                        const uint8_t* __restrict__ label, const dlb_aug_params* __restrict__ prm,
                                                        const unsigned int* __restrict__ present, float* __restrict__ X,
                                                        uint8_t* __restrict__ label_out) {
  pdl_prologue();
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const dlb_aug_params p = prm[b];
  const size_t ioff = static_cast<size_t>(b) * H * W;
  int v[3], lab;
  if (p.warp) {
    // cv::warpAffine: AB_BITS = 10, INTER_BITS = 5; every product is rounded to an int on its own (saturate_cast)
    const double AB = 1024.0;
    const int round_delta = 16;
    // (explicit rn multiplies / adds: an fma contraction would round differently from the host library)
    const double xd = static_cast<double>(x), yd = static_cast<double>(y);
    const int adelta = static_cast<int>(llrint(__dmul_rn(__dmul_rn(p.minv[0], xd), AB)));
    const int bdelta = static_cast<int>(llrint(__dmul_rn(__dmul_rn(p.minv[3], xd), AB)));
    const int X0 = static_cast<int>(llrint(__dmul_rn(__dadd_rn(__dmul_rn(p.minv[1], yd), p.minv[2]), AB))) + round_delta;
    const int Y0 = static_cast<int>(llrint(__dmul_rn(__dadd_rn(__dmul_rn(p.minv[4], yd), p.minv[5]), AB))) + round_delta;
    const int Xf = (X0 + adelta) >> 5, Yf = (Y0 + bdelta) >> 5;
    const int sx = Xf >> 5, sy = Yf >> 5;
    int w[4];
    bilinear_itab(Yf & 31, Xf & 31, w);
    int acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = sy + (k >> 1), xx = sx + (k & 1);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const size_t o = ioff + static_cast<size_t>(yy) * W + xx;
        const uint8_t* q = img + o * 3;
        acc[0] += w[k] * q[0]; acc[1] += w[k] * q[1]; acc[2] += w[k] * q[2];
        acc[3] += w[k] * label[o];
      }
    }
    v[0] = (acc[0] + 16384) >> 15; v[1] = (acc[1] + 16384) >> 15; v[2] = (acc[2] + 16384) >> 15;
    lab = (acc[3] + 16384) >> 15;
  } else {
    const size_t o = ioff + static_cast<size_t>(y) * W + x;
    v[0] = img[o * 3]; v[1] = img[o * 3 + 1]; v[2] = img[o * 3 + 2];
    lab = label[o];
  }
  const size_t o = ioff + static_cast<size_t>(y) * W + x;
  X[o * 3] = static_cast<float>(v[0]); X[o * 3 + 1] = static_cast<float>(v[1]); X[o * 3 + 2] = static_cast<float>(v[2]);
  // utils.py:360-365: values the interpolation invented (not in the decoded label image) and everything above
  // n_classes - 1 become the void label
  const bool known = (present[b * 8 + (lab >> 5)] >> (lab & 31)) & 1u;
  label_out[o] = static_cast<uint8_t>((known && lab < n_classes) ? lab : n_classes);
}

}  // namespace dlb

using namespace dlb;

extern "C" int dlb_augment_batch(int B, int H, int W, int n_classes, const uint8_t* img, const uint8_t* label,
                                 const dlb_aug_params* params_dev, const uint8_t* luts, uint8_t* tmp_img,
                                 uint8_t* tmp_label, unsigned int* present, float* X, uint8_t* label_out, void* stream) {
  DLB_REQUIRE(B > 0 && H > 0 && W > 0 && img && label && params_dev && tmp_img && tmp_label && present && X && label_out,
              "augment_batch: bad arguments");
  DLB_REQUIRE(n_classes >= 1 && n_classes <= 255, "augment_batch: 1 <= n_classes <= 255");
  DLB_REQUIRE(B <= 65535 && H <= 65535, "augment_batch: batch / height too large for the launch grid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DLB_CUDA(cudaMemsetAsync(present, 0, sizeof(unsigned int) * 8 * B, st));
  const long long npix = static_cast<long long>(H) * W;
  long long blocks = (npix + 255) / 256;
  if (blocks > 64) blocks = 64;
  launch_k(aug_present_kernel, dim3(static_cast<unsigned>(blocks), B), 256, 0, st, npix, label, present);
  const dim3 grid((W + 255) / 256, H, B);
  launch_k(aug_pass1_kernel, grid, 256, 0, st, H, W, img, label, params_dev, luts, tmp_img, tmp_label);
  launch_k(aug_pass2_kernel, grid, 256, 0, st, H, W, n_classes, (const uint8_t*)tmp_img, (const uint8_t*)tmp_label, params_dev,
           (const unsigned int*)present, X, label_out);
  g_launches += 3;
  return check_launch("augment_batch");
}

// Frame ingestion on the device: the reference's per-frame CPU transform
//     transforms.Resize((224, 224)) -> transforms.Grayscale(1) -> transforms.ToTensor() -> transforms.Normalize([0.5], [0.5])
// (/root/reference/main.py:103-108, applied frame by frame through PIL in DatasetBase.__getitem__,
// /root/reference/data_loader/data_loader.py:434-446) as byte kernels: uint8 RGB frames cross PCIe as stored (3 bytes per pixel
// instead of 4 bytes per grey pixel AFTER the CPU transform, and no CPU transform at all) and become the encoder's fp32 input here.
// Bit-exact with PIL 8-bit arithmetic:
//   resize     Pillow's ImagingResample (bilinear = triangle filter, support scaled by the down-scaling factor): horizontal pass to
//              uint8, then vertical pass to uint8; fixed-point coefficients with 22 fractional bits, accumulator starts at 2^21,
//              result = clip8(acc >> 22).  The coefficient tables are computed on the host (videocad_b200/ingest.py) exactly as
//              precompute_coeffs / normalize_coeffs_8bpc do and cached per (in, out) size.  A pass whose sizes agree is skipped.
//   grayscale  L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16   (Pillow ImagingConvert rgb2l)
//   ToTensor + Normalize   (float(L) / 255 - mean) / std, same fp32 operations in the same order as torch
// All kernels are HBM-bound byte movers: one thread per 4 output pixels (grey) / per output byte (resize passes).
#include <cuda_runtime.h>
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"
#include "host_util.h"

namespace vck {

namespace {

inline cudaStream_t cs(stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
constexpr int RS_BITS = 32 - 8 - 2;

__device__ __forceinline__ float gray_norm(uint32_t r, uint32_t g, uint32_t b, float mean, float std) {
  const uint32_t l = (r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16;
  return ((float)l / 255.0f - mean) / std;
}

// same-size path: 4 pixels (12 bytes = three 32-bit loads) per thread, one 128-bit store
__global__ void rgb_u8_gray_norm_kernel(const uint8_t* __restrict__ src, long long n_px, float mean, float std, float* __restrict__ dst) {
  pdl_grid_sync();
  const long long n4 = n_px >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(src + 12 * i);
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];  // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
    float4 o;
    o.x = gray_norm(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, mean, std);
    o.y = gray_norm(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u, mean, std);
    o.z = gray_norm((w1 >> 16) & 255u, w1 >> 24, w2 & 255u, mean, std);
    o.w = gray_norm((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24, mean, std);
    *reinterpret_cast<float4*>(dst + 4 * i) = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = n4 * 4; i < n_px; ++i) dst[i] = gray_norm(src[3 * i], src[3 * i + 1], src[3 * i + 2], mean, std);
  }
}

__device__ __forceinline__ uint32_t clip8(int v) { return (uint32_t)min(max(v, 0), 255); }

// horizontal pass: src [n, H, Win, 3] -> dst [n, H, Wout, 3]; one thread per output byte
__global__ void resample_h_kernel(const uint8_t* __restrict__ src, long long rows, int Win, int Wout, const int* __restrict__ kk,
                                  const int* __restrict__ bounds, int ks, uint8_t* __restrict__ dst) {
  pdl_grid_sync();
  const long long total = rows * Wout * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    const long long t = i / 3;
    const int xx = (int)(t % Wout);
    const long long row = t / Wout;
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const uint8_t* p = src + (row * Win + xmin) * 3 + c;
    const int* k = kk + (long long)xx * ks;
    int acc = 1 << (RS_BITS - 1);
    for (int x = 0; x < xmax; ++x) acc += (int)p[3 * x] * k[x];
    dst[i] = (uint8_t)clip8(acc >> RS_BITS);
  }
}

// vertical pass fused with grayscale + normalise: src [n, Hin, W, 3] -> dst fp32 [n, Hout, W]; one thread per output pixel
__global__ void resample_v_gray_norm_kernel(const uint8_t* __restrict__ src, long long n, int Hin, int Hout, int W, const int* __restrict__ kk,
                                            const int* __restrict__ bounds, int ks, float mean, float std, float* __restrict__ dst) {
  pdl_grid_sync();
  const long long total = n * Hout * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long t = i / W;
    const int yy = (int)(t % Hout);
    const long long f = t / Hout;
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    const uint8_t* p = src + ((f * Hin + ymin) * W + x) * 3;
    const int* k = kk + (long long)yy * ks;
    int a0 = 1 << (RS_BITS - 1), a1 = a0, a2 = a0;
    for (int y = 0; y < ymax; ++y) {
      const uint8_t* q = p + (long long)y * W * 3;
      a0 += (int)q[0] * k[y]; a1 += (int)q[1] * k[y]; a2 += (int)q[2] * k[y];
    }
    dst[i] = gray_norm(clip8(a0 >> RS_BITS), clip8(a1 >> RS_BITS), clip8(a2 >> RS_BITS), mean, std);
  }
}

int grid_for(long long work, int block) {
  long long g = (work + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int frames_rgb_u8_ingest(const uint8_t* src, int64_t n, int Hin, int Win, int Hout, int Wout, const int* kk_h, const int* bounds_h, int ks_h,
                         const int* kk_v, const int* bounds_v, int ks_v, uint8_t* tmp, float mean, float std, float* dst, stream_t s) {
  if (!src || !dst || n <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) return set_error("frames_rgb_u8_ingest: bad arguments");
  if (std == 0.f) return set_error("frames_rgb_u8_ingest: std must be non-zero");
  const uint8_t* cur = src;
  if (Win != Wout) {
    if (!kk_h || !bounds_h || ks_h <= 0 || !tmp) return set_error("frames_rgb_u8_ingest: horizontal pass needs coefficients and a temporary");
    const long long rows = (long long)n * Hin;
    VC_LAUNCH((resample_h_kernel), grid_for(rows * Wout * 3, 256), 256, 0, cs(s), src, rows, Win, Wout, kk_h, bounds_h, ks_h, tmp);
    if (int rc = check_launch("resample_h_kernel")) return rc;
    cur = tmp;
  }
  if (Hin != Hout) {
    if (!kk_v || !bounds_v || ks_v <= 0) return set_error("frames_rgb_u8_ingest: vertical pass needs coefficients");
    VC_LAUNCH((resample_v_gray_norm_kernel), grid_for((long long)n * Hout * Wout, 256), 256, 0, cs(s), cur, (long long)n, Hin, Hout, Wout, kk_v,
              bounds_v, ks_v, mean, std, dst);
    return check_launch("resample_v_gray_norm_kernel");
  }
  if ((reinterpret_cast<uintptr_t>(cur) & 3) != 0 || (reinterpret_cast<uintptr_t>(dst) & 15) != 0)
    return set_error("frames_rgb_u8_ingest: src must be 4-byte and dst 16-byte aligned");
  const long long n_px = (long long)n * Hout * Wout;
  VC_LAUNCH((rgb_u8_gray_norm_kernel), grid_for(n_px / 4, 256), 256, 0, cs(s), cur, n_px, mean, std, dst);
  return check_launch("rgb_u8_gray_norm_kernel");
}

}  // namespace vck

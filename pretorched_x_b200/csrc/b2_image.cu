// b2_image.cu -- GPU side of the reference's image preprocessing (pretorched/transforms/utils.py:34-81, TransformImage):
// Pillow-exact 8-bit bilinear resampling (horizontal pass, then vertical, each rounding to uint8 -- libImaging/Resample.c
// semantics, 22-bit fixed-point coefficients computed on the host by pretorched_x_b200/transforms.py), centre / random crop,
// flips, ToTensor (/255), ToSpaceBGR, ToRange255 and Normalize, writing the fp32 NCHW tensor the reference returns and / or
// the fp16 NDHWC4 matrix the stem convolution consumes (so the stand-alone layout pass disappears for image inputs).
// HBM-bound byte work: one thread per output element group, coalesced along the row.
#include "b2_host.h"

#include <cuda_fp16.h>

namespace b2 {

constexpr int kImgPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ unsigned char clip8(int acc) {
  const int v = acc >> kImgPrecisionBits;          // arithmetic shift, like Pillow's clip8 lookup
  return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// dst[y][xx][c] = clip8((1 << 21) + sum_x src[y][xmin + x][c] * k[xx][x])
__global__ void resample_h_u8_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int H, int W, int Wout,
                                     const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)H * Wout * 3;
  if (i >= total) return;
  const int c = (int)(i % 3);
  const int xx = (int)((i / 3) % Wout);
  const int y = (int)(i / (3LL * Wout));
  const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
  const int* k = kk + (size_t)xx * ksize;
  const unsigned char* row = src + ((size_t)y * W + xmin) * 3 + c;
  int acc = 1 << (kImgPrecisionBits - 1);
  for (int x = 0; x < n; ++x) acc += (int)row[3 * x] * k[x];
  dst[i] = clip8(acc);
}

struct ImgFinish {
  int Hs, Ws;              // source (after the horizontal pass) size
  int Hr;                  // rows after the vertical pass
  int top, left, ch, cw;   // crop window in the resized image, output size
  int hflip, vflip, bgr, range255;
  float mean[3], stdv[3];
};

// vertical pass + crop + flips + ToTensor + BGR + range + Normalize for ONE output pixel (3 channels) per thread
__global__ void resample_v_finish_kernel(const unsigned char* __restrict__ src, const int* __restrict__ bounds, const int* __restrict__ kk,
                                         int ksize, const ImgFinish p, float* __restrict__ out_f32, __half* __restrict__ out_h4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.ch * p.cw) return;
  const int ox = i % p.cw, oy = i / p.cw;
  // torchvision applies the flips AFTER the crop: output (oy, ox) shows crop pixel (ch-1-oy / cw-1-ox)
  const int cy = p.vflip ? p.ch - 1 - oy : oy, cx = p.hflip ? p.cw - 1 - ox : ox;
  const int ry = p.top + cy, rx = p.left + cx;
  unsigned char px[3];
  if (bounds) {
    const int ymin = bounds[2 * ry], n = bounds[2 * ry + 1];
    const int* k = kk + (size_t)ry * ksize;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int acc = 1 << (kImgPrecisionBits - 1);
      const unsigned char* col = src + ((size_t)ymin * p.Ws + rx) * 3 + c;
      for (int y = 0; y < n; ++y) acc += (int)col[(size_t)y * p.Ws * 3] * k[y];
      px[c] = clip8(acc);
    }
  } else {
    const unsigned char* s = src + ((size_t)ry * p.Ws + rx) * 3;
    px[0] = s[0]; px[1] = s[1]; px[2] = s[2];
  }
  float v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int sc = p.bgr ? 2 - c : c;                                   // ToSpaceBGR swaps channels 0 and 2 (utils.py:14-20)
    float t = __fdiv_rn((float)px[sc], 255.0f);                         // ToTensor: exact IEEE division, immune to --use_fast_math
    if (p.range255) t = __fmul_rn(t, 255.0f);                           // ToRange255 (utils.py:28-31)
    v[c] = __fdiv_rn(__fsub_rn(t, p.mean[c]), p.stdv[c]);               // Normalize
  }
  if (out_f32) {
    const size_t plane = (size_t)p.ch * p.cw;
    out_f32[i] = v[0]; out_f32[plane + i] = v[1]; out_f32[2 * plane + i] = v[2];
  }
  if (out_h4) {
    const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], 0.f);
    uint2 w;
    w.x = *reinterpret_cast<const unsigned*>(&a); w.y = *reinterpret_cast<const unsigned*>(&b);
    reinterpret_cast<uint2*>(out_h4)[i] = w;
  }
}

// uint8 frames [pixels][3] (decoded video: N*T*H*W RGB pixels, channels last) -> fp16 NDHWC4 with ToTensor (/255), optional BGR
// swap / x255 and Normalize: the clip-side counterpart of resample_v_finish_kernel.  4 pixels (12 bytes in, 32 bytes out) per thread.
__global__ void u8_to_ndhwc4_norm_kernel(const unsigned char* __restrict__ src, __half* __restrict__ dst, long long npx, int bgr,
                                         int range255, float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;      // group of 4 pixels
  const long long p0 = g * 4;
  if (p0 >= npx) return;
  unsigned char b[12];
  if (p0 + 4 <= npx) {
    const uint3 w = *reinterpret_cast<const uint3*>(src + p0 * 3);           // 12-byte aligned: p0 is a multiple of 4
    *reinterpret_cast<uint3*>(b) = w;
  } else {
    for (int i = 0; i < 12; ++i) b[i] = (p0 * 3 + i < npx * 3) ? src[p0 * 3 + i] : 0;
  }
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (p0 + k >= npx) break;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int sc = bgr ? 2 - c : c;
      float t = __fdiv_rn((float)b[k * 3 + sc], 255.0f);
      if (range255) t = __fmul_rn(t, 255.0f);
      v[c] = __fdiv_rn(__fsub_rn(t, mean[c]), stdv[c]);
    }
    const __half2 a = __floats2half2_rn(v[0], v[1]), c2 = __floats2half2_rn(v[2], 0.f);
    uint2 w;
    w.x = *reinterpret_cast<const unsigned*>(&a); w.y = *reinterpret_cast<const unsigned*>(&c2);
    reinterpret_cast<uint2*>(dst)[p0 + k] = w;
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_u8_frames_to_ndhwc4_f16(const uint8_t* frames, void* y, long long pixels, int flags, const float* mean,
                                          const float* stdv, void* stream) {
  B2_CHECK_ARG(frames && y && pixels > 0 && mean && stdv, "bad argument");
  B2_CHECK_ARG((reinterpret_cast<uintptr_t>(frames) & 3) == 0, "frame buffer must be 4-byte aligned");
  const long long groups = (pixels + 3) / 4;
  u8_to_ndhwc4_norm_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      frames, reinterpret_cast<__half*>(y), pixels, (flags >> 2) & 1, (flags >> 3) & 1, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2]);
  B2_CHECK_LAUNCH("u8_to_ndhwc4_norm");
  return B2_OK;
}

extern "C" int b2_transform_image_u8(const uint8_t* img, int H, int W, const int32_t* hbounds, const int32_t* hk, int hksize, int Wr,
                                     const int32_t* vbounds, const int32_t* vk, int vksize, int Hr, uint8_t* tmp, int top, int left,
                                     int crop_h, int crop_w, int flags, const float* mean, const float* stdv, float* out_f32,
                                     void* out_h4, void* stream) {
  B2_CHECK_ARG(img && H > 0 && W > 0 && Wr > 0 && Hr > 0 && crop_h > 0 && crop_w > 0 && mean && stdv, "bad argument");
  B2_CHECK_ARG(out_f32 || out_h4, "no output requested");
  B2_CHECK_ARG((hbounds != nullptr) == (Wr != W) || hbounds, "horizontal coefficient table missing");
  B2_CHECK_ARG(Wr == W || (hbounds && hk && hksize > 0 && tmp), "horizontal pass needs bounds, coefficients and a [H][Wr][3] scratch image");
  B2_CHECK_ARG(Hr == H || (vbounds && vk && vksize > 0), "vertical pass needs bounds and coefficients");
  B2_CHECK_ARG(top >= 0 && left >= 0 && top + crop_h <= Hr && left + crop_w <= Wr, "crop window outside the resized image");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint8_t* src = img;
  if (Wr != W) {
    const long long total = (long long)H * Wr * 3;
    resample_h_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(img, tmp, H, W, Wr, hbounds, hk, hksize);
    B2_CHECK_LAUNCH("resample_h_u8");
    src = tmp;
  }
  ImgFinish p;
  p.Hs = H; p.Ws = Wr; p.Hr = Hr; p.top = top; p.left = left; p.ch = crop_h; p.cw = crop_w;
  p.hflip = flags & 1; p.vflip = (flags >> 1) & 1; p.bgr = (flags >> 2) & 1; p.range255 = (flags >> 3) & 1;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean[c]; p.stdv[c] = stdv[c]; }
  const int px = crop_h * crop_w;
  resample_v_finish_kernel<<<(px + 127) / 128, 128, 0, st>>>(src, Hr != H ? vbounds : nullptr, vk, vksize, p, out_f32,
                                                          reinterpret_cast<__half*>(out_h4));
  B2_CHECK_LAUNCH("resample_v_finish");
  return B2_OK;
}

// b2_gan.cu -- HBM-bound helper kernels of the BigGAN-deep generator path (BASELINE.json configs[4]; the architecture
// is absent from /root/reference, see oracle/biggan.py): conditioning-vector assembly, the class-conditional
// BatchNorm + ReLU (+ nearest 2x upsampling) pass, and the final tanh + NHWC -> NCHW image write.  Same conventions
// as b2_aux.cu: fp16 channels-last rows, 16 bytes per thread per access, channel index fastest.
#include "b2_host.h"

#include <cuda_fp16.h>

namespace b2 {

static inline int gan_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// v[b] = [ table[label[b]][0:ds] | z[b][0:dz] | 0 ... ] (D = round_up(ds + dz, 8) entries).  split == 0: y[b] = fp16(v),
// pitch ldy >= D.  split != 0: y[b] = [ hi | lo | hi ] with hi = fp16(v), lo = fp16(v - hi), pitch ldy >= 3 * D: against a
// weight matrix stored as [ W_hi | W_hi | W_lo ] one fp16 GEMM of K = 3 * D evaluates v . W to ~fp32 accuracy (the
// dropped lo x lo term is 2^-22 relative), which the stacked class-conditional BatchNorm gains need.
__global__ void embed_concat_kernel(const float* __restrict__ z, const long long* __restrict__ labels,
                                    const float* __restrict__ table, const float* __restrict__ embedded, __half* __restrict__ y,
                                    int B, int dz, int ds, int n_classes, int D, int ldy, int split) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, c = i - b * D;
  float v = 0.f;
  if (c < ds) {
    if (embedded) v = embedded[(long long)b * ds + c];
    else {
      long long l = labels[b];
      l = l < 0 ? 0 : (l >= n_classes ? n_classes - 1 : l);
      v = table[l * ds + c];
    }
  } else if (c < ds + dz) {
    v = z[(long long)b * dz + (c - ds)];
  }
  const __half hi = __float2half_rn(v);
  __half* row = y + (long long)b * ldy;
  row[c] = hi;
  if (split) {
    row[D + c] = __float2half_rn(v - __half2float(hi));
    row[2 * D + c] = hi;
  }
}

// y[n, up*h + i, up*w + j, c] = act(x[n, h, w, c] * scale[n*lda + c] + shift[n*lda + c]),  i, j < up, c < C;
// channels [C, ldy) of y are written as zero.  scale == nullptr: identity (pure channel-slice / upsample copy).
//
// Thread (tx, ty) owns one 16-byte channel chunk (c8 = tx, tx + blockDim.x, ...) of kCcbnPix consecutive "pixel lanes":
// the 8 + 8 affine floats of its chunk stay in registers across pixels (they only change with the sample), so the
// kernel moves 16 B in + 16 B x up^2 out per iteration instead of re-reading 64 B of scale/shift for every 16 B of data.
// blockDim.x = min(row chunks rounded up to a power of two, 32) keeps every warp on whole 128-byte lines.
constexpr int kCcbnPix = 8;

template <int UP>
__global__ void __launch_bounds__(256)
ccbn_act_kernel(const __half* __restrict__ x, int ldx8, __half* __restrict__ y, int ldy8, const float* __restrict__ scale,
                const float* __restrict__ shift, int lda, int H, int W, int C8, int relu, long long pixels) {
  const int HW = H * W;
  const long long p0 = ((long long)blockIdx.x * blockDim.y + threadIdx.y) * kCcbnPix;
  if (p0 >= pixels) return;
  const long long Wo = (long long)W * UP;
  const long long n0 = p0 / HW;
  const int rem0 = (int)(p0 - n0 * HW);
  const int h0 = rem0 / W, w0 = rem0 - h0 * W;
  for (int c8 = threadIdx.x; c8 < ldy8; c8 += blockDim.x) {
    // all loads of the thread's pixels are issued before any is consumed (8 x 16 B in flight per thread)
    uint4 v[kCcbnPix];
#pragma unroll
    for (int j = 0; j < kCcbnPix; ++j) {
      v[j] = make_uint4(0, 0, 0, 0);
      if (c8 < C8 && p0 + j < pixels) v[j] = __ldg(reinterpret_cast<const uint4*>(x) + (p0 + j) * ldx8 + c8);
    }
    float sc[8], sh[8];
    long long n_cur = -1, n = n0;
    int h = h0, w = w0 - 1;
#pragma unroll
    for (int j = 0; j < kCcbnPix; ++j) {
      const long long q = p0 + j;
      if (q < pixels) {
        if (++w == W) { w = 0; if (++h == H) { h = 0; ++n; } }      // (n, h, w) of pixel q, without divisions
        uint4 out = v[j];
        if (c8 < C8) {
          if (scale) {
            if (n != n_cur) {
              n_cur = n;
              const float4* sp = reinterpret_cast<const float4*>(scale + n * lda + c8 * 8);
              const float4* tp = reinterpret_cast<const float4*>(shift + n * lda + c8 * 8);
              const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1), t0 = __ldg(tp), t1 = __ldg(tp + 1);
              sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
              sh[0] = t0.x; sh[1] = t0.y; sh[2] = t0.z; sh[3] = t0.w; sh[4] = t1.x; sh[5] = t1.y; sh[6] = t1.z; sh[7] = t1.w;
            }
            const uint32_t in[4] = {out.x, out.y, out.z, out.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&in[e]));
              float a0 = fmaf(f.x, sc[2 * e], sh[2 * e]);
              float a1 = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
              if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
              const __half2 hv = __floats2half2_rn(a0, a1);
              o[e] = *reinterpret_cast<const uint32_t*>(&hv);
            }
            out = make_uint4(o[0], o[1], o[2], o[3]);
          } else if (relu) {
            __half2* hp = reinterpret_cast<__half2*>(&out);
            const __half2 zero = __floats2half2_rn(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < 4; ++e) hp[e] = __hmax2(hp[e], zero);
          }
        }
        uint4* yo = reinterpret_cast<uint4*>(y) + ((n * H * UP + (long long)h * UP) * Wo + (long long)w * UP) * ldy8 + c8;
#pragma unroll
        for (int a = 0; a < UP; ++a)
#pragma unroll
          for (int b = 0; b < UP; ++b) yo[(a * Wo + b) * ldy8] = out;
      }
    }
  }
}

// phase-folded filters of conv3x3(nearest_up2(x)): out[k][ph*4 + 2a + b][c] = sum_{dh in R(py,a)} sum_{dw in R(px,b)} w[k][c][dh][dw]
// with R(0,0) = {0}, R(0,1) = {1,2}, R(1,0) = {0,1}, R(1,1) = {2}  (see b2_pack_upconv3x3_weight in the header)
__global__ void pack_upconv3x3_kernel(const float* __restrict__ w, __half* __restrict__ out, int K, int Cin, int C, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int t = (int)((i / C) % 16);
  const int k = (int)(i / ((long long)C * 16));
  const int ph = t >> 2, a = (t >> 1) & 1, b = t & 1, py = ph >> 1, px = ph & 1;
  const int h_lo = (py == 0) ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), h_hi = (py == 0) ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
  const int w_lo = (px == 0) ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), w_hi = (px == 0) ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
  float v = 0.f;
  if (c < Cin)
    for (int dh = h_lo; dh <= h_hi; ++dh)
      for (int dw = w_lo; dw <= w_hi; ++dw) v += w[(((long long)k * Cin + c) * 3 + dh) * 3 + dw];
  out[i] = __float2half_rn(v);
}

// y[n][c][s] = tanh(x[n*S + s][c])  (fp16 NHWC with pitch ldx -> NCHW, fp32 or fp16): 32 x 32 smem transpose is not
// needed for C = 3; each thread handles one pixel (one 16-byte load) and writes C scalars to C planes, consecutive
// threads -> consecutive addresses within each plane.
template <typename TOut>
__global__ void tanh_nhwc_to_nchw_kernel(const __half* __restrict__ x, int ldx, TOut* __restrict__ y, int C, long long S,
                                         long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / S, s = i - n * S;
  const __half* xr = x + i * ldx;
  for (int c0 = 0; c0 < C; c0 += 8) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(xr + c0));
    const __half* hv = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (c0 + e < C) {
        // tanh(v) = 1 - 2 / (exp(2v) + 1); |v| clamped so that exp stays finite (tanh(15) == 1 in fp32)
        const float v2 = 2.f * fminf(fmaxf(__half2float(hv[e]), -15.f), 15.f);
        y[(n * C + c0 + e) * S + s] = static_cast<TOut>(1.f - __fdividef(2.f, __expf(v2) + 1.f));
      }
  }
}

// RGB head, second half: the 3x3 convolution to K <= 4 channels was split into a 1x1 GEMM that produces, for every
// pixel q and tap (dh, dw), the partial products P[q][(dh*3 + dw)*4 + k] = sum_c x[q][c] * w[k][c][dh][dw] (one tensor
// core pass over x with 36 useful columns instead of 9 shifted passes with 3), and this gather:
//   y[n][k][h][w] = tanh(bias[k] + sum_{dh,dw} P[(n, h + dh - 1, w + dw - 1)][(dh*3 + dw)*4 + k])      (zero padding)
// One thread per output pixel, nine 8-byte loads (a tap's k-quadruple), neighbours share sectors through L1/L2.
template <typename TOut>
__global__ void __launch_bounds__(256)
rgb_head_gather_tanh_kernel(const __half* __restrict__ P, int ldp, const float* __restrict__ bias, TOut* __restrict__ y,
                            int H, int W, int K) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y;
  const long long n = blockIdx.z;
  if (w >= W) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dh = 0; dh < 3; ++dh) {
    const int hh = h + dh - 1;
    if (hh < 0 || hh >= H) continue;
#pragma unroll
    for (int dw = 0; dw < 3; ++dw) {
      const int ww = w + dw - 1;
      if (ww < 0 || ww >= W) continue;
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(P + ((n * H + hh) * W + ww) * ldp + (dh * 3 + dw) * 4));
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
    }
  }
  const long long S = (long long)H * W;
  for (int k = 0; k < K; ++k) {
    const float v2 = 2.f * fminf(fmaxf(acc[k] + __ldg(bias + k), -15.f), 15.f);
    y[(n * K + k) * S + (long long)h * W + w] = static_cast<TOut>(1.f - __fdividef(2.f, __expf(v2) + 1.f));
  }
}

}  // namespace b2

using namespace b2;

extern "C" {

int b2_embed_concat(const float* z, const long long* labels, const float* table, const float* embedded, void* y, int B,
                    int dz, int ds, int n_classes, int ldy, int split, void* stream) {
  B2_CHECK_ARG(z && y && (embedded || (labels && table)), "null pointer");
  const int D = (dz + ds + 7) / 8 * 8;
  B2_CHECK_ARG(B > 0 && dz >= 0 && ds >= 0 && ldy % 8 == 0 && ldy >= (split ? 3 * D : D), "bad dimensions");
  embed_concat_kernel<<<gan_div_up((long long)B * D, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      z, labels, table, embedded, (__half*)y, B, dz, ds, n_classes, D, ldy, split);
  B2_CHECK_LAUNCH("embed_concat");
  return B2_OK;
}

int b2_pack_upconv3x3_weight(const float* w_oihw, void* w_packed, int K, int Cin, int C, void* stream) {
  B2_CHECK_ARG(w_oihw && w_packed && K > 0 && Cin > 0 && C >= Cin && C % 8 == 0, "bad argument");
  const long long total = (long long)K * 16 * C;
  pack_upconv3x3_kernel<<<gan_div_up(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w_oihw, (__half*)w_packed, K, Cin, C, total);
  B2_CHECK_LAUNCH("pack_upconv3x3");
  return B2_OK;
}

int b2_ccbn_act_ndhwc(const void* x, int ldx, void* y, int ldy, const float* scale, const float* shift, int lda, int N,
                      int H, int W, int C, int up, int relu, void* stream) {
  B2_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0, "bad argument");
  B2_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && C % 8 == 0 && ldx >= C && ldy >= C, "channel counts / pitches must be multiples of 8");
  B2_CHECK_ARG((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  B2_CHECK_ARG(lda % 4 == 0 && (reinterpret_cast<uintptr_t>(scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(shift) & 15) == 0,
               "affine arrays must be 16-byte aligned with a pitch that is a multiple of 4");
  if (up != 1 && up != 2) return set_error(B2_ERR_UNSUPPORTED, "nearest upsampling by %d is not implemented (1 or 2)", up);
  const long long pixels = (long long)N * H * W;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int bx = 1;
  while (bx < ldy / 8 && bx < 32) bx *= 2;
  const dim3 block(bx, 256 / bx);
  const int grid = gan_div_up(pixels, (long long)block.y * kCcbnPix);
  if (up == 1)
    ccbn_act_kernel<1><<<grid, block, 0, st>>>((const __half*)x, ldx / 8, (__half*)y, ldy / 8, scale, shift, lda, H, W, C / 8, relu, pixels);
  else
    ccbn_act_kernel<2><<<grid, block, 0, st>>>((const __half*)x, ldx / 8, (__half*)y, ldy / 8, scale, shift, lda, H, W, C / 8, relu, pixels);
  B2_CHECK_LAUNCH("ccbn_act");
  return B2_OK;
}

int b2_rgb_head_gather_tanh(const void* partial, int ldp, const float* bias, void* y, int N, int H, int W, int K, int out_f32,
                            void* stream) {
  B2_CHECK_ARG(partial && bias && y && N > 0 && H > 0 && W > 0 && K >= 1 && K <= 4, "bad argument");
  B2_CHECK_ARG(ldp >= 36 && ldp % 4 == 0 && (reinterpret_cast<uintptr_t>(partial) & 7) == 0, "partial products need a pitch >= 36, multiple of 4");
  B2_CHECK_ARG(H <= 65535 && N <= 65535, "image height / batch exceed the grid limits");
  const dim3 block(W >= 256 ? 256 : (W + 31) / 32 * 32), grid((W + block.x - 1) / block.x, H, N);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_f32)
    rgb_head_gather_tanh_kernel<float><<<grid, block, 0, st>>>((const __half*)partial, ldp, bias, (float*)y, H, W, K);
  else
    rgb_head_gather_tanh_kernel<__half><<<grid, block, 0, st>>>((const __half*)partial, ldp, bias, (__half*)y, H, W, K);
  B2_CHECK_LAUNCH("rgb_head_gather_tanh");
  return B2_OK;
}

int b2_tanh_nhwc_to_nchw(const void* x, int ldx, void* y, int N, int C, long long S, int out_f32, void* stream) {
  B2_CHECK_ARG(x && y && N > 0 && C > 0 && S > 0 && ldx % 8 == 0 && ldx >= C, "bad argument");
  const long long total = (long long)N * S;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_f32)
    tanh_nhwc_to_nchw_kernel<float><<<gan_div_up(total, 256), 256, 0, st>>>((const __half*)x, ldx, (float*)y, C, S, total);
  else
    tanh_nhwc_to_nchw_kernel<__half><<<gan_div_up(total, 256), 256, 0, st>>>((const __half*)x, ldx, (__half*)y, C, S, total);
  B2_CHECK_LAUNCH("tanh_nhwc_to_nchw");
  return B2_OK;
}

}  // extern "C"

// b2_stemconv.cuh -- stem convolutions (kt x kh x 7, stride (1, sh, 2), padding (pt, ph, 3), Cin <= 4) as an
// implicit GEMM whose im2col matrix is never built: it is *described*.
//
// Input is NDHWC4 (8 bytes per pixel).  For a fixed temporal/vertical tap (dt, dh) the 7 horizontal taps of
// output pixel wo read the 8 consecutive pixels [2wo-4, 2wo+4) (pixel 2wo-4 carries a zero weight; it only keeps
// the run 16-byte aligned): 32 fp16 = 64 contiguous bytes, and the run of output pixel wo+1 starts exactly
// 16 bytes later.  A K-major SWIZZLE_NONE tcgen05 operand is addressed as
//     byte(row r, 16-byte chunk j) = start + (r % 8) * 16 + (r / 8) * SBO + j * LBO,
// so with LBO = 16 and SBO = 128 the 128 x 32 im2col tile of one (dt, dh) tap pair is an overlapping (Toeplitz)
// view of the raw input row sitting in shared memory (verified by tools/probe_umma.py, mode 1).  One TMA box per
// temporal tap brings 2G+5 input rows (zero-filled outside the image) for G output rows -- the tensor map treats a
// pixel as one 8-byte element so that a box row is a single 2 KB run (16-byte-wide boxes starve the TMA unit); every (dt, dh) tap is then
// two K=16 MMAs per output row straight out of that slab.  Activation traffic from L2 drops ~4.4x versus
// gathering 64 bytes per output pixel per tap, and no thread touches the data on its way to the tensor core.
//
// CTA = G output rows x 120 output columns x BN output channels (G accumulators of BN fp32 columns in TMEM).
// Warps 0-3: epilogue (BN + ReLU -> fp16 NDHWC), warp 4: TMA/bulk-copy producer, warp 5: MMA issuer.
#pragma once

#include "b2_ptx.cuh"

namespace b2 {

constexpr int kStemThreads = 192;
constexpr int kStemPitch = 2048;        // bytes per slab row: 256 pixels [2*w0-4, 2*w0+252) of 8 bytes
constexpr int kStemRowPx = 256;         // TMA box width (the maximum box extent), one 8-byte element per pixel
constexpr int kStemTileW = 120;         // output columns per CTA: rows r < 124 of the 128-row MMA tile see complete runs

struct StemParams {
  int T, H, W;             // input dims per clip (W even)
  int To, Ho, Wo;
  int kt, kh;              // kw == 7
  int sh;                  // vertical stride (horizontal stride is 2, temporal stride 1)
  int pt, ph;
  int G;                   // output rows per CTA
  int rows;                // slab rows = sh*(G-1) + kh
  int stage_bytes;         // slab + weights of one temporal tap, multiple of 128
  int w_bytes;             // kh * BN * 64: weight image of one temporal tap for this N tile
  int nstages;
  int Ncols;
  const __half* wimg;      // packed weight image [ntile][kt*kh][BN/8][4][8][8]
  const float* scale;
  const float* shift;
  __half* y;               // [N*To*Ho*Wo][ldy]
  int ldy;
  int relu;
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kStemThreads, 1)
stemconv_kernel(const __grid_constant__ CUtensorMap tmX,   // input as 8-byte pixels (W, H, N*T), box (256, rows, 1), no swizzle
                const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* tail = smem + p.nstages * p.stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);          // [nstages]
  uint64_t* empty = full + 4;                                  // [nstages]
  uint64_t* tmem_full = empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_scale = reinterpret_cast<float*>(tail + 128);
  float* s_shift = s_scale + BN;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int w0 = blockIdx.x * kStemTileW;             // first output column of the tile
  const int ho0 = blockIdx.y * p.G;
  const int ntile = blockIdx.z % ((p.ldy + BN - 1) / BN);
  const int plane_o = blockIdx.z / ((p.ldy + BN - 1) / BN);     // n*To + to   (temporal stride 1 => To == T)
  const int n0 = ntile * BN;
  const int to = plane_o % p.To, n = plane_o / p.To;
  const int dt_lo = max(0, p.pt - to), dt_hi = min(p.kt - 1, p.T - 1 - to + p.pt);
  const int n_dt = dt_hi - dt_lo + 1;
  const int g_valid = min(p.G, p.Ho - ho0);
  const int slab_bytes = p.rows * kStemPitch;

  if (tid == 128) {
    for (int s = 0; s < p.nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  for (int i = tid; i < BN; i += kStemThreads) {
    const int c = n0 + i;
    s_scale[i] = (c < p.Ncols) ? __ldg(&p.scale[c]) : 0.f;
    s_shift[i] = (c < p.Ncols) ? __ldg(&p.shift[c]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    for (int i = 0; i < n_dt; ++i) {
      const int dt = dt_lo + i;
      const int s = i % p.nstages;
      mbar_wait(&empty[s], ((i / p.nstages) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[s], static_cast<uint32_t>(slab_bytes + p.w_bytes));
        uint8_t* dst = smem + s * p.stage_bytes;
        tma_load_3d(dst, &tmX, &full[s], 2 * w0 - 4, p.sh * ho0 - p.ph, n * p.T + to + dt - p.pt);
        const __half* wsrc = p.wimg + (static_cast<size_t>(ntile) * p.kt + dt) * (p.w_bytes / 2);
        bulk_load_1d(dst + slab_bytes, wsrc, static_cast<uint32_t>(p.w_bytes), &full[s]);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    constexpr uint32_t idesc = make_idesc_f16(128, BN, 0);
    constexpr uint32_t kPairBytes = BN * 64;           // weight image of one (dt, dh) pair: BN rows x 32 k
    // SWIZZLE_NONE K-major descriptors: hi = SBO >> 4 | version; lo = addr >> 4 | (LBO >> 4) << 16
    constexpr uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = (512u >> 4) | (1u << 14);
    const uint32_t tm = warp_uniform(tmem_base);
    const uint32_t base = smem_u32(smem);
    for (int i = 0; i < n_dt; ++i) {
      const int s = i % p.nstages;
      mbar_wait(&full[s], (i / p.nstages) & 1);
      tc_fence_after();
      const uint32_t slab = base + s * p.stage_bytes;
      const uint32_t wbase = slab + slab_bytes;
      if (elect_one()) {
        for (int dh = 0; dh < p.kh; ++dh) {
          const uint32_t b_lo = ((wbase + dh * kPairBytes) >> 4) | ((128u >> 4) << 16);
          for (int g = 0; g < g_valid; ++g) {
            const uint32_t a_lo = ((slab + static_cast<uint32_t>(p.sh * g + dh) * kStemPitch) >> 4) | ((16u >> 4) << 16);
            umma_f16(tm + g * BN, desc_from(a_hi, a_lo), desc_from(b_hi, b_lo), idesc, (i | dh) != 0 ? 1u : 0u);
            umma_f16(tm + g * BN, desc_from(a_hi, a_lo + 2), desc_from(b_hi, b_lo + 16), idesc, 1u);
          }
        }
        umma_commit(&empty[s]);
        if (i == n_dt - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const int wo = w0 + tid;
    const bool col_ok = (tid < kStemTileW) && (wo < p.Wo);
    const int ncols_here = min(BN, p.ldy - n0);
    for (int g = 0; g < g_valid; ++g) {
      const size_t row = (static_cast<size_t>(plane_o) * p.Ho + (ho0 + g)) * p.Wo + wo;
      __half* yrow = p.y + row * p.ldy + n0;
#pragma unroll 1
      for (int jc = 0; jc < BN / 32; ++jc) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + g * BN + jc * 32, v);
        tmem_ld_wait();
        if (col_ok) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const int col = jc * 32 + c8 * 8;
            if (col < ncols_here) {
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int ci = col + e * 2;
                float a0 = __uint_as_float(v[c8 * 8 + e * 2]) * s_scale[ci] + s_shift[ci];
                float a1 = __uint_as_float(v[c8 * 8 + e * 2 + 1]) * s_scale[ci + 1] + s_shift[ci + 1];
                if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                o[e] = pack_half2(a0, a1);
              }
              *reinterpret_cast<uint4*>(yrow + col) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

}  // namespace b2

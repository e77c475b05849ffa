// Backward hot kernel (recompute form).  For X [nx,d], Y [ny,d] (bf16) and c = s*log2(e):
//   dX[i,:] += sum_j G[i,j] * Y[j,:],   G[i,j] = 2^(c x_i.y_j - off_row[i]) + 2^(c x_i.y_j - off_col[j])
// The logit tile S (128 x 256, fp32) is recomputed into tensor memory, turned into the bf16
// probability tile G in place (tcgen05.ld -> ex2 -> tcgen05.st), and consumed straight from
// tensor memory as the A operand of the second MMA (dX += G * Y, B = Y tile, MN-major).
// Neither S nor G ever leaves the SM.
//
// One CTA owns (128 rows of X) x (256 columns of dX) x (a contiguous range of Y tiles):
//   TMEM columns [0,256)   fp32 dX accumulators
//   TMEM columns [256,512) S tile; G (packed bf16) overwrites its first 128 columns
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
// Work is linearised (x-tile, d-chunk, y-tile) and cut into equal contiguous ranges, one per
// CTA; a CTA that stops in the middle of a (x-tile, d-chunk) item adds its partial accumulator
// into dX with fp32 reductions (red.global.add), so dX must be zero-initialised by the caller.
#include <cstdlib>

#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace mmcl {

namespace {

constexpr int BM = 128;   // X rows per tile
constexpr int BNY = 256;  // Y rows per step (UMMA N of the S MMA, K extent of the dX MMA)
constexpr int BK = 64;
constexpr int DC = 256;   // dX columns per CTA
constexpr int XCHUNK_BYTES = BM * BK * 2;   // 16 KB
constexpr int YBOX_BYTES = BNY * BK * 2;    // 32 KB: one [256 y, 64 d] box (S phase)
constexpr int YSUB_BYTES = 64 * BK * 2;     // 8 KB : one [64 y, 64 d] box (dX phase)
constexpr float LOG2E = 1.4426950408889634f;

struct BwdParams {
  int nx, ny, d, kchunks;
  int nxt, ndc, nyt;
  long long total;
  int per_cta;
  int nstages;
  const float* scale;
  const float* xinv;  // [nx] 1/||x_i|| or null
  const float* yinv;  // [ny] 1/||y_j|| or null
  const float* off_row;
  const float* off_col;
  float off_mul, row_add, col_add;  // offset = off_*[i] * off_mul + *_add (lets callers pass raw lse)
  float* dx;
};

// cy[e] = (off_col[j], 1/||y_j||) for the 32 columns of the chunk.  The logit is
// s * (x_i.y_j) * xinv_i * yinv_j; c = s*log2e*xinv_i.  G is multiplied by yinv_j so that the
// second MMA can use the raw (un-normalised) Y tile.
template <int MODE>  // 1: row term, 2: column term, 3: both
__device__ __forceinline__ void g_chunk(const uint32_t (&v)[32], uint32_t (&g)[16], float c,
                                        float offr, const float2* cy) {
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    const float4 f = *reinterpret_cast<const float4*>(cy + e);  // (offc0, yinv0, offc1, yinv1)
    const float x0 = __uint_as_float(v[e]) * f.y, x1 = __uint_as_float(v[e + 1]) * f.w;
    float g0 = 0.f, g1 = 0.f;
    if (MODE & 1) {
      g0 = ex2f(fmaf(x0, c, -offr));
      g1 = ex2f(fmaf(x1, c, -offr));
    }
    if (MODE & 2) {
      g0 += ex2f(fmaf(x0, c, -f.x));
      g1 += ex2f(fmaf(x1, c, -f.z));
    }
    g[e >> 1] = pack_bf16x2(g0 * f.y, g1 * f.w);
  }
}

// Fast form of MODE 3 (one ex2 per logit instead of two):
//   2^(L-offr_i) + 2^(L-offc_j) = 2^(L-offr_i) * (1 + R_i*C_j),  R_i = 2^(offr_i-mu), C_j = 2^(mu-offc_j)
// valid when every offset lies within +-kFastRange of mu (checked per launch, see the kernel
// prologue); cy[e] = (C_j * yinv_j, yinv_j).
constexpr float kFastRange = 40.f;
__device__ __forceinline__ void g_chunk_fast(const uint32_t (&v)[32], uint32_t (&g)[16], float c,
                                             float offr, float R, const float2* cy) {
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    const float4 f = *reinterpret_cast<const float4*>(cy + e);  // (cyB0, yinv0, cyB1, yinv1)
    const float e0 = ex2f(fmaf(__uint_as_float(v[e]) * f.y, c, -offr));
    const float e1 = ex2f(fmaf(__uint_as_float(v[e + 1]) * f.w, c, -offr));
    g[e >> 1] = pack_bf16x2(e0 * fmaf(R, f.x, f.y), e1 * fmaf(R, f.z, f.w));
  }
}

template <bool kXRes>
__global__ void __launch_bounds__(256, 1)
sim_grad_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY256,
                    const __grid_constant__ CUtensorMap tmY64, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operands need 1024-byte alignment
  const int stage_bytes = kXRes ? YBOX_BYTES : XCHUNK_BYTES + YBOX_BYTES;
  uint8_t* sX = smem;
  uint8_t* sStage = smem + (kXRes ? p.kchunks * XCHUNK_BYTES : 0);
  float2* sCol = reinterpret_cast<float2*>(sStage + p.nstages * stage_bytes);  // [256] (offc, yinv)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sCol + BNY);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.nstages;
  uint64_t* x_full = bars + 2 * p.nstages;
  uint64_t* x_empty = x_full + 1;
  uint64_t* s_full = x_full + 2;
  uint64_t* g_full = x_full + 3;  // [2]: G of key half 0 / 1 is in tensor memory
  uint64_t* acc_full = x_full + 5;
  uint64_t* acc_empty = x_full + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 7);
  float* s_range = reinterpret_cast<float*>(x_full + 8);  // [16] lo per warp, [16] hi per warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY256);
    tma_prefetch_desc(&tmY64);
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(&g_full[0], 128);
    mbar_init(&g_full[1], 128);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    fence_barrier_init();
  }
  const int mode = (p.off_row ? 1 : 0) | (p.off_col ? 2 : 0);
  if (mode == 3) {
    // range of all offsets: decides whether the one-exponential form of G is safe
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < p.nx; i += 256) {
      const float v = fmaf(p.off_row[i], p.off_mul, p.row_add);
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
    for (int i = threadIdx.x; i < p.ny; i += 256) {
      const float v = fmaf(p.off_col[i], p.off_mul, p.col_add);
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      s_range[warp] = lo;
      s_range[16 + warp] = hi;
    }
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_acc = tmem, tm_s = tmem + 256;

  const long long t_begin = (long long)blockIdx.x * p.per_cta;
  long long t_end = t_begin + p.per_cta;
  if (t_end > p.total) t_end = p.total;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, xphase = 0;
      for (long long t = t_begin; t < t_end;) {
        const long long item = t / p.nyt;
        const int yt0 = int(t % p.nyt);
        int yt1 = yt0 + int(t_end - t);
        if (yt1 > p.nyt) yt1 = p.nyt;
        const int xt = int(item / p.ndc), dc = int(item % p.ndc);
        int nb = (p.d - dc * DC + BK - 1) / BK;  // 64-wide column boxes in this d-chunk
        if (nb > DC / BK) nb = DC / BK;
        if (kXRes) {
          mbar_wait(x_empty, xphase ^ 1);
          mbar_arrive_expect_tx(x_full, p.kchunks * XCHUNK_BYTES);
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_2d(sX + kc * XCHUNK_BYTES, &tmX, kc * BK, xt * BM, x_full);
          xphase ^= 1;
        }
        for (int yt = yt0; yt < yt1; ++yt) {
          for (int kc = 0; kc < p.kchunks; ++kc) {  // S phase
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], stage_bytes);
            uint8_t* dst = sStage + stage * stage_bytes;
            if (!kXRes) {
              tma_load_2d(dst, &tmX, kc * BK, xt * BM, &full[stage]);
              dst += XCHUNK_BYTES;
            }
            tma_load_2d(dst, &tmY256, kc * BK, yt * BNY, &full[stage]);
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          }
          for (int js = 0; js < BNY / 64; ++js) {  // dX phase: [64 y, nb*64 d] per stage
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], nb * YSUB_BYTES);
            uint8_t* dst = sStage + stage * stage_bytes;
            for (int b = 0; b < nb; ++b)
              tma_load_2d(dst + b * YSUB_BYTES, &tmY64, dc * DC + b * BK, yt * BNY + js * 64,
                          &full[stage]);
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          }
        }
        t += yt1 - yt0;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(BM, BNY, 0, 0);
      int stage = 0;
      uint32_t phase = 0, xphase = 0, step = 0, seg = 0;
      for (long long t = t_begin; t < t_end;) {
        const long long item = t / p.nyt;
        const int yt0 = int(t % p.nyt);
        int yt1 = yt0 + int(t_end - t);
        if (yt1 > p.nyt) yt1 = p.nyt;
        const int dc = int(item % p.ndc);
        int ncols = p.d - dc * DC;
        if (ncols > DC) ncols = DC;
        ncols = (ncols + 63) & ~63;
        const uint32_t idesc_dx = make_idesc_bf16(BM, ncols, 0, 1);
        if (kXRes) {
          mbar_wait(x_full, xphase);
          xphase ^= 1;
        }
        // accumulators of the previous segment must have been drained
        mbar_wait(acc_empty, (seg & 1) ^ 1);
        tc_fence_after();
        for (int yt = yt0; yt < yt1; ++yt) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(sStage + stage * stage_bytes);
            const uint32_t xa = kXRes ? smem_u32(sX + kc * XCHUNK_BYTES) : st;
            const uint32_t ya = kXRes ? st : st + XCHUNK_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss(tm_s, make_smem_desc_sw128(xa + k * 32, 0, 1024),
                      make_smem_desc_sw128(ya + k * 32, 0, 1024), idesc_s, (kc | k) != 0);
            umma_commit(&empty[stage]);
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          }
          umma_commit(s_full);
          if (kXRes && yt == yt1 - 1) umma_commit(x_empty);
          for (int js = 0; js < BNY / 64; ++js) {
            if ((js & 1) == 0) {  // G of this key half has been written to tensor memory
              mbar_wait(&g_full[js >> 1], step & 1);
              tc_fence_after();
            }
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(sStage + stage * stage_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ts(tm_acc, tm_s + (js * 4 + k) * 8,
                      make_smem_desc_sw128(st + k * 2048, YSUB_BYTES, 1024), idesc_dx,
                      (yt != yt0) || (js | k) != 0);
            umma_commit(&empty[stage]);
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          }
          ++step;
        }
        umma_commit(acc_full);
        ++seg;
        t += yt1 - yt0;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;  // 0..127
    const float c0 = (*p.scale) * LOG2E;
    const uint32_t lane_base = uint32_t(q * 32) << 16;
    float mu = 0.f;
    bool fast = false;
    if (mode == 3) {
      float lo = INFINITY, hi = -INFINITY;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        lo = fminf(lo, s_range[w]);
        hi = fmaxf(hi, s_range[16 + w]);
      }
      mu = 0.5f * (lo + hi);
      fast = (hi - lo) <= 2.f * kFastRange;  // false for inf/NaN ranges as well
    }
    uint32_t step = 0, seg = 0;
    for (long long t = t_begin; t < t_end;) {
      const long long item = t / p.nyt;
      const int yt0 = int(t % p.nyt);
      int yt1 = yt0 + int(t_end - t);
      if (yt1 > p.nyt) yt1 = p.nyt;
      const int xt = int(item / p.ndc), dc = int(item % p.ndc);
      const int grow = xt * BM + row;
      const float offr = (p.off_row && grow < p.nx) ? fmaf(p.off_row[grow], p.off_mul, p.row_add) : INFINITY;
      const float c = c0 * ((p.xinv && grow < p.nx) ? p.xinv[grow] : 1.f);
      const float R = (fast && grow < p.nx) ? ex2f(offr - mu) : 0.f;
      for (int yt = yt0; yt < yt1; ++yt) {
        named_bar_sync(1, 128);  // everyone is done reading the previous step's column data
        for (int j = et; j < BNY; j += 128) {
          const int gy = yt * BNY + j;
          float2 v2;
          v2.x = (p.off_col && gy < p.ny) ? fmaf(p.off_col[gy], p.off_mul, p.col_add) : INFINITY;
          v2.y = (p.yinv && gy < p.ny) ? p.yinv[gy] : 1.f;
          if (fast) v2.x = ex2f(mu - v2.x) * v2.y;
          sCol[j] = v2;
        }
        named_bar_sync(2, 128);
        mbar_wait(s_full, step & 1);
        tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < BNY / 32; ++ch) {
          uint32_t v[32], g[16];
          tmem_ld_x32(tm_s + lane_base + ch * 32, v);
          tmem_ld_wait();
          if (fast)
            g_chunk_fast(v, g, c, offr, R, sCol + ch * 32);
          else if (mode == 3)
            g_chunk<3>(v, g, c, offr, sCol + ch * 32);
          else if (mode == 1)
            g_chunk<1>(v, g, c, offr, sCol + ch * 32);
          else
            g_chunk<2>(v, g, c, offr, sCol + ch * 32);
          tmem_st_x16(tm_s + lane_base + ch * 16, g);
          if ((ch & 3) == 3) {  // one key half (128 columns) of G is complete
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&g_full[ch >> 2]);
          }
        }
        ++step;
      }
      // flush this segment's accumulators
      mbar_wait(acc_full, seg & 1);
      tc_fence_after();
      int ncols = p.d - dc * DC;
      if (ncols > DC) ncols = DC;
#pragma unroll 1
      for (int ch = 0; ch < DC / 32; ++ch) {
        if (ch * 32 >= ncols) break;
        uint32_t v[32];
        tmem_ld_x32(tm_acc + lane_base + ch * 32, v);
        tmem_ld_wait();
        if (grow < p.nx) {
          float* dst = p.dx + (long long)grow * p.d + dc * DC + ch * 32;
          // d % 8 == 0 (checked by the ABI): groups of 4 columns are whole and 16-byte aligned
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            if (ch * 32 + e < ncols) red_add_v4(dst + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
      ++seg;
      t += yt1 - yt0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

}  // namespace

cudaError_t launch_sim_grad_bwd(int num_sms, const void* x, long long nx, const void* y,
                                long long ny, long long d, const float* scale, const float* xinv,
                                const float* yinv, const float* off_row, const float* off_col,
                                float off_mul, float row_add, float col_add, float* dx,
                                cudaStream_t stream, const char** err) {
  if (!off_row && !off_col) return cudaSuccess;  // G == 0
  // CTA-pair form first (every logit tile computed once); MMCL_BWD_SINGLE=1 forces this file's
  // single-CTA kernel (used by the tests to cover both).
  static const bool force_single = [] {
    const char* e = getenv("MMCL_BWD_SINGLE");
    return e && e[0] == '1';
  }();
  if (!force_single) {
    cudaError_t pe = launch_sim_grad_bwd_pair(num_sms, x, nx, y, ny, d, scale, xinv, yinv, off_row,
                                              off_col, off_mul, row_add, col_add, dx, stream, err);
    if (pe != cudaErrorNotSupported) return pe;
  }
  BwdParams p;
  p.nx = int(nx);
  p.ny = int(ny);
  p.d = int(d);
  p.kchunks = int((d + BK - 1) / BK);
  p.nxt = int((nx + BM - 1) / BM);
  p.ndc = int((d + DC - 1) / DC);
  p.nyt = int((ny + BNY - 1) / BNY);
  p.total = (long long)p.nxt * p.ndc * p.nyt;
  long long ctas = p.total < num_sms ? p.total : num_sms;
  p.per_cta = int((p.total + ctas - 1) / ctas);
  ctas = (p.total + p.per_cta - 1) / p.per_cta;
  p.scale = scale;
  p.xinv = xinv;
  p.yinv = yinv;
  p.off_row = off_row;
  p.off_col = off_col;
  p.off_mul = off_mul;
  p.row_add = row_add;
  p.col_add = col_add;
  p.dx = dx;
  const size_t budget = 227 * 1024 - 2 * BNY * sizeof(float) - 512;
  const bool xres = size_t(p.kchunks) * XCHUNK_BYTES + 3 * size_t(YBOX_BYTES) <= budget;
  const size_t stage_bytes = xres ? YBOX_BYTES : XCHUNK_BYTES + YBOX_BYTES;
  const size_t fixed = xres ? size_t(p.kchunks) * XCHUNK_BYTES : 0;
  int ns = int((budget - fixed) / stage_bytes);
  if (ns > 6) ns = 6;
  p.nstages = ns;
  const size_t smem = fixed + ns * stage_bytes + 2 * BNY * sizeof(float) + 512;
  CUtensorMap tmX, tmY256, tmY64;
  if (make_tmap_bf16_sw128(&tmX, x, nx, d, d, BM) ||
      make_tmap_bf16_sw128(&tmY256, y, ny, d, d, BNY) ||
      make_tmap_bf16_sw128(&tmY64, y, ny, d, d, 64)) {
    *err = "sim_grad_bwd: cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  cudaError_t e;
  if (xres) {
    e = cudaFuncSetAttribute(sim_grad_bwd_kernel<true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    sim_grad_bwd_kernel<true><<<int(ctas), 256, smem, stream>>>(tmX, tmY256, tmY64, p);
  } else {
    e = cudaFuncSetAttribute(sim_grad_bwd_kernel<false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    sim_grad_bwd_kernel<false><<<int(ctas), 256, smem, stream>>>(tmX, tmY256, tmY64, p);
  }
  count_launches(1);
  return cudaGetLastError();
}

}  // namespace mmcl

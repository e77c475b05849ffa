// Backward hot kernel, CTA-pair form (no logit recompute per d-chunk).
//
// Same contract as sim_bwd.cu (dX[i,:] += sum_j G[i,j] * yinv_j * Y[j,:]), but the fp32 dX
// accumulator of a 128-row X tile -- 128 x 512 fp32 = one SM's whole tensor memory at d = 512 --
// is split over the two CTAs of a 2-CTA cluster (two SMs): CTA c keeps the columns
// [c*d/2, (c+1)*d/2) in tensor-memory columns [0,256) and has columns [256,512) free for one
// 128 x 256 logit tile.  The pair walks the key tiles in lock step and ALTERNATES who recomputes
// the logit tile: for its own step a CTA runs S = X Y^T (SS MMA, N = 256), turns S into the bf16
// probability tile G in its epilogue warps, keeps G in tensor memory (A operand of its own
// dX += G Y MMA, TS form) and at the same time stores G through distributed shared memory into the
// partner's staging buffer, from where the partner feeds it to ITS dX MMA as a shared-memory A
// operand.  Every logit tile is therefore computed exactly once per pair, and each CTA's tensor
// pipe alternates  S(own) | dX(partner's G)  | dX(own G)  with the epilogue of S(own) hidden
// under dX(partner's G).
//
// Per CTA (384 threads, 1 CTA/SM): warp 0 TMA producer, warp 1 MMA issuer, warps 2-3 G send warps
// (warp 2 also allocates tensor memory), warps 4-7 epilogue of key half 0 (S columns 0-127), warps
// 8-11 epilogue of key half 1.  The epilogue warps hand every converted 32-key chunk to the send
// warps through a 2 KB shared-memory buffer per warp (mbarrier hand-shake); the send warps read it
// back transposed and issue the st.async stream into the partner's staging tile, so the
// distributed-shared-memory back-pressure (18-20 B/clk, tools/dsmem_rate.cu) never stalls the
// conversion.  Shared memory: 3 stages x 48 KB (X chunk 128x64 + Y box 256x64 for S; 64-key x d/2
// boxes for dX), 64 KB staging for the partner's G (K-major, 128-byte swizzle, 4 atoms of 64 keys),
// 16 KB of hand-off buffers, column data, barriers.  tools/bwd_trace.cu (MMCL_TRACE) stamps the SM
// clock of every role; DESIGN.md section 4 has the measured phase times.
// Work: one launch covers a LIST of problems (fused.h, BwdProb: every (pair, side) of the loss).
// Their (x-tile, d-pass, y-tile) spaces are laid end to end, linearised and cut into equal
// contiguous ranges, one per CTA pair.  d-passes exist only for d > 512 (each pass covers
// d/ndc <= 512 columns and recomputes S).  A range that starts inside an (x-tile, d-pass) item
// holds a partial sum of that item:
//   accumulate == 1 (ABI calls): every segment is added into dX with fp32 reductions; the caller
//                   zero-fills dX;
//   accumulate == 0 (fused step): segments that start an item are STORED to dX, the one segment
//                   per CTA pair that continues an item goes to that pair's scratch tile and the
//                   gradient-unpack kernel adds it -- no memset, no atomics, dX written once.
// Several ranks: the column offsets of other ranks' rows (their log-sum-exps) arrive by peer
// stores; the kernel waits for the per-source flags of channel kChanLse before it reads them.
#include <cstdlib>

#include "fused.h"
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace mmcl {

namespace {

constexpr int BM = 128;
constexpr int BNY = 256;
constexpr int BK = 64;
constexpr int XCHUNK_BYTES = BM * BK * 2;              // 16 KB
constexpr int YBOX_BYTES = BNY * BK * 2;               // 32 KB
constexpr int YSUB_BYTES = 64 * BK * 2;                // 8 KB: one [64 y, 64 d] box
constexpr int STAGE_BYTES = XCHUNK_BYTES + YBOX_BYTES; // 48 KB
constexpr int NSTAGES = 3;
constexpr int ATOM_BYTES = BM * 128;                   // 16 KB: 128 rows x 64 keys (bf16)
constexpr int STG_BYTES = 4 * ATOM_BYTES;              // 64 KB: G tile 128 x 256 (bf16)
constexpr int XFER_BYTES = 8 * 2048;                   // per epilogue warp: 32 rows x 64 B
constexpr int TAIL_BYTES = 512;                        // barriers + range scratch
constexpr float LOG2E = 1.4426950408889634f;
constexpr float kFastRange = 40.f;

#ifdef MMCL_TRACE
// Timeline instrumentation (tools/bwd_trace.cu): (event, step, SM clock) of one CTA pair.  Every
// tracing thread (role: 0 TMA producer, 1 MMA issuer, 2/3 first lane of the two epilogue
// warpgroups) appends to its own buffer with a register index -- no atomics, so a stamp costs a
// clock read and two stores and the traced pair stays representative.
#ifndef MMCL_TRACE_PAIR
#define MMCL_TRACE_PAIR 0
#endif
constexpr unsigned kTraceCap = 8192;
__device__ unsigned long long g_trace[2][4][kTraceCap][2];
__device__ __forceinline__ void trace_ev(int pair, uint32_t cta, int role, unsigned& i, unsigned ev,
                                         unsigned step) {
  if (pair != MMCL_TRACE_PAIR || i >= kTraceCap) return;
  g_trace[cta][role][i][0] = (static_cast<unsigned long long>(ev) << 32) | step;
  g_trace[cta][role][i][1] = clock64();
  ++i;
}
#define TRACE(ev, step) trace_ev(pair, c, trace_role, trace_i, ev, unsigned(step))
#else
#define TRACE(ev, step)
#endif
#if defined(MMCL_TRACE) && defined(MMCL_TRACE_FINE)  // stamps inside the conversion loop
#define TRACE_EPI(k) if (lane == 0 && q == 0) TRACE(32 + ci * 8 + (k), own_cnt)
#else
#define TRACE_EPI(k)
#endif

// (x tile, d pass, y tile) of consecutive linear tiles of one problem: one 64-bit division at the
// start of a range, increments after it
struct TileWalk {
  int xt = 0, dc = 0, yt = 0;
  __device__ __forceinline__ void start(long long local, int nyt, int ndc) {
    const long long item = local / nyt;
    yt = int(local - item * nyt);
    xt = int(item / ndc);
    dc = int(item - (long long)xt * ndc);
  }
  __device__ __forceinline__ void next(int nyt, int ndc) {
    if (++yt == nyt) {
      yt = 0;
      if (++dc == ndc) {
        dc = 0;
        ++xt;
      }
    }
  }
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// 16 logits of one row -> 8 packed bf16x2 values of G.  Exact form (see sim_bwd.cu):
// cy[e] = (off_col[j], 1/||y_j||).
template <int MODE>
__device__ __forceinline__ void conv16(const uint32_t (&v)[16], uint32_t* g, float c, float offr,
                                       const float2* cy) {
#pragma unroll
  for (int e = 0; e < 16; e += 2) {
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
// One-exponential form of MODE 3 (see sim_bwd.cu) on packed fp32x2 math (sm_100 FFMA2/FMUL2):
// per pair of logits 2 FMUL2 + 2 FFMA2 + 2 MUFU + 1 cvt.  cy4[i] = (cyB[2i], cyB[2i+1],
// yinv[2i], yinv[2i+1]).
__device__ __forceinline__ void conv16_fast(const uint32_t (&v)[16], uint32_t* g, float c,
                                            float offr, float R, const float4* cy4) {
  const float2 c2 = make_float2(c, c), o2 = make_float2(-offr, -offr), r2 = make_float2(R, R);
#pragma unroll
  for (int e = 0; e < 16; e += 2) {
    const float4 f = cy4[e >> 1];
    const float2 yi = make_float2(f.z, f.w);
    const float2 t = __fmul2_rn(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), yi);
    const float2 l = __ffma2_rn(t, c2, o2);
    const float2 ex = make_float2(ex2f(l.x), ex2f(l.y));
    const float2 w = __ffma2_rn(r2, make_float2(f.x, f.y), yi);
    const float2 gg = __fmul2_rn(ex, w);
    g[e >> 1] = pack_bf16x2(gg.x, gg.y);
  }
}

// BCE form (BwdProb::bce).  cy[e] = (band end of column j as int bits, 1/||y_j||); `jg0` = global
// index of the first of the 16 columns.  Three MUFU per logit (ex2, rcp for the sigmoid, rcp for
// the column coefficient); the alignment term is not the headline path.
__device__ __forceinline__ void conv16_bce(const uint32_t (&v)[16], uint32_t* g, float crow,
                                           long long ig, long long bh_i, float p_i, float n_i,
                                           float tt, long long jg0, long long ny,
                                           const float2* cy) {
#pragma unroll
  for (int e = 0; e < 16; e += 2) {
    float gg[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 f = cy[e + h];
      const long long jg = jg0 + e + h;
      const long long bh_j = (long long)__float_as_int(f.x);
      const float x = __uint_as_float(v[e + h]) * f.y * crow;
      const float u = ex2f(-fabsf(x) * LOG2E);
      const float r = __frcp_rn(1.f + u);
      const float sig = x >= 0.f ? r : u * r;
      const bool t_ij = jg >= ig && jg < bh_i;
      const bool t_ji = ig >= jg && ig < bh_j;
      const float np_j = float(bh_j - jg);                       // |band_j|
      const float a_ji = __frcp_rn(tt * (t_ji ? np_j : tt - np_j));
      const float a_ij = t_ij ? p_i : n_i;
      float gv = sig * (a_ij + a_ji) - ((t_ij ? p_i : 0.f) + (t_ji ? a_ji : 0.f));
      if (jg >= ny) gv = 0.f;
      gg[h] = gv * f.y;
    }
    g[e >> 1] = pack_bf16x2(gg[0], gg[1]);
  }
}

constexpr int NTHREADS = 384;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
sim_grad_bwd_pair_kernel(const __grid_constant__ TensorMaps maps,
                         const __grid_constant__ BwdLaunch L) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operands need 1024-byte alignment
  uint8_t* sStage = smem;
  uint8_t* sStg = smem + NSTAGES * STAGE_BYTES;
  float2* sCol = reinterpret_cast<float2*>(sStg + STG_BYTES);  // [256]
  uint8_t* sXfer = reinterpret_cast<uint8_t*>(sCol + BNY);  // [8 warps][2 KB] transposition
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXfer + XFER_BYTES);
  uint64_t* full = bars;                  // [NSTAGES] TMA -> MMA
  uint64_t* empty = bars + NSTAGES;       // [NSTAGES] MMA -> TMA
  uint64_t* s_full = bars + 2 * NSTAGES;  // S tile complete in tensor memory
  uint64_t* g_full = s_full + 1;          // [4] own G, 64-key quarter q, written to tensor memory
  uint64_t* stg_full = s_full + 5;        // partner's G has landed in my staging buffer
  uint64_t* stg_free = s_full + 6;        // the partner has consumed what I stored in ITS staging
  uint64_t* acc_full = s_full + 7;
  uint64_t* acc_empty = s_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 9);
  float* s_range = reinterpret_cast<float*>(s_full + 10);  // [16] lo per warp, [16] hi per warp
  float* s_mu = s_range + 32;                              // [kMaxBwdProbs] centre of the offsets
  int* s_fast = reinterpret_cast<int*>(s_mu + kMaxBwdProbs);  // [kMaxBwdProbs] one-exponential form
  // hand-off of a G chunk from epilogue warp w to the send warps: written / read out of sXfer[w]
  uint64_t* xf_full = reinterpret_cast<uint64_t*>(s_fast + kMaxBwdProbs);  // [8]
  uint64_t* xf_free = xf_full + 8;                                          // [8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef MMCL_TRACE
  unsigned trace_i = 0;
  const int trace_role = warp < 2 ? warp : 2 + ((warp - 4) >> 2);
#endif
  const uint32_t c = cluster_ctarank();
  const uint32_t peer = c ^ 1u;
  const int pair = blockIdx.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(s_full, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&g_full[i], 8);  // one arrival per epilogue warp
    mbar_init(stg_full, 1);  // armed with expect_tx by the MMA issuer; bytes arrive via st.async
    mbar_init(stg_free, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&xf_full[i], 1);
      mbar_init(&xf_free[i], 1);
    }
    fence_barrier_init();
  }
  if (L.world > 1 && int(threadIdx.x) < L.world && int(threadIdx.x) != L.rank) {
    // the log-sum-exp vectors of the other ranks (column offsets) have landed in this rank's arena
    const unsigned long long* f = L.flags + kChanLse * kMaxRanks + threadIdx.x;
    if (ld_acquire_sys_u64(f) < L.seq) {
      const unsigned long long t0 = global_timer_ns();
      unsigned spins = 0;
      while (ld_acquire_sys_u64(f) < L.seq)
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 4000000000ull) {
          if (L.err) atomicExch(L.err, kErrFlagTimeout);
          break;
        }
    }
  }
  if (L.world > 1) __syncthreads();
  // per problem: range of all offsets -- decides whether the one-exponential form of G is safe
  for (int pi = 0; pi < L.nprob; ++pi) {
    const BwdProb& p = L.prob[pi];
    if (!(p.off_row && p.off_col)) {
      if (threadIdx.x == 0) { s_mu[pi] = 0.f; s_fast[pi] = 0; }
      continue;
    }
    if (p.rng_row && p.rng_col) {
      // the caller supplies the ranges (the scan below costs ~N/384 dependent L2 round trips)
      if (threadIdx.x == 0) {
        float l2 = fmaf(__ldcg(p.rng_row), p.off_mul, p.row_add);
        float h2 = fmaf(__ldcg(p.rng_row + 1), p.off_mul, p.row_add);
        for (int r = 0; r < p.rng_col_count; ++r) {
          const float* q = p.rng_col + (long long)r * p.rng_col_stride;
          l2 = fminf(l2, fmaf(__ldcg(q), p.off_mul, p.col_add));
          h2 = fmaxf(h2, fmaf(__ldcg(q + 1), p.off_mul, p.col_add));
        }
        s_mu[pi] = 0.5f * (l2 + h2);
        s_fast[pi] = (h2 - l2) <= 2.f * kFastRange ? 1 : 0;
      }
      continue;
    }
    // (scan path: callers without precomputed ranges order the vectors before the launch, so
    // plain cached loads are safe here)
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < p.nx; i += NTHREADS) {
      const float v = fmaf(p.off_row[i], p.off_mul, p.row_add);
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
    for (int i = threadIdx.x; i < p.ny; i += NTHREADS) {
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
    __syncthreads();
    if (threadIdx.x == 0) {
      float l2 = INFINITY, h2 = -INFINITY;
      for (int w = 0; w < NTHREADS / 32; ++w) {
        l2 = fminf(l2, s_range[w]);
        h2 = fmaxf(h2, s_range[16 + w]);
      }
      s_mu[pi] = 0.5f * (l2 + h2);
      s_fast[pi] = (h2 - l2) <= 2.f * kFastRange ? 1 : 0;  // false for inf/NaN ranges as well
    }
    __syncthreads();
  }
  __syncthreads();
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the partner's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_acc = tmem, tm_s = tmem + 256;

  // this pair's share [t_begin, t_end) of problem pi
  auto prob_range = [&](int pi, long long* t_begin, long long* t_end) {
    const BwdProb& P = L.prob[pi];
    *t_begin = P.tile_begin + (long long)pair * P.per_pair;
    long long e = *t_begin + P.per_pair;
    if (e > P.tile_begin + P.tiles) e = P.tile_begin + P.tiles;
    *t_end = e;
  };
  const int nb = L.dcc / BK;  // 64-wide column boxes of this CTA's dX slice

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const CUtensorMap *tmX = nullptr, *tmY256 = nullptr, *tmY64 = nullptr;
#ifdef MMCL_TRACE
      unsigned prod_cnt = 0;
#endif
      auto load_s = [&](int xt, int yt) {
        TRACE(56, prod_cnt);
        for (int kc = 0; kc < L.kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (kc == 0 || kc == 3) TRACE(kc == 0 ? 8 : 9, prod_cnt);
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* dst = sStage + stage * STAGE_BYTES;
          tma_load_2d(dst, tmX, kc * BK, xt * BM, &full[stage]);
          tma_load_2d(dst + XCHUNK_BYTES, tmY256, kc * BK, yt * BNY, &full[stage]);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
      };
      auto load_dx = [&](int col0, int yt, int kind) {
#pragma unroll 1
        for (int js = 0; js < BNY / 64; ++js) {
          mbar_wait(&empty[stage], phase ^ 1);
          TRACE(24 + kind * 4 + js, prod_cnt);
          mbar_arrive_expect_tx(&full[stage], nb * YSUB_BYTES);
          uint8_t* dst = sStage + stage * STAGE_BYTES;
          for (int b = 0; b < nb; ++b)
            tma_load_2d(dst + b * YSUB_BYTES, tmY64, col0 + b * BK, yt * BNY + js * 64,
                        &full[stage]);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
          TRACE(52 + js, prod_cnt * 2 + kind);
        }
      };
      bool pend = false;
      int pend_yt = 0;
      for (int pi = 0; pi < L.nprob; ++pi) {
      long long t_begin, t_end;
      prob_range(pi, &t_begin, &t_end);
      const BwdProb& p = L.prob[pi];
      tmX = &maps.tm[p.map_x];
      tmY256 = &maps.tm[p.map_y256];
      tmY64 = &maps.tm[p.map_y64];
      TileWalk w;
      if (t_begin < t_end) w.start(t_begin - p.tile_begin, p.nyt, L.ndc);
      const int n = int(t_end - t_begin), nyt = p.nyt, ndc = L.ndc;
      for (int i = 0; i < n; ++i, w.next(nyt, ndc)) {
        const int yt = w.yt, xt = w.xt, dc = w.dc;
        const int col0 = dc * L.dp + int(c) * L.dcc;
        const bool own = (uint32_t(i) & 1u) == c;
        const bool seg_last = (i == n - 1) || (yt == nyt - 1);
        if (own) {
          load_s(xt, yt);
#ifdef MMCL_TRACE
          ++prod_cnt;
#endif
          if (pend) { load_dx(col0, pend_yt, 0); pend = false; }
          load_dx(col0, yt, 1);
        } else {
          pend = true;
          pend_yt = yt;
        }
        if (seg_last && pend) { load_dx(col0, pend_yt, 0); pend = false; }
      }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(BM, BNY, 0, 0);
      const uint32_t idesc_dx = make_idesc_bf16(BM, uint32_t(L.dcc), 0, 1);
      const uint32_t stg = smem_u32(sStg);
      int stage = 0;
      uint32_t phase = 0, own_cnt = 0, peer_cnt = 0, seg = 0, accf = 0;
      auto issue_s = [&]() {
        TRACE(0, own_cnt);
        for (int kc = 0; kc < L.kchunks; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          TRACE(16 + kc, own_cnt);
          const uint32_t st = smem_u32(sStage + stage * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_ss(tm_s, make_smem_desc_sw128(st + k * 32, 0, 1024),
                    make_smem_desc_sw128(st + XCHUNK_BYTES + k * 32, 0, 1024), idesc_s,
                    (kc | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(s_full);
        TRACE(1, own_cnt);
      };
      auto begin_dx = [&]() {
        if (!accf) {  // the previous segment's accumulators must have been drained
          mbar_wait(acc_empty, (seg & 1) ^ 1);
          tc_fence_after();
        }
      };
      auto issue_dx_own = [&]() {  // A = own G in tensor memory
        begin_dx();
#pragma unroll 1
        for (int js = 0; js < BNY / 64; ++js) {
          mbar_wait(&g_full[js], own_cnt & 1);
          tc_fence_after();
          if ((js & 1) == 0) TRACE(5 + (js >> 1), own_cnt);
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          TRACE(40 + js, own_cnt);
          const uint32_t st = smem_u32(sStage + stage * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_ts(tm_acc, tm_s + (js * 2 + (k >> 1)) * 32 + (k & 1) * 8,
                    make_smem_desc_sw128(st + k * 2048, YSUB_BYTES, 1024), idesc_dx, accf);
            accf = 1;
          }
          umma_commit(&empty[stage]);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
        TRACE(7, own_cnt);
        ++own_cnt;
      };
      auto issue_dx_peer = [&]() {  // A = the partner's G in my staging buffer
        begin_dx();
        TRACE(2, peer_cnt);
#ifdef MMCL_NO_DSM  // timing experiment only (tools/bwd_trace.cu): no G exchange, wrong results
        mbar_arrive_expect_tx(stg_full, 8 * 8 * 16);
#else
        mbar_arrive_expect_tx(stg_full, STG_BYTES);  // 128 rows x 512 B from the partner's st.async
#endif
        mbar_wait_cluster(stg_full, peer_cnt & 1);
        TRACE(3, peer_cnt);
        fence_proxy_async_smem();  // the partner's stores -> my tensor-core (async proxy) reads
        tc_fence_after();
#pragma unroll 1
        for (int js = 0; js < BNY / 64; ++js) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          TRACE(44 + js, peer_cnt);
          const uint32_t st = smem_u32(sStage + stage * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_ss(tm_acc, make_smem_desc_sw128(stg + js * ATOM_BYTES + k * 32, 0, 1024),
                    make_smem_desc_sw128(st + k * 2048, YSUB_BYTES, 1024), idesc_dx, accf);
            accf = 1;
          }
          umma_commit(&empty[stage]);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
        // tell the partner's epilogue that my staging buffer may be overwritten
        umma_commit_mcast(stg_free, uint16_t(1u << peer));
        TRACE(4, peer_cnt);
        ++peer_cnt;
      };
      bool pend = false;
      for (int pi = 0; pi < L.nprob; ++pi) {
      long long t_begin, t_end;
      prob_range(pi, &t_begin, &t_end);
      const int nyt = L.prob[pi].nyt;
      TileWalk w;
      if (t_begin < t_end) w.start(t_begin - L.prob[pi].tile_begin, nyt, L.ndc);
      const int n = int(t_end - t_begin), ndc = L.ndc;
      for (int i = 0; i < n; ++i, w.next(nyt, ndc)) {
        const bool own = (uint32_t(i) & 1u) == c;
        const bool seg_last = (i == n - 1) || (w.yt == nyt - 1);
        if (own) {
          issue_s();
          if (pend) { issue_dx_peer(); pend = false; }
          issue_dx_own();
        } else {
          pend = true;
        }
        if (seg_last) {
          if (pend) { issue_dx_peer(); pend = false; }
          umma_commit(acc_full);
          accf = 0;
          ++seg;
        }
      }
      }
    }
  } else if (warp < 4) {
    // ================================ G send warps ================================
    // Warp 2 serves epilogue warps 4-7, warp 3 serves 8-11: per chunk it waits for the warp's 2 KB
    // buffer, reads it back transposed (the buffer is free again as soon as the values are in
    // registers) and stores it into the partner's staging tile (128-byte swizzle, K-major): each
    // st.async covers 8 rows x 64 contiguous bytes -- distributed-shared-memory stores are
    // packet-rate bound.
    const int wg = warp - 2;
    const uint32_t peer_base = mapa_shared(smem_u32(sStg), peer);
    const uint32_t peer_stg_full = mapa_shared(smem_u32(stg_full), peer);
    uint32_t own_cnt = 0, xf_cnt = 0;
    for (int pi = 0; pi < L.nprob; ++pi) {
    long long t_begin, t_end;
    prob_range(pi, &t_begin, &t_end);
    const int n = int(t_end - t_begin);
    for (int i = 0; i < n; ++i) {
      if ((uint32_t(i) & 1u) != c) continue;
      // the partner has finished reading the G tile stored two steps ago (write-after-read on its
      // staging buffer; nothing is read here, so CTA-scope acquire is enough)
      mbar_wait(stg_free, (own_cnt & 1) ^ 1);
      ++own_cnt;
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
        const int ch = ci * 2 + wg;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          const int ew = wg * 4 + q;  // epilogue warp 4 + ew: lane quadrant q, key half wg
          const uint32_t xb = smem_u32(sXfer) + uint32_t(ew) * 2048u;
          mbar_wait(&xf_full[ew], xf_cnt & 1);
          uint32_t w[4][4];
#pragma unroll
          for (int s2 = 0; s2 < 4; ++s2)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w[s2][0]), "=r"(w[s2][1]), "=r"(w[s2][2]), "=r"(w[s2][3])
                         : "r"(xb + uint32_t(s2 * 512 + lane * 16))
                         : "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&xf_free[ew]);
          const uint32_t abase =
              peer_base + uint32_t(q * 32) * 128u + uint32_t(ch >> 1) * ATOM_BYTES;
#pragma unroll
          for (int s2 = 0; s2 < 4; ++s2) {
            const uint32_t r = uint32_t(s2 * 8 + (lane >> 2));  // row within the warp's 32 rows
            const uint32_t half = (uint32_t(ch) ^ (r >> 2)) & 1u;
#ifdef MMCL_NO_DSM
            if (s2 == 0 && ci == 0 && lane < 8)  // 8 warps x 8 lanes x 16 B: timing experiment only
#endif
            st_async_v4(abase + r * 128u + half * 64u + (uint32_t(lane & 3) << 4), w[s2][0],
                        w[s2][1], w[s2][2], w[s2][3], peer_stg_full);
          }
        }
        ++xf_cnt;
      }
    }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int wg = (warp - 4) >> 2;  // key half this warpgroup converts
    const int q = warp & 3;          // TMEM lane quadrant (== warp % 4)
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128 - wg * 128;  // 0..127 within the warpgroup
    const uint32_t lane_base = uint32_t(q * 32) << 16;
    const uint32_t xb = smem_u32(sXfer) + uint32_t(warp - 4) * 2048u;
    const uint32_t sw = uint32_t(row & 7);
    uint32_t own_cnt = 0, seg = 0, xf_cnt = 0;
    bool seg_head = false;  // the current segment continues an item another CTA pair started
    for (int pi = 0; pi < L.nprob; ++pi) {
    long long t_begin, t_end;
    prob_range(pi, &t_begin, &t_end);
    const BwdProb& p = L.prob[pi];
    TileWalk w;
    if (t_begin < t_end) w.start(t_begin - p.tile_begin, p.nyt, L.ndc);
    const int n = int(t_end - t_begin);
    for (int i = 0; i < n; ++i, w.next(p.nyt, L.ndc)) {
      const int yt = w.yt, xt = w.xt, dc = w.dc;
      const bool own = (uint32_t(i) & 1u) == c;
      const bool seg_last = (i == n - 1) || (yt == p.nyt - 1);
      const int grow = xt * BM + row;
      if (i == 0) seg_head = yt != 0;
      const int mode = p.bce ? 4 : ((p.off_row ? 1 : 0) | (p.off_col ? 2 : 0));
      const bool fast = s_fast[pi] != 0;
      const float mu = s_mu[pi];
      if (own) {
        const float c0 = (*p.scale) * LOG2E;
        const float offr = (p.off_row && grow < p.nx)
                               ? fmaf(__ldcg(p.off_row + grow), p.off_mul, p.row_add) : INFINITY;
        const float cc = c0 * ((p.xinv && grow < p.nx) ? p.xinv[grow] : 1.f);
        const float R = (fast && grow < p.nx) ? ex2f(offr - mu) : 0.f;
        // BCE form: this row's band and coefficients
        const float tt = float(p.ny);
        const long long ig = p.row_base + grow;
        long long bh_i = ig + 1;
        if (mode == 4 && grow < p.nx) {
          const long long b0 = (long long)p.band_hi[ig];
          if (b0 > bh_i) bh_i = b0;
        }
        const float crow = (*p.scale) * ((p.xinv && grow < p.nx) ? p.xinv[grow] : 1.f);
        const float p_i = __frcp_rn(tt * float(bh_i - ig));
        const float n_i = __frcp_rn(tt * (tt - float(bh_i - ig)));
        named_bar_sync(1, 256);  // everyone is done with the previous step's column data
        {
          const int j = wg * 128 + et;
          const int gy = yt * BNY + j;
          float2 v2;
          v2.x = (p.off_col && gy < p.ny) ? fmaf(__ldcg(p.off_col + gy), p.off_mul, p.col_add)
                                          : INFINITY;
          v2.y = (p.yinv && gy < p.ny) ? p.yinv[gy] : 1.f;
          if (mode == 4) {
            int bh = gy + 1;
            if (gy < p.ny) {
              const int b0 = p.band_hi[gy];
              if (b0 > bh) bh = b0;
            }
            v2.x = __int_as_float(bh);
          }
          if (fast) {
            // pair-planar layout for the packed-math path: (cyB0, cyB1, yinv0, yinv1)
            float* f = reinterpret_cast<float*>(sCol) + (j >> 1) * 4 + (j & 1);
            f[0] = ex2f(mu - v2.x) * v2.y;
            f[2] = v2.y;
          } else {
            sCol[j] = v2;
          }
        }
        named_bar_sync(2, 256);
        mbar_wait(s_full, own_cnt & 1);
        tc_fence_after();
        if (lane == 0 && q == 0) TRACE(10 + wg, own_cnt);
        if (lane == 0 && q == 0) TRACE(12 + wg, own_cnt);
        // Conversion order: step ci converts the 64-key quarter ci (32-column chunk 2*ci by
        // warpgroup 0, chunk 2*ci+1 by warpgroup 1) and hands it to the MMA issuer at once, so the
        // dX MMAs of quarter ci run under the conversion of the quarters after it.
        // Software pipeline over 16-column halves: the tensor-memory load of the next half is
        // in flight while the current one is converted.
        uint32_t va[16], vb[16], g[16];
        const uint32_t sbase = tm_s + lane_base;
        auto convert = [&](const uint32_t (&v)[16], uint32_t* gout, int col) {
          if (mode == 4)
            conv16_bce(v, gout, crow, ig, bh_i, p_i, n_i, tt, (long long)yt * BNY + col, p.ny,
                       sCol + col);
          else if (fast)
            conv16_fast(v, gout, cc, offr, R, reinterpret_cast<const float4*>(sCol) + (col >> 1));
          else if (mode == 3)
            conv16<3>(v, gout, cc, offr, sCol + col);
          else if (mode == 1)
            conv16<1>(v, gout, cc, offr, sCol + col);
          else
            conv16<2>(v, gout, cc, offr, sCol + col);
        };
        tmem_ld_x16(sbase + uint32_t(wg) * 32u, va);
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          const int ch = ci * 2 + wg;
          tmem_ld_wait();
          TRACE_EPI(0);
          tmem_ld_x16(sbase + ch * 32 + 16, vb);
          convert(va, g, ch * 32);
          TRACE_EPI(1);
          tmem_ld_wait();
          if (ci < 3) tmem_ld_x16(sbase + ((ci + 1) * 2 + wg) * 32, va);
          convert(vb, g + 8, ch * 32 + 16);
          TRACE_EPI(2);
          // G chunk ch overwrites the first 16 columns of S chunk ch (already consumed by this
          // very thread), so no other warp's reads are ever at risk
          tmem_st_x16(tm_s + lane_base + ch * 32, g);
          // this warpgroup's share of quarter ci is in tensor memory: hand it to the MMA issuer
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();  // every lane's store is fenced before the one arrival of this warp
          if (lane == 0) mbar_arrive(&g_full[ci]);
          TRACE_EPI(3);
          // ... and the same 32 keys (64 B per row) to the send warps (warps 2-3) through this warp's
          // 2 KB buffer: they read it back transposed and issue the st.async stream, so its
          // back-pressure never stalls the conversion
          const uint32_t j0 = uint32_t(ch & 1) * 4u;
          mbar_wait(&xf_free[warp - 4], (xf_cnt & 1) ^ 1);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const uint32_t a = xb + uint32_t(lane) * 64u + ((((j0 + qd) ^ sw) & 3u) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(g[4 * qd]),
                         "r"(g[4 * qd + 1]), "r"(g[4 * qd + 2]), "r"(g[4 * qd + 3])
                         : "memory");
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&xf_full[warp - 4]);
          ++xf_cnt;
          TRACE_EPI(4);
        }
        if (lane == 0 && q == 0) TRACE(14 + wg, own_cnt);
        ++own_cnt;
      }
      if (seg_last) {
        // flush this segment's accumulators
        mbar_wait(acc_full, seg & 1);
        tc_fence_after();
        const int col0 = dc * L.dp + int(c) * L.dcc;
#pragma unroll 1
        for (int ch = wg; ch < L.dcc / 32; ch += 2) {
          uint32_t v[32];
          tmem_ld_x32(tm_acc + lane_base + ch * 32, v);
          tmem_ld_wait();
          if (L.accumulate) {
            if (grow < p.nx) {
              float* dst = p.dx + (long long)grow * L.d + col0 + ch * 32;
#pragma unroll
              for (int e = 0; e < 32; e += 4)
                red_add_v4(dst + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
            }
          } else if (seg_head) {
            // continuation of an item another pair started: this pair's scratch tile
            float4* dst = reinterpret_cast<float4*>(
                L.scratch + (((long long)pi * L.pairs + pair) * BM + row) * L.dp + int(c) * L.dcc +
                ch * 32);
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              dst[e >> 2] = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                        __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
          } else if (grow < p.nx) {
            float4* dst =
                reinterpret_cast<float4*>(p.dx + (long long)grow * L.d + col0 + ch * 32);
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              dst[e >> 2] = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                        __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        ++seg;
        seg_head = false;
      }
    }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves while the partner may still store into / signal this CTA
  if (warp == 2) tmem_dealloc<512>(tmem);
}

}  // namespace

// ---- planning (host) ---------------------------------------------------------------------------

bool bwd_pair_supported(int num_sms, long long d) {
  const int ndc = int((d + 511) / 512);
  return num_sms >= 2 && d % (128 * ndc) == 0;
}

void bwd_begin(BwdLaunch* L, long long d) {
  *L = BwdLaunch();
  L->d = int(d);
  L->kchunks = int(d / BK);
  L->ndc = int((d + 511) / 512);
  L->dp = int(d / L->ndc);
  L->dcc = L->dp / 2;
  L->world = 1;
  L->accumulate = 1;
}

int bwd_add_prob(BwdLaunch* L, long long nx, long long ny) {
  if (L->nprob >= kMaxBwdProbs) return -1;
  BwdProb& p = L->prob[L->nprob];
  p = BwdProb();
  p.nx = int(nx);
  p.ny = int(ny);
  p.nxt = int((nx + BM - 1) / BM);
  p.nyt = int((ny + BNY - 1) / BNY);
  p.tile_begin = L->total;
  p.off_mul = 1.f;
  p.tiles = (long long)p.nxt * L->ndc * p.nyt;
  L->total += p.tiles;
  return L->nprob++;
}

// Fixes the split over CTA pairs; returns the scratch bytes an accumulate == 0 launch needs.
size_t bwd_finish(BwdLaunch* L, int num_sms) {
  long long widest = 1;
  for (int i = 0; i < L->nprob; ++i)
    if (L->prob[i].tiles > widest) widest = L->prob[i].tiles;
  long long pairs = num_sms / 2;
  if (widest < pairs) pairs = widest;
  if (pairs < 1) pairs = 1;
  L->pairs = int(pairs);
  for (int i = 0; i < L->nprob; ++i) {
    BwdProb& p = L->prob[i];
    p.per_pair = int((p.tiles + pairs - 1) / pairs);
    if (p.per_pair < 1) p.per_pair = 1;
  }
  return size_t(L->nprob) * size_t(pairs) * BM * size_t(L->dp) * sizeof(float);
}

int bwd_cta_pairs(const BwdLaunch& L) { return L.pairs; }

int bwd_make_maps(TensorMaps* maps, int slot, const void* rows, long long n, long long d) {
  if (slot < 0 || 3 * slot + 2 >= kMaxBwdMaps) return -1;
  if (make_tmap_bf16_sw128(&maps->tm[3 * slot], rows, n, d, d, BM)) return -1;
  if (make_tmap_bf16_sw128(&maps->tm[3 * slot + 1], rows, n, d, d, BNY)) return -1;
  if (make_tmap_bf16_sw128(&maps->tm[3 * slot + 2], rows, n, d, d, 64)) return -1;
  return 0;
}

cudaError_t launch_bwd_fused(const TensorMaps& maps, const BwdLaunch& L, cudaStream_t stream) {
  if (L.total == 0) return cudaSuccess;
  const size_t smem =
      size_t(NSTAGES) * STAGE_BYTES + STG_BYTES + BNY * sizeof(float2) + XFER_BYTES + TAIL_BYTES;
  cudaError_t e = cudaFuncSetAttribute(sim_grad_bwd_pair_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  sim_grad_bwd_pair_kernel<<<2 * bwd_cta_pairs(L), NTHREADS, smem, stream>>>(maps, L);
  count_launches(1);
  return cudaGetLastError();
}

// BCE form of the recompute backward (one problem, accumulating into a zero-filled dx).
cudaError_t launch_sim_bce_bwd_pair(int num_sms, const void* x, long long nx, const void* y,
                                    long long ny, long long d, const float* scale,
                                    const float* xinv, const float* yinv, long long row_base,
                                    const int* band_hi, float* dx, cudaStream_t stream,
                                    const char** err) {
  if (!bwd_pair_supported(num_sms, d)) return cudaErrorNotSupported;
  BwdLaunch L;
  bwd_begin(&L, d);
  bwd_add_prob(&L, nx, ny);
  bwd_finish(&L, num_sms);
  BwdProb& p = L.prob[0];
  p.scale = scale;
  p.xinv = xinv;
  p.yinv = yinv;
  p.bce = 1;
  p.row_base = row_base;
  p.band_hi = band_hi;
  p.dx = dx;
  TensorMaps maps;
  if (make_tmap_bf16_sw128(&maps.tm[0], x, nx, d, d, BM) ||
      make_tmap_bf16_sw128(&maps.tm[1], y, ny, d, d, BNY) ||
      make_tmap_bf16_sw128(&maps.tm[2], y, ny, d, d, 64)) {
    *err = "sim_bce_bwd: cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  p.map_x = 0;
  p.map_y256 = 1;
  p.map_y64 = 2;
  return launch_bwd_fused(maps, L, stream);
}

// Returns cudaErrorNotSupported (without launching) when the shape does not fit the pair form;
// the caller then uses the single-CTA kernel of sim_bwd.cu.
cudaError_t launch_sim_grad_bwd_pair(int num_sms, const void* x, long long nx, const void* y,
                                     long long ny, long long d, const float* scale,
                                     const float* xinv, const float* yinv, const float* off_row,
                                     const float* off_col, float off_mul, float row_add,
                                     float col_add, float* dx, cudaStream_t stream,
                                     const char** err) {
  if (!off_row && !off_col) return cudaSuccess;  // G == 0
  if (!bwd_pair_supported(num_sms, d)) return cudaErrorNotSupported;
  BwdLaunch L;
  bwd_begin(&L, d);
  bwd_add_prob(&L, nx, ny);
  bwd_finish(&L, num_sms);
  BwdProb& p = L.prob[0];
  p.scale = scale;
  p.xinv = xinv;
  p.yinv = yinv;
  p.off_row = off_row;
  p.off_col = off_col;
  p.off_mul = off_mul;
  p.row_add = row_add;
  p.col_add = col_add;
  p.dx = dx;
  TensorMaps maps;
  if (make_tmap_bf16_sw128(&maps.tm[0], x, nx, d, d, BM) ||
      make_tmap_bf16_sw128(&maps.tm[1], y, ny, d, d, BNY) ||
      make_tmap_bf16_sw128(&maps.tm[2], y, ny, d, d, 64)) {
    *err = "sim_grad_bwd_pair: cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  p.map_x = 0;
  p.map_y256 = 1;
  p.map_y64 = 2;
  return launch_bwd_fused(maps, L, stream);
}

}  // namespace mmcl

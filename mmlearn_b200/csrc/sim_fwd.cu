// Forward hot kernel: S = s * Q K^T in 128x256 fp32 tiles that live only in tensor memory,
// reduced on the fly to per-row (max, sum-exp, sum-exp*logit).
//
// Roles (384 threads, 1 CTA / SM, persistent):
//   warp 0    TMA producer  (Q tile resident in smem when it fits, K streamed in 64-wide k-chunks)
//   warp 1    tcgen05.mma issuer (one elected thread), double-buffered TMEM accumulators
//   warp 2    TMEM allocator
//   warps 4-11 epilogue: tcgen05.ld -> online log-sum-exp in registers.  Two warps per SM
//             sub-partition: warps 4-7 reduce columns 0-127 of each 128x256 tile, warps 8-11
//             columns 128-255 (one row per thread, each half gives its own partial statistics).
//             With one warp per sub-partition the epilogue (dependent ALU latency, MUFU) took
//             ~6500 cycles per tile against ~4200 for the MMAs (ncu stall samples, round 1).
// Work distribution: the nqt x nkt tile grid is linearised (q-tile major) and cut into
// equal contiguous ranges, one per CTA ("stream-K" over the key axis), so all 148 SMs finish
// together whatever nq is; a row's partial statistics from different CTAs are merged by
// lse_combine_kernel.
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace mmcl {

namespace {

constexpr int BM = 128;      // query rows per tile (UMMA M)
constexpr int BN = 256;      // key rows per tile (UMMA N); N=256 is the only shape that amortises
                             // the ~128-cycle per-instruction floor measured in tools/umma_rate*.cu
constexpr int BK = 64;       // k-chunk: 64 bf16 = one 128-byte swizzle atom
constexpr int QCHUNK_BYTES = BM * BK * 2;  // 16 KB
constexpr int KCHUNK_BYTES = BN * BK * 2;  // 32 KB
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct FwdParams {
  int nq, nk, kchunks;
  int nqt, nkt;
  long long total_tiles;
  int tiles_per_cta;
  int nstages;
  int max_slots;
  long long part_stride;  // floats between slots (= nqt * BM)
  const float* scale;
  const float* qinv;  // [nq] 1/||q_i|| or null (rows already unit-norm / no normalisation)
  const float* kinv;  // [nk] likewise for the keys
  float* part_m;
  float* part_l;
  float* part_t;
  const float* thr;  // rank mode: per-row threshold on  qinv_i * kinv_j * (q_i . k_j)
  const long long* tgt;  // rank mode: per-row key index that is never counted (the positive) or null
};

// kRank == false: online log-sum-exp epilogue (the loss).  kRank == true: the same mainloop with a
// counting epilogue -- per row, the number of keys whose similarity exceeds the row's threshold
// (retrieval rank of the positive, mmcl_sim_rank_fwd); one multiply + compare per logit, no MUFU.
template <bool kQRes, bool kRank>
__global__ void __launch_bounds__(384, 1)
sim_lse_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operands need 1024-byte alignment
  const int stage_bytes = kQRes ? KCHUNK_BYTES : QCHUNK_BYTES + KCHUNK_BYTES;
  uint8_t* sQ = smem;
  uint8_t* sStage = smem + (kQRes ? p.kchunks * QCHUNK_BYTES : 0);
  float* sKinv = reinterpret_cast<float*>(sStage + p.nstages * stage_bytes);  // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKinv + 2 * BN);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.nstages;
  uint64_t* q_full = bars + 2 * p.nstages;
  uint64_t* q_empty = q_full + 1;
  uint64_t* acc_full = q_full + 2;   // [2]
  uint64_t* acc_empty = q_full + 4;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const long long t_begin = (long long)blockIdx.x * p.tiles_per_cta;
  long long t_end = t_begin + p.tiles_per_cta;
  if (t_end > p.total_tiles) t_end = p.total_tiles;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, qphase = 0;
      for (long long t = t_begin; t < t_end;) {
        const int qt = int(t / p.nkt), kt0 = int(t % p.nkt);
        int kt1 = kt0 + int(t_end - t);
        if (kt1 > p.nkt) kt1 = p.nkt;
        if (kQRes) {
          mbar_wait(q_empty, qphase ^ 1);
          mbar_arrive_expect_tx(q_full, p.kchunks * QCHUNK_BYTES);
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_2d(sQ + kc * QCHUNK_BYTES, &tmQ, kc * BK, qt * BM, q_full);
          qphase ^= 1;
        }
        for (int kt = kt0; kt < kt1; ++kt) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], stage_bytes);
            uint8_t* dst = sStage + stage * stage_bytes;
            if (!kQRes) {
              tma_load_2d(dst, &tmQ, kc * BK, qt * BM, &full[stage]);
              dst += QCHUNK_BYTES;
            }
            tma_load_2d(dst, &tmK, kc * BK, kt * BN, &full[stage]);
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          }
        }
        t += kt1 - kt0;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0, qphase = 0;
      uint32_t acc = 0;
      for (long long t = t_begin; t < t_end;) {
        const int kt0 = int(t % p.nkt);
        int kt1 = kt0 + int(t_end - t);
        if (kt1 > p.nkt) kt1 = p.nkt;
        if (kQRes) {
          mbar_wait(q_full, qphase);
          qphase ^= 1;
        }
        for (int kt = kt0; kt < kt1; ++kt) {
          const uint32_t ab = acc & 1, aphase = (acc >> 1) & 1;
          mbar_wait(&acc_empty[ab], aphase ^ 1);
          tc_fence_after();
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(sStage + stage * stage_bytes);
            const uint32_t qa = kQRes ? smem_u32(sQ + kc * QCHUNK_BYTES) : st;
            const uint32_t ka = kQRes ? st : st + QCHUNK_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(qa + k * 32, 0, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(ka + k * 32, 0, 1024);
              umma_ss(tmem + ab * BN, adesc, bdesc, idesc, (kc | k) != 0);
            }
            umma_commit(&empty[stage]);
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&acc_full[ab]);
          ++acc;
        }
        if (kQRes) umma_commit(q_empty);
        t += kt1 - kt0;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int wg = (warp - 4) >> 2;  // column half of every tile this warpgroup reduces
    const int q = warp & 3;          // == warp % 4: TMEM lane quadrant
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;  // 0..255: one key factor per thread and tile
    const int ch0 = wg * (BN / 64);    // first of this warpgroup's BN/64 32-column chunks
    const float s = kRank ? 1.f : *p.scale;
    const bool neg = s < 0.f;
    uint32_t acc = 0;
    for (long long t = t_begin; t < t_end;) {
      const int qt = int(t / p.nkt), kt0 = int(t % p.nkt);
      int kt1 = kt0 + int(t_end - t);
      if (kt1 > p.nkt) kt1 = p.nkt;
      float m = -INFINITY, l = 0.f, tt = 0.f;
      const int grow = qt * BM + row;
      float knext;  // key factor of column `et` of the next tile (prefetched one tile ahead)
      {
        const int gk = kt0 * BN + et;
        knext = (p.kinv && gk < p.nk) ? p.kinv[gk] : 1.f;
      }
      if (kRank) {
        // count_j [ qinv_i * kinv_j * (q_i . k_j) > thr_i ]  ==  [ x_ij * kinv_j > thr_i / qinv_i ]
        const float qi = (p.qinv && grow < p.nq) ? p.qinv[grow] : 1.f;
        const float th = (grow < p.nq ? p.thr[grow] : INFINITY) / qi;
        const long long tg = (p.tgt && grow < p.nq) ? p.tgt[grow] : -1;
        int cnt = 0;
        for (int kt = kt0; kt < kt1; ++kt) {
          const uint32_t ab = acc & 1, aphase = (acc >> 1) & 1;
          float* kv = sKinv + ab * BN;
          kv[et] = knext;  // fetched while the previous tile was being reduced
          if (kt + 1 < kt1) {
            const int gk = (kt + 1) * BN + et;
            knext = (p.kinv && gk < p.nk) ? p.kinv[gk] : 1.f;
          }
          named_bar_sync(1, 256);
          mbar_wait(&acc_full[ab], aphase);
          tc_fence_after();
          const uint32_t taddr = tmem + ab * BN + (uint32_t(q * 32) << 16);
          const int valid = p.nk - kt * BN;
#pragma unroll 1
          for (int ch = ch0; ch < ch0 + BN / 64; ++ch) {
            uint32_t v[32];
            tmem_ld_x32(taddr + ch * 32, v);
            tmem_ld_wait();
            if (ch == ch0 + BN / 64 - 1) {
              tc_fence_before();
              mbar_arrive(&acc_empty[ab]);
            }
            const int cv = valid - ch * 32;
            const float* kc = kv + ch * 32;
            const long long tl = tg - ((long long)kt * BN + ch * 32);  // the positive's slot, if here
            if (cv >= 32 && (tl < 0 || tl >= 32)) {
              // common case: whole chunk, positive elsewhere -> FMUL + FSETP + predicated FADD
              float c0 = 0.f, c1 = 0.f;
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                if (__uint_as_float(v[e]) * kc[e] > th) c0 += 1.f;
                if (__uint_as_float(v[e + 1]) * kc[e + 1] > th) c1 += 1.f;
              }
              cnt += int(c0 + c1);
            } else {
              const int tloc = (tl >= 0 && tl < 32) ? int(tl) : -1;
#pragma unroll
              for (int e = 0; e < 32; ++e)
                cnt += (e < cv && e != tloc && __uint_as_float(v[e]) * kc[e] > th) ? 1 : 0;
            }
          }
          ++acc;
        }
        m = float(cnt);  // exact below 2^24 keys per segment
      } else {
      // logit = s * (q_i.k_j) * qinv_i * kinv_j ; c folds everything that depends on the row
      const float c = s * LOG2E * ((p.qinv && grow < p.nq) ? p.qinv[grow] : 1.f);
      for (int kt = kt0; kt < kt1; ++kt) {
        const uint32_t ab = acc & 1, aphase = (acc >> 1) & 1;
        float* kv = sKinv + ab * BN;
        kv[et] = knext;  // fetched while the previous tile was being reduced
        if (kt + 1 < kt1) {
          const int gk = (kt + 1) * BN + et;
          knext = (p.kinv && gk < p.nk) ? p.kinv[gk] : 1.f;
        }
        named_bar_sync(1, 256);
        mbar_wait(&acc_full[ab], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem + ab * BN + (uint32_t(q * 32) << 16);
        const int valid = p.nk - kt * BN;  // columns of this tile that exist (>= 1)
        // One 32-column chunk: multiply by the key factors, update the running maximum, add the
        // exponentials.  Packed fp32x2 arithmetic (FMUL2 / FFMA2 / FADD2) with two interleaved
        // partial sums halves the FP32 instruction count next to the MUFU work.
        auto process = [&](uint32_t (&v)[32], int ch) {
          const int cv = valid - ch * 32;
          float2 x2[16];
          {
            const float4* k4 = reinterpret_cast<const float4*>(kv + ch * 32);
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const float4 f = k4[e4];
              x2[2 * e4] = __fmul2_rn(make_float2(__uint_as_float(v[4 * e4]),
                                                  __uint_as_float(v[4 * e4 + 1])),
                                      make_float2(f.x, f.y));
              x2[2 * e4 + 1] = __fmul2_rn(make_float2(__uint_as_float(v[4 * e4 + 2]),
                                                      __uint_as_float(v[4 * e4 + 3])),
                                          make_float2(f.z, f.w));
            }
          }
          if (cv >= 32) {
            float ext = x2[0].x;
            if (!neg) {
#pragma unroll
              for (int e = 0; e < 16; ++e) ext = fmaxf(ext, fmaxf(x2[e].x, x2[e].y));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) ext = fminf(ext, fminf(x2[e].x, x2[e].y));
            }
            const float mn = fmaxf(m, ext * c);
            const float alpha = ex2f(m - mn);
            m = mn;
            const float2 c2 = make_float2(c, c), nm2 = make_float2(-m, -m);
            float2 l2 = make_float2(l * alpha, 0.f), t2 = make_float2(tt * alpha, 0.f);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float2 a = __ffma2_rn(x2[e], c2, nm2);
              const float2 pe = make_float2(ex2f(a.x), ex2f(a.y));
              l2 = __fadd2_rn(l2, pe);
              t2 = __ffma2_rn(pe, x2[e], t2);
            }
            l = l2.x + l2.y;
            tt = t2.x + t2.y;
          } else if (cv > 0) {
            // ragged last tile: only the first cv columns exist
            float ext = neg ? INFINITY : -INFINITY;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float x = (e & 1) ? x2[e >> 1].y : x2[e >> 1].x;
              if (e < cv) ext = neg ? fminf(ext, x) : fmaxf(ext, x);
            }
            const float mn = fmaxf(m, ext * c);
            const float alpha = ex2f(m - mn);
            l *= alpha;
            tt *= alpha;
            m = mn;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float x = (e & 1) ? x2[e >> 1].y : x2[e >> 1].x;
              if (e < cv) {
                const float pe = ex2f(fmaf(x, c, -m));
                l += pe;
                tt = fmaf(pe, x, tt);
              }
            }
          }
        };
        // Software pipeline over chunk pairs: the tensor-memory load of the next chunk is in
        // flight while the current one is reduced.
        uint32_t va[32], vb[32];
        tmem_ld_x32(taddr + ch0 * 32, va);
#pragma unroll 1
        for (int ch = ch0; ch < ch0 + BN / 64; ch += 2) {
          tmem_ld_wait();
          tmem_ld_x32(taddr + (ch + 1) * 32, vb);
          process(va, ch);
          tmem_ld_wait();
          if (ch + 2 < ch0 + BN / 64) {
            tmem_ld_x32(taddr + (ch + 2) * 32, va);
          } else {  // every column of this accumulator has been read
            tc_fence_before();
            mbar_arrive(&acc_empty[ab]);
          }
          process(vb, ch + 1);
        }
        ++acc;
      }
      }
      // partial statistics of this (q-tile, key range) segment
      const long long first_tile = (long long)qt * p.nkt;
      const int c_first = int(first_tile / p.tiles_per_cta);
      const int slot = 2 * (int(blockIdx.x) - c_first) + wg;  // one partial per column half
      const long long o = (long long)slot * p.part_stride + (long long)qt * BM + row;
      p.part_m[o] = m;
      p.part_l[o] = l;
      p.part_t[o] = tt;
      t += kt1 - kt0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// Merge the per-CTA partial statistics of each row.
__global__ void lse_combine_kernel(FwdParams p, float* __restrict__ lse, float* __restrict__ esum) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= p.nq) return;
  const int qt = row / BM;
  const long long first_tile = (long long)qt * p.nkt;
  const int c_first = int(first_tile / p.tiles_per_cta);
  const int c_last = int((first_tile + p.nkt - 1) / p.tiles_per_cta);
  const int nslots = 2 * (c_last - c_first + 1);  // two column halves per CTA segment
  float M = -INFINITY;
  for (int s = 0; s < nslots; ++s)
    M = fmaxf(M, p.part_m[(long long)s * p.part_stride + row]);
  float L = 0.f, T = 0.f;
  for (int s = 0; s < nslots; ++s) {
    const long long o = (long long)s * p.part_stride + row;
    const float w = exp2f(p.part_m[o] - M);
    L += p.part_l[o] * w;
    T += p.part_t[o] * w;
  }
  const float sc = (*p.scale) * (p.qinv ? p.qinv[row] : 1.f);
  lse[row] = (M + log2f(L)) * LN2;
  if (esum) esum[row] = sc * T / L;
}

// Rank mode: total count of each row over the key segments of different CTAs.
__global__ void rank_combine_kernel(FwdParams p, long long* __restrict__ rank) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= p.nq) return;
  const int qt = row / BM;
  const long long first_tile = (long long)qt * p.nkt;
  const int c_first = int(first_tile / p.tiles_per_cta);
  const int c_last = int((first_tile + p.nkt - 1) / p.tiles_per_cta);
  long long total = 0;
  for (int s = 0; s < 2 * (c_last - c_first + 1); ++s)
    total += (long long)p.part_m[(long long)s * p.part_stride + row];
  rank[row] = total;
}

int fwd_plan(int num_sms, long long nq, long long nk, long long d, FwdParams* p, bool* qres,
             size_t* smem_bytes) {
  p->nq = int(nq);
  p->nk = int(nk);
  p->kchunks = int((d + BK - 1) / BK);
  p->nqt = int((nq + BM - 1) / BM);
  p->nkt = int((nk + BN - 1) / BN);
  p->total_tiles = (long long)p->nqt * p->nkt;
  long long ctas = p->total_tiles < num_sms ? p->total_tiles : num_sms;
  p->tiles_per_cta = int((p->total_tiles + ctas - 1) / ctas);
  p->max_slots = 2 * ((p->nkt + p->tiles_per_cta - 1) / p->tiles_per_cta + 1);  // x2: column halves
  p->part_stride = (long long)p->nqt * BM;
  const size_t budget = 227 * 1024 - 256 /*barriers*/ - 2 * BN * sizeof(float) /*kinv*/;
  *qres = size_t(p->kchunks) * QCHUNK_BYTES + 3 * size_t(KCHUNK_BYTES) <= budget;
  const size_t stage_bytes = *qres ? KCHUNK_BYTES : QCHUNK_BYTES + KCHUNK_BYTES;
  const size_t fixed = *qres ? size_t(p->kchunks) * QCHUNK_BYTES : 0;
  int ns = int((budget - fixed) / stage_bytes);
  if (ns > 8) ns = 8;
  p->nstages = ns;
  *smem_bytes = fixed + ns * stage_bytes + 2 * BN * sizeof(float) + 256;
  return 0;
}

}  // namespace

size_t sim_lse_workspace_bytes(int num_sms, long long nq, long long nk, long long d) {
  FwdParams p;
  bool qres;
  size_t smem;
  fwd_plan(num_sms, nq, nk, d, &p, &qres, &smem);
  return size_t(3) * p.max_slots * p.part_stride * sizeof(float);
}

namespace {

template <bool kRank>
cudaError_t launch_fwd(int num_sms, const void* q, long long nq, const void* k, long long nk,
                       long long d, FwdParams p, bool qres, size_t smem, cudaStream_t stream,
                       const char** err) {
  CUtensorMap tmQ, tmK;
  if (make_tmap_bf16_sw128(&tmQ, q, nq, d, d, BM) || make_tmap_bf16_sw128(&tmK, k, nk, d, d, BN)) {
    *err = "sim_fwd: cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  const int ctas = int((p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta);
  cudaError_t e;
  if (qres) {
    e = cudaFuncSetAttribute(sim_lse_fwd_kernel<true, kRank>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    sim_lse_fwd_kernel<true, kRank><<<ctas, 384, smem, stream>>>(tmQ, tmK, p);
  } else {
    e = cudaFuncSetAttribute(sim_lse_fwd_kernel<false, kRank>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    sim_lse_fwd_kernel<false, kRank><<<ctas, 384, smem, stream>>>(tmQ, tmK, p);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_sim_lse_fwd(int num_sms, const void* q, long long nq, const void* k,
                               long long nk, long long d, const float* scale, const float* qinv,
                               const float* kinv, float* lse, float* esum, void* workspace,
                               size_t workspace_bytes, cudaStream_t stream, const char** err) {
  FwdParams p;
  bool qres;
  size_t smem;
  fwd_plan(num_sms, nq, nk, d, &p, &qres, &smem);
  const size_t need = size_t(3) * p.max_slots * p.part_stride * sizeof(float);
  if (workspace_bytes < need) {
    *err = "sim_lse_fwd: workspace too small";
    return cudaErrorInvalidValue;
  }
  p.scale = scale;
  p.qinv = qinv;
  p.kinv = kinv;
  p.thr = nullptr;
  p.tgt = nullptr;
  p.part_m = static_cast<float*>(workspace);
  p.part_l = p.part_m + size_t(p.max_slots) * p.part_stride;
  p.part_t = p.part_l + size_t(p.max_slots) * p.part_stride;
  cudaError_t e = launch_fwd<false>(num_sms, q, nq, k, nk, d, p, qres, smem, stream, err);
  if (e != cudaSuccess) return e;
  lse_combine_kernel<<<(p.nq + 255) / 256, 256, 0, stream>>>(p, lse, esum);
  count_launches(2);
  return cudaGetLastError();
}

cudaError_t launch_sim_rank_fwd(int num_sms, const void* q, long long nq, const void* k,
                                long long nk, long long d, const float* qinv, const float* kinv,
                                const float* thr, const long long* tgt, long long* rank,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream,
                                const char** err) {
  FwdParams p;
  bool qres;
  size_t smem;
  fwd_plan(num_sms, nq, nk, d, &p, &qres, &smem);
  const size_t need = size_t(3) * p.max_slots * p.part_stride * sizeof(float);
  if (workspace_bytes < need) {
    *err = "sim_rank_fwd: workspace too small";
    return cudaErrorInvalidValue;
  }
  p.scale = nullptr;
  p.qinv = qinv;
  p.kinv = kinv;
  p.thr = thr;
  p.tgt = tgt;
  p.part_m = static_cast<float*>(workspace);
  p.part_l = p.part_m + size_t(p.max_slots) * p.part_stride;
  p.part_t = p.part_l + size_t(p.max_slots) * p.part_stride;
  cudaError_t e = launch_fwd<true>(num_sms, q, nq, k, nk, d, p, qres, smem, stream, err);
  if (e != cudaSuccess) return e;
  rank_combine_kernel<<<(p.nq + 255) / 256, 256, 0, stream>>>(p, rank);
  count_launches(2);
  return cudaGetLastError();
}

}  // namespace mmcl

// Forward hot kernel: S = s * Q K^T in 128x256 fp32 tiles that live only in tensor memory,
// reduced on the fly to per-row (max, sum-exp, sum-exp*logit) and the positive logit.
//
// One launch covers a LIST of sub-problems (fused.h, FwdSub): every (pair, direction) of the
// loss, each optionally cut into "local keys first" and "remote keys" parts, laid end to end in
// one linear tile space that is cut into equal contiguous ranges, one per CTA ("stream-K" over
// the key axis), so all 148 SMs finish together whatever the shapes are; a row's partial
// statistics from different CTAs are merged afterwards (lse_combine_kernel / the step's
// post-forward kernel).
//
// Roles (384 threads, 1 CTA / SM, persistent):
//   warp 0    TMA producer  (Q tile resident in smem when it fits, K streamed in 64-wide k-chunks).
//             With several ranks the keys of other ranks arrive in this rank's arena by peer
//             stores; before the first load from source rank r the producer acquires r's ready
//             flag (ld.acquire.sys + fence.proxy.async) -- there is no fence kernel in front.
//   warp 1    tcgen05.mma issuer (one elected thread), double-buffered TMEM accumulators
//   warps 2-3 several ranks: the all-gather itself.  They copy this rank's packed rows / norms /
//             ids from the local arena slot to the same offset of every peer (16-byte stores over
//             NVLink) while the tensor cores work on the local keys, then release the per-channel
//             ready flags.  (warp 2 also owns the TMEM allocation.)
//   warps 4-11 epilogue: tcgen05.ld -> online log-sum-exp in registers.  Two warps per SM
//             sub-partition: warps 4-7 reduce columns 0-127 of each 128x256 tile, warps 8-11
//             columns 128-255 (one row per thread, each half gives its own partial statistics).
//             With one warp per sub-partition the epilogue (dependent ALU latency, MUFU) took
//             ~6500 cycles per tile against ~4200 for the MMAs (ncu stall samples, round 1).
#include "fused.h"
#include "fwd_combine.cuh"
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

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Spin until *flag >= seq; gives up after ~4 s (records kErrFlagTimeout) so that a rank that never
// shows up turns into a failed verification instead of a hung GPU.
__device__ __noinline__ void wait_flag(const unsigned long long* flag, unsigned long long seq,
                                       unsigned long long* err) {
  if (ld_acquire_sys_u64(flag) >= seq) return;
  const unsigned long long t0 = global_timer_ns();
  unsigned spins = 0;
  while (ld_acquire_sys_u64(flag) < seq) {
    if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 4000000000ull) {
      if (err) atomicExch(err, kErrFlagTimeout);
      return;
    }
  }
}

// sub-problem that holds linear tile t (t only grows: `s` is carried along)
__device__ __forceinline__ int find_sub(const FwdLaunch& L, long long t, int s) {
  while (s + 1 < L.nsub && t >= L.sub[s + 1].tile_begin) ++s;
  return s;
}

// kMode == 3: binary-cross-entropy row statistics of the modality-alignment term
// (contrastive.py:344-413): the all-vs-all logits x_ij = s * cos_ij with targets t_ij = 1 on the
// band j in [i, band_hi[i]) of row i (diagonal + the reference's upper-triangular same-modality
// blocks, built from tile coordinates).  Per row: sum softplus(x), and over the band: sum
// softplus(x), sum x; plus sum sigmoid(x) x over all columns and over the band (d loss / d scale).
// kMode == 0: online log-sum-exp epilogue (the loss).  kMode == 2: the logits themselves are
// stored, out[i, j] = out_scale * qinv_i * kinv_j * (q_i . k_j) (zero-shot classification scores:
// few keys, every score needed by the metric collection).  kMode == 1: the same mainloop with a
// counting epilogue -- per row, the number of keys whose similarity exceeds the row's threshold
// (retrieval rank of the positive, mmcl_sim_rank_fwd); one multiply + compare per logit, no MUFU.
template <bool kQRes, int kMode>
__global__ void __launch_bounds__(384, 1)
sim_lse_fwd_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ FwdLaunch L) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operands need 1024-byte alignment
  const int stage_bytes = kQRes ? KCHUNK_BYTES : QCHUNK_BYTES + KCHUNK_BYTES;
  uint8_t* sQ = smem;
  uint8_t* sStage = smem + (kQRes ? L.kchunks * QCHUNK_BYTES : 0);
  float* sKinv = reinterpret_cast<float*>(sStage + L.nstages * stage_bytes);  // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKinv + 2 * BN);
  uint64_t* full = bars;
  uint64_t* empty = bars + L.nstages;
  uint64_t* q_full = bars + 2 * L.nstages;
  uint64_t* q_empty = q_full + 1;
  uint64_t* acc_full = q_full + 2;   // [2]
  uint64_t* acc_empty = q_full + 4;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < L.nstages; ++i) {
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

  // this CTA's share [t_begin, t_end) of phase ph
  auto phase_range = [&](int ph, long long* t_begin, long long* t_end) {
    const FwdPhase& P = L.phase[ph];
    *t_begin = P.tile_begin + (long long)blockIdx.x * P.tiles_per_cta;
    long long e = *t_begin + P.tiles_per_cta;
    if (e > P.tile_begin + P.total) e = P.tile_begin + P.total;
    *t_end = e;
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0, si = 0;
      uint32_t phase = 0, qphase = 0;
      int seen_chan = -1;
      uint32_t seen = 0;
      for (int ph = 0; ph < L.nphase; ++ph) {
      long long t_begin, t_end;
      phase_range(ph, &t_begin, &t_end);
      for (long long t = t_begin; t < t_end;) {
        si = find_sub(L, t, si);
        const FwdSub& S = L.sub[si];
        const long long local = t - S.tile_begin;
        const int qt = int(local / S.nkt_sub), j0 = int(local % S.nkt_sub);
        int j1 = j0 + int(t_end - t);
        if (j1 > S.nkt_sub) j1 = S.nkt_sub;
        const CUtensorMap* tmQ = &maps.tm[S.map_q];
        const CUtensorMap* tmK = &maps.tm[S.map_k];
        if (kQRes) {
          mbar_wait(q_empty, qphase ^ 1);
          mbar_arrive_expect_tx(q_full, L.kchunks * QCHUNK_BYTES);
          for (int kc = 0; kc < L.kchunks; ++kc)
            tma_load_2d(sQ + kc * QCHUNK_BYTES, tmQ, kc * BK, qt * BM, q_full);
          qphase ^= 1;
        }
        if (S.chan != seen_chan) {
          seen_chan = S.chan;
          seen = 0;
        }
        for (int j = j0; j < j1; ++j) {
          int kt = S.kt0 + j;
          if (kt >= S.nkt_all) kt -= S.nkt_all;
          if (L.world > 1) {
            // rows [kt*BN, kt*BN + BN) of the gathered keys: which ranks wrote them?
            const long long g0 = (long long)kt * BN;
            long long g1 = g0 + BN - 1;
            if (g1 > S.nk - 1) g1 = S.nk - 1;
            const int r0 = int(g0 / L.rows_per_rank), r1 = int(g1 / L.rows_per_rank);
            bool waited = false;
            for (int r = r0; r <= r1 && r < L.world; ++r) {
              if (r == L.rank || ((seen >> r) & 1u)) continue;
              wait_flag(L.flags + S.chan * kMaxRanks + r, L.seq, L.err);
              seen |= 1u << r;
              waited = true;
            }
            // peer stores (generic proxy, observed through the acquire) -> TMA (async proxy)
            if (waited) fence_proxy_async_all();
          }
          for (int kc = 0; kc < L.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], stage_bytes);
            uint8_t* dst = sStage + stage * stage_bytes;
            if (!kQRes) {
              tma_load_2d(dst, tmQ, kc * BK, qt * BM, &full[stage]);
              dst += QCHUNK_BYTES;
            }
            tma_load_2d(dst, tmK, kc * BK, kt * BN, &full[stage]);
            if (++stage == L.nstages) { stage = 0; phase ^= 1; }
          }
        }
        t += j1 - j0;
      }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0, si = 0;
      uint32_t phase = 0, qphase = 0;
      uint32_t acc = 0;
      for (int ph = 0; ph < L.nphase; ++ph) {
      long long t_begin, t_end;
      phase_range(ph, &t_begin, &t_end);
      for (long long t = t_begin; t < t_end;) {
        si = find_sub(L, t, si);
        const int nkt_sub = L.sub[si].nkt_sub;
        const int j0 = int((t - L.sub[si].tile_begin) % nkt_sub);
        int j1 = j0 + int(t_end - t);
        if (j1 > nkt_sub) j1 = nkt_sub;
        if (kQRes) {
          mbar_wait(q_full, qphase);
          qphase ^= 1;
        }
        for (int j = j0; j < j1; ++j) {
          const uint32_t ab = acc & 1, aphase = (acc >> 1) & 1;
          mbar_wait(&acc_empty[ab], aphase ^ 1);
          tc_fence_after();
          for (int kc = 0; kc < L.kchunks; ++kc) {
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
            if (++stage == L.nstages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&acc_full[ab]);
          ++acc;
        }
        if (kQRes) umma_commit(q_empty);
        t += j1 - j0;
      }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ================================ peer pushes ==================================
    if (L.npush > 0) {
      const long long w = (long long)blockIdx.x * 2 + (warp - 2);
      const long long nw = (long long)gridDim.x * 2;
      const char* src_base = L.peers[L.rank] + L.buf_off;
      for (int sgi = 0; sgi < L.npush; ++sgi) {
        const PushSeg& sg = L.push[sgi];
        const long long units = (long long)(sg.bytes >> 4);
        const uint4* src = reinterpret_cast<const uint4*>(src_base + sg.off);
        constexpr int U = 8;  // 8 x 512 B per warp in flight before the 7-peer store burst
        for (long long base = w * (32 * U); base < units; base += nw * (32 * U)) {
          uint4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const long long i = base + u * 32 + lane;
            if (i < units) v[u] = __ldcg(src + i);
          }
          for (int r = 0; r < L.world; ++r) {
            if (r == L.rank) continue;
            uint4* dst = reinterpret_cast<uint4*>(L.peers[r] + L.buf_off + sg.off);
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const long long i = base + u * 32 + lane;
              if (i < units) dst[i] = v[u];
            }
          }
        }
        if (sg.last_of_chan) {
          // the warp's stores happen-before the barrier; lane 0's fence is cumulative over them
          __syncwarp();
          int last = 0;
          if (lane == 0) {
            __threadfence_system();
            const unsigned long long done = atomicAdd(L.tickets + sg.chan, 1ull) + 1ull;
            last = done == L.ticket_target ? 1 : 0;  // the last push warp of this launch
            if (last) __threadfence_system();
          }
          last = __shfl_sync(0xffffffffu, last, 0);
          if (last) {
            if (sg.chan == L.status_chan && lane < L.world) {
              unsigned long long sv = L.status_val;
              for (int b = 0; b < L.nbad; ++b)
                if (__ldcg(L.bad + b) == (unsigned long long)L.gen) sv &= ~1ull;
              *reinterpret_cast<unsigned long long*>(L.peers[lane] + L.buf_off + L.status_off +
                                                     8ull * L.rank) = sv;
            }
            // one lane per peer: fence + relaxed store = a release (eight release stores from one
            // thread would serialise eight NVLink round trips)
            if (lane < L.world) {
              __threadfence_system();
              st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(L.peers[lane]) +
                                     sg.chan * kMaxRanks + L.rank,
                                 L.seq);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ epilogue ====================================
    const int wg = (warp - 4) >> 2;  // column half of every tile this warpgroup reduces
    const int q = warp & 3;          // == warp % 4: TMEM lane quadrant
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;  // 0..255: one key factor per thread and tile
    const int ch0 = wg * (BN / 64);    // first of this warpgroup's BN/64 32-column chunks
    uint32_t acc = 0;
    int si = 0;
    int kseen_chan = -1;
    uint32_t kseen = 0;
    for (int ph = 0; ph < L.nphase; ++ph) {
    long long t_begin, t_end;
    phase_range(ph, &t_begin, &t_end);
    for (long long t = t_begin; t < t_end;) {
      si = find_sub(L, t, si);
      const FwdSub& S = L.sub[si];
      const long long local = t - S.tile_begin;
      const int qt = int(local / S.nkt_sub), j0 = int(local % S.nkt_sub);
      int j1 = j0 + int(t_end - t);
      if (j1 > S.nkt_sub) j1 = S.nkt_sub;
      const float* kinv = S.kinv;
      const int nk = S.nk;
      constexpr bool kRank = kMode == 1;
      const float s = kMode == 0 ? *S.scale : 1.f;  // (modes 1-3 read what they need themselves)
      const bool neg = s < 0.f;
      float m = -INFINITY, l = 0.f, tt = 0.f;
      const int grow = qt * BM + row;
      auto tile_of = [&](int j) {
        int kt = S.kt0 + j;
        return kt >= S.nkt_all ? kt - S.nkt_all : kt;
      };
      // Key factor of column `et` of tile kt.  With several ranks the factors of another rank's
      // rows arrive by peer stores like the rows themselves: this thread acquires the source
      // rank's ready flag before its first read (the TMA producer waits on the same flag, so a
      // stall here never delays a tile that could otherwise run).
      if (S.chan != kseen_chan) {
        kseen_chan = S.chan;
        kseen = 0;
      }
      auto key_factor = [&](int kt) -> float {
        const int gk = kt * BN + et;
        if (!(kinv && gk < nk)) return 1.f;
        if (L.world > 1) {
          const int r = gk / int(L.rows_per_rank);
          if (r != L.rank && r < L.world && !((kseen >> r) & 1u)) {
            wait_flag(L.flags + S.chan * kMaxRanks + r, L.seq, L.err);
            kseen |= 1u << r;
          }
        }
        return __ldcg(kinv + gk);
      };
      float knext = key_factor(tile_of(j0));  // prefetched one tile ahead
      if (kMode == 3) {
        const float crow = (*S.scale) * ((S.qinv && grow < S.nq) ? S.qinv[grow] : 1.f);
        const long long ig = S.row_base + grow;  // this row's index among the keys
        long long bh = (grow < S.nq) ? (long long)S.band_hi[ig] : 0;
        if (bh < ig + 1) bh = ig + 1;  // the diagonal is always a positive
        float sA = 0.f, sB = 0.f, sC = 0.f, sE = 0.f, sF = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int kt = tile_of(j);
          const uint32_t ab = acc & 1, aphase = (acc >> 1) & 1;
          float* kv = sKinv + ab * BN;
          kv[et] = knext;
          if (j + 1 < j1) knext = key_factor(tile_of(j + 1));
          named_bar_sync(1, 256);
          mbar_wait(&acc_full[ab], aphase);
          tc_fence_after();
          const uint32_t taddr = tmem + ab * BN + (uint32_t(q * 32) << 16);
          const int valid = nk - kt * BN;
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
            const long long c0 = (long long)kt * BN + ch * 32;  // first column of this chunk
            float a = 0.f, b = 0.f, c = 0.f, e2 = 0.f, f2 = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float x = __uint_as_float(v[e]) * kc[e] * crow;
              const float u = ex2f(-fabsf(x) * LOG2E);        // exp(-|x|)
              const float sp = fmaxf(x, 0.f) + lg2f(1.f + u) * LN2;  // softplus(x)
              const float r = __frcp_rn(1.f + u);
              const float sx = (x >= 0.f ? r : u * r) * x;     // sigmoid(x) * x
              const bool ok = e < cv;
              const bool pos = ok && (c0 + e >= ig) && (c0 + e < bh);
              a += ok ? sp : 0.f;
              e2 += ok ? sx : 0.f;
              b += pos ? sp : 0.f;
              c += pos ? x : 0.f;
              f2 += pos ? sx : 0.f;
            }
            sA += a; sB += b; sC += c; sE += e2; sF += f2;
          }
          ++acc;
        }
        const long long first_tile = S.tile_begin - S.phase_begin + (long long)qt * S.nkt_sub;
        const int c_first = int(first_tile / S.tpc);
        const int slot = 2 * (int(blockIdx.x) - c_first) + wg;
        const long long stride = (long long)S.nqt * BM;
        const long long o = (long long)slot * stride + (long long)qt * BM + row;
        const long long plane = (long long)S.max_slots * stride;
        S.part[o] = sA;
        S.part[plane + o] = sB;
        S.part[2 * plane + o] = sC;
        S.part[3 * plane + o] = sE;
        S.part[4 * plane + o] = sF;
        t += j1 - j0;
        continue;
      }
      if (kMode == 2) {
        const float cq = S.out_scale * ((S.qinv && grow < S.nq) ? S.qinv[grow] : 1.f);
        for (int j = j0; j < j1; ++j) {
          const int kt = tile_of(j);
          const uint32_t ab = acc & 1, aphase = (acc >> 1) & 1;
          float* kv = sKinv + ab * BN;
          kv[et] = knext;
          if (j + 1 < j1) knext = key_factor(tile_of(j + 1));
          named_bar_sync(1, 256);
          mbar_wait(&acc_full[ab], aphase);
          tc_fence_after();
          const uint32_t taddr = tmem + ab * BN + (uint32_t(q * 32) << 16);
          const int valid = nk - kt * BN;
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
            if (grow < S.nq && cv > 0) {
              const float* kc = kv + ch * 32;
              float* dst = S.out + (long long)grow * S.out_ld + (long long)kt * BN + ch * 32;
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (e < cv) dst[e] = __uint_as_float(v[e]) * kc[e] * cq;
            }
          }
          ++acc;
        }
        t += j1 - j0;
        continue;  // nothing to merge afterwards
      }
      if (kRank) {
        // count_j [ qinv_i * kinv_j * (q_i . k_j) > thr_i ]  ==  [ x_ij * kinv_j > thr_i / qinv_i ]
        const float qi = (S.qinv && grow < S.nq) ? S.qinv[grow] : 1.f;
        const float th = (grow < S.nq ? S.thr[grow] : INFINITY) / qi;
        const long long tg = (S.tgt && grow < S.nq) ? S.tgt[grow] : -1;
        int cnt = 0;
        for (int j = j0; j < j1; ++j) {
          const int kt = tile_of(j);
          const uint32_t ab = acc & 1, aphase = (acc >> 1) & 1;
          float* kv = sKinv + ab * BN;
          kv[et] = knext;  // fetched while the previous tile was being reduced
          if (j + 1 < j1) knext = key_factor(tile_of(j + 1));
          named_bar_sync(1, 256);
          mbar_wait(&acc_full[ab], aphase);
          tc_fence_after();
          const uint32_t taddr = tmem + ab * BN + (uint32_t(q * 32) << 16);
          const int valid = nk - kt * BN;
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
      const float c = s * LOG2E * ((S.qinv && grow < S.nq) ? S.qinv[grow] : 1.f);
      // the positive of this row is key row label_off + grow: picked out of the tile that holds it
      const long long tg = (S.label_off >= 0 && S.pos && grow < S.nq) ? S.label_off + grow : -1;
      for (int j = j0; j < j1; ++j) {
        const int kt = tile_of(j);
        const uint32_t ab = acc & 1, aphase = (acc >> 1) & 1;
        float* kv = sKinv + ab * BN;
        kv[et] = knext;  // fetched while the previous tile was being reduced
        if (j + 1 < j1) knext = key_factor(tile_of(j + 1));
        named_bar_sync(1, 256);
        mbar_wait(&acc_full[ab], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem + ab * BN + (uint32_t(q * 32) << 16);
        const int valid = nk - kt * BN;  // columns of this tile that exist (>= 1)
        const long long tcol = tg - (long long)kt * BN;  // the positive's column in this tile, if any
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
          const long long tl = tcol - ch * 32;
          if (tl >= 0 && tl < 32 && tl < cv) {  // (rare) this chunk holds the row's positive
            float px = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e == int(tl)) px = (e & 1) ? x2[e >> 1].y : x2[e >> 1].x;
            S.pos[grow] = px * c * LN2;
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
      // partial statistics of this (sub-problem, q-tile, key range) segment
      const long long first_tile = S.tile_begin - S.phase_begin + (long long)qt * S.nkt_sub;
      const int c_first = int(first_tile / S.tpc);
      const int slot = 2 * (int(blockIdx.x) - c_first) + wg;  // one partial per column half
      const long long stride = (long long)S.nqt * BM;
      const long long o = (long long)slot * stride + (long long)qt * BM + row;
      S.part[o] = m;
      S.part[(long long)S.max_slots * stride + o] = l;
      S.part[2ll * S.max_slots * stride + o] = tt;
      t += j1 - j0;
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// Merge the per-CTA partial statistics of each row (single sub-problem launches of the ABI).
__global__ void lse_combine_kernel(FwdSub S, float* __restrict__ lse, float* __restrict__ esum) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= S.nq) return;
  float M, Lsum, T;
  fwd_combine_row(S, row, &M, &Lsum, &T);
  const float sc = (*S.scale) * (S.qinv ? S.qinv[row] : 1.f);
  lse[row] = (M + log2f(Lsum)) * LN2;
  if (esum) esum[row] = sc * T / Lsum;
}

// Rank mode: total count of each row over the key segments of different CTAs.
__global__ void rank_combine_kernel(FwdSub S, long long* __restrict__ rank) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= S.nq) return;
  const int qt = row / BM;
  const long long first_tile = S.tile_begin - S.phase_begin + (long long)qt * S.nkt_sub;
  const int c_first = int(first_tile / S.tpc);
  const int c_last = int((first_tile + S.nkt_sub - 1) / S.tpc);
  const long long stride = (long long)S.nqt * BM;
  long long total = 0;
  for (int s = 0; s < 2 * (c_last - c_first + 1); ++s)
    total += (long long)S.part[(long long)s * stride + row];
  rank[row] = total;
}

}  // namespace

// ---- planning (host) ---------------------------------------------------------------------------

void fwd_begin(FwdLaunch* L, long long d) {
  *L = FwdLaunch();
  L->nsub = 0;
  L->nphase = 0;
  L->total_tiles = 0;
  L->kchunks = int((d + BK - 1) / BK);
  L->world = 1;
  L->rank = 0;
  L->rows_per_rank = 1;
  L->npush = 0;
  L->status_chan = -1;
}

// Sub-problems must be added in phase order (phase = 0, 0, ..., 1, 1, ...).
int fwd_add_sub(FwdLaunch* L, long long nq, long long nk, int kt0, int nkt_sub, int phase) {
  if (L->nsub >= kMaxFwdSubs || phase < 0 || phase >= kMaxFwdPhases) return -1;
  if (phase + 1 < L->nphase || phase > L->nphase) return -1;
  if (phase == L->nphase) {
    FwdPhase& P = L->phase[L->nphase++];
    P.tile_begin = L->total_tiles;
    P.total = 0;
    P.tiles_per_cta = 1;
  }
  FwdSub& S = L->sub[L->nsub];
  S = FwdSub();
  S.nq = int(nq);
  S.nk = int(nk);
  S.nqt = int((nq + BM - 1) / BM);
  S.nkt_all = int((nk + BN - 1) / BN);
  S.kt0 = kt0;
  S.nkt_sub = nkt_sub < 0 ? S.nkt_all : nkt_sub;
  S.tile_begin = L->total_tiles;
  S.phase_begin = L->phase[phase].tile_begin;
  S.label_off = -1;
  S.chan = kChanRows;
  const long long tiles = (long long)S.nqt * S.nkt_sub;
  L->total_tiles += tiles;
  L->phase[phase].total += tiles;
  return L->nsub++;
}

static size_t fwd_smem_budget() { return 227 * 1024 - 256 /*barriers*/ - 2 * BN * sizeof(float); }

// Fixes the CTA split; returns the bytes of partial statistics the launch needs and the offset of
// every sub-problem's partials inside that block (part_off[i], in floats).
size_t fwd_finish(FwdLaunch* L, int num_sms, size_t* part_off) {
  long long widest = 1;
  for (int ph = 0; ph < L->nphase; ++ph)
    if (L->phase[ph].total > widest) widest = L->phase[ph].total;
  const long long ctas = widest < num_sms ? widest : num_sms;
  L->ctas = int(ctas);
  for (int ph = 0; ph < L->nphase; ++ph) {
    FwdPhase& P = L->phase[ph];
    P.tiles_per_cta = int((P.total + ctas - 1) / ctas);
    if (P.tiles_per_cta < 1) P.tiles_per_cta = 1;
  }
  const size_t budget = fwd_smem_budget();
  const bool qres = size_t(L->kchunks) * QCHUNK_BYTES + 3 * size_t(KCHUNK_BYTES) <= budget;
  const size_t stage_bytes = qres ? KCHUNK_BYTES : QCHUNK_BYTES + KCHUNK_BYTES;
  const size_t fixed = qres ? size_t(L->kchunks) * QCHUNK_BYTES : 0;
  int ns = int((budget - fixed) / stage_bytes);
  if (ns > 8) ns = 8;
  L->nstages = ns;
  size_t floats = 0;
  for (int i = 0; i < L->nsub; ++i) {
    FwdSub& S = L->sub[i];
    for (int ph = 0; ph < L->nphase; ++ph)
      if (S.phase_begin == L->phase[ph].tile_begin) S.tpc = L->phase[ph].tiles_per_cta;
    // x2: column halves; +1: a range may start inside the sub-problem's first CTA share
    S.max_slots = 2 * ((S.nkt_sub + S.tpc - 1) / S.tpc + 1);
    if (part_off) part_off[i] = floats;
    floats += size_t(3) * S.max_slots * size_t(S.nqt) * BM;
  }
  return floats * sizeof(float);
}

static bool fwd_qres(const FwdLaunch& L) {
  return size_t(L.kchunks) * QCHUNK_BYTES + 3 * size_t(KCHUNK_BYTES) <= fwd_smem_budget();
}

template <int kMode>
static cudaError_t launch_fwd(const TensorMaps& maps, const FwdLaunch& L, cudaStream_t stream) {
  const bool qres = fwd_qres(L);
  const size_t stage_bytes = qres ? KCHUNK_BYTES : QCHUNK_BYTES + KCHUNK_BYTES;
  const size_t smem = (qres ? size_t(L.kchunks) * QCHUNK_BYTES : 0) + L.nstages * stage_bytes +
                      2 * BN * sizeof(float) + 256;
  const int ctas = L.ctas;
  cudaError_t e;
  if (qres) {
    e = cudaFuncSetAttribute(sim_lse_fwd_kernel<true, kMode>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    sim_lse_fwd_kernel<true, kMode><<<ctas, 384, smem, stream>>>(maps, L);
  } else {
    e = cudaFuncSetAttribute(sim_lse_fwd_kernel<false, kMode>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    sim_lse_fwd_kernel<false, kMode><<<ctas, 384, smem, stream>>>(maps, L);
  }
  count_launches(1);
  return cudaGetLastError();
}

int fwd_grid(const FwdLaunch& L) { return L.ctas; }

cudaError_t launch_fwd_fused(const TensorMaps& maps, const FwdLaunch& L, cudaStream_t stream) {
  if (L.total_tiles == 0) return cudaSuccess;
  return launch_fwd<0>(maps, L, stream);
}

int fwd_make_maps(TensorMaps* maps, int slot, const void* rows, long long n, long long d) {
  if (slot < 0 || 2 * slot + 1 >= kMaxFwdMaps) return -1;
  if (make_tmap_bf16_sw128(&maps->tm[2 * slot], rows, n, d, d, BM)) return -1;
  if (make_tmap_bf16_sw128(&maps->tm[2 * slot + 1], rows, n, d, d, BN)) return -1;
  return 0;
}

size_t sim_lse_workspace_bytes(int num_sms, long long nq, long long nk, long long d) {
  FwdLaunch L;
  fwd_begin(&L, d);
  fwd_add_sub(&L, nq, nk, 0, -1, 0);
  return fwd_finish(&L, num_sms, nullptr);
}

namespace {

cudaError_t single_problem(int num_sms, const void* q, long long nq, const void* k, long long nk,
                           long long d, void* workspace, size_t workspace_bytes, FwdLaunch* L,
                           TensorMaps* maps, const char** err, const char* who) {
  fwd_begin(L, d);
  fwd_add_sub(L, nq, nk, 0, -1, 0);
  size_t off = 0;
  const size_t need = fwd_finish(L, num_sms, &off);
  if (workspace_bytes < need) {
    *err = who;
    return cudaErrorInvalidValue;
  }
  if (make_tmap_bf16_sw128(&maps->tm[0], q, nq, d, d, BM) ||
      make_tmap_bf16_sw128(&maps->tm[1], k, nk, d, d, BN)) {
    *err = "sim_fwd: cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  L->sub[0].map_q = 0;
  L->sub[0].map_k = 1;
  L->sub[0].part = static_cast<float*>(workspace);
  return cudaSuccess;
}

}  // namespace

cudaError_t launch_sim_lse_fwd(int num_sms, const void* q, long long nq, const void* k,
                               long long nk, long long d, const float* scale, const float* qinv,
                               const float* kinv, float* lse, float* esum, void* workspace,
                               size_t workspace_bytes, cudaStream_t stream, const char** err) {
  FwdLaunch L;
  TensorMaps maps;
  cudaError_t e = single_problem(num_sms, q, nq, k, nk, d, workspace, workspace_bytes, &L, &maps,
                                 err, "sim_lse_fwd: workspace too small");
  if (e != cudaSuccess) return e;
  L.sub[0].scale = scale;
  L.sub[0].qinv = qinv;
  L.sub[0].kinv = kinv;
  e = launch_fwd<0>(maps, L, stream);
  if (e != cudaSuccess) return e;
  lse_combine_kernel<<<(L.sub[0].nq + 255) / 256, 256, 0, stream>>>(L.sub[0], lse, esum);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_sim_rank_fwd(int num_sms, const void* q, long long nq, const void* k,
                                long long nk, long long d, const float* qinv, const float* kinv,
                                const float* thr, const long long* tgt, long long* rank,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream,
                                const char** err) {
  FwdLaunch L;
  TensorMaps maps;
  cudaError_t e = single_problem(num_sms, q, nq, k, nk, d, workspace, workspace_bytes, &L, &maps,
                                 err, "sim_rank_fwd: workspace too small");
  if (e != cudaSuccess) return e;
  L.sub[0].qinv = qinv;
  L.sub[0].kinv = kinv;
  L.sub[0].thr = thr;
  L.sub[0].tgt = tgt;
  e = launch_fwd<1>(maps, L, stream);
  if (e != cudaSuccess) return e;
  rank_combine_kernel<<<(L.sub[0].nq + 255) / 256, 256, 0, stream>>>(L.sub[0], rank);
  count_launches(1);
  return cudaGetLastError();
}

namespace {

// Sum of the per-CTA partial BCE statistics of each row: stats[k * nq + row], k = 0..4.
__global__ void bce_combine_kernel(FwdSub S, float* __restrict__ stats) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= S.nq) return;
  const int qt = row / BM;
  const long long first_tile = S.tile_begin - S.phase_begin + (long long)qt * S.nkt_sub;
  const int c_first = int(first_tile / S.tpc);
  const int c_last = int((first_tile + S.nkt_sub - 1) / S.tpc);
  const long long stride = (long long)S.nqt * BM;
  const long long plane = (long long)S.max_slots * stride;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int sl = 0; sl < 2 * (c_last - c_first + 1); ++sl)
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[k] += S.part[k * plane + (long long)sl * stride + row];
#pragma unroll
  for (int k = 0; k < 5; ++k) stats[(long long)k * S.nq + row] = acc[k];
}

}  // namespace

size_t sim_bce_workspace_bytes(int num_sms, long long nq, long long nk, long long d) {
  return sim_lse_workspace_bytes(num_sms, nq, nk, d) / 3 * 5;  // five statistics instead of three
}

cudaError_t launch_sim_bce_fwd(int num_sms, const void* q, long long nq, const void* k,
                               long long nk, long long d, const float* scale, const float* qinv,
                               const float* kinv, long long row_base, const int* band_hi,
                               float* stats, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream, const char** err) {
  FwdLaunch L;
  TensorMaps maps;
  cudaError_t e = single_problem(num_sms, q, nq, k, nk, d, workspace, workspace_bytes / 5 * 3, &L,
                                 &maps, err, "sim_bce_fwd: workspace too small");
  if (e != cudaSuccess) return e;
  L.sub[0].scale = scale;
  L.sub[0].qinv = qinv;
  L.sub[0].kinv = kinv;
  L.sub[0].row_base = row_base;
  L.sub[0].band_hi = band_hi;
  e = launch_fwd<3>(maps, L, stream);
  if (e != cudaSuccess) return e;
  bce_combine_kernel<<<(L.sub[0].nq + 255) / 256, 256, 0, stream>>>(L.sub[0], stats);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_sim_logits_fwd(int num_sms, const void* q, long long nq, const void* k,
                                  long long nk, long long d, float scale, const float* qinv,
                                  const float* kinv, float* out, long long out_ld,
                                  cudaStream_t stream, const char** err) {
  FwdLaunch L;
  TensorMaps maps;
  fwd_begin(&L, d);
  fwd_add_sub(&L, nq, nk, 0, -1, 0);
  fwd_finish(&L, num_sms, nullptr);
  if (make_tmap_bf16_sw128(&maps.tm[0], q, nq, d, d, BM) ||
      make_tmap_bf16_sw128(&maps.tm[1], k, nk, d, d, BN)) {
    *err = "sim_logits_fwd: cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  L.sub[0].map_q = 0;
  L.sub[0].map_k = 1;
  L.sub[0].qinv = qinv;
  L.sub[0].kinv = kinv;
  L.sub[0].out = out;
  L.sub[0].out_ld = out_ld;
  L.sub[0].out_scale = scale;
  return launch_fwd<2>(maps, L, stream);
}

}  // namespace mmcl

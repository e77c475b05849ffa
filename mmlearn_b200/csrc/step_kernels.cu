// The small kernels of the fused step (step.cu) around the two tensor-core launches:
//   step_pack_kernel      every modality's rows -> bf16 + fp32 inverse norms into this rank's slot
//                         of the gather buffer, the example ids into their slot, and the elementwise
//                         id comparison of every pair (one launch for the whole step)
//   step_post_fwd_kernel  merges the forward kernel's partial row statistics of both directions of
//                         every pair, writes lse / esum / pos, pushes the log-sum-exp vectors to
//                         every rank, reduces the loss and d(loss)/d(scale) in a fixed order,
//                         checks that the gathered key ids are distinct (generation-stamped hash
//                         table: no memset), and its last block writes the step's verification
//                         word and releases the log-sum-exp ready flags
//   step_motions_kernel   what a rank whose batch does not fit the assumed structure runs instead:
//                         publishes "not ok", releases and awaits every flag, so nobody hangs
//   step_wait_flags_kernel  stream-ordered wait on one flag channel (kernels without flag support)
//   step_unpack_kernel    gradient unpack of several (pair, side) problems in one launch, adding
//                         the backward kernel's scratch tiles (segments that continue an item)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "fwd_combine.cuh"
#include "kernels.h"
#include "step_kernels.h"

namespace mmcl {

namespace {

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
__device__ void wait_flag(const unsigned long long* flag, unsigned long long seq,
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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void load8(int dt, const void* base, long long elem, float (&f)[8]) {
  if (dt == 0) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else if (dt == 1) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const __half*>(base) + elem));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}

__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void store8(int odt, void* base, long long elem, const float (&o)[8]) {
  if (odt == 0) {
    float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + elem);
    p[0] = make_float4(o[0], o[1], o[2], o[3]);
    p[1] = make_float4(o[4], o[5], o[6], o[7]);
  } else if (odt == 1) {
    uint4 u;
    u.x = bf16x2(o[0], o[1]);
    u.y = bf16x2(o[2], o[3]);
    u.z = bf16x2(o[4], o[5]);
    u.w = bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + elem) = u;
  } else {
    uint4 u;
    __half2 h;
    h = __floats2half2_rn(o[0], o[1]); u.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(o[2], o[3]); u.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(o[4], o[5]); u.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(o[6], o[7]); u.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(static_cast<__half*>(base) + elem) = u;
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void step_pack_kernel(const __grid_constant__ PackLaunch L) {
  __shared__ int s_last;
  const int m = blockIdx.y;
  const PackMod& md = L.mod[m];
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && m == 0 && threadIdx.x == 0) {
    L.hdr[kHdrErr] = 0ull;
    L.hdr[kHdrVerify] = 0ull;
  }
  const int d = L.d;
  if (row < L.rows) {
    float inv = 1.f;
    if (L.normalize) {
      float ss = 0.f;
      for (int c = lane * 8; c < d; c += 256) {
        float f[8];
        load8(md.dtype, md.src, row * d + c, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss = fmaf(f[i], f[i], ss);
      }
      ss = warp_sum(ss);
      inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
    }
    for (int c = lane * 8; c < d; c += 256) {
      float f[8];
      load8(md.dtype, md.src, row * d + c, f);
      uint4 o;
      o.x = bf16x2(f[0], f[1]);
      o.y = bf16x2(f[2], f[3]);
      o.z = bf16x2(f[4], f[5]);
      o.w = bf16x2(f[6], f[7]);
      if (md.fanout) {
        for (int r = 0; r < L.world; ++r)
          *reinterpret_cast<uint4*>(L.peers[r] + L.buf_off + md.off_rows + (size_t(row) * d + c) * 2) = o;
      } else {
        *reinterpret_cast<uint4*>(md.dst + row * d + c) = o;
      }
    }
    const longlong2 mine = __ldg(reinterpret_cast<const longlong2*>(md.ids) + row);
    if (md.fanout) {
      if (lane < L.world) {
        char* base = L.peers[lane] + L.buf_off;
        if (md.inv) reinterpret_cast<float*>(base + md.off_inv)[row] = inv;
        if (md.ids_dst) reinterpret_cast<longlong2*>(base + md.off_ids)[row] = mine;
      }
    } else if (lane == 0) {
      if (md.inv) md.inv[row] = inv;
      if (md.ids_dst) reinterpret_cast<longlong2*>(md.ids_dst)[row] = mine;
    }
    if (lane == 0) {
      // every pair whose first modality this is: the local match is the identity only if the ids
      // of both modalities agree row by row (distinctness is checked on the gathered key ids later)
      for (int p = 0; p < L.P; ++p) {
        if (L.pair_a[p] != m) continue;
        const longlong2 other =
            __ldg(reinterpret_cast<const longlong2*>(L.mod[L.pair_b[p]].ids) + row);
        if (other.x != mine.x || other.y != mine.y) L.hdr[kHdrBad + p] = L.gen;
      }
    }
  }
  if (!md.fanout) return;
  // last block of this modality: release its ready flags on every rank
  __shared__ int s_last_all;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned long long done = atomicAdd(L.tickets + md.chan, 1ull) + 1ull;
    s_last = done == L.ticket_target ? 1 : 0;
    s_last_all = 0;
    if (L.status_chan >= 0)
      s_last_all = atomicAdd(L.all_ticket, 1ull) + 1ull == L.all_target ? 1 : 0;
    if (s_last || s_last_all) __threadfence_system();
  }
  __syncthreads();
  if (s_last && md.chan != L.status_chan && int(threadIdx.x) < L.world) {
    __threadfence_system();
    st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(L.peers[threadIdx.x]) +
                           md.chan * kMaxRanks + L.rank,
                       L.seq);
  }
  if (s_last_all && int(threadIdx.x) < L.world) {
    // every block of every modality is done: the id comparisons are final
    unsigned long long sv = L.status_val;
    for (int p = 0; p < L.P; ++p)
      if (__ldcg(L.hdr + kHdrBad + p) == (unsigned long long)L.gen) sv &= ~1ull;
    *reinterpret_cast<unsigned long long*>(L.peers[threadIdx.x] + L.buf_off + L.status_off +
                                           8ull * L.rank) = sv;
    __threadfence_system();
    st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(L.peers[threadIdx.x]) +
                           L.status_chan * kMaxRanks + L.rank,
                       L.seq);
  }
}

// ------------------------------------------------------------------------------------------------
// generation-stamped hash insert with duplicate detection: entry = gen << 32 | row
__device__ __forceinline__ unsigned long long mix64(long long a, long long b) {
  unsigned long long x = (unsigned long long)a * 0x9E3779B97F4A7C15ull;
  x ^= (unsigned long long)b + 0xD6E8FEB86659FD93ull + (x << 6) + (x >> 2);
  x ^= x >> 32;
  x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 29;
  return x;
}

__device__ void id_insert(const PostIds& I, long long j, unsigned gen, unsigned long long* hdr) {
  const longlong2* ids = reinterpret_cast<const longlong2*>(I.ids);
  const longlong2 mine = __ldcg(ids + j);
  const unsigned long long T = I.table_size;
  unsigned long long h = mix64(mine.x, mine.y) & (T - 1);
  const unsigned long long want = (static_cast<unsigned long long>(gen) << 32) | (unsigned long long)j;
  for (unsigned long long probes = 0; probes < T; ++probes) {
    unsigned long long cur = __ldcg(I.table + h);
    if (unsigned(cur >> 32) != gen) {  // empty for this generation
      const unsigned long long old = atomicCAS(I.table + h, cur, want);
      if (old == cur) return;  // inserted
      cur = old;
      if (unsigned(cur >> 32) != gen) {
        --probes;  // another stale value: retry the same slot
        continue;
      }
    }
    const long long o = (long long)(cur & 0xffffffffull);
    if (o == j) return;
    if (o < I.n) {
      const longlong2 other = __ldcg(ids + o);
      if (other.x == mine.x && other.y == mine.y) {
        hdr[kHdrDup] = gen;  // duplicate example id (benign race: same value from every writer)
        return;
      }
    }
    h = (h + 1) & (T - 1);
  }
  hdr[kHdrDup] = gen;  // table full of foreign entries: cannot vouch for distinctness
}

__global__ void __launch_bounds__(256)
step_post_fwd_kernel(const __grid_constant__ FwdLaunch F, const __grid_constant__ PostLaunch L) {
  __shared__ double sh[2][8];
  __shared__ float shr[4][8];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (int(blockIdx.x) < L.main_blocks) {
    // ---- combine + loss terms of 256 rows of one pair ----
    const int p = blockIdx.x / L.blocks_per_pair;
    const long long row = (long long)(blockIdx.x % L.blocks_per_pair) * 256 + tid;
    const PostPair& pp = L.pair[p];
    double lterm = 0.0, eterm = 0.0;
    float rmin[2] = {INFINITY, INFINITY}, rmax[2] = {-INFINITY, -INFINITY};
    if (row < L.rows) {
      float lse[2], es[2];
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        float M = -INFINITY, S = 0.f, T = 0.f;
        for (int k = 0; k < pp.nsub[dir]; ++k)
          fwd_merge_row(F.sub[pp.sub[dir][k]], int(row), &M, &S, &T);
        const FwdSub& s0 = F.sub[pp.sub[dir][0]];
        const float sc = (*s0.scale) * (s0.qinv ? s0.qinv[row] : 1.f);
        lse[dir] = (M + log2f(S)) * LN2;
        es[dir] = sc * T / S;
      }
      float* vec = pp.vec;
      const float pos = vec[4 * L.rows + row];
      rmin[0] = rmax[0] = lse[0];
      rmin[1] = rmax[1] = lse[1];
      vec[row] = lse[0];
      vec[L.rows + row] = lse[1];
      vec[2 * L.rows + row] = es[0];
      vec[3 * L.rows + row] = es[1];
      const double p2 = 2.0 * (double)pos;
      lterm = (double)lse[0] + (double)lse[1] - p2;
      eterm = (double)es[0] + (double)es[1] - p2;
      if (pp.push_lse) {  // the column terms of every rank's backward need these
        for (int r = 0; r < L.world; ++r) {
          char* base = L.peers[r] + L.buf_off;
          reinterpret_cast<float*>(base + pp.lse_off[0])[L.row_off + row] = lse[0];
          reinterpret_cast<float*>(base + pp.lse_off[1])[L.row_off + row] = lse[1];
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lterm += __shfl_xor_sync(0xffffffffu, lterm, o);
      eterm += __shfl_xor_sync(0xffffffffu, eterm, o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        rmin[k] = fminf(rmin[k], __shfl_xor_sync(0xffffffffu, rmin[k], o));
        rmax[k] = fmaxf(rmax[k], __shfl_xor_sync(0xffffffffu, rmax[k], o));
      }
    if (lane == 0) {
      sh[0][w] = lterm;
      sh[1][w] = eterm;
      shr[0][w] = rmin[0]; shr[1][w] = rmax[0]; shr[2][w] = rmin[1]; shr[3][w] = rmax[1];
    }
    __syncthreads();
    if (tid == 0) {
      double l = 0.0, e = 0.0;
      float r4[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
      for (int i = 0; i < 8; ++i) {
        l += sh[0][i];
        e += sh[1][i];
        r4[0] = fminf(r4[0], shr[0][i]); r4[1] = fmaxf(r4[1], shr[1][i]);
        r4[2] = fminf(r4[2], shr[2][i]); r4[3] = fmaxf(r4[3], shr[3][i]);
      }
      L.blk_part[2 * blockIdx.x] = l;
      L.blk_part[2 * blockIdx.x + 1] = e;
      for (int k = 0; k < 4; ++k) L.blk_range[4 * blockIdx.x + k] = r4[k];
    }
  } else {
    // ---- distinctness of the (gathered) key ids ----
    int blk = blockIdx.x - L.main_blocks;
    for (int k = 0; k < L.nids; ++k) {
      const PostIds& I = L.ids[k];
      const int nb = int((I.n + 255) / 256);
      if (blk < nb) {
        const long long j = (long long)blk * 256 + tid;
        if (j < I.n) id_insert(I, j, L.gen, L.hdr);
        break;
      }
      blk -= nb;
    }
  }
  // ---- last block: loss, verification word, log-sum-exp ready flags ----
  // (the block's peer stores happen-before the barrier; thread 0's fence is cumulative over them)
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    const unsigned long long done = atomicAdd(L.ticket, 1ull) + 1ull;
    s_last = done == L.ticket_target ? 1 : 0;
    if (s_last) __threadfence_system();
  }
  __syncthreads();
  if (!s_last) return;
  // per pair: block partials summed by all 256 threads in a FIXED order (thread t takes blocks
  // t, t + 256, ...; then a fixed shuffle / shared-memory tree): bit-reproducible loss
  for (int p = 0; p < L.P; ++p) {
    const PostPair& pp = L.pair[p];
    const int b0 = p * L.blocks_per_pair;
    double l = 0.0, e = 0.0;
    float r4[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
    for (int i = tid; i < L.blocks_per_pair; i += 256) {
      l += __ldcg(L.blk_part + 2 * (b0 + i));
      e += __ldcg(L.blk_part + 2 * (b0 + i) + 1);
      const float4 q = __ldcg(reinterpret_cast<const float4*>(L.blk_range) + (b0 + i));
      r4[0] = fminf(r4[0], q.x); r4[1] = fmaxf(r4[1], q.y);
      r4[2] = fminf(r4[2], q.z); r4[3] = fmaxf(r4[3], q.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, o);
      e += __shfl_xor_sync(0xffffffffu, e, o);
      r4[0] = fminf(r4[0], __shfl_xor_sync(0xffffffffu, r4[0], o));
      r4[1] = fmaxf(r4[1], __shfl_xor_sync(0xffffffffu, r4[1], o));
      r4[2] = fminf(r4[2], __shfl_xor_sync(0xffffffffu, r4[2], o));
      r4[3] = fmaxf(r4[3], __shfl_xor_sync(0xffffffffu, r4[3], o));
    }
    __syncthreads();  // (shared arrays are reused from pair to pair)
    if (lane == 0) {
      sh[0][w] = l;
      sh[1][w] = e;
      shr[0][w] = r4[0]; shr[1][w] = r4[1]; shr[2][w] = r4[2]; shr[3][w] = r4[3];
    }
    __syncthreads();
    if (tid == 0) {
      l = e = 0.0;
      for (int i = 0; i < 8; ++i) {
        l += sh[0][i];
        e += sh[1][i];
        r4[0] = fminf(r4[0], shr[0][i]); r4[1] = fmaxf(r4[1], shr[1][i]);
        r4[2] = fminf(r4[2], shr[2][i]); r4[3] = fmaxf(r4[3], shr[3][i]);
      }
      L.out2[2 * p] = (float)(pp.coef * l);
      L.out2[2 * p + 1] = (float)(pp.coef * e / (double)(*pp.scale));
      // (min, max) of this rank's log-sum-exps: the backward kernel decides from them (and from
      // the other ranks') whether its one-exponential form is safe, instead of scanning the vectors
      for (int k = 0; k < 4; ++k) pp.range[k] = r4[k];
      shr[0][0] = r4[0]; shr[1][0] = r4[1]; shr[2][0] = r4[2]; shr[3][0] = r4[3];
    }
    __syncthreads();
    if (pp.push_lse && tid < L.world) {
      float* q = reinterpret_cast<float*>(L.peers[tid] + L.buf_off + pp.range_off) + 4 * L.rank;
      for (int k = 0; k < 4; ++k) q[k] = shr[k][0];
    }
  }
  __syncthreads();  // the range pushes happen-before the flag release below
  if (tid == 0) {
    bool ok = L.local_ok != 0 && __ldcg(L.hdr + kHdrErr) == 0ull &&
              __ldcg(L.hdr + kHdrDup) != (unsigned long long)L.gen;
    for (int p = 0; p < L.P; ++p) ok = ok && __ldcg(L.hdr + kHdrBad + p) != (unsigned long long)L.gen;
    if (L.world > 1) {
      const unsigned long long* st = reinterpret_cast<const unsigned long long*>(
          L.peers[L.rank] + L.buf_off + L.status_off);
      for (int r = 0; r < L.world; ++r) ok = ok && __ldcg(st + r) == L.status_ok;
    }
    L.hdr[kHdrVerify] = (static_cast<unsigned long long>(L.gen) << 1) | (ok ? 1ull : 0ull);
  }
  // log-sum-exp ready flags: one thread per peer (fence + relaxed store = a release; a release
  // store per peer from one thread would serialise eight round trips)
  if (L.world > 1 && L.release_lse && tid < L.world) {
    __threadfence_system();
    st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(L.peers[tid]) +
                           kChanLse * kMaxRanks + L.rank,
                       L.seq);
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void step_motions_kernel(const __grid_constant__ MotionsLaunch L) {
  const int r = threadIdx.x;
  if (r < L.world) {
    *reinterpret_cast<unsigned long long*>(L.peers[r] + L.buf_off + L.status_off + 8ull * L.rank) =
        L.status_val;
    __threadfence_system();
    for (int c = 0; c < kFlagChannels; ++c)
      if ((L.chan_mask >> c) & 1u)
        st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(L.peers[r]) + c * kMaxRanks + L.rank,
                           L.seq);
    // nobody may overwrite the other parity buffer before every rank has left the previous step
    for (int c = 0; c < kFlagChannels; ++c)
      if ((L.wait_mask >> c) & 1u) wait_flag(L.flags + c * kMaxRanks + r, L.seq, L.hdr + kHdrErr);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    L.hdr[kHdrErr] = 0ull;
    L.hdr[kHdrVerify] = static_cast<unsigned long long>(L.gen) << 1;  // not ok
  }
}

__global__ void step_wait_flags_kernel(const unsigned long long* flags, int chan,
                                       unsigned long long seq, int world, int rank,
                                       unsigned long long* err) {
  const int r = threadIdx.x;
  if (r < world && r != rank) wait_flag(flags + chan * kMaxRanks + r, seq, err);
}

// ------------------------------------------------------------------------------------------------
// Gradient unpack of several (pair, side) problems: see aux_kernels.cu unpack_grad_kernel for the
// arithmetic.  MAXV = ceil(d / 256) register-resident 8-wide column groups per lane.
template <int MAXV>
__global__ void step_unpack_kernel(const __grid_constant__ UnpackLaunch L) {
  int blk = blockIdx.x, pi = 0;
  while (pi + 1 < L.nprob && blk >= L.prob[pi].blocks) blk -= L.prob[pi++].blocks;
  const UnpackProb& P = L.prob[pi];
  const long long row = (long long)blk * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= L.rows) return;
  const int d = L.d;
  const float coef = P.weight * (L.gout ? *L.gout : 1.f) * (*P.scale);
  float g[MAXV][8];
  float x[MAXV][8];
  float dot = 0.f;
  const float inv = L.normalize ? P.inv_norm[row] : 1.f;
  const float pc = P.pos_coef * (P.partner_inv ? P.partner_inv[row] : 1.f);
  const int xt = int(row >> 7), rin = int(row & 127);
  // (converged walk with a clamped address and a 0/1 mask: see aux_kernels.cu unpack_grad_kernel)
#pragma unroll
  for (int nv = 0; nv < MAXV; ++nv) {
    const int c = lane * 8 + nv * 256;
    const bool act = c < d;
    const int cc = act ? c : 0;
    float pa[8];
    float4 a0 = __ldcg(reinterpret_cast<const float4*>(P.dacc + row * d + cc));
    float4 a1 = __ldcg(reinterpret_cast<const float4*>(P.dacc + row * d + cc + 4));
    if (P.scratch) {
      // segments of this (x tile, d pass) item that other CTA pairs continued
      const int dc = cc / L.dp, cin = cc - dc * L.dp;
      const long long t0 = ((long long)xt * L.ndc + dc) * P.nyt;
      const long long c_first = t0 / P.per_pair, c_last = (t0 + P.nyt - 1) / P.per_pair;
      for (long long cp = c_first + 1; cp <= c_last; ++cp) {
        const float* sp = P.scratch + ((cp * 128 + rin) * (long long)L.dp + cin);
        const float4 b0 = __ldcg(reinterpret_cast<const float4*>(sp));
        const float4 b1 = __ldcg(reinterpret_cast<const float4*>(sp + 4));
        a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
        a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
      }
    }
    load8(1, P.partner, row * d + cc, pa);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) g[nv][i] = act ? coef * (a[i] - pc * pa[i]) : 0.f;
    if (L.normalize) {
      load8(P.src_dtype, P.src, row * d + cc, x[nv]);
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(g[nv][i], x[nv][i] * inv, dot);
    }
  }
  __syncwarp();
  if (L.normalize) dot = warp_sum(dot);
  const bool clamped = inv >= 0.99e12f;  // ||x|| < eps: x_hat = x / eps, a plain scaling
#pragma unroll
  for (int nv = 0; nv < MAXV; ++nv) {
    const int c = lane * 8 + nv * 256;
    if (c < d) {
      float o[8];
      if (L.normalize) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = clamped ? g[nv][i] * inv : inv * (g[nv][i] - x[nv][i] * inv * dot);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = g[nv][i];
      }
      if (P.direct) {
        store8(P.grad_dtype, P.grad, row * d + c, o);
      } else {
        float* dst = static_cast<float*>(P.grad) + row * d + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dst + i, o[i]);
      }
    }
  }
}

}  // namespace

cudaError_t launch_step_pack(const PackLaunch& L, cudaStream_t stream) {
  const int wpb = 8;
  dim3 grid(unsigned((L.rows + wpb - 1) / wpb), unsigned(L.M));
  step_pack_kernel<<<grid, wpb * 32, 0, stream>>>(L);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_step_post_fwd(const FwdLaunch& F, const PostLaunch& L, cudaStream_t stream) {
  step_post_fwd_kernel<<<L.total_blocks, 256, 0, stream>>>(F, L);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_step_motions(const MotionsLaunch& L, cudaStream_t stream) {
  step_motions_kernel<<<1, 32, 0, stream>>>(L);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_step_wait_flags(const unsigned long long* flags, int chan,
                                   unsigned long long seq, int world, int rank,
                                   unsigned long long* err, cudaStream_t stream) {
  step_wait_flags_kernel<<<1, 32, 0, stream>>>(flags, chan, seq, world, rank, err);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_step_unpack(const UnpackLaunch& L, cudaStream_t stream) {
  if (L.nprob == 0) return cudaSuccess;
  int blocks = 0;
  for (int i = 0; i < L.nprob; ++i) blocks += L.prob[i].blocks;
  const int nv = (L.d + 255) / 256;
  if (nv <= 1) step_unpack_kernel<1><<<blocks, 128, 0, stream>>>(L);
  else if (nv == 2) step_unpack_kernel<2><<<blocks, 128, 0, stream>>>(L);
  else if (nv == 3) step_unpack_kernel<3><<<blocks, 128, 0, stream>>>(L);
  else if (nv == 4) step_unpack_kernel<4><<<blocks, 128, 0, stream>>>(L);
  else step_unpack_kernel<8><<<blocks, 128, 0, stream>>>(L);
  count_launches(1);
  return cudaGetLastError();
}

}  // namespace mmcl

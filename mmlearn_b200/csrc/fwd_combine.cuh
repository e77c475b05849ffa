// Merging the forward kernel's partial row statistics (device code shared by sim_fwd.cu and the
// step's post-forward kernel).  A (sub-problem, q-tile) item is covered by the CTAs
// c_first..c_last of its phase's linear tile split; CTA c wrote its partials of the item into
// slots 2*(c - c_first) and 2*(c - c_first) + 1 (one per column half of the 128x256 tiles).
#pragma once
#include "fused.h"

namespace mmcl {

// folds the partials of `row` of sub-problem S into the running (M, L, T): max (log2 units), sum
// of 2^(x - M), sum of 2^(x - M) * cos
__device__ __forceinline__ void fwd_merge_row(const FwdSub& S, int row, float* M, float* L,
                                              float* T) {
  const int qt = row / 128;
  const long long first_tile = S.tile_begin - S.phase_begin + (long long)qt * S.nkt_sub;
  const int c_first = int(first_tile / S.tpc);
  const int c_last = int((first_tile + S.nkt_sub - 1) / S.tpc);
  const int nslots = 2 * (c_last - c_first + 1);
  const long long stride = (long long)S.nqt * 128;
  const float* pm = S.part;
  const float* pl = S.part + (long long)S.max_slots * stride;
  const float* pt = S.part + 2ll * S.max_slots * stride;
  float m = *M, l = *L, t = *T;
  // batches of four slots: twelve independent loads in flight, then the (sequential) merge
  for (int s0 = 0; s0 < nslots; s0 += 4) {
    float ms[4], ls[4], ts[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool ok = s0 + k < nslots;
      const long long o = (long long)(ok ? s0 + k : s0) * stride + row;
      ms[k] = ok ? __ldcg(pm + o) : -INFINITY;
      ls[k] = ok ? __ldcg(pl + o) : 0.f;
      ts[k] = ok ? __ldcg(pt + o) : 0.f;
    }
    float mn = m;
#pragma unroll
    for (int k = 0; k < 4; ++k) mn = fmaxf(mn, ms[k]);
    if (mn == -INFINITY) continue;  // nothing seen so far
    const float a = exp2f(m - mn);
    l *= a;
    t *= a;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float b = exp2f(ms[k] - mn);
      l = fmaf(ls[k], b, l);
      t = fmaf(ts[k], b, t);
    }
    m = mn;
  }
  *M = m;
  *L = l;
  *T = t;
}

__device__ __forceinline__ void fwd_combine_row(const FwdSub& S, int row, float* M, float* L,
                                                float* T) {
  *M = -INFINITY;
  *L = 0.f;
  *T = 0.f;
  fwd_merge_row(S, row, M, L, T);
}

}  // namespace mmcl

// Example-id matching on the device (replaces mmlearn/datasets/core/example.py:101-166).
// The reference builds an [n, m] boolean matrix and calls torch.where; here the second id list
// is inserted into an open-addressing hash table (one slot per distinct id, owners found with
// a 32-bit CAS on the row index -- the 128-bit key itself is compared through the input
// array), rows with equal ids are bucketed and sorted, and each row of the first list looks
// its bucket up.  Two exclusive scans give torch.where's row-major output order:
// ascending i, then ascending j, with the full cross product for duplicates.  O(n + m).
#include "kernels.h"

namespace mmcl {

namespace {

struct MatchWs {
  int* owner;        // [T] index into ids_b owning the slot, -1 = empty
  int* count;        // [T] rows of b with this id
  int* cursor;       // [T]
  long long* start;  // [T] exclusive scan of count
  int* slot_b;       // [m]
  int* bucket;       // [m] b-row indices grouped by slot, ascending inside a group
  int* cnt_a;        // [n]
  int* slot_a;       // [n]
  long long* off_a;  // [n] exclusive scan of cnt_a
  long long* bsum;   // [max(T, n) / SCAN_TILE + 1] per-tile sums of the running scan
  long long T;
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;  // 2048 elements per block

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

long long table_size(long long m) {
  long long T = 64;
  while (T < 2 * m) T <<= 1;
  return T;
}

size_t carve(MatchWs* w, void* base, long long n, long long m) {
  const long long T = table_size(m);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + o : nullptr;
    o += align16(bytes);
    return p;
  };
  w->T = T;
  w->owner = static_cast<int*>(take(T * 4));
  w->count = static_cast<int*>(take(T * 4));
  w->cursor = static_cast<int*>(take(T * 4));
  w->start = static_cast<long long*>(take(T * 8));
  w->slot_b = static_cast<int*>(take(m * 4));
  w->bucket = static_cast<int*>(take(m * 4));
  w->cnt_a = static_cast<int*>(take(n * 4));
  w->slot_a = static_cast<int*>(take(n * 4));
  w->off_a = static_cast<long long*>(take(n * 8));
  w->bsum = static_cast<long long*>(take(((T > n ? T : n) / SCAN_TILE + 1) * 8));
  return o;
}

__device__ __forceinline__ unsigned long long mix(long long a, long long b) {
  unsigned long long h = (unsigned long long)a * 0x9E3779B97F4A7C15ull;
  h ^= (unsigned long long)b + 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
  h ^= h >> 31;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 29;
  return h;
}

__global__ void init_kernel(MatchWs w) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w.T) {
    w.owner[i] = -1;
    w.count[i] = 0;
    w.cursor[i] = 0;
  }
}

__global__ void insert_kernel(MatchWs w, const long long* __restrict__ ids_b, long long m) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const long long k0 = ids_b[2 * j], k1 = ids_b[2 * j + 1];
  long long h = (long long)(mix(k0, k1) & (unsigned long long)(w.T - 1));
  while (true) {
    int o = atomicCAS(&w.owner[h], -1, (int)j);
    if (o == -1) o = (int)j;
    if (ids_b[2 * (long long)o] == k0 && ids_b[2 * (long long)o + 1] == k1) {
      w.slot_b[j] = (int)h;
      atomicAdd(&w.count[h], 1);
      return;
    }
    h = (h + 1) & (w.T - 1);
  }
}

// Exclusive scan int -> long long in three small launches: per-tile sums, scan of the tile sums
// (one block), per-tile scan with the tile offset.  Tiles are read with coalesced accesses.
__global__ void scan_tile_sums_kernel(const int* __restrict__ in, long long n,
                                      long long* __restrict__ bsum) {
  __shared__ long long wsum[SCAN_THREADS / 32];
  const long long base = (long long)blockIdx.x * SCAN_TILE;
  long long s = 0;
#pragma unroll
  for (int e = 0; e < SCAN_PER_THREAD; ++e) {
    const long long i = base + e * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += wsum[w];
    bsum[blockIdx.x] = t;
  }
}

// one block: bsum[0..nb) -> exclusive scan in place; *total = sum
__global__ void scan_bsums_kernel(long long* __restrict__ bsum, long long nb,
                                  long long* __restrict__ total) {
  __shared__ long long part[1024];
  __shared__ long long carry_s;
  const int t = threadIdx.x;
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (long long base = 0; base < nb; base += 1024) {
    const long long i = base + t;
    const long long v = i < nb ? bsum[i] : 0;
    part[t] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const long long a = t >= o ? part[t - o] : 0;
      __syncthreads();
      part[t] += a;
      __syncthreads();
    }
    const long long carry = carry_s;
    if (i < nb) bsum[i] = carry + part[t] - v;
    __syncthreads();
    if (t == 1023) carry_s = carry + part[t];
    __syncthreads();
  }
  if (total && t == 0) *total = carry_s;
}

__global__ void scan_tiles_kernel(const int* __restrict__ in, long long* __restrict__ out,
                                  long long n, const long long* __restrict__ bsum) {
  __shared__ long long wsum[SCAN_THREADS / 32];
  // thread t owns the SCAN_PER_THREAD consecutive elements [base + t*P, base + (t+1)*P)
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_PER_THREAD;
  int v[SCAN_PER_THREAD];
  long long s = 0;
#pragma unroll
  for (int e = 0; e < SCAN_PER_THREAD; ++e) {
    v[e] = base + e < n ? in[base + e] : 0;
    s += v[e];
  }
  long long incl = s;  // inclusive scan of the per-thread sums across the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long a = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += a;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  long long run = bsum[blockIdx.x] + incl - s;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) run += wsum[w];
#pragma unroll
  for (int e = 0; e < SCAN_PER_THREAD; ++e) {
    if (base + e < n) out[base + e] = run;
    run += v[e];
  }
}

void launch_scan(const int* in, long long* out, long long n, long long* total, long long* bsum,
                 cudaStream_t stream) {
  const long long nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_tile_sums_kernel<<<unsigned(nb), SCAN_THREADS, 0, stream>>>(in, n, bsum);
  scan_bsums_kernel<<<1, 1024, 0, stream>>>(bsum, nb, total);
  scan_tiles_kernel<<<unsigned(nb), SCAN_THREADS, 0, stream>>>(in, out, n, bsum);
}

__global__ void fill_kernel(MatchWs w, long long m) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int h = w.slot_b[j];
  const int pos = atomicAdd(&w.cursor[h], 1);
  w.bucket[w.start[h] + pos] = (int)j;
}

__global__ void sort_buckets_kernel(MatchWs w) {
  const long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= w.T) return;
  const int c = w.count[h];
  if (c < 2) return;
  int* b = w.bucket + w.start[h];
  for (int i = 1; i < c; ++i) {  // insertion sort; duplicate groups are tiny in practice
    const int v = b[i];
    int k = i - 1;
    while (k >= 0 && b[k] > v) {
      b[k + 1] = b[k];
      --k;
    }
    b[k + 1] = v;
  }
}

__global__ void lookup_kernel(MatchWs w, const long long* __restrict__ ids_a, long long n,
                              const long long* __restrict__ ids_b) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long k0 = ids_a[2 * i], k1 = ids_a[2 * i + 1];
  long long h = (long long)(mix(k0, k1) & (unsigned long long)(w.T - 1));
  int cnt = 0, slot = -1;
  while (true) {
    const int o = w.owner[h];
    if (o == -1) break;
    if (ids_b[2 * (long long)o] == k0 && ids_b[2 * (long long)o + 1] == k1) {
      cnt = w.count[h];
      slot = (int)h;
      break;
    }
    h = (h + 1) & (w.T - 1);
  }
  w.cnt_a[i] = cnt;
  w.slot_a[i] = slot;
}

__global__ void emit_kernel(MatchWs w, long long n, long long m, long long* __restrict__ out_a,
                            long long* __restrict__ out_b, long long capacity,
                            long long* __restrict__ info) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cnt = w.cnt_a[i];
  const int slot = w.slot_a[i];
  const long long o = w.off_a[i];
  bool self = false;
  if (cnt > 0) {
    const int* b = w.bucket + w.start[slot];
    for (int t = 0; t < cnt; ++t) {
      if (o + t < capacity) {
        out_a[o + t] = i;
        out_b[o + t] = b[t];
      }
    }
    self = (cnt == 1 && b[0] == (int)i);
  }
  if (!self) info[1] = 0;  // benign race: every writer stores 0
}

__global__ void info_init_kernel(long long* info, long long n, long long m, int quick) {
  info[0] = (quick && n == m) ? n : 0;
  info[1] = (n == m) ? 1 : 0;
  info[2] = n;
  info[3] = m;
}

// Identity check: ids_a[j] == ids_b[j] for every j and all ids_b distinct (hash insert with
// duplicate detection).  On any violation info = {0, 0, n, m}.
__global__ void identity_check_kernel(int* __restrict__ owner, long long T,
                                      const long long* __restrict__ ids_a,
                                      const long long* __restrict__ ids_b, long long n,
                                      long long* __restrict__ info) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const long long k0 = ids_b[2 * j], k1 = ids_b[2 * j + 1];
  bool ok = ids_a[2 * j] == k0 && ids_a[2 * j + 1] == k1;
  long long h = (long long)(mix(k0, k1) & (unsigned long long)(T - 1));
  while (ok) {
    const int o = atomicCAS(&owner[h], -1, (int)j);
    if (o == -1) break;  // inserted
    if (ids_b[2 * (long long)o] == k0 && ids_b[2 * (long long)o + 1] == k1) ok = false;  // duplicate
    h = (h + 1) & (T - 1);
  }
  if (!ok) {  // benign race: every writer stores the same values
    info[0] = 0;
    info[1] = 0;
  }
}

}  // namespace

size_t match_workspace_bytes(long long n, long long m) {
  MatchWs w;
  return carve(&w, nullptr, n > 0 ? n : 1, m > 0 ? m : 1);
}

cudaError_t launch_match_ids(const int64_t* ids_a, long long n, const int64_t* ids_b, long long m,
                             int64_t* out_a, int64_t* out_b, long long capacity, int64_t* out_info,
                             void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  auto* info = reinterpret_cast<long long*>(out_info);
  info_init_kernel<<<1, 1, 0, stream>>>(info, n, m, 0);
  count_launches(1);
  if (n == 0 || m == 0) return cudaGetLastError();
  MatchWs w;
  if (carve(&w, workspace, n, m) > workspace_bytes) return cudaErrorInvalidValue;
  const int B = 256;
  auto blocks = [&](long long x) { return unsigned((x + B - 1) / B); };
  const auto* a = reinterpret_cast<const long long*>(ids_a);
  const auto* b = reinterpret_cast<const long long*>(ids_b);
  init_kernel<<<blocks(w.T), B, 0, stream>>>(w);
  insert_kernel<<<blocks(m), B, 0, stream>>>(w, b, m);
  launch_scan(w.count, w.start, w.T, nullptr, w.bsum, stream);
  fill_kernel<<<blocks(m), B, 0, stream>>>(w, m);
  sort_buckets_kernel<<<blocks(w.T), B, 0, stream>>>(w);
  lookup_kernel<<<blocks(n), B, 0, stream>>>(w, a, n, b);
  launch_scan(w.cnt_a, w.off_a, n, info, w.bsum, stream);
  emit_kernel<<<blocks(n), B, 0, stream>>>(w, n, m, reinterpret_cast<long long*>(out_a),
                                           reinterpret_cast<long long*>(out_b), capacity, info);
  count_launches(12);
  return cudaGetLastError();
}

// The identity check is a single dependent-probe pass, so its duration is the LONGEST probe chain:
// a sparser table (load factor 1/8 instead of 1/2) keeps chains to 1-3 steps.
static long long identity_table_size(long long m) { return 4 * table_size(m > 0 ? m : 1); }

size_t identity_check_workspace_bytes(long long m) { return size_t(identity_table_size(m)) * 4; }

cudaError_t launch_identity_check(const int64_t* ids_a, long long n, const int64_t* ids_b,
                                  long long m, int64_t* out_info, void* workspace,
                                  size_t workspace_bytes, cudaStream_t stream) {
  auto* info = reinterpret_cast<long long*>(out_info);
  info_init_kernel<<<1, 1, 0, stream>>>(info, n, m, 1);
  count_launches(1);
  if (n != m || n == 0) return cudaGetLastError();
  const long long T = identity_table_size(m);
  if (size_t(T) * 4 > workspace_bytes) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(workspace, 0xFF, size_t(T) * 4, stream);  // owner = -1
  if (e != cudaSuccess) return e;
  identity_check_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(
      static_cast<int*>(workspace), T, reinterpret_cast<const long long*>(ids_a),
      reinterpret_cast<const long long*>(ids_b), n, info);
  count_launches(1);
  return cudaGetLastError();
}

}  // namespace mmcl

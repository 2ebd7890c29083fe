// Exact order statistics by radix selection (sm_100a; memory-bound: 4 passes over the data instead of a full sort).
//
// replaces: the sorts behind torch.quantile / Tensor.sort in the candidate seeding of the reference --
// quant_layers/linear.py:432-481 (calculate_percentile_*_candidates), :763-814 (positive_percentile),
// matmul.py:211-240 -- where up to 77 M (per GPU; 620 M for DeiT-B / 1024 images before sharding) activations are fully
// sorted to read two order statistics per quantile.  The k-th smallest element is found digit by digit (8 bits per
// pass, most significant first) on the order-preserving integer image of the floats: a histogram of the next digit
// among the elements that match the digits found so far, then a 256-bin scan.  Between the two kernels of a pass the
// caller may all-reduce the histograms over the ranks of a data-parallel job: every rank then walks the same digits
// and ends with the k-th element of the UNION of the shards (what sort(all_gather(x))[k] would hold) after 4 small
// all-reduces, without moving the data.
#include "common.cuh"
#include "../../include/adalog_b200.h"

namespace adalog {
namespace sel {

constexpr int kMaxT = 8;            // target ranks per row
constexpr int kBins = 256;

struct State { unsigned int prefix; unsigned int pad; unsigned long long k; };   // 16 bytes per (row, target)

// float bits -> key whose unsigned order is torch.sort's order: negatives reversed below the positives, every NaN last
__device__ __forceinline__ unsigned int float_key(float v, int positive_only) {
  unsigned int u = __float_as_uint(v);
  if (positive_only && !(v > 0.0f)) return 0xFF800000u;                  // as +inf (linear.py:763-798 ranks the positives only)
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;               // NaN: last
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void select_init_kernel(State* __restrict__ st, const long long* __restrict__ ranks, long long rows, int T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * T) return;
  st[i].prefix = 0u; st[i].pad = 0u;
  st[i].k = (unsigned long long)ranks[i % T];
}

// pass p (0..3): digit = bits [24-8p, 32-8p) of the key; hist[row][rep(t)][digit] += 1 for every element whose higher
// digits equal the target's prefix.  Targets with equal prefixes share one histogram (all of them in pass 0).
__global__ void __launch_bounds__(256) select_hist_kernel(const float* __restrict__ x, long long n, long long row_stride,
                                                         const State* __restrict__ st, unsigned int* __restrict__ hist,
                                                         int T, int pass, int positive_only) {
  __shared__ unsigned int sh[kMaxT][kBins];
  __shared__ unsigned int pfx[kMaxT];
  __shared__ int rep[kMaxT], nrep_s;
  const long long row = blockIdx.y;
  const int shift = 24 - 8 * pass;
  const unsigned int hi_mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
  if (threadIdx.x == 0) {
    int nrep = 0;
    for (int t = 0; t < T; ++t) {
      const unsigned int p = st[row * T + t].prefix & hi_mask;
      int r = -1;
      for (int q = 0; q < nrep; ++q) if (pfx[q] == p) { r = q; break; }
      if (r < 0) { r = nrep; pfx[nrep++] = p; }
      rep[t] = r;
    }
    nrep_s = nrep;
  }
  for (int i = threadIdx.x; i < kMaxT * kBins; i += blockDim.x) (&sh[0][0])[i] = 0u;
  __syncthreads();
  const int nrep = nrep_s;
  const float* xr = x + row * row_stride;
  // contiguous slice of the row per CTA, float4 loads where aligned
  const long long per = (((n + gridDim.x - 1) / gridDim.x + 3) / 4) * 4;
  const long long i0 = (long long)blockIdx.x * per, i1 = min(n, i0 + per);
  const bool vec = ((reinterpret_cast<uintptr_t>(xr) & 15) == 0) && ((row_stride & 3) == 0 || row == 0);
  auto visit = [&](float v) {
    const unsigned int key = float_key(v, positive_only);
    const unsigned int d = (key >> shift) & 0xFFu;
    const unsigned int top = key & hi_mask;
    for (int q = 0; q < nrep; ++q)
      if (top == pfx[q]) atomicAdd(&sh[q][d], 1u);
  };
  if (vec) {
    for (long long i = i0 + 4 * (long long)threadIdx.x; i < i1; i += 4 * (long long)blockDim.x) {
      if (i + 3 < i1) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xr + i));
        visit(v.x); visit(v.y); visit(v.z); visit(v.w);
      } else {
        for (long long j = i; j < i1; ++j) visit(__ldg(xr + j));
      }
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) visit(__ldg(xr + i));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * kBins; i += blockDim.x) {
    const int t = i / kBins, d = i - t * kBins;
    // every target gets its own (copied) histogram: the scan then needs no representative map
    const unsigned int c = sh[rep[t]][d];
    if (c) atomicAdd(hist + (row * T + t) * kBins + d, c);
  }
}

// one warp per (row, target): find the digit whose cumulative count passes k, descend, clear the histogram
__global__ void select_scan_kernel(State* __restrict__ st, unsigned int* __restrict__ hist, long long n_rt, int pass) {
  const long long rt = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (rt >= n_rt) return;
  const int lane = threadIdx.x & 31;
  unsigned int* h = hist + rt * kBins;
  // lane l owns bins 8l .. 8l+7
  unsigned int c[8];
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = h[lane * 8 + j]; s += c[j]; h[lane * 8 + j] = 0u; }
  unsigned long long incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned long long excl = incl - s;
  const unsigned long long k = st[rt].k;
  const bool mine = k >= excl && k < incl;
  const unsigned int who = __ballot_sync(0xffffffffu, mine);
  if (who == 0) return;                     // rank beyond the element count: leave the state (caller validates ranks)
  if (lane == __ffs(who) - 1) {
    unsigned long long cum = excl;
    int d = 0;
    for (; d < 8; ++d) { if (k < cum + c[d]) break; cum += c[d]; }
    const int shift = 24 - 8 * pass;
    st[rt].prefix |= (unsigned int)(lane * 8 + d) << shift;
    st[rt].k = k - cum;
  }
}

__global__ void select_finish_kernel(const State* __restrict__ st, float* __restrict__ out, long long n_rt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rt) out[i] = key_float(st[i].prefix) + 0.0f;     // (a selected zero is returned as +0.0)
}

}  // namespace sel
}  // namespace adalog

using namespace adalog;

extern "C" {

int64_t adalog_select_workspace_bytes(int64_t rows, int T) {
  if (rows <= 0 || T <= 0 || T > sel::kMaxT) return -1;
  return rows * T * (int64_t)(sizeof(sel::State) + sel::kBins * sizeof(unsigned int));
}

// workspace layout: State[rows*T] then hist[rows*T*256] (uint32).  adalog_select_hist_ptr gives the histogram base so a
// data-parallel caller can all-reduce it between adalog_select_hist and adalog_select_scan.
void* adalog_select_hist_ptr(void* workspace, int64_t rows, int T) {
  return reinterpret_cast<uint8_t*>(workspace) + rows * T * sizeof(sel::State);
}

int adalog_select_init(void* workspace, int64_t rows, int T, const long long* ranks, void* stream) {
  ADALOG_REQUIRE(workspace && ranks && rows > 0 && T > 0 && T <= sel::kMaxT, -1, "select_init: bad arguments (T <= 8)");
  cudaStream_t s = (cudaStream_t)stream;
  auto* st = reinterpret_cast<sel::State*>(workspace);
  cudaMemsetAsync(adalog_select_hist_ptr(workspace, rows, T), 0, rows * T * sel::kBins * sizeof(unsigned int), s);
  const long long n = rows * T;
  sel::select_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(st, ranks, rows, T);
  return check_launch("select_init");
}

// histogram of pass `pass` over this rank's rows [row0, row0 + local_rows) of the workspace's `rows`
int adalog_select_hist(const float* x, int64_t local_rows, int64_t n, int64_t row_stride, void* workspace, int64_t rows,
                       int64_t row0, int T, int pass, int positive_only, void* stream) {
  ADALOG_REQUIRE(x && workspace && local_rows > 0 && n > 0 && row0 >= 0 && row0 + local_rows <= rows && pass >= 0 &&
                     pass < 4 && T > 0 && T <= sel::kMaxT, -1, "select_hist: bad arguments");
  auto* st = reinterpret_cast<sel::State*>(workspace) + row0 * T;
  auto* hist = reinterpret_cast<unsigned int*>(adalog_select_hist_ptr(workspace, rows, T)) + row0 * T * sel::kBins;
  // a few CTAs per SM over the longest rows, fewer when there are many rows
  long long bx = (n + 16383) / 16384;
  const long long cap = (long long)kNumSMs * 8 / local_rows;
  bx = bx < 1 ? 1 : bx;
  if (bx > (cap < 1 ? 1 : cap)) bx = cap < 1 ? 1 : cap;
  ADALOG_REQUIRE(local_rows <= 65535, -1, "select_hist: at most 65535 rows per call");
  dim3 grid((unsigned)bx, (unsigned)local_rows);
  sel::select_hist_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, row_stride, st, hist, T, pass, positive_only);
  return check_launch("select_hist");
}

int adalog_select_scan(void* workspace, int64_t rows, int T, int pass, void* stream) {
  ADALOG_REQUIRE(workspace && rows > 0 && T > 0 && T <= sel::kMaxT && pass >= 0 && pass < 4, -1, "select_scan: bad arguments");
  const long long n_rt = rows * T;
  auto* st = reinterpret_cast<sel::State*>(workspace);
  auto* hist = reinterpret_cast<unsigned int*>(adalog_select_hist_ptr(workspace, rows, T));
  sel::select_scan_kernel<<<(unsigned)((n_rt + 7) / 8), 256, 0, (cudaStream_t)stream>>>(st, hist, n_rt, pass);
  return check_launch("select_scan");
}

int adalog_select_finish(const void* workspace, int64_t rows, int T, float* out, void* stream) {
  ADALOG_REQUIRE(workspace && out && rows > 0 && T > 0 && T <= sel::kMaxT, -1, "select_finish: bad arguments");
  const long long n_rt = rows * T;
  sel::select_finish_kernel<<<(unsigned)((n_rt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const sel::State*>(workspace), out, n_rt);
  return check_launch("select_finish");
}

}  // extern "C"

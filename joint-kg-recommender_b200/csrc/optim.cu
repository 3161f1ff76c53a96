// Sparse-row optimizer: the first "next" row of SURVEY 8(f).  Replaces the reference's dense
// torch.optim step + clip_grad_norm (utils/trainer.py:63-81, knowledge_representation.py:213),
// whose cost is O(table) per step, by kernels whose cost is O(rows touched by the batch) plus one
// 4-byte flag per table row.
//
// Layout.  Gradients arrive accumulated per row in a persistent dense accumulator `acc` (the training
// kernels' "dense" gradient mode; it is all-zero outside a step).  Which rows a step touched is recorded
// as an EPOCH MARK: marks[row] = step number, written by k_rows_mark from the batch's own id arrays (plain
// idempotent stores -- duplicates cost nothing, nothing is ever cleared, no atomics).  Two sweeps then
// visit the tables, ALL of them in one launch each:
//   k_rows_sqnorm   sum of |acc[row]|^2 over marked rows            (clip_grad_norm's total norm)
//   k_rows_update   clip scale, weight decay, SGD / Adagrad / Adam on the marked rows, acc row := 0
// A warp reads 32 marks with one coalesced load, ballots, and walks the set bits; a marked row is
// processed with 128-bit loads / stores (lane = 16-byte chunk).  Per step the cost is
// rows * 4 B of marks + touched rows * (2..4 reads + 2..4 writes) of d floats -- at configs[1]
// (100k entities, every row touched) ~0.25 GB, against ~1.3 GB for the id-list + CAS version it replaces.
#include "common.cuh"

namespace kgrec {

enum { OPT_SGD = 0, OPT_ADAGRAD = 1, OPT_ADAM = 2 };
constexpr int kMaxOptTables = 8;
constexpr int kMaxMarkSegs = 8;

struct MarkArgs {
  kgrec_mark_seg seg[kMaxMarkSegs];
  int64_t begin[kMaxMarkSegs + 1];      // prefix sums of seg[].n
  int n_segs;
  int32_t epoch;
};

// marks[id] = epoch for every id of every segment.  compact: the group-compact corrupted-id format
// (v < 0 names entity ~v).  remap: ids are looked up first (KTUP: item -> aligned entity row).
__global__ void __launch_bounds__(256) k_rows_mark(const MarkArgs A, int32_t* status) {
  const int64_t total = A.begin[A.n_segs];
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int s = 0;
#pragma unroll
    for (int k = 1; k < kMaxMarkSegs; ++k) s += (k < A.n_segs && i >= A.begin[k]) ? 1 : 0;
    const kgrec_mark_seg& S = A.seg[s];
    int64_t v = load_idx(S.ids, i - A.begin[s], S.idx_bytes == 8);
    if (S.compact && v < 0) v = ~v;
    // out-of-range ids: the training kernels clamp them to row 0 (and raise the status word), so that is where
    // their gradient went -- mark the same row, or its accumulator would never be cleared
    if (S.remap) {
      if (static_cast<uint64_t>(v) >= static_cast<uint64_t>(S.n_remap)) { if (status) *status = 1; v = 0; }
      v = __ldg(S.remap + v);
    }
    if (static_cast<uint64_t>(v) >= static_cast<uint64_t>(S.rows)) { if (status) *status = 1; v = 0; }
    S.marks[v] = A.epoch;
  }
}

struct SweepArgs {
  kgrec_opt_table tab[kMaxOptTables];
  int64_t chunk_begin[kMaxOptTables + 1];    // prefix sums of ceil(rows / 32)
  int32_t div[kMaxOptTables];                // sweep rows per table row (wide rows are swept in segments)
  int n_tabs;
  int32_t epoch;
  int kind;
  float lr, eps, beta1, beta2, wd, bias1, bias2_sqrt;
  const float* sqnorm;
  float max_norm;
};

// A warp owns 32 consecutive rows of one table: one coalesced load of their marks, a ballot, and the marked rows are
// handed out kRowsInFlight at a time -- f gets the batch so that it can issue the loads of all its rows before the
// first use (the sweeps are latency-bound: a row is three dependent-free loads, a few flops, three stores).
constexpr int kRowsInFlight = 4;

template <typename F>
__device__ __forceinline__ void sweep_rows(const SweepArgs& A, F&& f) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t total = A.chunk_begin[A.n_tabs];
  for (int64_t c = warp; c < total; c += n_warps) {
    int t = 0;
#pragma unroll
    for (int k = 1; k < kMaxOptTables; ++k) t += (k < A.n_tabs && c >= A.chunk_begin[k]) ? 1 : 0;
    const kgrec_opt_table& T = A.tab[t];
    const int64_t row0 = (c - A.chunk_begin[t]) * 32;
    const int64_t row = row0 + lane;
    bool mine = row < T.rows;
    if (mine && T.marks) mine = __ldg(T.marks + row / A.div[t]) == A.epoch;
    unsigned m = __ballot_sync(FULL, mine);
    while (m) {
      int64_t rows[kRowsInFlight];
      int n = 0;
#pragma unroll
      for (int i = 0; i < kRowsInFlight; ++i) {
        rows[i] = row0;
        if (m) { rows[i] = row0 + (__ffs(m) - 1); m &= m - 1; n = i + 1; }
      }
      f(T, rows, n, lane);
    }
  }
}

__global__ void __launch_bounds__(256) k_rows_sqnorm(const SweepArgs A, float* out) {
  float local = 0.f;
  sweep_rows(A, [&](const kgrec_opt_table& T, const int64_t (&rows)[kRowsInFlight], int n, int lane) {
    const int nch = (T.dim + 3) >> 2;
    for (int ch = lane; ch < nch; ch += 32) {
      if (T.vec) {
        float4 v[kRowsInFlight];
#pragma unroll
        for (int i = 0; i < kRowsInFlight; ++i)
          v[i] = i < n ? *reinterpret_cast<const float4*>(T.acc + rows[i] * T.dim + ch * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < kRowsInFlight; ++i)
          local = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, local))));
      } else {
        for (int i = 0; i < n; ++i)
          for (int e = 0; e < 4 && ch * 4 + e < T.dim; ++e) { const float g = T.acc[rows[i] * T.dim + ch * 4 + e]; local = fmaf(g, g, local); }
      }
    }
  });
  local = warp_sum(local);
  __shared__ float part[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) part[wid] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w];
    if (t != 0.f) atomicAdd(out, t);
  }
}

__device__ __forceinline__ float opt_elem(const SweepArgs& A, float pv, float gv, float* s1, float* s2) {
  if (A.wd != 0.f) gv = fmaf(A.wd, pv, gv);          // weight_decay = l2_lambda, on touched rows only
  if (A.kind == OPT_SGD) return pv - A.lr * gv;
  if (A.kind == OPT_ADAGRAD) {                        // torch.optim.Adagrad: sum += g^2 ; p -= lr g / (sqrt(sum) + eps)
    const float s = fmaf(gv, gv, *s1);
    *s1 = s;
    return pv - A.lr * gv / (sqrtf(s) + A.eps);
  }
  const float m = A.beta1 * *s1 + (1.f - A.beta1) * gv;          // torch.optim.Adam on the touched rows ("lazy")
  const float v = A.beta2 * *s2 + (1.f - A.beta2) * gv * gv;
  *s1 = m;
  *s2 = v;
  return pv - (A.lr / A.bias1) * m / (sqrtf(v) / A.bias2_sqrt + A.eps);
}

__global__ void __launch_bounds__(256) k_rows_update(const SweepArgs A) {
  float scale = 1.f;
  if (A.sqnorm) scale = fminf(1.f, A.max_norm / (sqrtf(__ldg(A.sqnorm)) + 1e-6f));   // clip_grad_norm's coefficient
  sweep_rows(A, [&](const kgrec_opt_table& T, const int64_t (&rows)[kRowsInFlight], int n, int lane) {
    const int nch = (T.dim + 3) >> 2;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ch = lane; ch < nch; ch += 32) {
      if (T.vec) {
        float4 g[kRowsInFlight], p[kRowsInFlight], a[kRowsInFlight], b[kRowsInFlight];
#pragma unroll
        for (int i = 0; i < kRowsInFlight; ++i) {            // every load of the batch leaves before the first use
          const int64_t o = rows[i] * T.dim + ch * 4;
          g[i] = p[i] = a[i] = b[i] = z4;
          if (i < n) {
            g[i] = *reinterpret_cast<const float4*>(T.acc + o);
            p[i] = *reinterpret_cast<const float4*>(T.table + o);
            if (A.kind != OPT_SGD) a[i] = *reinterpret_cast<const float4*>(T.state1 + o);
            if (A.kind == OPT_ADAM) b[i] = *reinterpret_cast<const float4*>(T.state2 + o);
          }
        }
#pragma unroll
        for (int i = 0; i < kRowsInFlight; ++i) {
          if (i >= n) continue;
          const int64_t o = rows[i] * T.dim + ch * 4;
          p[i].x = opt_elem(A, p[i].x, g[i].x * scale, &a[i].x, &b[i].x);
          p[i].y = opt_elem(A, p[i].y, g[i].y * scale, &a[i].y, &b[i].y);
          p[i].z = opt_elem(A, p[i].z, g[i].z * scale, &a[i].z, &b[i].z);
          p[i].w = opt_elem(A, p[i].w, g[i].w * scale, &a[i].w, &b[i].w);
          if (!T.keep_acc) *reinterpret_cast<float4*>(T.acc + o) = z4;                   // zero again after the step
          *reinterpret_cast<float4*>(T.table + o) = p[i];
          if (A.kind != OPT_SGD) *reinterpret_cast<float4*>(T.state1 + o) = a[i];
          if (A.kind == OPT_ADAM) *reinterpret_cast<float4*>(T.state2 + o) = b[i];
        }
      } else {
        for (int i = 0; i < n; ++i)
          for (int e = 0; e < 4 && ch * 4 + e < T.dim; ++e) {
            const int64_t o = rows[i] * T.dim + ch * 4 + e;
            const float g = T.acc[o];
            if (!T.keep_acc) T.acc[o] = 0.f;
            float a = A.kind != OPT_SGD ? T.state1[o] : 0.f, b = A.kind == OPT_ADAM ? T.state2[o] : 0.f;
            T.table[o] = opt_elem(A, T.table[o], g * scale, &a, &b);
            if (A.kind != OPT_SGD) T.state1[o] = a;
            if (A.kind == OPT_ADAM) T.state2[o] = b;
          }
      }
    }
  });
}

}  // namespace kgrec

using namespace kgrec;

static int sweep_args(const kgrec_opt_table* tabs, int n_tabs, int32_t epoch, int need_table, SweepArgs& A) {
  if (!tabs || n_tabs < 1 || n_tabs > kMaxOptTables) {
    set_error("sparse row optimizer: 1..%d tables per call", kMaxOptTables);
    return KGREC_ERR_INVALID;
  }
  A = SweepArgs{};
  A.n_tabs = n_tabs;
  A.epoch = epoch;
  for (int t = 0; t < n_tabs; ++t) {
    kgrec_opt_table T = tabs[t];
    if (!T.acc || T.rows <= 0 || T.dim <= 0 || (need_table && !T.table)) {
      set_error("sparse row optimizer: table %d has no accumulator / parameter / shape", t);
      return KGREC_ERR_INVALID;
    }
    T.vec = (T.dim % 4 == 0) && (reinterpret_cast<uintptr_t>(T.acc) % 16 == 0) &&
            (!T.table || reinterpret_cast<uintptr_t>(T.table) % 16 == 0) &&
            (!T.state1 || reinterpret_cast<uintptr_t>(T.state1) % 16 == 0) &&
            (!T.state2 || reinterpret_cast<uintptr_t>(T.state2) % 16 == 0);
    // wide rows (TransR's d x d matrices: 10^4 floats per row, a few hundred rows) are swept as rows of a segment each, or
    // a handful of warps would walk the whole table
    A.div[t] = 1;
    if (T.dim > 512)
      for (int seg = 512; seg >= 64; seg -= 4)
        if (T.dim % seg == 0) { A.div[t] = T.dim / seg; T.rows *= A.div[t]; T.dim = seg; break; }
    A.tab[t] = T;
    A.chunk_begin[t + 1] = A.chunk_begin[t] + (T.rows + 31) / 32;
  }
  for (int t = n_tabs; t < kMaxOptTables; ++t) A.chunk_begin[t + 1] = A.chunk_begin[n_tabs];
  return KGREC_OK;
}

static int sweep_grid(const SweepArgs& A) {
  const int64_t warps = A.chunk_begin[A.n_tabs], ctas = (warps + 7) / 8, cap = static_cast<int64_t>(sm_count()) * 8;
  return static_cast<int>(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}

extern "C" int kgrec_rows_mark(const kgrec_mark_seg* segs, int n_segs, int32_t epoch, int32_t* status,
                               kgrec_stream_t stream) {
  if (!segs || n_segs < 1 || n_segs > kMaxMarkSegs) {
    set_error("kgrec_rows_mark: 1..%d id segments per call", kMaxMarkSegs);
    return KGREC_ERR_INVALID;
  }
  MarkArgs A{};
  A.n_segs = n_segs;
  A.epoch = epoch;
  for (int s = 0; s < n_segs; ++s) {
    const kgrec_mark_seg& S = segs[s];
    if (S.n < 0 || (S.n > 0 && (!S.ids || !S.marks)) || (S.idx_bytes != 4 && S.idx_bytes != 8) || S.rows <= 0) {
      set_error("kgrec_rows_mark: bad segment %d", s);
      return KGREC_ERR_INVALID;
    }
    A.seg[s] = S;
    A.begin[s + 1] = A.begin[s] + S.n;
  }
  for (int s = n_segs; s < kMaxMarkSegs; ++s) A.begin[s + 1] = A.begin[n_segs];
  const int64_t total = A.begin[n_segs];
  if (total == 0) return KGREC_OK;
  const int64_t blocks = (total + 255) / 256, cap = static_cast<int64_t>(sm_count()) * 16;
  k_rows_mark<<<static_cast<int>(blocks < cap ? blocks : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(A, status);
  KGREC_CUDA_OK(cudaGetLastError());
  return KGREC_OK;
}

extern "C" int kgrec_rows_sqnorm(const kgrec_opt_table* tabs, int n_tabs, int32_t epoch, float* sqnorm,
                                 kgrec_stream_t stream) {
  SweepArgs A;
  int rc = sweep_args(tabs, n_tabs, epoch, 0, A);
  if (rc) return rc;
  if (!sqnorm) { set_error("sqnorm is NULL"); return KGREC_ERR_INVALID; }
  k_rows_sqnorm<<<sweep_grid(A), 256, 0, static_cast<cudaStream_t>(stream)>>>(A, sqnorm);
  KGREC_CUDA_OK(cudaGetLastError());
  return KGREC_OK;
}

extern "C" int kgrec_rows_update(const kgrec_opt_table* tabs, int n_tabs, int32_t epoch, int kind, float lr, float eps,
                                 float beta1, float beta2, int64_t step, float weight_decay, const float* sqnorm,
                                 float max_norm, kgrec_stream_t stream) {
  SweepArgs A;
  int rc = sweep_args(tabs, n_tabs, epoch, 1, A);
  if (rc) return rc;
  if (kind < OPT_SGD || kind > OPT_ADAM) { set_error("sparse row optimizer: unknown kind %d", kind); return KGREC_ERR_INVALID; }
  for (int t = 0; t < n_tabs; ++t)
    if ((kind != OPT_SGD && !tabs[t].state1) || (kind == OPT_ADAM && !tabs[t].state2)) {
      set_error("sparse row optimizer: state missing for optimizer kind %d (table %d)", kind, t);
      return KGREC_ERR_INVALID;
    }
  A.kind = kind; A.lr = lr; A.eps = eps; A.beta1 = beta1; A.beta2 = beta2; A.wd = weight_decay;
  A.bias1 = 1.f - powf(beta1, static_cast<float>(step));
  A.bias2_sqrt = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
  A.sqnorm = sqnorm; A.max_norm = max_norm;
  k_rows_update<<<sweep_grid(A), 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  KGREC_CUDA_OK(cudaGetLastError());
  return KGREC_OK;
}

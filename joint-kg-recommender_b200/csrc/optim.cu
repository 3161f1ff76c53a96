// Sparse-row optimizer: the first "next" row of SURVEY 8(f).  Replaces the reference's dense
// torch.optim step + clip_grad_norm (utils/trainer.py:63-81, knowledge_representation.py:213),
// whose cost is O(table) per step, by kernels whose cost is O(rows touched by the batch).
//
// Gradients arrive accumulated per row in a persistent dense accumulator `acc` (the kernels'
// "dense" gradient mode; it is all-zero outside a step), together with the id list of the
// batch (duplicates allowed).  A row is owned by the first id-list entry that claims it
// (flags: 0 free -> 1 claimed for the norm -> 2 claimed for the update); the owner applies the
// update and clears the accumulator row, a last pass frees the flags.
#include "common.cuh"

namespace kgrec {

enum { OPT_SGD = 0, OPT_ADAGRAD = 1, OPT_ADAM = 2 };

struct OptArgs {
  float* table; float* acc; float* s1; float* s2;   // s1: Adagrad sum / Adam m ; s2: Adam v
  int32_t* flags;
  const void* idx; int is64; int64_t n; int64_t rows; int d;
  int kind; float lr, eps, beta1, beta2, wd, bias1, bias2;
  const float* sqnorm; float max_norm;               // optional clip: scale = min(1, max_norm / (sqrt(*sqnorm) + 1e-6))
};

__device__ __forceinline__ int64_t opt_row(const OptArgs& A, int64_t i) {
  const int64_t r = load_idx(A.idx, i, A.is64);
  return (static_cast<uint64_t>(r) < static_cast<uint64_t>(A.rows)) ? r : -1;
}

// sum over distinct touched rows of |acc[row]|^2  (+= into *out)
__global__ void __launch_bounds__(kThreads) k_rows_sqnorm(const OptArgs A, float* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float local = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + wid; i < A.n; i += static_cast<int64_t>(gridDim.x) * kWarpsPerCta) {
    const int64_t r = opt_row(A, i);
    if (r < 0) continue;
    int own = 0;
    if (lane == 0) own = atomicCAS(A.flags + r, 0, 1) == 0;
    own = __shfl_sync(FULL, own, 0);
    if (!own) continue;
    const float* g = A.acc + r * A.d;
    for (int j = lane; j < A.d; j += 32) { const float v = g[j]; local = fmaf(v, v, local); }
  }
  local = warp_sum(local);
  __shared__ float part[kWarpsPerCta];
  if (lane == 0) part[wid] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kWarpsPerCta; ++w) t += part[w];
    if (t != 0.f) atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(kThreads) k_rows_step(const OptArgs A, const int claimed_from) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float scale = 1.f;
  if (A.sqnorm) scale = fminf(1.f, A.max_norm / (sqrtf(__ldg(A.sqnorm)) + 1e-6f));   // clip_grad_norm's coefficient
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + wid; i < A.n; i += static_cast<int64_t>(gridDim.x) * kWarpsPerCta) {
    const int64_t r = opt_row(A, i);
    if (r < 0) continue;
    int own = 0;
    if (lane == 0) own = atomicCAS(A.flags + r, claimed_from, 2) == claimed_from;
    own = __shfl_sync(FULL, own, 0);
    if (!own) continue;
    float* p = A.table + r * A.d;
    float* g = A.acc + r * A.d;
    for (int j = lane; j < A.d; j += 32) {
      float gv = g[j] * scale;
      g[j] = 0.f;                                     // the accumulator is zero again after the step
      float pv = p[j];
      if (A.wd != 0.f) gv = fmaf(A.wd, pv, gv);       // weight_decay = l2_lambda, on touched rows only
      if (A.kind == OPT_SGD) {
        pv -= A.lr * gv;
      } else if (A.kind == OPT_ADAGRAD) {             // torch.optim.Adagrad: sum += g^2 ; p -= lr g / (sqrt(sum) + eps)
        const float s = fmaf(gv, gv, A.s1[r * A.d + j]);
        A.s1[r * A.d + j] = s;
        pv -= A.lr * gv / (sqrtf(s) + A.eps);
      } else {                                        // torch.optim.Adam on the touched rows ("lazy")
        const float m = A.beta1 * A.s1[r * A.d + j] + (1.f - A.beta1) * gv;
        const float v = A.beta2 * A.s2[r * A.d + j] + (1.f - A.beta2) * gv * gv;
        A.s1[r * A.d + j] = m;
        A.s2[r * A.d + j] = v;
        pv -= (A.lr / A.bias1) * m / (sqrtf(v) / sqrtf(A.bias2) + A.eps);
      }
      p[j] = pv;
    }
  }
}

__global__ void __launch_bounds__(256) k_rows_release(const OptArgs A) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < A.n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = opt_row(A, i);
    if (r >= 0) A.flags[r] = 0;
  }
}

}  // namespace kgrec

using namespace kgrec;

static int opt_check(const float* acc, const int32_t* flags, const void* idx, int idx_bytes, int64_t n, int64_t rows, int d) {
  if (!acc || !flags || !idx || (idx_bytes != 4 && idx_bytes != 8) || n < 0 || rows <= 0 || d <= 0) {
    set_error("sparse row optimizer: bad arguments");
    return KGREC_ERR_INVALID;
  }
  return KGREC_OK;
}

static int opt_grid(int64_t n) {
  const int64_t ctas = (n + kWarpsPerCta - 1) / kWarpsPerCta, cap = static_cast<int64_t>(sm_count()) * 8;
  return static_cast<int>(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}

extern "C" int kgrec_rows_sqnorm(const float* acc, int32_t* flags, const void* idx, int idx_bytes, int64_t n,
                                 int64_t rows, int32_t dim, float* sqnorm, kgrec_stream_t stream) {
  int rc = opt_check(acc, flags, idx, idx_bytes, n, rows, dim);
  if (rc) return rc;
  if (!sqnorm) { set_error("sqnorm is NULL"); return KGREC_ERR_INVALID; }
  if (n == 0) return KGREC_OK;
  OptArgs A{};
  A.acc = const_cast<float*>(acc); A.flags = flags; A.idx = idx; A.is64 = idx_bytes == 8; A.n = n; A.rows = rows; A.d = dim;
  k_rows_sqnorm<<<opt_grid(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(A, sqnorm);
  KGREC_CUDA_OK(cudaGetLastError());
  return KGREC_OK;
}

extern "C" int kgrec_rows_step(float* table, float* acc, float* state1, float* state2, int32_t* flags, const void* idx,
                               int idx_bytes, int64_t n, int64_t rows, int32_t dim, int kind, float lr, float eps,
                               float beta1, float beta2, int64_t step, float weight_decay, const float* sqnorm,
                               float max_norm, int norm_claimed, kgrec_stream_t stream) {
  int rc = opt_check(acc, flags, idx, idx_bytes, n, rows, dim);
  if (rc) return rc;
  if (!table || kind < OPT_SGD || kind > OPT_ADAM || (kind != OPT_SGD && !state1) || (kind == OPT_ADAM && !state2)) {
    set_error("sparse row optimizer: table / state missing for optimizer kind %d", kind);
    return KGREC_ERR_INVALID;
  }
  if (n == 0) return KGREC_OK;
  OptArgs A{};
  A.table = table; A.acc = acc; A.s1 = state1; A.s2 = state2; A.flags = flags; A.idx = idx; A.is64 = idx_bytes == 8;
  A.n = n; A.rows = rows; A.d = dim; A.kind = kind; A.lr = lr; A.eps = eps; A.beta1 = beta1; A.beta2 = beta2; A.wd = weight_decay;
  A.bias1 = 1.f - powf(beta1, static_cast<float>(step));
  A.bias2 = 1.f - powf(beta2, static_cast<float>(step));
  A.sqnorm = sqnorm; A.max_norm = max_norm;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_rows_step<<<opt_grid(n), kThreads, 0, st>>>(A, norm_claimed ? 1 : 0);
  KGREC_CUDA_OK(cudaGetLastError());
  const int64_t blocks = (n + 255) / 256, cap = static_cast<int64_t>(sm_count()) * 8;
  k_rows_release<<<static_cast<int>(blocks < cap ? blocks : cap), 256, 0, st>>>(A);
  KGREC_CUDA_OK(cudaGetLastError());
  return KGREC_OK;
}

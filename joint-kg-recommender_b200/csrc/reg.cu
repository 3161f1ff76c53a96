// The drivers' regularisers that are not tied to a scoring kernel's rows: the recommendation side of
// item_recommendation.py:177-180 and knowledgable_recommendation.py:343-344,
//   normLoss(user rows of the batch) + normLoss(item rows of cat[pos, neg]) + normLoss(pref table)
//   + orthogonalLoss(pref table, pref_norm table)
// (utils/loss.py:18-23).  Values are added to a device scalar, gradients to the persistent dense
// accumulators the sparse-row optimizer consumes (csrc/optim.cu).  The KG side's normLoss / orthogonalLoss
// over the gathered triples' rows is fused into the step kernels (train_group.cu, REG).
#include "common.cuh"

namespace kgrec {

// normLoss over table[ids[i]] (ids == NULL: rows 0..n-1): loss += scale * max(|row|^2 - 1, 0),
// acc[row] += scale * 2 row where |row|^2 > 1 (each LISTED occurrence counts, as the reference's gather does).
__global__ void __launch_bounds__(kThreads)
k_reg_norm_rows(const float* __restrict__ table, int64_t rows, int d, const void* ids, int is64, int64_t n,
                float scale, float* loss_out, float* acc, int32_t* status) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(acc)) % 16 == 0);
  float lsum = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + wid; i < n; i += static_cast<int64_t>(gridDim.x) * kWarpsPerCta) {
    int64_t r = ids ? load_idx(ids, i, is64) : i;
    if (static_cast<uint64_t>(r) >= static_cast<uint64_t>(rows)) { if (status) *status = 1; continue; }
    const float* x = table + r * d;
    float n2 = 0.f;
    if (vec) {
      for (int c = lane; c * 4 < d; c += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + c);
        n2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, n2))));
      }
    } else {
      for (int j = lane; j < d; j += 32) n2 = fmaf(x[j], x[j], n2);
    }
    n2 = warp_sum(n2);
    if (n2 > 1.f) {
      lsum += n2 - 1.f;
      if (acc) {
        const float s2 = 2.f * scale;
        float* g = acc + r * d;
        if (vec) {
          for (int c = lane; c * 4 < d; c += 32) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x) + c);
            red_add_f4(g + 4 * c, s2 * v.x, s2 * v.y, s2 * v.z, s2 * v.w);
          }
        } else {
          for (int j = lane; j < d; j += 32) atomicAdd(g + j, s2 * x[j]);
        }
      }
    }
  }
  __shared__ float part[kWarpsPerCta];
  if (lane == 0) part[wid] = lsum;
  __syncthreads();
  if (threadIdx.x == 0 && loss_out) {
    float t = 0.f;
    for (int w = 0; w < kWarpsPerCta; ++w) t += part[w];
    if (t != 0.f) atomicAdd(loss_out, scale * t);
  }
}

// orthogonalLoss(rel, norm) = sum_rows (w.r)^2 / |r|^2 over two whole [rows, d] tables;
// d/dr = 2 q w - 2 q^2 r, d/dw = 2 q r with q = (w.r) / |r|^2.
__global__ void __launch_bounds__(kThreads)
k_reg_orth(const float* __restrict__ rel, const float* __restrict__ nrm, int64_t rows, int d, float scale,
           float* loss_out, float* acc_rel, float* acc_nrm) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float lsum = 0.f;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + wid; r < rows; r += static_cast<int64_t>(gridDim.x) * kWarpsPerCta) {
    const float* x = rel + r * d;
    const float* w = nrm + r * d;
    float wr = 0.f, n2 = 0.f;
    for (int j = lane; j < d; j += 32) { wr = fmaf(w[j], x[j], wr); n2 = fmaf(x[j], x[j], n2); }
    warp_sum2(wr, n2);
    lsum += wr * wr / n2;
    const float q = wr / n2;
    for (int j = lane; j < d; j += 32) {
      if (acc_rel) atomicAdd(acc_rel + r * d + j, scale * (2.f * q * w[j] - 2.f * q * q * x[j]));
      if (acc_nrm) atomicAdd(acc_nrm + r * d + j, scale * 2.f * q * x[j]);
    }
  }
  __shared__ float part[kWarpsPerCta];
  if (lane == 0) part[wid] = lsum;
  __syncthreads();
  if (threadIdx.x == 0 && loss_out) {
    float t = 0.f;
    for (int w = 0; w < kWarpsPerCta; ++w) t += part[w];
    atomicAdd(loss_out, scale * t);
  }
}

}  // namespace kgrec

using namespace kgrec;

static int reg_grid(int64_t n) {
  const int64_t ctas = (n + kWarpsPerCta - 1) / kWarpsPerCta, cap = static_cast<int64_t>(sm_count()) * 8;
  return static_cast<int>(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}

extern "C" int kgrec_reg_norm_rows(const float* table, int64_t rows, int32_t dim, const void* ids, int idx_bytes,
                                   int64_t n, float scale, float* loss_out, float* acc, int32_t* status,
                                   kgrec_stream_t stream) {
  if (!table || rows <= 0 || dim <= 0 || n < 0 || (ids && idx_bytes != 4 && idx_bytes != 8)) {
    set_error("kgrec_reg_norm_rows: bad arguments");
    return KGREC_ERR_INVALID;
  }
  if (!ids && n > rows) { set_error("kgrec_reg_norm_rows: n > rows without an id list"); return KGREC_ERR_INVALID; }
  if (n == 0) return KGREC_OK;
  k_reg_norm_rows<<<reg_grid(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(table, rows, dim, ids, idx_bytes == 8, n,
                                                                                    scale, loss_out, acc, status);
  KGREC_CUDA_OK(cudaGetLastError());
  return KGREC_OK;
}

extern "C" int kgrec_reg_orth_tables(const float* rel, const float* norm, int64_t rows, int32_t dim, float scale,
                                     float* loss_out, float* acc_rel, float* acc_norm, kgrec_stream_t stream) {
  if (!rel || !norm || rows <= 0 || dim <= 0) { set_error("kgrec_reg_orth_tables: bad arguments"); return KGREC_ERR_INVALID; }
  k_reg_orth<<<reg_grid(rows), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(rel, norm, rows, dim, scale, loss_out, acc_rel, acc_norm);
  KGREC_CUDA_OK(cudaGetLastError());
  return KGREC_OK;
}

// Hungarian-matched instance loss of the training step (networks/evaluator.py:19-74, called twice per iteration by
// train_dmsr.py:38-45): the two [ins x ins] cost matrices from ONE pass over the N rays of the batch, and the gradient of the
// matched loss.  The assignment itself (scipy linear_sum_assignment on the [valid x ins] matrix, evaluator.py:45-47) stays on
// the host like in the reference: it is a few microseconds on a <= 94 x 94 matrix.
//
// gt is one-hot (evaluator.py:21-25: column v of gt_ins marks the rays whose label is the v-th smallest label present), so the
// dense [ins x ins x N] broadcast of the reference collapses to per-row sums.  With row(n) = index of ray n's label:
//   A[p]    = sum_n          log(1 - pred[n,p] + 1e-8)
//   B[g,p]  = sum_{row(n)=g} log(pred[n,p] + 1e-8)       C[g,p] = sum_{row(n)=g} log(1 - pred[n,p] + 1e-8)
//   TP[g,p] = sum_{row(n)=g} pred[n,p]                   S[p]   = sum_n pred[n,p]           cnt[g] = #{n : row(n) = g}
//   cost_ce[g,p]   = -(B[g,p] + A[p] - C[g,p]) / N                                           (evaluator.py:60)
//   cost_siou[g,p] = 1 - TP / (TP + (S[p] - TP) + (cnt[g] - TP) + 1e-6)                      (evaluator.py:63-67)
// fp32 terms, fp64 accumulation (the reference sums fp32 terms pairwise: both are ~1e-7 from the exact sum).
#include <cstdint>

#include "common.cuh"
#include "ray_ops.cuh"

namespace dmnerf {

constexpr int EV_MAX_K = DMNERF_MAX_INS + 1;

// One block per prediction column p.
__global__ void hungarian_cost_kernel(const float* __restrict__ pred, const int32_t* __restrict__ gt_row, int64_t n, int k,
                                      float* __restrict__ cost_ce, float* __restrict__ cost_siou, float* __restrict__ tp_out,
                                      float* __restrict__ s_out, float* __restrict__ cnt_out) {
  __shared__ double sB[EV_MAX_K], sC[EV_MAX_K], sTP[EV_MAX_K];
  __shared__ double sA, sS;
  __shared__ unsigned int sCnt[EV_MAX_K];
  const int p = blockIdx.x;
  for (int g = threadIdx.x; g < k; g += blockDim.x) { sB[g] = 0.0; sC[g] = 0.0; sTP[g] = 0.0; sCnt[g] = 0u; }
  if (threadIdx.x == 0) { sA = 0.0; sS = 0.0; }
  __syncthreads();
  double a = 0.0, s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = pred[i * k + p];
    const int g = gt_row[i];
    const float l1 = logf(__fadd_rn(__fsub_rn(1.0f, v), 1e-8f));
    a += (double)l1;
    s += (double)v;
    if (g >= 0 && g < k) {
      atomicAdd(&sB[g], (double)logf(__fadd_rn(v, 1e-8f)));
      atomicAdd(&sC[g], (double)l1);
      atomicAdd(&sTP[g], (double)v);
      atomicAdd(&sCnt[g], 1u);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(FULL, a, d); s += __shfl_xor_sync(FULL, s, d); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sA, a); atomicAdd(&sS, s); }
  __syncthreads();
  for (int g = threadIdx.x; g < k; g += blockDim.x) {
    const double tp = sTP[g], cnt = (double)sCnt[g];
    cost_ce[(size_t)g * k + p] = (float)(-(sB[g] + sA - sC[g]) / (double)n);
    // the reference evaluates TP, FP = sum(pred) - TP, FN = sum(gt) - TP in fp32 and then TP / (TP + FP + FN + 1e-6)
    const float tpf = (float)tp, fp = __fsub_rn((float)sS, tpf), fn = __fsub_rn((float)cnt, tpf);
    const float den = __fadd_rn(__fadd_rn(__fadd_rn(tpf, fp), fn), 1e-6f);
    cost_siou[(size_t)g * k + p] = __fsub_rn(1.0f, __fdiv_rn(tpf, den));
    tp_out[(size_t)g * k + p] = tpf;
    if (p == 0) cnt_out[g] = (float)cnt;
  }
  if (threadIdx.x == 0) s_out[p] = (float)sS;
}

// d loss / d pred for  loss = g_ce * valid_ce + g_inv * invalid_ce + g_siou * valid_siou  (evaluator.py:27-36):
//   valid_ce = mean_{g < V} cost_ce[g, col(g)],  valid_siou likewise,  invalid_ce = mean(pred[:, unmatched columns]).
// row_of_col[p] = matched gt row of prediction column p, or -1 (unmatched).  g3 = the three upstream gradients (device).
__global__ void ins_loss_grad_kernel(const float* __restrict__ pred, const int32_t* __restrict__ gt_row, int64_t n, int k,
                                     const int32_t* __restrict__ row_of_col, int n_valid, const float* __restrict__ tp,
                                     const float* __restrict__ s_sum, const float* __restrict__ cnt,
                                     const float* __restrict__ g3, float* __restrict__ d_pred) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * k) return;
  const int64_t i = idx / k;
  const int p = (int)(idx % k);
  const int g = row_of_col[p];
  const float v = pred[idx];
  float d;
  if (g >= 0) {
    const bool on = gt_row[i] == g;
    const float inv_v = 1.0f / (float)n_valid;
    const float dce = on ? -1.0f / (v + 1e-8f) : 1.0f / ((1.0f - v) + 1e-8f);
    const float t = tp[(size_t)g * k + p];
    const float den = (s_sum[p] + cnt[g] - t) + 1e-6f;
    const float dsi = on ? -1.0f / den : t / (den * den);      // -(gt D - TP (1 - gt)) / D^2
    d = g3[0] * inv_v * dce / (float)n + g3[2] * inv_v * dsi;
  } else {
    d = g3[1] / ((float)n * (float)(k - n_valid));
  }
  d_pred[idx] = d;
}

int launch_hungarian_costs(const float* pred, const int32_t* gt_row, int64_t n, int k, float* cost_ce, float* cost_siou,
                           float* tp, float* s_sum, float* cnt, cudaStream_t st) {
  DMN_CHECK(k >= 1 && k <= EV_MAX_K, "hungarian_costs: ins_num %d out of range (max %d)", k, EV_MAX_K);
  DMN_CHECK(n >= 1, "hungarian_costs: empty batch");
  hungarian_cost_kernel<<<(unsigned)k, 256, 0, st>>>(pred, gt_row, n, k, cost_ce, cost_siou, tp, s_sum, cnt);
  DMN_LAUNCH_OK();
  return 0;
}

int launch_ins_loss_grad(const float* pred, const int32_t* gt_row, int64_t n, int k, const int32_t* row_of_col, int n_valid,
                         const float* tp, const float* s_sum, const float* cnt, const float* g3, float* d_pred, cudaStream_t st) {
  DMN_CHECK(k >= 1 && k <= EV_MAX_K && n_valid >= 1 && n_valid <= k, "ins_loss_grad: bad sizes (k %d, valid %d)", k, n_valid);
  const int64_t total = n * k;
  if (total == 0) return 0;
  ins_loss_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pred, gt_row, n, k, row_of_col, n_valid, tp, s_sum, cnt,
                                                                        g3, d_pred);
  DMN_LAUNCH_OK();
  return 0;
}

}  // namespace dmnerf

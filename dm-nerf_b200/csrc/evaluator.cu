// Hungarian-matched instance loss of the training step (networks/evaluator.py:19-74, called twice per iteration by
// train_dmsr.py:38-45): the two [ins x ins] cost matrices from ONE pass over the N rays of the batch, and the gradient of the
// matched loss.  The assignment itself (scipy linear_sum_assignment on the [valid x ins] matrix, evaluator.py:45-47) either stays
// on the host like in the reference (dmnerf_hungarian_costs + scipy: one device->host hop per call) or runs on the device
// (label_rows_kernel + hungarian_assign_kernel: the same shortest-augmenting-path algorithm with the same tie rule, no hop at
// all -- the training iteration stays asynchronous end to end).
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
                                     const int32_t* __restrict__ row_of_col, int n_valid, const int32_t* __restrict__ n_valid_dev,
                                     const float* __restrict__ tp, const float* __restrict__ s_sum, const float* __restrict__ cnt,
                                     const float* __restrict__ g3, float* __restrict__ d_pred) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * k) return;
  if (n_valid_dev) {                         // device-side assignment: the number of distinct labels never left the device
    n_valid = *n_valid_dev;
    if (n_valid < 1) { d_pred[idx] = 0.0f; return; }        // rejected labels: NaN loss, zero gradient, error on the next call
  }
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

// ---------------------------------------------------------------------------------------------------- device-side assignment
// Row of every ray = rank of its label among the distinct labels of the batch (ascending: torch.unique order, evaluator.py:21-25),
// from a presence bitmap over label values [0, 65536); n_valid = number of distinct labels.  Labels outside that range or more
// distinct labels than prediction channels: n_valid = -1 (the loss comes out NaN, its gradient zero) and the code goes to the
// status word (mapped host memory: the next ins_criterion call raises without any synchronisation).
constexpr int LBL_WORDS = 2048;            // 65536 label values
__global__ void __launch_bounds__(1024) label_rows_kernel(const int32_t* __restrict__ labels, int64_t n, int k,
                                                          int32_t* __restrict__ gt_row, int32_t* __restrict__ n_valid_out,
                                                          int32_t* status) {
  __shared__ uint32_t bitmap[LBL_WORDS], prefix[LBL_WORDS];
  __shared__ uint32_t warp_tot[32];
  __shared__ int bad_s;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  bitmap[2 * t] = 0u; bitmap[2 * t + 1] = 0u;
  if (t == 0) bad_s = 0;
  __syncthreads();
  for (int64_t i = t; i < n; i += 1024) {
    const int l = labels[i];
    if (l < 0 || l >= LBL_WORDS * 32) bad_s = 1;
    else atomicOr(&bitmap[l >> 5], 1u << (l & 31));
  }
  __syncthreads();
  const uint32_t c0 = __popc(bitmap[2 * t]), c1 = __popc(bitmap[2 * t + 1]);
  uint32_t incl = c0 + c1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_tot[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL, w, d); if (lane >= d) w += o; }
    warp_tot[lane] = w;                                      // inclusive totals of the warps
  }
  __syncthreads();
  const uint32_t excl = incl - (c0 + c1) + (warp ? warp_tot[warp - 1] : 0u);
  prefix[2 * t] = excl; prefix[2 * t + 1] = excl + c0;
  __syncthreads();
  const int n_valid = (int)warp_tot[31];
  const bool bad = bad_s != 0 || n_valid > k || n_valid < 1;
  for (int64_t i = t; i < n; i += 1024) {
    const int l = labels[i];
    int row = -1;
    if (!bad && l >= 0 && l < LBL_WORDS * 32) row = (int)(prefix[l >> 5] + __popc(bitmap[l >> 5] & ((1u << (l & 31)) - 1u)));
    gt_row[i] = row;
  }
  if (t == 0) {
    *n_valid_out = bad ? -1 : n_valid;
    if (bad && status) atomicCAS(status, 0, bad_s ? 701 : 702);
  }
}

// scipy.optimize.linear_sum_assignment (the rectangular shortest-augmenting-path solver of Crouse 2016 that scipy implements) on
// rows 0..V-1 of cost_ce + cost_siou, by ONE warp: fp64 duals like scipy, the same evaluation order of every sum, and the same
// choice among equal path costs (scipy scans the remaining columns in order and lets a later column replace an equal earlier one
// only if it is unassigned: the winner is the LAST unassigned column among the minima if there is one, else the FIRST minimum --
// encoded below as a unique integer score per scan position, so the lane-parallel scan picks exactly scipy's column; checked
// against scipy on tie-heavy matrices in tests/test_gpu_train.py).  Then the matched loss terms (evaluator.py:27-36):
// loss3 = { mean cost_ce[g, col(g)], mean pred[:, unmatched columns] (0 if none), mean cost_siou[g, col(g)] }.
__global__ void __launch_bounds__(32) hungarian_assign_kernel(const float* __restrict__ cost_ce, const float* __restrict__ cost_siou,
                                                              const float* __restrict__ s_sum, const int32_t* __restrict__ n_valid_dev,
                                                              int64_t n, int k, int32_t* __restrict__ row_of_col,
                                                              float* __restrict__ loss3) {
  __shared__ double u[EV_MAX_K], v[EV_MAX_K], sp[EV_MAX_K];
  __shared__ int path[EV_MAX_K], col4row[EV_MAX_K], row4col[EV_MAX_K], remaining[EV_MAX_K];
  __shared__ unsigned char SR[EV_MAX_K], SC[EV_MAX_K];
  const int lane = threadIdx.x;
  const int V = *n_valid_dev;
  const float qnan = __int_as_float(0x7fc00000);
  if (V < 1 || V > k) {                                       // label_rows_kernel rejected the labels
    for (int p = lane; p < k; p += 32) row_of_col[p] = -1;
    if (lane < 3) loss3[lane] = qnan;
    return;
  }
  for (int j = lane; j < k; j += 32) { v[j] = 0.0; path[j] = -1; row4col[j] = -1; }
  for (int i = lane; i < V; i += 32) { u[i] = 0.0; col4row[i] = -1; }
  __syncwarp();
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  bool failed = false;
  for (int cur = 0; cur < V && !failed; ++cur) {
    for (int j = lane; j < k; j += 32) { remaining[j] = k - j - 1; SC[j] = 0; sp[j] = INF; }
    for (int i = lane; i < V; i += 32) SR[i] = 0;
    __syncwarp();
    double min_val = 0.0;
    int i = cur, sink = -1, R = k;
    while (sink < 0) {
      if (lane == 0) SR[i] = 1;
      const double ui = u[i];
      double best = INF;
      int bscore = -1, bit = -1;
      for (int it = lane; it < R; it += 32) {
        const int j = remaining[it];
        const double c = (double)__fadd_rn(__ldg(cost_ce + i * k + j), __ldg(cost_siou + i * k + j));      // evaluator.py:70 (fp32 sum)
        const double r = ((min_val + c) - ui) - v[j];
        double s = sp[j];
        if (r < s) { path[j] = i; sp[j] = r; s = r; }
        const int score = (row4col[j] < 0) ? R + it : R - 1 - it;
        if (s < best || (s == best && score > bscore)) { best = s; bscore = score; bit = it; }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const double ob = __shfl_xor_sync(FULL, best, d);
        const int os = __shfl_xor_sync(FULL, bscore, d), oi = __shfl_xor_sync(FULL, bit, d);
        if (ob < best || (ob == best && os > bscore)) { best = ob; bscore = os; bit = oi; }
      }
      if (bit < 0 || !(best < INF)) { failed = true; break; }       // NaN / infinite costs: no finite augmenting path
      min_val = best;
      const int j = remaining[bit];
      const int r4c = row4col[j];
      if (r4c < 0) sink = j; else i = r4c;
      __syncwarp();
      if (lane == 0) { SC[j] = 1; remaining[bit] = remaining[R - 1]; }
      --R;
      __syncwarp();
    }
    if (failed) break;
    if (lane == 0) u[cur] += min_val;
    for (int i2 = lane; i2 < V; i2 += 32)
      if (SR[i2] && i2 != cur) u[i2] += min_val - sp[col4row[i2]];
    for (int j = lane; j < k; j += 32)
      if (SC[j]) v[j] -= min_val - sp[j];
    __syncwarp();
    if (lane == 0) {                                                // augment along the path
      int j = sink;
      while (true) {
        const int i2 = path[j];
        row4col[j] = i2;
        const int t = col4row[i2];
        col4row[i2] = j;
        j = t;
        if (i2 == cur) break;
      }
    }
    __syncwarp();
  }
  if (failed) {
    for (int p = lane; p < k; p += 32) row_of_col[p] = -1;
    if (lane < 3) loss3[lane] = qnan;
    return;
  }
  double ce = 0.0, si = 0.0, inv = 0.0;
  for (int g = lane; g < V; g += 32) { ce += (double)cost_ce[g * k + col4row[g]]; si += (double)cost_siou[g * k + col4row[g]]; }
  for (int p = lane; p < k; p += 32) {
    row_of_col[p] = row4col[p];
    if (row4col[p] < 0) inv += (double)s_sum[p];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    ce += __shfl_xor_sync(FULL, ce, d); si += __shfl_xor_sync(FULL, si, d); inv += __shfl_xor_sync(FULL, inv, d);
  }
  if (lane == 0) {
    loss3[0] = (float)(ce / (double)V);
    loss3[1] = (k > V) ? (float)(inv / ((double)n * (double)(k - V))) : 0.0f;
    loss3[2] = (float)(si / (double)V);
  }
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
                         const int32_t* n_valid_dev, const float* tp, const float* s_sum, const float* cnt, const float* g3,
                         float* d_pred, cudaStream_t st) {
  DMN_CHECK(k >= 1 && k <= EV_MAX_K && (n_valid_dev || (n_valid >= 1 && n_valid <= k)), "ins_loss_grad: bad sizes (k %d, valid %d)", k,
            n_valid);
  const int64_t total = n * k;
  if (total == 0) return 0;
  ins_loss_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pred, gt_row, n, k, row_of_col, n_valid, n_valid_dev, tp,
                                                                        s_sum, cnt, g3, d_pred);
  DMN_LAUNCH_OK();
  return 0;
}

// Error word of the device-side assignment in mapped host memory (one per process): written by label_rows_kernel, read and
// cleared by ins_status_take() on the host without any synchronisation.
static volatile int32_t* g_ins_h_status = nullptr;
static int32_t* g_ins_d_status = nullptr;
static int ins_status_init() {
  if (g_ins_h_status) return 0;
  int32_t* h = nullptr;
  DMN_CUDA(cudaHostAlloc((void**)&h, sizeof(int32_t), cudaHostAllocMapped | cudaHostAllocPortable));
  *h = 0;
  DMN_CUDA(cudaHostGetDevicePointer((void**)&g_ins_d_status, (void*)h, 0));
  g_ins_h_status = h;
  return 0;
}
int ins_status_take() {
  if (!g_ins_h_status) return 0;
  const int code = (int)*g_ins_h_status;
  if (code) *g_ins_h_status = 0;
  return code;
}

int launch_label_rows(const int32_t* labels, int64_t n, int k, int32_t* gt_row, int32_t* n_valid, cudaStream_t st) {
  DMN_CHECK(k >= 1 && k <= EV_MAX_K, "ins_label_rows: ins_num %d out of range (max %d)", k, EV_MAX_K);
  DMN_CHECK(n >= 1, "ins_label_rows: empty batch");
  if (ins_status_init()) return 1;
  label_rows_kernel<<<1, 1024, 0, st>>>(labels, n, k, gt_row, n_valid, g_ins_d_status);
  DMN_LAUNCH_OK();
  return 0;
}

int launch_hungarian_assign(const float* cost_ce, const float* cost_siou, const float* s_sum, const int32_t* n_valid, int64_t n, int k,
                            int32_t* row_of_col, float* loss3, cudaStream_t st) {
  DMN_CHECK(k >= 1 && k <= EV_MAX_K, "hungarian_assign: ins_num %d out of range (max %d)", k, EV_MAX_K);
  DMN_CHECK(n >= 1, "hungarian_assign: empty batch");
  hungarian_assign_kernel<<<1, 32, 0, st>>>(cost_ce, cost_siou, s_sum, n_valid, n, k, row_of_col, loss3);
  DMN_LAUNCH_OK();
  return 0;
}

}  // namespace dmnerf

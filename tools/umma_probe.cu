// Hardware probe for the tcgen05 building blocks used by dm-nerf_b200/csrc/mlp_umma.cu.
// Runs small single-CTA GEMMs D[128,N] = A[128,K] * B[N,K]^T on one SM and checks them against a CPU reference:
//   mode 0  SS   : A and B from shared memory (K-major SWIZZLE_128B slabs written with generic stores)
//   mode 1  TS   : A from tensor memory (bf16 pairs packed along K, written with tcgen05.st), B from shared memory
//   mode 2  x3   : bf16x3 split product  Ahi*Bhi (SS) + Ahi*Blo (SS) + Alo*Bhi (TS)  vs exact fp64 -> precision of the scheme
//   mode 3  bulk : as mode 0, B slabs fetched from a pre-swizzled global image with cp.async.bulk + mbarrier
// plus a timing loop (cycles per MMA for SS and TS issue).  Every wait is bounded, so a wrong assumption yields
// FAIL lines, never a hung GPU.   Build: tools/build_tools.py   Run: tools/bin/umma_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../dm-nerf_b200/csrc/umma.cuh"

using namespace umma;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e = (x);                                                                       \
    if (e != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);           \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

struct Params {
  const float* A;      // [128][K]
  const float* B;      // [N][K]
  const uint8_t* Bimg; // pre-swizzled bf16(B) slabs [K/64][N*128 B]   (mode 3)
  float* D;            // [128][N]
  int N, K, mode, reps;
  long long* cycles;   // [1]
  int* status;         // 0 ok, else failure code
};

// Generic SWIZZLE_128B descriptor with explicit leading / stride byte offsets (MN-major experiments, modes 7 / 8).
__device__ __forceinline__ uint64_t make_sdesc_generic(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_mma, bar_load;
  __shared__ uint32_t tmem_base_s;
  // 1024 B align the dynamic region
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = p.N, K = p.K, KS = K / 64;
  uint8_t* sA = smem;                               // KS slabs of 128 rows
  uint8_t* sBhi = sA + (size_t)KS * 128 * 128;      // KS slabs of N rows
  uint8_t* sBlo = sBhi + (size_t)KS * N * 128;      // (mode 2)

  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_load, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t tD = tbase;                         // accumulator columns [0, N)
  const uint32_t tA = tbase + 256;                   // A operand columns [256, 256 + K/2)
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

  // ---- stage A: row = tid
  if (p.mode >= 7) {
    // MN-major operands: shared-memory rows are K indices, 64 consecutive M (or N) elements per 128-byte row, one slab of K
    // rows per group of 64 M / N elements.
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16_rn(p.A[(size_t)tid * K + k]);
      *reinterpret_cast<__nv_bfloat16*>(sA + (size_t)(tid >> 6) * K * 128 + sw128_offset(k, tid & 63)) = h;
    }
    for (int n = tid; n < N; n += 128)
      for (int k = 0; k < K; ++k) {
        __nv_bfloat16 h = __float2bfloat16_rn(p.B[(size_t)n * K + k]);
        *reinterpret_cast<__nv_bfloat16*>(sBhi + (size_t)(n >> 6) * K * 128 + sw128_offset(k, n & 63)) = h;
      }
  } else {
    const float* arow = p.A + (size_t)tid * K;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t hi[8], lo[8];
      for (int j = 0; j < 8; ++j) split_bf16x2(arow[k0 + 2 * j], arow[k0 + 2 * j + 1], hi[j], lo[j]);
      // shared-memory copy of A_hi (two 16-byte units)
      const int slab = k0 / 64, kk = k0 % 64;
      uint8_t* base = sA + (size_t)slab * 128 * 128;
      const uint32_t* sm = (p.mode == 4) ? lo : hi;      // mode 4: A_lo in shared memory, A_hi in tensor memory
      *reinterpret_cast<uint4*>(base + sw128_offset(tid, kk)) = make_uint4(sm[0], sm[1], sm[2], sm[3]);
      *reinterpret_cast<uint4*>(base + sw128_offset(tid, kk + 8)) = make_uint4(sm[4], sm[5], sm[6], sm[7]);
      if (p.mode == 1 || p.mode == 4 || p.mode == 5) tmem_st_x8(tA + lane_sel + k0 / 2, hi);
      if (p.mode == 2) tmem_st_x8(tA + lane_sel + k0 / 2, lo);
    }
  }
  // ---- stage B
  if (p.mode >= 7) {
  } else if (p.mode == 3) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_load, (uint32_t)(KS * N * 128));
      for (int s = 0; s < KS; ++s)
        bulk_g2s(sBhi + (size_t)s * N * 128, p.Bimg + (size_t)s * N * 128, (uint32_t)(N * 128), &bar_load);
    }
  } else {
    for (int n = tid; n < N; n += 128) {
      const float* brow = p.B + (size_t)n * K;
      for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t hi[4], lo[4];
        for (int j = 0; j < 4; ++j) split_bf16x2(brow[k0 + 2 * j], brow[k0 + 2 * j + 1], hi[j], lo[j]);
        const int slab = k0 / 64, kk = k0 % 64;
        *reinterpret_cast<uint4*>(sBhi + (size_t)slab * N * 128 + sw128_offset(n, kk)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (p.mode == 2 || p.mode == 4)
          *reinterpret_cast<uint4*>(sBlo + (size_t)slab * N * 128 + sw128_offset(n, kk)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  fence_proxy_async_smem();
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();

  const uint32_t idesc = make_idesc_bf16(128, N) | (p.mode >= 7 ? ((1u << 15) | (1u << 16)) : 0u);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {                                   // converged warp, one elected lane issues (uniform-register operands)
    bool ok = true;
    if (p.mode == 3) ok = mbar_wait(&bar_load, 0);
    if (!ok) atomicExch(p.status, 10);
    tc_fence_after();
    t0 = clock64();
    if (elect_one()) {
    if (p.mode >= 7) {
      const uint32_t slab = (uint32_t)K * 128u;                    // bytes between groups of 64 M / N elements
      for (int rep = 0; rep < p.reps; ++rep)
        for (int k0 = 0; k0 < K; k0 += 16) {
          const uint32_t lbo = (p.mode == 7) ? slab : 1024u, sbo = (p.mode == 7) ? 1024u : slab;
          const uint64_t da = make_sdesc_generic(smem_u32(sA) + (k0 >> 3) * 1024, lbo, sbo);
          const uint64_t db = make_sdesc_generic(smem_u32(sBhi) + (k0 >> 3) * 1024, lbo, sbo);
          mma_ss(tD, da, db, idesc, (rep > 0 || k0 > 0) ? 1u : 0u);
        }
    } else if (p.mode == 5 || p.mode == 6) {
      // raw tensor-pipe rate: 32 MMAs per iteration, operands fixed, no address arithmetic between issues
      const uint64_t db = make_sdesc_sw128(smem_u32(sBhi));
      const uint64_t da0 = make_sdesc_sw128(smem_u32(sA));
      for (int rep = 0; rep < p.reps; ++rep) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          if (p.mode == 5) mma_ts(tD, tA + (u & 3) * 8, db + 2 * (u & 3), idesc, (rep | u) ? 1u : 0u);
          else mma_ss(tD, da0 + 2 * (u & 3), db + 2 * (u & 3), idesc, (rep | u) ? 1u : 0u);
        }
      }
    } else
    for (int rep = 0; rep < p.reps; ++rep) {
      for (int k0 = 0; k0 < K; k0 += 16) {
        const int slab = k0 / 64, kk = k0 % 64;
        const uint32_t acc = (rep > 0 || k0 > 0) ? 1u : 0u;
        const uint64_t da = make_sdesc_sw128(smem_u32(sA + (size_t)slab * 128 * 128) + kk * 2);
        const uint64_t dbh = make_sdesc_sw128(smem_u32(sBhi + (size_t)slab * N * 128) + kk * 2);
        if (p.mode == 0 || p.mode == 3) {
          mma_ss(tD, da, dbh, idesc, acc);
        } else if (p.mode == 1) {
          mma_ts(tD, tA + k0 / 2, dbh, idesc, acc);
        } else if (p.mode == 2) {
          const uint64_t dbl = make_sdesc_sw128(smem_u32(sBlo + (size_t)slab * N * 128) + kk * 2);
          mma_ss(tD, da, dbh, idesc, acc);
          mma_ss(tD, da, dbl, idesc, 1u);
          mma_ts(tD, tA + k0 / 2, dbh, idesc, 1u);
        } else {
          // mode 4: the production schedule of mlp_umma.cu -- per 64-wide chunk: 4x (A_hi*B_hi TS, A_lo*B_hi SS), then 4x A_hi*B_lo TS
          if (kk == 0) {
            for (int q2 = 0; q2 < 4; ++q2) {
              const uint64_t dh = make_sdesc_sw128(smem_u32(sBhi + (size_t)slab * N * 128) + q2 * 32);
              const uint64_t dl = make_sdesc_sw128(smem_u32(sA + (size_t)slab * 128 * 128) + q2 * 32);
              mma_ts(tD, tA + (k0 + q2 * 16) / 2, dh, idesc, (rep > 0 || k0 > 0 || q2 > 0) ? 1u : 0u);
              mma_ss(tD, dl, dh, idesc, 1u);
            }
            for (int q2 = 0; q2 < 4; ++q2) {
              const uint64_t dlo = make_sdesc_sw128(smem_u32(sBlo + (size_t)slab * N * 128) + q2 * 32);
              mma_ts(tD, tA + (k0 + q2 * 16) / 2, dlo, idesc, 1u);
            }
          }
        }
      }
    }
    mma_commit(&bar_mma);
    }
    __syncwarp();
  }
  const bool done = mbar_wait(&bar_mma, 0);
  if (tid == 0) {
    t1 = clock64();
    p.cycles[0] = t1 - t0;
  }
  if (!done) atomicExch(p.status, 11);
  tc_fence_after();
  if (done) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld_x16(tD + lane_sel + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) p.D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  const uint32_t r = 0x7FFFu + ((u >> 16) & 1u);
  u = (u + r) & 0xFFFF0000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

static int run_case(const char* name, int mode, int N, int K, int reps, bool check) {
  std::vector<float> A((size_t)128 * K), B((size_t)N * K);
  srand(1234 + mode * 7 + N + K);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.125f;
  // pre-swizzled image of bf16(B)
  std::vector<uint8_t> img((size_t)(K / 64) * N * 128);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      float r = bf16_round(B[(size_t)n * K + k]);
      uint32_t u;
      memcpy(&u, &r, 4);
      uint16_t h = (uint16_t)(u >> 16);
      memcpy(&img[(size_t)(k / 64) * N * 128 + sw128_offset(n, k % 64)], &h, 2);
    }
  float *dA, *dB, *dD;
  uint8_t* dImg;
  long long* dCyc;
  int* dStatus;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, (size_t)128 * N * 4));
  CK(cudaMalloc(&dImg, img.size())); CK(cudaMalloc(&dCyc, 8)); CK(cudaMalloc(&dStatus, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dImg, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, (size_t)128 * N * 4)); CK(cudaMemset(dStatus, 0, 4)); CK(cudaMemset(dCyc, 0, 8));
  Params p{dA, dB, dImg, dD, N, K, mode, reps, dCyc, dStatus};
  const size_t smem = (size_t)(K / 64) * 128 * 128 + (size_t)(K / 64) * N * 128 * ((mode == 2 || mode == 4) ? 2 : 1) + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-34s FAIL  kernel error: %s\n", name, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> D((size_t)128 * N);
  long long cyc = 0;
  int status = 0;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&status, dStatus, 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dImg); cudaFree(dCyc); cudaFree(dStatus);
  const int n_mma = (mode == 5 || mode == 6) ? reps * 32 : reps * (K / 16) * ((mode == 2 || mode == 4) ? 3 : 1);
  if (status) {
    printf("%-34s FAIL  status=%d (barrier timeout)\n", name, status);
    return 1;
  }
  if (!check) {
    printf("%-34s TIME  %lld cycles / %d MMA = %.1f cyc per 128x%dx16 MMA\n", name, cyc, n_mma, (double)cyc / n_mma, N);
    return 0;
  }
  double max_err = 0, max_ref = 0, sum_sq_err = 0, sum_sq_ref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) {
        double a = A[(size_t)m * K + k], b = B[(size_t)n * K + k];
        if (mode != 2 && mode != 4) { a = bf16_round((float)a); b = bf16_round((float)b); }
        ref += a * b;
      }
      ref *= reps;
      const double err = fabs((double)D[(size_t)m * N + n] - ref);
      max_err = fmax(max_err, err); max_ref = fmax(max_ref, fabs(ref));
      sum_sq_err += err * err; sum_sq_ref += ref * ref;
    }
  const double rel_l2 = sqrt(sum_sq_err / sum_sq_ref);
  const double tol = (mode == 2 || mode == 4) ? 5e-5 : 2e-5;   // vs bf16-exact inputs (modes 0,1,3) / vs fp64 of fp32 inputs (mode 2)
  const bool ok = std::isfinite(rel_l2) && rel_l2 < tol;
  printf("%-34s %s  relL2=%.3e maxabs=%.3e (max|ref|=%.3f)  %lld cyc for %d MMA\n", name, ok ? "PASS" : "FAIL", rel_l2,
         max_err, max_ref, cyc, n_mma);
  return ok ? 0 : 1;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  int fails = 0;
  fails += run_case("SS   N=256 K=64", 0, 256, 64, 1, true);
  fails += run_case("SS   N=256 K=256 (4 slabs)", 0, 256, 256, 1, true);
  fails += run_case("SS   N=128 K=128", 0, 128, 128, 1, true);
  fails += run_case("SS   N=16  K=128", 0, 16, 128, 1, true);
  fails += run_case("SS   N=112 K=64", 0, 112, 64, 1, true);
  fails += run_case("TS   N=256 K=64", 1, 256, 64, 1, true);
  fails += run_case("TS   N=256 K=256", 1, 256, 256, 1, true);
  fails += run_case("x3   N=256 K=128 (bf16x3 vs fp64)", 2, 256, 128, 1, true);
  fails += run_case("bulk N=256 K=256 (cp.async.bulk)", 3, 256, 256, 1, true);
  run_case("time SS N=256 K=256 x64", 0, 256, 256, 64, false);
  run_case("time TS N=256 K=256 x64", 1, 256, 256, 64, false);
  run_case("time SS N=128 K=256 x64", 0, 128, 256, 64, false);
  run_case("time x3 N=256 K=128 x64", 2, 256, 128, 64, false);
  fails += run_case("prod N=128 K=128 (TS hi, SS lo)", 4, 128, 128, 1, true);
  fails += run_case("MN-major A,B N=256 K=64  (LBO=slab)", 7, 256, 64, 1, true);
  fails += run_case("MN-major A,B N=256 K=128 (LBO=slab)", 7, 256, 128, 1, true);
  fails += run_case("MN-major A,B N=128 K=64  (LBO=slab)", 7, 128, 64, 1, true);
  run_case("MN-major A,B N=256 K=64  (SBO=slab: expected FAIL)", 8, 256, 64, 1, true);
  run_case("time MN-major SS N=256 K=128 x64", 7, 256, 128, 64, false);
  run_case("time TS   N=128 K=256 x64", 1, 128, 256, 64, false);
  run_case("time prod N=128 K=128 x64", 4, 128, 128, 64, false);
  run_case("time prod N=256 K=128 x64", 4, 256, 128, 64, false);
  run_case("time prod N=64  K=128 x64", 4, 64, 128, 64, false);
  run_case("raw TS N=256 (unrolled x32)", 5, 256, 64, 64, false);
  run_case("raw TS N=128 (unrolled x32)", 5, 128, 64, 64, false);
  run_case("raw TS N=64  (unrolled x32)", 5, 64, 64, 64, false);
  run_case("raw SS N=256 (unrolled x32)", 6, 256, 64, 64, false);
  run_case("raw SS N=128 (unrolled x32)", 6, 128, 64, 64, false);
  run_case("raw SS N=64  (unrolled x32)", 6, 64, 64, 64, false);
  printf("umma_probe: %d failing case(s)\n", fails);
  return fails ? 1 : 0;
}

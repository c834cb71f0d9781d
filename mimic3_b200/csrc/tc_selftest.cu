// Self-test of the tcgen05 building blocks (descriptor conventions, row-shifted A operand,
// TMEM st / accumulate-on-top / ld) against a host reference.  Exposed as m3_selftest();
// run by tests/test_gpu_tc.py so a descriptor mistake is caught in isolation.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../include/m3b200.h"
#include "tc_common.cuh"

namespace m3 {
namespace {

template <int FMT>
__global__ void __launch_bounds__(128) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                       const float* __restrict__ init, float* __restrict__ D, int K,
                                                       int N, int taps, int dil, int rows, uint32_t lboA,
                                                       uint32_t sboA, uint32_t lboB, uint32_t sboB) {
  extern __shared__ __align__(128) uint8_t smem[];
  using E = tc::Elem<FMT>;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* sA = reinterpret_cast<uint32_t*>(smem);                      // [K/8][rows][8] elements
  uint32_t* sW = reinterpret_cast<uint32_t*>(smem + size_t(K / 8) * rows * 16);  // [taps][K/8][N][8]

  for (int idx = tid; idx < rows * (K / 2); idx += 128) {
    const int r = idx / (K / 2), k2 = idx - r * (K / 2);
    const int k = k2 * 2;
    const uint32_t v = E::pack2(A[r * K + k], A[r * K + k + 1]);
    sA[((k >> 3) * rows + r) * 4 + ((k & 7) >> 1)] = v;
  }
  for (int idx = tid; idx < taps * N * (K / 2); idx += 128) {
    const int tap = idx / (N * (K / 2));
    const int rem = idx - tap * N * (K / 2);
    const int n = rem / (K / 2), k = (rem - n * (K / 2)) * 2;
    const uint32_t v = E::pack2(W[(tap * N + n) * K + k], W[(tap * N + n) * K + k + 1]);
    sW[((tap * (K / 8) + (k >> 3)) * N + n) * 4 + ((k & 7) >> 1)] = v;
  }
  if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);

  if (init) {  // pre-load the accumulator: the MMAs must add on top of it
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      for (int j = 0; j < 16; ++j) v[j] = init[(warp * 32 + lane) * N + c0 + j];
      tc::tmem_st16(trow + c0, v);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(128, N, FMT);
    const uint32_t aBase = tc::smem_u32(sA), wBase = tc::smem_u32(sW);
    uint32_t acc = init ? 1u : 0u;
    for (int tap = 0; tap < taps; ++tap) {
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = tc::make_desc(aBase + uint32_t(tap * dil) * 16u + uint32_t(ks) * 2u * uint32_t(rows) * 16u, lboA, sboA);
        const uint64_t bd = tc::make_desc(wBase + uint32_t(tap * (K / 8) + ks * 2) * uint32_t(N) * 16u, lboB, sboB);
        tc::mma_f16_ss(tmem, ad, bd, idesc, acc);
        acc = 1u;
      }
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(trow + c0, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

double run_probe(int fmt, int K, int N, int taps, int dil, bool with_init, bool swap_lbo_sbo) {
  const int rows = 128 + (taps - 1) * dil;
  std::vector<float> A(size_t(rows) * K), W(size_t(taps) * N * K), I(size_t(128) * N), D(size_t(128) * N);
  uint32_t s = 12345u + K * 7 + N * 13 + taps;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return float(int((s >> 9) & 0xFFFF) - 32768) / 32768.0f;
  };
  auto q = [&](float x) { return fmt == 1 ? tc::Elem<1>::round(x) : tc::Elem<0>::round(x); };
  for (auto& v : A) v = q(rnd());
  for (auto& v : W) v = q(rnd() * 0.25f);
  for (auto& v : I) v = rnd() * 3.0f;
  float *dA, *dW, *dI, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&dI, I.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dI, I.data(), I.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  uint32_t lboA = rows * 16, sboA = 128, lboB = N * 16, sboB = 128;
  if (swap_lbo_sbo) {
    std::swap(lboA, sboA);
    std::swap(lboB, sboB);
  }
  const size_t smem = size_t(K / 8) * rows * 16 + size_t(taps) * (K / 8) * N * 16;
  double err = 1e30;
  cudaError_t e;
  if (smem > 227 * 1024) return -1000.0;
  if (fmt == 1) {
    cudaFuncSetAttribute(tc_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    tc_probe_kernel<1><<<1, 128, smem>>>(dA, dW, with_init ? dI : nullptr, dD, K, N, taps, dil, rows, lboA, sboA, lboB, sboB);
  } else {
    cudaFuncSetAttribute(tc_probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    tc_probe_kernel<0><<<1, 128, smem>>>(dA, dW, with_init ? dI : nullptr, dD, K, N, taps, dil, rows, lboA, sboA, lboB, sboB);
  }
  e = cudaDeviceSynchronize();
  if (e == cudaSuccess) {
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    err = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double acc = with_init ? I[size_t(m) * N + n] : 0.0;
        for (int tap = 0; tap < taps; ++tap)
          for (int k = 0; k < K; ++k) acc += double(A[size_t(m + tap * dil) * K + k]) * W[(size_t(tap) * N + n) * K + k];
        err = std::fmax(err, std::fabs(acc - D[size_t(m) * N + n]));
      }
  } else {
    err = -double(int(e));
    cudaGetLastError();
  }
  cudaFree(dA);
  cudaFree(dW);
  cudaFree(dI);
  cudaFree(dD);
  return err;
}

// ---------------------------------------------------------------------------------------
// Micro-benchmarks (cycles, single CTA): what the MRF design depends on but no document states
// for sm_100a -- TMEM ld/st rate, SS-mode MMA rate for small N, commit->mbarrier latency.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_ubench_kernel(int mode, int N, int reps, int warps_active,
                                                        long long* out, const uint8_t* gsrc) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16);
  long long t0 = 0, t1 = 0;
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = float(tid + i);
  if (mode == 0 || mode == 1 || mode == 6) {  // TMEM ld (0: wait per ld, 6: 4 lds per wait) / st (1)
    __syncthreads();
    t0 = clock64();
    if (warp < warps_active) {
      float acc = 0.f;
      for (int r = 0; r < reps; ++r) {
        if (mode == 0) {
          tc::tmem_ld16(lane_base + ((r * 16) & 255), v);
          tc::tmem_ld_wait();
          acc += v[0];
        } else if (mode == 6) {
          float a[16], b[16], c[16], d[16];
          tc::tmem_ld16(lane_base + 0, a);
          tc::tmem_ld16(lane_base + 16, b);
          tc::tmem_ld16(lane_base + 32, c);
          tc::tmem_ld16(lane_base + 48, d);
          tc::tmem_ld_wait();
          acc += a[0] + b[0] + c[0] + d[0];
        } else {
          tc::tmem_st16(lane_base + ((r * 16) & 255), v);
        }
      }
      if (mode == 1) tc::tmem_st_wait();
      if (acc == 123.456f) out[63] = 1;
    }
    __syncthreads();
    t1 = clock64();
  } else if (mode == 2 || mode == 3) {  // MMA stream: reps MMAs then one commit (2), or commit+wait each (3)
    const uint32_t idesc = tc::make_idesc(128, N, 0);
    const uint64_t ad = tc::make_desc(tc::smem_u32(smem), 128u * 16u, 128u);
    const uint64_t bd = tc::make_desc(tc::smem_u32(smem) + 16384u, uint32_t(N) * 16u, 128u);
    __syncthreads();
    t0 = clock64();
    if (tid == 0) {
      uint32_t ph = 0;
      for (int r = 0; r < reps; ++r) {
        tc::mma_f16_ss(tmem, ad, bd, idesc, r ? 1u : 0u);
        if (mode == 3) {
          tc::mma_commit(&bar);
          tc::mbar_wait(&bar, ph);
          ph ^= 1u;
        }
      }
      if (mode == 2) {
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
      }
    }
    __syncthreads();
    t1 = clock64();
  } else if (mode >= 7 && mode <= 9) {  // warp-uniform branch + elect.sync issue (no waterfall loop)
    // mode 7: one accumulator; 8: four accumulators round robin; 9: like 8 with M=64
    const uint32_t idesc = tc::make_idesc(mode == 9 ? 64 : 128, N, 0);
    const uint64_t ad = tc::make_desc(tc::smem_u32(smem), 128u * 16u, 128u);
    const uint64_t bd = tc::make_desc(tc::smem_u32(smem) + 16384u, uint32_t(N) * 16u, 128u);
    __syncthreads();
    long long e0 = 0, e1 = 0;
    if (warp == 0) {
      if (tc::elect_one()) {
        e0 = clock64();
        for (int r = 0; r < reps; r += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            tc::mma_f16_ss(tmem + (mode >= 8 ? uint32_t(u * 128) : 0u), ad, bd, idesc, r ? 1u : 0u);
        }
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        e1 = clock64();
        out[1] = e1 - e0;
      }
      __syncwarp();
    }
    __syncthreads();
    t0 = 0;
    t1 = 0;
  } else if (mode == 10) {
    // What a conv-by-descriptor-shift kernel really issues: `warps_active` is a bit mask of deviations from the
    // idealised stream of modes 7-9 (same aligned operands every time):
    //   1: A starts one row (16 B) into its first core matrix    2: odd row pitch (LBO = 133 rows)
    //   4: four different A start rows (0, 3, 6, 9) round robin   8: four different B blocks round robin
    //  16: warps 1-4 stream 128-bit STS/LDS on another region    32: warp 5 streams 8 KB bulk copies into smem
    //  64: A starts half a core matrix (64 B) in
    const int var = warps_active;
    __shared__ volatile int stop_flag;
    __shared__ __align__(8) uint64_t bbar;
    if (tid == 0) {
      stop_flag = 0;
      tc::mbar_init(&bbar, 1);
      tc::mbar_fence_init();
    }
    const uint32_t pitch = (var & 2) ? 133u : 128u;
    const uint32_t idesc = tc::make_idesc(128, N, 0);
    const uint32_t a0 = tc::smem_u32(smem) + ((var & 1) ? 16u : 0u) + ((var & 64) ? 64u : 0u);
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ad[u] = tc::make_desc(a0 + ((var & 4) ? uint32_t(u * 3 * 16) : 0u), pitch * 16u, 128u);
      bd[u] = tc::make_desc(tc::smem_u32(smem) + 16384u + ((var & 8) ? uint32_t(u * N * 32) : 0u), uint32_t(N) * 16u, 128u);
    }
    __syncthreads();
    if (warp == 0) {
      if (tc::elect_one()) {
        const long long e0 = clock64();
        for (int r = 0; r < reps; r += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) tc::mma_f16_ss(tmem + uint32_t(u * 128), ad[u], bd[u], idesc, r ? 1u : 0u);
        }
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        out[1] = clock64() - e0;
        stop_flag = 1;
      }
      __syncwarp();
    } else if ((var & 16) && warp >= 1 && warp <= 4) {
      uint4* region = reinterpret_cast<uint4*>(smem + 40960) + (warp - 1) * 256 + (tid & 31);
      uint4 x = make_uint4(tid, 1u, 2u, 3u);
      while (!stop_flag) {
#pragma unroll
        for (int i = 0; i < 8; ++i) region[i * 32] = x;
#pragma unroll
        for (int i = 0; i < 4; ++i) x.x += region[i * 32].y;
      }
      if (x.x == 0x12345u) out[62] = 1;
    } else if ((var & 32) && warp == 5) {
      if (tc::elect_one()) {
        uint32_t ph = 0;
        while (!stop_flag) {
          tc::mbar_expect_tx(&bbar, 8192u);
          tc::bulk_g2s(smem + 57344, gsrc, 8192u, &bbar);
          tc::mbar_wait(&bbar, ph);
          ph ^= 1u;
        }
      }
      __syncwarp();
    }
    __syncthreads();
    t0 = 0;
    t1 = 0;
  } else if (mode == 11 || mode == 12) {
    // the per-stage pattern of the streaming kernels: `warps_active` MMAs, then tcgen05.commit to an mbarrier
    // (mode 12: plus a try_wait on an mbarrier whose phase is already complete and tcgen05.fence::after_thread_sync,
    // i.e. the "weights have landed" check that precedes every stage)
    const int period = warps_active > 0 ? warps_active : 12;
    __shared__ __align__(8) uint64_t cbar[2];
    if (tid == 0) {
      tc::mbar_init(&cbar[0], 1);
      tc::mbar_init(&cbar[1], 1);
      tc::mbar_fence_init();
    }
    const uint32_t idesc = tc::make_idesc(128, N, 0);
    const uint32_t a0 = tc::smem_u32(smem) + 16u, b0 = tc::smem_u32(smem) + 16384u;
    __syncthreads();
    if (tid == 0) tc::mbar_arrive(&cbar[1]);  // phase 0 of cbar[1] is complete from now on
    __syncthreads();
    if (warp == 0) {
      if (tc::elect_one()) {
        uint64_t ad4[4], bd4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ad4[u] = tc::make_desc(a0 + uint32_t(u) * 48u, 133u * 16u, 128u);
          bd4[u] = tc::make_desc(b0 + uint32_t(u) * uint32_t(N) * 32u, uint32_t(N) * 16u, 128u);
        }
        int stage = 0;
        const long long e0 = clock64();
        for (int r = 0; r < reps; r += period) {
          if (mode == 12) {
            tc::mbar_wait(&cbar[1], 0u);
            tc::fence_after_sync();
          }
          const uint32_t dcol = tmem + uint32_t(stage++ & 3) * 128u;
          for (int u = 0; u < period; ++u) {
            const int w = u & 3;
            const uint64_t ad = w == 0 ? ad4[0] : w == 1 ? ad4[1] : w == 2 ? ad4[2] : ad4[3];
            const uint64_t bd = w == 0 ? bd4[0] : w == 1 ? bd4[1] : w == 2 ? bd4[2] : bd4[3];
            tc::mma_f16_ss(dcol, ad, bd, idesc, (r | u) ? 1u : 0u);
          }
          tc::mma_commit(&cbar[0]);
        }
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        out[1] = clock64() - e0;
      }
      __syncwarp();
    }
    __syncthreads();
    t0 = 0;
    t1 = 0;
  } else if (mode == 13) {
    // epilogue-style TMEM reads (warps 1-7, three x16 loads per round) with or without a concurrent MMA stream (warp 0):
    // warps_active bit 0: all three loads before ONE tcgen05.wait::ld (else a wait after each load);
    // bit 1: warp 0 streams N-column MMAs into columns [256, 512) meanwhile.  Result: cycles per round.
    const int var = warps_active;
    __shared__ volatile int stop13;
    if (tid == 0) stop13 = 0;
    const uint32_t idesc = tc::make_idesc(128, N, 0);
    const uint64_t ad = tc::make_desc(tc::smem_u32(smem) + 16u, 133u * 16u, 128u);
    const uint64_t bd = tc::make_desc(tc::smem_u32(smem) + 16384u, uint32_t(N) * 16u, 128u);
    __syncthreads();
    if (warp == 0) {
      if ((var & 2) && tc::elect_one()) {
        long long n = 0;
        while (!stop13) {
#pragma unroll
          for (int u = 0; u < 8; ++u) tc::mma_f16_ss(tmem + 256u + uint32_t(u & 1) * 128u, ad, bd, idesc, 1u);
          tc::mma_commit(&bar);
          n += 8;
          if ((n & 63) == 0) {  // keep the queue bounded: wait for every 8th batch
            tc::mbar_wait(&bar, uint32_t((n / 8 - 1) & 1));
          } else {
            tc::mbar_wait(&bar, uint32_t((n / 8 - 1) & 1));
          }
        }
        out[2] = n;
      }
      __syncwarp();
    } else {
      float a[16], b[16], c[16];
      float acc = 0.f;
      const long long e0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (var & 1) {
          tc::tmem_ld16(lane_base + 0, a);
          tc::tmem_ld16(lane_base + 32, b);
          tc::tmem_ld16(lane_base + 64, c);
          tc::tmem_ld_wait();
        } else {
          tc::tmem_ld16(lane_base + 0, a);
          tc::tmem_ld_wait();
          tc::tmem_ld16(lane_base + 32, b);
          tc::tmem_ld_wait();
          tc::tmem_ld16(lane_base + 64, c);
          tc::tmem_ld_wait();
        }
        acc += a[0] + b[1] + c[2];
      }
      const long long e1 = clock64();
      if (acc == 123.456f) out[63] = 1;
      asm volatile("bar.sync 3, 224;\n" ::: "memory");  // warps 1-7
      if (tid == 32) {
        out[1] = e1 - e0;
        stop13 = 1;
      }
    }
    __syncthreads();
    t0 = 0;
    t1 = 0;
  } else if (mode == 4) {  // one conv-on-one-tile round trip: st -> sync -> MMA(s) -> commit -> wait -> ld
    const uint32_t idesc = tc::make_idesc(128, N, 0);
    const uint64_t ad = tc::make_desc(tc::smem_u32(smem), 128u * 16u, 128u);
    const uint64_t bd = tc::make_desc(tc::smem_u32(smem) + 16384u, uint32_t(N) * 16u, 128u);
    uint32_t ph = 0;
    __syncthreads();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      tc::tmem_st16(lane_base, v);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      if (tid == 0) {
        for (int i = 0; i < 6; ++i) tc::mma_f16_ss(tmem, ad, bd, idesc, 1u);
        tc::mma_commit(&bar);
      }
      tc::mbar_wait(&bar, ph);
      ph ^= 1u;
      tc::fence_after_sync();
      tc::tmem_ld16(lane_base, v);
      tc::tmem_ld_wait();
    }
    __syncthreads();
    t1 = clock64();
    if (v[0] == 123.456f) out[63] = 1;
  } else if (mode == 5) {  // __syncthreads cost
    __syncthreads();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) __syncthreads();
    t1 = clock64();
  }
  if (tid == 0 && !(mode >= 7 && mode <= 13)) out[1] = t1 - t0;
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

double run_ubench(int mode, int N, int reps, int warps) {
  long long* d;
  cudaMalloc(&d, 64 * sizeof(long long));
  cudaMemset(d, 0, 64 * sizeof(long long));
  cudaFuncSetAttribute(tc_ubench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  uint8_t* gsrc = nullptr;
  cudaMalloc(&gsrc, 64 * 1024);
  cudaMemset(gsrc, 0, 64 * 1024);
  tc_ubench_kernel<<<1, 256, 64 * 1024>>>(mode, N, reps, warps, d, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(gsrc);
  long long h = -1;
  if (e == cudaSuccess) cudaMemcpy(&h, d + 1, sizeof h, cudaMemcpyDeviceToHost);
  else {
    h = -(long long)e;
    cudaGetLastError();
  }
  cudaFree(d);
  return double(h) / reps;
}

}  // namespace
}  // namespace m3

extern "C" int32_t m3_selftest(int32_t which, double* result) {
  if (!result) return M3_ERR_INVALID;
  using m3::run_probe;
  switch (which) {
    case 0: *result = run_probe(1, 32, 32, 3, 1, false, false); break;    // bf16, MRF stage-3 shape
    case 1: *result = run_probe(1, 64, 64, 7, 12, false, false); break;   // dilated k=7
    case 2: *result = run_probe(1, 128, 128, 5, 6, true, false); break;   // accumulate on top of tcgen05.st
    case 3: *result = run_probe(0, 32, 32, 3, 2, true, false); break;     // fp16 operands
    case 4: *result = run_probe(1, 192, 256, 1, 1, false, false); break;  // flow WN in_layer shape (N=256)
    case 5: *result = run_probe(1, 32, 32, 3, 1, false, true); break;     // diagnostic: LBO/SBO swapped
    case 6: *result = run_probe(1, 96, 192, 1, 1, false, false); break;   // 1x1 conv
    case 7: *result = run_probe(0, 64, 64, 7, 3, true, false); break;
    // micro-benchmarks: cycles per repetition
    case 100: *result = m3::run_ubench(0, 0, 256, 4); break;   // tcgen05.ld x16 + wait, 4 warps
    case 101: *result = m3::run_ubench(0, 0, 256, 8); break;   // 8 warps
    case 102: *result = m3::run_ubench(0, 0, 256, 1); break;   // 1 warp (latency)
    case 103: *result = m3::run_ubench(6, 0, 256, 4); break;   // 4 x ld16 per wait, 4 warps
    case 104: *result = m3::run_ubench(6, 0, 256, 8); break;
    case 105: *result = m3::run_ubench(1, 0, 256, 4); break;   // tcgen05.st x16, 4 warps
    case 106: *result = m3::run_ubench(1, 0, 256, 8); break;
    case 110: *result = m3::run_ubench(2, 32, 512, 0); break;  // SS MMA stream, N=32
    case 111: *result = m3::run_ubench(2, 64, 512, 0); break;
    case 112: *result = m3::run_ubench(2, 128, 512, 0); break;
    case 113: *result = m3::run_ubench(2, 256, 512, 0); break;
    case 114: *result = m3::run_ubench(7, 32, 512, 0); break;   // elect-issued, one accumulator, N=32
    case 115: *result = m3::run_ubench(7, 128, 512, 0); break;
    case 116: *result = m3::run_ubench(8, 32, 512, 0); break;   // elect-issued, 4 accumulators, N=32
    case 117: *result = m3::run_ubench(8, 64, 512, 0); break;
    case 118: *result = m3::run_ubench(8, 128, 512, 0); break;
    case 119: *result = m3::run_ubench(9, 32, 512, 0); break;   // M=64, 4 accumulators
    case 120: *result = m3::run_ubench(3, 32, 64, 0); break;   // MMA + commit + wait latency
    case 121: *result = m3::run_ubench(3, 128, 64, 0); break;
    case 130: *result = m3::run_ubench(4, 32, 64, 0); break;   // st/sync/6 MMA/commit/wait/ld round trip
    case 131: *result = m3::run_ubench(4, 128, 64, 0); break;
    case 140: *result = m3::run_ubench(5, 0, 256, 0); break;   // __syncthreads
    default:
      // 200 + 100 * {0: N=32, 1: N=64, 2: N=128} + variant mask (mode 10): MMA stream as the conv kernels issue it
      if (which >= 700 && which < 800) {  // 700 + variant: TMEM reads of 7 warps, sequential / batched, without / with an N=32 MMA stream
        *result = m3::run_ubench(13, 32, 256, (which - 700) & 3);
        return M3_OK;
      }
      if (which >= 500 && which < 700) {  // 500 + period: N=64, commit every `period` MMAs; 600 + period: + wait/fence per stage
        *result = m3::run_ubench(which < 600 ? 11 : 12, 64, 480, (which - 500) % 100);
        return M3_OK;
      }
      if (which >= 200 && which < 500) {
        const int n = which < 300 ? 32 : which < 400 ? 64 : 128;
        *result = m3::run_ubench(10, n, 512, (which - 200) % 100);
        return M3_OK;
      }
      return M3_ERR_INVALID;
  }
  return M3_OK;
}

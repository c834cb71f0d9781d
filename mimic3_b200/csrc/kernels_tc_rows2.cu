// Token-level Conv1d-as-GEMM with fp16 hi/lo split operands, A-STATIONARY variant.
// STAGED FOR ROUND 2 -- off by default (M3B200_ROWGEMM_V2=1 selects it; tests/test_gpu_experimental.py); it has
// not run on hardware yet.  Same arithmetic, same packed weights ([chunk][K block][tap][hi|lo][4][64][8]) and the
// same per-accumulator MMA order as rowgemm_tc_kernel (kernels_tc_rows.cu), so results are bit-identical; what
// changes is who pays for the A operand:
//   * rowgemm_tc_kernel: one CTA per 128 x 64 output tile; every tile re-loads and re-splits its 128 x K fp32 rows
//     (9 x for q|k|v, 12 x for FFN-1) and lives for one short load -> convert -> MMA -> epilogue chain (ncu, round 1:
//     tensor pipe 9-24 %, 26-34 % of the stall samples on the A loads);
//   * here: one CTA per 128-row tile x a RANGE of 64-column chunks, organised like conv_tc_kernel: the A rows are
//     loaded and split ONCE into shared memory (all K blocks resident), warp 0 streams the weight blocks of
//     (chunk, K block) through a bulk-copy ring, warp 1 issues the MMAs into main+correction accumulators that are
//     double-buffered in TMEM (2 x 128 columns), and eight epilogue warps drain chunk c while chunk c+1 is computed.
// K is limited by the resident A (K <= 288 with k = 3): the 768-channel FFN-2 keeps the per-tile kernel.
#include <algorithm>
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

namespace {
constexpr int R2_KB = 32;        // K block (matches the packed weights)
constexpr int R2_NC = 64;        // output columns per chunk (matches the packed weights)
constexpr int R2_THREADS = 320;  // warp 0 producer, warp 1 issuer, warps 2-9 epilogue
constexpr int R2_MAXSLOT = 6;
constexpr int R2_SMEM_MAX = 224 * 1024;
constexpr int R2_LOADS = 5;      // A items in flight per thread while staging

__global__ void __launch_bounds__(R2_THREADS, 1) rowgemm2_kernel(RowGemmTcParams p, int chunks_per_cta, int nslot) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t full_bar[R2_MAXSLOT], empty_bar[R2_MAXSLOT], acc_full[2], acc_empty[2];

  const int v0 = blockIdx.x * 128;
  const int taps = p.taps;
  const int RA = (128 + taps - 1) | 1;                    // A rows per K block (odd pitch)
  const uint32_t a_bytes = 2u * (R2_KB / 8) * RA * 16;    // hi + lo of one K block
  const uint32_t a_blk = (a_bytes + 127u) & ~127u;
  const uint32_t w_bytes = uint32_t(taps) * 2u * R2_KB * R2_NC * 2;  // one (chunk, K block): all taps, hi + lo
  const int nkb = p.K / R2_KB;
  const int nchunks = (p.N + R2_NC - 1) / R2_NC;
  const int c_begin = blockIdx.y * chunks_per_cta;
  const int c_end = min(nchunks, c_begin + chunks_per_cta);
  if (c_begin >= c_end) return;
  uint8_t* const bufA = smem;
  uint8_t* const ring = smem + size_t(nkb) * a_blk;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
  if (tid == 32) {
    for (int s = 0; s < nslot; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 8);
    }
    tc::mbar_fence_init();
  }

  // ---- stage A once: virtual rows [v0 - pad_left, v0 - pad_left + RA) x K, fp32 -> fp16 hi | lo -----------------
  // item = (K block, row, 8-channel chunk); per K block the layout of rowgemm_tc_kernel's stage:
  // hi [4 chunks][RA rows][8 halfs], then lo -- so the MMA descriptors are the same with a per-block base.
  {
    const int items_kb = (R2_KB / 8) * RA;
    const int total = nkb * items_kb;
    for (int i0 = tid; i0 < total; i0 += R2_LOADS * R2_THREADS) {
      float va[R2_LOADS][8];
      int dst[R2_LOADS];
#pragma unroll
      for (int u = 0; u < R2_LOADS; ++u) {
        const int i = i0 + u * R2_THREADS;
        dst[u] = -1;
#pragma unroll
        for (int e = 0; e < 8; ++e) va[u][e] = 0.f;
        if (i < total) {
          const int kb = i / items_kb, g = i - kb * items_kb;
          const int rr = g / (R2_KB / 8), c8 = g - rr * (R2_KB / 8);
          dst[u] = kb * int(a_blk / 16) + c8 * RA + rr;   // in 16-byte units
          const int v = v0 - p.pad_left + rr;
          const int phys = (v >= 0 && v < p.vrows) ? p.vmap[v] : -1;
          if (phys >= 0) {
            const float* src = p.in + (long long)phys * p.in_stride + kb * R2_KB + c8 * 8;
            if (p.wide) {
              tc::ldg256(src, va[u]);
            } else {
              const float4 a = *reinterpret_cast<const float4*>(src);
              const float4 b = *reinterpret_cast<const float4*>(src + 4);
              va[u][0] = a.x, va[u][1] = a.y, va[u][2] = a.z, va[u][3] = a.w;
              va[u][4] = b.x, va[u][5] = b.y, va[u][6] = b.z, va[u][7] = b.w;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < R2_LOADS; ++u) {
        if (dst[u] < 0) continue;
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 hh = __floats2half2_rn(va[u][2 * e], va[u][2 * e + 1]);
          const float2 back = __half22float2(hh);
          const __half2 ll = __floats2half2_rn((va[u][2 * e] - back.x) * 2048.f, (va[u][2 * e + 1] - back.y) * 2048.f);
          h[e] = *reinterpret_cast<const uint32_t*>(&hh);
          l[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        uint8_t* d = bufA + size_t(dst[u]) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(d + size_t(R2_KB / 8) * RA * 16) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (tc::elect_one()) {
      const size_t blk_elems = size_t(taps) * 2 * R2_KB * R2_NC;
      int it = 0;
      for (int c = c_begin; c < c_end; ++c)
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % nslot;
          tc::mbar_wait(&empty_bar[s], (((it / nslot) & 1) ^ 1));
          tc::mbar_expect_tx(&full_bar[s], w_bytes);
          tc::bulk_g2s(ring + size_t(s) * w_bytes, p.w + (size_t(c) * nkb + kb) * blk_elems, w_bytes, &full_bar[s]);
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, R2_NC, 0);
      const uint32_t abase0 = tc::smem_u32(bufA);
      int it = 0;
      for (int c = c_begin, ci = 0; c < c_end; ++c, ++ci) {
        const int b = ci & 1;
        tc::mbar_wait(&acc_empty[b], (((ci >> 1) & 1) ^ 1));
        tc::fence_after_sync();
        const uint32_t d_main = tmem + uint32_t(b) * 2u * R2_NC, d_corr = d_main + R2_NC;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % nslot;
          tc::mbar_wait(&full_bar[s], ((it / nslot) & 1));
          tc::fence_after_sync();
          const uint32_t abase = abase0 + uint32_t(kb) * a_blk;
          const uint32_t wbase = tc::smem_u32(ring + size_t(s) * w_bytes);
          for (int tap = 0; tap < taps; ++tap) {
#pragma unroll
            for (int ks = 0; ks < R2_KB / 16; ++ks) {
              const uint32_t a_hi = abase + uint32_t((ks * 2) * RA + tap) * 16u;
              const uint32_t a_lo = abase + uint32_t((R2_KB / 8 + ks * 2) * RA + tap) * 16u;
              const uint32_t w_hi = wbase + uint32_t(((tap * 2 + 0) * (R2_KB / 8) + ks * 2) * R2_NC) * 16u;
              const uint32_t w_lo = wbase + uint32_t(((tap * 2 + 1) * (R2_KB / 8) + ks * 2) * R2_NC) * 16u;
              const uint64_t dah = tc::make_desc(a_hi, uint32_t(RA) * 16u, 128u), dal = tc::make_desc(a_lo, uint32_t(RA) * 16u, 128u);
              const uint64_t dwh = tc::make_desc(w_hi, R2_NC * 16u, 128u), dwl = tc::make_desc(w_lo, R2_NC * 16u, 128u);
              const uint32_t first = (kb | tap | ks) ? 1u : 0u;
              tc::mma_f16_ss(d_main, dah, dwh, idesc, first);
              tc::mma_f16_ss(d_corr, dah, dwl, idesc, first);
              tc::mma_f16_ss(d_corr, dal, dwh, idesc, 1u);
            }
          }
          tc::mma_commit(&empty_bar[s]);
        }
        tc::mma_commit(&acc_full[b]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (8 warps) =====================
    const int q = warp & 3;           // TMEM lane quarter this warp may touch
    const int hh = (warp - 2) >> 2;   // column half of the chunk (32 columns)
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
    const int v = v0 + q * 32 + lane;
    const int phys = (v < p.vrows) ? p.vmap[v] : -1;
    const int seg = (p.ubias && phys >= 0) ? p.rowinfo[phys].z : 0;
    for (int c = c_begin, ci = 0; c < c_end; ++c, ++ci) {
      __syncwarp();  // lanes that skipped the previous chunk's stores rejoin before the warp-collective TMEM loads
      const int b = ci & 1;
      tc::mbar_wait(&acc_full[b], ((ci >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t acc = lane_base + uint32_t(b) * 2u * R2_NC + uint32_t(hh * 32);
      float m[32], cr[32];
      tc::tmem_ld16(acc, m);
      tc::tmem_ld16(acc + 16, m + 16);
      tc::tmem_ld16(acc + R2_NC, cr);
      tc::tmem_ld16(acc + R2_NC + 16, cr + 16);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[b]);  // the buffer is free as soon as its TMEM reads are complete
      if (phys < 0) continue;
      const int n0 = c * R2_NC + hh * 32;
      if (n0 >= p.N) continue;
      float* dst = p.out + (long long)phys * p.out_stride + n0;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int n = n0 + e < p.N ? n0 + e : p.N - 1;
        float val = m[e] + cr[e] * (1.0f / 2048.0f);
        if (p.bias) val += p.bias[n];
        if (p.ubias) val += p.ubias[(long long)seg * p.ub_stride + n];
        if (p.act == 1) val = fmaxf(val, 0.f);
        m[e] = val;
      }
      if (n0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) tc::stg256(dst + e, m + e);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (n0 + e < p.N) dst[e] = m[e];
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

size_t rows2_a_bytes(int K, int taps) {
  const int RA = (128 + taps - 1) | 1;
  const size_t a_blk = (size_t(2) * (R2_KB / 8) * RA * 16 + 127) & ~size_t(127);
  return size_t(K / R2_KB) * a_blk;
}
}  // namespace

bool rowgemm2_enabled() {
  const char* e = getenv("M3B200_ROWGEMM_V2");
  return e && *e && *e != '0';
}

// Shapes the resident-A variant takes: the weights must be packed for 64-column chunks (the default), and the
// A rows of all K blocks plus at least two weight slots must fit shared memory.
bool rowgemm2_supported(const RowGemmTcParams& p) {
  if (p.nc != R2_NC || (p.taps != 1 && p.taps != 3) || p.K % R2_KB || p.K < R2_KB) return false;
  const size_t w_bytes = size_t(p.taps) * 2 * R2_KB * R2_NC * 2;
  return rows2_a_bytes(p.K, p.taps) + 2 * w_bytes <= size_t(R2_SMEM_MAX);
}

void launch_rowgemm2(const RowGemmTcParams& p_in, cudaStream_t st) {
  RowGemmTcParams p = p_in;
  if (p.vrows <= 0) return;
  p.wide = (p.in_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.in) & 31u) == 0) ? 1 : 0;
  const size_t a_bytes = rows2_a_bytes(p.K, p.taps);
  const size_t w_bytes = size_t(p.taps) * 2 * R2_KB * R2_NC * 2;
  int nslot = int((size_t(R2_SMEM_MAX) - a_bytes) / w_bytes);
  nslot = std::min(nslot, R2_MAXSLOT);
  if (nslot < 2) throw std::runtime_error("rowgemm2: shape does not fit shared memory");
  static const int n_sm = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  const int row_tiles = (p.vrows + 127) / 128;
  const int nchunks = (p.N + R2_NC - 1) / R2_NC;
  int nsplit = (2 * n_sm + row_tiles - 1) / row_tiles;  // at least two waves of CTAs
  nsplit = std::max(1, std::min(nsplit, nchunks));
  const int cpc = (nchunks + nsplit - 1) / nsplit;
  nsplit = (nchunks + cpc - 1) / cpc;
  // >= 120 KB keeps a single CTA per SM
  const size_t smem = std::max(a_bytes + size_t(nslot) * w_bytes, size_t(120 * 1024));
  ensure_max_dynamic_smem(reinterpret_cast<const void*>(rowgemm2_kernel));
  rowgemm2_kernel<<<dim3(row_tiles, nsplit), R2_THREADS, smem, st>>>(p, cpc, nslot);
  post_launch("rowgemm2_kernel", st);
}

}  // namespace m3

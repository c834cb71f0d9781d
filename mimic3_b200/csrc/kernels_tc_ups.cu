// Polyphase ConvTranspose1d upsampler of the HiFi-GAN generator as an implicit GEMM (sm_100a), second generation.
// Default for the upsamplers since round 2 (1.22 -> 0.87 ms per step at batch 256, bit-identical on hardware:
// profiles/r02a_ab_staged_kernels.txt).  Same math, operand layout, weight packing and CTA roles as conv_tc_kernel with the TC_UPS
// epilogue (kernels_tc_conv.cu):
//     D[t, ph*Cout + co] = sum_{tap} lrelu(x[t - taps + 1 + tap, :]) . W_tap[:, ph*Cout + co];   y[t*u + ph - pad, co] = D + b[co]
// What changes is the epilogue, which is what bounds conv_tc_kernel on this shape: the ncu source page of
// round 1 (profiles/r01e_ncu_full_frame_kernels.md) puts the stall samples on the add that waits for the
// per-column bias load from GLOBAL memory inside a rolled store loop, and the accumulator buffer is only handed
// back to the MMA warp after the last store of a chunk, so the tensor pipe idles behind ~2000-cycle store rounds.
// Here: bias staged once per CTA in shared memory, the 32 columns a thread owns per round fetched with two
// tcgen05.ld in flight, four 256-bit stores per round, the division by Cout hoisted out of the store loop, and
// the accumulator released as soon as its TMEM reads have completed (before the stores are issued).
#include <algorithm>
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

namespace {
constexpr int CT_NT = 2;           // M tiles per CTA
constexpr int CT_R = CT_NT * 128;  // rows per CTA
constexpr int CT_THREADS = 320;
constexpr int CT_SMEM_MAX = 225 * 1024;

template <int FMT>
__global__ void __launch_bounds__(CT_THREADS, 1) ups_tc_kernel(TcConvParams p, int stages, int rows_a, int bias_bytes) {
  using E = tc::Elem<FMT>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], acc_full[2], acc_empty[2];

  const int seg = blockIdx.y;
  const int len_units = p.seg_len[seg];
  const int in_len = len_units * p.in_scale;
  const int rows = in_len + p.rows_extra;
  const int t0 = blockIdx.x * CT_R;
  if (t0 >= rows) return;
  const long long in_base = (long long)p.seg_off[seg] * p.in_scale;
  const int out_len = len_units * p.out_scale;
  const long long out_base = (long long)p.seg_off[seg] * p.out_scale;
  const int CH = p.K / 8;
  const int NC = p.NC;
  const uint32_t stage_bytes = uint32_t(p.K) * NC * 2;
  float* s_bias = reinterpret_cast<float*>(smem);  // [ups_cout]: read once per CTA, not once per store
  uint8_t* bufA = smem + bias_bytes;
  uint8_t* wring = bufA + ((size_t(CH) * rows_a * 16 + 127) & ~size_t(127));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  if (tid == 32) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 8);
    }
    tc::mbar_fence_init();
  }
  for (int i = threadIdx.x; i < p.ups_cout; i += CT_THREADS) s_bias[i] = p.bias[i];
  // ---- stage the A operand: input rows [t0 - halo_l, t0 - halo_l + rows_a) x K ------------------
  // Item = (row, 8-channel chunk), chunk fastest: a warp reads whole rows (coalesced 32 B per lane);
  // rows_a is odd so the 16-byte smem stores of 8 neighbouring chunks fall into distinct banks.
  // Four items per thread are in flight at once (the loop is a chain of L2 round trips otherwise).
  {
    const int halo_l = p.pad_left * p.dil;
    const float slope = p.in_slope;
    const int items = CH * rows_a;
    auto lr = [slope](float v) { return v >= 0.f ? v : slope * v; };
    for (int i0 = tid; i0 < items; i0 += 4 * CT_THREADS) {
      float4 a[4], b[4];
      int dsti[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i0 + u * CT_THREADS;
        a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        dsti[u] = -1;
        if (idx < items) {
          const int rr = idx / CH, c8 = idx - rr * CH;
          dsti[u] = c8 * rows_a + rr;
          const int ti = t0 - halo_l + rr;
          if (ti >= 0 && ti < in_len) {
            const float* src = p.in + (in_base + ti) * (long long)p.in_stride + p.in_coff + c8 * 8;
            a[u] = *reinterpret_cast<const float4*>(src);
            b[u] = *reinterpret_cast<const float4*>(src + 4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (dsti[u] < 0) continue;
        uint4 pk;
        pk.x = E::pack2(lr(a[u].x), lr(a[u].y));
        pk.y = E::pack2(lr(a[u].z), lr(a[u].w));
        pk.z = E::pack2(lr(b[u].x), lr(b[u].y));
        pk.w = E::pack2(lr(b[u].z), lr(b[u].w));
        *reinterpret_cast<uint4*>(bufA + size_t(dsti[u]) * 16) = pk;
      }
    }
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ===================== producer =====================
    if (tc::elect_one()) {
      int it = 0;
      for (int c = 0; c < p.n_chunks; ++c)
        for (int tap = 0; tap < p.taps; ++tap, ++it) {
          const int s = it % stages;
          tc::mbar_wait(&empty_bar[s], (((it / stages) & 1) ^ 1));
          tc::mbar_expect_tx(&full_bar[s], stage_bytes);
          tc::bulk_g2s(wring + size_t(s) * stage_bytes, p.w + (size_t(c) * p.taps + tap) * size_t(p.K) * NC,
                       stage_bytes, &full_bar[s]);
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, NC, FMT);
      const uint32_t abase = tc::smem_u32(bufA);
      int it = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        const int b = c & 1;
        tc::mbar_wait(&acc_empty[b], (((c >> 1) & 1) ^ 1));
        tc::fence_after_sync();
        for (int tap = 0; tap < p.taps; ++tap, ++it) {
          const int s = it % stages;
          tc::mbar_wait(&full_bar[s], ((it / stages) & 1));
          tc::fence_after_sync();
          const uint32_t wbase = tc::smem_u32(wring + size_t(s) * stage_bytes);
          // k-step outer, tile inner: consecutive MMAs alternate between the two accumulators
          // (a dependent accumulate chain costs ~115 cycles per instruction, tools/ubench.py)
          for (int ks = 0; ks < p.K / 16; ++ks) {
            const uint64_t bd = tc::make_desc(wbase + uint32_t(ks * 2 * NC) * 16u, uint32_t(NC) * 16u, 128u);
#pragma unroll
            for (int m = 0; m < CT_NT; ++m) {
              const int arow = m * 128 + tap * p.dil;
              const uint64_t ad =
                  tc::make_desc(abase + uint32_t((ks * 2) * rows_a + arow) * 16u, uint32_t(rows_a) * 16u, 128u);
              tc::mma_f16_ss(tmem + uint32_t(b * CT_NT + m) * NC, ad, bd, idesc, (tap | ks) ? 1u : 0u);
            }
          }
          tc::mma_commit(&empty_bar[s]);
        }
        tc::mma_commit(&acc_full[b]);
      }
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    const int q = warp & 3;            // TMEM lane quarter this warp may touch
    const int hh = (warp - 2) >> 2;    // column half
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
    const int cpt = NC / 2;            // columns per thread and chunk (multiple of 16)
    for (int c = 0; c < p.n_chunks; ++c) {
      const int b = c & 1;
      tc::mbar_wait(&acc_full[b], ((c >> 1) & 1));
      tc::fence_after_sync();
      for (int m = 0; m < CT_NT; ++m) {
        const int t = t0 + m * 128 + q * 32 + lane;
        const bool row_ok = t < rows;
        const uint32_t acc = lane_base + uint32_t(b * CT_NT + m) * NC;
        const bool last_tile = m + 1 == CT_NT;
        for (int cc = 0; cc < cpt; cc += 32) {
          __syncwarp();
          const int j0 = hh * cpt + cc;
          const bool two = cc + 16 < cpt;  // warp-uniform
          float v[32];
          tc::tmem_ld16(acc + j0, v);
          if (two) tc::tmem_ld16(acc + j0 + 16, v + 16);
          tc::tmem_ld_wait();
          if (last_tile && cc + 32 >= cpt) {  // every TMEM read of this buffer is complete: give it back now
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[b]);
          }
          if (!row_ok) continue;
          const int n0 = c * NC + j0;
          const int ngroups = two ? 4 : 2;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            if (g8 >= ngroups) break;
            const int n = n0 + g8 * 8;
            if (n >= p.N) break;
            const int phase = n / p.ups_cout, co = n - phase * p.ups_cout;
            const int po = t * p.ups_u + phase - p.ups_pad;
            if (po < 0 || po >= out_len) continue;
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + co);
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + co + 4);
            float o[8];
            o[0] = v[g8 * 8 + 0] + b0.x;
            o[1] = v[g8 * 8 + 1] + b0.y;
            o[2] = v[g8 * 8 + 2] + b0.z;
            o[3] = v[g8 * 8 + 3] + b0.w;
            o[4] = v[g8 * 8 + 4] + b1.x;
            o[5] = v[g8 * 8 + 5] + b1.y;
            o[6] = v[g8 * 8 + 6] + b1.z;
            o[7] = v[g8 * 8 + 7] + b1.w;
            tc::stg256(p.out + (out_base + po) * p.out_stride + co, o);
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}
}  // namespace

// Default since round 2 (upsample stage 1.22 -> 0.87 ms, bit-identical; profiles/r02a_ab_staged_kernels.txt).
// M3B200_UPS_V1=1 selects the generic conv_tc_kernel TC_UPS epilogue again (A/B and the bit-identity test).
bool ups_tc_enabled() {
  const char* e = getenv("M3B200_UPS_V1");
  return !(e && *e && *e != '0');
}

// Same shapes as conv_tc_supported plus the alignment the 256-bit stores and the float4 bias reads need.
bool ups_tc_supported(const TcConvParams& p) {
  if (p.epi != TC_UPS || !p.bias || p.ubias) return false;
  if (p.ups_cout % 8 || p.ups_cout > 1024 || p.out_stride % 8 || p.out_coff) return false;
  if ((reinterpret_cast<uintptr_t>(p.out) & 31u) != 0) return false;
  if (p.NC % 32 || p.NC < 32 || p.NC > 128 || p.K % 16 || p.K < 16 || p.K > 512) return false;
  const size_t bias_bytes = (size_t(p.ups_cout) * 4 + 127) & ~size_t(127);
  const size_t a_bytes = (size_t(p.K / 8) * ((CT_R + (p.taps - 1) * p.dil) | 1) * 16 + 127) & ~size_t(127);
  return bias_bytes + a_bytes + 2 * size_t(p.K) * p.NC * 2 <= size_t(CT_SMEM_MAX);
}

void launch_ups_tc(const TcConvParams& p, int fmt, int n_seg, int max_seg_len, cudaStream_t st) {
  const int rows = max_seg_len * p.in_scale + p.rows_extra;
  if (rows <= 0 || n_seg <= 0) return;
  const int rows_a = (CT_R + (p.taps - 1) * p.dil) | 1;  // odd row pitch: conflict-free chunk-major smem stores
  const size_t bias_bytes = (size_t(p.ups_cout) * 4 + 127) & ~size_t(127);
  const size_t a_bytes = (size_t(p.K / 8) * rows_a * 16 + 127) & ~size_t(127);
  const size_t stage = size_t(p.K) * p.NC * 2;
  int stages = int((size_t(CT_SMEM_MAX) - bias_bytes - a_bytes) / stage);
  stages = std::min(4, stages);
  if (stages < 2) throw std::runtime_error("ups_tc: shape does not fit shared memory");
  // >= 120 KB keeps a single CTA per SM (each CTA owns all 512 TMEM columns)
  const size_t smem = std::max(bias_bytes + a_bytes + stages * stage, size_t(120 * 1024));
  dim3 grid((rows + CT_R - 1) / CT_R, n_seg);
  ensure_max_dynamic_smem(fmt ? reinterpret_cast<const void*>(ups_tc_kernel<1>) : reinterpret_cast<const void*>(ups_tc_kernel<0>));
  if (fmt) ups_tc_kernel<1><<<grid, CT_THREADS, smem, st>>>(p, stages, rows_a, int(bias_bytes));
  else ups_tc_kernel<0><<<grid, CT_THREADS, smem, st>>>(p, stages, rows_a, int(bias_bytes));
  post_launch("ups_tc_kernel", st);
}

}  // namespace m3

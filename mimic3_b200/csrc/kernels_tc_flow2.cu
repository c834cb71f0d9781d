// Fused residual-coupling layer of the VITS flow, reverse direction, second generation:
//     h = pre(x0);  for i < nl: a = in_i(h) (k=5) + cond_i;  act = tanh(a[:H]) * sigmoid(a[H:]);
//                               rs = res_skip_i(act);  h += rs[:H];  skip += rs[H:]   (last: skip += rs)
//     x1 -= post(skip)
// Same window per CTA, same operand layout and the same TMEM-resident fp32 h as flow_tc_kernel (kernels_tc_flow.cu).
// What changed, and why (per-role cycle counters, profiles/r02h_role_cycles.md): in the first kernel the ONE issuing
// thread was busy 90 % of a window -- 1734 N=64 MMAs in 145 weight stages, each stage costing it a wait, a fence and a
// tcgen05.commit (~320 cycles, m3_selftest 5xx/6xx) on top of ~50 cycles per MMA -- while the tensor pipe needed
// only 48 % of the window.  So this kernel issues FEWER, WIDER MMAs in fewer stages:
//   * post(skip) is linear, so skip is never formed: m = post(skip) = sum_i (W_post W_skip_i) act_i + const.  The host
//     pre-multiplies W'_i = W_post . W_skip_i (96 x 192 per layer, fp64 on the host, Flip folded in); m accumulates
//     directly in 96 TMEM columns.  No 192-column skip region, no skip -> fp16 epilogue, no post stage;
//   * gate chunks are 48 channels wide: N = 96 MMAs (56 cycles for 1.5 x the columns of an N = 64 one), 4 chunks per
//     layer, ping-pong between two 96-column accumulators;
//   * the residual update h += W_res act is ONE N = 192 MMA per k-step (two 6-MMA stages, K halves), the m update ONE
//     N = 96 MMA per k-step: 1062 MMAs in 93 stages per window instead of 1734 in 145;
//   * every weight block is 36 864 B ([K/8][N][8] with K x N = 192 x 96 or 96 x 192): one slot size, 3-slot ring;
//   * the two epilogue warps of a TMEM lane quarter alternate whole gate chunks (chunk c belongs to accumulator and
//     warp group c & 1) instead of splitting every chunk's columns: all tcgen05.ld are 16-column aligned.
// TMEM (512 columns): H = [0, 192), M = [192, 288), ACC0 = [288, 384), ACC1 = [384, 480).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <type_traits>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

namespace {
constexpr int F2_THREADS = 320;  // warp 0 weight producer, warp 1 MMA issuer, warps 2-9 epilogue
constexpr int F2_STAGES = 3;
constexpr int F2_H = 192, F2_HALF = 96, F2_GC = 48;     // WN hidden channels, coupling half, gated channels per chunk
constexpr uint32_t F2_BLOCK = 192u * 96u * 2u;          // bytes of every weight block
constexpr int F2_ROWS_H = 133, F2_ROWS_A = 129;         // operand row pitches (odd: conflict-free chunk-major stores)
constexpr uint32_t F2_TH = 0, F2_TM = 192, F2_TACC = 288;
}  // namespace

template <int FMT>
__global__ void __launch_bounds__(F2_THREADS, 1) flow2_tc_kernel(FlowTcParams p) {
  using E = tc::Elem<FMT>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t full_bar[F2_STAGES], empty_bar[F2_STAGES], acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t h_full, h_ready, act_ready;

  constexpr int Hc = F2_H, half = F2_HALF;
  const int nl = p.nl;
  const int HALO = 2 * nl;  // frames lost per side over the nl k=5 layers
  const int seg = blockIdx.y;
  const int L = p.seg_len[seg];
  const int o0 = blockIdx.x * (128 - 2 * HALO);
  if (o0 >= L) return;
  const long long base = p.seg_off[seg];
  const int w0 = o0 - HALO;
  constexpr int KH = Hc / 8, KX = half / 8;
  uint8_t* bufH = smem;
  uint8_t* bufA = bufH + ((size_t(KH) * F2_ROWS_H * 16 + 127) & ~size_t(127));  // act / x0 operand
  uint8_t* wring = bufA + ((size_t(KH) * F2_ROWS_A * 16 + 127) & ~size_t(127));
  float* sb = reinterpret_cast<float*>(wring + size_t(F2_STAGES) * F2_BLOCK);
  // sb layout: in_bias[nl][2 Hc] (+ this utterance's conditioning) | cum_bias[nl][Hc] | m_bias[half]
  float* s_inb = sb;
  float* s_cb = s_inb + nl * 2 * Hc;
  float* s_mb = s_cb + nl * Hc;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  if (tid == 32) {
    for (int s = 0; s < F2_STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 4);  // the four warps (one per lane quarter) of the group that owns accumulator b
    }
    tc::mbar_init(&h_full, 1);
    tc::mbar_init(&h_ready, 8);
    tc::mbar_init(&act_ready, 8);
    tc::mbar_fence_init();
  }
  for (int i = tid; i < nl * 2 * Hc; i += F2_THREADS)
    s_inb[i] = p.in_bias[i] + (p.cond ? p.cond[(long long)seg * p.cond_stride + i] : 0.f);
  for (int i = tid; i < nl * Hc; i += F2_THREADS) s_cb[i] = p.cum_bias[i];
  for (int i = tid; i < half; i += F2_THREADS) s_mb[i] = p.post_bias[i];
  {  // x0 window (no halo: pre is 1x1) -> bufA as the 16-bit A operand, coalesced, 4 items in flight
    const int items = KX * 128;
    for (int i0 = tid; i0 < items; i0 += 4 * F2_THREADS) {
      float4 a[4], b[4];
      int dsti[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i0 + u * F2_THREADS;
        a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        dsti[u] = -1;
        if (idx < items) {
          const int rr = idx / KX, c8 = idx - rr * KX;
          dsti[u] = c8 * F2_ROWS_A + rr;
          const int g = w0 + rr;
          if (g >= 0 && g < L) {
            const float* src = p.z + (base + g) * (long long)p.z_stride + p.x0_coff + c8 * 8;
            a[u] = *reinterpret_cast<const float4*>(src);
            b[u] = *reinterpret_cast<const float4*>(src + 4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (dsti[u] < 0) continue;
        uint4 pk;
        pk.x = E::pack2(a[u].x, a[u].y);
        pk.y = E::pack2(a[u].z, a[u].w);
        pk.z = E::pack2(b[u].x, b[u].y);
        pk.w = E::pack2(b[u].z, b[u].w);
        *reinterpret_cast<uint4*>(bufA + size_t(dsti[u]) * 16) = pk;
      }
    }
    // halo rows of bufH that no epilogue ever writes must still be finite
    for (int i = tid; i < KH * F2_ROWS_H; i += F2_THREADS) {
      const int rr = i % F2_ROWS_H;
      if (rr < 2 || rr >= 130) *reinterpret_cast<uint4*>(bufH + size_t(i) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  constexpr int n_gate_chunks = Hc / F2_GC;  // 4

  if (warp == 0) {
    // ===================== producer: walk the packed weight stream (all blocks the same size) =====================
    if (tc::elect_one()) {
      const int n_blocks = 1 + nl * (n_gate_chunks * 5 + 1) + (nl - 1) * 2;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w);
      for (int it = 0; it < n_blocks; ++it) {
        const int s = it % F2_STAGES;
        tc::mbar_wait(&empty_bar[s], (((it / F2_STAGES) & 1) ^ 1));
        tc::mbar_expect_tx(&full_bar[s], F2_BLOCK);
        tc::bulk_g2s(wring + size_t(s) * F2_BLOCK, src, F2_BLOCK, &full_bar[s]);
        src += F2_BLOCK;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      const uint32_t id96 = tc::make_idesc(128, 96, FMT), id192 = tc::make_idesc(128, 192, FMT);
      const uint32_t aH = tc::smem_u32(bufH), aA = tc::smem_u32(bufA);
      // descriptors of a stage differ only in their 14-bit start-address field: templates + additive constants
      const uint32_t d_hi = uint32_t(tc::make_desc(0u, 16u, 128u) >> 32);  // SBO + version bit (same for every operand)
      const uint32_t a_loH = uint32_t(tc::make_desc(aH, uint32_t(F2_ROWS_H) * 16u, 128u));
      const uint32_t a_loA = uint32_t(tc::make_desc(aA, uint32_t(F2_ROWS_A) * 16u, 128u));
      uint32_t b96[F2_STAGES], b192[F2_STAGES];
#pragma unroll
      for (int s = 0; s < F2_STAGES; ++s) {
        const uint32_t wb = tc::smem_u32(wring + size_t(s) * F2_BLOCK);
        b96[s] = uint32_t(tc::make_desc(wb, 96u * 16u, 128u));
        b192[s] = uint32_t(tc::make_desc(wb, 192u * 16u, 128u));
      }
      int slot = 0;
      uint32_t slot_par = 0;
      // one weight block: KS k-steps of one MMA each (N = NCOL columns) into `dst`
      auto stage_mma = [&](auto ks_tag, auto n_tag, uint32_t a_lo, uint32_t a_kstep, uint32_t dst, bool first) {
        constexpr int KS = decltype(ks_tag)::value, NCOL = decltype(n_tag)::value;
        tc::mbar_wait(&full_bar[slot], slot_par);
        tc::fence_after_sync();
        const uint32_t bl = NCOL == 96 ? (slot == 0 ? b96[0] : (slot == 1 ? b96[1] : b96[2]))
                                       : (slot == 0 ? b192[0] : (slot == 1 ? b192[1] : b192[2]));
        const uint32_t idesc = NCOL == 96 ? id96 : id192;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint64_t ad = (uint64_t(d_hi) << 32) | uint64_t(a_lo + uint32_t(ks) * a_kstep);
          const uint64_t bd = (uint64_t(d_hi) << 32) | uint64_t(bl + uint32_t(ks * 2 * NCOL));
          tc::mma_f16_ss(tmem + dst, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
        }
        tc::mma_commit(&empty_bar[slot]);
        if (++slot == F2_STAGES) {
          slot = 0;
          slot_par ^= 1u;
        }
      };
      using K6 = std::integral_constant<int, 6>;
      using K12 = std::integral_constant<int, 12>;
      using N96 = std::integral_constant<int, 96>;
      using N192 = std::integral_constant<int, 192>;
      constexpr uint32_t kstepH = uint32_t(2 * F2_ROWS_H), kstepA = uint32_t(2 * F2_ROWS_A);
      uint32_t ph_hready = 0, ph_act = 0;
      int acc_it = 0;
      // ---- pre: H = x0 . Wpre (K = 96, N = 192) ----
      stage_mma(K6{}, N192{}, a_loA, kstepA, F2_TH, true);
      tc::mma_commit(&h_full);
      for (int i = 0; i < nl; ++i) {
        // ---- in_layer i: 48-channel gate chunks into the accumulator ping-pong; needs the fp16 h of this layer ----
        tc::mbar_wait(&h_ready, ph_hready);
        ph_hready ^= 1u;
        tc::fence_after_sync();
        for (int c = 0; c < n_gate_chunks; ++c, ++acc_it) {
          const int b = acc_it & 1;
          tc::mbar_wait(&acc_empty[b], (((acc_it >> 1) & 1) ^ 1));
          tc::fence_after_sync();
          for (int tap = 0; tap < 5; ++tap)
            stage_mma(K12{}, N96{}, a_loH + uint32_t(tap), kstepH, F2_TACC + uint32_t(b) * 96u, tap == 0);
          tc::mma_commit(&acc_full[b]);
        }
        // ---- h += W_res act (two K halves, N = 192) and m += W'_i act (N = 96): need the whole fp16 act ----
        tc::mbar_wait(&act_ready, ph_act);
        ph_act ^= 1u;
        tc::fence_after_sync();
        if (i < nl - 1) {
          stage_mma(K6{}, N192{}, a_loA, kstepA, F2_TH, false);
          stage_mma(K6{}, N192{}, a_loA + uint32_t(12 * F2_ROWS_A), kstepA, F2_TH, false);
        }
        stage_mma(K12{}, N96{}, a_loA, kstepA, F2_TM, i == 0);
        tc::mma_commit(&h_full);
      }
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;  // column half for H / M, owner of the gate chunks with (c & 1) == hh
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
    const int r = q * 32 + lane;  // window row of this thread
    const int g = w0 + r;         // frame inside the utterance
    const bool inside = g >= 0 && g < L;
    uint32_t ph_hfull = 0;

    for (int i = 0; i < nl; ++i) {
      // ---- h of layer i (pre or previous residual update) + cumulated bias -> masked 16-bit rows of bufH ----
      tc::mbar_wait(&h_full, ph_hfull);
      ph_hfull ^= 1u;
      tc::fence_after_sync();
      {
        const float* bias = s_cb + i * Hc;
        for (int cc = 0; cc < Hc / 2; cc += 32) {
          __syncwarp();
          float v0[16], v1[16];
          const int col = hh * (Hc / 2) + cc;
          tc::tmem_ld16(lane_base + F2_TH + col, v0);
          tc::tmem_ld16(lane_base + F2_TH + col + 16, v1);
          tc::tmem_ld_wait();
#pragma unroll
          for (int hpart = 0; hpart < 2; ++hpart) {
            const float* v = hpart ? v1 : v0;
            const int c0 = col + hpart * 16;
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              pk[e] = inside ? E::pack2(v[2 * e] + bias[c0 + 2 * e], v[2 * e + 1] + bias[c0 + 2 * e + 1]) : 0u;
            uint8_t* d = bufH + (size_t(c0 / 8) * F2_ROWS_H + r + 2) * 16;
            *reinterpret_cast<uint4*>(d) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(d + size_t(F2_ROWS_H) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncwarp();
      }
      if (lane == 0) tc::mbar_arrive(&h_ready);
      // ---- gate chunks of this warp group -> bufA (48 channels = 3 x 16 "a" columns + 3 x 16 "b" columns) ----
      for (int c = hh; c < n_gate_chunks; c += 2) {
        const int k = (i * n_gate_chunks + c) >> 1;  // how many chunks accumulator hh has held before this one
        tc::mbar_wait(&acc_full[hh], uint32_t(k & 1));
        tc::fence_after_sync();
        const uint32_t acc = lane_base + F2_TACC + uint32_t(hh) * 96u;
        const float* ba = s_inb + i * 2 * Hc + c * F2_GC;
        const float* bb = ba + Hc;
#pragma unroll
        for (int part = 0; part < 3; ++part) {  // 16 channels at a time: registers stay within the 320-thread budget
          float va[16], vb[16];
          tc::tmem_ld16(acc + uint32_t(16 * part), va);
          tc::tmem_ld16(acc + uint32_t(F2_GC + 16 * part), vb);
          tc::tmem_ld_wait();
          if (part == 2) {  // the accumulator is in registers now
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[hh]);
          }
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int ch = 16 * part + 2 * e;
            pk[e] = E::pack2(tc::gate_tanh_sigmoid(va[2 * e] + ba[ch], vb[2 * e] + bb[ch]),
                             tc::gate_tanh_sigmoid(va[2 * e + 1] + ba[ch + 1], vb[2 * e + 1] + bb[ch + 1]));
          }
          const int ch0 = c * F2_GC + 16 * part;
          uint8_t* d = bufA + (size_t(ch0 / 8) * F2_ROWS_A + r) * 16;
          *reinterpret_cast<uint4*>(d) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(d + size_t(F2_ROWS_A) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&act_ready);
    }
    // ---- x1 -= m + bias (m = post(skip), accumulated layer by layer in TMEM) ----
    tc::mbar_wait(&h_full, ph_hfull);
    tc::fence_after_sync();
    const bool store = inside && r >= HALO && r < 128 - HALO;
    {
      const int n0 = hh * (half / 2);  // 48 columns per thread
      float v[3][16];
#pragma unroll
      for (int k = 0; k < 3; ++k) tc::tmem_ld16(lane_base + F2_TM + uint32_t(n0 + 16 * k), v[k]);
      tc::tmem_ld_wait();
      if (store) {
        float* row = p.z + (base + g) * (long long)p.z_stride + p.x1_coff + n0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float4 cur[4];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) cur[e4] = *reinterpret_cast<const float4*>(row + 16 * k + 4 * e4);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const int n = n0 + 16 * k + 4 * e4;
            float4 o = cur[e4];
            o.x -= v[k][4 * e4] + s_mb[n];
            o.y -= v[k][4 * e4 + 1] + s_mb[n + 1];
            o.z -= v[k][4 * e4 + 2] + s_mb[n + 2];
            o.w -= v[k][4 * e4 + 3] + s_mb[n + 3];
            *reinterpret_cast<float4*>(row + 16 * k + 4 * e4) = o;
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

size_t flow2_tc_smem_bytes(int nl) {
  const size_t bufH = (size_t(F2_H / 8) * F2_ROWS_H * 16 + 127) & ~size_t(127);
  const size_t bufA = (size_t(F2_H / 8) * F2_ROWS_A * 16 + 127) & ~size_t(127);
  const size_t ring = size_t(F2_STAGES) * F2_BLOCK;
  const size_t bias = sizeof(float) * (size_t(nl) * 2 * F2_H + size_t(nl) * F2_H + F2_HALF);
  return bufH + bufA + ring + bias + 64;
}

bool flow2_tc_supported(int Hc, int half, int nl, int kernel) {
  if (kernel != 5 || nl < 1 || nl > 8) return false;
  if (Hc != F2_H || half != F2_HALF) return false;  // every shipped voice: 192 hidden channels, 96-channel halves
  if (128 - 4 * nl < 32) return false;
  return flow2_tc_smem_bytes(nl) + 1024 <= size_t(227 * 1024);
}

size_t flow2_tc_weight_elems(int nl) {  // 16-bit elements of one coupling layer's packed stream
  return size_t(F2_BLOCK / 2) * size_t(1 + nl * 21 + (nl - 1) * 2);
}

void launch_flow2_tc(const FlowTcParams& p, int fmt, int n_seg, int max_len, cudaStream_t st) {
  if (n_seg <= 0 || max_len <= 0) return;
  const size_t smem = flow2_tc_smem_bytes(p.nl);
  ensure_max_dynamic_smem(fmt ? reinterpret_cast<const void*>(flow2_tc_kernel<1>) : reinterpret_cast<const void*>(flow2_tc_kernel<0>));
  const int stride = 128 - 4 * p.nl;
  dim3 grid((max_len + stride - 1) / stride, n_seg);
  if (fmt) flow2_tc_kernel<1><<<grid, F2_THREADS, smem, st>>>(p);
  else flow2_tc_kernel<0><<<grid, F2_THREADS, smem, st>>>(p);
  post_launch("flow2_tc_kernel", st);
}

}  // namespace m3

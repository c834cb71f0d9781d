// Fused residual-coupling layer of the VITS flow, reverse direction:
//     h = pre(x0);  for i < nl: a = in_i(h) (k=5) + cond_i;  act = tanh(a[:H]) * sigmoid(a[H:]);
//                               rs = res_skip_i(act);  h += rs[:H];  skip += rs[H:]   (last: skip += rs)
//     x1 -= post(skip)
// One CTA owns a window of 128 frames of one utterance for the WHOLE coupling layer; nothing
// but x0 (read) and x1 (read-modify-write) touches HBM:
//   * h and skip live in TMEM as fp32 (192 + 192 columns); the res/skip 1x1 convs are
//     tcgen05.mma's that accumulate straight into those regions (the residual adds are free);
//   * the 16-bit A operands (h, act, skip) live in shared memory and are rewritten by the
//     epilogue warps after each stage; a k=5 tap is a descriptor start-address shift;
//   * gate chunks (32 channels: 32 "a" + 32 "b" columns) ping-pong between two 64-column TMEM
//     accumulators so tanh*sigmoid of chunk c overlaps the MMAs of chunk c+1;
//   * weights stream through a 3-stage ring of 1-D bulk copies in one fixed schedule order
//     (host packs them in exactly that order);
//   * each WN layer shrinks the exact region by 2 frames per side: windows advance by 128 - 4*nl.
// The channel Flip between coupling layers is folded into the packing of pre/post weights.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <type_traits>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

constexpr int FL_THREADS = 320;
constexpr int FL_NC = 64;       // columns per weight stage / MMA
constexpr int FL_STAGES = 3;

template <int FMT>
__global__ void __launch_bounds__(FL_THREADS, 1) flow_tc_kernel(FlowTcParams p) {
  using E = tc::Elem<FMT>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t full_bar[FL_STAGES], empty_bar[FL_STAGES], acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t h_full, h_ready, act_ready;

  const int Hc = p.Hc;          // WN hidden channels (192)
  const int half = p.half;      // coupling half (96)
  const int nl = p.nl;
  const int HALO = 2 * nl;      // frames lost per side over the nl k=5 layers ((k-1)/2 = 2 each)
  const int seg = blockIdx.y;
  const int L = p.seg_len[seg];
  const int o0 = blockIdx.x * (128 - 2 * HALO);
  if (o0 >= L) return;
  const long long base = p.seg_off[seg];
  const int w0 = o0 - HALO;
  const int KH = Hc / 8, KX = half / 8;
  const int ROWS_H = 133;       // 128 + 2*2 taps halo, odd pitch
  const int ROWS_A = 129;
  uint8_t* bufH = smem;
  uint8_t* bufA = bufH + ((size_t(KH) * ROWS_H * 16 + 127) & ~size_t(127));   // act / skip / x0 operand
  uint8_t* wring = bufA + ((size_t(KH) * ROWS_A * 16 + 127) & ~size_t(127));
  const uint32_t slot_bytes = uint32_t(Hc) * FL_NC * 2;
  float* sb = reinterpret_cast<float*>(wring + size_t(FL_STAGES) * slot_bytes);
  // sb layout: in_bias[nl][2Hc] (+cond) | cb[nl][Hc] | skipb[Hc] | postb[half]
  float* s_inb = sb;
  float* s_cb = s_inb + nl * 2 * Hc;
  float* s_skb = s_cb + nl * Hc;
  float* s_pob = s_skb + Hc;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef M3B200_KERNEL_PROFILE  // per-role cycle counters (M3B200_FLOW_PROFILE=1); compiled out by default (+5 % kernel time when present)
  const bool prof = p.prof != nullptr;
#else
  constexpr bool prof = false;
#endif
  const long long k_start = prof ? clock64() : 0;
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  if (tid == 32) {
    for (int s = 0; s < FL_STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 8);
    }
    tc::mbar_init(&h_full, 1);
    tc::mbar_init(&h_ready, 8);
    tc::mbar_init(&act_ready, 8);
    tc::mbar_fence_init();
  }
  // biases (+ this utterance's conditioning) -> smem
  for (int i = tid; i < nl * 2 * Hc; i += FL_THREADS)
    s_inb[i] = p.in_bias[i] + (p.cond ? p.cond[(long long)seg * p.cond_stride + i] : 0.f);
  for (int i = tid; i < nl * Hc; i += FL_THREADS) s_cb[i] = p.cum_bias[i];
  for (int i = tid; i < Hc; i += FL_THREADS) s_skb[i] = p.skip_bias[i];
  for (int i = tid; i < half; i += FL_THREADS) s_pob[i] = p.post_bias[i];
  // x0 window (no halo: pre is 1x1) -> bufA as the 16-bit A operand, coalesced, 4 items in flight
  {
    const int items = KX * 128;
    for (int i0 = tid; i0 < items; i0 += 4 * FL_THREADS) {
      float4 a[4], b[4];
      int dsti[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i0 + u * FL_THREADS;
        a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        dsti[u] = -1;
        if (idx < items) {
          const int rr = idx / KX, c8 = idx - rr * KX;
          dsti[u] = c8 * ROWS_A + rr;
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
    for (int i = tid; i < KH * ROWS_H; i += FL_THREADS) {
      const int rr = i % ROWS_H;
      if (rr < 2 || rr >= 130) *reinterpret_cast<uint4*>(bufH + size_t(i) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const long long k_pro = prof ? clock64() : 0;
  auto timed_wait = [&](uint64_t* bar, uint32_t parity, long long& acc) {
    if (prof && !tc::mbar_test(bar, parity)) {  // only waits that actually block are timed
      const long long t = clock64();
      tc::mbar_wait(bar, parity);
      acc += clock64() - t;
    } else {
      tc::mbar_wait(bar, parity);
    }
  };
  const uint32_t T_H = 0, T_SKIP = uint32_t(Hc), T_ACC = 2u * uint32_t(Hc);  // ACC0 / ACC1: +0 / +64
  const int n_gate_chunks = Hc / 32;    // 32 gated channels per chunk (64 MMA columns)
  const int n_h_chunks = Hc / FL_NC;    // 64-column chunks of an Hc-wide output
  const int n_post_chunks = (half + FL_NC - 1) / FL_NC;

  if (warp == 0) {
    // ===================== producer: walk the packed weight stream =====================
    if (tc::elect_one()) {
      const uint16_t* src = p.w;
      int it = 0;
      long long w_empty = 0;
      auto push = [&](int K) {
        const int s = it % FL_STAGES;
        const uint32_t bytes = uint32_t(K) * FL_NC * 2;
        timed_wait(&empty_bar[s], (((it / FL_STAGES) & 1) ^ 1), w_empty);
        tc::mbar_expect_tx(&full_bar[s], bytes);
        tc::bulk_g2s(wring + size_t(s) * slot_bytes, src, bytes, &full_bar[s]);
        src += size_t(K) * FL_NC;
        ++it;
      };
      for (int c = 0; c < n_h_chunks; ++c) push(half);                 // pre
      for (int i = 0; i < nl; ++i) {
        for (int c = 0; c < n_gate_chunks; ++c)
          for (int tap = 0; tap < 5; ++tap) push(Hc);                  // in_layer i
        const int nrs = (i < nl - 1 ? 2 : 1) * n_h_chunks;
        for (int c = 0; c < nrs; ++c) push(Hc);                        // res_skip i
      }
      for (int c = 0; c < n_post_chunks; ++c) push(Hc);                // post
      if (prof) atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 8), (unsigned long long)w_empty);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, FL_NC, FMT);
      const uint32_t aH = tc::smem_u32(bufH), aA = tc::smem_u32(bufA);
      // one weight stage = one (chunk[, tap]) block: K/16 MMAs into `dst`.  The descriptors of a stage differ only
      // in their 14-bit start-address field, so the k-steps are fully unrolled around two descriptor templates:
      // one independent 32-bit add per operand and MMA instead of a shift/mask/or chain on the uniform datapath
      // (the rolled loop issued one MMA per ~100 cycles -- ncu: 1734 MMAs in 176 k cycles per window -- while
      // the tensor pipe needs 48.6).
      const uint32_t a_hiH = uint32_t(tc::make_desc(0u, uint32_t(ROWS_H) * 16u, 128u) >> 32);  // same for both buffers:
      const uint32_t b_hi = uint32_t(tc::make_desc(0u, uint32_t(FL_NC) * 16u, 128u) >> 32);    // SBO + version bit
      const uint32_t a_loH = uint32_t(tc::make_desc(aH, uint32_t(ROWS_H) * 16u, 128u));
      const uint32_t a_loA = uint32_t(tc::make_desc(aA, uint32_t(ROWS_A) * 16u, 128u));
      uint32_t b_lo[FL_STAGES];
#pragma unroll
      for (int s = 0; s < FL_STAGES; ++s)
        b_lo[s] = uint32_t(tc::make_desc(tc::smem_u32(wring + size_t(s) * slot_bytes), uint32_t(FL_NC) * 16u, 128u));
      int slot = 0;
      uint32_t slot_par = 0;
      long long c_full = 0, c_acc = 0, c_h = 0, c_act = 0;
      const long long c_start = prof ? clock64() : 0;
      auto stage_mma = [&](auto ks_tag, uint32_t a_lo, uint32_t a_kstep, uint32_t dst, bool first) {
        constexpr int KS = decltype(ks_tag)::value;
        timed_wait(&full_bar[slot], slot_par, c_full);
        tc::fence_after_sync();  // (measured free: r02l A/B with and without it)
        const uint32_t bl = slot == 0 ? b_lo[0] : (slot == 1 ? b_lo[1] : b_lo[2]);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint64_t ad = (uint64_t(a_hiH) << 32) | uint64_t(a_lo + uint32_t(ks) * a_kstep);
          const uint64_t bd = (uint64_t(b_hi) << 32) | uint64_t(bl + uint32_t(ks * 2 * FL_NC));
          tc::mma_f16_ss(tmem + dst, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
        }
        tc::mma_commit(&empty_bar[slot]);
        if (++slot == FL_STAGES) {
          slot = 0;
          slot_par ^= 1u;
        }
      };
      using KX6 = std::integral_constant<int, 6>;
      using KH12 = std::integral_constant<int, 12>;
      const uint32_t kstepH = uint32_t(2 * ROWS_H), kstepA = uint32_t(2 * ROWS_A);
      uint32_t ph_hready = 0, ph_act = 0;
      int acc_it = 0;
      // ---- pre: H = x0 . Wpre ----
      for (int c = 0; c < n_h_chunks; ++c) stage_mma(KX6{}, a_loA, kstepA, T_H + c * FL_NC, true);
      tc::mma_commit(&h_full);
      for (int i = 0; i < nl; ++i) {
        // ---- in_layer i: gate chunks into ACC ping-pong; needs the fp16 h of this layer ----
        timed_wait(&h_ready, ph_hready, c_h);
        ph_hready ^= 1u;
        tc::fence_after_sync();
        for (int c = 0; c < n_gate_chunks; ++c, ++acc_it) {
          const int b = acc_it & 1;
          timed_wait(&acc_empty[b], (((acc_it >> 1) & 1) ^ 1), c_acc);
          tc::fence_after_sync();
          for (int tap = 0; tap < 5; ++tap) stage_mma(KH12{}, a_loH + uint32_t(tap), kstepH, T_ACC + b * FL_NC, tap == 0);
          tc::mma_commit(&acc_full[b]);
        }
        // ---- res_skip i: accumulate straight into H / SKIP; needs the whole fp16 act ----
        timed_wait(&act_ready, ph_act, c_act);
        ph_act ^= 1u;
        tc::fence_after_sync();
        if (i < nl - 1)
          for (int c = 0; c < n_h_chunks; ++c) stage_mma(KH12{}, a_loA, kstepA, T_H + c * FL_NC, false);
        for (int c = 0; c < n_h_chunks; ++c) stage_mma(KH12{}, a_loA, kstepA, T_SKIP + c * FL_NC, i == 0);
        tc::mma_commit(&h_full);
      }
      // ---- post: m = skip . Wpost (fp16 skip staged in bufA by the epilogue) ----
      timed_wait(&act_ready, ph_act, c_act);
      ph_act ^= 1u;
      tc::fence_after_sync();
      for (int c = 0; c < n_post_chunks; ++c, ++acc_it) {
        const int b = acc_it & 1;
        tc::mbar_wait(&acc_empty[b], (((acc_it >> 1) & 1) ^ 1));
        tc::fence_after_sync();
        stage_mma(KH12{}, a_loA, kstepA, T_ACC + b * FL_NC, true);
        tc::mma_commit(&acc_full[b]);
      }
      if (prof) {
        unsigned long long* q = reinterpret_cast<unsigned long long*>(p.prof);
        atomicAdd(q + 0, (unsigned long long)(clock64() - c_start));
        atomicAdd(q + 1, (unsigned long long)c_full);
        atomicAdd(q + 2, (unsigned long long)c_acc);
        atomicAdd(q + 3, (unsigned long long)c_h);
        atomicAdd(q + 4, (unsigned long long)c_act);
        atomicAdd(q + 5, 1ull);
        atomicAdd(q + 6, (unsigned long long)(k_pro - k_start));
      }
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
    const int r = q * 32 + lane;       // window row of this thread
    const int g = w0 + r;              // frame inside the utterance
    const bool inside = g >= 0 && g < L;
    uint32_t ph_hfull = 0;
    int acc_it = 0;
    long long e_h = 0, e_acc = 0;
    const int hcols = Hc / 2;          // columns of H / SKIP per thread

    // TMEM region (Hc fp32 columns) + bias -> masked 16-bit A operand rows in `dst`
    auto region_to_smem = [&](uint32_t region, const float* bias, uint8_t* dst, int pitch, int row_off) {
      // (two loads per tcgen05.wait::ld; putting all six of a thread's loads in flight was tried on the decoder
      // kernels and made them slower, r02k)
      for (int cc = 0; cc < hcols; cc += 32) {
        __syncwarp();
        float v0[16], v1[16];
        const int col = hh * hcols + cc;
        tc::tmem_ld16(lane_base + region + col, v0);
        tc::tmem_ld16(lane_base + region + col + 16, v1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int hpart = 0; hpart < 2; ++hpart) {
          const float* v = hpart ? v1 : v0;
          const int c0 = col + hpart * 16;
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            pk[e] = inside ? E::pack2(v[2 * e] + bias[c0 + 2 * e], v[2 * e + 1] + bias[c0 + 2 * e + 1]) : 0u;
          uint8_t* d = dst + (size_t(c0 / 8) * pitch + r + row_off) * 16;
          *reinterpret_cast<uint4*>(d) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(d + size_t(pitch) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncwarp();
    };

    for (int i = 0; i < nl; ++i) {
      // ---- h of layer i (pre or previous res update) -> bufH ----
      timed_wait(&h_full, ph_hfull, e_h);
      ph_hfull ^= 1u;
      tc::fence_after_sync();
      region_to_smem(T_H, s_cb + i * Hc, bufH, ROWS_H, 2);
      if (lane == 0) tc::mbar_arrive(&h_ready);
      // ---- gate chunks -> bufA ----
      for (int c = 0; c < n_gate_chunks; ++c, ++acc_it) {
        const int b = acc_it & 1;
        timed_wait(&acc_full[b], ((acc_it >> 1) & 1), e_acc);
        tc::fence_after_sync();
        float va[16], vb[16];
        const int j0 = hh * 16;
        tc::tmem_ld16(lane_base + T_ACC + b * FL_NC + j0, va);
        tc::tmem_ld16(lane_base + T_ACC + b * FL_NC + 32 + j0, vb);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[b]);  // accumulator is in registers now
        const int ch0 = c * 32 + j0;
        const float* ba = s_inb + i * 2 * Hc + ch0;
        const float* bb = ba + Hc;
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float a0 = va[2 * e] + ba[2 * e], a1 = va[2 * e + 1] + ba[2 * e + 1];
          const float g0 = vb[2 * e] + bb[2 * e], g1 = vb[2 * e + 1] + bb[2 * e + 1];
          pk[e] = E::pack2(tc::gate_tanh_sigmoid(a0, g0), tc::gate_tanh_sigmoid(a1, g1));
        }
        uint8_t* d = bufA + (size_t(ch0 / 8) * ROWS_A + r) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(d + size_t(ROWS_A) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&act_ready);
    }
    // ---- skip -> bufA (A operand of post) ----
    timed_wait(&h_full, ph_hfull, e_h);
    ph_hfull ^= 1u;
    tc::fence_after_sync();
    region_to_smem(T_SKIP, s_skb, bufA, ROWS_A, 0);
    if (lane == 0) tc::mbar_arrive(&act_ready);
    // ---- post chunks: x1 -= m ----
    const bool store = inside && r >= HALO && r < 128 - HALO;
    for (int c = 0; c < n_post_chunks; ++c, ++acc_it) {
      const int b = acc_it & 1;
      tc::mbar_wait(&acc_full[b], ((acc_it >> 1) & 1));
      tc::fence_after_sync();
      float v0[16], v1[16];
      const int j0 = hh * 32;
      tc::tmem_ld16(lane_base + T_ACC + b * FL_NC + j0, v0);
      tc::tmem_ld16(lane_base + T_ACC + b * FL_NC + j0 + 16, v1);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[b]);
      if (store) {
        float* row = p.z + (base + g) * (long long)p.z_stride + p.x1_coff;
        float4 cur[8];
        const int n0 = c * FL_NC + j0;
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4)
          if (n0 + e4 * 4 < half) cur[e4] = *reinterpret_cast<const float4*>(row + n0 + e4 * 4);
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const int n = n0 + e4 * 4;
          if (n >= half) continue;
          const float* v = e4 < 4 ? v0 + e4 * 4 : v1 + (e4 - 4) * 4;
          float4 o = cur[e4];
          o.x -= v[0] + s_pob[n + 0];
          o.y -= v[1] + s_pob[n + 1];
          o.z -= v[2] + s_pob[n + 2];
          o.w -= v[3] + s_pob[n + 3];
          *reinterpret_cast<float4*>(row + n) = o;
        }
      }
    }
    if (prof && warp == 2 && lane == 0) {
      unsigned long long* q = reinterpret_cast<unsigned long long*>(p.prof);
      atomicAdd(q + 9, (unsigned long long)(clock64() - k_pro));   // epilogue warp: whole window after the prologue
      atomicAdd(q + 10, (unsigned long long)e_h);
      atomicAdd(q + 11, (unsigned long long)e_acc);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
  if (prof && tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 7), (unsigned long long)(clock64() - k_start));
}

size_t flow_tc_smem_bytes(int Hc, int half, int nl) {
  const size_t bufH = (size_t(Hc / 8) * 133 * 16 + 127) & ~size_t(127);
  const size_t bufA = (size_t(Hc / 8) * 129 * 16 + 127) & ~size_t(127);
  const size_t ring = size_t(FL_STAGES) * Hc * FL_NC * 2;
  const size_t bias = sizeof(float) * (size_t(nl) * 2 * Hc + size_t(nl) * Hc + Hc + half);
  return bufH + bufA + ring + bias + 64;
}

bool flow_tc_supported(int Hc, int half, int nl, int kernel) {
  if (kernel != 5 || nl < 1 || nl > 8) return false;
  if (Hc != 192 || half != 96) return false;  // the issuer's k-step loops are unrolled for these (every shipped voice)
  if (2 * Hc + 2 * FL_NC > 512) return false;  // TMEM: H + SKIP + two accumulators
  if (128 - 4 * nl < 32) return false;
  return flow_tc_smem_bytes(Hc, half, nl) <= size_t(225 * 1024);
}

void launch_flow_tc(const FlowTcParams& p, int fmt, int n_seg, int max_len, cudaStream_t st) {
  if (n_seg <= 0 || max_len <= 0) return;
  const size_t smem = flow_tc_smem_bytes(p.Hc, p.half, p.nl);
  ensure_max_dynamic_smem(fmt ? reinterpret_cast<const void*>(flow_tc_kernel<1>) : reinterpret_cast<const void*>(flow_tc_kernel<0>));
  const int stride = 128 - 4 * p.nl;
  dim3 grid((max_len + stride - 1) / stride, n_seg);
  static const bool want_prof = getenv("M3B200_FLOW_PROFILE") != nullptr;
  static long long* d_prof = nullptr;
  FlowTcParams q = p;
  if (want_prof) {
    if (!d_prof) cudaMalloc(&d_prof, 16 * sizeof(long long));
    cudaMemsetAsync(d_prof, 0, 16 * sizeof(long long), st);
    q.prof = d_prof;
  }
  if (fmt) flow_tc_kernel<1><<<grid, FL_THREADS, smem, st>>>(q);
  else flow_tc_kernel<0><<<grid, FL_THREADS, smem, st>>>(q);
  post_launch("flow_tc_kernel", st);
  if (want_prof) {  // debug only: synchronous read-back of the per-role cycle counters (summed over windows)
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost);
    const double n = double(h[5] > 0 ? h[5] : 1);
    fprintf(stderr,
            "[flow profile] windows %lld | CTA cycles %.0f (prologue %.0f) | issuer: total %.0f wait_weights %.0f wait_acc_empty %.0f "
            "wait_h_ready %.0f wait_act_ready %.0f | producer wait_empty %.0f | epilogue warp: total %.0f wait_h_full %.0f "
            "wait_acc_full %.0f\n",
            h[5], h[7] / n, h[6] / n, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n);
  }
}

}  // namespace m3

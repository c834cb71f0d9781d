// MRF stage (sum of three ResBlock2 over the same input) for C = 128 channels:
// persistent + warp-specialised.  This is the one MRF stage whose bound IS the tensor pipe (N = 128 SS-mode
// MMAs run at 64.5 cycles = the math rate), so the kernel is organised around never letting the issuers wait:
//     out = 1/nk * sum_j [ x1_j + conv2_j(lrelu(x1_j)) ],   x1_j = x + b1_j + conv1_j(lrelu(x))
//   * one CTA per SM loops over (utterance, 256-row window) items;
//   * TMEM (512 columns): T = [0, 256) -- two 128-row tiles of the chain in flight: the epilogue writes x + b1_j
//     (tcgen05.st), conv1_j accumulates on top, so T ends as x1_j;  S = [256, 512) -- the epilogue adds x1_j into it
//     (read-modify-write while the tensor pipe is idle on S) and every conv2 accumulates on top: after the third
//     chain S = sum_j x1_j + sum_j conv2_j, and the final epilogue only adds the late bias and scales;
//   * chains run one after the other through T ("conv1_j | epi_j | conv2_j, conv1_j+1 | ..."): T is re-initialised
//     for chain j+1 while conv2_j runs, the next window's input is staged while the last conv2 runs, the final
//     epilogue of window i runs under conv1 of window i+1 -- the tensor pipe idles only while an epilogue turns
//     a finished conv1 into the conv2 operand (3 short bubbles per window);
//   * weights stream through a ring of 16 KB half-tap blocks (cp.async.bulk from L2: the stage's 960 KB of
//     weights stay L2-resident), full/empty mbarriers, one elected loader thread;
//   * one issuer warp per 128-row tile (a single thread's instruction stream cannot feed the pipe), every
//     descriptor differs from its template by a small additive constant.
// Shared memory: X (lrelu(x) fp16, 256 + 2 HX rows) 70 KB + Y (lrelu(x1_j) fp16, 256 + 2 HY rows, FIXED layout for
// all chains so its halo rows stay zero) 84 KB + ring 4 x 16 KB = 218 KB.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

// per-role cycle counters (M3B200_MRF_PROFILE=1) exist only in builds made with M3B200_KERNEL_PROFILE=1 (see kernels_tc_mrf2.cu)
#ifdef M3B200_KERNEL_PROFILE
#define M3_PROF3(p) ((p).prof != nullptr)
#else
#define M3_PROF3(p) false
#endif

namespace {
constexpr int bC = 128, bNT = 2, bR = bNT * 128, bCH = bC / 8, bKS = bC / 16;
constexpr int bSegTable = 1024;
constexpr uint32_t bSlotBytes = bC * bC * 2 / 2;  // half a tap: K-chunks [0, 8) or [8, 16) of [K/8][C][8]
constexpr uint32_t bT0 = 0, bS0 = 256;
constexpr int bMaxSlots = 6;
enum BBar { BXT_READY = 0, BT_READY, BC1_DONE = BT_READY + 2, BY_READY = BC1_DONE + 3, BC2_DONE = BY_READY + 3, BFULL,
            BEMPTY = BFULL + bMaxSlots, BNBAR = BEMPTY + bMaxSlots };

struct BGeo {
  int rows_x, rows_y;
  size_t off_ring, off_x, off_y, total;
};
__host__ __device__ inline BGeo make_bgeo(const MrfParams& p, int nslot) {
  BGeo g;
  g.rows_x = (bR + 2 * p.HX) | 1;  // odd pitches: conflict-free chunk-major stores
  g.rows_y = (bR + 2 * p.HY) | 1;
  size_t o = 0;
  g.off_ring = o;
  o += size_t(nslot) * bSlotBytes;
  g.off_x = o;
  o += size_t(bCH) * g.rows_x * 16;
  g.off_y = o;
  o += size_t(bCH) * g.rows_y * 16;
  g.total = o;
  return g;
}
__device__ __forceinline__ float blrelu(float v) { return fmaxf(v, 0.1f * v); }
__device__ __forceinline__ void bprefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ int bchain_at(int win, int pos) { return (win + pos) % 3; }
}  // namespace

// NEW = epilogue warps (8 or 16): lane quarter q = warp & 3, column group cg = warp >> 2 of 128 / (NEW / 4) channels
template <int FMT, int NEW>
__global__ void __launch_bounds__((NEW + 3) * 32, 1) mrf_ws128_kernel(MrfParams p) {
  using E = tc::Elem<FMT>;
  constexpr int kIssuer = NEW, kLoader = NEW + 2, kThreads = (NEW + 3) * 32, kEpiThreads = NEW * 32;
  constexpr int NCG = NEW / 4, HC = bC / NCG, NCC = HC / 16;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[BNBAR];
  __shared__ __align__(16) float sbias[4][bC];  // [j] first-conv bias of chain j, [3] summed second-conv bias
  __shared__ int s_rows[bSegTable];

  const int nslot = p.nslot;
  const BGeo g = make_bgeo(p, nslot);
  uint8_t* const ring = smem + g.off_ring;
  uint8_t* const bufX = smem + g.off_x;
  uint8_t* const bufY = smem + g.off_y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  auto seg_rows = [&](int seg) { return seg < bSegTable ? s_rows[seg] : p.seg_len[seg] * p.scale; };
  for (int i = tid; i < p.n_seg && i < bSegTable; i += kThreads) s_rows[i] = p.seg_len[i] * p.scale;
  __syncthreads();
  const int total = p.n_seg * p.max_win;
  auto valid = [&](int idx) {
    const int seg = idx / p.max_win, win = idx - seg * p.max_win;
    return win * p.stride < seg_rows(seg);
  };
  auto next_item = [&](int idx) {
    idx += int(gridDim.x);
    while (idx < total && !valid(idx)) idx += int(gridDim.x);
    return idx;
  };
  const int first = next_item(int(blockIdx.x) - int(gridDim.x));

  if (tid == 0) {
    for (int i = 0; i < BNBAR; ++i) {
      const bool from_issuers = (i >= BC1_DONE && i < BC1_DONE + 3) || i == BC2_DONE || i >= BEMPTY;  // one commit per issuer
      const bool from_loader = i >= BFULL && i < BEMPTY;
      tc::mbar_init(&bars[i], from_issuers ? 2 : (from_loader ? 1 : NEW));
    }
    tc::mbar_fence_init();
  }
  for (int i = tid; i < 4 * bC; i += kThreads) {
    const int j = i / bC, c = i - j * bC;
    sbias[j][c] = j < 3 ? p.bias[j][0][c] : p.late_bias[c];
  }
  {  // operand buffers start as zeros: the halo rows of Y are never written afterwards
    uint4* z = reinterpret_cast<uint4*>(smem + g.off_x);
    const int n16 = int((g.total - g.off_x) / 16);
    for (int i = tid; i < n16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == kIssuer) tc::tmem_alloc<512>(&tmem_slot);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const bool prof = M3_PROF3(p);
  auto timed_wait = [&](uint64_t* bar, uint32_t parity, long long& acc) {
    if (prof && !tc::mbar_test(bar, parity)) {
      const long long t = clock64();
      tc::mbar_wait(bar, parity);
      acc += clock64() - t;
    } else {
      tc::mbar_wait(bar, parity);
    }
  };

  if (warp == kLoader) {
    // =================================== weight loader ============================================
    // tap order of a window = the issuers' order: conv1, conv2 of the chain at position 0, then 1, then 2
    if (tc::elect_one()) {
      uint32_t slot = 0, eparity = 1u;  // parity 1 of a fresh barrier passes at once (first use of a slot)
      for (int idx = first; idx < total; idx = next_item(idx)) {
        const int win = idx % p.max_win;
        for (int pos = 0; pos < 3; ++pos) {
          const int j = bchain_at(win, pos);
          for (int d = 0; d < 2; ++d) {
            const uint16_t* src = p.w16 + p.woff[j][d];
            for (int t = 0; t < 2 * p.k[j]; ++t) {  // half-taps are contiguous in [tap][K/8][C][8]
              tc::mbar_wait(&bars[BEMPTY + slot], eparity);
              tc::mbar_expect_tx(&bars[BFULL + slot], bSlotBytes);
              tc::bulk_g2s(ring + size_t(slot) * bSlotBytes, src + size_t(t) * (bSlotBytes / 2), bSlotBytes, &bars[BFULL + slot]);
              if (++slot == uint32_t(nslot)) {
                slot = 0;
                eparity ^= 1u;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= kIssuer && warp < kIssuer + 2) {
    // =================================== MMA issuers ==============================================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, bC, FMT);
      const uint64_t b_tmpl = tc::make_desc(tc::smem_u32(ring), uint32_t(bC) * 16u, 128u);
      const uint32_t b_hi = uint32_t(b_tmpl >> 32), b_lo0 = uint32_t(b_tmpl);
      const int my_tile = warp - kIssuer;
      uint32_t slot = 0, fparity = 0u;
      long long c_full = 0, c_x = 0, c_t = 0, c_y = 0;
      const long long c_start = prof ? clock64() : 0;
      // one conv = k taps x 2 half-taps x 4 K-steps on this issuer's tile, all accumulating (the accumulator was
      // initialised by the epilogue warps)
      auto conv = [&](uint32_t abase, int rows_in, int halo, int k, int dil, uint32_t dcol) {
        const uint64_t a_tmpl = tc::make_desc(abase, uint32_t(rows_in) * 16u, 128u);
        const uint32_t a_hi = uint32_t(a_tmpl >> 32);
        uint32_t at = uint32_t(a_tmpl) + uint32_t(halo - ((k - 1) / 2) * dil) + uint32_t(my_tile * 128);
        const uint32_t kstep = uint32_t(2 * rows_in);  // 2 chunks of 16 B x rows_in per K-step of 16, in 16-byte units
#pragma unroll 1
        for (int t = 0; t < k; ++t, at += uint32_t(dil)) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            timed_wait(&bars[BFULL + slot], fparity, c_full);
            tc::fence_after_sync();  // (measured free: r02l A/B with and without it)
            const uint32_t bt = b_lo0 + slot * (bSlotBytes >> 4);
#pragma unroll
            for (int ks = 0; ks < bKS / 2; ++ks) {
              const uint64_t bd = (uint64_t(b_hi) << 32) | uint64_t(bt + uint32_t(ks * 2 * bC));
              const uint64_t ad = (uint64_t(a_hi) << 32) | uint64_t(at + uint32_t(hf * (bKS / 2) + ks) * kstep);
              tc::mma_f16_ss(tmem + dcol + uint32_t(my_tile * bC), ad, bd, idesc, 1u);
            }
            tc::mma_commit(&bars[BEMPTY + slot]);
            if (++slot == uint32_t(nslot)) {
              slot = 0;
              fparity ^= 1u;
            }
          }
        }
      };
      int it = 0;
      for (int idx = first; idx < total; idx = next_item(idx), ++it) {
        const uint32_t par = uint32_t(it) & 1u;
        const int win = idx % p.max_win;
        timed_wait(&bars[BXT_READY], par, c_x);
        tc::fence_after_sync();
        for (int pos = 0; pos < 3; ++pos) {
          const int j = bchain_at(win, pos);
          if (pos > 0) {
            timed_wait(&bars[BT_READY + pos - 1], par, c_t);
            tc::fence_after_sync();
          }
          conv(tc::smem_u32(bufX), g.rows_x, p.HX, p.k[j], p.dil[j][0], bT0);
          tc::mma_commit(&bars[BC1_DONE + pos]);
          timed_wait(&bars[BY_READY + pos], par, c_y);
          tc::fence_after_sync();
          conv(tc::smem_u32(bufY), g.rows_y, p.HY, p.k[j], p.dil[j][1], bS0);
        }
        tc::mma_commit(&bars[BC2_DONE]);
      }
      if (prof && my_tile == 0) {
        unsigned long long* q = reinterpret_cast<unsigned long long*>(p.prof);
        atomicAdd(q + 0, (unsigned long long)(clock64() - c_start));
        atomicAdd(q + 1, (unsigned long long)c_full);
        atomicAdd(q + 2, (unsigned long long)c_x);
        atomicAdd(q + 3, (unsigned long long)c_t);
        atomicAdd(q + 4, (unsigned long long)c_y);
        atomicAdd(q + 5, (unsigned long long)it);
      }
    }
    __syncwarp();
  } else {
    // =================================== epilogue warps ===========================================
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
    const int col0 = cg * HC;

    auto arrive = [&](int b) {
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[b]);
    };
    // 16 channels of one row (lrelu -> 16-bit) -> two 16-byte operand chunks; rows outside the utterance are zero
    auto store_ops = [&](uint8_t* buf, int pitch, int row, int col, const float* v, bool inside) {
#pragma unroll
      for (int c8 = 0; c8 < 2; ++c8) {
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (inside) {
          pk.x = E::pack2(blrelu(v[8 * c8]), blrelu(v[8 * c8 + 1]));
          pk.y = E::pack2(blrelu(v[8 * c8 + 2]), blrelu(v[8 * c8 + 3]));
          pk.z = E::pack2(blrelu(v[8 * c8 + 4]), blrelu(v[8 * c8 + 5]));
          pk.w = E::pack2(blrelu(v[8 * c8 + 6]), blrelu(v[8 * c8 + 7]));
        }
        *reinterpret_cast<uint4*>(buf + (size_t(col / 8 + c8) * pitch + row) * 16) = pk;
      }
    };
    // T <- x + b1_j for this thread's rows / columns (x from global: L2-hot, it was just read for the operand)
    auto init_T = [&](int idx, int j, bool also_x) {
      const int seg = idx / p.max_win, win = idx - seg * p.max_win;
      const int L = seg_rows(seg);
      const long long base = (long long)p.seg_off[seg] * p.scale;
      const int w0 = win * p.stride - p.H;
#pragma unroll
      for (int m = 0; m < bNT; ++m) {
        const int r = m * 128 + q * 32 + lane;
        const int gi = w0 + r;
        const bool inside = gi >= 0 && gi < L;
        const float* src = p.x + (base + gi) * bC + col0;
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) {
          float v[16];
          if (inside) {
            tc::ldg256(src + 16 * cc, v);
            tc::ldg256(src + 16 * cc + 8, v + 8);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0.f;
          }
          if (also_x) store_ops(bufX, g.rows_x, r + p.HX, col0 + 16 * cc, v, inside);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 bb = *reinterpret_cast<const float4*>(&sbias[j][col0 + 16 * cc + 4 * c]);
            v[4 * c] += bb.x;
            v[4 * c + 1] += bb.y;
            v[4 * c + 2] += bb.z;
            v[4 * c + 3] += bb.w;
          }
          tc::tmem_st16(lane_base + bT0 + uint32_t(m * bC + col0 + 16 * cc), v);
        }
      }
      if (also_x) {  // the HX halo rows on both sides of the window
        const int items = 2 * p.HX * bCH;
        for (int i = tid; i < items; i += kEpiThreads) {
          const int rr = i / bCH, c8 = i - rr * bCH;
          const int hrow = rr < p.HX ? rr : bR + rr;  // bufX rows [0, HX) and [R + HX, R + 2 HX)
          const int gi = w0 - p.HX + hrow;
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (gi >= 0 && gi < L) {
            float hv[8];
            tc::ldg256(p.x + (base + gi) * bC + c8 * 8, hv);
            pk.x = E::pack2(blrelu(hv[0]), blrelu(hv[1]));
            pk.y = E::pack2(blrelu(hv[2]), blrelu(hv[3]));
            pk.z = E::pack2(blrelu(hv[4]), blrelu(hv[5]));
            pk.w = E::pack2(blrelu(hv[6]), blrelu(hv[7]));
          }
          *reinterpret_cast<uint4*>(bufX + (size_t(c8) * g.rows_x + hrow) * 16) = pk;
        }
      }
      tc::tmem_st_wait();
    };
    auto prefetch_x = [&](int idx) {
      const int seg = idx / p.max_win, win = idx - seg * p.max_win;
      const int L = seg_rows(seg);
      const long long base = (long long)p.seg_off[seg] * p.scale;
      const int w0 = win * p.stride - p.H;
#pragma unroll
      for (int m = 0; m < bNT; ++m) {
        const int gi = w0 + m * 128 + q * 32 + lane;
        if (gi >= 0 && gi < L) bprefetch_l2(p.x + (base + gi) * bC + col0);
      }
    };

    long long e_c1 = 0, e_c2 = 0;
    const long long e_start = prof ? clock64() : 0;
    if (first < total) {
      init_T(first, bchain_at(first % p.max_win, 0), true);
      arrive(BXT_READY);
    }
    int it = 0;
    for (int idx = first; idx < total; ++it) {
      const int nxt = next_item(idx);
      const uint32_t par = uint32_t(it) & 1u;
      const int seg = idx / p.max_win, win = idx - seg * p.max_win;
      const int w0 = win * p.stride - p.H;
      const int L = seg_rows(seg);
      if (nxt < total) {
        const int nn = next_item(nxt);
        if (nn < total) prefetch_x(nn);
      }
#pragma unroll 1
      for (int pos = 0; pos < 3; ++pos) {
        // ---- conv1 of the chain is complete: S += x1, Y <- lrelu(x1) ----
        timed_wait(&bars[BC1_DONE + pos], par, e_c1);
        tc::fence_after_sync();
#pragma unroll
        for (int m = 0; m < bNT; ++m) {
          const int r = m * 128 + q * 32 + lane;
          const int gi = w0 + r;
          const bool inside = gi >= 0 && gi < L;
#pragma unroll
          for (int cc = 0; cc < NCC; ++cc) {
            float v[16], s[16];
            const uint32_t col = uint32_t(m * bC + col0 + 16 * cc);
            tc::tmem_ld16(lane_base + bT0 + col, v);
            if (pos > 0) tc::tmem_ld16(lane_base + bS0 + col, s);
            tc::tmem_ld_wait();
            if (pos > 0) {
#pragma unroll
              for (int e = 0; e < 16; ++e) s[e] += v[e];
              tc::tmem_st16(lane_base + bS0 + col, s);
            } else {
              tc::tmem_st16(lane_base + bS0 + col, v);
            }
            store_ops(bufY, g.rows_y, r + p.HY, col0 + 16 * cc, v, inside);
          }
        }
        tc::tmem_st_wait();
        arrive(BY_READY + pos);
        // ---- T for the next chain of this window (runs under conv2 of this chain) ----
        if (pos < 2) {
          init_T(idx, bchain_at(win, pos + 1), false);
          arrive(BT_READY + pos);
        }
      }
      // ---- next window: operand X + T of its first chain, while the last conv2 of this window runs ----
      if (nxt < total) {
        init_T(nxt, bchain_at(nxt % p.max_win, 0), true);
        arrive(BXT_READY);
      }
      // ---- out = (S + late bias) / nk ----
      timed_wait(&bars[BC2_DONE], par, e_c2);
      tc::fence_after_sync();
      {
        const long long base = (long long)p.seg_off[seg] * p.scale;
#pragma unroll
        for (int m = 0; m < bNT; ++m) {
          const int r = m * 128 + q * 32 + lane;
          const int gi = w0 + r;
          const bool store = r >= p.H && r < bR - p.H && gi < L;
          float* dst = p.out + (base + gi) * bC + col0;
#pragma unroll
          for (int cc = 0; cc < NCC; ++cc) {
            float v[16];
            tc::tmem_ld16(lane_base + bS0 + uint32_t(m * bC + col0 + 16 * cc), v);
            tc::tmem_ld_wait();
            if (store) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float4 bb = *reinterpret_cast<const float4*>(&sbias[3][col0 + 16 * cc + 4 * c]);
                v[4 * c] = (v[4 * c] + bb.x) * p.inv_nk;
                v[4 * c + 1] = (v[4 * c + 1] + bb.y) * p.inv_nk;
                v[4 * c + 2] = (v[4 * c + 2] + bb.z) * p.inv_nk;
                v[4 * c + 3] = (v[4 * c + 3] + bb.w) * p.inv_nk;
              }
              tc::stg256(dst + 16 * cc, v);
              tc::stg256(dst + 16 * cc + 8, v + 8);
            }
          }
        }
      }
      idx = nxt;
    }
    if (prof && tid == 0) {
      unsigned long long* q = reinterpret_cast<unsigned long long*>(p.prof);
      atomicAdd(q + 8, (unsigned long long)(clock64() - e_start));
      atomicAdd(q + 9, (unsigned long long)e_c1);
      atomicAdd(q + 10, (unsigned long long)e_c2);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kIssuer) tc::tmem_dealloc<512>(tmem);
}

static int mrf_ws128_slots(const MrfParams& p, size_t* smem_out) {
  int optin = 227 * 1024, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t budget = size_t(optin) - 2560 - 4 * bSegTable - 128;  // static shared memory (barriers, biases, row table) + alignment
  const BGeo g0 = make_bgeo(p, 0);
  if (g0.total + 3 * bSlotBytes > budget) return 0;
  int n = int((budget - g0.total) / bSlotBytes);
  if (n > bMaxSlots) n = bMaxSlots;
  if (smem_out) *smem_out = make_bgeo(p, n).total + 128;
  return n;
}

bool mrf_ws128_supported(const MrfParams& p, int C) {
  if (C != bC || p.nk != 3 || p.nd != 2) return false;
  for (int j = 0; j < 3; ++j)
    if (p.k[j] < 1 || p.k[j] > 11 || !(p.k[j] & 1)) return false;
  if (p.H != p.HY || bR - 2 * p.H < 64) return false;
  return mrf_ws128_slots(p, nullptr) >= 3;
}

void launch_mrf_ws128(const MrfParams& p_in, int fmt, int n_seg, int max_len, cudaStream_t st) {
  MrfParams p = p_in;

  p.stride = bR - 2 * p.H;
  const int L = max_len * p.scale;
  p.n_seg = n_seg;
  p.max_win = (L + p.stride - 1) / p.stride;
  if (p.max_win <= 0 || n_seg <= 0) return;
  size_t smem = 0;
  p.nslot = mrf_ws128_slots(p, &smem);
  if (p.nslot < 3) throw std::runtime_error("mrf_ws128: shared memory budget");
  static const int n_sm = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  static const int new_warps = [] { const char* e = getenv("M3B200_MRF128_WARPS"); return e ? atoi(e) : 16; }();
  const long long items = (long long)n_seg * p.max_win;
  const int grid = int(items < n_sm ? items : n_sm);
  static const bool want_prof = getenv("M3B200_MRF_PROFILE") != nullptr;
  static long long* d_prof = nullptr;
  if (want_prof) {
    if (!d_prof) cudaMalloc(&d_prof, 16 * sizeof(long long));
    cudaMemsetAsync(d_prof, 0, 16 * sizeof(long long), st);
    p.prof = d_prof;
  }
#define M3_WS128(F, W)                                                                        \
  {                                                                                           \
    ensure_max_dynamic_smem(reinterpret_cast<const void*>(mrf_ws128_kernel<F, W>));           \
    mrf_ws128_kernel<F, W><<<grid, (W + 3) * 32, smem, st>>>(p);                              \
  }
  if (new_warps == 8) {
    if (fmt) M3_WS128(1, 8) else M3_WS128(0, 8)
  } else {
    if (fmt) M3_WS128(1, 16) else M3_WS128(0, 16)
  }
#undef M3_WS128
  post_launch("mrf_ws128_kernel", st);
  if (want_prof) {  // debug only
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost);
    const double w = h[5] > 0 ? double(h[5]) : 1.0;
    fprintf(stderr,
            "[mrf_ws128 profile] windows %lld grid %d | issuer cycles/window: total %.0f wait_full %.0f wait_xt %.0f wait_t %.0f wait_y %.0f | "
            "epilogue warp0: total %.0f wait_c1 %.0f wait_c2 %.0f\n",
            h[5], grid, h[0] / w, h[1] / w, h[2] / w, h[3] / w, h[4] / w, h[8] / w, h[9] / w, h[10] / w);
  }
}

}  // namespace m3

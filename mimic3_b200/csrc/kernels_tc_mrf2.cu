// MRF stage (sum of three ResBlock2 over the same input) for C = 64 channels,
// persistent + warp-specialised.  Same math as mrf_tc_kernel (kernels_tc.cu):
//     out = 1/nk * sum_j [ x1_j + conv2_j(lrelu(x1_j)) ],   x1_j = x + conv1_j(lrelu(x))
// but organised like dec_fused_kernel (kernels_tc_dec2.cu):
//   * one CTA per SM loops over (utterance, window) items; warps 0-7 are epilogue warps, warps 8-9 issue
//     the tcgen05.mma of one 128-row tile each, warp 10 streams the weights tap by tap (8 KB cp.async.bulk blocks, L2-resident)
//     through a ring of shared-memory slots guarded by full/empty mbarriers;
//   * the three resblocks are independent chains: the tensor pipe runs conv1 of chain j+1 while the
//     epilogue warps turn chain j's accumulator into its second conv's operand, and all second convs
//     accumulate into one TMEM tile S (x stays in registers; sum_j x1_j is parked in TMEM tiles that are
//     idle at that point);
//   * the NEXT window's input is fetched and published (lrelu -> fp16 operand) while the second convs of
//     the current window still run, so the issuer never waits for global memory.
// TMEM (512 columns): T_j = [128 j, 128 j + 128) for the two 128-row tiles of chain j, S = [384, 512).
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

// per-role cycle counters (M3B200_MRF_PROFILE=1) exist only in builds made with M3B200_KERNEL_PROFILE=1 in the
// environment of `python -m mimic3_b200.build`: merely compiled in they cost the persistent kernels 5-9 %
#ifdef M3B200_KERNEL_PROFILE
#define M3_PROF(p) ((p).prof != nullptr)
#else
#define M3_PROF(p) false
#endif

namespace {
constexpr int wC = 64, wNT = 2, wR = wNT * 128, wCH = wC / 8, wKS = wC / 16;
constexpr int wIssuers = wNT;  // epilogue warps NEW (8 or 16) come first, then the issuers, then the weight loader
constexpr int wSegTable = 1024;  // per-utterance row counts cached in shared memory (larger batches read global memory)
constexpr uint32_t wTapBytes = wC * wC * 2;
constexpr int wStageTaps = 2;  // taps per ring slot / issuer stage: every stage costs the issuing thread a wait and a
                               // tcgen05.commit (~140 cycles, m3_selftest 501-560) on top of its 4 MMAs per tap
constexpr uint32_t wSlotBytes = wStageTaps * wTapBytes;
constexpr uint32_t wS0 = 384;
constexpr int wMaxSlots = 6;
enum WBar { WX_READY = 0, WC1_DONE, WY_READY = WC1_DONE + 3, WC2_DONE = WY_READY + 3, WF_DONE, WFULL, WEMPTY = WFULL + wMaxSlots, WNBAR = WEMPTY + wMaxSlots };

struct WGeo {
  int rows_x, rows_y[3], hy[3];
  size_t off_ring, off_x, off_y[3], total;
};
__host__ __device__ inline WGeo make_wgeo(const MrfParams& p, int nslot) {
  WGeo g;
  g.rows_x = (wR + 2 * p.HX) | 1;
  size_t o = 0;
  g.off_ring = o;
  o += size_t(nslot) * wSlotBytes;
  g.off_x = o;
  o += size_t(wCH) * g.rows_x * 16;
  for (int j = 0; j < 3; ++j) {
    g.hy[j] = p.dil[j][1] * (p.k[j] - 1) / 2;
    g.rows_y[j] = (wR + 2 * g.hy[j]) | 1;
    g.off_y[j] = o;
    o += size_t(wCH) * g.rows_y[j] * 16;
  }
  g.total = o;
  return g;
}
__device__ __forceinline__ float wlrelu(float v, float s) { return fmaxf(v, s * v); }
using tc::ldg256;
using tc::stg256;
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
// chain order of a window: rotated by the window index, so that the SMs (which work on different windows at
// any moment) do not all pull the same weight block from the same L2 slices at the same time; a function of
// the window only, so results do not depend on the batch composition
__device__ __forceinline__ int chain_at(int win, int pos) { return (win + pos) % 3; }
}  // namespace

// NEW = epilogue warps: 8 (each thread owns 32 channels of its rows) or 16 (16 channels: half the dependent
// tmem_ld -> math -> st.shared chain per thread and twice the warps to hide it -- the M3B200_MRF_PROFILE counters show
// the epilogue warps, not the tensor pipe, are the busiest resource of this kernel)
template <int FMT, int NEW>
__global__ void __maxnreg__(NEW == 16 ? 96 : 168) mrf_ws_kernel(MrfParams p) {  // 19 warps x 96 x 32 registers fit one SM
  using E = tc::Elem<FMT>;
  constexpr int wEpiWarps = NEW, wIssuer = NEW, wLoader = wIssuer + wIssuers, wThreads = 32 * (wLoader + 1);
  constexpr int NCG = NEW / 4, HC = wC / NCG, NH = HC / 16;  // column groups, channels and 16-channel halves per thread
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[WNBAR];
  __shared__ __align__(16) float sbias[4][wC];  // [j] first-conv bias of chain j, [3] summed second-conv bias

  const int nslot = p.nslot;
  const WGeo g = make_wgeo(p, nslot);
  uint8_t* const ring = smem + g.off_ring;
  uint8_t* const bufX = smem + g.off_x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  __shared__ int s_rows[wSegTable];
  auto seg_rows = [&](int seg) { return seg < wSegTable ? s_rows[seg] : p.seg_len[seg] * p.scale; };
  for (int i = threadIdx.x; i < p.n_seg && i < wSegTable; i += wThreads) s_rows[i] = p.seg_len[i] * p.scale;
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
    for (int i = 0; i < WNBAR; ++i) {
      const bool many = i == WX_READY || (i >= WY_READY && i < WY_READY + 3) || i == WF_DONE;
      const bool from_issuers = (i >= WC1_DONE && i < WC1_DONE + 3) || i == WC2_DONE || i >= WEMPTY;  // one commit per issuer
      tc::mbar_init(&bars[i], many ? wEpiWarps : (from_issuers ? wIssuers : 1));
    }
    tc::mbar_fence_init();
  }
  for (int i = tid; i < 4 * wC; i += wThreads) {
    const int j = i / wC, c = i - j * wC;
    sbias[j][c] = j < 3 ? p.bias[j][0][c] : p.late_bias[c];
  }
  {  // operand buffers start as zeros (halo rows of bufY that no epilogue writes stay zero)
    uint4* z = reinterpret_cast<uint4*>(smem + g.off_x);
    const int n16 = int((g.total - g.off_x) / 16);
    for (int i = tid; i < n16; i += wThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == wIssuer) tc::tmem_alloc<512>(&tmem_slot);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == wLoader) {
    // =================================== weight loader ============================================
    // one elected thread: the tap sequence of every window is (conv1 of chains 0..2, conv2 of chains 0..2)
    if (tc::elect_one()) {
      uint32_t slot = 0, eparity = 1u;  // waiting on parity 1 of a fresh barrier passes at once (first use of a slot)
      long long l_empty = 0, l_blocked = 0;
      for (int idx = first; idx < total; idx = next_item(idx)) {
        const int win = idx % p.max_win;
        for (int d = 0; d < 2; ++d)
          for (int jj = 0; jj < 3; ++jj) {
            const int j = chain_at(win, jj);
            const uint16_t* src = p.w16 + p.woff[j][d];
            for (int t = 0; t < p.k[j]; t += wStageTaps) {
              const uint32_t bytes = uint32_t(p.k[j] - t < wStageTaps ? p.k[j] - t : wStageTaps) * wTapBytes;
              if (M3_PROF(p) && !tc::mbar_test(&bars[WEMPTY + slot], eparity)) {
                const long long t0 = clock64();
                tc::mbar_wait(&bars[WEMPTY + slot], eparity);
                l_empty += clock64() - t0;
                ++l_blocked;
              } else {
                tc::mbar_wait(&bars[WEMPTY + slot], eparity);
              }
              tc::mbar_expect_tx(&bars[WFULL + slot], bytes);
              tc::bulk_g2s(ring + size_t(slot) * wSlotBytes, src + size_t(t) * wC * wC, bytes, &bars[WFULL + slot]);
              if (++slot == uint32_t(nslot)) {
                slot = 0;
                eparity ^= 1u;
              }
            }
          }
      }
      if (M3_PROF(p)) {
        atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 14), (unsigned long long)l_empty);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 15), (unsigned long long)l_blocked);
      }
    }
    __syncwarp();
  } else if (warp >= wIssuer && warp < wIssuer + wIssuers) {
    // =================================== MMA issuers ==============================================
    // One issuer per 128-row tile (independent accumulators): a single thread could not keep the tensor pipe
    // busy (its per-tap instruction stream took ~640 cycles against 390 cycles of MMA time, measured with the
    // M3B200_MRF_PROFILE counters).  In each issuer warp ONE thread runs the whole schedule (waits included): no per-tap warp reconvergence, no divisions, and
    // the descriptors of a conv differ only by small additive constants in their low word (the 14-bit start
    // address never overflows: shared memory is < 256 KB), so each MMA costs one add and the instruction.
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, wC, FMT);
      const uint64_t b_tmpl = tc::make_desc(tc::smem_u32(ring), uint32_t(wC) * 16u, 128u);
      const uint32_t b_hi = uint32_t(b_tmpl >> 32), b_lo0 = uint32_t(b_tmpl);
      const int my_tile = warp - wIssuer;
      uint32_t slot = 0, fparity = 0u;
      long long c_full = 0, c_x = 0, c_y = 0, c_f = 0;
      const bool prof = M3_PROF(p);
      const long long c_start = clock64();
      long long n_blocked = 0;
      auto timed_wait = [&](uint64_t* bar, uint32_t parity, long long& acc) {
        if (prof) {  // only waits that actually block are timed (a probe that succeeds costs nothing)
          if (tc::mbar_test(bar, parity)) return;
          const long long t0 = clock64();
          tc::mbar_wait(bar, parity);
          acc += clock64() - t0;
          if (&acc == &c_full) ++n_blocked;
        } else {
          tc::mbar_wait(bar, parity);
        }
      };
      auto conv = [&](uint32_t abase, int rows_in, int halo, int k, int dil, uint32_t dcol, bool acc0) {
        const uint64_t a_tmpl = tc::make_desc(abase, uint32_t(rows_in) * 16u, 128u);
        const uint32_t a_hi = uint32_t(a_tmpl >> 32);
        uint32_t at = uint32_t(a_tmpl) + uint32_t(halo - ((k - 1) / 2) * dil);
        const uint32_t kstep = uint32_t(2 * rows_in);
#pragma unroll 1
        for (int t = 0; t < k; t += wStageTaps) {
          timed_wait(&bars[WFULL + slot], fparity, c_full);
          tc::fence_after_sync();  // (measured free: r02l A/B with and without it)
          const uint32_t bt0 = b_lo0 + slot * (wSlotBytes >> 4);
#pragma unroll
          for (int tt = 0; tt < wStageTaps; ++tt) {
            if (t + tt < k) {
              const uint32_t bt = bt0 + uint32_t(tt) * (wTapBytes >> 4);
#pragma unroll
              for (int ks = 0; ks < wKS; ++ks) {
                const uint64_t bd = (uint64_t(b_hi) << 32) | uint64_t(bt + uint32_t(ks * 2 * wC));
                const uint32_t acc = (ks > 0 || acc0 || t + tt > 0) ? 1u : 0u;
                const uint64_t ad = (uint64_t(a_hi) << 32) | uint64_t(at + uint32_t(ks) * kstep + uint32_t(my_tile * 128));
                tc::mma_f16_ss(tmem + dcol + uint32_t(my_tile * wC), ad, bd, idesc, acc);
              }
              at += uint32_t(dil);
            }
          }
          tc::mma_commit(&bars[WEMPTY + slot]);
          if (++slot == uint32_t(nslot)) {
            slot = 0;
            fparity ^= 1u;
          }
        }
      };
      int it = 0;
      for (int idx = first; idx < total; idx = next_item(idx), ++it) {
        const uint32_t par = uint32_t(it) & 1u;
        const int win = idx % p.max_win;
        timed_wait(&bars[WX_READY], par, c_x);
        tc::fence_after_sync();
        for (int jj = 0; jj < 3; ++jj) {
          const int j = chain_at(win, jj);
          if (jj == 2 && it > 0) {  // this T tile still holds the previous window's sum until its final epilogue has read it
            timed_wait(&bars[WF_DONE], uint32_t(it - 1) & 1u, c_f);
            tc::fence_after_sync();
          }
          conv(tc::smem_u32(bufX), g.rows_x, p.HX, p.k[j], p.dil[j][0], uint32_t(j) * 128u, false);
          tc::mma_commit(&bars[WC1_DONE + j]);
        }
        for (int jj = 0; jj < 3; ++jj) {
          const int j = chain_at(win, jj);
          timed_wait(&bars[WY_READY + j], par, c_y);
          tc::fence_after_sync();
          conv(tc::smem_u32(smem + g.off_y[j]), g.rows_y[j], g.hy[j], p.k[j], p.dil[j][1], wS0, jj > 0);
        }
        tc::mma_commit(&bars[WC2_DONE]);
      }
      if (prof && my_tile == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 0), (unsigned long long)(clock64() - c_start));
        atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 1), (unsigned long long)c_full);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 2), (unsigned long long)c_x);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 3), (unsigned long long)c_y);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 4), (unsigned long long)c_f);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 5), (unsigned long long)it);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 6), (unsigned long long)n_blocked);
      }
    }
    __syncwarp();
  } else {
    // =================================== epilogue warps ===========================================
    // warp w: TMEM lane quarter q = w & 3 (hardware restriction), channel half hh = w >> 2 (32 channels)
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
    const int col0 = cg * HC;
    float xr[wNT][HC];  // this thread's x (fp32 residual source): rows {m*128 + q*32 + lane}, HC channels

    auto arrive = [&](int b) {
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[b]);
    };
    // 16 channels (half h of this thread's 32) of one row -> two 16-byte operand chunks
    auto store_ops = [&](uint8_t* buf, int pitch, int row, int h, const float* v, bool inside) {
#pragma unroll
      for (int c8 = 0; c8 < 2; ++c8) {
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (inside) {
          pk.x = E::pack2(wlrelu(v[8 * c8], 0.1f), wlrelu(v[8 * c8 + 1], 0.1f));
          pk.y = E::pack2(wlrelu(v[8 * c8 + 2], 0.1f), wlrelu(v[8 * c8 + 3], 0.1f));
          pk.z = E::pack2(wlrelu(v[8 * c8 + 4], 0.1f), wlrelu(v[8 * c8 + 5], 0.1f));
          pk.w = E::pack2(wlrelu(v[8 * c8 + 6], 0.1f), wlrelu(v[8 * c8 + 7], 0.1f));
        }
        *reinterpret_cast<uint4*>(buf + (size_t(col0 / 8 + h * 2 + c8) * pitch + row) * 16) = pk;
      }
    };
    // fetch the window's x: own rows into registers (fp32 residual) and, as lrelu -> 16-bit, into bufX
    // together with the HX halo rows on both sides
    auto load_x = [&](int idx) {
      const int seg = idx / p.max_win, win = idx - seg * p.max_win;
      const int L = seg_rows(seg);
      const long long base = (long long)p.seg_off[seg] * p.scale;
      const int w0 = win * p.stride - p.H;
#pragma unroll
      for (int m = 0; m < wNT; ++m) {
        const int gi = w0 + m * 128 + q * 32 + lane;
        const float* src = p.x + (base + gi) * wC + col0;
        if (gi >= 0 && gi < L) {
#pragma unroll
          for (int e = 0; e < HC / 8; ++e) ldg256(src + 8 * e, &xr[m][8 * e]);
        } else {
#pragma unroll
          for (int e = 0; e < HC; ++e) xr[m][e] = 0.f;
        }
      }
      float hv[8];
      const bool has_halo = tid < 2 * p.HX * wCH;  // halo item: (row rr of the 2*HX halo rows, chunk c8)
      int hrow = 0, hc8 = 0;
      bool hin = false;
      if (has_halo) {
        const int rr = tid / wCH;
        hc8 = tid - rr * wCH;
        hrow = rr < p.HX ? rr : wR + rr;  // bufX rows [0, HX) and [R + HX, R + 2 HX)
        const int gi = w0 - p.HX + hrow;
        hin = gi >= 0 && gi < L;
        if (hin) ldg256(p.x + (base + gi) * wC + hc8 * 8, hv);
      }
#pragma unroll
      for (int m = 0; m < wNT; ++m) {
        const int r = m * 128 + q * 32 + lane;
        const int gi = w0 + r;
#pragma unroll
        for (int h = 0; h < NH; ++h) store_ops(bufX, g.rows_x, r + p.HX, h, xr[m] + 16 * h, gi >= 0 && gi < L);
      }
      if (has_halo) {
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (hin) {
          pk.x = E::pack2(wlrelu(hv[0], 0.1f), wlrelu(hv[1], 0.1f));
          pk.y = E::pack2(wlrelu(hv[2], 0.1f), wlrelu(hv[3], 0.1f));
          pk.z = E::pack2(wlrelu(hv[4], 0.1f), wlrelu(hv[5], 0.1f));
          pk.w = E::pack2(wlrelu(hv[6], 0.1f), wlrelu(hv[7], 0.1f));
        }
        *reinterpret_cast<uint4*>(bufX + (size_t(hc8) * g.rows_x + hrow) * 16) = pk;
      }
    };
    // pull a window's x rows into L2 one window ahead of their use (no registers involved)
    auto prefetch_x = [&](int idx) {
      const int seg = idx / p.max_win, win = idx - seg * p.max_win;
      const int L = seg_rows(seg);
      const long long base = (long long)p.seg_off[seg] * p.scale;
      const int w0 = win * p.stride - p.H;
#pragma unroll
      for (int m = 0; m < wNT; ++m) {
        const int gi = w0 + m * 128 + q * 32 + lane;
        if (gi >= 0 && gi < L) prefetch_l2(p.x + (base + gi) * wC + col0);
      }
    };

    // profile counters of thread 0 live in shared memory (seven 64-bit registers per epilogue thread would spill)
    const bool prof = M3_PROF(p) && tid == 0;
    __shared__ long long s_prof[8];
    long long& e_c1w = s_prof[0];
    long long& e_c1 = s_prof[1];
    long long& e_lx = s_prof[2];
    long long& e_c2w = s_prof[3];
    long long& e_fin = s_prof[4];
    long long& e_t = s_prof[5];
    long long& e_start = s_prof[6];
    if (prof) {
      e_c1w = e_c1 = e_lx = e_c2w = e_fin = e_t = 0;
      e_start = clock64();
    }
    if (first < total) {
      load_x(first);
      arrive(WX_READY);
    }
    int it = 0;
    for (int idx = first; idx < total; ++it) {
      const int nxt = next_item(idx);
      const uint32_t par = uint32_t(it) & 1u;
      const int seg = idx / p.max_win, win = idx - seg * p.max_win;
      const int w0 = win * p.stride - p.H;
      if (nxt < total) {
        const int nn = next_item(nxt);
        if (nn < total) prefetch_x(nn);
      }
      // TMEM tile that keeps sum_j x1_j until the final epilogue: the chain the NEXT window runs last, so the
      // issuer (which waits for WF_DONE before that chain's first conv) is never held up by it
      const int park = nxt < total ? chain_at(nxt % p.max_win, 2) : chain_at(win, 2);

      // ---- first conv of each chain: x1_j = x + b + conv(lrelu x); second conv's operand = lrelu(x1_j).
      // The running sum of the x1_j is parked in TMEM tiles that are idle at that point.
#pragma unroll 1
      for (int jj = 0; jj < 3; ++jj) {
        const int j = chain_at(win, jj), j0 = chain_at(win, 0);
        if (prof) e_t = clock64();
        tc::mbar_wait(&bars[WC1_DONE + j], par);
        if (prof) {
          const long long t1 = clock64();
          e_c1w += t1 - e_t;
          e_t = t1;
        }
        tc::fence_after_sync();
        uint8_t* by = smem + g.off_y[j];
        const int pitch = g.rows_y[j], hy = g.hy[j];
        const int L = seg_rows(seg);
        const uint32_t dst_tile = uint32_t(jj == 2 ? park : j0) * 128u;
#pragma unroll
        for (int m = 0; m < wNT; ++m) {
          const int r = m * 128 + q * 32 + lane;
          const int gi = w0 + r;
          const bool inside = gi >= 0 && gi < L;
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            float v[16], acc[16];
            tc::tmem_ld16(lane_base + uint32_t(j * 128 + m * wC + col0 + 16 * h), v);
            if (jj > 0) tc::tmem_ld16(lane_base + uint32_t(j0 * 128 + m * wC + col0 + 16 * h), acc);
            tc::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 bb = *reinterpret_cast<const float4*>(&sbias[j][col0 + 16 * h + 4 * c]);
              v[4 * c] += xr[m][16 * h + 4 * c] + bb.x;
              v[4 * c + 1] += xr[m][16 * h + 4 * c + 1] + bb.y;
              v[4 * c + 2] += xr[m][16 * h + 4 * c + 2] + bb.z;
              v[4 * c + 3] += xr[m][16 * h + 4 * c + 3] + bb.w;
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = jj > 0 ? acc[c] + v[c] : v[c];
            tc::tmem_st16(lane_base + dst_tile + uint32_t(m * wC + col0 + 16 * h), acc);
            store_ops(by, pitch, r + hy, h, v, inside);
          }
        }
        tc::tmem_st_wait();
        arrive(WY_READY + j);
        if (prof) e_c1 += clock64() - e_t;
      }
      if (prof) e_t = clock64();

      // ---- next window's input while this window's second convs run ----
      if (nxt < total) {
        load_x(nxt);
        arrive(WX_READY);
      }

      // ---- out = (sum_j x1_j + S + late bias) / nk ----
      if (prof) {
        const long long t1 = clock64();
        e_lx += t1 - e_t;
        e_t = t1;
      }
      tc::mbar_wait(&bars[WC2_DONE], par);
      if (prof) {
        const long long t1 = clock64();
        e_c2w += t1 - e_t;
        e_t = t1;
      }
      tc::fence_after_sync();
      {
        const int L = seg_rows(seg);
        const long long base = (long long)p.seg_off[seg] * p.scale;
#pragma unroll
        for (int m = 0; m < wNT; ++m) {
          const int r = m * 128 + q * 32 + lane;
          const int gi = w0 + r;
          const bool store = r >= p.H && r < wR - p.H && gi < L;
          float* dst = p.out + (base + gi) * wC + col0;
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            float v[16], acc[16];
            tc::tmem_ld16(lane_base + wS0 + uint32_t(m * wC + col0 + 16 * h), v);
            tc::tmem_ld16(lane_base + uint32_t(park * 128 + m * wC + col0 + 16 * h), acc);
            tc::tmem_ld_wait();
            if (m == wNT - 1 && h == NH - 1) {  // the parked sum and S are consumed
              tc::fence_before_sync();
              __syncwarp();
              if (lane == 0) tc::mbar_arrive(&bars[WF_DONE]);
            }
            if (store) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float4 bb = *reinterpret_cast<const float4*>(&sbias[3][col0 + 16 * h + 4 * c]);
                v[4 * c] = (v[4 * c] + acc[4 * c] + bb.x) * p.inv_nk;
                v[4 * c + 1] = (v[4 * c + 1] + acc[4 * c + 1] + bb.y) * p.inv_nk;
                v[4 * c + 2] = (v[4 * c + 2] + acc[4 * c + 2] + bb.z) * p.inv_nk;
                v[4 * c + 3] = (v[4 * c + 3] + acc[4 * c + 3] + bb.w) * p.inv_nk;
              }
              stg256(dst + 16 * h, v);
              stg256(dst + 16 * h + 8, v + 8);
            }
          }
        }
      }
      if (prof) e_fin += clock64() - e_t;
      idx = nxt;
    }
    if (prof) {
      atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 8), (unsigned long long)(clock64() - e_start));
      atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 9), (unsigned long long)e_c1w);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 10), (unsigned long long)e_c1);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 11), (unsigned long long)e_lx);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 12), (unsigned long long)e_c2w);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 13), (unsigned long long)e_fin);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == wIssuer) tc::tmem_dealloc<512>(tmem);
}

static int mrf_ws_slots(const MrfParams& p, size_t* smem_out) {
  int optin = 227 * 1024, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t budget = size_t(optin) - 2048 - 4 * wSegTable - 128;  // static shared memory (barriers, biases, row table) + alignment
  const WGeo g0 = make_wgeo(p, 0);
  if (g0.total + 3 * wSlotBytes > budget) return 0;
  int n = int((budget - g0.total) / wSlotBytes);
  if (n > wMaxSlots) n = wMaxSlots;
  if (smem_out) *smem_out = make_wgeo(p, n).total + 128;
  return n;
}

bool mrf_ws_supported(const MrfParams& p, int C) {
  if (C != wC || p.nk != 3 || p.nd != 2) return false;
  for (int j = 0; j < 3; ++j)
    if (p.k[j] < 1 || p.k[j] > 11 || !(p.k[j] & 1)) return false;
  if (2 * p.HX * wCH > 8 * 32) return false;  // halo items: one per epilogue thread (8 or 16 epilogue warps)
  if (wR - 2 * p.H < 64) return false;
  return mrf_ws_slots(p, nullptr) >= 3;
}

void launch_mrf_ws(const MrfParams& p_in, int fmt, int n_seg, int max_len, cudaStream_t st) {
  MrfParams p = p_in;

  p.stride = wR - 2 * p.H;
  const int L = max_len * p.scale;
  p.n_seg = n_seg;
  p.max_win = (L + p.stride - 1) / p.stride;
  if (p.max_win <= 0 || n_seg <= 0) return;
  size_t smem = 0;
  p.nslot = mrf_ws_slots(p, &smem);
  if (p.nslot < 3) throw std::runtime_error("mrf_ws: shared memory budget");
  static const int n_sm = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  static const int new_warps = [] { const char* e = getenv("M3B200_MRF64_WARPS"); return e ? atoi(e) : 8; }();
  const bool w16 = new_warps == 16;
  const void* kern = w16 ? (fmt ? reinterpret_cast<const void*>(mrf_ws_kernel<1, 16>) : reinterpret_cast<const void*>(mrf_ws_kernel<0, 16>))
                         : (fmt ? reinterpret_cast<const void*>(mrf_ws_kernel<1, 8>) : reinterpret_cast<const void*>(mrf_ws_kernel<0, 8>));
  ensure_max_dynamic_smem(kern);
  const long long items = (long long)n_seg * p.max_win;
  const int grid = int(items < n_sm ? items : n_sm);
  static const bool want_prof = getenv("M3B200_MRF_PROFILE") != nullptr;
  static long long* d_prof = nullptr;
  if (want_prof) {
    if (!d_prof) cudaMalloc(&d_prof, 16 * sizeof(long long));
    cudaMemsetAsync(d_prof, 0, 16 * sizeof(long long), st);
    p.prof = d_prof;
  }
  const int threads = 32 * ((w16 ? 16 : 8) + wIssuers + 1);
  if (w16) {
    if (fmt) mrf_ws_kernel<1, 16><<<grid, threads, smem, st>>>(p);
    else mrf_ws_kernel<0, 16><<<grid, threads, smem, st>>>(p);
  } else {
    if (fmt) mrf_ws_kernel<1, 8><<<grid, threads, smem, st>>>(p);
    else mrf_ws_kernel<0, 8><<<grid, threads, smem, st>>>(p);
  }
  post_launch("mrf_ws_kernel", st);
  if (want_prof) {  // debug only: synchronous read-back of the per-role cycle counters (summed over CTAs)
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost);
    const double w = h[5] > 0 ? double(h[5]) : 1.0;
    fprintf(stderr,
            "[mrf_ws profile] windows %lld grid %d | issuer cycles/window: total %.0f wait_full %.0f wait_x %.0f wait_y %.0f wait_f %.0f | "
            "epilogue warp0: total %.0f wait_c1 %.0f c1_body %.0f load_x %.0f wait_c2 %.0f final %.0f | blocked FULL waits/window %.1f, "
            "loader: EMPTY-blocked cycles/window %.0f (%.1f waits)\n",
            h[5], grid, h[0] / w, h[1] / w, h[2] / w, h[3] / w, h[4] / w, h[8] / w, h[9] / w, h[10] / w, h[11] / w, h[12] / w, h[13] / w,
            h[6] / w, h[14] / w, h[15] / w);
  }
}

}  // namespace m3

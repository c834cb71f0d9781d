// MRF stage (sum of three ResBlock2 over the same input) for C = 64 channels,
// persistent + warp-specialised.  Same math as mrf_tc_kernel (kernels_tc.cu):
//     out = 1/nk * sum_j [ x1_j + conv2_j(lrelu(x1_j)) ],   x1_j = x + conv1_j(lrelu(x))
// but organised like dec_fused_kernel (kernels_tc_dec2.cu):
//   * one CTA per SM loops over (utterance, window) items; warps 0-15 are epilogue warps, warp 16 issues
//     every tcgen05.mma, warp 17 streams the weights tap by tap (8 KB cp.async.bulk blocks, L2-resident)
//     through a ring of shared-memory slots guarded by full/empty mbarriers;
//   * the three resblocks are independent chains: the tensor pipe runs conv1 of chain j+1 while the
//     epilogue warps turn chain j's accumulator into its second conv's operand, and all second convs
//     accumulate into one TMEM tile S (x stays in registers; sum_j x1_j is parked in TMEM tiles that are
//     idle at that point);
//   * the NEXT window's input is fetched and published (lrelu -> fp16 operand) while the second convs of
//     the current window still run, so the issuer never waits for global memory.
// TMEM (512 columns): T_j = [128 j, 128 j + 128) for the two 128-row tiles of chain j, S = [384, 512).
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

namespace {
constexpr int wC = 64, wNT = 2, wR = wNT * 128, wCH = wC / 8, wKS = wC / 16;
constexpr int wEpiWarps = 16, wIssuer = 16, wLoader = 17, wThreads = 32 * 18;
constexpr int wSegTable = 1024;  // per-utterance row counts cached in shared memory (larger batches read global memory)
constexpr uint32_t wTapBytes = wC * wC * 2;
constexpr uint32_t wS0 = 384;
constexpr int wMaxSlots = 12;
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
  o += size_t(nslot) * wTapBytes;
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
}  // namespace

template <int FMT>
__global__ void __launch_bounds__(wThreads, 1) mrf_ws_kernel(MrfParams p) {
  using E = tc::Elem<FMT>;
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
      tc::mbar_init(&bars[i], many ? wEpiWarps : 1);
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
      for (int idx = first; idx < total; idx = next_item(idx)) {
        for (int d = 0; d < 2; ++d)
          for (int j = 0; j < 3; ++j) {
            const uint16_t* src = p.w16 + p.woff[j][d];
            for (int t = 0; t < p.k[j]; ++t) {
              tc::mbar_wait(&bars[WEMPTY + slot], eparity);
              tc::mbar_expect_tx(&bars[WFULL + slot], wTapBytes);
              tc::bulk_g2s(ring + size_t(slot) * wTapBytes, src + size_t(t) * wC * wC, wTapBytes, &bars[WFULL + slot]);
              if (++slot == uint32_t(nslot)) {
                slot = 0;
                eparity ^= 1u;
              }
            }
          }
      }
    }
    __syncwarp();
  } else if (warp == wIssuer) {
    // =================================== MMA issuer ===============================================
    // ONE thread runs the whole schedule (waits included): no per-tap warp reconvergence, no divisions, and
    // the descriptors of a conv differ only by small additive constants in their low word (the 14-bit start
    // address never overflows: shared memory is < 256 KB), so each MMA costs one add and the instruction.
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, wC, FMT);
      const uint64_t b_tmpl = tc::make_desc(tc::smem_u32(ring), uint32_t(wC) * 16u, 128u);
      const uint32_t b_hi = uint32_t(b_tmpl >> 32), b_lo0 = uint32_t(b_tmpl);
      uint32_t slot = 0, fparity = 0u;
      auto conv = [&](uint32_t abase, int rows_in, int halo, int k, int dil, uint32_t dcol, bool acc0) {
        const uint64_t a_tmpl = tc::make_desc(abase, uint32_t(rows_in) * 16u, 128u);
        const uint32_t a_hi = uint32_t(a_tmpl >> 32);
        uint32_t at = uint32_t(a_tmpl) + uint32_t(halo - ((k - 1) / 2) * dil);
        const uint32_t kstep = uint32_t(2 * rows_in);
#pragma unroll 1
        for (int t = 0; t < k; ++t, at += uint32_t(dil)) {
          tc::mbar_wait(&bars[WFULL + slot], fparity);
          tc::fence_after_sync();
          const uint32_t bt = b_lo0 + slot * (wTapBytes >> 4);
#pragma unroll
          for (int ks = 0; ks < wKS; ++ks) {
            const uint64_t bd = (uint64_t(b_hi) << 32) | uint64_t(bt + uint32_t(ks * 2 * wC));
            const uint32_t acc = (ks > 0 || acc0 || t > 0) ? 1u : 0u;
#pragma unroll
            for (int m = 0; m < wNT; ++m) {
              const uint64_t ad = (uint64_t(a_hi) << 32) | uint64_t(at + uint32_t(ks) * kstep + uint32_t(m * 128));
              tc::mma_f16_ss(tmem + dcol + uint32_t(m * wC), ad, bd, idesc, acc);
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
        tc::mbar_wait(&bars[WX_READY], par);
        tc::fence_after_sync();
        for (int j = 0; j < 3; ++j) {
          if (j == 2 && it > 0) {  // T_2 still holds the previous window's sum until its final epilogue has read it
            tc::mbar_wait(&bars[WF_DONE], uint32_t(it - 1) & 1u);
            tc::fence_after_sync();
          }
          conv(tc::smem_u32(bufX), g.rows_x, p.HX, p.k[j], p.dil[j][0], uint32_t(j) * 128u, false);
          tc::mma_commit(&bars[WC1_DONE + j]);
        }
        for (int j = 0; j < 3; ++j) {
          tc::mbar_wait(&bars[WY_READY + j], par);
          tc::fence_after_sync();
          conv(tc::smem_u32(smem + g.off_y[j]), g.rows_y[j], g.hy[j], p.k[j], p.dil[j][1], wS0, j > 0);
        }
        tc::mma_commit(&bars[WC2_DONE]);
      }
    }
    __syncwarp();
  } else {
    // =================================== epilogue warps ===========================================
    // warp w: TMEM lane quarter q = w & 3 (hardware restriction), column group cg = w >> 2 (16 channels)
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
    const int col0 = cg * 16;
    float xr[wNT][16];  // this thread's x (fp32 residual source): rows {m*128 + q*32 + lane}, 16 channels

    auto arrive = [&](int b) {
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[b]);
    };
    // 16 channels of one row -> two 16-byte operand chunks
    auto store_ops = [&](uint8_t* buf, int pitch, int row, const float* v, bool inside) {
#pragma unroll
      for (int c8 = 0; c8 < 2; ++c8) {
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (inside) {
          pk.x = E::pack2(wlrelu(v[8 * c8], 0.1f), wlrelu(v[8 * c8 + 1], 0.1f));
          pk.y = E::pack2(wlrelu(v[8 * c8 + 2], 0.1f), wlrelu(v[8 * c8 + 3], 0.1f));
          pk.z = E::pack2(wlrelu(v[8 * c8 + 4], 0.1f), wlrelu(v[8 * c8 + 5], 0.1f));
          pk.w = E::pack2(wlrelu(v[8 * c8 + 6], 0.1f), wlrelu(v[8 * c8 + 7], 0.1f));
        }
        *reinterpret_cast<uint4*>(buf + (size_t(cg * 2 + c8) * pitch + row) * 16) = pk;
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
        const int r = m * 128 + q * 32 + lane;
        const int gi = w0 + r;
        const bool inside = gi >= 0 && gi < L;
        const float4* src = reinterpret_cast<const float4*>(p.x + (base + gi) * wC + col0);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (inside) t = __ldg(src + e);
          xr[m][4 * e] = t.x;
          xr[m][4 * e + 1] = t.y;
          xr[m][4 * e + 2] = t.z;
          xr[m][4 * e + 3] = t.w;
        }
      }
      float4 ha = make_float4(0.f, 0.f, 0.f, 0.f), hb = ha;
      const int hi = tid;  // halo item: (row rr of the 2*HX halo rows, chunk c8)
      const bool has_halo = hi < 2 * p.HX * wCH;
      int hrow = 0, hc8 = 0;
      if (has_halo) {
        const int rr = hi / wCH;
        hc8 = hi - rr * wCH;
        hrow = rr < p.HX ? rr : wR + rr;  // bufX rows [0, HX) and [R + HX, R + 2 HX)
        const int gi = w0 - p.HX + hrow;
        if (gi >= 0 && gi < L) {
          const float4* src = reinterpret_cast<const float4*>(p.x + (base + gi) * wC + hc8 * 8);
          ha = __ldg(src);
          hb = __ldg(src + 1);
        }
      }
#pragma unroll
      for (int m = 0; m < wNT; ++m) {
        const int r = m * 128 + q * 32 + lane;
        const int gi = w0 + r;
        store_ops(bufX, g.rows_x, r + p.HX, xr[m], gi >= 0 && gi < L);
      }
      if (has_halo) {
        uint4 pk;
        pk.x = E::pack2(wlrelu(ha.x, 0.1f), wlrelu(ha.y, 0.1f));
        pk.y = E::pack2(wlrelu(ha.z, 0.1f), wlrelu(ha.w, 0.1f));
        pk.z = E::pack2(wlrelu(hb.x, 0.1f), wlrelu(hb.y, 0.1f));
        pk.w = E::pack2(wlrelu(hb.z, 0.1f), wlrelu(hb.w, 0.1f));
        *reinterpret_cast<uint4*>(bufX + (size_t(hc8) * g.rows_x + hrow) * 16) = pk;
      }
    };

    if (first < total) {
      load_x(first);
      arrive(WX_READY);
    }
    int it = 0;
    for (int idx = first; idx < total; ++it) {
      const int nxt = next_item(idx);
      const uint32_t par = uint32_t(it) & 1u;
      const int seg = idx / p.max_win, win = idx - seg * p.max_win;
      const int L = seg_rows(seg);
      const long long base = (long long)p.seg_off[seg] * p.scale;
      const int w0 = win * p.stride - p.H;

      // ---- first conv of each chain: x1_j = x + b + conv(lrelu x); second conv's operand = lrelu(x1_j).
      // The running sum of the x1_j lives in TMEM: in T_0 after chains 0 and 1, in T_2 after chain 2
      // (T_0 is overwritten by the next window's first conv before the final epilogue runs, T_2 is not:
      // the issuer waits for WF_DONE before it touches T_2 again).
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        tc::mbar_wait(&bars[WC1_DONE + j], par);
        tc::fence_after_sync();
        uint8_t* by = smem + g.off_y[j];
        const int pitch = g.rows_y[j], hy = g.hy[j];
#pragma unroll
        for (int m = 0; m < wNT; ++m) {
          float v[16], acc[16];
          tc::tmem_ld16(lane_base + uint32_t(j * 128 + m * wC + col0), v);
          if (j > 0) tc::tmem_ld16(lane_base + uint32_t(m * wC + col0), acc);
          tc::tmem_ld_wait();
          const int r = m * 128 + q * 32 + lane;
          const int gi = w0 + r;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 bb = *reinterpret_cast<const float4*>(&sbias[j][col0 + 4 * c]);
            v[4 * c] += xr[m][4 * c] + bb.x;
            v[4 * c + 1] += xr[m][4 * c + 1] + bb.y;
            v[4 * c + 2] += xr[m][4 * c + 2] + bb.z;
            v[4 * c + 3] += xr[m][4 * c + 3] + bb.w;
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = j > 0 ? acc[c] + v[c] : v[c];
          tc::tmem_st16(lane_base + uint32_t((j == 2 ? 256 : 0) + m * wC + col0), acc);
          store_ops(by, pitch, r + hy, v, gi >= 0 && gi < L);
        }
        tc::tmem_st_wait();
        arrive(WY_READY + j);
      }

      // ---- next window's input while this window's second convs run ----
      if (nxt < total) {
        load_x(nxt);
        arrive(WX_READY);
      }

      // ---- out = (sum_j x1_j + S + late bias) / nk ----
      tc::mbar_wait(&bars[WC2_DONE], par);
      tc::fence_after_sync();
      {
#pragma unroll
        for (int m = 0; m < wNT; ++m) {
          float v[16], acc[16];
          tc::tmem_ld16(lane_base + wS0 + uint32_t(m * wC + col0), v);
          tc::tmem_ld16(lane_base + 256u + uint32_t(m * wC + col0), acc);
          tc::tmem_ld_wait();
          if (m == wNT - 1) {  // T_2 and S are consumed
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars[WF_DONE]);
          }
          const int r = m * 128 + q * 32 + lane;
          const int gi = w0 + r;
          if (r >= p.H && r < wR - p.H && gi < L) {
            float4* dst = reinterpret_cast<float4*>(p.out + (base + gi) * wC + col0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 bb = *reinterpret_cast<const float4*>(&sbias[3][col0 + 4 * c]);
              float4 o;
              o.x = (v[4 * c] + acc[4 * c] + bb.x) * p.inv_nk;
              o.y = (v[4 * c + 1] + acc[4 * c + 1] + bb.y) * p.inv_nk;
              o.z = (v[4 * c + 2] + acc[4 * c + 2] + bb.z) * p.inv_nk;
              o.w = (v[4 * c + 3] + acc[4 * c + 3] + bb.w) * p.inv_nk;
              dst[c] = o;
            }
          }
        }
      }
      idx = nxt;
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
  if (g0.total + 4 * wTapBytes > budget) return 0;
  int n = int((budget - g0.total) / wTapBytes);
  if (n > wMaxSlots) n = wMaxSlots;
  if (smem_out) *smem_out = make_wgeo(p, n).total + 128;
  return n;
}

bool mrf_ws_supported(const MrfParams& p, int C) {
  if (C != wC || p.nk != 3 || p.nd != 2) return false;
  for (int j = 0; j < 3; ++j)
    if (p.k[j] < 1 || p.k[j] > 11 || !(p.k[j] & 1)) return false;
  if (2 * p.HX * wCH > wEpiWarps * 32) return false;  // halo items: one per epilogue thread
  if (wR - 2 * p.H < 64) return false;
  return mrf_ws_slots(p, nullptr) >= 4;
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
  if (p.nslot < 4) throw std::runtime_error("mrf_ws: shared memory budget");
  static const int n_sm = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  const void* kern = fmt ? reinterpret_cast<const void*>(mrf_ws_kernel<1>) : reinterpret_cast<const void*>(mrf_ws_kernel<0>);
  ensure_max_dynamic_smem(kern);
  const long long items = (long long)n_seg * p.max_win;
  const int grid = int(items < n_sm ? items : n_sm);
  if (fmt) mrf_ws_kernel<1><<<grid, wThreads, smem, st>>>(p);
  else mrf_ws_kernel<0><<<grid, wThreads, smem, st>>>(p);
  post_launch("mrf_ws_kernel", st);
}

}  // namespace m3

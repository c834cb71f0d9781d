// tcgen05 tensor-core kernels (sm_100a).  Operands: bf16 or fp16 in the interleaved
// no-swizzle layout of tc_common.cuh; accumulators and the residual stream: fp32 in TMEM.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

// =====================================================================================
// Fused MRF stage of the HiFi-GAN generator:
//     out = 1/nk * sum_j ResBlock2_j(x),   ResBlock2(x): for d: x = x + conv_d(lrelu(x, 0.1))
//
// One CTA owns a window of R = NT*128 consecutive samples of one utterance:
//   * lrelu(x) of the window (+halo) is staged once in smem as the MMA A operand (bufX);
//   * per resblock the residual stream lives in TMEM: T <- x + b1 (tcgen05.st), the first
//     conv's taps are tcgen05.mma's accumulating ON TOP of it (residual add for free), the
//     epilogue reads T, writes lrelu(T) as bf16 into bufY, the second conv accumulates into
//     the same T, and T is folded into the running sum S (also TMEM);
//   * a conv tap is the same smem tile with the descriptor start advanced by tap*dil rows.
// Only rows [H, R-H) of a window are exact after the second conv (H = its halo); windows
// therefore advance by R - 2H.  Rows outside the utterance are zeroed in bufX/bufY, which
// reproduces the reference's per-layer zero padding (batch-1 edge semantics).
// =====================================================================================
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// v2 pipeline inside the CTA:
//   * x of the window is read from global exactly twice (once as the 16-bit A operand, once into
//     registers for the three T <- x + b initialisations);
//   * conv weights stream through a 2-deep cp.async ring, one tap-group ahead of the MMAs;
//   * the last tap-group of every conv commits one mbarrier per 128-row tile, so the epilogue of
//     tile m runs while the tensor pipe is still working on tiles m+1.. .
template <int C, int NT, int FMT, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) mrf_tc_kernel(MrfParams p) {
  constexpr int NTHR = NW * 32;
  constexpr int R = NT * 128;
  constexpr int CH = C / 8;        // 16-byte K-chunks per row
  constexpr int HC = C / (NW / 4); // columns per epilogue thread (NW/4 column groups)
  constexpr int NCC = HC / 16;     // 16-column groups per thread
  constexpr int TCOLS_RAW = 2 * NT * C;
  constexpr int TCOLS = TCOLS_RAW <= 32 ? 32 : TCOLS_RAW <= 64 ? 64 : TCOLS_RAW <= 128 ? 128 : TCOLS_RAW <= 256 ? 256 : 512;
  static_assert(TCOLS_RAW <= 512, "TMEM budget");
  using E = tc::Elem<FMT>;

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t gbar[2], tbar[NT];
  __shared__ float sbias[5][C];  // first-conv bias of each resblock (<= 4) + the summed late bias

  const int seg = blockIdx.y;
  const int L = p.seg_len[seg] * p.scale;
  const int o0 = blockIdx.x * p.stride;
  if (o0 >= L) return;
  const long long base = (long long)p.seg_off[seg] * p.scale;
  const int w0 = o0 - p.H;
  const int ROWSX = (R + 2 * p.HX) | 1, ROWSY = (R + 2 * p.HY) | 1;  // odd pitches: conflict-free chunk-major stores
  const uint32_t wb_bytes = uint32_t(p.wg) * C * C * 2;
  uint8_t* bufX = smem;
  uint8_t* bufY = bufX + size_t(CH) * ROWSX * 16;
  uint8_t* wbuf = bufY + size_t(CH) * ROWSY * 16;  // two buffers of wb_bytes

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hhalf = warp >> 2;
  const float* __restrict__ xg = p.x;

  // weight tap-groups in kernel order: (resblock j, conv d, first tap g0)
  auto group_src = [&](int j, int d, int g0) { return p.w16 + p.woff[j][d] + size_t(g0) * C * C; };
  auto prefetch = [&](int j, int d, int g0, int buf) {
    const int ntap = min(p.wg, p.k[j] - g0);
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(group_src(j, d, g0));
    uint4* dst = reinterpret_cast<uint4*>(wbuf + size_t(buf) * wb_bytes);
    const int n16 = ntap * C * C / 8;
    for (int i = tid; i < n16; i += NTHR) cp_async16(dst + i, src + i);
    cp_async_commit();
  };
  prefetch(0, 0, 0, 0);

  for (int i = tid; i < 5 * C; i += NTHR) {
    const int j = i / C, c = i - j * C;
    sbias[j][c] = j < 4 ? (j < p.nk ? p.bias[j][0][c] : 0.f) : p.late_bias[c];
  }
  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_slot);
  if (tid == 0) {
    tc::mbar_init(&gbar[0], 1);
    tc::mbar_init(&gbar[1], 1);
    for (int m = 0; m < NT; ++m) tc::mbar_init(&tbar[m], 1);
    tc::mbar_fence_init();
  }
  // ---- x window -> registers (fp32, for the residual stream) ---------------------------------
  float xr[NT][NCC][16];
#pragma unroll
  for (int m = 0; m < NT; ++m) {
    const int g = w0 + m * 128 + q * 32 + lane;
    const bool inside = g >= 0 && g < L;
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) {
      const int col = hhalf * HC + cc * 16;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inside && !(p.dbg & 8)) t = *reinterpret_cast<const float4*>(xg + (base + g) * C + col + e * 4);
        xr[m][cc][e * 4 + 0] = t.x;
        xr[m][cc][e * 4 + 1] = t.y;
        xr[m][cc][e * 4 + 2] = t.z;
        xr[m][cc][e * 4 + 3] = t.w;
      }
    }
  }
  // ---- lrelu(x) of the window (+halo) as the 16-bit A operand ----------------------------------
  // item = (row, 8-channel chunk), chunk fastest (coalesced reads); four items in flight per thread
  {
    const int items = CH * ROWSX;
    auto lr = [](float v) { return v >= 0.f ? v : 0.1f * v; };
    for (int i0 = tid; i0 < items; i0 += 4 * NTHR) {
      float4 a[4], b[4];
      int dsti[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i0 + u * NTHR;
        a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        dsti[u] = -1;
        if (idx < items) {
          const int rr = idx / CH, c8 = idx - rr * CH;
          dsti[u] = c8 * ROWSX + rr;
          const int g = w0 - p.HX + rr;
          if (g >= 0 && g < L && !(p.dbg & 8)) {
            const float* src = xg + (base + g) * C + c8 * 8;
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
        *reinterpret_cast<uint4*>(bufX + size_t(dsti[u]) * 16) = pk;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
  const uint32_t T0 = 0, S0 = NT * C;  // column offsets of the two TMEM regions
  const uint32_t idesc = tc::make_idesc(128, C, FMT);
  uint32_t gphase[2] = {0u, 0u}, tphase = 0;  // gbar[b] guards weight buffer b (one pending arrival max)
  int gi = 0;                 // running tap-group index (selects the weight buffer)
  bool prev_nonlast = false;  // previous group committed to gbar and has not been awaited yet

  for (int j = 0; j < p.nk; ++j) {
    const int k = p.k[j];
    const int half = (k - 1) / 2;
    // ---- T <- x + bias of the first conv (registers -> TMEM) -----------------------------------
    {
      const float* b0 = sbias[j];  // warp-uniform addresses: broadcast LDS, no global latency
#pragma unroll
      for (int m = 0; m < ((p.dbg & 16) ? 0 : NT); ++m) {
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) {
          const int col = hhalf * HC + cc * 16;
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = xr[m][cc][e] + b0[col + e];
          tc::tmem_st16(lane_base + T0 + m * C + col, v);
        }
      }
      tc::tmem_st_wait();
    }
    for (int d = 0; d < p.nd; ++d) {
      const int dil = p.dil[j][d];
      const uint8_t* inbuf = d == 0 ? bufX : bufY;
      const int rows_in = d == 0 ? ROWSX : ROWSY;
      const int halo_in = d == 0 ? p.HX : p.HY;
      for (int g0 = 0; g0 < k; g0 += p.wg, ++gi) {
        const int ntap = min(p.wg, k - g0);
        const bool last_group = g0 + p.wg >= k;
        cp_async_wait_all();
        tc::fence_async_smem();  // bufX / bufY / weight ring writes -> async proxy
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (warp == 0 && tc::elect_one()) {  // warp-uniform branch + elect.sync: no per-MMA waterfall loop
          const uint32_t abase = tc::smem_u32(inbuf), wbase = tc::smem_u32(wbuf + size_t(gi & 1) * wb_bytes);
          // Issue order: tap / k-step outer, tile inner.  Consecutive MMAs then hit DIFFERENT TMEM
          // accumulators: a dependent accumulate chain costs ~115 cycles per instruction (measured,
          // tools/ubench.py), NT independent chains bring it down to ~115/NT.
#pragma unroll 1
          for (int t = 0; t < ((p.dbg & 1) ? 0 : ntap); ++t) {
            const int shift = halo_in + (g0 + t - half) * dil;
#pragma unroll
            for (int ks = 0; ks < C / 16; ++ks) {
              const uint64_t bd = tc::make_desc(wbase + uint32_t((t * CH + ks * 2) * C) * 16u, uint32_t(C) * 16u, 128u);
#pragma unroll
              for (int m = 0; m < NT; ++m) {
                const uint64_t ad = tc::make_desc(abase + uint32_t((ks * 2) * rows_in + m * 128 + shift) * 16u,
                                                  uint32_t(rows_in) * 16u, 128u);
                tc::mma_f16_ss(tmem + T0 + m * C, ad, bd, idesc, 1u);
              }
            }
          }
          if (last_group) {
#pragma unroll
            for (int m = 0; m < NT; ++m) tc::mma_commit(&tbar[m]);
          } else {
            tc::mma_commit(&gbar[gi & 1]);
          }
        }
        // the buffer the NEXT group will land in was last read by the PREVIOUS group
        if (prev_nonlast) {
          const int pb = (gi - 1) & 1;
          tc::mbar_wait(&gbar[pb], gphase[pb]);
          gphase[pb] ^= 1u;
        }
        prev_nonlast = !last_group;
        {  // prefetch the next tap-group of the whole kernel into the other buffer
          int nj = j, nd2 = d, ng = g0 + p.wg;
          if (ng >= k) {
            ng = 0;
            if (++nd2 >= p.nd) {
              nd2 = 0;
              ++nj;
            }
          }
          if (nj < p.nk) prefetch(nj, nd2, ng, (gi + 1) & 1);
        }
      }
      // ---- epilogue, tile by tile as the per-tile barriers fire -----------------------------------
      if (d + 1 < p.nd) {
#pragma unroll 1
        for (int m = 0; m < NT; ++m) {
          tc::mbar_wait(&tbar[m], tphase);
          tc::fence_after_sync();
          const int r = m * 128 + q * 32 + lane;
          const int g = w0 + r;
          const bool inside = g >= 0 && g < L;
          if (p.dbg & 2) continue;
          float v[NCC][16];
#pragma unroll
          for (int cc = 0; cc < NCC; ++cc) tc::tmem_ld16(lane_base + T0 + m * C + hhalf * HC + cc * 16, v[cc]);
          tc::tmem_ld_wait();
#pragma unroll
          for (int cc = 0; cc < NCC; ++cc) {
            const int col = hhalf * HC + cc * 16;
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float a = v[cc][2 * e], b = v[cc][2 * e + 1];
              a = a >= 0.f ? a : 0.1f * a;
              b = b >= 0.f ? b : 0.1f * b;
              pk[e] = inside ? E::pack2(a, b) : 0u;
            }
            uint8_t* dst = bufY + (size_t(col / 8) * ROWSY + r + p.HY) * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dst + size_t(ROWSY) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      } else {
        const bool first = j == 0, last = j == p.nk - 1;
        const float* lbv = sbias[4];
#pragma unroll 1
        for (int m = 0; m < NT; ++m) {
          tc::mbar_wait(&tbar[m], tphase);
          tc::fence_after_sync();
          const int r = m * 128 + q * 32 + lane;
          const int g = w0 + r;
          const bool store = r >= p.H && r < R - p.H && g < L;
          if (p.dbg & 4) continue;
#pragma unroll
          for (int cc = 0; cc < NCC; ++cc) {
            const int col = hhalf * HC + cc * 16;
            float v[16];
            tc::tmem_ld16(lane_base + T0 + m * C + col, v);
            if (!first) {
              float s[16];
              tc::tmem_ld16(lane_base + S0 + m * C + col, s);
              tc::tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] += s[e];
            } else {
              tc::tmem_ld_wait();
            }
            if (!last) {
              tc::tmem_st16(lane_base + S0 + m * C + col, v);
            } else if (store) {
              float4* dst = reinterpret_cast<float4*>(p.out + (base + g) * C + col);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float4 o;
                o.x = (v[e * 4 + 0] + lbv[col + e * 4 + 0]) * p.inv_nk;
                o.y = (v[e * 4 + 1] + lbv[col + e * 4 + 1]) * p.inv_nk;
                o.z = (v[e * 4 + 2] + lbv[col + e * 4 + 2]) * p.inv_nk;
                o.w = (v[e * 4 + 3] + lbv[col + e * 4 + 3]) * p.inv_nk;
                dst[e] = o;
              }
            }
          }
        }
        if (!last) tc::tmem_st_wait();
      }
      tphase ^= 1u;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

template <int C, int NT, int FMT, int NW, int MINB>
static void launch_mrf_inst(const MrfParams& p, int n_seg, int max_len, cudaStream_t st) {
  const int R = NT * 128;
  MrfParams q = p;
  q.stride = R - 2 * p.H;
  if (q.stride <= 0) throw std::runtime_error("mrf_tc: receptive field exceeds the window");
  const size_t smem = size_t(C / 8) * 16 * (size_t((R + 2 * p.HX) | 1) + size_t((R + 2 * p.HY) | 1)) + size_t(2) * p.wg * C * C * 2;
  auto kern = mrf_tc_kernel<C, NT, FMT, NW, MINB>;
  ensure_max_dynamic_smem(reinterpret_cast<const void*>(kern));
  const int L = max_len * p.scale;
  dim3 grid((L + q.stride - 1) / q.stride, n_seg);
  kern<<<grid, NW * 32, smem, st>>>(q);
  post_launch("mrf_tc_kernel", st);
}

bool mrf_tc_supported(int C, int nk, int nd, const int* k, int max_halo) {
  if (!(C == 32 || C == 64 || C == 128)) return false;
  if (nk < 1 || nk > 4 || nd < 1 || nd > 2) return false;
  for (int j = 0; j < nk; ++j)
    if (k[j] < 1 || k[j] > 11 || !(k[j] & 1)) return false;
  return max_halo <= 60;
}

void launch_mrf_tc(const MrfParams& p, int C, int fmt, int n_seg, int max_len, cudaStream_t st) {
  MrfParams q = p;
  static const int dbg = [] { const char* e = getenv("M3B200_MRF_DEBUG"); return e ? atoi(e) : 0; }();
  q.dbg = dbg;
  auto pick_wg = [&](int bytes_budget) {
    int kmax = 1;
    for (int j = 0; j < p.nk; ++j) kmax = max(kmax, p.k[j]);
    return max(1, min(kmax, bytes_budget / (C * C * 2)));
  };
#define M3_MRF(CC, NT, NW, MINB)                                                          \
  {                                                                                       \
    if (fmt == 1) launch_mrf_inst<CC, NT, 1, NW, MINB>(q, n_seg, max_len, st);            \
    else launch_mrf_inst<CC, NT, 0, NW, MINB>(q, n_seg, max_len, st);                     \
  }
  static const int nt32 = [] { const char* e = getenv("M3B200_MRF_NT32"); return e ? atoi(e) : 4; }();
  static const int mrf_warps = [] { const char* e = getenv("M3B200_MRF_WARPS"); return e ? atoi(e) : 8; }();
  static const int nt64 = [] { const char* e = getenv("M3B200_MRF_NT64"); return e ? atoi(e) : 2; }();
  if (C == 32) {
    q.wg = pick_wg(16 * 1024);  // whole conv (<= 14 KB) per buffer; ~100 KB/CTA -> 2 CTAs/SM
    if (nt32 == 2) M3_MRF(32, 2, 4, 4)       // 128-thread CTAs, 4 per SM (more latency chains in flight)
    else if (nt32 == 3) M3_MRF(32, 3, 4, 3)  // 128-thread CTAs, 3 per SM
    else M3_MRF(32, 4, 8, 2)
  } else if (C == 64) {
    q.wg = pick_wg(16 * 1024);  // 2 taps per buffer; ~109 KB/CTA -> 2 CTAs/SM
    if (nt64 == 4) M3_MRF(64, 4, 8, 1)
    else if (mrf_warps == 16) M3_MRF(64, 2, 16, 2)
    else M3_MRF(64, 2, 8, 2)
  } else if (C == 128) {
    q.wg = pick_wg(32 * 1024);  // 1 tap per buffer
    if (mrf_warps == 16) M3_MRF(128, 2, 16, 1)
    else M3_MRF(128, 2, 8, 1)
  } else {
    throw std::runtime_error("mrf_tc: unsupported channel count");
  }
#undef M3_MRF
}

}  // namespace m3

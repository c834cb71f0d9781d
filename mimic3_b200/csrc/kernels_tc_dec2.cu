// Fused LAST stage of the HiFi-GAN generator, second generation: persistent + warp-specialised.
//     x   = ConvTranspose1d(lrelu(y_prev, 0.1))                       (stride u, kernel k = 2u)
//     out = 1/nk * sum_j ResBlock2_j(x)                               (MRF, nk = 3)
//     y   = tanh(conv_post(lrelu(out, 0.01)))  (+ per-utterance max|y| for the int16 scaling)
// (reference graph: the tail of generator.onnx run by
//  mimic3_tts/voice.py:230.)  Differences to dec_last_kernel (kernels_tc_dec.cu):
//   * one CTA per SM loops over (utterance, window) work items; ALL weights of the stage (~100 KB of
//     fp16) are pulled into shared memory once per CTA with cp.async.bulk instead of once per window;
//   * roles: warps 0-7 epilogue (TMEM -> registers -> activations back to smem), warp 8 issues every
//     tcgen05.mma, warps 9-11 stage the next window's input; they only meet on mbarriers, so the tensor
//     pipe runs resblock j+1 while the epilogue warps turn resblock j's accumulator into its next
//     operand, and the next window's transposed conv / this window's conv_post fill the remaining gaps;
//   * the transposed conv is POLYPHASE: D[t, ph*C+co] = sum_{d=0,1} lrelu(y[t-d]) . W[:, co, u*d+ph], one
//     M=128 x N=u*C GEMM over y_prev rows -- no zero-stuffed rows, 1/7 of the MMA time.  Its output is
//     transposed into sample order through an fp32 staging tile that aliases the (then idle) resblock
//     operand buffers;
//   * no running-sum round trips through TMEM: x and sum_j x1_j live in registers, all second convs
//     accumulate into one TMEM tile S.
// TMEM (512 columns): T_j = [96 j, 96 j + 96) conv1 accumulators, S = [288, 384) (conv_post reuses it),
// D = [384, 512) transposed-conv result.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <type_traits>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

namespace {
constexpr int kC = 32, kNT = 3, kR = kNT * 128, kCH = kC / 8;
constexpr int kLoaders = 3;  // warp roles: NEW epilogue warps (8 or 16), one issuer, kLoaders loader warps
constexpr int kLoadDepth = 9;   // y_prev rows-chunks in flight per loader lane (one round for u = 4)
constexpr uint32_t kT0 = 0, kS0 = 288, kD0 = 384;
constexpr int kSegTable = 1024;
enum Bar { W_FULL = 0, A_FULL, U_DONE, X_READY, C1_DONE, Y_READY = C1_DONE + 3, C2_DONE = Y_READY + 3, O_READY, P_DONE, NBAR };

struct Geo {
  int rows_x, rows_y[3], rows_a, nmu;
  size_t off_w, off_x, off_y[3], off_a, total;
};
__host__ __device__ inline Geo make_geo(const DecFusedParams& p) {
  Geo g;
  g.rows_x = (kR + 2 * p.HX) | 1;
  for (int j = 0; j < 3; ++j) g.rows_y[j] = (kR + 2 * p.HYb[j]) | 1;
  g.nmu = (kR / p.up_u + 1 + 127) / 128;
  g.rows_a = (g.nmu * 128 + 1) | 1;
  size_t o = 0;
  g.off_w = o;
  o += (size_t(p.w_bytes) + 127) & ~size_t(127);
  g.off_x = o;
  o += size_t(kCH) * g.rows_x * 16;
  for (int j = 0; j < 3; ++j) {
    g.off_y[j] = o;
    o += size_t(kCH) * g.rows_y[j] * 16;
  }
  g.off_a = o;
  o += size_t(p.cin / 8) * g.rows_a * 16;
  g.total = o;
  return g;
}

template <int NTHR>
__device__ __forceinline__ void epi_bar_n() { asm volatile("bar.sync 1, %0;\n" ::"n"(NTHR) : "memory"); }
__device__ __forceinline__ float lrelu(float v, float s) { return fmaxf(v, s * v); }
}  // namespace

// NEW = epilogue warps: 8 (16 of the 32 channels per thread) or 16 (8 channels per thread: half the dependent
// tmem_ld -> math -> st.shared chain per thread and twice the warps to hide its latency; the M3B200_DEC_PROFILE counters
// show the epilogue warps are busy 86 % of a window while the issuer waits for them 28 % of it)
template <int FMT, int NEW>
__global__ void __maxnreg__(NEW == 16 ? 96 : 168) dec_fused_kernel(DecFusedParams p) {
  using E = tc::Elem<FMT>;
  constexpr int kEpiWarps = NEW, kIssuer = NEW, kLoader = NEW + 1, kThreads = 32 * (kLoader + kLoaders);
  constexpr int NCG = NEW / 4, G = kC / NCG;  // column groups and channels per thread (16 or 8)
  constexpr int kLD = NEW == 16 ? 5 : kLoadDepth;  // loads in flight per loader lane (96 registers per thread with 16 epilogue warps)
  auto epi_bar = [] { epi_bar_n<32 * NEW>(); };
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[NBAR];
  __shared__ __align__(16) float sbias[5][kC];  // [0] up bias, [1..3] first-conv bias of resblock j, [4] summed second-conv bias

  const Geo g = make_geo(p);
  uint8_t* const wts = smem + g.off_w;
  uint8_t* const bufX = smem + g.off_x;
  uint8_t* const bufA = smem + g.off_a;
  uint8_t* const XS = smem + g.off_y[1];  // fp32 [kR][32] staging, aliases bufY1 + bufY2
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u = p.up_u, CHI = p.cin / 8, NUP = u * kC;

  // ---- work items: (utterance, window), the same deterministic sequence in every role ------------
  __shared__ int s_rows[kSegTable];  // per-utterance sample counts (larger batches fall back to global memory)
  auto seg_rows = [&](int seg) { return seg < kSegTable ? s_rows[seg] : p.seg_len[seg] * p.scale; };
  for (int i = threadIdx.x; i < p.n_seg && i < kSegTable; i += kThreads) s_rows[i] = p.seg_len[i] * p.scale;
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

  // ---- one-time setup ----------------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < NBAR; ++i) {
      const bool many = i == X_READY || (i >= Y_READY && i < Y_READY + 3) || i == O_READY;
      tc::mbar_init(&bars[i], many ? kEpiWarps : (i == A_FULL ? kLoaders : 1));
    }
    tc::mbar_fence_init();
  }
  for (int i = tid; i < 5 * kC; i += kThreads) {
    const int j = i / kC, c = i - j * kC;
    float v;
    if (j == 0) v = p.up_bias[c];
    else if (j <= 3) v = p.bias1[j - 1][c];
    else v = p.late_bias[c];
    sbias[j][c] = v;
  }
  {  // activations start as zeros: conv halos that no epilogue ever writes stay zero for the whole kernel
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
#ifdef M3B200_KERNEL_PROFILE  // per-role cycle counters (M3B200_DEC_PROFILE=1); compiled out by default: +9 % kernel time when merely present
  const bool prof = p.prof != nullptr;
#else
  constexpr bool prof = false;
#endif
  auto timed_wait = [&](uint64_t* bar, uint32_t parity, long long& acc) {
    if (prof && !tc::mbar_test(bar, parity)) {  // only waits that actually block are timed
      const long long t = clock64();
      tc::mbar_wait(bar, parity);
      acc += clock64() - t;
    } else {
      tc::mbar_wait(bar, parity);
    }
  };

  if (warp >= kLoader) {
    // =================================== loader warp ===============================================
    if (first < total) {
      if (warp == kLoader && tc::elect_one()) {
        tc::mbar_expect_tx(&bars[W_FULL], p.w_bytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wblob);
        for (uint32_t o = 0; o < p.w_bytes; o += 32768u) {
          const uint32_t n = p.w_bytes - o < 32768u ? p.w_bytes - o : 32768u;
          tc::bulk_g2s(wts + o, src + o, n, &bars[W_FULL]);
        }
      }
      __syncwarp();
      const int rows_need = kR / u + 3 < g.rows_a ? kR / u + 3 : g.rows_a;
      const int items = rows_need * CHI;
      int it = 0;
      for (int idx = first; idx < total; idx = next_item(idx), ++it) {
        if (it > 0) tc::mbar_wait(&bars[U_DONE], uint32_t(it - 1) & 1u);  // previous window's transposed conv has read bufA
        const int seg = idx / p.max_win, win = idx - seg * p.max_win;
        const int Lprev = seg_rows(seg) / p.scale * p.prev_scale;
        const long long base_prev = (long long)p.seg_off[seg] * p.prev_scale;
        const int w0 = win * p.stride - p.H;
        const int e = w0 + p.up_pad;
        const int tq0 = e >= 0 ? e / u : -((-e + u - 1) / u);
        const int lt = (warp - kLoader) * 32 + lane;
        for (int i0 = lt; i0 < items; i0 += 32 * kLoaders * kLD) {  // all global loads of a round in flight together
          float4 a[kLD], b[kLD];
          bool ok[kLD];
#pragma unroll
          for (int k = 0; k < kLD; ++k) {
            const int i = i0 + 32 * kLoaders * k;
            const int ra = i / CHI, c8 = i - ra * CHI;
            const int t = tq0 - 1 + ra;
            ok[k] = i < items && t >= 0 && t < Lprev;
            if (ok[k]) {
              const float4* src = reinterpret_cast<const float4*>(p.yprev + (base_prev + t) * (long long)p.cin + c8 * 8);
              a[k] = __ldg(src);
              b[k] = __ldg(src + 1);
            }
          }
#pragma unroll
          for (int k = 0; k < kLD; ++k) {
            const int i = i0 + 32 * kLoaders * k;
            if (i >= items) continue;
            const int ra = i / CHI, c8 = i - ra * CHI;
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (ok[k]) {
              pk.x = E::pack2(lrelu(a[k].x, 0.1f), lrelu(a[k].y, 0.1f));
              pk.y = E::pack2(lrelu(a[k].z, 0.1f), lrelu(a[k].w, 0.1f));
              pk.z = E::pack2(lrelu(b[k].x, 0.1f), lrelu(b[k].y, 0.1f));
              pk.w = E::pack2(lrelu(b[k].z, 0.1f), lrelu(b[k].w, 0.1f));
            }
            *reinterpret_cast<uint4*>(bufA + (size_t(c8) * g.rows_a + ra) * 16) = pk;
          }
        }
        tc::fence_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[A_FULL]);
      }
    }
  } else if (warp == kIssuer) {
    // =================================== MMA issuer warp ===========================================
    if (first < total) {
      const uint32_t wbase = tc::smem_u32(wts);
      const uint32_t idC = tc::make_idesc(128, kC, FMT), idU = tc::make_idesc(128, NUP, FMT), idP = tc::make_idesc(128, 16, FMT);
      // one conv = taps x (K/16) x ntile MMAs; a tap is a descriptor start-address shift.  The descriptors of
      // one conv differ only in their 14-bit start-address field, so the (k-step, tile) grid of each tap is
      // fully unrolled: independent address adds instead of a serial uniform-datapath loop per MMA
      // (the rolled loop issued one MMA per ~100 cycles, 2.5x slower than the tensor pipe executes them).
      auto issue = [&](auto ks_tag, auto nt_tag, uint32_t abase, int rows_in, int halo, const DecFusedConv& cv, int N,
                       uint32_t dcol, int dstep, uint32_t idesc, bool acc0) {
        constexpr int KS = decltype(ks_tag)::value, NTL = decltype(nt_tag)::value;
        const uint64_t a_tmpl = tc::make_desc(abase, uint32_t(rows_in) * 16u, 128u);
        const uint64_t b_tmpl = tc::make_desc(wbase + cv.woff, uint32_t(N) * 16u, 128u);
        const uint32_t a_hi = uint32_t(a_tmpl >> 32), b_hi = uint32_t(b_tmpl >> 32);
        uint32_t at = uint32_t(a_tmpl) + uint32_t(halo - cv.pad_left * cv.dil), bt = uint32_t(b_tmpl);
        const uint32_t bstep = uint32_t(2 * KS * N);  // 16-byte units per tap: (K/8) * N
        const uint32_t kstep = uint32_t(2 * rows_in);
#pragma unroll 1
        for (int t = 0; t < cv.taps; ++t, at += uint32_t(cv.dil), bt += bstep) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t bd = (uint64_t(b_hi) << 32) | uint64_t(bt + uint32_t(ks * 2 * N));
            const uint32_t acc = (ks > 0 || acc0 || t > 0) ? 1u : 0u;
#pragma unroll
            for (int m = 0; m < NTL; ++m) {
              const uint64_t ad = (uint64_t(a_hi) << 32) | uint64_t(at + uint32_t(ks) * kstep + uint32_t(m * 128));
              tc::mma_f16_ss(tmem + dcol + uint32_t(m * dstep), ad, bd, idesc, acc);
            }
          }
        }
      };
      using I1 = std::integral_constant<int, 1>;
      using I2 = std::integral_constant<int, 2>;
      using I3 = std::integral_constant<int, 3>;
      using I4 = std::integral_constant<int, 4>;
      auto issue_up = [&] {  // K = cin (64 -> 4 k-steps, else generic), 1 or 2 row tiles
        if (p.cin == 64 && g.nmu == 1) issue(I4{}, I1{}, tc::smem_u32(bufA), g.rows_a, 1, p.up, NUP, kD0, NUP, idU, false);
        else if (p.cin == 64) issue(I4{}, I2{}, tc::smem_u32(bufA), g.rows_a, 1, p.up, NUP, kD0, NUP, idU, false);
        else {
          const uint64_t a_tmpl = tc::make_desc(0u, uint32_t(g.rows_a) * 16u, 128u);
          const uint64_t b_tmpl = tc::make_desc(0u, uint32_t(NUP) * 16u, 128u);
          for (int t = 0; t < p.up.taps; ++t)
            for (int k0 = 0; k0 < p.cin / 16; ++k0)
              for (int m = 0; m < g.nmu; ++m) {
                const uint32_t a = (tc::smem_u32(bufA) >> 4) + uint32_t(1 - t + k0 * 2 * g.rows_a + m * 128);
                const uint32_t b = ((wbase + p.up.woff) >> 4) + uint32_t((t * (p.cin / 8) + k0 * 2) * NUP);
                tc::mma_f16_ss(tmem + kD0 + uint32_t(m * NUP), a_tmpl | uint64_t(a & 0x3FFFu), b_tmpl | uint64_t(b & 0x3FFFu), idU,
                               (k0 || t) ? 1u : 0u);
              }
        }
      };
      const uint32_t aX = tc::smem_u32(bufX);
      // ONE elected thread runs the whole schedule, waits included: no warp reconvergence between convs
      if (tc::elect_one()) {
        long long c_x = 0, c_y = 0, c_a = 0, c_o = 0;
        tc::mbar_wait(&bars[W_FULL], 0u);
        tc::mbar_wait(&bars[A_FULL], 0u);
        const long long c_start = prof ? clock64() : 0;
        tc::fence_after_sync();
        issue_up();
        tc::mma_commit(&bars[U_DONE]);
        int it = 0;
        for (int idx = first; idx < total; ++it) {
          const int nxt = next_item(idx);
          const uint32_t par = uint32_t(it) & 1u;
          timed_wait(&bars[X_READY], par, c_x);
          tc::fence_after_sync();
          for (int j = 0; j < 3; ++j) {
            issue(I2{}, I3{}, aX, g.rows_x, p.HX, p.c1[j], kC, kT0 + uint32_t(j) * 96u, kC, idC, false);
            tc::mma_commit(&bars[C1_DONE + j]);
          }
          for (int j = 0; j < 3; ++j) {
            timed_wait(&bars[Y_READY + j], par, c_y);
            tc::fence_after_sync();
            issue(I2{}, I3{}, tc::smem_u32(smem + g.off_y[j]), g.rows_y[j], p.HYb[j], p.c2[j], kC, kS0, kC, idC, j > 0);
          }
          tc::mma_commit(&bars[C2_DONE]);
          if (nxt < total) {  // next window's transposed conv fills the gap while the epilogue reduces this one
            timed_wait(&bars[A_FULL], uint32_t(it + 1) & 1u, c_a);
            tc::fence_after_sync();
            issue_up();
            tc::mma_commit(&bars[U_DONE]);
          }
          timed_wait(&bars[O_READY], par, c_o);
          tc::fence_after_sync();
          issue(I2{}, I3{}, tc::smem_u32(smem + g.off_y[0]), g.rows_y[0], p.HYb[0], p.post, 16, kS0, 16, idP, false);
          tc::mma_commit(&bars[P_DONE]);
          idx = nxt;
        }
        if (prof) {
          unsigned long long* q = reinterpret_cast<unsigned long long*>(p.prof);
          atomicAdd(q + 0, (unsigned long long)(clock64() - c_start));
          atomicAdd(q + 1, (unsigned long long)c_x);
          atomicAdd(q + 2, (unsigned long long)c_y);
          atomicAdd(q + 3, (unsigned long long)c_a);
          atomicAdd(q + 4, (unsigned long long)c_o);
          atomicAdd(q + 5, (unsigned long long)it);
        }
      }
      __syncwarp();
    }
  } else {
    // =================================== epilogue warps ============================================
    const int q = warp & 3, hhalf = warp >> 2;  // hhalf: column group 0..NCG-1
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
    const int col0 = hhalf * G;
    float xr[kNT][G], xs[kNT][G];

    auto store_y = [&](uint8_t* buf, int pitch, int row, const uint32_t* pk) {  // G columns = G / 8 chunks of 8
      uint8_t* dst = buf + (size_t(col0 / 8) * pitch + row) * 16;
      *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if constexpr (G == 16) *reinterpret_cast<uint4*>(dst + size_t(pitch) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };
    auto arrive = [&](int b) {
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[b]);
    };
    long long e_u = 0, e_p = 0, e_c1 = 0, e_c2 = 0;
    const long long e_start = prof ? clock64() : 0;
    auto post_epi = [&](int seg, int w0, int L, long long base, uint32_t par) {
      timed_wait(&bars[P_DONE], par, e_p);
      tc::fence_after_sync();
      if (hhalf == 0) {
        float mx = 0.f;
#pragma unroll
        for (int m = 0; m < kNT; ++m) {
          float v[8];
          tc::tmem_ld8(lane_base + kS0 + uint32_t(m * 16), v);
          tc::tmem_ld_wait();
          const int r = m * 128 + q * 32 + lane;
          const int gi = w0 + r;
          if (r >= p.H && r < kR - p.H && gi < L) {
            const float y = tanhf(v[0]);
            p.audio[base + gi] = y;
            mx = fmaxf(mx, fabsf(y));
          }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0 && mx > 0.f) atomicMax(p.peak_bits + seg, __float_as_uint(mx));
      }
      tc::fence_before_sync();
    };

    int it = 0;
    int pseg = 0, pw0 = 0, pL = 0;
    long long pbase = 0;
    for (int idx = first; idx < total; idx = next_item(idx), ++it) {
      const uint32_t par = uint32_t(it) & 1u;
      const int seg = idx / p.max_win, win = idx - seg * p.max_win;
      const int L = seg_rows(seg);
      const long long base = (long long)p.seg_off[seg] * p.scale;
      const int w0 = win * p.stride - p.H;
      const int e = w0 + p.up_pad;
      const int tq0 = e >= 0 ? e / u : -((-e + u - 1) / u);
      const int ph0 = e - tq0 * u;

      // ---- transposed-conv epilogue: D[t][ph*C + co] -> x in sample order ----
      // pass 1 (lane = y_prev row): x = D + b into the fp32 staging tile, 16-byte chunks XOR-swizzled so that
      // both this strided write (rows u apart) and the row-per-lane read below are bank-conflict free
      timed_wait(&bars[U_DONE], par, e_u);
      tc::fence_after_sync();
      for (int mt = 0; mt < g.nmu; ++mt) {
        const int lq = mt * 128 + q * 32 + lane;
        for (int pp = 0; pp < u / NCG; ++pp) {
          const int ph = hhalf * (u / NCG) + pp;
          float v[32];
          tc::tmem_ld16(lane_base + kD0 + uint32_t(mt * NUP + ph * kC), v);
          tc::tmem_ld16(lane_base + kD0 + uint32_t(mt * NUP + ph * kC + 16), v + 16);
          tc::tmem_ld_wait();
          const int r = u * lq + ph - ph0;
          if (r >= 0 && r < kR) {
            const int key = (r & 7) ^ ((r >> 3) & 3);
            uint8_t* xrow = XS + size_t(r) * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 bb = *reinterpret_cast<const float4*>(&sbias[0][4 * c]);
              float4 f;
              f.x = v[4 * c] + bb.x;
              f.y = v[4 * c + 1] + bb.y;
              f.z = v[4 * c + 2] + bb.z;
              f.w = v[4 * c + 3] + bb.w;
              *reinterpret_cast<float4*>(xrow + ((c ^ key) << 4)) = f;
            }
          }
        }
      }
      tc::fence_before_sync();
      epi_bar();  // every x row is staged
      // pass 2 (lane = sample row, as in every later epilogue): keep x in registers, publish lrelu(x) as the
      // first convs' operand
#pragma unroll
      for (int m = 0; m < kNT; ++m) {
        const int r = m * 128 + q * 32 + lane;
        const int gi = w0 + r;
        const bool inside = gi >= 0 && gi < L;
        const int key = (r & 7) ^ ((r >> 3) & 3);
        const uint8_t* xrow = XS + size_t(r) * 128;
#pragma unroll
        for (int c = 0; c < G / 4; ++c) {
          const float4 f = *reinterpret_cast<const float4*>(xrow + (((hhalf * (G / 4) + c) ^ key) << 4));
          xr[m][4 * c] = f.x;
          xr[m][4 * c + 1] = f.y;
          xr[m][4 * c + 2] = f.z;
          xr[m][4 * c + 3] = f.w;
        }
        uint32_t pk[G / 2];
#pragma unroll
        for (int c = 0; c < G / 2; ++c) pk[c] = inside ? E::pack2(lrelu(xr[m][2 * c], 0.1f), lrelu(xr[m][2 * c + 1], 0.1f)) : 0u;
        store_y(bufX, g.rows_x, r + p.HX, pk);
      }
      arrive(X_READY);
      epi_bar();  // staging tile is dead: its bytes are bufY1 / bufY2 again
      for (int j = 1; j < 3; ++j) {  // re-zero the halo rows the staging tile overwrote
        uint8_t* by = smem + g.off_y[j];
        const int hy = p.HYb[j], n = 2 * hy * kCH;
        for (int i = tid; i < n; i += kEpiWarps * 32) {
          const int c8 = i / (2 * hy), rr = i - c8 * 2 * hy;
          const int row = rr < hy ? rr : kR + rr;  // [0, hy) and [kR + hy, kR + 2 hy)
          *reinterpret_cast<uint4*>(by + (size_t(c8) * g.rows_y[j] + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
      }  // (made visible to the tensor pipe by the Y_READY arrivals below)

      // ---- previous window's conv_post result (its MMAs ran during the code above) ----
      if (it > 0) post_epi(pseg, pw0, pL, pbase, uint32_t(it - 1) & 1u);

      // ---- first conv of each resblock: x1 = x + b + conv(lrelu x); operand of the second conv = lrelu(x1) ----
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        timed_wait(&bars[C1_DONE + j], par, e_c1);
        tc::fence_after_sync();
        uint8_t* by = smem + g.off_y[j];
        const int pitch = g.rows_y[j], hy = p.HYb[j];
        float bj[G];
#pragma unroll
        for (int c = 0; c < G / 4; ++c) {
          const float4 f = *reinterpret_cast<const float4*>(&sbias[1 + j][col0 + 4 * c]);
          bj[4 * c] = f.x;
          bj[4 * c + 1] = f.y;
          bj[4 * c + 2] = f.z;
          bj[4 * c + 3] = f.w;
        }
#pragma unroll
        for (int m = 0; m < kNT; ++m) {
          float v[G];
          tc::tmem_ldg<G>(lane_base + kT0 + uint32_t(j * 96 + m * kC + col0), v);
          tc::tmem_ld_wait();
          const int r = m * 128 + q * 32 + lane;
          const int gi = w0 + r;
          const bool inside = gi >= 0 && gi < L;
          uint32_t pk[G / 2];
#pragma unroll
          for (int c = 0; c < G; ++c) {
            v[c] += xr[m][c] + bj[c];
            xs[m][c] = j == 0 ? v[c] : xs[m][c] + v[c];
          }
#pragma unroll
          for (int c = 0; c < G / 2; ++c) pk[c] = inside ? E::pack2(lrelu(v[2 * c], 0.1f), lrelu(v[2 * c + 1], 0.1f)) : 0u;
          store_y(by, pitch, r + hy, pk);
        }
        arrive(Y_READY + j);
      }

      // ---- out = (sum_j x1_j + S + late bias) / nk; operand of conv_post = lrelu(out, 0.01) in bufY0 ----
      timed_wait(&bars[C2_DONE], par, e_c2);
      tc::fence_after_sync();
      {
        uint8_t* by = smem + g.off_y[0];
        const int pitch = g.rows_y[0], hy = p.HYb[0];
        float bl[G];
#pragma unroll
        for (int c = 0; c < G / 4; ++c) {
          const float4 f = *reinterpret_cast<const float4*>(&sbias[4][col0 + 4 * c]);
          bl[4 * c] = f.x;
          bl[4 * c + 1] = f.y;
          bl[4 * c + 2] = f.z;
          bl[4 * c + 3] = f.w;
        }
#pragma unroll
        for (int m = 0; m < kNT; ++m) {
          float v[G];
          tc::tmem_ldg<G>(lane_base + kS0 + uint32_t(m * kC + col0), v);
          tc::tmem_ld_wait();
          const int r = m * 128 + q * 32 + lane;
          const int gi = w0 + r;
          const bool inside = gi >= 0 && gi < L;
          uint32_t pk[G / 2];
#pragma unroll
          for (int c = 0; c < G; ++c) v[c] = lrelu((v[c] + xs[m][c] + bl[c]) * p.inv_nk, 0.01f);
#pragma unroll
          for (int c = 0; c < G / 2; ++c) pk[c] = inside ? E::pack2(v[2 * c], v[2 * c + 1]) : 0u;
          store_y(by, pitch, r + hy, pk);
        }
      }
      arrive(O_READY);
      pseg = seg;
      pw0 = w0;
      pL = L;
      pbase = base;
    }
    if (it > 0) post_epi(pseg, pw0, pL, pbase, uint32_t(it - 1) & 1u);
    if (prof && tid == 0) {
      unsigned long long* q = reinterpret_cast<unsigned long long*>(p.prof);
      atomicAdd(q + 8, (unsigned long long)(clock64() - e_start));
      atomicAdd(q + 9, (unsigned long long)e_u);
      atomicAdd(q + 10, (unsigned long long)e_p);
      atomicAdd(q + 11, (unsigned long long)e_c1);
      atomicAdd(q + 12, (unsigned long long)e_c2);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kIssuer) tc::tmem_dealloc<512>(tmem);
}

size_t dec_fused_smem_bytes(const DecFusedParams& p) { return make_geo(p).total + 128; }

bool dec_fused_supported(int C, int cin, int up_k, int up_u, int nk, int nd, int HX, const int* HYb, size_t w_bytes) {
  if (C != kC || nk != 3 || nd != 2) return false;
  if (cin % 16 || cin > 256 || up_u < 2 || (up_u & 1) || up_k != 2 * up_u) return false;
  const int NUP = up_u * kC, nmu = (kR / up_u + 1 + 127) / 128;
  if (NUP > 256 || nmu * NUP > 128) return false;  // TMEM columns [384, 512)
  DecFusedParams p;
  p.cin = cin;
  p.up_u = up_u;
  p.HX = HX;
  for (int j = 0; j < 3; ++j) p.HYb[j] = HYb[j];
  p.w_bytes = unsigned(w_bytes);
  const Geo g = make_geo(p);
  if (size_t(kCH) * (g.rows_y[1] + g.rows_y[2]) * 16 < size_t(kR) * 128) return false;  // fp32 staging alias
  int HYmax = 0;
  for (int j = 0; j < 3; ++j) HYmax = HYb[j] > HYmax ? HYb[j] : HYmax;
  if (kR - 2 * (HX + HYmax + 3) < 64) return false;
  return g.total + 128 + 1024 + 4 * kSegTable <= size_t(227) * 1024;
}

void launch_dec_fused(const DecFusedParams& p_in, int fmt, int n_seg, int max_len, cudaStream_t st) {
  DecFusedParams p = p_in;
  p.stride = kR - 2 * p.H;
  if (p.stride <= 0) throw std::runtime_error("dec_fused: receptive field exceeds the window");
  const int L = max_len * p.scale;
  p.n_seg = n_seg;
  p.max_win = (L + p.stride - 1) / p.stride;
  if (p.max_win <= 0 || n_seg <= 0) return;
  static const int n_sm = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  const size_t smem = dec_fused_smem_bytes(p);
  static const int new_warps = [] { const char* e = getenv("M3B200_DEC_WARPS2"); return e ? atoi(e) : 16; }();  // 16: 4.58 -> 4.44 ms (r02j, r02m)
  const bool w16 = new_warps == 16 && p.up_u % 4 == 0;  // 16 epilogue warps: one transposed-conv phase per column group
  const void* kern = w16 ? (fmt ? reinterpret_cast<const void*>(dec_fused_kernel<1, 16>) : reinterpret_cast<const void*>(dec_fused_kernel<0, 16>))
                         : (fmt ? reinterpret_cast<const void*>(dec_fused_kernel<1, 8>) : reinterpret_cast<const void*>(dec_fused_kernel<0, 8>));
  ensure_max_dynamic_smem(kern);
  const long long items = (long long)n_seg * p.max_win;
  const int grid = int(items < n_sm ? items : n_sm);
  static const bool want_prof = getenv("M3B200_DEC_PROFILE") != nullptr;
  static long long* d_prof = nullptr;
  if (want_prof) {
    if (!d_prof) cudaMalloc(&d_prof, 16 * sizeof(long long));
    cudaMemsetAsync(d_prof, 0, 16 * sizeof(long long), st);
    p.prof = d_prof;
  }
  const int threads = 32 * ((w16 ? 16 : 8) + 1 + kLoaders);
  if (w16) {
    if (fmt) dec_fused_kernel<1, 16><<<grid, threads, smem, st>>>(p);
    else dec_fused_kernel<0, 16><<<grid, threads, smem, st>>>(p);
  } else {
    if (fmt) dec_fused_kernel<1, 8><<<grid, threads, smem, st>>>(p);
    else dec_fused_kernel<0, 8><<<grid, threads, smem, st>>>(p);
  }
  post_launch("dec_fused_kernel", st);
  if (want_prof) {  // debug only: synchronous read-back of the per-role cycle counters (summed over CTAs)
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost);
    const double n = double(h[5] > 0 ? h[5] : 1);
    fprintf(stderr,
            "[dec_fused profile] windows %lld grid %d | issuer cycles/window: total %.0f wait_x %.0f wait_y %.0f wait_a %.0f wait_o %.0f | "
            "epilogue warp 0: total %.0f wait_u %.0f wait_p %.0f wait_c1 %.0f wait_c2 %.0f\n",
            h[5], grid, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n, h[12] / n);
  }
}

}  // namespace m3

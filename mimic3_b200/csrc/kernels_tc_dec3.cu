// Fused LAST stage of the HiFi-GAN generator, third generation: PHASE-MAJOR PLANES.
//     x   = ConvTranspose1d(lrelu(y_prev, 0.1))                       (stride u = 4, kernel 8)
//     out = 1/3 * sum_j ResBlock2_j(x)                                (MRF)
//     y   = tanh(conv_post(lrelu(out, 0.01)))  (+ per-utterance max|y| for the int16 scaling)
// (reference graph: the tail of generator.onnx run by mimic3_tts/voice.py:230.)
//
// What dec_fused_kernel (kernels_tc_dec2.cu) paid for and this kernel does not:
//   * the polyphase transposed conv leaves its result as D[t][ph*C + co] (lane = y_prev row t, one column block per output
//     phase ph), while every later conv wants lane = SAMPLE row 4t+ph.  v2 transposed through an fp32 staging tile:
//     two full epilogue passes and two 512-thread barriers per window (~25 % of the epilogue warps' time, all of it on
//     the window's critical path).  Here the WHOLE stage lives in phase-major order: a 512-sample window is four planes
//     of 128 rows, sample 4t+ph = row t of plane ph, in shared memory and in TMEM alike.  D is then already the x tile
//     of plane ph for lane t -- the transposed-conv epilogue is one pass like any other.  A conv tap at sample offset o
//     maps plane ph -> plane (ph+o) mod 4 shifted by floor((ph+o)/4) rows: still one descriptor start address per MMA,
//     same MMA count per row as before;
//   * 512-sample windows (v2: 384): 412 useful rows of 512 (80 %) instead of 288 of 384 (75 %), per-window hand-offs
//     amortised over 1.43x the samples.  Shared memory pays for it by double-buffering the second-conv operand
//     (chains 0 and 2 share YA, chain 1 and conv_post share YB) instead of one buffer per chain;
//   * the next window's transposed conv (into T_0, free once chain 0's epilogue has read it) and its epilogue run
//     BEFORE this window's final epilogue, so the tensor pipe goes straight from c2 of window w to c1 of window w+1
//     while the epilogue warps reduce window w and conv_post(w) slots in between c1_1 and c1_2 of w+1;
//   * two issuer warps (planes 0-1 / planes 2-3): ~270 MMAs per window would make one issuing thread the limiter;
//   * conv_post (ONE output channel) no longer runs as 7 taps x 4 planes of N=16 MMAs: its output COLUMN is the output plane
//     ph', and the input side is grouped by (row shift sh, input plane pi) -- 10 groups x 2 k-steps = 20 MMAs per window
//     instead of 56 (each costs its 4.6 KB operand fetch whatever it computes), and a thread reads the four samples
//     4t..4t+3 of its row with one 4-column TMEM load.  The two issuers accumulate five groups each into their own
//     16-column tile; the epilogue adds the two.
// Window origins are chosen so that (w0 + up_pad) % 4 == 0: plane ph IS polyphase ph, no lane shift anywhere.
// TMEM (512 columns): T_j = [128 j, 128 j + 128) conv1 accumulators of chain j, plane ph at +32 ph (T_0 doubles as the
// transposed-conv result D of the NEXT window), S = [384, 512) all second convs (conv_post reuses [384, 416)).
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

namespace {
constexpr int kC = 32, kCH = kC / 8, kNP = 4, kPR = 128, kW = kNP * kPR;  // channels, planes, rows per plane, window samples
constexpr int kEpiWarps = 16, kIssA = 16, kLoader = 18, kLoaders = 2;  // warps 0-15 epilogue, 16-17 issuers A / B, 18-19 loaders
// (registers are allocated in groups of four warps: 19 warps cost as much as 20, so the limit is 65536 / 640 -> 96 per thread)
constexpr int kThreads = 32 * (kLoader + kLoaders);
constexpr int kLD = 5;  // 32-byte y_prev chunks in flight per loader lane
constexpr uint32_t kT0 = 0, kS0 = 384;
constexpr int kSegTable = 256;
enum Bar { W_FULL = 0, A_FULL, U_DONE, X_READY, C1_DONE, Y_READY = C1_DONE + 3, YA_FREE = Y_READY + 3, C2_DONE, O_READY, P_DONE, S_FREE, NBAR };

__host__ __device__ inline int plane_halo(int h) { return (h + 3) / 4; }  // rows of a plane a +-h sample shift can reach

struct Geo {
  int hx, ha, hb, px, pa, pb, rows_x, rows_ya, rows_yb, rows_a;
  size_t off_w, off_x, off_ya, off_yb, off_a, total;
};
__host__ __device__ inline Geo make_geo(const DecFusedParams& p) {
  Geo g;
  const int h2a = p.HYb[0] > p.HYb[2] ? p.HYb[0] : p.HYb[2];
  const int h2b = p.HYb[1] > 3 ? p.HYb[1] : 3;  // conv_post (k7) reads YB too
  g.hx = plane_halo(p.HX);
  g.ha = plane_halo(h2a);
  g.hb = plane_halo(h2b);
  g.px = kPR + 2 * g.hx;
  g.pa = kPR + 2 * g.ha;
  g.pb = kPR + 2 * g.hb;
  g.rows_x = (kNP * g.px) | 1;
  g.rows_ya = (kNP * g.pa) | 1;
  g.rows_yb = (kNP * g.pb) | 1;
  g.rows_a = kPR + 1;
  size_t o = 0;
  g.off_w = o;
  o += (size_t(p.w_bytes) + 127) & ~size_t(127);
  g.off_x = o;
  o += size_t(kCH) * g.rows_x * 16;
  g.off_ya = o;
  o += size_t(kCH) * g.rows_ya * 16;
  g.off_yb = o;
  o += size_t(kCH) * g.rows_yb * 16;
  g.off_a = o;
  o += size_t(p.cin / 8) * g.rows_a * 16;
  g.total = o;
  return g;
}
__device__ __forceinline__ float lrelu(float v, float s) { return fmaxf(v, s * v); }

// Tuning knobs (compile-time; tools/build_variant.sh builds A/B libraries):
#ifndef DEC3_BATCH       // accumulator planes pulled out of TMEM per tcgen05.wait::ld (1, 2 or 4).  Measured (r02u, dec_last ms):
#define DEC3_BATCH 1      // 1: 2.89, 2: 3.00, 4: 3.99 (spills) -- the 16 warps hide the load latency, registers are the scarcer resource
#endif
#ifndef DEC3_F32X2       // packed fp32 pair arithmetic (FADD2 / FMUL2) in the epilogues.  Measured (r02u): 1: 3.00, 0: 3.08
#define DEC3_F32X2 1
#endif
#ifndef DEC3_H2          // fp16 operands only: leaky-relu and the running sum of x1_j on packed half pairs (HMUL2 / HMNMX2 / HADD2)
#define DEC3_H2 1         // instead of fp32 + converts: ~60 -> ~38 instructions per plane row of a chain epilogue
#endif
#ifndef DEC3_KO           // timing knock-outs (results wrong, attribution only): 1 no MMAs, 2 epilogues without TMEM loads and
#define DEC3_KO 0         // math (zeros stored), 3 epilogues with TMEM loads but no math
#endif
#ifndef DEC3_POLL_ALL    // 1: every lane of a waiting epilogue warp polls its mbarrier; 0: lane 0 polls + __syncwarp.  Measured (r02u):
#define DEC3_POLL_ALL 1   // 1: 2.89, 0: 3.00 -- the single poller wakes its warp later than the hardware wakes 32 suspended lanes
#endif

// fp32 pairs in one 64-bit register: sm_100 executes add / mul on both halves in one instruction (FADD2 / FMUL2);
// the epilogue warps are issue-bound (ncu r02s: ~1400 instructions per warp and window, 97 % busy), so halving the
// fp32 adds of the residual / bias / running-sum path is time, not style.
__device__ __forceinline__ unsigned long long pk64(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void un64(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
#if DEC3_F32X2
  unsigned long long r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
#else
  float a0, a1, b0, b1;
  un64(a, a0, a1);
  un64(b, b0, b1);
  return pk64(a0 + b0, a1 + b1);
#endif
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
#if DEC3_F32X2
  unsigned long long r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
#else
  float a0, a1, b0, b1;
  un64(a, a0, a1);
  un64(b, b0, b1);
  return pk64(a0 * b0, a1 * b1);
#endif
}
// 16-bit operand pair of lrelu(v, slope) for an fp32 pair.  Half-pair variant (fp16 operands): round first, then
// max(h, slope * h) on the pair -- identical for v >= 0; for v < 0 the result carries two more fp16 roundings of a value
// that is 10x (100x) smaller than its positive-side neighbours' rounding errors, i.e. noise below the operand format's own.
template <class E, int FMT>
__device__ __forceinline__ uint32_t lrelu_op(unsigned long long v, unsigned long long slope2, float slope) {
  if constexpr (DEC3_H2 && FMT == 0) {
    float a, b;
    un64(v, a, b);
    return E::lrelu2(E::pack2(a, b), slope);
  } else {
    float a, b, c, d;
    un64(v, a, b);
    un64(mul2(v, slope2), c, d);
    return E::pack2(fmaxf(a, c), fmaxf(b, d));
  }
}
template <class E>
__device__ __forceinline__ uint32_t lrelu_pack(unsigned long long v, unsigned long long slope2) {
  float a, b, c, d;
  un64(v, a, b);
  un64(mul2(v, slope2), c, d);
  return E::pack2(fmaxf(a, c), fmaxf(b, d));
}
}  // namespace

template <int FMT>
__global__ void __maxnreg__(96) dec_planes_kernel(DecFusedParams p) {
  using E = tc::Elem<FMT>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[NBAR];
  __shared__ __align__(16) float sbias[5][kC];  // [0] up bias, [1..3] first-conv bias of chain j, [4] summed second-conv bias
  __shared__ int s_rows[kSegTable];             // per-utterance sample counts (larger batches read global memory)

  const Geo g = make_geo(p);
  uint8_t* const wts = smem + g.off_w;
  uint8_t* const bufX = smem + g.off_x;
  uint8_t* const bufYA = smem + g.off_ya;
  uint8_t* const bufYB = smem + g.off_yb;
  uint8_t* const bufA = smem + g.off_a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int u = kNP, CHI = 8, NUP = u * kC;  // cin = 64

  // ---- work items: (utterance, window), the same deterministic sequence in every role ------------
  auto seg_rows = [&](int seg) { return seg < kSegTable ? s_rows[seg] : p.seg_len[seg] * p.scale; };
  for (int i = tid; i < p.n_seg && i < kSegTable; i += kThreads) s_rows[i] = p.seg_len[i] * p.scale;
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
      uint32_t n = 2;  // commits of the two issuers
      if (i == W_FULL || i == U_DONE) n = 1;
      else if (i == A_FULL) n = kLoaders;
      else if (i == X_READY || (i >= Y_READY && i < Y_READY + 3) || i == O_READY) n = kEpiWarps;
      else if (i == S_FREE) n = 4;
      tc::mbar_init(&bars[i], n);
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
  {  // activations start as zeros: plane halos that no epilogue ever writes stay zero for the whole kernel
    uint4* z = reinterpret_cast<uint4*>(smem + g.off_x);
    const int n16 = int((g.total - g.off_x) / 16);
    for (int i = tid; i < n16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == kIssA) tc::tmem_alloc<512>(&tmem_slot);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
#ifdef M3B200_KERNEL_PROFILE  // per-role cycle counters (M3B200_DEC_PROFILE=1); compiled out of the shipped library
  const bool prof = p.prof != nullptr;
#else
  constexpr bool prof = false;
#endif
  auto warp_wait = [&](uint64_t* bar, uint32_t parity, long long& acc) {  // all lanes of a warp call it; lane 0 polls
    if (prof && !tc::mbar_test(bar, parity)) {
      const long long t = clock64();
      DEC3_POLL_ALL ? tc::mbar_wait(bar, parity) : tc::mbar_wait_warp(bar, parity);
      acc += clock64() - t;
    } else {
      DEC3_POLL_ALL ? tc::mbar_wait(bar, parity) : tc::mbar_wait_warp(bar, parity);
    }
  };
  auto timed_wait = [&](uint64_t* bar, uint32_t parity, long long& acc) {  // single-thread wait (the elected issuers)
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
    // bufA row ra = lrelu(y_prev[tq0 - 1 + ra]) as 16-bit operands, zero outside the utterance
    if (first < total) {
      if (warp == kLoader && tc::elect_one()) {
        tc::mbar_expect_tx(&bars[W_FULL], p.w_bytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wblob);
        const uint32_t main_bytes = p.post.woff;  // blob[0, post.woff): up | c1 / c2 of the three chains
        for (uint32_t o = 0; o < main_bytes; o += 32768u) {
          const uint32_t n = main_bytes - o < 32768u ? main_bytes - o : 32768u;
          tc::bulk_g2s(wts + o, src + o, n, &bars[W_FULL]);
        }
        tc::bulk_g2s(wts + main_bytes, src + p.post_planes_src, p.w_bytes - main_bytes, &bars[W_FULL]);  // regrouped conv_post
      }
      __syncwarp();
      const int items = g.rows_a * CHI;
      int it = 0;
      for (int idx = first; idx < total; idx = next_item(idx), ++it) {
        if (it > 0) tc::mbar_wait_warp(&bars[U_DONE], uint32_t(it - 1) & 1u);  // previous window's transposed conv has read bufA
        const int seg = idx / p.max_win, win = idx - seg * p.max_win;
        const int Lprev = seg_rows(seg) / p.scale * p.prev_scale;
        const long long base_prev = (long long)p.seg_off[seg] * p.prev_scale;
        const int tq0 = (win * p.stride - p.HL + p.up_pad) / u;  // exact: the launcher keeps the numerator a multiple of u
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
  } else if (warp >= kIssA) {
    // =================================== MMA issuer warps ==========================================
    // issuer A: planes 0-1 of every conv + the transposed conv; issuer B: planes 2-3.  Every *_DONE barrier
    // collects one tcgen05.commit from each.
    if (first < total && tc::elect_one()) {
      const bool isA = warp == kIssA;
      const int p0 = isA ? 0 : 2;
      const uint32_t wbase16 = tc::smem_u32(wts) >> 4;
      const uint32_t idC = tc::make_idesc(128, kC, FMT), idU = tc::make_idesc(128, NUP, FMT), idP = tc::make_idesc(128, 16, FMT);
      // One conv over my two planes: per tap 2 k-steps x 2 planes; a tap at sample offset o reads plane (ph + o) mod 4
      // shifted by floor((ph + o) / 4) rows -- one descriptor start address per MMA.
      auto conv = [&](const uint8_t* abuf, int rows_in, int pitch, int halo, const DecFusedConv& cv, int N, uint32_t dcol,
                      uint32_t idesc, bool acc0) {
        const uint64_t a_tmpl = tc::make_desc(0u, uint32_t(rows_in) * 16u, 128u);
        const uint64_t b_tmpl = tc::make_desc(0u, uint32_t(N) * 16u, 128u);
        const uint32_t abase = (tc::smem_u32(abuf) >> 4) + uint32_t(halo);
        uint32_t bt = wbase16 + (cv.woff >> 4);
        int o = -cv.pad_left * cv.dil + 256;  // + 256: keeps ph + o positive, 256 % 4 == 0
#pragma unroll 1
        for (int t = 0; t < cv.taps; ++t, o += cv.dil, bt += uint32_t(4 * N)) {
          uint32_t a0[2];
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const int q = p0 + pp + o;
            a0[pp] = abase + uint32_t((q & 3) * pitch + (q >> 2) - 64);
          }
          const uint32_t acc = (acc0 || t > 0) ? 1u : 0u;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t bd = b_tmpl | uint64_t(bt + uint32_t(ks * 2 * N));
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              const uint64_t ad = a_tmpl | uint64_t(a0[pp] + uint32_t(ks * 2 * rows_in));
              if (DEC3_KO != 1) tc::mma_f16_ss(tmem + dcol + uint32_t((p0 + pp) * N), ad, bd, idesc, ks > 0 ? 1u : acc);
            }
          }
        }
      };
      auto issue_up = [&] {  // D[t][ph*C + co] = sum_{d=0,1} A[t + 1 - d] . W_d : K = 64, N = 128, into T_0
        const uint64_t a_tmpl = tc::make_desc(0u, uint32_t(g.rows_a) * 16u, 128u);
        const uint64_t b_tmpl = tc::make_desc(0u, uint32_t(NUP) * 16u, 128u);
        const uint32_t abase = tc::smem_u32(bufA) >> 4;
        const uint32_t bbase = wbase16 + (p.up.woff >> 4);
#pragma unroll
        for (int d = 0; d < 2; ++d)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = a_tmpl | uint64_t(abase + uint32_t(1 - d + ks * 2 * g.rows_a));
            const uint64_t bd = b_tmpl | uint64_t(bbase + uint32_t((d * CHI + ks * 2) * NUP));
            if (DEC3_KO != 1) tc::mma_f16_ss(tmem + kT0, ad, bd, idU, (d || ks) ? 1u : 0u);
          }
      };
      // conv_post: five (row shift, input plane) groups per issuer into its own 16-column tile of S
      auto issue_post = [&] {
        const uint64_t a_tmpl = tc::make_desc(0u, uint32_t(g.rows_yb) * 16u, 128u);
        const uint64_t b_tmpl = tc::make_desc(0u, 16u * 16u, 128u);
        const uint32_t abase = (tc::smem_u32(bufYB) >> 4) + uint32_t(g.hb);
        const uint32_t bbase = wbase16 + (p.post.woff >> 4);
        const uint32_t dcol = tmem + kS0 + (isA ? 0u : 16u);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int pair = (isA ? 0 : 5) + k;          // order: sh = -1: pi 1..3 | sh = 0: pi 0..3 | sh = +1: pi 0..2
          const int sh = pair < 3 ? -1 : (pair < 7 ? 0 : 1);
          const int pi = pair < 3 ? pair + 1 : (pair < 7 ? pair - 3 : pair - 7);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ad = a_tmpl | uint64_t(abase + uint32_t(pi * g.pb + sh + ks * 2 * g.rows_yb));
            const uint64_t bd = b_tmpl | uint64_t(bbase + uint32_t(pair * 64 + ks * 32));
            if (DEC3_KO != 1) tc::mma_f16_ss(dcol, ad, bd, idP, (k || ks) ? 1u : 0u);
          }
        }
      };
      long long c_x = 0, c_y = 0, c_o = 0, c_s = 0;
      tc::mbar_wait(&bars[W_FULL], 0u);
      if (isA) {
        tc::mbar_wait(&bars[A_FULL], 0u);
        tc::fence_after_sync();
        issue_up();
        tc::mma_commit(&bars[U_DONE]);
      }
      const long long c_start = prof ? clock64() : 0;
      int it = 0;
      for (int idx = first; idx < total; ++it) {
        const int nxt = next_item(idx);
        const uint32_t par = uint32_t(it) & 1u, ppar = par ^ 1u;
        timed_wait(&bars[X_READY], par, c_x);
        tc::fence_after_sync();
        conv(bufX, g.rows_x, g.px, g.hx, p.c1[0], kC, kT0, idC, false);
        tc::mma_commit(&bars[C1_DONE + 0]);
        conv(bufX, g.rows_x, g.px, g.hx, p.c1[1], kC, kT0 + 128u, idC, false);
        tc::mma_commit(&bars[C1_DONE + 1]);
        if (it > 0) {  // previous window's conv_post: its operand was published while c1_0 / c1_1 were being issued
          timed_wait(&bars[O_READY], ppar, c_o);
          tc::fence_after_sync();
          issue_post();
          tc::mma_commit(&bars[P_DONE]);
        }
        conv(bufX, g.rows_x, g.px, g.hx, p.c1[2], kC, kT0 + 256u, idC, false);
        tc::mma_commit(&bars[C1_DONE + 2]);
        timed_wait(&bars[Y_READY + 0], par, c_y);
        if (it > 0) timed_wait(&bars[S_FREE], ppar, c_s);  // conv_post's result has been read out of S
        tc::fence_after_sync();
        conv(bufYA, g.rows_ya, g.pa, g.ha, p.c2[0], kC, kS0, idC, false);
        tc::mma_commit(&bars[YA_FREE]);
        if (isA && nxt < total) {  // next window's transposed conv: T_0 is free once chain 0's epilogue has read it
          tc::mbar_wait(&bars[A_FULL], uint32_t(it + 1) & 1u);
          tc::fence_after_sync();
          issue_up();
          tc::mma_commit(&bars[U_DONE]);
        }
        timed_wait(&bars[Y_READY + 1], par, c_y);
        tc::fence_after_sync();
        conv(bufYB, g.rows_yb, g.pb, g.hb, p.c2[1], kC, kS0, idC, true);
        timed_wait(&bars[Y_READY + 2], par, c_y);
        tc::fence_after_sync();
        conv(bufYA, g.rows_ya, g.pa, g.ha, p.c2[2], kC, kS0, idC, true);
        tc::mma_commit(&bars[C2_DONE]);
        idx = nxt;
      }
      timed_wait(&bars[O_READY], uint32_t(it - 1) & 1u, c_o);
      tc::fence_after_sync();
      issue_post();
      tc::mma_commit(&bars[P_DONE]);
      if (prof && isA) {
        unsigned long long* q = reinterpret_cast<unsigned long long*>(p.prof);
        atomicAdd(q + 0, (unsigned long long)(clock64() - c_start));
        atomicAdd(q + 1, (unsigned long long)c_x);
        atomicAdd(q + 2, (unsigned long long)c_y);
        atomicAdd(q + 3, (unsigned long long)c_s);
        atomicAdd(q + 4, (unsigned long long)c_o);
        atomicAdd(q + 5, (unsigned long long)it);
      }
    }
    __syncwarp();
  } else {
    // =================================== epilogue warps ============================================
    // thread = (plane row t = 32 q + lane, 8 channels cg*8..); it owns that row of all four planes
    const int q = warp & 3, cg = warp >> 2, col0 = cg * 8;
    const int tp = q * 32 + lane;
    const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16) + uint32_t(col0);
    // x (fp32) and sum_j x1_j.  The running sum is kept as packed fp16 pairs (added in fp32, rounded once per chain): its
    // only consumer rounds (sum + S) / 3 to a 16-bit conv_post operand anyway, and 64 fp32 registers of per-thread state
    // do not fit next to the epilogue temporaries at 96 registers per thread.
    unsigned long long xr[kNP][4];  // fp32 pairs
    uint32_t xs[kNP][4];            // fp16 pairs
    using u64 = unsigned long long;
    const u64 slope01 = pk64(0.1f, 0.1f);

    auto store_op = [&](uint8_t* buf, int rows_total, int row, const uint32_t* pk) {
      *reinterpret_cast<uint4*>(buf + (size_t(cg) * rows_total + row) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    };
    auto load_bias = [&](int j, u64* b) {  // 8 channels = 4 pairs
      const ulonglong2 f0 = *reinterpret_cast<const ulonglong2*>(&sbias[j][col0]);
      const ulonglong2 f1 = *reinterpret_cast<const ulonglong2*>(&sbias[j][col0 + 4]);
      b[0] = f0.x, b[1] = f0.y, b[2] = f1.x, b[3] = f1.y;
    };
    // one pass over the thread's four plane rows: DEC3_BATCH accumulator tiles per tcgen05.wait::ld, body(ph, pairs)
    auto for_planes = [&](uint32_t tcol, auto&& body) {
#pragma unroll
      for (int pq = 0; pq < kNP; pq += DEC3_BATCH) {
        float v[DEC3_BATCH][8];
        if (DEC3_KO != 2) {
#pragma unroll
          for (int h = 0; h < DEC3_BATCH; ++h) tc::tmem_ld8(lane_base + tcol + uint32_t((pq + h) * kC), v[h]);
          tc::tmem_ld_wait();
        }
#pragma unroll
        for (int h = 0; h < DEC3_BATCH; ++h) {
          if (DEC3_KO >= 2) {  // knock-out: publish something, skip the arithmetic
            const uint32_t z[4] = {DEC3_KO == 3 ? __float_as_uint(v[h][0]) : 0u, 0u, 0u, 0u};
            if (tcol == kT0) store_op(bufX, g.rows_x, (pq + h) * g.px + g.hx + tp, z);
            else store_op(bufYB, g.rows_yb, (pq + h) * g.pb + g.hb + tp, z);
            continue;
          }
          u64 a[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) a[c] = pk64(v[h][2 * c], v[h][2 * c + 1]);
          body(pq + h, a);
        }
      }
    };
    auto arrive = [&](int b) {
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[b]);
    };
    long long e_u = 0, e_p = 0, e_c1 = 0, e_c2 = 0, e_f = 0;
    const long long e_start = prof ? clock64() : 0;

    // transposed-conv epilogue of a window: x = D + b stays in registers, lrelu(x) becomes the first convs' operand
    auto up_epi = [&](int w0, int L, uint32_t par) {
      warp_wait(&bars[U_DONE], par, e_u);
      tc::fence_after_sync();
      u64 ub[4];
      load_bias(0, ub);
      for_planes(kT0, [&](int ph, const u64* a) {
        const int gi = w0 + u * tp + ph;
        const bool inside = gi >= 0 && gi < L;
        uint32_t pk[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          xr[ph][c] = add2(a[c], ub[c]);
          pk[c] = inside ? lrelu_op<E, FMT>(xr[ph][c], slope01, 0.1f) : 0u;
        }
        store_op(bufX, g.rows_x, ph * g.px + g.hx + tp, pk);
      });
      arrive(X_READY);
    };
    // conv_post result of a window: the row's four samples, tanh, store, per-utterance peak
    auto post_epi = [&](int seg, int w0, int L, long long base, uint32_t par) {
      warp_wait(&bars[P_DONE], par, e_p);  // every warp: YB may be rewritten once conv_post has read it
      if (cg == 0) {
        tc::fence_after_sync();
        float v[kNP], v2[kNP];
        tc::tmem_ld4(lane_base + kS0, v);        // issuer A's groups: column ph' = sample 4t + ph'
        tc::tmem_ld4(lane_base + kS0 + 16u, v2);  // issuer B's groups
        tc::tmem_ld_wait();
#pragma unroll
        for (int ph = 0; ph < kNP; ++ph) v[ph] += v2[ph];
        float mx = 0.f;
#pragma unroll
        for (int ph = 0; ph < kNP; ++ph) {
          const int r = u * tp + ph;
          const int gi = w0 + r;
          if (r >= p.HL && r < p.HL + p.stride && gi < L) {
            const float y = tanhf(v[ph]);
            p.audio[base + gi] = y;
            mx = fmaxf(mx, fabsf(y));
          }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0 && mx > 0.f) atomicMax(p.peak_bits + seg, __float_as_uint(mx));
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[S_FREE]);
      }
    };

    int it = 0;
    int pseg = 0, pw0 = 0, pL = 0;
    long long pbase = 0;
    if (first < total) {
      const int seg = first / p.max_win, win = first - seg * p.max_win;
      up_epi(win * p.stride - p.HL, seg_rows(seg), 0u);
    }
    for (int idx = first; idx < total; ++it) {
      const int nxt = next_item(idx);
      const uint32_t par = uint32_t(it) & 1u;
      const int seg = idx / p.max_win, win = idx - seg * p.max_win;
      const int L = seg_rows(seg);
      const long long base = (long long)p.seg_off[seg] * p.scale;
      const int w0 = win * p.stride - p.HL;

      // ---- first conv of each chain: x1 = x + b + conv(lrelu x); operand of the second conv = lrelu(x1) ----
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        warp_wait(&bars[C1_DONE + j], par, e_c1);
        if (j == 1 && it > 0) post_epi(pseg, pw0, pL, pbase, par ^ 1u);  // previous window's audio; frees S and YB
        if (j == 2) warp_wait(&bars[YA_FREE], par, e_c2);                // chain 0's second conv has read YA
        tc::fence_after_sync();
        uint8_t* const by = j == 1 ? bufYB : bufYA;
        const int rows_t = j == 1 ? g.rows_yb : g.rows_ya, pitch = j == 1 ? g.pb : g.pa, hy = j == 1 ? g.hb : g.ha;
        u64 bj[4];
        load_bias(1 + j, bj);
        for_planes(kT0 + uint32_t(j * 128), [&](int ph, const u64* a) {
          const int gi = w0 + u * tp + ph;
          const bool inside = gi >= 0 && gi < L;
          uint32_t pk[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const u64 v = add2(add2(a[c], xr[ph][c]), bj[c]);  // x1 = conv + x + b
            float v0, v1;
            un64(v, v0, v1);
            if constexpr (DEC3_H2 && FMT == 0) {
              const uint32_t pv = E::pack2(v0, v1);
              if (j > 0) {
                const __half2 h = __hadd2(*reinterpret_cast<const __half2*>(&xs[ph][c]), *reinterpret_cast<const __half2*>(&pv));
                xs[ph][c] = *reinterpret_cast<const uint32_t*>(&h);
              } else {
                xs[ph][c] = pv;
              }
              pk[c] = inside ? E::lrelu2(pv, 0.1f) : 0u;
            } else {
              float s0 = 0.f, s1 = 0.f;
              if (j > 0) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xs[ph][c]));
                s0 = f.x, s1 = f.y;
              }
              const __half2 h = __floats2half2_rn(s0 + v0, s1 + v1);
              xs[ph][c] = *reinterpret_cast<const uint32_t*>(&h);
              pk[c] = inside ? lrelu_pack<E>(v, slope01) : 0u;
            }
          }
          store_op(by, rows_t, ph * pitch + hy + tp, pk);
        });
        arrive(Y_READY + j);
      }

      // ---- next window's x (its transposed conv ran in T_0 under this window's second convs) ----
      if (nxt < total) {
        const int nseg = nxt / p.max_win, nwin = nxt - nseg * p.max_win;
        up_epi(nwin * p.stride - p.HL, seg_rows(nseg), par ^ 1u);
      }

      // ---- out = (sum_j x1_j + S + late bias) / 3; operand of conv_post = lrelu(out, 0.01) in YB ----
      warp_wait(&bars[C2_DONE], par, e_f);
      tc::fence_after_sync();
      {
        u64 bl[4];
        load_bias(4, bl);
        const u64 inv2 = pk64(p.inv_nk, p.inv_nk), slope001 = pk64(0.01f, 0.01f);
        for_planes(kS0, [&](int ph, const u64* a) {
          const int gi = w0 + u * tp + ph;
          const bool inside = gi >= 0 && gi < L;
          uint32_t pk[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xs[ph][c]));
            const u64 o = mul2(add2(add2(a[c], pk64(f.x, f.y)), bl[c]), inv2);
            pk[c] = inside ? lrelu_op<E, FMT>(o, slope001, 0.01f) : 0u;
          }
          store_op(bufYB, g.rows_yb, ph * g.pb + g.hb + tp, pk);
        });
      }
      arrive(O_READY);
      pseg = seg;
      pw0 = w0;
      pL = L;
      pbase = base;
      idx = nxt;
    }
    if (it > 0) post_epi(pseg, pw0, pL, pbase, uint32_t(it - 1) & 1u);
    if (prof && tid == 0) {
      unsigned long long* q8 = reinterpret_cast<unsigned long long*>(p.prof);
      atomicAdd(q8 + 8, (unsigned long long)(clock64() - e_start));
      atomicAdd(q8 + 9, (unsigned long long)e_u);
      atomicAdd(q8 + 10, (unsigned long long)e_p);
      atomicAdd(q8 + 11, (unsigned long long)e_c1);
      atomicAdd(q8 + 12, (unsigned long long)e_c2);
      atomicAdd(q8 + 13, (unsigned long long)e_f);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kIssA) tc::tmem_dealloc<512>(tmem);
}

size_t dec_planes_smem_bytes(const DecFusedParams& p) { return make_geo(p).total + 128; }

bool dec_planes_supported(int C, int cin, int up_k, int up_u, int nk, int nd, int HX, const int* HYb, size_t w_bytes) {
  if (C != kC || nk != 3 || nd != 2 || cin != 64 || up_u != kNP || up_k != 2 * up_u) return false;
  DecFusedParams p;
  p.cin = cin;
  p.up_u = up_u;
  p.HX = HX;
  int hymax = 3;
  for (int j = 0; j < 3; ++j) {
    p.HYb[j] = HYb[j];
    hymax = HYb[j] > hymax ? HYb[j] : hymax;
  }
  p.w_bytes = unsigned(w_bytes);
  if (kW - 2 * (HX + hymax + 3 + kNP) < 128) return false;
  const size_t stat = sizeof(int) * kSegTable + 5 * kC * 4 + NBAR * 8 + 64;
  return make_geo(p).total + 128 + stat <= size_t(227) * 1024;
}

void launch_dec_planes(const DecFusedParams& p_in, int fmt, int n_seg, int max_len, cudaStream_t st) {
  DecFusedParams p = p_in;
  // window origin w0 = win * stride - HL with (w0 + up_pad) % u == 0 and stride % u == 0: plane ph is polyphase ph
  const int u = kNP;
  p.HL = p.H + (((p.up_pad - p.H) % u) + u) % u;
  int hr = p.H;
  while ((kW - p.HL - hr) % u) ++hr;
  p.stride = kW - p.HL - hr;
  if (p.stride <= 0) throw std::runtime_error("dec_planes: receptive field exceeds the window");
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
  const size_t smem = dec_planes_smem_bytes(p);
  const void* kern = fmt ? reinterpret_cast<const void*>(dec_planes_kernel<1>) : reinterpret_cast<const void*>(dec_planes_kernel<0>);
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
  if (fmt) dec_planes_kernel<1><<<grid, kThreads, smem, st>>>(p);
  else dec_planes_kernel<0><<<grid, kThreads, smem, st>>>(p);
  post_launch("dec_planes_kernel", st);
  if (want_prof) {  // debug only: synchronous read-back of the per-role cycle counters (summed over CTAs)
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost);
    const double n = double(h[5] > 0 ? h[5] : 1);
    fprintf(stderr,
            "[dec_planes profile] windows %lld grid %d stride %d | issuer A cycles/window: total %.0f wait_x %.0f wait_y %.0f wait_s %.0f "
            "wait_o %.0f | epilogue warp 0: total %.0f wait_u %.0f wait_p %.0f wait_c1 %.0f wait_ya %.0f wait_c2 %.0f\n",
            h[5], grid, p.stride, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n,
            h[12] / n, h[13] / n);
  }
}

}  // namespace m3

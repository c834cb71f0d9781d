// Fused LAST stage of the HiFi-GAN generator, one kernel:
//     x   = ConvTranspose1d(lrelu(y_prev, 0.1))                       (stride u, kernel k)
//     out = 1/nk * sum_j ResBlock2_j(x)                               (MRF)
//     y   = tanh(conv_post(lrelu(out, 0.01)))  (+ per-utterance max|y| for the int16 scaling)
// Neither x nor the C-channel `out` ever touches HBM: the kernel reads y_prev (2C channels at 1/u
// of the rate) and writes one float per sample.
//
// Everything is an implicit-GEMM conv on the same 128-row TMEM tiles (lane = output sample):
//   * the transposed conv is a plain conv over the ZERO-STUFFED input (rows s = t*u hold lrelu(y[t]),
//     the rest are zero): taps are descriptor shifts like everywhere else and the result lands in
//     the residual tile T in output-row order, ready to be the resblocks' fp32 residual stream
//     (3/4 of those MMA rows multiply zeros -- the price for not transposing through memory);
//   * resblock convs accumulate on top of T (residual add for free), running sum S in TMEM,
//     exactly as in mrf_tc_kernel (kernels_tc.cu);
//   * conv_post (C -> 1, k = 7) is one more tensor-core conv with N = 16 (column 0 is real).
// x is only computed on the window's own rows, so every conv shrinks the exact region:
// H = HX + HY + 3 rows per side are recomputed by the neighbouring windows.
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

namespace {
__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
}  // namespace

template <int C, int NT, int FMT, int NW>
__global__ void __launch_bounds__(NW * 32, 2) dec_last_kernel(DecStageParams p) {
  constexpr int NTHR = NW * 32;
  constexpr int R = NT * 128;
  constexpr int CH = C / 8;
  constexpr int HC = C / (NW / 4);   // columns per epilogue thread (NW warps = 4 lane quarters x NW/4 column groups)
  constexpr int G = HC >= 16 ? 16 : 8;  // columns per tcgen05.ld / st
  constexpr int NCC = HC / G;
  static_assert(HC % 8 == 0 && HC >= 8, "column split");
  constexpr int TCOLS_RAW = 2 * NT * C;
  constexpr int TCOLS = TCOLS_RAW <= 64 ? 64 : TCOLS_RAW <= 128 ? 128 : TCOLS_RAW <= 256 ? 256 : 512;
  static_assert(TCOLS_RAW <= 512, "TMEM budget");
  using E = tc::Elem<FMT>;

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t gbar[2], tbar[NT];
  __shared__ float sbias[6][C];  // [0] up bias, [1..4] first-conv bias of resblock j, [5] summed late bias

  const int seg = blockIdx.y;
  const int L = p.seg_len[seg] * p.scale;          // samples of this utterance at this level
  const int Lprev = p.seg_len[seg] * p.prev_scale; // rows of y_prev
  const int o0 = blockIdx.x * p.stride;
  if (o0 >= L) return;
  const long long base_prev = (long long)p.seg_off[seg] * p.prev_scale;
  const long long base = (long long)p.seg_off[seg] * p.scale;
  const int w0 = o0 - p.H;
  const int CHI = p.cin / 8;
  const int ROWSX = (R + 2 * p.HX) | 1, ROWSY = (R + 2 * p.HY) | 1, ROWSU = (R + p.up.taps - 1) | 1;
  uint8_t* bufX = smem;
  uint8_t* bufU = bufX + size_t(CH) * ROWSX * 16;   // zero-stuffed lrelu(y_prev); later reused as bufY
  uint8_t* bufY = bufU;
  const size_t u_bytes = size_t(CHI) * ROWSU * 16, y_bytes = size_t(CH) * ROWSY * 16;
  uint8_t* wbuf = bufU + (((u_bytes > y_bytes ? u_bytes : y_bytes) + 15) & ~size_t(15));
  const uint32_t wb_bytes = uint32_t(p.wb_bytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hhalf = warp >> 2;

  // ---- weight tap-groups of the whole kernel, in order: step s, first tap g0 ------------------
  auto step_wg = [&](int s) { return max(1, int(wb_bytes) / (p.steps[s].K * p.steps[s].N * 2)); };
  auto prefetch = [&](int s, int g0, int buf) {
    const DecConv& cv = p.steps[s];
    const int ntap = min(step_wg(s), cv.taps - g0);
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(p.w16 + cv.woff + size_t(g0) * cv.K * cv.N);
    uint4* dst = reinterpret_cast<uint4*>(wbuf + size_t(buf) * wb_bytes);
    const int n16 = ntap * cv.K * cv.N / 8;
    for (int i = tid; i < n16; i += NTHR) cpa16(dst + i, src + i);
    cpa_commit();
  };
  prefetch(0, 0, 0);

  for (int i = tid; i < 6 * C; i += NTHR) {
    const int j = i / C, c = i - j * C;
    float v = 0.f;
    if (j == 0) v = p.up_bias[c];
    else if (j <= 4) v = (j - 1 < p.nk) ? p.bias0[j - 1][c] : 0.f;
    else v = p.late_bias[c];
    sbias[j][c] = v;
  }
  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_slot);
  if (tid == 0) {
    tc::mbar_init(&gbar[0], 1);
    tc::mbar_init(&gbar[1], 1);
    for (int m = 0; m < NT; ++m) tc::mbar_init(&tbar[m], 1);
    tc::mbar_fence_init();
  }
  // ---- zero-stuffed lrelu(y_prev): bufU row sr <-> stuffed index s = w0 - pl + sr ---------------
  {
    const int s0 = w0 - p.up.pad_left;
    const int items = CHI * ROWSU;
    auto lr = [](float v) { return v >= 0.f ? v : 0.1f * v; };
    for (int idx = tid; idx < items; idx += NTHR) {
      const int sr = idx / CHI, c8 = idx - sr * CHI;
      const int s = s0 + sr;
      uint4 pk = make_uint4(0u, 0u, 0u, 0u);
      if (s >= 0 && (s % p.up_u) == 0) {
        const int t = s / p.up_u;
        if (t < Lprev) {
          const float* src = p.yprev + (base_prev + t) * (long long)p.cin + c8 * 8;
          const float4 a = *reinterpret_cast<const float4*>(src);
          const float4 b = *reinterpret_cast<const float4*>(src + 4);
          pk.x = E::pack2(lr(a.x), lr(a.y));
          pk.y = E::pack2(lr(a.z), lr(a.w));
          pk.z = E::pack2(lr(b.x), lr(b.y));
          pk.w = E::pack2(lr(b.z), lr(b.w));
        }
      }
      *reinterpret_cast<uint4*>(bufU + (size_t(c8) * ROWSU + sr) * 16) = pk;
    }
    // bufX rows outside the window's own [0, R) are never produced: keep them zero
    for (int idx = tid; idx < CH * ROWSX; idx += NTHR) {
      const int rr = idx % ROWSX;
      if (rr < p.HX || rr >= R + p.HX) *reinterpret_cast<uint4*>(bufX + size_t(idx) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = tmem + (uint32_t(q * 32) << 16);
  const uint32_t T0 = 0, S0 = NT * C;
  uint32_t gphase[2] = {0u, 0u}, tphase = 0;
  int gi = 0;
  bool prev_nonlast = false;
  float xr[NT][NCC][G];  // x of this thread's rows / columns, fp32 (residual stream source)

  for (int s = 0; s < p.nsteps; ++s) {
    const DecConv cv = p.steps[s];
    const int kind = cv.kind;  // 0 up, 1 resblock conv (not last of its block), 2 last conv of a resblock, 3 post
    const uint8_t* inbuf = kind == 0 ? bufU : (kind == 2 ? bufY : bufX);
    const int rows_in = kind == 0 ? ROWSU : (kind == 2 ? ROWSY : ROWSX);
    const int halo_in = kind == 0 ? cv.pad_left : (kind == 2 ? p.HY : p.HX);  // row of output 0, tap pad_left
    const uint32_t idesc = tc::make_idesc(128, cv.N, FMT);
    const int wg = step_wg(s);
    const int KC = cv.K / 8;
    for (int g0 = 0; g0 < cv.taps; g0 += wg, ++gi) {
      const int ntap = min(wg, cv.taps - g0);
      const bool last_group = g0 + wg >= cv.taps;
      cpa_wait_all();
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      if (warp == 0 && tc::elect_one()) {
        const uint32_t abase = tc::smem_u32(inbuf), wbase = tc::smem_u32(wbuf + size_t(gi & 1) * wb_bytes);
#pragma unroll 1
        for (int t = 0; t < ntap; ++t) {
          const int shift = halo_in + (g0 + t - cv.pad_left) * cv.dil;
#pragma unroll 1
          for (int ks = 0; ks < cv.K / 16; ++ks) {
            const uint64_t bd = tc::make_desc(wbase + uint32_t((t * KC + ks * 2) * cv.N) * 16u, uint32_t(cv.N) * 16u, 128u);
            const uint32_t acc = (kind == 0 || kind == 3) ? ((g0 + t) | ks ? 1u : 0u) : 1u;
#pragma unroll
            for (int m = 0; m < NT; ++m) {
              const uint64_t ad = tc::make_desc(abase + uint32_t((ks * 2) * rows_in + m * 128 + shift) * 16u,
                                                uint32_t(rows_in) * 16u, 128u);
              tc::mma_f16_ss(tmem + T0 + m * C, ad, bd, idesc, acc);
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
      if (prev_nonlast) {
        const int pb = (gi - 1) & 1;
        tc::mbar_wait(&gbar[pb], gphase[pb]);
        gphase[pb] ^= 1u;
      }
      prev_nonlast = !last_group;
      {  // prefetch the next tap-group of the whole kernel
        int ns = s, ng = g0 + wg;
        if (ng >= cv.taps) {
          ng = 0;
          ++ns;
        }
        if (ns < p.nsteps) prefetch(ns, ng, (gi + 1) & 1);
      }
    }
    // ------------------------------ epilogues ------------------------------
    // thread <-> (row = lane quarter q*32+lane of tile m, column group hhalf: HC columns in groups of G)
    auto store_ops = [&](uint8_t* buf, int pitch, int row, int col, const uint32_t* pk) {
      uint8_t* dst = buf + (size_t(col / 8) * pitch + row) * 16;
#pragma unroll
      for (int i = 0; i < G / 8; ++i)
        *reinterpret_cast<uint4*>(dst + size_t(i) * pitch * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
    };
#pragma unroll
    for (int m = 0; m < NT; ++m) {
      tc::mbar_wait(&tbar[m], tphase);
      tc::fence_after_sync();
      const int r = m * 128 + q * 32 + lane;
      const int g = w0 + r;
      const bool inside = g >= 0 && g < L;
      if (kind == 0) {
        // x = T + b_up: keep in registers, publish lrelu(x) as the resblocks' A operand, seed T for resblock 0
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) {
          const int col = hhalf * HC + cc * G;
          float v[G];
          tc::tmem_ldg<G>(lane_base + T0 + m * C + col, v);
          tc::tmem_ld_wait();
          uint32_t pk[G / 2];
#pragma unroll
          for (int e = 0; e < G; ++e) xr[m][cc][e] = v[e] + sbias[0][col + e];
#pragma unroll
          for (int e = 0; e < G / 2; ++e) {
            float a = xr[m][cc][2 * e], b = xr[m][cc][2 * e + 1];
            a = a >= 0.f ? a : 0.1f * a;
            b = b >= 0.f ? b : 0.1f * b;
            pk[e] = inside ? E::pack2(a, b) : 0u;
          }
          store_ops(bufX, ROWSX, r + p.HX, col, pk);
#pragma unroll
          for (int e = 0; e < G; ++e) v[e] = xr[m][cc][e] + sbias[1][col + e];
          tc::tmem_stg<G>(lane_base + T0 + m * C + col, v);
        }
      } else if (kind == 1) {
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) {
          const int col = hhalf * HC + cc * G;
          float v[G];
          tc::tmem_ldg<G>(lane_base + T0 + m * C + col, v);
          tc::tmem_ld_wait();
          uint32_t pk[G / 2];
#pragma unroll
          for (int e = 0; e < G / 2; ++e) {
            float a = v[2 * e], b = v[2 * e + 1];
            a = a >= 0.f ? a : 0.1f * a;
            b = b >= 0.f ? b : 0.1f * b;
            pk[e] = inside ? E::pack2(a, b) : 0u;
          }
          store_ops(bufY, ROWSY, r + p.HY, col, pk);
        }
      } else if (kind == 2) {
        const int j = cv.rb;
        const bool first = j == 0, last = j == p.nk - 1;
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) {
          const int col = hhalf * HC + cc * G;
          float v[G];
          tc::tmem_ldg<G>(lane_base + T0 + m * C + col, v);
          if (!first) {
            float sv[G];
            tc::tmem_ldg<G>(lane_base + S0 + m * C + col, sv);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < G; ++e) v[e] += sv[e];
          } else {
            tc::tmem_ld_wait();
          }
          if (!last) {
            tc::tmem_stg<G>(lane_base + S0 + m * C + col, v);
            float t2[G];  // seed T for the next resblock: x + its first-conv bias
#pragma unroll
            for (int e = 0; e < G; ++e) t2[e] = xr[m][cc][e] + sbias[j + 2][col + e];
            tc::tmem_stg<G>(lane_base + T0 + m * C + col, t2);
          } else {
            // out = (sum + late bias) / nk ; A operand of conv_post = lrelu(out, 0.01)
            uint32_t pk[G / 2];
#pragma unroll
            for (int e = 0; e < G / 2; ++e) {
              float a = (v[2 * e] + sbias[5][col + 2 * e]) * p.inv_nk;
              float b = (v[2 * e + 1] + sbias[5][col + 2 * e + 1]) * p.inv_nk;
              a = a >= 0.f ? a : 0.01f * a;
              b = b >= 0.f ? b : 0.01f * b;
              pk[e] = inside ? E::pack2(a, b) : 0u;
            }
            store_ops(bufX, ROWSX, r + p.HX, col, pk);
          }
        }
      } else {  // kind == 3: y = tanh(column 0), per-utterance peak
        float v[8];
        tc::tmem_ld8(lane_base + T0 + m * C, v);
        tc::tmem_ld_wait();
        float y = 0.f;
        const bool store = hhalf == 0 && r >= p.H && r < R - p.H && g < L;
        if (store) {
          y = tanhf(v[0]);
          p.audio[base + g] = y;
        }
        float mx = fabsf(y);
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0 && hhalf == 0 && mx > 0.f) atomicMax(p.peak_bits + seg, __float_as_uint(mx));
      }
    }
    if (kind == 0 || kind == 2) tc::tmem_st_wait();
    tphase ^= 1u;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

size_t dec_last_smem_bytes(int C, int NT, int cin, int up_taps, int HX, int HY, int wb_bytes) {
  const int R = NT * 128;
  const size_t x = size_t(C / 8) * ((R + 2 * HX) | 1) * 16;
  const size_t u = size_t(cin / 8) * ((R + up_taps - 1) | 1) * 16;
  const size_t y = size_t(C / 8) * ((R + 2 * HY) | 1) * 16;
  return x + ((std::max(u, y) + 15) & ~size_t(15)) + 2 * size_t(wb_bytes) + 64;
}

template <int C, int NT, int FMT, int NW>
static void launch_dec_inst(const DecStageParams& p, int n_seg, int max_len, cudaStream_t st) {
  DecStageParams q = p;
  const int R = NT * 128;
  q.stride = R - 2 * p.H;
  if (q.stride <= 0) throw std::runtime_error("dec_last: receptive field exceeds the window");
  const size_t smem = dec_last_smem_bytes(C, NT, p.cin, p.up.taps, p.HX, p.HY, p.wb_bytes);
  auto kern = dec_last_kernel<C, NT, FMT, NW>;
  ensure_max_dynamic_smem(reinterpret_cast<const void*>(kern));
  const int L = max_len * p.scale;
  dim3 grid((L + q.stride - 1) / q.stride, n_seg);
  kern<<<grid, NW * 32, smem, st>>>(q);
  post_launch("dec_last_kernel", st);
}

bool dec_last_supported(int C, int cin, int up_k, int up_u, int nk, int nd, int HX, int HY) {
  if (C != 32) return false;  // instantiated shape (the *_low voices); others use the unfused kernels
  if (cin % 16 || nk < 1 || nk > 4 || nd != 2 || up_k > 16 || up_u < 1) return false;
  const int H = HX + HY + 3;
  if (3 * 128 - 2 * H < 64) return false;
  return dec_last_smem_bytes(C, 3, cin, up_k, HX, HY, 16 * 1024) <= size_t(113 * 1024);
}

void launch_dec_last(const DecStageParams& p, int C, int fmt, int n_seg, int max_len, cudaStream_t st) {
  static const int nw = [] { const char* e = getenv("M3B200_DEC_WARPS"); return e ? atoi(e) : 8; }();  // 8 measured faster than 16 (profiles/r01_notes.md)
  if (C == 32) {
    if (nw == 8) {
      if (fmt) launch_dec_inst<32, 3, 1, 8>(p, n_seg, max_len, st);
      else launch_dec_inst<32, 3, 0, 8>(p, n_seg, max_len, st);
    } else {
      if (fmt) launch_dec_inst<32, 3, 1, 16>(p, n_seg, max_len, st);
      else launch_dec_inst<32, 3, 0, 16>(p, n_seg, max_len, st);
    }
  } else {
    throw std::runtime_error("dec_last: unsupported channel count");
  }
}

}  // namespace m3

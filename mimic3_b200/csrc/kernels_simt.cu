// fp32 SIMT kernels of the engine: everything on the duration-deciding text side
// (TextEncoder + duration predictor must stay fp32 FFMA: a 1-ulp change in logw near an
// integer boundary adds a whole frame, SURVEY.md §7 hard part 2) plus the generic
// segment-aware Conv1d-as-GEMM used wherever no tensor-core kernel applies.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <set>
#include <stdexcept>

#include "kernels.h"

namespace m3 {

thread_local int64_t g_launch_count = 0;

// Counts the launch; with M3B200_SYNC_DEBUG=1 also synchronises and names the failing kernel.
void post_launch(const char* what, cudaStream_t st) {
  ++g_launch_count;
  static const bool dbg = [] {
    const char* e = getenv("M3B200_SYNC_DEBUG");
    return e && *e && *e != '0';
  }();
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess && dbg) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) {
    char buf[256];
    snprintf(buf, sizeof buf, "kernel %s failed: %s", what, cudaGetErrorString(err));
    throw std::runtime_error(buf);
  }
}
// 256-bit global loads/stores in the row GEMM and the polyphase upsampler epilogue: on by default (measured
// 17.30 -> 16.49 ms per step, profiles/r01e_ab_wide_io.txt; bit-identical results); M3B200_WIDE_IO=0 restores the
// 128-bit accesses.
bool wide_io_enabled() {
  const char* e = getenv("M3B200_WIDE_IO");
  return !(e && *e == '0');
}

void ensure_max_dynamic_smem(const void* func) {
  static std::mutex mu;
  static std::set<const void*> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(func)) return;
  // 227 KB opt-in maximum minus room for static shared memory; the limit is only ever raised, never
  // lowered, so concurrent launches with different sizes cannot invalidate each other.
  cudaFuncAttributes fa;
  int optin = 227 * 1024, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (cudaFuncGetAttributes(&fa, func) != cudaSuccess) fa.sharedSizeBytes = 4096;
  const int dyn_max = optin - int(fa.sharedSizeBytes);  // static + dynamic must fit the opt-in maximum
  if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max) != cudaSuccess) {
    cudaGetLastError();
    throw std::runtime_error("cannot raise the dynamic shared-memory limit of a kernel");
  }
  done.insert(func);
}
#define M3_LAUNCHED() post_launch(__func__, st)

// -------------------------------------------------------------------------------------
// Philox4x32-10 noise source; specification in oracle/philox.py (kept bit-compatible).
// -------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t& o0, uint32_t& o1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  o0 = c0;
  o1 = c1;
}

__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t stream, uint32_t row, uint32_t pos,
                                               uint32_t chan) {
  uint32_t x0, x1;
  philox4x32_10(pos, chan, stream, row, uint32_t(seed), uint32_t(seed >> 32), x0, x1);
  float u1 = (float(x0 >> 9) + 0.5f) * 1.1920928955078125e-07f;  // 2^-23
  float u2 = (float(x1 >> 9) + 0.5f) * 1.1920928955078125e-07f;
  float r = sqrtf(-2.0f * logf(u1));
  return r * cosf(6.283185307179586f * u2);
}

// -------------------------------------------------------------------------------------
// Generic segment-aware Conv1d / ConvTranspose1d(polyphase) as a tiled fp32 GEMM.
// -------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16, APAD = 4;

__device__ __forceinline__ float lrelu(float x, float slope) { return x >= 0.f ? x : x * slope; }

template <bool GATE>
__global__ void __launch_bounds__(256) conv_gemm_kernel(ConvParams p) {
  const int seg = blockIdx.z / p.phases, phase = blockIdx.z % p.phases;
  const int len_units = p.seg_len[seg];
  const int in_len = len_units * p.in_scale;
  const int rows = in_len + p.rows_extra;
  const int t0 = blockIdx.x * BM;
  if (t0 >= rows) return;
  const int n0 = blockIdx.y * BN;
  const long long in_base = (long long)p.seg_off[seg] * p.in_scale;
  const int out_len = len_units * p.out_scale;
  const long long out_base = (long long)p.seg_off[seg] * p.out_scale;
  const int ldw = GATE ? 2 * p.Cout : p.Cout;
  const float* __restrict__ W = p.W + (long long)phase * p.w_phase_stride;

  __shared__ __align__(16) float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[GATE ? 2 : 1][BK][BN];

  float acc[4][4], acc2[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f, acc2[i][j] = 0.f;

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int a_r = tid >> 2, a_q = tid & 3;
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;
  const bool vecA = ((p.in_stride | p.in_coff) & 3) == 0;
  const bool vecB = (ldw & 3) == 0 && (p.Cout & 3) == 0;
  const float slope = p.in_slope;

  for (int tap = 0; tap < p.taps; ++tap) {
    const int ti = t0 + a_r + (tap - p.pad_left) * p.dil;
    const bool rvalid = ti >= 0 && ti < in_len;
    const float* __restrict__ arow = p.in + (in_base + ti) * (long long)p.in_stride + p.in_coff;
    for (int k0 = 0; k0 < p.Cin; k0 += BK) {
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kc = k0 + a_q * 4;
      if (rvalid) {
        if (vecA && kc + 3 < p.Cin) {
          av = *reinterpret_cast<const float4*>(arow + kc);
        } else {
          if (kc + 0 < p.Cin) av.x = arow[kc + 0];
          if (kc + 1 < p.Cin) av.y = arow[kc + 1];
          if (kc + 2 < p.Cin) av.z = arow[kc + 2];
          if (kc + 3 < p.Cin) av.w = arow[kc + 3];
        }
        if (slope != 1.f) {
          av.x = lrelu(av.x, slope); av.y = lrelu(av.y, slope);
          av.z = lrelu(av.z, slope); av.w = lrelu(av.w, slope);
        }
      }
      As[a_q * 4 + 0][a_r] = av.x;
      As[a_q * 4 + 1][a_r] = av.y;
      As[a_q * 4 + 2][a_r] = av.z;
      As[a_q * 4 + 3][a_r] = av.w;

      const int kk = k0 + b_k;
#pragma unroll
      for (int g = 0; g < (GATE ? 2 : 1); ++g) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kk < p.Cin) {
          const float* __restrict__ wrow = W + ((long long)tap * p.Cin + kk) * ldw + g * p.Cout;
          const int n = n0 + b_n;
          if (vecB && n + 3 < p.Cout) {
            bv = *reinterpret_cast<const float4*>(wrow + n);
          } else {
            if (n + 0 < p.Cout) bv.x = wrow[n + 0];
            if (n + 1 < p.Cout) bv.y = wrow[n + 1];
            if (n + 2 < p.Cout) bv.z = wrow[n + 2];
            if (n + 3 < p.Cout) bv.w = wrow[n + 3];
          }
        }
        *reinterpret_cast<float4*>(&Bs[g][b_k][b_n]) = bv;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[0][k][tx * 4]);
        const float ar[4] = {a.x, a.y, a.z, a.w};
        const float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        if (GATE) {
          const float4 b2 = *reinterpret_cast<const float4*>(&Bs[GATE ? 1 : 0][k][tx * 4]);
          const float br2[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[i][j] = fmaf(ar[i], br2[j], acc2[i][j]);
        }
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= rows) continue;
    const int po = t * p.out_mul + p.out_add + phase * p.out_add_phase;
    if (po < 0 || po >= out_len) continue;
    const long long orow = out_base + po;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.ubias) v += p.ubias[(long long)seg * p.ub_stride + n];
      if (GATE) {
        float vb = acc2[i][j];
        if (p.bias) vb += p.bias[n + p.Cout];
        if (p.ubias) vb += p.ubias[(long long)seg * p.ub_stride + n + p.Cout];
        v = tanhf(v) * (1.f / (1.f + expf(-vb)));
      }
      if (p.res) v += p.res[orow * p.res_stride + p.res_coff + n];
      v *= p.scale;
      if (p.act == 1) v = fmaxf(v, 0.f);
      float* dst = n < p.split ? p.out + orow * p.out_stride + p.out_coff + n
                               : p.out2 + orow * p.out2_stride + (n - p.split);
      if (p.mode == 0) *dst = v;
      else if (p.mode == 1) *dst += v;
      else *dst -= v;
    }
  }
}

void launch_conv(const ConvParams& p, int n_seg, int max_seg_len, cudaStream_t st) {
  const int rows = max_seg_len * p.in_scale + p.rows_extra;
  if (rows <= 0 || n_seg <= 0) return;
  dim3 grid((rows + BM - 1) / BM, (p.Cout + BN - 1) / BN, n_seg * p.phases);
  if (p.gate) conv_gemm_kernel<true><<<grid, 256, 0, st>>>(p);
  else conv_gemm_kernel<false><<<grid, 256, 0, st>>>(p);
  M3_LAUNCHED();
}

// -------------------------------------------------------------------------------------
// Token-level conv-as-GEMM over packed rows: 128x64 tile, 8x4 outputs per thread, BK = 16,
// register-prefetch double buffering (global loads of tile k+1 overlap the FMAs of tile k).
// fp32 FFMA on purpose: this feeds the duration predictor (SURVEY.md hard part 2).
// -------------------------------------------------------------------------------------
constexpr int RB_M = 128, RB_N = 64, RB_K = 16;

__global__ void fill_rowinfo_kernel(int4* rowinfo, const int* seg_off, const int* seg_len) {
  const int seg = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int len = seg_len[seg];
  if (t >= len) return;
  const int lo = seg_off[seg];
  rowinfo[lo + t] = make_int4(lo, lo + len, seg, 0);
}

void launch_fill_rowinfo(int4* rowinfo, const int* seg_off, const int* seg_len, int n_seg, int max_len,
                         cudaStream_t st) {
  if (max_len <= 0) return;
  fill_rowinfo_kernel<<<dim3((max_len + 127) / 128, n_seg), 128, 0, st>>>(rowinfo, seg_off, seg_len);
  M3_LAUNCHED();
}

__global__ void __launch_bounds__(256, 2) row_conv_kernel(RowConvParams p) {
  __shared__ __align__(16) float As[RB_K][RB_M + 4];
  __shared__ __align__(16) float Bs[RB_K][RB_N];
  const int r0 = blockIdx.x * RB_M, n0 = blockIdx.y * RB_N;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads: rows ty*8.., cols tx*4..
  // A loader: 128 rows x 16 k = 512 float4 -> 2 per thread (row = idx>>2, kq = idx&3)
  const int a_row0 = tid >> 2, a_kq = tid & 3;
  // B loader: 16 k x 64 n = 256 float4 -> 1 per thread
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;
  const bool vecB = (p.Cout & 3) == 0;

  int4 ri[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = r0 + a_row0 + h * 64;
    ri[h] = r < p.rows ? p.rowinfo[r] : make_int4(0, 0, 0, 0);
  }
  const int ksteps = (p.Cin + RB_K - 1) / RB_K;
  const int total = p.taps * ksteps;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 pa[2], pb;
  auto fetch = [&](int it) {
    const int tap = it / ksteps, k0 = (it - tap * ksteps) * RB_K;
    const int shift = tap - p.pad_left;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = r0 + a_row0 + h * 64 + shift;
      const int kc = k0 + a_kq * 4;
      pa[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r >= ri[h].x && r < ri[h].y && kc < p.Cin)  // Cin % 4 == 0 (checked on the host)
        pa[h] = *reinterpret_cast<const float4*>(p.in + (long long)r * p.in_stride + kc);
    }
    const int kk = k0 + b_k;
    pb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kk < p.Cin) {
      const float* wrow = p.W + ((long long)tap * p.Cin + kk) * p.Cout;
      const int n = n0 + b_n;
      if (vecB && n + 3 < p.Cout) pb = *reinterpret_cast<const float4*>(wrow + n);
      else {
        if (n + 0 < p.Cout) pb.x = wrow[n + 0];
        if (n + 1 < p.Cout) pb.y = wrow[n + 1];
        if (n + 2 < p.Cout) pb.z = wrow[n + 2];
        if (n + 3 < p.Cout) pb.w = wrow[n + 3];
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = a_row0 + h * 64;
      As[a_kq * 4 + 0][m] = pa[h].x;
      As[a_kq * 4 + 1][m] = pa[h].y;
      As[a_kq * 4 + 2][m] = pa[h].z;
      As[a_kq * 4 + 3][m] = pa[h].w;
    }
    *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = pb;
  };

  fetch(0);
  stash();
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    if (it + 1 < total) fetch(it + 1);
#pragma unroll
    for (int k = 0; k < RB_K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
    if (it + 1 < total) {
      stash();
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty * 8 + i;
    if (r >= p.rows) continue;
    const int seg = p.ubias ? p.rowinfo[r].z : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.ubias) v += p.ubias[(long long)seg * p.ub_stride + n];
      if (p.act == 1) v = fmaxf(v, 0.f);
      p.out[(long long)r * p.out_stride + n] = v;
    }
  }
}

void launch_row_conv(const RowConvParams& p, cudaStream_t st) {
  if (p.rows <= 0) return;
  if ((p.Cin & 3) || (p.in_stride & 3)) throw std::runtime_error("row_conv: Cin and in_stride must be multiples of 4");
  dim3 grid((p.rows + RB_M - 1) / RB_M, (p.Cout + RB_N - 1) / RB_N);
  row_conv_kernel<<<grid, 256, 0, st>>>(p);
  M3_LAUNCHED();
}

// -------------------------------------------------------------------------------------
// LayerNorm over channels (modules.LayerNorm [EXT]: biased variance, eps 1e-5).
// -------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }


__device__ __forceinline__ void warp_layernorm(float* v, int nper, int C, int lane, const float* gamma,
                                               const float* beta, int act) {
  float s = 0.f;
  for (int i = 0; i < nper; ++i)
    if (lane + 32 * i < C) s += v[i];
  const float mean = warp_sum(s) / float(C);
  float q = 0.f;
  for (int i = 0; i < nper; ++i)
    if (lane + 32 * i < C) {
      const float d = v[i] - mean;
      q += d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) / float(C) + 1e-5f);
  for (int i = 0; i < nper; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
      if (act == 1) y = gelu_erf(y);
      v[i] = y;
    }
  }
}

template <int NPER>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        const float* res, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* out, int rows, int C,
                                                        int act) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[NPER];
  const long long base = (long long)row * C;
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < C ? a[base + c] + (b ? b[base + c] : 0.f) : 0.f;
  }
  warp_layernorm(v, NPER, C, lane, gamma, beta, act);
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int c = lane + 32 * i;
    if (c < C) out[base + c] = v[i] + (res ? res[base + c] : 0.f);
  }
}

void launch_layernorm(const float* a, const float* b, const float* res, const float* gamma, const float* beta,
                      float* out, int rows, int C, int act, cudaStream_t st) {
  if (rows <= 0) return;
  dim3 grid((rows + 7) / 8);
  const int nper = (C + 31) / 32;
  if (nper <= 2) layernorm_kernel<2><<<grid, 256, 0, st>>>(a, b, res, gamma, beta, out, rows, C, act);
  else if (nper <= 6) layernorm_kernel<6><<<grid, 256, 0, st>>>(a, b, res, gamma, beta, out, rows, C, act);
  else if (nper <= 8) layernorm_kernel<8><<<grid, 256, 0, st>>>(a, b, res, gamma, beta, out, rows, C, act);
  else layernorm_kernel<32><<<grid, 256, 0, st>>>(a, b, res, gamma, beta, out, rows, C, act);
  M3_LAUNCHED();
}

// DDSConv first half: depthwise k=3 dilated conv -> LN -> GELU, one warp per row.
template <int NPER>
__global__ void __launch_bounds__(256) dds_sep_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias,
                                                      const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float* out, int C, int dil,
                                                      const int* seg_off, const int* seg_len) {
  const int seg = blockIdx.y;
  const int T = seg_len[seg];
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const long long base = seg_off[seg];
  float v[NPER];
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int c = lane + 32 * i;
    float s = 0.f;
    if (c < C) {
      s = bias[c];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int ti = t + (j - 1) * dil;
        if (ti >= 0 && ti < T) s = fmaf(w[j * C + c], x[(base + ti) * C + c], s);
      }
    }
    v[i] = s;
  }
  warp_layernorm(v, NPER, C, lane, gamma, beta, 1);
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int c = lane + 32 * i;
    if (c < C) out[(base + t) * C + c] = v[i];
  }
}

void launch_dds_sep(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                    float* out, int C, int dil, const int* seg_off, const int* seg_len, int n_seg, int max_len,
                    cudaStream_t st) {
  if (max_len <= 0) return;
  dim3 grid((max_len + 7) / 8, n_seg);
  const int nper = (C + 31) / 32;
  if (nper <= 2) dds_sep_kernel<2><<<grid, 256, 0, st>>>(x, w, bias, gamma, beta, out, C, dil, seg_off, seg_len);
  else if (nper <= 6) dds_sep_kernel<6><<<grid, 256, 0, st>>>(x, w, bias, gamma, beta, out, C, dil, seg_off, seg_len);
  else if (nper <= 8) dds_sep_kernel<8><<<grid, 256, 0, st>>>(x, w, bias, gamma, beta, out, C, dil, seg_off, seg_len);
  else dds_sep_kernel<32><<<grid, 256, 0, st>>>(x, w, bias, gamma, beta, out, C, dil, seg_off, seg_len);
  M3_LAUNCHED();
}

// -------------------------------------------------------------------------------------
// Embedding / gathers
// -------------------------------------------------------------------------------------
__global__ void embedding_kernel(const int64_t* __restrict__ ids, int t_stride, const float* __restrict__ emb,
                                 int num_symbols, float scale, float* out, int H, const int* seg_off,
                                 const int* seg_len) {
  const int seg = blockIdx.y, t = blockIdx.x;
  if (t >= seg_len[seg]) return;
  int64_t id = ids[(long long)seg * t_stride + t];
  id = id < 0 ? 0 : (id >= num_symbols ? num_symbols - 1 : id);  // host validates; device ids are clamped
  const long long row = seg_off[seg] + t;
  for (int c = threadIdx.x; c < H; c += blockDim.x) out[row * H + c] = emb[id * H + c] * scale;
}

void launch_embedding(const int64_t* ids, int t_stride, const float* emb, int num_symbols, float scale, float* out,
                      int H, const int* seg_off, const int* seg_len, int n_seg, int max_len, cudaStream_t st) {
  if (max_len <= 0) return;
  embedding_kernel<<<dim3(max_len, n_seg), 64, 0, st>>>(ids, t_stride, emb, num_symbols, scale, out, H, seg_off,
                                                        seg_len);
  M3_LAUNCHED();
}

__global__ void add_ubias_kernel(const float* __restrict__ in, const float* __restrict__ ubias, int ub_stride,
                                 float* out, int C, const int* seg_off, const int* seg_len) {
  const int seg = blockIdx.y, t = blockIdx.x;
  if (t >= seg_len[seg]) return;
  const long long row = seg_off[seg] + t;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    out[row * C + c] = in[row * C + c] + ubias[(long long)seg * ub_stride + c];
}

void launch_add_ubias(const float* in, const float* ubias, int ub_stride, float* out, int C, const int* seg_off,
                      const int* seg_len, int n_seg, int max_len, cudaStream_t st) {
  if (max_len <= 0) return;
  add_ubias_kernel<<<dim3(max_len, n_seg), 64, 0, st>>>(in, ubias, ub_stride, out, C, seg_off, seg_len);
  M3_LAUNCHED();
}

__global__ void gather_rows_kernel(const int64_t* __restrict__ idx, const float* __restrict__ table, float* out,
                                   int C) {
  const int r = blockIdx.x;
  const int64_t id = idx[r];
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[(long long)r * C + c] = table[id * C + c];
}

void launch_gather_rows(const int64_t* idx, const float* table, float* out, int n, int C, cudaStream_t st) {
  if (n <= 0) return;
  gather_rows_kernel<<<n, 128, 0, st>>>(idx, table, out, C);
  M3_LAUNCHED();
}

// -------------------------------------------------------------------------------------
// Relative-position multi-head attention (attentions.MultiHeadAttention [EXT], window W,
// relative embeddings shared by heads).  Flash-style: 64 queries x 64-key tiles, online
// softmax, fp32 throughout.
// -------------------------------------------------------------------------------------
constexpr int AQ = 64, AKT = 64;

template <int NE>
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv,
                                                        const float* __restrict__ Ek,
                                                        const float* __restrict__ Ev, float* out, int H, int dk,
                                                        int W, const int* seg_off, const int* seg_len) {
  extern __shared__ __align__(16) float sm[];
  const int seg = blockIdx.z, h = blockIdx.y;
  const int T = seg_len[seg];
  const int q0 = blockIdx.x * AQ;
  if (q0 >= T) return;
  const long long base = seg_off[seg];
  const int nrel = 2 * W + 1;
  const int ldq = AQ + 4;
  float* Qs = sm;                      // [dk][AQ+4]
  float* Ks = Qs + dk * ldq;           // [dk][AKT+4]
  float* Vs = Ks + dk * ldq;           // [AKT][dk]
  float* Ss = Vs + AKT * dk;           // [AQ][AKT+1]
  float* RQ = Ss + AQ * (AKT + 1);     // [AQ][nrel]
  float* Eks = RQ + AQ * nrel;         // [nrel][dk]
  float* Evs = Eks + nrel * dk;        // [nrel][dk]
  float* rowm = Evs + nrel * dk;       // [AQ]
  float* rowl = rowm + AQ;
  float* rowa = rowl + AQ;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, warp = tid >> 5, lane = tid & 31;
  const int H3 = 3 * H;
  const float sq = sqrtf(float(dk));

  for (int idx = tid; idx < AQ * dk; idx += 256) {
    const int i = idx / dk, d = idx - i * dk;
    float v = 0.f;
    if (q0 + i < T) v = qkv[(base + q0 + i) * H3 + h * dk + d] / sq;
    Qs[d * ldq + i] = v;
  }
  for (int idx = tid; idx < nrel * dk; idx += 256) {
    Eks[idx] = Ek[idx];
    Evs[idx] = Ev[idx];
  }
  if (tid < AQ) {
    rowm[tid] = -INFINITY;
    rowl[tid] = 0.f;
  }
  __syncthreads();
  for (int idx = tid; idx < AQ * nrel; idx += 256) {
    const int i = idx / nrel, r = idx - i * nrel;
    float s = 0.f;
    for (int d = 0; d < dk; ++d) s = fmaf(Qs[d * ldq + i], Eks[r * dk + d], s);
    RQ[idx] = s;
  }

  float acc[4][NE];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < NE; ++e) acc[i][e] = 0.f;

  for (int k0 = 0; k0 < T; k0 += AKT) {
    __syncthreads();
    for (int idx = tid; idx < AKT * dk; idx += 256) {
      const int j = idx / dk, d = idx - j * dk;
      float kv = 0.f, vv = 0.f;
      if (k0 + j < T) {
        const float* row = qkv + (base + k0 + j) * H3 + h * dk + d;
        kv = row[H];
        vv = row[2 * H];
      }
      Ks[d * ldq + j] = kv;
      Vs[j * dk + d] = vv;
    }
    __syncthreads();
    {
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
      for (int d = 0; d < dk; ++d) {
        const float4 a = *reinterpret_cast<const float4*>(&Qs[d * ldq + ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Ks[d * ldq + tx * 4]);
        const float ar[4] = {a.x, a.y, a.z, a.w};
        const float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) s[i][j] = fmaf(ar[i], br[j], s[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int li = ty * 4 + i, lj = tx * 4 + j;
          const int gi = q0 + li, gj = k0 + lj;
          const int rel = gj - gi;
          float v = s[i][j];
          if (rel >= -W && rel <= W) v += RQ[li * nrel + rel + W];
          if (gj >= T) v = -INFINITY;
          Ss[li * (AKT + 1) + lj] = v;
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int rr = 0; rr < 8; ++rr) {
      const int i = warp * 8 + rr;
      const float v0 = Ss[i * (AKT + 1) + lane], v1 = Ss[i * (AKT + 1) + lane + 32];
      const float mx = warp_max(fmaxf(v0, v1));
      const float m_old = rowm[i];
      const float m_new = fmaxf(m_old, mx);
      const float p0 = expf(v0 - m_new), p1 = expf(v1 - m_new);
      const float sum = warp_sum(p0 + p1);
      Ss[i * (AKT + 1) + lane] = p0;
      Ss[i * (AKT + 1) + lane + 32] = p1;
      __syncwarp();
      if (lane == 0) {
        const float alpha = expf(m_old - m_new);
        rowa[i] = alpha;
        rowl[i] = rowl[i] * alpha + sum;
        rowm[i] = m_new;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = rowa[ty * 4 + i];
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[i][e] *= a;
    }
    for (int j = 0; j < AKT; ++j) {
      float vv[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int d = tx + 16 * e;
        vv[e] = d < dk ? Vs[j * dk + d] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pv = Ss[(ty * 4 + i) * (AKT + 1) + j];
#pragma unroll
        for (int e = 0; e < NE; ++e) acc[i][e] = fmaf(pv, vv[e], acc[i][e]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int li = ty * 4 + i, gi = q0 + li;
      for (int r = -W; r <= W; ++r) {
        const int gj = gi + r, jl = gj - k0;
        if (jl >= 0 && jl < AKT && gj < T) {
          const float pv = Ss[li * (AKT + 1) + jl];
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const int d = tx + 16 * e;
            if (d < dk) acc[i][e] = fmaf(pv, Evs[(r + W) * dk + d], acc[i][e]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int li = ty * 4 + i, gi = q0 + li;
    if (gi >= T) continue;
    const float l = rowl[li];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int d = tx + 16 * e;
      if (d < dk) out[(base + gi) * H + h * dk + d] = acc[i][e] / l;
    }
  }
}

void launch_attention(const float* qkv, const float* emb_rel_k, const float* emb_rel_v, float* out, int H,
                      int n_heads, int window, const int* seg_off, const int* seg_len, int n_seg, int max_len,
                      cudaStream_t st) {
  if (max_len <= 0) return;
  // short sequences: whole (utterance, head) resident in shared memory (kernels_attn.cu)
  if (launch_attention_short(qkv, emb_rel_k, emb_rel_v, out, H, n_heads, window, seg_off, seg_len, n_seg, max_len, st)) return;
  const int dk = H / n_heads;
  const int nrel = 2 * window + 1;
  const size_t smem = sizeof(float) * (size_t(2) * dk * (AQ + 4) + AKT * dk + AQ * (AKT + 1) + AQ * nrel +
                                       2 * nrel * dk + 3 * AQ);
  dim3 grid((max_len + AQ - 1) / AQ, n_heads, n_seg);
  const int ne = (dk + 15) / 16;
#define M3_ATTN(NE)                                                                                          \
  {                                                                                                          \
    ensure_max_dynamic_smem(reinterpret_cast<const void*>(attention_kernel<NE>));                           \
    attention_kernel<NE><<<grid, 256, smem, st>>>(qkv, emb_rel_k, emb_rel_v, out, H, dk, window, seg_off,    \
                                                   seg_len);                                                 \
  }
  if (ne <= 1) M3_ATTN(1)
  else if (ne <= 2) M3_ATTN(2)
  else if (ne <= 4) M3_ATTN(4)
  else if (ne <= 6) M3_ATTN(6)
  else M3_ATTN(8)
#undef M3_ATTN
  M3_LAUNCHED();
}

// -------------------------------------------------------------------------------------
// Stochastic duration predictor pieces
// -------------------------------------------------------------------------------------
__global__ void convflow_pre_kernel(const float* __restrict__ z, int zc, const float* __restrict__ w,
                                    const float* __restrict__ b, const float* __restrict__ h, float* out, int rows,
                                    int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * C) return;
  const int t = int(idx / C), c = int(idx - (long long)t * C);
  out[idx] = fmaf(z[t * 2 + zc], w[c], b[c]) + h[idx];
}

void launch_convflow_pre(const float* z, int zc, const float* w, const float* b, const float* h, float* out,
                         int rows, int C, cudaStream_t st) {
  if (rows <= 0) return;
  const long long n = (long long)rows * C;
  convflow_pre_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(z, zc, w, b, h, out, rows, C);
  M3_LAUNCHED();
}

constexpr int RQ_BINS = 10;

__device__ void rqs_knots(const float* u, float* cum, float* width) {
  float mx = u[0];
#pragma unroll
  for (int i = 1; i < RQ_BINS; ++i) mx = fmaxf(mx, u[i]);
  float e[RQ_BINS], sum = 0.f;
#pragma unroll
  for (int i = 0; i < RQ_BINS; ++i) {
    e[i] = expf(u[i] - mx);
    sum += e[i];
  }
  float c = 0.f;
  cum[0] = -5.f;
#pragma unroll
  for (int i = 0; i < RQ_BINS; ++i) {
    const float wv = 1e-3f + (1.f - 1e-3f * RQ_BINS) * (e[i] / sum);
    c += wv;
    cum[i + 1] = 10.f * c + -5.f;
  }
  cum[RQ_BINS] = 5.f;
#pragma unroll
  for (int i = 0; i < RQ_BINS; ++i) width[i] = cum[i + 1] - cum[i];
}

__global__ void rqs_inverse_kernel(float* z, int zc, const float* __restrict__ params, int pstride,
                                   float inv_sqrt_c, int rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows) return;
  const float y = z[t * 2 + zc];
  if (!(y >= -5.f && y <= 5.f)) return;  // linear tails: identity
  const float* pr = params + (long long)t * pstride;
  float uw[RQ_BINS], uh[RQ_BINS];
#pragma unroll
  for (int i = 0; i < RQ_BINS; ++i) {
    uw[i] = pr[i] * inv_sqrt_c;
    uh[i] = pr[RQ_BINS + i] * inv_sqrt_c;
  }
  float cumw[RQ_BINS + 1], widths[RQ_BINS], cumh[RQ_BINS + 1], heights[RQ_BINS];
  rqs_knots(uw, cumw, widths);
  rqs_knots(uh, cumh, heights);
  int bin = -1;
#pragma unroll
  for (int k = 0; k <= RQ_BINS; ++k) {
    const float loc = k == RQ_BINS ? cumh[k] + 1e-6f : cumh[k];
    bin += (y >= loc) ? 1 : 0;
  }
  bin = min(max(bin, 0), RQ_BINS - 1);
  const float cst = 0.5397424101829529f;  // fp32(log(exp(1 - 1e-3) - 1)), as the reference stores it
  auto deriv = [&](int k) -> float {
    const float u = (k == 0 || k == RQ_BINS) ? cst : pr[2 * RQ_BINS + k - 1];
    const float sp = u > 20.f ? u : log1pf(expf(u));
    return 1e-3f + sp;
  };
  const float d0 = deriv(bin), d1 = deriv(bin + 1);
  const float in_w = widths[bin], in_cumw = cumw[bin], in_cumh = cumh[bin], in_h = heights[bin];
  const float delta = in_h / in_w;
  const float dy = y - in_cumh;
  const float s = d0 + d1 - 2.f * delta;
  const float a = dy * s + in_h * (delta - d0);
  const float b = in_h * d0 - dy * s;
  const float c = -delta * dy;
  const float disc = b * b - 4.f * a * c;
  const float root = (2.f * c) / (-b - sqrtf(disc));
  z[t * 2 + zc] = root * in_w + in_cumw;
}

void launch_rqs_inverse(float* z, int zc, const float* params, int pstride, float inv_sqrt_c, int rows,
                        cudaStream_t st) {
  if (rows <= 0) return;
  rqs_inverse_kernel<<<(rows + 127) / 128, 128, 0, st>>>(z, zc, params, pstride, inv_sqrt_c, rows);
  M3_LAUNCHED();
}

__global__ void sdp_noise_kernel(float* z, float noise_w, const float* __restrict__ row_scales, uint64_t seed,
                                 const int* seg_off, const int* seg_len) {
  const int seg = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= seg_len[seg]) return;
  if (row_scales) noise_w = row_scales[seg * 3 + 2];
  const long long row = seg_off[seg] + t;
#pragma unroll
  for (int c = 0; c < 2; ++c)
    z[row * 2 + c] = noise_w != 0.f ? philox_normal(seed, 0u, uint32_t(seg), uint32_t(t), uint32_t(c)) * noise_w : 0.f;
}

void launch_sdp_noise(float* z, float noise_w, const float* row_scales, uint64_t seed, const int* seg_off,
                      const int* seg_len, int n_seg, int max_len, cudaStream_t st) {
  if (max_len <= 0) return;
  sdp_noise_kernel<<<dim3((max_len + 127) / 128, n_seg), 128, 0, st>>>(z, noise_w, row_scales, seed, seg_off, seg_len);
  M3_LAUNCHED();
}

__global__ void sdp_finish_kernel(const float* __restrict__ z, int zc, float m, float logs, float* logw, int rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows) return;
  logw[t] = (z[t * 2 + zc] - m) * expf(-logs);
}

void launch_sdp_finish(const float* z, int zc, float m, float logs, float* logw, int rows, cudaStream_t st) {
  if (rows <= 0) return;
  sdp_finish_kernel<<<(rows + 255) / 256, 256, 0, st>>>(z, zc, m, logs, logw, rows);
  M3_LAUNCHED();
}

// -------------------------------------------------------------------------------------
// Durations -> monotonic alignment -> expanded prior
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) durations_kernel(const float* __restrict__ logw, int logw_stride,
                                                        float length_scale, const float* __restrict__ row_scales,
                                                        int* cum, int* frames, const int* seg_off,
                                                        const int* seg_len) {
  const int seg = blockIdx.x;
  const int T = seg_len[seg];
  if (row_scales) length_scale = row_scales[seg * 3 + 1];
  const long long base = seg_off[seg];
  __shared__ int warp_tot[8];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = 0; t0 < T; t0 += 256) {
    const int t = t0 + threadIdx.x;
    int d = 0;
    if (t < T) {
      const float w = expf(logw[(base + t) * logw_stride]) * length_scale;
      const float wc = ceilf(w);
      d = wc > 0.f ? (wc < 1.0e6f ? int(wc) : 1000000) : 0;  // NaN / negative -> 0
    }
    int v = d;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    int pre = carry_s;
    for (int wv = 0; wv < warp; ++wv) pre += warp_tot[wv];
    if (t < T) cum[base + t] = pre + v;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = pre + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) frames[seg] = max(1, carry_s);
}

void launch_durations(const float* logw, int logw_stride, float length_scale, const float* row_scales, int* cum,
                      int* frames, const int* seg_off, const int* seg_len, int n_seg, cudaStream_t st) {
  if (n_seg <= 0) return;
  durations_kernel<<<n_seg, 256, 0, st>>>(logw, logw_stride, length_scale, row_scales, cum, frames, seg_off, seg_len);
  M3_LAUNCHED();
}

__global__ void __launch_bounds__(256) expand_kernel(const float* __restrict__ stats, int I,
                                                     const int* __restrict__ cum, const int* tok_off,
                                                     const int* tok_len, const int* frm_off, const int* frm_len,
                                                     float noise_scale, const float* __restrict__ row_scales,
                                                     uint64_t seed, float* zp) {
  const int seg = blockIdx.y;
  if (row_scales) noise_scale = row_scales[seg * 3];
  const int F = frm_len[seg];
  const int y = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (y >= F) return;
  const int T = tok_len[seg];
  const int* c = cum + tok_off[seg];
  int lo = 0, hi = T;  // first t with cum[t] > y
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (c[mid] > y) hi = mid;
    else lo = mid + 1;
  }
  const long long orow = frm_off[seg] + y;
  const float* st = lo < T ? stats + (long long)(tok_off[seg] + lo) * 2 * I : nullptr;
  for (int ch = lane; ch < I; ch += 32) {
    float m = 0.f, ls = 0.f;
    if (st) {
      m = st[ch];
      ls = st[I + ch];
    }
    float v = m;
    if (noise_scale != 0.f)
      v = m + philox_normal(seed, 1u, uint32_t(seg), uint32_t(y), uint32_t(ch)) * expf(ls) * noise_scale;
    zp[orow * I + ch] = v;
  }
}

void launch_expand(const float* stats, int I, const int* cum, const int* tok_off, const int* tok_len,
                   const int* frm_off, const int* frm_len, int n_seg, int max_frames, float noise_scale,
                   const float* row_scales, uint64_t seed, float* zp, cudaStream_t st) {
  if (max_frames <= 0) return;
  expand_kernel<<<dim3((max_frames + 7) / 8, n_seg), 256, 0, st>>>(stats, I, cum, tok_off, tok_len, frm_off,
                                                                     frm_len, noise_scale, row_scales, seed, zp);
  M3_LAUNCHED();
}

__global__ void flip_channels_kernel(float* x, int rows, int C) {
  const int half = C / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * half) return;
  const long long r = idx / half;
  const int c = int(idx - r * half);
  float* row = x + r * C;
  const float a = row[c], b = row[C - 1 - c];
  row[c] = b;
  row[C - 1 - c] = a;
}

void launch_flip_channels(float* x, int rows, int C, cudaStream_t st) {
  if (rows <= 0) return;
  const long long n = (long long)rows * (C / 2);
  flip_channels_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(x, rows, C);
  M3_LAUNCHED();
}

// -------------------------------------------------------------------------------------
// conv_post (C -> 1, no bias) + tanh + per-utterance peak; int16 conversion
// -------------------------------------------------------------------------------------
// One CTA: 256 consecutive samples.  The (256 + k - 1) x C input tile is staged in smem with
// coalesced float4 loads (lrelu applied on the way in); rows are padded to C+1 floats so the
// per-thread row walk is bank-conflict free.
__global__ void __launch_bounds__(256) conv_post_kernel(const float* __restrict__ x, int C,
                                                        const float* __restrict__ w, int k, float slope,
                                                        float* audio, unsigned* peak_bits, const int* seg_off,
                                                        const int* seg_len, int scale) {
  extern __shared__ float sm[];
  const int seg = blockIdx.y;
  const int L = seg_len[seg] * scale;
  const int t0 = blockIdx.x * 256;
  if (t0 >= L) return;
  const long long base = (long long)seg_off[seg] * scale;
  const int pad = (k - 1) / 2;
  const int rows = 256 + k - 1;
  const int ld = C + 1;
  float* ws = sm;             // [k][C]
  float* tile = sm + k * C;   // [rows][C+1]
  for (int i = threadIdx.x; i < k * C; i += 256) ws[i] = w[i];
  const int c4n = C / 4;
  for (int i = threadIdx.x; i < rows * c4n; i += 256) {
    const int r = i / c4n, c4 = i - r * c4n;
    const int ti = t0 + r - pad;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ti >= 0 && ti < L) v = *reinterpret_cast<const float4*>(x + (base + ti) * C + c4 * 4);
    float* d = tile + r * ld + c4 * 4;
    d[0] = lrelu(v.x, slope);
    d[1] = lrelu(v.y, slope);
    d[2] = lrelu(v.z, slope);
    d[3] = lrelu(v.w, slope);
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  float y = 0.f;
  if (t < L) {
    float s = 0.f;
    for (int j = 0; j < k; ++j) {
      const float* row = tile + (threadIdx.x + j) * ld;
      const float* wj = ws + j * C;
#pragma unroll 8
      for (int c = 0; c < C; ++c) s = fmaf(wj[c], row[c], s);
    }
    y = tanhf(s);
    audio[base + t] = y;
  }
  float m = warp_max(fabsf(y));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, wm[i]);
    atomicMax(peak_bits + seg, __float_as_uint(m));
  }
}

void launch_conv_post(const float* x, int C, const float* w, int k, float slope, float* audio, unsigned* peak_bits,
                      const int* seg_off, const int* seg_len, int scale, int n_seg, int max_len, cudaStream_t st) {
  if (max_len <= 0) return;
  if (C % 4) throw std::runtime_error("conv_post: channel count must be a multiple of 4");
  const size_t smem = sizeof(float) * (size_t(k) * C + size_t(256 + k - 1) * (C + 1));
  if (smem > 48 * 1024) ensure_max_dynamic_smem(reinterpret_cast<const void*>(conv_post_kernel));
  conv_post_kernel<<<dim3((max_len * scale + 255) / 256, n_seg), 256, smem, st>>>(x, C, w, k, slope, audio,
                                                                                  peak_bits, seg_off, seg_len, scale);
  M3_LAUNCHED();
}

__global__ void to_int16_kernel(const float* __restrict__ audio, const unsigned* __restrict__ peak_bits,
                                int16_t* pcm, const int* seg_off, const int* seg_len, int scale) {
  const int seg = blockIdx.y;
  const int L = seg_len[seg] * scale;
  const long long base = (long long)seg_off[seg] * scale;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const float peak = fmaxf(0.01f, __uint_as_float(peak_bits[seg]));
  const float s = __fdiv_rn(32767.0f, peak);
  float y = __fmul_rn(audio[base + t], s);
  y = fminf(fmaxf(y, -32767.0f), 32767.0f);
  pcm[base + t] = int16_t(int(y));  // truncation toward zero, like ndarray.astype("int16")
}

void launch_to_int16(const float* audio, const unsigned* peak_bits, int16_t* pcm, const int* seg_off,
                     const int* seg_len, int scale, int n_seg, int max_len, cudaStream_t st) {
  if (max_len <= 0) return;
  to_int16_kernel<<<dim3((max_len * scale + 255) / 256, n_seg), 256, 0, st>>>(audio, peak_bits, pcm, seg_off,
                                                                              seg_len, scale);
  M3_LAUNCHED();
}

// int16 conversion + the PCM post chain of Mimic3TextToSpeechSystem._speak_sentence_phonemes
// (mimic3_tts/tts.py:536-543): utterance b lands at out_off[b] of the output stream (the gaps are the
// silences of add_break, tts.py:452-465, zero-filled by the caller) and, when volume != NULL, goes through
// audioop.mul(bytes, 2, factor): val = sample * factor in double; > 32767 -> 32767; < -32767 -> -32768;
// floor.  (CPython Modules/audioop.c, fbound(); pinned against the real module in tests/test_post_chain.py.)
__global__ void to_int16_post_kernel(const float* __restrict__ audio, const unsigned* __restrict__ peak_bits,
                                     int16_t* stream, const int* seg_off, const int* seg_len, int scale,
                                     const long long* __restrict__ out_off, const double* __restrict__ volume) {
  const int seg = blockIdx.y;
  const int L = seg_len[seg] * scale;
  const long long base = (long long)seg_off[seg] * scale;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const float peak = fmaxf(0.01f, __uint_as_float(peak_bits[seg]));
  const float s = __fdiv_rn(32767.0f, peak);
  float y = __fmul_rn(audio[base + t], s);
  y = fminf(fmaxf(y, -32767.0f), 32767.0f);
  int v = int(y);
  if (volume) {
    double d = __dmul_rn(double(v), volume[seg]);
    if (d > 32767.0) d = 32767.0;
    else if (d < -32767.0) d = -32768.0;
    v = int(floor(d));
  }
  stream[out_off[seg] + t] = int16_t(v);
}

void launch_to_int16_post(const float* audio, const unsigned* peak_bits, int16_t* stream, const int* seg_off,
                          const int* seg_len, int scale, const long long* out_off, const double* volume, int n_seg,
                          int max_len, cudaStream_t st) {
  if (max_len <= 0) return;
  to_int16_post_kernel<<<dim3((max_len * scale + 255) / 256, n_seg), 256, 0, st>>>(audio, peak_bits, stream, seg_off,
                                                                                   seg_len, scale, out_off, volume);
  M3_LAUNCHED();
}

}  // namespace m3

// Token-level Conv1d-as-GEMM on tensor cores with fp32-equivalent products ("fp16 x 3"):
//     a = a_hi + a_lo / 2048,   a_hi = fp16(a),  a_lo = fp16((a - a_hi) * 2048)      (same for w)
//     a*w ~= a_hi*w_hi + (a_hi*w_lo + a_lo*w_hi) / 2048          (dropped term: 2^-22 relative)
// Two fp32 TMEM accumulators per tile (main, correction), three tcgen05.mma per K-step.  This is what
// lets the text encoder / duration predictor leave the fp32 FFMA pipe without moving `logw` (a 1-ulp
// class change there can add a whole frame, SURVEY.md hard part 2).
//
// Rows are the packed tokens of all utterances with ONE virtual zero row after each utterance, so a
// k=3 tap is a descriptor shift even across utterance boundaries (vmap[v] = physical row or -1).
// CTA = 128 virtual rows x NC output columns (64, or 128 for the 1x1 layers with M3B200_ROWGEMM_NC=128: half as
// many CTAs re-stage the same A rows); K streamed in blocks of 32 through a 2- or 3-stage ring:
//   warps 0-3  convert the fp32 A block to hi/lo fp16 (interleaved layout) -- and run the epilogue;
//   warp 4     one thread: bulk-copies the pre-split weight block of all taps (one copy per stage);
//   warp 5     one thread: issues the MMAs, frees stages with tcgen05.commit.
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"
#include "tc_common.cuh"

namespace m3 {

constexpr int RG_KB = 32;       // K block
constexpr int RG_THREADS = 192;

__global__ void fill_vmap_kernel(int* vmap, const int* seg_off, const int* seg_len, int n_seg) {
  const int seg = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int len = seg_len[seg];
  if (t > len) return;
  const int v = seg_off[seg] + seg + t;  // one gap row after every utterance
  vmap[v] = t < len ? seg_off[seg] + t : -1;
}

void launch_fill_vmap(int* vmap, const int* seg_off, const int* seg_len, int n_seg, int max_len, cudaStream_t st) {
  if (max_len <= 0) return;
  fill_vmap_kernel<<<dim3((max_len + 1 + 127) / 128, n_seg), 128, 0, st>>>(vmap, seg_off, seg_len, n_seg);
  post_launch("fill_vmap_kernel", st);
}

template <int RG_NC, int RG_STAGES>
__global__ void __launch_bounds__(RG_THREADS, 4) rowgemm_tc_kernel(RowGemmTcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t a_full[RG_STAGES], w_full[RG_STAGES], empty_bar[RG_STAGES], acc_full;

  const int v0 = blockIdx.x * 128;
  const int chunk = blockIdx.y;
  const int taps = p.taps;
  const int RA = (128 + taps - 1) | 1;                 // A rows per stage (odd pitch)
  const uint32_t a_bytes = 2u * (RG_KB / 8) * RA * 16; // hi + lo
  const uint32_t w_bytes = uint32_t(taps) * 2u * RG_KB * RG_NC * 2;
  const uint32_t a_pad = (a_bytes + 127u) & ~127u;
  const uint32_t stage_bytes = a_pad + w_bytes;
  const int nkb = p.K / RG_KB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tc::tmem_alloc<2 * RG_NC>(&tmem_slot);
  if (tid == 32) {
    for (int s = 0; s < RG_STAGES; ++s) {
      tc::mbar_init(&a_full[s], 4);
      tc::mbar_init(&w_full[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(&acc_full, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ===================== A stagers (then epilogue) =====================
    const int pad_left = p.pad_left;
    const int items = (RG_KB / 8) * RA;  // (row, 8-channel chunk) per stage, 128 threads
    constexpr int MAXI = 5;              // items per thread: ceil(4 * 131 / 128)
    // physical row of each of this thread's items (same for every K block)
    int physr[MAXI], dsto[MAXI];
#pragma unroll
    for (int u = 0; u < MAXI; ++u) {
      const int idx = tid + u * 128;
      physr[u] = -2;  // no item
      dsto[u] = 0;
      if (idx < items) {
        const int rr = idx / (RG_KB / 8), c8 = idx - rr * (RG_KB / 8);
        const int v = v0 - pad_left + rr;
        physr[u] = (v >= 0 && v < p.vrows) ? p.vmap[v] : -1;
        dsto[u] = c8 * RA + rr;
      }
    }
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % RG_STAGES;
      // global loads first (all in flight together), THEN wait for the ring slot: the L2 latency of block
      // kb overlaps the MMAs that are still draining the slot
      float4 la[MAXI], lb[MAXI];
#pragma unroll
      for (int u = 0; u < MAXI; ++u) {
        la[u] = lb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (physr[u] >= 0) {
          const int c8 = dsto[u] / RA;
          const float* src = p.in + (long long)physr[u] * p.in_stride + kb * RG_KB + c8 * 8;
          if (p.wide) {  // one 32-byte sector per lane and instruction
            float t8[8];
            tc::ldg256(src, t8);
            la[u] = make_float4(t8[0], t8[1], t8[2], t8[3]);
            lb[u] = make_float4(t8[4], t8[5], t8[6], t8[7]);
          } else {
            la[u] = *reinterpret_cast<const float4*>(src);
            lb[u] = *reinterpret_cast<const float4*>(src + 4);
          }
        }
      }
      tc::mbar_wait(&empty_bar[s], (((kb / RG_STAGES) & 1) ^ 1));
      uint8_t* abuf = smem + size_t(s) * stage_bytes;
#pragma unroll
      for (int u = 0; u < MAXI; ++u) {
        if (physr[u] == -2) continue;
        const float f[8] = {la[u].x, la[u].y, la[u].z, la[u].w, lb[u].x, lb[u].y, lb[u].z, lb[u].w};
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 hh = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
          const float2 back = __half22float2(hh);
          const __half2 ll = __floats2half2_rn((f[2 * e] - back.x) * 2048.f, (f[2 * e + 1] - back.y) * 2048.f);
          h[e] = *reinterpret_cast<const uint32_t*>(&hh);
          l[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(abuf + size_t(dsto[u]) * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(abuf + (size_t(RG_KB / 8) * RA + dsto[u]) * 16) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&a_full[s]);
    }
    // ---- epilogue: out = main + corr/2048 + bias ----
    tc::mbar_wait(&acc_full, 0);
    tc::fence_after_sync();
    const int v = v0 + warp * 32 + lane;
    const int phys = (v < p.vrows) ? p.vmap[v] : -1;
    const uint32_t lane_base = tmem + (uint32_t(warp * 32) << 16);
    const int seg = (p.ubias && phys >= 0) ? p.rowinfo[phys].z : 0;
    for (int c0 = 0; c0 < RG_NC; c0 += 16) {
      __syncwarp();
      float m[16], c[16];
      tc::tmem_ld16(lane_base + c0, m);
      tc::tmem_ld16(lane_base + RG_NC + c0, c);
      tc::tmem_ld_wait();
      if (phys < 0) continue;
      const int n0 = chunk * RG_NC + c0;
      if (n0 >= p.N) continue;
      float* dst = p.out + (long long)phys * p.out_stride + n0;
      float val[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int n = n0 + e < p.N ? n0 + e : p.N - 1;
        float v = m[e] + c[e] * (1.0f / 2048.0f);
        if (p.bias) v += p.bias[n];
        if (p.ubias) v += p.ubias[(long long)seg * p.ub_stride + n];
        if (p.act == 1) v = fmaxf(v, 0.f);
        val[e] = v;
      }
      if (p.wide && n0 + 16 <= p.N && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {  // 2 x 256-bit stores per row
        tc::stg256(dst, val);
        tc::stg256(dst + 8, val + 8);
      } else if (n0 + 16 <= p.N && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {  // 4 x 128-bit stores per row
#pragma unroll
        for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(dst + e) = make_float4(val[e], val[e + 1], val[e + 2], val[e + 3]);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (n0 + e < p.N) dst[e] = val[e];
      }
    }
  } else if (warp == 4) {
    // ===================== weight producer =====================
    if (tc::elect_one()) {
      const uint16_t* src = p.w + size_t(chunk) * nkb * (size_t(taps) * 2 * RG_KB * RG_NC);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % RG_STAGES;
        tc::mbar_wait(&empty_bar[s], (((kb / RG_STAGES) & 1) ^ 1));
        tc::mbar_expect_tx(&w_full[s], w_bytes);
        tc::bulk_g2s(smem + size_t(s) * stage_bytes + a_pad, src + size_t(kb) * (size_t(taps) * 2 * RG_KB * RG_NC), w_bytes,
                     &w_full[s]);
      }
    }
  } else {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc(128, RG_NC, 0);
      const uint32_t d_main = tmem, d_corr = tmem + RG_NC;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % RG_STAGES;
        const uint32_t ph = (kb / RG_STAGES) & 1;
        tc::mbar_wait(&a_full[s], ph);
        tc::mbar_wait(&w_full[s], ph);
        tc::fence_after_sync();
        const uint32_t abase = tc::smem_u32(smem + size_t(s) * stage_bytes);
        const uint32_t wbase = abase + a_pad;
        for (int tap = 0; tap < taps; ++tap) {
#pragma unroll
          for (int ks = 0; ks < RG_KB / 16; ++ks) {
            const uint32_t a_hi = abase + uint32_t((ks * 2) * RA + tap) * 16u;
            const uint32_t a_lo = abase + uint32_t((RG_KB / 8 + ks * 2) * RA + tap) * 16u;
            const uint32_t w_hi = wbase + uint32_t(((tap * 2 + 0) * (RG_KB / 8) + ks * 2) * RG_NC) * 16u;
            const uint32_t w_lo = wbase + uint32_t(((tap * 2 + 1) * (RG_KB / 8) + ks * 2) * RG_NC) * 16u;
            const uint64_t dah = tc::make_desc(a_hi, uint32_t(RA) * 16u, 128u), dal = tc::make_desc(a_lo, uint32_t(RA) * 16u, 128u);
            const uint64_t dwh = tc::make_desc(w_hi, RG_NC * 16u, 128u), dwl = tc::make_desc(w_lo, RG_NC * 16u, 128u);
            const uint32_t first = (kb | tap | ks) ? 1u : 0u;
            tc::mma_f16_ss(d_main, dah, dwh, idesc, first);
            tc::mma_f16_ss(d_corr, dah, dwl, idesc, first);
            tc::mma_f16_ss(d_corr, dal, dwh, idesc, 1u);
          }
        }
        tc::mma_commit(&empty_bar[s]);
      }
      tc::mma_commit(&acc_full);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<2 * RG_NC>(tmem);
}

bool rowgemm_tc_supported(int K, int taps) { return K % RG_KB == 0 && K >= RG_KB && (taps == 1 || taps == 3); }

// Output columns per CTA for a layer (fixed at voice-load time: the weight packing depends on it).
int rowgemm_tc_nc(int N, int taps) {
  const char* e = getenv("M3B200_ROWGEMM_NC");  // read when a voice is packed
  const int want = e && atoi(e) == 128 ? 128 : 64;
  return (want == 128 && taps == 1 && N > 64) ? 128 : 64;
}

size_t rowgemm_tc_weight_elems(int K, int N, int taps, int nc) {
  const int chunks = (N + nc - 1) / nc;
  return size_t(chunks) * (K / RG_KB) * taps * 2 * RG_KB * nc;
}

template <int NC, int ST>
static void launch_rowgemm_inst(const RowGemmTcParams& p, size_t stage_bytes, cudaStream_t st) {
  ensure_max_dynamic_smem(reinterpret_cast<const void*>(rowgemm_tc_kernel<NC, ST>));
  dim3 grid((p.vrows + 127) / 128, (p.N + NC - 1) / NC);
  rowgemm_tc_kernel<NC, ST><<<grid, RG_THREADS, ST * stage_bytes, st>>>(p);
}

void launch_rowgemm_tc(const RowGemmTcParams& p_in, cudaStream_t st) {
  RowGemmTcParams p = p_in;
  if (p.vrows <= 0) return;
  p.wide = (wide_io_enabled() && p.in_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.in) & 31u) == 0) ? 1 : 0;
  if (p.nc != 64 && p.nc != 128) throw std::runtime_error("rowgemm_tc: unsupported column chunk");
  const int RA = (128 + p.taps - 1) | 1;
  const size_t a_pad = (size_t(2) * (RG_KB / 8) * RA * 16 + 127) & ~size_t(127);
  const size_t stage = a_pad + size_t(p.taps) * 2 * RG_KB * p.nc * 2;
  const char* es = getenv("M3B200_ROWGEMM_STAGES");
  const int want_stages = es ? atoi(es) : 2;
  const bool deep = want_stages >= 3 && 3 * stage <= 100 * 1024;  // a third stage only while two CTAs still share an SM
  if (p.nc == 64) {
    if (deep) launch_rowgemm_inst<64, 3>(p, stage, st);
    else launch_rowgemm_inst<64, 2>(p, stage, st);
  } else {
    if (deep) launch_rowgemm_inst<128, 3>(p, stage, st);
    else launch_rowgemm_inst<128, 2>(p, stage, st);
  }
  post_launch("rowgemm_tc_kernel", st);
}

}  // namespace m3

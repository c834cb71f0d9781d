// sm_100a building blocks: tcgen05.mma / TMEM / mbarrier wrappers (inline PTX) and the
// "interleaved" shared-memory operand layout used by every tensor-core kernel here.
//
// Operand layout (bf16/fp16, K-major, SWIZZLE_NONE): [K/8][rows][8 elements].
// A core matrix is 8 rows x 16 B; with SBO = 128 B the 8-row groups are contiguous, so row r
// of K-chunk c sits at  c*LBO + r*16 B  (LBO = rows*16 B) -- fully linear in r.  A Conv1d
// tap is then just the same tile with the descriptor start address advanced by
// (tap_shift * 16 B): no im2col, no re-staging, any dilation.  (Canonical layout
// ((8,n),2):((1,SBO),LBO) in 16-byte units.)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace m3 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// ---- shared-memory matrix descriptor (SWIZZLE_NONE, version 1) -------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;  // descriptor version (Blackwell)
  return d;
}

// ---- instruction descriptor: kind::f16, fp32 accumulate, A and B K-major ----------------
// ab_format: 0 = fp16, 1 = bf16
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int ab_format) {
  return (1u << 4) | (uint32_t(ab_format) << 7) | (uint32_t(ab_format) << 10) | (uint32_t(N >> 3) << 17) |
         (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// commit all prior tcgen05.mma of this thread to an mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- TMEM allocation (one full warp) ------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(COLS) : "memory");
}

// ---- TMEM <-> registers: 32 lanes x 32-bit, N consecutive columns per thread ----------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
      "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
      "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
// G = 8 or 16 consecutive columns
template <int G>
__device__ __forceinline__ void tmem_ldg(uint32_t taddr, float* v) {
  if constexpr (G == 16) tmem_ld16(taddr, v);
  else tmem_ld8(taddr, v);
}
template <int G>
__device__ __forceinline__ void tmem_stg(uint32_t taddr, const float* v) {
  if constexpr (G == 16) tmem_st16(taddr, v);
  else tmem_st8(taddr, v);
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
// Waits for the phase with the given parity.  A waiter may lag its barrier by at most ONE phase
// (never leave two un-awaited arrivals on a barrier).  Watchdog: a protocol bug traps after ~2 s
// instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && (spins & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}

// Whole-warp wait with ONE polling lane.  Tried in dec_planes_kernel because ncu (r02s) attributed 14 % of the shared-memory
// data path to lsu_wavefronts_mem_shared_op_ld with 24 M bank conflicts while 16 epilogue warps polled with all lanes;
// measured SLOWER than all-lane polling (dec_last 3.00 vs 2.89 ms, r02u): kept as an option only.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31u) == 0u) mbar_wait(bar, parity);
  __syncwarp();
}

// single non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// 1-D bulk copy global -> shared (TMA engine), completion signalled on an mbarrier (complete_tx)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// WaveNet gate tanh(a) * sigmoid(g) on the SFU: with t = 2^(-2|a| log2 e) and u = 2^(-g log2 e),
//     tanh(a) * sigmoid(g) = sign(a) * (1 - t) / ((1 + t) * (1 + u))
// -- two ex2.approx + one rcp.approx (a few ulp each) instead of tanhf + expf + an IEEE division (~60 instructions per
// element; the ncu source page of flow_tc_kernel put 35 % of the epilogue warps' samples on those two lines).
// t <= 1 always; u = +inf (g << 0) gives rcp(inf) = 0, the correct limit.
__device__ __forceinline__ float gate_tanh_sigmoid(float a, float g) {
  float t, u, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-2.885390081777927f * fabsf(a)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(-1.4426950408889634f * g));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((1.f + t) * (1.f + u)));
  return copysignf((1.f - t) * r, a);
}

// ---- operand element types ------------------------------------------------------------------------
template <int FMT>
struct Elem;
template <>
struct Elem<1> {  // bf16
  using T = __nv_bfloat16;
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __host__ float round(float x) { return __bfloat162float(__float2bfloat16(x)); }
  // leaky-relu on a packed pair: max(x, slope*x)  (slope < 1)
  __device__ static __forceinline__ uint32_t lrelu2(uint32_t pk, float slope) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&pk);
    const __nv_bfloat162 r = __hmax2(h, __hmul2(h, __float2bfloat162_rn(slope)));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
};
template <>
struct Elem<0> {  // fp16
  using T = __half;
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __host__ float round(float x) { return __half2float(__float2half(x)); }
  __device__ static __forceinline__ uint32_t lrelu2(uint32_t pk, float slope) {
    const __half2 h = *reinterpret_cast<const __half2*>(&pk);
    const __half2 r = __hmax2(h, __hmul2(h, __float2half2_rn(slope)));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
};

// 256-bit global accesses: one full 32-byte sector per lane and instruction (the row-per-lane pattern of the
// epilogue warps touches 32 different lines per instruction; L1 handles one line per cycle)
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

}  // namespace tc
}  // namespace m3

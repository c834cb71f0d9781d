// Relative-position multi-head attention of the text encoder for SHORT sequences, fp32 FFMA
// (attentions.MultiHeadAttention [EXT], window W, relative embeddings shared by heads; SURVEY.md
// Appendix A.1; reference graph: generator.onnx run at mimic3_tts/voice.py:230).
//
// Sentences are tens to a few hundred ids, so one CTA keeps a whole (utterance, head) in shared memory:
// all keys and values, a block of QB queries and their full score rows.  No online softmax, no key tiles,
// no __syncthreads after the load: every warp owns a set of query rows from the scores to the output.
//   * the relative-key term q_i.Ek[r] is computed as 2W+1 EXTRA KEY ROWS appended to K, the relative-value
//     term p[i,i+r].Ev[r] as 2W+1 EXTRA VALUE ROWS appended to V: both are plain columns of the same
//     register-tiled products as the rest (RB query rows x 3 keys per lane; RB rows x 3 channels per lane);
//   * products are accumulated in ascending channel / key order with one fp32 accumulator per output, like
//     the generic kernel (attention_kernel, kernels_simt.cu), which remains the path for long utterances.
// The text side decides the integer durations, so everything here stays fp32 (SURVEY.md hard part 2).
#include <cstdlib>
#include <stdexcept>

#include "kernels.h"

namespace m3 {

namespace {
constexpr int AT_THREADS = 256, AT_WARPS = 8, AT_MAXRB = 8;

__device__ __forceinline__ float at_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float at_warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct AttnGeo {
  int P;      // row pitch of Q / K / V in floats (dk + pad, P/4 odd: conflict-free 128-bit row-per-lane loads)
  int SP;     // row pitch of the score rows
  int kcap;   // key rows incl. the relative rows
  int vcap;   // value rows incl. the relative rows, rounded up to 4
  size_t off_k, off_v, off_s, off_l, floats;
};
__host__ __device__ inline AttnGeo attn_geo(int dk, int nrel, int tcap, int qb) {
  AttnGeo g;
  g.P = dk + (((dk / 4) & 1) ? 8 : 4);
  g.kcap = tcap + nrel;
  g.vcap = (g.kcap + 3) & ~3;
  g.SP = g.vcap;
  size_t o = size_t(qb) * g.P;
  g.off_k = o;
  o += size_t(g.kcap) * g.P;
  g.off_v = o;
  o += size_t(g.vcap) * g.P;
  g.off_s = o;
  o += size_t(qb) * g.SP;
  g.off_l = o;
  o += size_t(qb);
  g.floats = o;
  return g;
}

// scores of RB query rows [i0, i0+RB) against all Tk extended keys: S[i][j] = q_i . k_j
template <int RB>
__device__ __forceinline__ void at_scores(const float* Qs, const float* Ks, float* S, int P, int SP, int dk4, int i0,
                                          int nq, int Tk, int lane) {
  int qrow[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) qrow[r] = min(i0 + r, nq - 1) * P;
  for (int j0 = 0; j0 < Tk; j0 += 96) {
    int krow[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) krow[c] = min(j0 + lane + 32 * c, Tk - 1) * P;
    float acc[RB][3];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[r][c] = 0.f;
#pragma unroll 2
    for (int d4 = 0; d4 < dk4; ++d4) {
      float4 k[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) k[c] = *reinterpret_cast<const float4*>(Ks + krow[c] + 4 * d4);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float4 q = *reinterpret_cast<const float4*>(Qs + qrow[r] + 4 * d4);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float a = acc[r][c];
          a = fmaf(q.x, k[c].x, a);
          a = fmaf(q.y, k[c].y, a);
          a = fmaf(q.z, k[c].z, a);
          a = fmaf(q.w, k[c].w, a);
          acc[r][c] = a;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int j = j0 + lane + 32 * c;
        if (i0 + r < nq && j < Tk) S[(i0 + r) * SP + j] = acc[r][c];
      }
  }
}

// out_i = (sum_j p[i][j] v_j) / l_i over the extended value rows, RB query rows at a time; lane = channel
template <int RB>
__device__ __forceinline__ void at_output(const float* S, const float* Vs, const float* linv, float* out, int P, int SP,
                                          int dk, int i0, int nq, int Tv, long long out_row0, int H, int lane) {
  int srow[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) srow[r] = min(i0 + r, nq - 1) * SP;
  int dcol[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) dcol[c] = min(lane + 32 * c, dk - 1);
  float acc[RB][3];
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[r][c] = 0.f;
  for (int j = 0; j < Tv; j += 4) {
    float v[4][3];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int c = 0; c < 3; ++c) v[jj][c] = Vs[(j + jj) * P + dcol[c]];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const float4 p = *reinterpret_cast<const float4*>(S + srow[r] + j);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float a = acc[r][c];
        a = fmaf(p.x, v[0][c], a);
        a = fmaf(p.y, v[1][c], a);
        a = fmaf(p.z, v[2][c], a);
        a = fmaf(p.w, v[3][c], a);
        acc[r][c] = a;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    if (i0 + r >= nq) continue;
    const float l = linv[i0 + r];
    float* dst = out + (out_row0 + i0 + r) * (long long)H;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int d = lane + 32 * c;
      if (d < dk) dst[d] = acc[r][c] / l;
    }
  }
}

__global__ void __launch_bounds__(AT_THREADS, 2)
    attention_short_kernel(const float* __restrict__ qkv, const float* __restrict__ Ek, const float* __restrict__ Ev,
                           float* __restrict__ out, int H, int dk, int W, const int* __restrict__ seg_off,
                           const int* __restrict__ seg_len, int tcap, int qb) {
  extern __shared__ __align__(16) float at_sm[];
  const int seg = blockIdx.z, h = blockIdx.y;
  const int T = seg_len[seg];
  const int q0 = blockIdx.x * qb;
  if (q0 >= T) return;
  const int nq = min(qb, T - q0);
  const int nrel = 2 * W + 1;
  const AttnGeo g = attn_geo(dk, nrel, tcap, qb);
  float* Qs = at_sm;
  float* Ks = at_sm + g.off_k;
  float* Vs = at_sm + g.off_v;
  float* S = at_sm + g.off_s;
  float* linv = at_sm + g.off_l;
  const int P = g.P, SP = g.SP;
  const int Tk = T + nrel, Tv = (Tk + 3) & ~3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dk4 = dk >> 2, H3 = 3 * H;
  const long long base = seg_off[seg];

  // ---- stage Q (scaled), K|Ek, V|Ev -------------------------------------------------------------
  {
    const float sq = sqrtf(float(dk));
    const float* src = qkv + (base + q0) * (long long)H3 + h * dk;
    for (int idx = tid; idx < nq * dk4; idx += AT_THREADS) {
      const int i = idx / dk4, c4 = idx - i * dk4;
      float4 v = *reinterpret_cast<const float4*>(src + (long long)i * H3 + 4 * c4);
      v.x /= sq;
      v.y /= sq;
      v.z /= sq;
      v.w /= sq;
      *reinterpret_cast<float4*>(Qs + i * P + 4 * c4) = v;
    }
    const float* ksrc = qkv + base * (long long)H3 + H + h * dk;
    for (int idx = tid; idx < T * dk4; idx += AT_THREADS) {
      const int j = idx / dk4, c4 = idx - j * dk4;
      const float* row = ksrc + (long long)j * H3 + 4 * c4;
      *reinterpret_cast<float4*>(Ks + j * P + 4 * c4) = *reinterpret_cast<const float4*>(row);
      *reinterpret_cast<float4*>(Vs + j * P + 4 * c4) = *reinterpret_cast<const float4*>(row + H);
    }
    for (int idx = tid; idx < (Tv - T) * dk4; idx += AT_THREADS) {
      const int r = idx / dk4, c4 = idx - r * dk4;
      float4 ek = make_float4(0.f, 0.f, 0.f, 0.f), ev = ek;
      if (r < nrel) {  // scalar loads: the embedding tables carry no 16-byte alignment guarantee
        const float* a = Ek + r * dk + 4 * c4;
        const float* b = Ev + r * dk + 4 * c4;
        ek = make_float4(a[0], a[1], a[2], a[3]);
        ev = make_float4(b[0], b[1], b[2], b[3]);
      }
      if (r < nrel) *reinterpret_cast<float4*>(Ks + (T + r) * P + 4 * c4) = ek;
      *reinterpret_cast<float4*>(Vs + (T + r) * P + 4 * c4) = ev;  // rows [Tk, Tv) are zero padding
    }
  }
  __syncthreads();

  // ---- row blocks: rounds x 8 warps x RB rows cover the nq query rows --------------------------
  const int rounds = (nq + AT_WARPS * AT_MAXRB - 1) / (AT_WARPS * AT_MAXRB);
  const int rb = (nq + AT_WARPS * rounds - 1) / (AT_WARPS * rounds);  // 1..8
  const int nblk = (nq + rb - 1) / rb;
  const long long out_row0 = base + q0;
  float* const outh = out + h * dk;
  for (int blk = warp; blk < nblk; blk += AT_WARPS) {
    const int i0 = blk * rb;
    switch (rb) {
      case 1: at_scores<1>(Qs, Ks, S, P, SP, dk4, i0, nq, Tk, lane); break;
      case 2: at_scores<2>(Qs, Ks, S, P, SP, dk4, i0, nq, Tk, lane); break;
      case 3: at_scores<3>(Qs, Ks, S, P, SP, dk4, i0, nq, Tk, lane); break;
      case 4: at_scores<4>(Qs, Ks, S, P, SP, dk4, i0, nq, Tk, lane); break;
      case 5: at_scores<5>(Qs, Ks, S, P, SP, dk4, i0, nq, Tk, lane); break;
      case 6: at_scores<6>(Qs, Ks, S, P, SP, dk4, i0, nq, Tk, lane); break;
      case 7: at_scores<7>(Qs, Ks, S, P, SP, dk4, i0, nq, Tk, lane); break;
      default: at_scores<8>(Qs, Ks, S, P, SP, dk4, i0, nq, Tk, lane); break;
    }
    __syncwarp();
    // softmax over the T real keys; scores[i][j] += q_i.Ek[j-i+W] for |j-i| <= W (columns T.. of the row)
    const int iend = min(i0 + rb, nq);
    for (int i = i0; i < iend; ++i) {
      float* row = S + i * SP;
      const int gi = q0 + i;
      float mx = -INFINITY;
      for (int j = lane; j < T; j += 32) {
        float v = row[j];
        const int rel = j - gi;
        if (rel >= -W && rel <= W) v += row[T + rel + W];
        row[j] = v;
        mx = fmaxf(mx, v);
      }
      mx = at_warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < T; j += 32) {
        const float e = expf(row[j] - mx);
        row[j] = e;
        sum += e;
      }
      sum = at_warp_sum(sum);
      __syncwarp();
      // columns T.. now become the probabilities of the relative VALUE rows: p[i][i + r - W] (0 outside)
      if (lane < Tv - T) {
        const int j = gi + lane - W;
        row[T + lane] = (lane < nrel && j >= 0 && j < T) ? row[j] : 0.f;
      }
      if (lane == 0) linv[i] = sum;
    }
    __syncwarp();
    switch (rb) {
      case 1: at_output<1>(S, Vs, linv, outh, P, SP, dk, i0, nq, Tv, out_row0, H, lane); break;
      case 2: at_output<2>(S, Vs, linv, outh, P, SP, dk, i0, nq, Tv, out_row0, H, lane); break;
      case 3: at_output<3>(S, Vs, linv, outh, P, SP, dk, i0, nq, Tv, out_row0, H, lane); break;
      case 4: at_output<4>(S, Vs, linv, outh, P, SP, dk, i0, nq, Tv, out_row0, H, lane); break;
      case 5: at_output<5>(S, Vs, linv, outh, P, SP, dk, i0, nq, Tv, out_row0, H, lane); break;
      case 6: at_output<6>(S, Vs, linv, outh, P, SP, dk, i0, nq, Tv, out_row0, H, lane); break;
      case 7: at_output<7>(S, Vs, linv, outh, P, SP, dk, i0, nq, Tv, out_row0, H, lane); break;
      default: at_output<8>(S, Vs, linv, outh, P, SP, dk, i0, nq, Tv, out_row0, H, lane); break;
    }
    __syncwarp();
  }
}

int attn_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}
}  // namespace

// Picks the query-block size; returns false (generic kernel) when the sequence does not fit shared memory or
// the layout assumptions (dk % 4, dk <= 96, 2W+1 <= 32) do not hold.
bool launch_attention_short(const float* qkv, const float* emb_rel_k, const float* emb_rel_v, float* out, int H,
                            int n_heads, int window, const int* seg_off, const int* seg_len, int n_seg, int max_len,
                            cudaStream_t st) {
  const int forced_v1 = attn_env("M3B200_ATTN_V1", 0);  // read per launch: tests and A/B runs toggle it
  const int forced_nb = attn_env("M3B200_ATTN_NB", 0);
  if (forced_v1 || max_len <= 0 || H % n_heads) return false;
  const int dk = H / n_heads, nrel = 2 * window + 1;
  if (dk % 4 || dk > 96 || dk < 4 || H % 4 || nrel > 28) return false;
  static const int optin = [] {
    int dev = 0, v = 227 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return v;
  }();
  const size_t limit = size_t(optin) - 2048;
  const size_t two_per_sm = size_t(optin) / 2 - 2048;  // two CTAs per SM (1 KB of reserved shared memory each)
  auto block_rows = [&](int nb) { return (((max_len + nb - 1) / nb) + 7) & ~7; };
  auto bytes_for = [&](int qb) { return attn_geo(dk, nrel, max_len, qb).floats * sizeof(float); };
  int best_qb = 0;
  if (forced_nb > 0) {
    if (bytes_for(block_rows(forced_nb)) <= limit) best_qb = block_rows(forced_nb);
  } else {
    for (int nb = 1; nb <= 4 && !best_qb; ++nb)   // prefer a shape that lets two CTAs share an SM
      if (bytes_for(block_rows(nb)) <= two_per_sm) best_qb = block_rows(nb);
    for (int nb = 1; nb <= 16 && !best_qb; ++nb)
      if (bytes_for(block_rows(nb)) <= limit) best_qb = block_rows(nb);
  }
  if (!best_qb) return false;
  const size_t smem = attn_geo(dk, nrel, max_len, best_qb).floats * sizeof(float);
  ensure_max_dynamic_smem(reinterpret_cast<const void*>(attention_short_kernel));
  dim3 grid((max_len + best_qb - 1) / best_qb, n_heads, n_seg);
  attention_short_kernel<<<grid, AT_THREADS, smem, st>>>(qkv, emb_rel_k, emb_rel_v, out, H, dk, window, seg_off, seg_len,
                                                         max_len, best_qb);
  post_launch("attention_short_kernel", st);
  return true;
}

}  // namespace m3

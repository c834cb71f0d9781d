// Engine internals shared by engine.cu and m3_api.cc.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "voice_model.h"

namespace m3 {

struct EngineError : std::runtime_error {
  int code;
  EngineError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define M3_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      throw ::m3::EngineError(4, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr); \
  } while (0)

struct TcConvW {  // 16-bit tensor-core packing [chunk][tap][K/8][NC][8] (kernels_tc_conv.cu)
  bool ok = false;
  unsigned long long woff = 0;  // element offset into DeviceVoice::slab16
  int K = 0, NC = 0, n_chunks = 0, N = 0, taps = 1;
};

struct RowTcW {  // fp16 hi/lo split packing for rowgemm_tc_kernel (kernels_tc_rows.cu)
  bool ok = false;
  unsigned long long woff = 0;
  int nc = 64;  // output columns per CTA the block layout was built for (rowgemm_tc_nc)
};

struct Lin {  // a Conv1d packed as [taps][Cin][ldw] (+ bias[ldw])
  const float* w = nullptr;
  const float* b = nullptr;
  int cin = 0, cout = 0, taps = 1;
  TcConvW tc;
  RowTcW rtc;
};

struct DDSW {
  const float* sep_w[3];
  const float* sep_b[3];
  Lin c1x1[3];
  const float *n1g[3], *n1b[3], *n2g[3], *n2b[3];
};

struct EncLayerW {
  Lin qkv, o, ffn1, ffn2;
  const float *ek, *ev, *g1, *b1, *g2, *b2;
};

struct ConvFlowW {
  const float *pre_w, *pre_b;
  DDSW dds;
  Lin proj;
};

struct FlowTcW {  // fused coupling-layer packing (kernels_tc_flow.cu)
  bool ok = false;
  unsigned long long woff = 0;
  const float *in_bias = nullptr, *cum_bias = nullptr, *skip_bias = nullptr, *post_bias = nullptr;
  int x0_coff = 0, x1_coff = 0;
  // second-generation kernel (kernels_tc_flow2.cu): 36 864-byte blocks in its schedule order, post folded into the skips
  bool ok2 = false;
  unsigned long long woff2 = 0;
  const float* m_bias = nullptr;  // [half]: post bias + W_post . (sum of the skip biases), flip-permuted
};

struct CouplingW {
  Lin pre, post;
  std::vector<Lin> in, rs;
  int cond_off = -1;
  FlowTcW ftc;
};

struct UpW {
  const float* w = nullptr;  // [phase u][ntaps][Cin][Cout]
  const float* b = nullptr;
  int cin = 0, cout = 0, k = 0, u = 0, ntaps = 0, pad = 0;
  TcConvW tc;
};

struct ResBlockW {
  int k = 0;
  std::vector<int> dil;
  std::vector<Lin> c1, c2;  // c2 empty for resblock "2"
};

struct MrfStageW {  // tensor-core packing of one MRF stage (kernels_tc.cu)
  bool ok = false;
  unsigned long long woff[4][2] = {};
  const float* late_bias = nullptr;
  int H = 0, HX = 0, HY = 0, nk = 0, nd = 0;
};

constexpr unsigned kDecPostPlanesBytes = 10 * 4 * 16 * 8 * 2;  // regrouped conv_post of dec_planes_kernel (kernels.h)

struct DecLastW {  // fused last generator stage (kernels_tc_dec.cu)
  bool ok = false;
  unsigned long long up_woff = 0, post_woff = 0;
  // persistent kernel (kernels_tc_dec2.cu): one contiguous blob, byte offsets relative to blob_off
  bool fused_ok = false;
  bool planes_ok = false;  // phase-major kernel (kernels_tc_dec3.cu) can run this stage
  unsigned long long blob_off = 0;  // element offset in slab16
  unsigned blob_bytes = 0, f_up = 0, f_post = 0, f_c1[3] = {}, f_c2[3] = {};
  unsigned f_postp = 0;  // conv_post regrouped by (row shift, input plane) for the phase-major kernel; lies past blob_bytes
  int HYb[3] = {};
};

struct DeviceVoice {
  VoiceConfig cfg;
  DecLastW dec_last;
  int device = 0;
  float* slab = nullptr;  // all weights, one allocation
  uint16_t* slab16 = nullptr;  // 16-bit tensor-core operands
  int tc_fmt = 1;              // 0 fp16, 1 bf16
  bool use_tc = true;
  bool use_rows_tc = true;  // text-side GEMMs on tensor cores (fp16 x 3 split); M3B200_TEXT_SIMT=1 disables
  std::vector<MrfStageW> mrf;
  size_t slab_floats = 0;
  int64_t n_params = 0;
  bool has_emb_g = false;
  int window = 4;

  const float* emb = nullptr;
  std::vector<EncLayerW> enc;
  Lin enc_proj;
  // duration predictor
  bool use_sdp = true;
  int dp_ch = 0;
  Lin dp_pre, dp_proj;
  DDSW dp_dds;
  float ea_m[2] = {0, 0}, ea_logs[2] = {0, 0};
  std::vector<ConvFlowW> cflows;  // in application order (flows.7, .5, .3)
  Lin dpp_c1, dpp_c2, dpp_proj;   // plain DurationPredictor
  const float *dpp_g1 = nullptr, *dpp_b1 = nullptr, *dpp_g2 = nullptr, *dpp_b2 = nullptr;
  int dp_cond_off = -1;
  // flow
  std::vector<CouplingW> couplings;  // application order (flows.6, .4, .2, .0)
  int flow_hidden = 0, flow_layers = 0, flow_kernel = 5;
  // decoder
  Lin dec_pre;
  int dec_cond_off = -1;
  std::vector<UpW> ups;
  std::vector<ResBlockW> rbs;
  const float* post_w = nullptr;
  int post_k = 7, post_c = 0;
  // speaker conditioning: all cond layers as one GEMM (G -> n_cond)
  const float* emb_g = nullptr;
  Lin cond_all;
  int n_cond = 0;

  ~DeviceVoice();
};

// Host-only result of binding + packing a voice: the DeviceVoice with every `const float*` slot still unresolved
// (`fix` lists slot -> element offset into `f32`), the fp32 slab and the 16-bit tensor-core operand slab.
struct PackedVoice {
  std::unique_ptr<DeviceVoice> dv;
  std::vector<float> f32;
  std::vector<uint16_t> h16;
  std::vector<std::pair<const float**, size_t>> fix;
  std::string pack_flags;
};
PackedVoice pack_voice(const HostVoice& hv);                               // no CUDA call
std::unique_ptr<DeviceVoice> upload_voice(PackedVoice&& pv, int device);   // M3_ERR_NOGPU without an sm_100 device
void upload_slabs(DeviceVoice& dv, int device, const float* f32, size_t n_f32, const uint16_t* h16, size_t n_h16);
std::string pack_flags_string();
std::unique_ptr<DeviceVoice> build_device_voice(const HostVoice& hv, int device);

struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  void reset() { used = 0; }
  void reserve(size_t bytes);  // grows (cudaFree + cudaMalloc) if needed; contents lost
  template <typename T>
  T* alloc(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    if (used + bytes > cap) throw EngineError(4, "internal: workspace arena overflow");
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
  ~Arena();
};

struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes);
  ~PinnedBuf();
};

struct Context {  // per concurrent call: stream + workspaces (SURVEY.md §8b threading row)
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  Arena a1, a2;
  PinnedBuf h_in, h_meta, h_pcm, h_audio, h_opts;
  std::vector<cudaEvent_t> marks;  // stage-timing events (M3_FLAG_STAGE_TIMING), created lazily
  explicit Context(int dev);
  ~Context();
};

struct DebugTensor {
  std::vector<float> data;
  int64_t rows = 0, cols = 0;
};

struct Result {
  int batch = 0;
  std::vector<int64_t> sample_off, frames;
  std::vector<float> peaks;
  const int16_t* pcm = nullptr;   // pinned, owned by ctx
  const uint8_t* stream = nullptr;  // pinned: [44-byte WAV header if asked] + pcm (silences included)
  int64_t stream_bytes = 0;
  const float* audio = nullptr;   // pinned, owned by ctx
  const void* d_pcm = nullptr;    // device, owned by ctx
  double device_ms = 0;
  int64_t launches = 0;
  std::map<std::string, DebugTensor> debug;
  struct Voice* owner = nullptr;
  Context* ctx = nullptr;  // leased until the result is freed
};

struct Voice {
  std::unique_ptr<DeviceVoice> dv;
  std::mutex mu;
  std::vector<Context*> idle;
  std::vector<std::unique_ptr<Context>> all;
  Context* acquire();
  void release(Context* c);
};

// Per-utterance settings and the PCM post chain (m3_infer_opts in include/m3b200.h); all optional.
struct InferOpts {
  const float* row_scales = nullptr;     // [batch][3] {noise_scale, length_scale, noise_w}
  const double* volume = nullptr;        // [batch] audioop.mul factor
  const int64_t* lead_silence = nullptr;   // [batch] zero samples before utterance b
  const int64_t* trail_silence = nullptr;  // [batch] zero samples after utterance b
  int sample_rate_override = 0;
  bool wav_header = false;
  bool post_chain() const { return volume || lead_silence || trail_silence || wav_header; }
};

// 44-byte RIFF/WAVE header of 16-bit mono PCM (what Python's wave module writes for
// opentts_abc.AudioResult.to_wav_bytes, opentts_abc/__init__.py:117-127)
void write_wav_header(uint8_t* out44, int sample_rate, int64_t n_samples);

Result* run_inference(Voice& v, const int64_t* ids, const int64_t* lengths, int batch, int t_stride,
                      const float* scales, const int64_t* sid, uint64_t seed, uint32_t flags,
                      const InferOpts* opts = nullptr);

}  // namespace m3

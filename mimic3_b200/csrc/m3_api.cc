// C ABI of libm3b200.so (declared in include/m3b200.h).
#include <cuda_runtime.h>

#include <cstddef>
#include <cstring>
#include <string>

#include "../../include/m3b200.h"
#include "engine.h"

struct m3_voice {
  m3::Voice v;
  int64_t n_params = 0;
};
struct m3_result {
  m3::Result* r = nullptr;
};

namespace {
thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

template <typename F>
int guarded(F&& f) {
  try {
    g_err.clear();
    f();
    return M3_OK;
  } catch (const m3::EngineError& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(M3_ERR_CUDA, "out of host memory");
  } catch (const std::exception& e) {
    std::string m = e.what();
    int code = M3_ERR_MODEL;
    if (m.rfind("cannot open", 0) == 0 || m.rfind("short read", 0) == 0) code = M3_ERR_IO;
    return fail(code, m);
  }
}

int count_sm100() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}
}  // namespace

extern "C" {

const char* m3_version(void) { return "m3b200 0.1.0 (sm_100a)"; }
const char* m3_last_error(void) { return g_err.c_str(); }
int32_t m3_device_count(void) { return count_sm100(); }

int32_t m3_voice_load(const char* path, int32_t device, m3_voice** out) {
  if (!out || !path) return fail(M3_ERR_INVALID, "m3_voice_load: NULL argument");
  *out = nullptr;
  return guarded([&] {
    // Parse + bind + pack on the host first: file / model errors are reported even on a box
    // without a GPU; the device is only touched at upload time (M3_ERR_NOGPU if absent).
    m3::HostVoice hv = m3::load_host_voice(path);
    std::unique_ptr<m3_voice> mv(new m3_voice());
    mv->v.dv = m3::build_device_voice(hv, device);
    mv->n_params = mv->v.dv->n_params;
    *out = mv.release();
  });
}

void m3_voice_free(m3_voice* voice) { delete voice; }

int32_t m3_voice_get_info(const m3_voice* voice, m3_voice_info* info) {
  if (!voice || !info) return fail(M3_ERR_INVALID, "m3_voice_get_info: NULL argument");
  const m3::DeviceVoice& dv = *voice->v.dv;
  memset(info, 0, sizeof *info);
  info->num_symbols = dv.cfg.num_symbols;
  info->n_speakers = dv.cfg.n_speakers;
  info->is_multispeaker = dv.cfg.multispeaker ? 1 : 0;
  info->has_speaker_embedding = dv.has_emb_g ? 1 : 0;
  info->sample_rate = dv.cfg.sample_rate;
  info->hop_length = dv.cfg.hop();
  info->hidden_channels = dv.cfg.hidden;
  info->inter_channels = dv.cfg.inter;
  info->noise_scale = dv.cfg.noise_scale;
  info->length_scale = dv.cfg.length_scale;
  info->noise_w = dv.cfg.noise_w;
  info->n_params = voice->n_params;
  info->device = dv.device;
  return M3_OK;
}

int32_t m3_infer(m3_voice* voice, const int64_t* ids, const int64_t* lengths, int32_t batch, int32_t t_stride,
                 const float* scales, const int64_t* sid, uint64_t seed, uint32_t flags, m3_result** out) {
  if (!voice || !out) return fail(M3_ERR_INVALID, "m3_infer: NULL argument");
  *out = nullptr;
  return guarded([&] {
    m3::Result* r = m3::run_inference(voice->v, ids, lengths, batch, t_stride, scales, sid, seed, flags);
    m3_result* mr = new m3_result();
    mr->r = r;
    *out = mr;
  });
}

int32_t m3_infer_ex(m3_voice* voice, const int64_t* ids, const int64_t* lengths, int32_t batch, int32_t t_stride,
                    const float* scales, const int64_t* sid, const m3_infer_opts* opts, m3_result** out) {
  if (!voice || !out) return fail(M3_ERR_INVALID, "m3_infer_ex: NULL argument");
  *out = nullptr;
  if (opts && opts->struct_size < offsetof(m3_infer_opts, reserved))
    return fail(M3_ERR_INVALID, "m3_infer_ex: opts->struct_size does not describe an m3_infer_opts");
  return guarded([&] {
    m3::InferOpts o;
    uint64_t seed = 0;
    uint32_t flags = 0;
    if (opts) {
      o.row_scales = opts->row_scales;
      o.volume = opts->volume;
      o.lead_silence = opts->lead_silence;
      o.trail_silence = opts->trail_silence;
      o.wav_header = opts->wav_header != 0;
      seed = opts->seed;
      flags = opts->flags;
    }
    m3::Result* r = m3::run_inference(voice->v, ids, lengths, batch, t_stride, scales, sid, seed, flags, &o);
    m3_result* mr = new m3_result();
    mr->r = r;
    *out = mr;
  });
}

const uint8_t* m3_result_stream(const m3_result* r, int64_t* n_bytes) {
  if (n_bytes) *n_bytes = r ? r->r->stream_bytes : 0;
  return r ? r->r->stream : nullptr;
}

int32_t m3_wav_header(int32_t sample_rate, int64_t n_samples, uint8_t out[44]) {
  if (!out || sample_rate <= 0 || n_samples < 0 || n_samples > ((int64_t(1) << 31) - 64))
    return fail(M3_ERR_INVALID, "m3_wav_header: bad argument");
  m3::write_wav_header(out, sample_rate, n_samples);
  return M3_OK;
}

int32_t m3_result_batch(const m3_result* r) { return r ? r->r->batch : 0; }
const int64_t* m3_result_sample_offsets(const m3_result* r) { return r ? r->r->sample_off.data() : nullptr; }
const int64_t* m3_result_num_frames(const m3_result* r) { return r ? r->r->frames.data() : nullptr; }
const int16_t* m3_result_pcm(const m3_result* r) { return r ? r->r->pcm : nullptr; }
const float* m3_result_audio(const m3_result* r) { return r ? r->r->audio : nullptr; }
const float* m3_result_peaks(const m3_result* r) { return r ? r->r->peaks.data() : nullptr; }
const void* m3_result_device_pcm(const m3_result* r) { return r ? r->r->d_pcm : nullptr; }
double m3_result_device_ms(const m3_result* r) { return r ? r->r->device_ms : 0.0; }
int64_t m3_result_kernel_launches(const m3_result* r) { return r ? r->r->launches : 0; }

int32_t m3_result_tensor(const m3_result* r, const char* name, const float** data, int64_t* rows, int64_t* cols) {
  if (!r || !name || !data || !rows || !cols) return fail(M3_ERR_INVALID, "m3_result_tensor: NULL argument");
  auto it = r->r->debug.find(name);
  if (it == r->r->debug.end()) return fail(M3_ERR_INVALID, std::string("no debug tensor named '") + name + "'");
  *data = it->second.data.data();
  *rows = it->second.rows;
  *cols = it->second.cols;
  return M3_OK;
}

void m3_result_free(m3_result* r) {
  if (!r) return;
  if (r->r) {
    if (r->r->owner && r->r->ctx) r->r->owner->release(r->r->ctx);
    delete r->r;
  }
  delete r;
}

}  // extern "C"

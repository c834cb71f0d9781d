// C ABI of libm3b200.so (declared in include/m3b200.h).
#include <cuda_runtime.h>

#include <cstddef>
#include <cstring>
#include <string>

#include "../../include/m3b200.h"
#include "engine.h"
#include "weight_cache.h"

#include <sys/stat.h>

#include <chrono>

struct m3_voice {
  m3::Voice v;
  int64_t n_params = 0;
  m3_load_stats stats{};
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

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

std::string cache_dir_of(const char* arg) {
  if (arg && *arg) return arg;
  const char* e = getenv("M3B200_WEIGHT_CACHE");
  return e ? e : "";
}

// "" or 64 lower-case hex digits (upper case folded); anything else is a caller error
std::string check_hex(const char* s) {
  if (!s || !*s) return "";
  std::string h(s);
  if (h.size() != 64) throw m3::EngineError(M3_ERR_INVALID, "expected_sha256 must be 64 hex digits");
  for (char& c : h) {
    if (c >= 'A' && c <= 'F') c = char(c - 'A' + 'a');
    if (!((c >= '0' && c <= '9') || (c >= 'a' && c <= 'f')))
      throw m3::EngineError(M3_ERR_INVALID, "expected_sha256 must be 64 hex digits");
  }
  return h;
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

namespace m3 {
void set_last_error(const std::string& m) { g_err = m; }
}  // namespace m3

extern "C" {

const char* m3_version(void) { return "m3b200 0.1.0 (sm_100a)"; }
const char* m3_last_error(void) { return g_err.c_str(); }
int32_t m3_device_count(void) { return count_sm100(); }

int32_t m3_voice_load(const char* path, int32_t device, m3_voice** out) {
  return m3_voice_load_ex(path, device, nullptr, out);
}

int32_t m3_voice_load_ex(const char* path, int32_t device, const m3_load_opts* opts, m3_voice** out) {
  if (!out || !path) return fail(M3_ERR_INVALID, "m3_voice_load: NULL argument");
  *out = nullptr;
  if (opts && opts->struct_size < offsetof(m3_load_opts, expected_sha256) + sizeof(const char*))
    return fail(M3_ERR_INVALID, "m3_voice_load_ex: opts->struct_size does not describe an m3_load_opts");
  return guarded([&] {
    const double t_start = now_ms();
    std::unique_ptr<m3_voice> mv(new m3_voice());
    m3_load_stats& st = mv->stats;
    const uint32_t flags = opts ? opts->flags : 0u;
    const std::string dir = cache_dir_of(opts ? opts->cache_dir : nullptr);
    const std::string want = check_hex(opts ? opts->expected_sha256 : nullptr);
    std::string sha;  // digest of generator.onnx, once known
    auto hash_now = [&](const std::string& onnx) {
      const double t = now_ms();
      sha = m3::sha256_file(onnx);
      st.hash_ms += now_ms() - t;
    };
    auto check_manifest = [&] {
      if (!want.empty() && sha != want)
        throw m3::EngineError(M3_ERR_MODEL, "generator.onnx sha256 is " + sha + " but the voice registry lists " + want +
                                                " (mimic3_tts/voices.json; the downloader would fetch it again)");
    };
    m3::CacheKey key;
    std::string file;
    if (!dir.empty()) {
      key = m3::make_cache_key(path);
      file = dir + "/" + key.file_name();
      snprintf(st.cache_file, sizeof st.cache_file, "%s", file.c_str());
      const double t = now_ms();
      std::string why;
      std::unique_ptr<m3::CacheImage> img = m3::read_cache_file(file, &key, &why);
      st.cache_read_ms = now_ms() - t;
      if (img) {
        sha = img->onnx_sha256;
        if (flags & M3_LOAD_VERIFY_SHA256) hash_now(key.onnx_path);
        if (sha != img->onnx_sha256)
          img.reset();  // the file changed under an unchanged size + mtime: convert again
      }
      if (img) {
        check_manifest();
        const double tu = now_ms();
        mv->v.dv = m3::upload_cached(*img, device);
        st.upload_ms = now_ms() - tu;
        st.from_cache = 1;
      }
    }
    if (!mv->v.dv) {
      // Parse + bind + pack on the host first: file / model errors are reported even on a box
      // without a GPU; the device is only touched at upload time (M3_ERR_NOGPU if absent).
      double t = now_ms();
      m3::HostVoice hv = m3::load_host_voice(path);
      st.parse_ms = now_ms() - t;
      if (!want.empty() || !dir.empty()) {
        if (sha.empty()) hash_now(hv.onnx_path);
        check_manifest();
      }
      t = now_ms();
      m3::PackedVoice pv = m3::pack_voice(hv);
      st.pack_ms = now_ms() - t;
      if (!dir.empty() && !(flags & M3_LOAD_NO_CACHE_WRITE)) {
        try {
          mkdir(dir.c_str(), 0755);  // one level, like the reference's voice directories; EEXIST is fine
          m3::write_cache_file(file, key, sha, pv);
          st.cache_written = 1;
        } catch (const std::exception&) {
          st.cache_written = 0;  // an unwritable cache directory must not fail the load
        }
      }
      t = now_ms();
      mv->v.dv = m3::upload_voice(std::move(pv), device);
      st.upload_ms = now_ms() - t;
    }
    snprintf(st.onnx_sha256, sizeof st.onnx_sha256, "%s", sha.c_str());
    mv->n_params = mv->v.dv->n_params;
    st.total_ms = now_ms() - t_start;
    *out = mv.release();
  });
}

int32_t m3_voice_load_stats(const m3_voice* voice, m3_load_stats* stats) {
  if (!voice || !stats) return fail(M3_ERR_INVALID, "m3_voice_load_stats: NULL argument");
  *stats = voice->stats;
  return M3_OK;
}

int32_t m3_weight_cache_build(const char* path, const char* cache_dir, const char* expected_sha256, char* out_file,
                              int32_t out_cap) {
  if (!path || !cache_dir || !*cache_dir) return fail(M3_ERR_INVALID, "m3_weight_cache_build: NULL argument");
  if (out_file && out_cap > 0) out_file[0] = 0;
  return guarded([&] {
    const std::string want = check_hex(expected_sha256);
    m3::CacheKey key = m3::make_cache_key(path);
    m3::HostVoice hv = m3::load_host_voice(path);
    const std::string sha = m3::sha256_file(hv.onnx_path);
    if (!want.empty() && sha != want)
      throw m3::EngineError(M3_ERR_MODEL, "generator.onnx sha256 is " + sha + " but the voice registry lists " + want);
    m3::PackedVoice pv = m3::pack_voice(hv);
    mkdir(cache_dir, 0755);
    const std::string file = std::string(cache_dir) + "/" + key.file_name();
    m3::write_cache_file(file, key, sha, pv);
    if (out_file && out_cap > 0) snprintf(out_file, size_t(out_cap), "%s", file.c_str());
  });
}

int32_t m3_weight_cache_check(const char* cache_file, char onnx_sha256_out[65]) {
  if (!cache_file) return fail(M3_ERR_INVALID, "m3_weight_cache_check: NULL argument");
  if (onnx_sha256_out) onnx_sha256_out[0] = 0;
  return guarded([&] {
    std::string why;
    std::unique_ptr<m3::CacheImage> img = m3::read_cache_file(cache_file, nullptr, &why);
    if (!img) throw m3::EngineError(why == "absent" ? M3_ERR_IO : M3_ERR_MODEL, std::string("weight cache blob ") + why);
    if (onnx_sha256_out) snprintf(onnx_sha256_out, 65, "%s", img->onnx_sha256.c_str());
  });
}

int32_t m3_sha256_file(const char* path, char out[65]) {
  if (!path || !out) return fail(M3_ERR_INVALID, "m3_sha256_file: NULL argument");
  out[0] = 0;
  return guarded([&] { snprintf(out, 65, "%s", m3::sha256_file(path).c_str()); });
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

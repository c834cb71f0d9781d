#include "voice_model.h"

#include <sys/stat.h>

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "json_min.h"

namespace m3 {
namespace {

bool is_dir(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

std::string read_file(const std::string& p) {
  std::ifstream f(p, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + p);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

std::vector<int> int_list(const JsonValue* v, const std::vector<int>& dflt) {
  if (!v || v->type != JsonValue::Array) return dflt;
  std::vector<int> out;
  for (auto& e : v->arr) {
    if (e.type != JsonValue::Number || !(e.num >= -2147483648.0 && e.num <= 2147483647.0))
      throw std::runtime_error("config.json: integer list holds a non-integer");
    out.push_back(int(e.num));
  }
  return out;
}

bool ends_with(const std::string& s, const std::string& suf) {
  return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

std::string dims_str(const std::vector<int64_t>& d) {
  std::string s = "(";
  for (size_t i = 0; i < d.size(); ++i) s += (i ? "," : "") + std::to_string(d[i]);
  return s + ")";
}

}  // namespace

const OnnxTensor& HostVoice::need(const std::string& name, const std::vector<int64_t>& dims) const {
  auto it = params.find(name);
  if (it == params.end())
    throw std::runtime_error("generator.onnx: missing initializer '" + name + "' (expected shape " + dims_str(dims) +
                             "); see SURVEY.md Appendix B");
  if (it->second.data_type != 1) throw std::runtime_error("generator.onnx: '" + name + "' is not float32");
  if (!dims.empty() && it->second.dims != dims)
    throw std::runtime_error("generator.onnx: '" + name + "' has shape " + dims_str(it->second.dims) + ", expected " +
                             dims_str(dims) + " from config.json");
  return it->second;
}

HostVoice load_host_voice(const std::string& path) {
  HostVoice hv;
  std::string dir, onnx;
  if (is_dir(path)) {
    dir = path;
    onnx = path + "/generator.onnx";
  } else {
    onnx = path;
    size_t slash = path.find_last_of('/');
    dir = slash == std::string::npos ? "." : path.substr(0, slash);
  }
  hv.onnx_path = onnx;

  // ---- config.json -> VoiceConfig (fields of ModelConfig / AudioConfig / InferenceConfig)
  {
    std::string text = read_file(dir + "/config.json");
    JsonValue root = JsonParser(text).parse();
    VoiceConfig& c = hv.cfg;
    if (const JsonValue* m = root.get("model")) {
      c.num_symbols = m->int_or("num_symbols", 0);
      c.n_speakers = m->int_or("n_speakers", 1);
      c.inter = m->int_or("inter_channels", 192);
      c.hidden = m->int_or("hidden_channels", 192);
      c.filter = m->int_or("filter_channels", 768);
      c.n_heads = m->int_or("n_heads", 2);
      c.n_layers = m->int_or("n_layers", 6);
      c.kernel_size = m->int_or("kernel_size", 3);
      c.resblock = m->string_or("resblock", "1");
      c.rb_kernels = int_list(m->get("resblock_kernel_sizes"), c.rb_kernels);
      if (const JsonValue* d = m->get("resblock_dilation_sizes"); d && d->type == JsonValue::Array) {
        c.rb_dils.clear();
        for (auto& e : d->arr) c.rb_dils.push_back(int_list(&e, {}));
      }
      c.up_rates = int_list(m->get("upsample_rates"), c.up_rates);
      c.up_kernels = int_list(m->get("upsample_kernel_sizes"), c.up_kernels);
      c.up_init = m->int_or("upsample_initial_channel", 512);
      c.gin = m->int_or("gin_channels", 0);
      c.use_sdp = m->number_or("use_sdp", 1) != 0;
    } else {
      throw std::runtime_error("config.json: no \"model\" section");
    }
    if (const JsonValue* a = root.get("audio")) {
      c.sample_rate = a->int_or("sample_rate", 22050);
      c.hop_length = a->int_or("hop_length", 256);
    }
    if (const JsonValue* i = root.get("inference")) {
      c.length_scale = float(i->number_or("length_scale", 1.0));
      c.noise_scale = float(i->number_or("noise_scale", 0.667));
      c.noise_w = float(i->number_or("noise_w", 0.8));
    }
    c.multispeaker = c.n_speakers > 1;
    if (const JsonValue* ds = root.get("datasets"); ds && ds->type == JsonValue::Array)
      for (auto& d : ds->arr)
        if (d.number_or("multispeaker", 0) != 0) c.multispeaker = true;
    if (c.rb_kernels.size() != c.rb_dils.size())
      throw std::runtime_error("config.json: resblock_kernel_sizes / resblock_dilation_sizes length mismatch");
    if (c.up_rates.size() != c.up_kernels.size())
      throw std::runtime_error("config.json: upsample_rates / upsample_kernel_sizes length mismatch");
    // ranges BEFORE any arithmetic on them: a crafted config must come back as M3_ERR_MODEL, never as SIGFPE / bad_alloc
    auto in_range = [](const char* what, long long v, long long lo, long long hi) {
      if (v < lo || v > hi)
        throw std::runtime_error(std::string("config.json: ") + what + " = " + std::to_string(v) + " outside [" +
                                 std::to_string(lo) + ", " + std::to_string(hi) + "]");
    };
    in_range("model.num_symbols", c.num_symbols, 0, 1 << 20);
    in_range("model.n_speakers", c.n_speakers, 0, 1 << 20);
    in_range("model.inter_channels", c.inter, 2, 4096);
    in_range("model.hidden_channels", c.hidden, 1, 4096);
    in_range("model.filter_channels", c.filter, 1, 16384);
    in_range("model.n_heads", c.n_heads, 1, 64);
    in_range("model.n_layers", c.n_layers, 1, 64);
    in_range("model.kernel_size", c.kernel_size, 1, 15);
    in_range("model.upsample_initial_channel", c.up_init, 2, 4096);
    in_range("model.gin_channels", c.gin, 0, 8192);
    in_range("audio.sample_rate", c.sample_rate, 1, 768000);
    in_range("audio.hop_length", c.hop_length, 1, 1 << 16);
    in_range("model.upsample_rates (count)", (long long)c.up_rates.size(), 1, 8);
    in_range("model.resblock_kernel_sizes (count)", (long long)c.rb_kernels.size(), 1, 8);
    long long hop = 1;
    for (size_t i = 0; i < c.up_rates.size(); ++i) {
      in_range("model.upsample_rates[i]", c.up_rates[i], 1, 64);
      in_range("model.upsample_kernel_sizes[i]", c.up_kernels[i], 1, 256);
      hop *= c.up_rates[i];
    }
    in_range("product of model.upsample_rates", hop, 1, 1 << 16);
    if ((c.up_init >> c.up_rates.size()) < 1)
      throw std::runtime_error("config.json: upsample_initial_channel too small for the number of upsample stages");
    for (size_t j = 0; j < c.rb_kernels.size(); ++j) {
      in_range("model.resblock_kernel_sizes[j]", c.rb_kernels[j], 1, 31);
      if (!(c.rb_kernels[j] & 1)) throw std::runtime_error("config.json: resblock kernel sizes must be odd");
      in_range("model.resblock_dilation_sizes[j] (count)", (long long)c.rb_dils[j].size(), 1, 8);
      for (int d : c.rb_dils[j]) in_range("model.resblock_dilation_sizes[j][d]", d, 1, 256);
    }
    if (c.inter & 1) throw std::runtime_error("config.json: inter_channels must be even (coupling layers split it)");
    if (c.hidden % c.n_heads) throw std::runtime_error("config.json: hidden_channels not divisible by n_heads");
    if (!std::isfinite(c.length_scale) || !std::isfinite(c.noise_scale) || !std::isfinite(c.noise_w))
      throw std::runtime_error("config.json: inference scales must be finite");
  }

  // ---- generator.onnx -> named parameters
  OnnxModel om = load_onnx(onnx);
  for (auto& t : om.tensors)
    if (!t.name.empty() && t.data_type == 1) hv.params[t.name] = t;

  // (a) weight_g / weight_v pairs: w = g * v / ||v||, norm over all dims but 0 (PyTorch weight_norm dim=0)
  std::vector<std::string> gnames;
  for (auto& kv : hv.params)
    if (ends_with(kv.first, ".weight_g")) gnames.push_back(kv.first);
  for (auto& gn : gnames) {
    std::string base = gn.substr(0, gn.size() - strlen(".weight_g"));
    auto vit = hv.params.find(base + ".weight_v");
    if (vit == hv.params.end()) continue;
    const OnnxTensor& g = hv.params[gn];
    const OnnxTensor& v = vit->second;
    if (v.dims.empty() || v.dims[0] <= 0 || v.numel() <= 0 || g.numel() != v.dims[0]) continue;
    OnnxTensor w = v;
    w.name = base + ".weight";
    int64_t inner = v.numel() / v.dims[0];
    for (int64_t o = 0; o < v.dims[0]; ++o) {
      double ss = 0;
      for (int64_t i = 0; i < inner; ++i) ss += double(v.f32[o * inner + i]) * v.f32[o * inner + i];
      float norm = float(std::sqrt(ss));
      for (int64_t i = 0; i < inner; ++i) w.f32[o * inner + i] = g.f32[o] * (v.f32[o * inner + i] / norm);
    }
    if (!hv.params.count(w.name)) {
      hv.params[w.name] = std::move(w);
      hv.notes.push_back("fused weight_g/weight_v -> " + base + ".weight");
    }
  }
  // (b) constant-folded conv weights: anonymous W reachable through the Conv node that
  //     consumes the named "<module>.bias".
  for (auto& n : om.nodes) {
    if ((n.op_type != "Conv" && n.op_type != "ConvTranspose") || n.inputs.size() < 3) continue;
    const std::string& w = n.inputs[1];
    const std::string& b = n.inputs[2];
    if (!ends_with(b, ".bias")) continue;
    std::string wname = b.substr(0, b.size() - 5) + ".weight";
    if (hv.params.count(wname)) continue;
    auto it = hv.params.find(w);
    if (it == hv.params.end()) continue;
    OnnxTensor t = it->second;
    t.name = wname;
    hv.params[wname] = std::move(t);
    hv.notes.push_back("bound " + w + " -> " + wname + " via node " + n.name);
  }
  if (hv.cfg.num_symbols <= 0) {
    if (auto* e = hv.maybe("enc_p.emb.weight"); e && e->dims.size() == 2) hv.cfg.num_symbols = int(e->dims[0]);
  }
  return hv;
}

}  // namespace m3

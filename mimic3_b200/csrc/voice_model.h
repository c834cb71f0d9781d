// Host-side voice: config.json + generator.onnx -> named fp32 parameters.
// Mirrors what Mimic3Voice.load_from_directory reads (reference mimic3_tts/voice.py:246-299)
// and the ModelConfig fields (mimic3_tts/config.py:113-139).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "onnx_reader.h"

namespace m3 {

struct VoiceConfig {
  int num_symbols = 0, n_speakers = 1;
  int inter = 192, hidden = 192, filter = 768, n_heads = 2, n_layers = 6, kernel_size = 3;
  std::string resblock = "1";
  std::vector<int> rb_kernels{3, 7, 11};
  std::vector<std::vector<int>> rb_dils{{1, 3, 5}, {1, 3, 5}, {1, 3, 5}};
  std::vector<int> up_rates{8, 8, 2, 2}, up_kernels{16, 16, 4, 4};
  int up_init = 512;
  int gin = 0;
  bool use_sdp = true;
  int sample_rate = 22050, hop_length = 256;
  float noise_scale = 0.667f, length_scale = 1.0f, noise_w = 0.8f;  // config.py:260-262
  bool multispeaker = false;  // TrainingConfig.is_multispeaker, config.py:316-318
  int hop() const {
    int h = 1;
    for (int r : up_rates) h *= r;
    return h;
  }
};

struct HostVoice {
  VoiceConfig cfg;
  std::string onnx_path;
  std::map<std::string, OnnxTensor> params;  // keyed by PyTorch module path, resolved
  std::vector<std::string> notes;            // how non-trivial names were resolved

  const OnnxTensor& need(const std::string& name, const std::vector<int64_t>& dims) const;
  const OnnxTensor* maybe(const std::string& name) const {
    auto it = params.find(name);
    return it == params.end() ? nullptr : &it->second;
  }
};

// `path` is a voice directory (config.json + generator.onnx) or a generator.onnx whose
// directory holds config.json. Throws std::runtime_error with a precise message.
HostVoice load_host_voice(const std::string& path);

}  // namespace m3

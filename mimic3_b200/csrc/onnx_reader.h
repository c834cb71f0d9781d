// Dependency-free ONNX protobuf reader: just enough of onnx.proto to pull a VITS
// voice's weights out of generator.onnx (the file Mimic3Voice._load_model hands to
// onnxruntime.InferenceSession, reference mimic3_tts/voice.py:273,403-405).
// Field numbers: SURVEY.md Appendix D.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace m3 {

struct OnnxTensor {
  std::string name;
  int data_type = 0;  // 1 FLOAT, 7 INT64
  std::vector<int64_t> dims;
  std::vector<float> f32;    // populated for FLOAT
  std::vector<int64_t> i64;  // populated for INT64
  // element count; negative dims or a product beyond 2^40 (no voice comes near) -> -1, which no caller accepts
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : dims) {
      if (d < 0 || (d > 0 && n > (int64_t(1) << 40) / d)) return -1;
      n *= d;
    }
    return n;
  }
};

struct OnnxNode {
  std::string name, op_type;
  std::vector<std::string> inputs, outputs;
  int const_tensor = -1;  // index into OnnxModel::tensors for Constant nodes' "value"
};

struct OnnxModel {
  std::vector<OnnxTensor> tensors;  // initializers + Constant-node tensors
  std::map<std::string, int> by_name;
  std::vector<OnnxNode> nodes;
  std::vector<std::string> inputs, outputs;
  std::string producer;
  int64_t ir_version = 0, opset = 0;

  const OnnxTensor* find(const std::string& n) const {
    auto it = by_name.find(n);
    return it == by_name.end() ? nullptr : &tensors[it->second];
  }
};

// Throws std::runtime_error on malformed input.
OnnxModel load_onnx(const std::string& path);

}  // namespace m3

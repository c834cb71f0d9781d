#include "onnx_reader.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace m3 {
namespace {

struct Span {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
};

uint64_t varint(Span& s) {
  uint64_t r = 0;
  int shift = 0;
  while (true) {
    if (s.p >= s.end) throw std::runtime_error("onnx: truncated varint");
    uint8_t b = *s.p++;
    r |= uint64_t(b & 0x7F) << shift;
    if (!(b & 0x80)) return r;
    shift += 7;
    if (shift > 63) throw std::runtime_error("onnx: varint too long");
  }
}

struct Field {
  uint32_t number = 0;
  uint32_t wire = 0;
  uint64_t value = 0;          // wire 0 / 1 / 5
  Span bytes{nullptr, nullptr};  // wire 2
};

bool next(Span& s, Field& f) {
  if (s.done()) return false;
  uint64_t key = varint(s);
  f.number = uint32_t(key >> 3);
  f.wire = uint32_t(key & 7);
  switch (f.wire) {
    case 0: f.value = varint(s); break;
    case 1:
      if (s.end - s.p < 8) throw std::runtime_error("onnx: truncated fixed64");
      memcpy(&f.value, s.p, 8);
      s.p += 8;
      break;
    case 2: {
      uint64_t n = varint(s);
      if (uint64_t(s.end - s.p) < n) throw std::runtime_error("onnx: truncated bytes field");
      f.bytes = Span{s.p, s.p + n};
      s.p += n;
      break;
    }
    case 5: {
      if (s.end - s.p < 4) throw std::runtime_error("onnx: truncated fixed32");
      uint32_t v;
      memcpy(&v, s.p, 4);
      f.value = v;
      s.p += 4;
      break;
    }
    default: throw std::runtime_error("onnx: unsupported wire type");
  }
  return true;
}

std::string str(const Span& b) { return std::string(reinterpret_cast<const char*>(b.p), b.end - b.p); }

OnnxTensor parse_tensor(Span s) {
  OnnxTensor t;
  Span raw{nullptr, nullptr};
  Field f;
  while (next(s, f)) {
    switch (f.number) {
      case 1:  // dims (packed or not)
        if (f.wire == 0) t.dims.push_back(int64_t(f.value));
        else {
          Span d = f.bytes;
          while (!d.done()) t.dims.push_back(int64_t(varint(d)));
        }
        break;
      case 2: t.data_type = int(f.value); break;
      case 4:  // float_data
        if (f.wire == 5) {
          float v;
          uint32_t u = uint32_t(f.value);
          memcpy(&v, &u, 4);
          t.f32.push_back(v);
        } else {
          size_t n = (f.bytes.end - f.bytes.p) / 4;
          size_t o = t.f32.size();
          t.f32.resize(o + n);
          memcpy(t.f32.data() + o, f.bytes.p, n * 4);
        }
        break;
      case 7:  // int64_data
        if (f.wire == 0) t.i64.push_back(int64_t(f.value));
        else {
          Span d = f.bytes;
          while (!d.done()) t.i64.push_back(int64_t(varint(d)));
        }
        break;
      case 8: t.name = str(f.bytes); break;
      case 9: raw = f.bytes; break;
      case 13: case 14:
        if (f.number == 14 && f.value == 1)
          throw std::runtime_error("onnx: external tensor data is not supported (" + t.name + ")");
        break;
      default: break;
    }
  }
  if (raw.p) {
    size_t nbytes = raw.end - raw.p;
    if (t.data_type == 1) {
      t.f32.resize(nbytes / 4);
      memcpy(t.f32.data(), raw.p, t.f32.size() * 4);
    } else if (t.data_type == 7) {
      t.i64.resize(nbytes / 8);
      memcpy(t.i64.data(), raw.p, t.i64.size() * 8);
    }
  }
  if (t.data_type == 1 && int64_t(t.f32.size()) != t.numel())
    throw std::runtime_error("onnx: tensor '" + t.name + "' size does not match dims");
  return t;
}

}  // namespace

OnnxModel load_onnx(const std::string& path) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) throw std::runtime_error("cannot open " + path);
  fseek(fp, 0, SEEK_END);
  long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  std::vector<uint8_t> buf(sz > 0 ? sz : 0);
  if (sz > 0 && fread(buf.data(), 1, sz, fp) != size_t(sz)) {
    fclose(fp);
    throw std::runtime_error("short read on " + path);
  }
  fclose(fp);

  OnnxModel m;
  Span top{buf.data(), buf.data() + buf.size()};
  Field f;
  bool saw_graph = false;
  while (next(top, f)) {
    if (f.number == 1 && f.wire == 0) m.ir_version = int64_t(f.value);
    else if (f.number == 2 && f.wire == 2) m.producer = str(f.bytes);
    else if (f.number == 8 && f.wire == 2) {
      Span o = f.bytes;
      Field g;
      while (next(o, g))
        if (g.number == 2 && g.wire == 0) m.opset = int64_t(g.value);
    } else if (f.number == 7 && f.wire == 2) {
      saw_graph = true;
      Span g = f.bytes;
      Field gf;
      while (next(g, gf)) {
        if (gf.wire != 2) continue;
        if (gf.number == 5) {
          OnnxTensor t = parse_tensor(gf.bytes);
          m.by_name[t.name] = int(m.tensors.size());
          m.tensors.push_back(std::move(t));
        } else if (gf.number == 1) {
          OnnxNode n;
          Span ns = gf.bytes;
          Field nf;
          while (next(ns, nf)) {
            if (nf.wire != 2) continue;
            if (nf.number == 1) n.inputs.push_back(str(nf.bytes));
            else if (nf.number == 2) n.outputs.push_back(str(nf.bytes));
            else if (nf.number == 3) n.name = str(nf.bytes);
            else if (nf.number == 4) n.op_type = str(nf.bytes);
            else if (nf.number == 5) {  // AttributeProto: keep tensors (Constant.value)
              Span as = nf.bytes;
              Field af;
              while (next(as, af)) {
                if (af.number == 5 && af.wire == 2) {
                  OnnxTensor t = parse_tensor(af.bytes);
                  n.const_tensor = int(m.tensors.size());
                  m.tensors.push_back(std::move(t));
                }
              }
            }
          }
          m.nodes.push_back(std::move(n));
        } else if (gf.number == 11 || gf.number == 12) {
          Span vs = gf.bytes;
          Field vf;
          while (next(vs, vf))
            if (vf.number == 1 && vf.wire == 2) (gf.number == 11 ? m.inputs : m.outputs).push_back(str(vf.bytes));
        }
      }
    }
  }
  if (!saw_graph) throw std::runtime_error("onnx: no graph in " + path);
  // Constant nodes: expose their tensor under the node's output name.
  for (auto& n : m.nodes)
    if (n.op_type == "Constant" && n.const_tensor >= 0 && !n.outputs.empty() && !m.by_name.count(n.outputs[0])) {
      m.tensors[n.const_tensor].name = n.outputs[0];
      m.by_name[n.outputs[0]] = n.const_tensor;
    }
  return m;
}

}  // namespace m3

// Minimal JSON DOM reader for a voice's config.json (reference: TrainingConfig,
// mimic3_tts/config.py:274-327 -- the engine only needs model.*, audio.*, inference.*).
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace m3 {

struct JsonValue {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<JsonValue> arr;
  std::map<std::string, JsonValue> obj;

  const JsonValue* get(const std::string& k) const {
    if (type != Object) return nullptr;
    auto it = obj.find(k);
    return it == obj.end() ? nullptr : &it->second;
  }
  // number_or() narrowed to int with a range check: NaN / huge doubles are an error, not undefined behaviour
  int int_or(const std::string& k, int d) const {
    const double v = number_or(k, double(d));
    if (!(v >= -2147483648.0 && v <= 2147483647.0))
      throw std::runtime_error("config.json: \"" + k + "\" is not a representable integer");
    return int(v);
  }
  double number_or(const std::string& k, double d) const {
    const JsonValue* v = get(k);
    if (!v) return d;
    if (v->type == Number) return v->num;
    if (v->type == Bool) return v->b ? 1 : 0;
    if (v->type == String) return atof(v->str.c_str());
    return d;
  }
  std::string string_or(const std::string& k, const std::string& d) const {
    const JsonValue* v = get(k);
    if (!v) return d;
    if (v->type == String) return v->str;
    if (v->type == Number) {
      char buf[32];
      snprintf(buf, sizeof buf, "%g", v->num);
      return buf;
    }
    return d;
  }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string& s) : s_(s) {}
  JsonValue parse() {
    JsonValue v = value();
    ws();
    if (p_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;
  int depth_ = 0;
  static constexpr int kMaxDepth = 64;  // config.json nests 3 deep; bounded recursion = no stack overflow on "[[[[..."
  struct Nest {
    JsonParser& p;
    explicit Nest(JsonParser& q) : p(q) {
      if (++p.depth_ > kMaxDepth) p.fail("nesting too deep");
    }
    ~Nest() { --p.depth_; }
  };
  [[noreturn]] void fail(const char* m) const {
    throw std::runtime_error(std::string("config.json: ") + m + " at byte " + std::to_string(p_));
  }
  void ws() {
    while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) ++p_;
  }
  bool lit(const char* t) {
    size_t n = strlen(t);
    if (s_.compare(p_, n, t) == 0) {
      p_ += n;
      return true;
    }
    return false;
  }
  JsonValue value() {
    Nest nest(*this);
    ws();
    if (p_ >= s_.size()) fail("unexpected end");
    JsonValue v;
    char c = s_[p_];
    if (c == '{') {
      v.type = JsonValue::Object;
      ++p_;
      ws();
      if (p_ < s_.size() && s_[p_] == '}') {
        ++p_;
        return v;
      }
      while (true) {
        ws();
        std::string k = string();
        ws();
        if (p_ >= s_.size() || s_[p_] != ':') fail("expected ':'");
        ++p_;
        v.obj[k] = value();
        ws();
        if (p_ < s_.size() && s_[p_] == ',') {
          ++p_;
          continue;
        }
        if (p_ < s_.size() && s_[p_] == '}') {
          ++p_;
          return v;
        }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      v.type = JsonValue::Array;
      ++p_;
      ws();
      if (p_ < s_.size() && s_[p_] == ']') {
        ++p_;
        return v;
      }
      while (true) {
        v.arr.push_back(value());
        ws();
        if (p_ < s_.size() && s_[p_] == ',') {
          ++p_;
          continue;
        }
        if (p_ < s_.size() && s_[p_] == ']') {
          ++p_;
          return v;
        }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"') {
      v.type = JsonValue::String;
      v.str = string();
      return v;
    }
    if (lit("true")) {
      v.type = JsonValue::Bool;
      v.b = true;
      return v;
    }
    if (lit("false")) {
      v.type = JsonValue::Bool;
      return v;
    }
    if (lit("null")) return v;
    if (lit("NaN") || lit("Infinity") || lit("-Infinity")) {  // python json.dump extensions
      v.type = JsonValue::Number;
      return v;
    }
    char* end = nullptr;
    v.num = strtod(s_.c_str() + p_, &end);
    if (end == s_.c_str() + p_) fail("bad value");
    p_ = end - s_.c_str();
    v.type = JsonValue::Number;
    return v;
  }
  std::string string() {
    if (p_ >= s_.size() || s_[p_] != '"') fail("expected string");
    ++p_;
    std::string out;
    while (p_ < s_.size() && s_[p_] != '"') {
      char c = s_[p_++];
      if (c == '\\' && p_ < s_.size()) {
        char e = s_[p_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            unsigned cp = strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) out += char(cp);
            else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
            else { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += e;
        }
      } else {
        out += c;
      }
    }
    if (p_ >= s_.size()) fail("unterminated string");
    ++p_;
    return out;
  }
};

}  // namespace m3

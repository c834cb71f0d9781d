// Packed-weight cache: see weight_cache.h.  Host-only except upload_cached().
#include "weight_cache.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <type_traits>

#include "../../include/m3b200.h"
#include "sha256.h"

namespace m3 {
namespace {

// ---- tripwires: a field added to one of these structs must also be added to its io() below (and
// kCacheLayoutVersion bumped); the sizes are those of the one platform this library is built for (x86-64 Linux).
static_assert(sizeof(TcConvW) == 40 && sizeof(RowTcW) == 24 && sizeof(Lin) == 96, "update io(Lin) + layout version");
static_assert(sizeof(FlowTcW) == 80 && sizeof(MrfStageW) == 104 && sizeof(DecLastW) == 96, "update io() + version");
static_assert(sizeof(UpW) == 80 && sizeof(DDSW) == 432 && sizeof(EncLayerW) == 432, "update io() + layout version");

struct Writer {
  static constexpr bool reading = false;
  std::vector<uint8_t>& out;
  const float* base;
  bool ok = true;
  template <typename T>
  void pod(T& v) {
    static_assert(std::is_arithmetic<T>::value, "pod");
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
    out.insert(out.end(), p, p + sizeof(T));
  }
  void ptr(const float*& p) {
    int64_t off = p ? int64_t(p - base) : -1;
    pod(off);
  }
  void str(std::string& s) {
    uint32_t n = uint32_t(s.size());
    pod(n);
    out.insert(out.end(), s.begin(), s.end());
  }
  template <typename T>
  uint32_t count(std::vector<T>& v) {
    uint32_t n = uint32_t(v.size());
    pod(n);
    return n;
  }
};

struct Reader {
  static constexpr bool reading = true;
  const uint8_t* p;
  size_t left;
  const float* base;
  size_t n_f32, n_h16;
  bool ok = true;
  template <typename T>
  void pod(T& v) {
    static_assert(std::is_arithmetic<T>::value, "pod");
    if (!ok || left < sizeof(T)) {
      ok = false;
      return;
    }
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    left -= sizeof(T);
  }
  void ptr(const float*& q) {
    int64_t off = -1;
    pod(off);
    if (!ok) return;
    if (off == -1) {
      q = nullptr;
    } else if (off < 0 || size_t(off) >= n_f32) {
      ok = false;
    } else {
      q = base + off;
    }
  }
  void str(std::string& s) {
    uint32_t n = 0;
    pod(n);
    if (!ok || n > 4096 || left < n) {
      ok = false;
      return;
    }
    s.assign(reinterpret_cast<const char*>(p), n);
    p += n;
    left -= n;
  }
  template <typename T>
  uint32_t count(std::vector<T>& v) {
    uint32_t n = 0;
    pod(n);
    if (!ok || n > 4096) {  // no voice has more than a few dozen of anything
      ok = false;
      return 0;
    }
    v.clear();
    v.resize(n);
    return n;
  }
  void woff(unsigned long long off) {
    if (off > n_h16) ok = false;
  }
};

template <typename Ar>
void check_woff(Ar&, unsigned long long) {}
template <>
void check_woff<Reader>(Reader& a, unsigned long long off) {
  a.woff(off);
}

template <typename Ar>
void io(Ar& a, std::vector<int>& v) {
  const uint32_t n = a.count(v);
  for (uint32_t i = 0; i < n && a.ok; ++i) a.pod(v[i]);
}
template <typename Ar>
void io(Ar& a, VoiceConfig& c) {
  a.pod(c.num_symbols); a.pod(c.n_speakers); a.pod(c.inter); a.pod(c.hidden); a.pod(c.filter); a.pod(c.n_heads);
  a.pod(c.n_layers); a.pod(c.kernel_size);
  a.str(c.resblock);
  io(a, c.rb_kernels);
  {
    const uint32_t n = a.count(c.rb_dils);
    for (uint32_t i = 0; i < n && a.ok; ++i) io(a, c.rb_dils[i]);
  }
  io(a, c.up_rates);
  io(a, c.up_kernels);
  a.pod(c.up_init); a.pod(c.gin); a.pod(c.use_sdp); a.pod(c.sample_rate); a.pod(c.hop_length);
  a.pod(c.noise_scale); a.pod(c.length_scale); a.pod(c.noise_w); a.pod(c.multispeaker);
}
template <typename Ar>
void io(Ar& a, TcConvW& t) {
  a.pod(t.ok); a.pod(t.woff); a.pod(t.K); a.pod(t.NC); a.pod(t.n_chunks); a.pod(t.N); a.pod(t.taps);
  check_woff(a, t.woff);
}
template <typename Ar>
void io(Ar& a, RowTcW& t) {
  a.pod(t.ok); a.pod(t.woff); a.pod(t.nc);
  check_woff(a, t.woff);
}
template <typename Ar>
void io(Ar& a, Lin& l) {
  a.ptr(l.w); a.ptr(l.b); a.pod(l.cin); a.pod(l.cout); a.pod(l.taps);
  io(a, l.tc);
  io(a, l.rtc);
}
template <typename Ar>
void io(Ar& a, std::vector<Lin>& v) {
  const uint32_t n = a.count(v);
  for (uint32_t i = 0; i < n && a.ok; ++i) io(a, v[i]);
}
template <typename Ar>
void io(Ar& a, DDSW& d) {
  for (int i = 0; i < 3; ++i) {
    a.ptr(d.sep_w[i]); a.ptr(d.sep_b[i]);
    io(a, d.c1x1[i]);
    a.ptr(d.n1g[i]); a.ptr(d.n1b[i]); a.ptr(d.n2g[i]); a.ptr(d.n2b[i]);
  }
}
template <typename Ar>
void io(Ar& a, EncLayerW& e) {
  io(a, e.qkv); io(a, e.o); io(a, e.ffn1); io(a, e.ffn2);
  a.ptr(e.ek); a.ptr(e.ev); a.ptr(e.g1); a.ptr(e.b1); a.ptr(e.g2); a.ptr(e.b2);
}
template <typename Ar>
void io(Ar& a, ConvFlowW& c) {
  a.ptr(c.pre_w); a.ptr(c.pre_b);
  io(a, c.dds);
  io(a, c.proj);
}
template <typename Ar>
void io(Ar& a, FlowTcW& f) {
  a.pod(f.ok); a.pod(f.woff);
  a.ptr(f.in_bias); a.ptr(f.cum_bias); a.ptr(f.skip_bias); a.ptr(f.post_bias);
  a.pod(f.x0_coff); a.pod(f.x1_coff);
  a.pod(f.ok2); a.pod(f.woff2); a.ptr(f.m_bias);
  check_woff(a, f.woff);
  check_woff(a, f.woff2);
}
template <typename Ar>
void io(Ar& a, CouplingW& c) {
  io(a, c.pre); io(a, c.post); io(a, c.in); io(a, c.rs);
  a.pod(c.cond_off);
  io(a, c.ftc);
}
template <typename Ar>
void io(Ar& a, UpW& u) {
  a.ptr(u.w); a.ptr(u.b);
  a.pod(u.cin); a.pod(u.cout); a.pod(u.k); a.pod(u.u); a.pod(u.ntaps); a.pod(u.pad);
  io(a, u.tc);
}
template <typename Ar>
void io(Ar& a, ResBlockW& r) {
  a.pod(r.k);
  io(a, r.dil);
  io(a, r.c1);
  io(a, r.c2);
}
template <typename Ar>
void io(Ar& a, MrfStageW& m) {
  a.pod(m.ok);
  for (int j = 0; j < 4; ++j)
    for (int d = 0; d < 2; ++d) {
      a.pod(m.woff[j][d]);
      check_woff(a, m.woff[j][d]);
    }
  a.ptr(m.late_bias);
  a.pod(m.H); a.pod(m.HX); a.pod(m.HY); a.pod(m.nk); a.pod(m.nd);
}
template <typename Ar>
void io(Ar& a, DecLastW& d) {
  a.pod(d.ok); a.pod(d.up_woff); a.pod(d.post_woff); a.pod(d.fused_ok); a.pod(d.planes_ok); a.pod(d.blob_off); a.pod(d.blob_bytes);
  a.pod(d.f_up); a.pod(d.f_post); a.pod(d.f_postp);
  for (int i = 0; i < 3; ++i) { a.pod(d.f_c1[i]); a.pod(d.f_c2[i]); a.pod(d.HYb[i]); }
  check_woff(a, d.up_woff);
  check_woff(a, d.post_woff);
  check_woff(a, d.blob_off + (d.blob_bytes + 1) / 2);
  if (d.planes_ok) check_woff(a, d.blob_off + (d.f_postp + kDecPostPlanesBytes + 1) / 2);
}
template <typename Ar, typename T>
void io_vec(Ar& a, std::vector<T>& v) {
  const uint32_t n = a.count(v);
  for (uint32_t i = 0; i < n && a.ok; ++i) io(a, v[i]);
}
template <typename Ar>
void io(Ar& a, DeviceVoice& v) {
  io(a, v.cfg);
  io(a, v.dec_last);
  a.pod(v.tc_fmt); a.pod(v.use_tc); a.pod(v.use_rows_tc);
  io_vec(a, v.mrf);
  a.pod(v.n_params); a.pod(v.has_emb_g); a.pod(v.window);
  a.ptr(v.emb);
  io_vec(a, v.enc);
  io(a, v.enc_proj);
  a.pod(v.use_sdp); a.pod(v.dp_ch);
  io(a, v.dp_pre); io(a, v.dp_proj); io(a, v.dp_dds);
  a.pod(v.ea_m[0]); a.pod(v.ea_m[1]); a.pod(v.ea_logs[0]); a.pod(v.ea_logs[1]);
  io_vec(a, v.cflows);
  io(a, v.dpp_c1); io(a, v.dpp_c2); io(a, v.dpp_proj);
  a.ptr(v.dpp_g1); a.ptr(v.dpp_b1); a.ptr(v.dpp_g2); a.ptr(v.dpp_b2);
  a.pod(v.dp_cond_off);
  io_vec(a, v.couplings);
  a.pod(v.flow_hidden); a.pod(v.flow_layers); a.pod(v.flow_kernel);
  io(a, v.dec_pre);
  a.pod(v.dec_cond_off);
  io_vec(a, v.ups);
  io_vec(a, v.rbs);
  a.ptr(v.post_w);
  a.pod(v.post_k); a.pod(v.post_c);
  a.ptr(v.emb_g);
  io(a, v.cond_all);
  a.pod(v.n_cond);
}

#pragma pack(push, 1)
struct CacheHeader {
  char magic[8];
  uint32_t layout_version, header_bytes;
  uint64_t onnx_size;
  int64_t onnx_mtime_ns;
  uint64_t meta_bytes, f32_count, h16_count, checksum, meta_len;
  char onnx_sha256[64], config_sha256[64];
  char lib_version[32];
  char pack_flags[160];
  char pad[512 - 8 - 8 - 16 - 40 - 128 - 32 - 160];
};
#pragma pack(pop)
static_assert(sizeof(CacheHeader) == 512, "header is 512 bytes");
const char kMagic[8] = {'M', '3', 'B', '2', '0', '0', 'W', 'C'};

// 64-bit multiply-xorshift over 8-byte words (4 independent lanes so it runs at memory speed); integrity, not security
uint64_t checksum64(const uint8_t* p, size_t n) {
  uint64_t h[4] = {0x9e3779b97f4a7c15ull, 0xc2b2ae3d27d4eb4full, 0x165667b19e3779f9ull, 0x27d4eb2f165667c5ull};
  size_t i = 0;
  for (; i + 32 <= n; i += 32)
    for (int l = 0; l < 4; ++l) {
      uint64_t w;
      memcpy(&w, p + i + 8 * l, 8);
      h[l] = (h[l] ^ w) * 0x100000001b3ull;
      h[l] ^= h[l] >> 29;
    }
  uint64_t t = n;
  for (; i < n; ++i) t = (t ^ p[i]) * 0x100000001b3ull;
  uint64_t r = t;
  for (int l = 0; l < 4; ++l) r = (r ^ h[l]) * 0xff51afd7ed558ccdull, r ^= r >> 33;
  return r;
}

bool is_dir(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

void put_str(char* dst, size_t cap, const std::string& s) {
  memset(dst, 0, cap);
  memcpy(dst, s.data(), s.size() < cap ? s.size() : cap);
}
std::string get_str(const char* src, size_t cap) {
  size_t n = 0;
  while (n < cap && src[n]) ++n;
  return std::string(src, n);
}

}  // namespace

void resolve_voice_paths(const std::string& path, std::string* onnx, std::string* config) {
  std::string dir;
  if (is_dir(path)) {
    dir = path;
    *onnx = path + "/generator.onnx";
  } else {
    *onnx = path;
    const size_t slash = path.find_last_of('/');
    dir = slash == std::string::npos ? "." : path.substr(0, slash);
  }
  *config = dir + "/config.json";
}

std::string sha256_file(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  Sha256 h;
  std::vector<uint8_t> buf(1 << 20);
  size_t n;
  while ((n = fread(buf.data(), 1, buf.size(), f)) > 0) h.update(buf.data(), n);
  fclose(f);
  return h.hex();
}

CacheKey make_cache_key(const std::string& path) {
  CacheKey k;
  resolve_voice_paths(path, &k.onnx_path, &k.config_path);
  struct stat st;
  if (stat(k.onnx_path.c_str(), &st) != 0) throw std::runtime_error("cannot open " + k.onnx_path);
  k.onnx_size = uint64_t(st.st_size);
  k.onnx_mtime_ns = int64_t(st.st_mtim.tv_sec) * 1000000000ll + st.st_mtim.tv_nsec;
  k.config_sha256 = sha256_file(k.config_path);
  k.pack_flags = pack_flags_string();
  char* real = realpath(k.onnx_path.c_str(), nullptr);
  if (real) {
    k.onnx_path = real;
    free(real);
  }
  return k;
}

std::string CacheKey::file_name() const {
  Sha256 h;
  const std::string s = onnx_path + "|" + std::to_string(onnx_size) + "|" + std::to_string(onnx_mtime_ns) + "|" +
                        config_sha256 + "|" + pack_flags + "|" + std::to_string(kCacheLayoutVersion) + "|" + m3_version();
  h.update(s.data(), s.size());
  return h.hex().substr(0, 32) + ".m3w";
}

void serialize_voice(const DeviceVoice& dv, const float* base, std::vector<uint8_t>& out) {
  Writer w{out, base};
  io(w, const_cast<DeviceVoice&>(dv));
}

bool deserialize_voice(DeviceVoice& dv, const float* base, size_t n_f32, size_t n_h16, const uint8_t* p, size_t n) {
  Reader r{p, n, base, n_f32, n_h16};
  io(r, dv);
  return r.ok && r.left == 0;
}

void write_cache_file(const std::string& file, const CacheKey& key, const std::string& onnx_sha256, PackedVoice& pv) {
  for (auto& f : pv.fix) *f.first = pv.f32.data() + f.second;
  std::vector<uint8_t> meta;
  serialize_voice(*pv.dv, pv.f32.data(), meta);
  for (auto& f : pv.fix) *f.first = nullptr;
  const size_t meta_len = meta.size();
  while (meta.size() % 64) meta.push_back(0);

  CacheHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, kMagic, 8);
  h.layout_version = kCacheLayoutVersion;
  h.header_bytes = sizeof h;
  h.onnx_size = key.onnx_size;
  h.onnx_mtime_ns = key.onnx_mtime_ns;
  h.meta_bytes = meta.size();
  h.meta_len = meta_len;
  h.f32_count = pv.f32.size();
  h.h16_count = pv.h16.size();
  put_str(h.onnx_sha256, 64, onnx_sha256);
  put_str(h.config_sha256, 64, key.config_sha256);
  put_str(h.lib_version, 32, m3_version());
  if (key.pack_flags.size() >= sizeof h.pack_flags) throw std::runtime_error("weight cache: pack switches too long");
  put_str(h.pack_flags, sizeof h.pack_flags, key.pack_flags);
  // checksum over meta | f32 | h16 in file order
  std::vector<uint8_t> body;
  body.reserve(meta.size() + pv.f32.size() * 4 + pv.h16.size() * 2);
  body.insert(body.end(), meta.begin(), meta.end());
  const uint8_t* pf = reinterpret_cast<const uint8_t*>(pv.f32.data());
  body.insert(body.end(), pf, pf + pv.f32.size() * 4);
  const uint8_t* ph = reinterpret_cast<const uint8_t*>(pv.h16.data());
  body.insert(body.end(), ph, ph + pv.h16.size() * 2);
  h.checksum = checksum64(body.data(), body.size());

  const std::string tmp = file + ".tmp." + std::to_string(long(getpid()));
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot open " + tmp + " for writing");
  const bool okw = fwrite(&h, sizeof h, 1, f) == 1 && (body.empty() || fwrite(body.data(), body.size(), 1, f) == 1);
  const bool okc = fclose(f) == 0;
  if (!okw || !okc || rename(tmp.c_str(), file.c_str()) != 0) {
    remove(tmp.c_str());
    throw std::runtime_error("cannot open " + file + " (write failed)");
  }
}

std::unique_ptr<CacheImage> read_cache_file(const std::string& file, const CacheKey* key, std::string* why) {
  auto miss = [&](const char* m) {
    if (why) *why = m;
    return std::unique_ptr<CacheImage>();
  };
  FILE* f = fopen(file.c_str(), "rb");
  if (!f) return miss("absent");
  std::unique_ptr<CacheImage> img(new CacheImage());
  struct stat st;
  if (fstat(fileno(f), &st) != 0 || st.st_size < off_t(sizeof(CacheHeader)) || st.st_size > (off_t(1) << 34)) {
    fclose(f);
    return miss("damaged: size");
  }
  img->bytes.resize(size_t(st.st_size));
  const bool okr = fread(img->bytes.data(), img->bytes.size(), 1, f) == 1;
  fclose(f);
  if (!okr) return miss("damaged: short read");
  CacheHeader h;
  memcpy(&h, img->bytes.data(), sizeof h);
  if (memcmp(h.magic, kMagic, 8) != 0 || h.header_bytes != sizeof h) return miss("damaged: magic");
  if (h.layout_version != kCacheLayoutVersion) return miss("stale: layout version");
  if (get_str(h.lib_version, 32) != m3_version()) return miss("stale: library version");
  if (key) {
    if (h.onnx_size != key->onnx_size || h.onnx_mtime_ns != key->onnx_mtime_ns) return miss("stale: generator.onnx changed");
    if (get_str(h.config_sha256, 64) != key->config_sha256) return miss("stale: config.json changed");
    if (get_str(h.pack_flags, sizeof h.pack_flags) != key->pack_flags) return miss("stale: pack switches");
  }
  const size_t body = img->bytes.size() - sizeof h;
  if (h.meta_bytes > body || h.f32_count > body / 4 || h.h16_count > body / 2 ||
      h.meta_bytes + h.f32_count * 4 + h.h16_count * 2 != body || h.meta_bytes % 64 || h.meta_len > h.meta_bytes)
    return miss("damaged: section sizes");
  if (checksum64(img->bytes.data() + sizeof h, body) != h.checksum) return miss("damaged: checksum");
  img->meta_off = sizeof h;
  img->meta_bytes = size_t(h.meta_bytes);
  img->meta_len = size_t(h.meta_len);
  img->f32_off = img->meta_off + img->meta_bytes;
  img->f32_count = size_t(h.f32_count);
  img->h16_off = img->f32_off + img->f32_count * 4;
  img->h16_count = size_t(h.h16_count);
  img->onnx_sha256 = get_str(h.onnx_sha256, 64);
  {  // the meta section must parse (against a stand-in base: only offsets are checked here)
    DeviceVoice probe;
    const float* fake = reinterpret_cast<const float*>(img->bytes.data() + img->f32_off);
    if (!deserialize_voice(probe, fake, img->f32_count, img->h16_count, img->bytes.data() + img->meta_off, img->meta_len))
      return miss("damaged: meta section");
  }
  if (why) *why = "hit";
  return img;
}

std::unique_ptr<DeviceVoice> upload_cached(const CacheImage& img, int device) {
  std::unique_ptr<DeviceVoice> dv(new DeviceVoice());
  upload_slabs(*dv, device, reinterpret_cast<const float*>(img.bytes.data() + img.f32_off), img.f32_count,
               reinterpret_cast<const uint16_t*>(img.bytes.data() + img.h16_off), img.h16_count);
  if (!deserialize_voice(*dv, dv->slab, img.f32_count, img.h16_count, img.bytes.data() + img.meta_off, img.meta_len))
    throw EngineError(M3_ERR_MODEL, "weight cache: meta section does not parse");
  dv->slab_floats = img.f32_count;
  return dv;
}

}  // namespace m3

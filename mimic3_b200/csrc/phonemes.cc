// Native phonemes -> ids: the voice's phonemes.txt / phoneme_map.txt as hash tables and the
// id-sequence rules of Mimic3Voice.phonemes_to_ids (mimic3_tts/voice.py:126-152, which calls the third-party
// phonemes2ids.phonemes2ids with the PhonemesConfig of the voice, mimic3_tts/config.py:147-176; the files are read
// at voice.py:268-271 and 302-307).  Plain C++, no CUDA.  Behaviour is the one documented in
// mimic3_b200/phonemes.py (the Python restatement these entry points are fuzzed against); parity with the
// phonemes2ids package itself is pinned wherever that package can be imported (tests/golden/make_golden_phonemes2ids.py).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/m3b200.h"

namespace m3 {
void set_last_error(const std::string& m);  // m3_api.cc: the thread-local message behind m3_last_error()
}

struct m3_phoneme_table {
  std::unordered_map<std::string, int64_t> ids;
  std::unordered_map<std::string, std::vector<std::string>> map;
};

namespace {

struct Range {
  uint32_t lo, hi;
};
const Range kMarks[] = {
#include "unicode_marks.inc"
};

bool is_mark(uint32_t cp) {
  size_t lo = 0, hi = sizeof(kMarks) / sizeof(kMarks[0]);
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (cp < kMarks[mid].lo) hi = mid;
    else if (cp > kMarks[mid].hi) lo = mid + 1;
    else return true;
  }
  return false;
}

// one UTF-8 sequence starting at s[i] (lenient: a malformed byte is a code point of its own)
uint32_t decode(const std::string& s, size_t i, size_t* len) {
  const unsigned char c = static_cast<unsigned char>(s[i]);
  int n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
  if (i + n > s.size()) n = 1;
  uint32_t cp = n == 1 ? c : c & (0xFF >> (n + 1));
  for (int k = 1; k < n; ++k) {
    const unsigned char d = static_cast<unsigned char>(s[i + k]);
    if ((d >> 6) != 2) {
      *len = 1;
      return c;
    }
    cp = (cp << 6) | (d & 0x3F);
  }
  *len = size_t(n);
  return cp;
}

constexpr uint32_t kZwj = 0x200D;

void graphemes(const std::string& s, std::vector<std::string>& out) {
  bool prev_zwj = false, any = false;
  for (size_t i = 0; i < s.size();) {
    size_t n;
    const uint32_t cp = decode(s, i, &n);
    if (any && (is_mark(cp) || cp == kZwj || prev_zwj)) out.back().append(s, i, n);
    else out.emplace_back(s, i, n);
    any = true;
    prev_zwj = cp == kZwj;
    i += n;
  }
}

bool is_tone_letter(uint32_t cp) { return cp >= 0x2E5 && cp <= 0x2E9; }

// "^(.*?)([0-9]+|[tone letters]+)$" with a non-empty prefix
bool split_tone(const std::string& s, std::string* head, std::string* tone) {
  std::vector<std::pair<size_t, uint32_t>> cps;  // (byte offset, code point)
  for (size_t i = 0; i < s.size();) {
    size_t n;
    const uint32_t cp = decode(s, i, &n);
    cps.emplace_back(i, cp);
    i += n;
  }
  if (cps.size() < 2) return false;
  auto cls = [](uint32_t cp) { return (cp >= '0' && cp <= '9') ? 1 : is_tone_letter(cp) ? 2 : 0; };
  const int c = cls(cps.back().second);
  if (!c) return false;
  size_t k = cps.size();
  while (k > 0 && cls(cps[k - 1].second) == c) --k;
  if (k == 0) return false;  // the whole phoneme is a tone: the lazy prefix would be empty
  *head = s.substr(0, cps[k].first);
  *tone = s.substr(cps[k].first);
  return true;
}

struct Opts {
  bool auto_bos_eos = false, blank_at_start = true, blank_at_end = true, simple_punct = false, sep_graphemes = false,
       sep_tones = false, tone_before = false;
  int blank_between = 1;
  std::unordered_map<std::string, std::string> punct;
  std::vector<std::string> separate;  // longest first
};

void split_phoneme(const std::string& p, const Opts& o, std::vector<std::string>& out) {
  std::vector<std::string> parts{p};
  if (o.sep_tones) {
    std::string head, tone;
    if (split_tone(p, &head, &tone)) {
      parts.clear();
      if (o.tone_before) { parts.push_back(tone); parts.push_back(head); }
      else { parts.push_back(head); parts.push_back(tone); }
    }
  }
  if (!o.separate.empty()) {
    std::vector<std::string> next;
    for (const std::string& part : parts) {
      std::string cur;
      for (size_t i = 0; i < part.size();) {
        const std::string* hit = nullptr;
        for (const std::string& s : o.separate)
          if (!s.empty() && part.compare(i, s.size(), s) == 0) {
            hit = &s;
            break;
          }
        if (hit) {
          if (!cur.empty()) next.push_back(cur);
          cur.clear();
          next.push_back(*hit);
          i += hit->size();
        } else {
          size_t n;
          decode(part, i, &n);
          cur.append(part, i, n);
          i += n;
        }
      }
      if (!cur.empty()) next.push_back(cur);
    }
    parts.swap(next);
  }
  if (o.sep_graphemes) {
    std::vector<std::string> next;
    for (const std::string& part : parts) graphemes(part, next);
    parts.swap(next);
  }
  for (auto& s : parts) out.push_back(std::move(s));
}

int ph_fail(int code, const std::string& m) {
  m3::set_last_error(m);
  return code;
}

}  // namespace

extern "C" {

int32_t m3_phoneme_table_create(m3_phoneme_table** out) {
  if (!out) return ph_fail(M3_ERR_INVALID, "m3_phoneme_table_create: NULL argument");
  *out = new m3_phoneme_table();
  return M3_OK;
}

void m3_phoneme_table_free(m3_phoneme_table* t) { delete t; }

int32_t m3_phoneme_table_add(m3_phoneme_table* t, const char* phoneme, int64_t id) {
  if (!t || !phoneme) return ph_fail(M3_ERR_INVALID, "m3_phoneme_table_add: NULL argument");
  t->ids[phoneme] = id;
  return M3_OK;
}

int32_t m3_phoneme_table_add_map(m3_phoneme_table* t, const char* from, const char* const* to, int32_t n_to) {
  if (!t || !from || n_to < 0 || (n_to && !to)) return ph_fail(M3_ERR_INVALID, "m3_phoneme_table_add_map: bad argument");
  std::vector<std::string> v;
  for (int i = 0; i < n_to; ++i) v.emplace_back(to[i] ? to[i] : "");
  t->map[from] = std::move(v);
  return M3_OK;
}

// phonemes.txt: "<id><space><phoneme>" per line; blank lines and '#' comment lines skipped; the phoneme keeps inner
// and trailing spaces (voice.py:268-271 -> phonemes2ids.load_phoneme_ids)
int32_t m3_phoneme_table_load_ids(m3_phoneme_table* t, const char* phonemes_txt) {
  if (!t || !phonemes_txt) return ph_fail(M3_ERR_INVALID, "m3_phoneme_table_load_ids: NULL argument");
  std::ifstream f(phonemes_txt, std::ios::binary);
  if (!f) return ph_fail(M3_ERR_IO, std::string("cannot open ") + phonemes_txt);
  std::string line;
  int lineno = 0;
  while (std::getline(f, line)) {
    ++lineno;
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
    if (line.find_first_not_of(" \t\v\f") == std::string::npos || line[0] == '#') continue;
    const size_t sp = line.find(' ');
    const std::string num = line.substr(0, sp);
    char* end = nullptr;
    const long long id = strtoll(num.c_str(), &end, 10);
    if (num.empty() || *end)
      return ph_fail(M3_ERR_MODEL, std::string(phonemes_txt) + ":" + std::to_string(lineno) + ": id is not an integer");
    t->ids[sp == std::string::npos ? std::string() : line.substr(sp + 1)] = id;
  }
  return M3_OK;
}

// phoneme_map.txt: "<from> <to> [<to> ...]" per line (voice.py:302-307 -> phonemes2ids.utils.load_phoneme_map)
int32_t m3_phoneme_table_load_map(m3_phoneme_table* t, const char* phoneme_map_txt) {
  if (!t || !phoneme_map_txt) return ph_fail(M3_ERR_INVALID, "m3_phoneme_table_load_map: NULL argument");
  std::ifstream f(phoneme_map_txt, std::ios::binary);
  if (!f) return ph_fail(M3_ERR_IO, std::string("cannot open ") + phoneme_map_txt);
  std::string line;
  while (std::getline(f, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
    size_t b = 0;
    while (b < line.size() && (line[b] == '\r' || line[b] == '\n')) ++b;
    std::vector<std::string> parts;
    size_t i = b;
    while (true) {
      const size_t sp = line.find(' ', i);
      parts.push_back(line.substr(i, sp == std::string::npos ? std::string::npos : sp - i));
      if (sp == std::string::npos) break;
      i = sp + 1;
    }
    if (parts.size() >= 2 && !parts[0].empty()) t->map[parts[0]] = std::vector<std::string>(parts.begin() + 1, parts.end());
  }
  return M3_OK;
}

int64_t m3_phoneme_table_size(const m3_phoneme_table* t) { return t ? int64_t(t->ids.size()) : 0; }

int32_t m3_phoneme_table_lookup(const m3_phoneme_table* t, const char* phoneme, int64_t* id) {
  if (!t || !phoneme || !id) return ph_fail(M3_ERR_INVALID, "m3_phoneme_table_lookup: NULL argument");
  auto it = t->ids.find(phoneme);
  if (it == t->ids.end()) return ph_fail(M3_ERR_INVALID, std::string("no phoneme '") + phoneme + "'");
  *id = it->second;
  return M3_OK;
}

int32_t m3_phonemes_to_ids(const m3_phoneme_table* t, const m3_phoneme_opts* po, const char* const* phonemes,
                           const int32_t* word_lengths, int32_t n_words, int64_t* out_ids, int64_t out_cap,
                           int64_t* n_out) {
  if (!t || !n_out || n_words < 0 || (n_words && (!phonemes || !word_lengths)))
    return ph_fail(M3_ERR_INVALID, "m3_phonemes_to_ids: bad argument");
  *n_out = 0;
  if (po && po->struct_size < offsetof(m3_phoneme_opts, n_separate) + sizeof(int32_t))
    return ph_fail(M3_ERR_INVALID, "m3_phonemes_to_ids: opts->struct_size does not describe an m3_phoneme_opts");
  Opts o;
  const char *bos = nullptr, *eos = nullptr, *blank = nullptr, *blank_word = nullptr;
  if (po) {
    o.auto_bos_eos = po->flags & M3_PH_AUTO_BOS_EOS;
    o.blank_at_start = po->flags & M3_PH_BLANK_AT_START;
    o.blank_at_end = po->flags & M3_PH_BLANK_AT_END;
    o.simple_punct = po->flags & M3_PH_SIMPLE_PUNCTUATION;
    o.sep_graphemes = po->flags & M3_PH_SEPARATE_GRAPHEMES;
    o.sep_tones = po->flags & M3_PH_SEPARATE_TONES;
    o.tone_before = po->flags & M3_PH_TONE_BEFORE;
    if (po->blank_between < 0 || po->blank_between > 2) return ph_fail(M3_ERR_INVALID, "m3_phonemes_to_ids: blank_between");
    o.blank_between = po->blank_between;
    bos = po->bos;
    eos = po->eos;
    blank = po->blank;
    blank_word = po->blank_word;
    if (po->n_punctuation < 0) {
      o.punct = {{";", ","}, {":", ","}, {"?", "."}, {"!", "."}};  // phonemes2ids' default punctuation_map
    } else {
      for (int i = 0; i < po->n_punctuation; ++i)
        if (po->punctuation_from && po->punctuation_to && po->punctuation_from[i] && po->punctuation_to[i])
          o.punct[po->punctuation_from[i]] = po->punctuation_to[i];
    }
    for (int i = 0; i < po->n_separate; ++i)
      if (po->separate && po->separate[i] && *po->separate[i]) o.separate.emplace_back(po->separate[i]);
    // longest first (stable): "(" + "|".join(sorted(separate, key=len, reverse=True)) + ")"
    for (size_t i = 1; i < o.separate.size(); ++i)
      for (size_t j = i; j > 0 && o.separate[j - 1].size() < o.separate[j].size(); --j) std::swap(o.separate[j - 1], o.separate[j]);
  } else {
    o.punct = {{";", ","}, {":", ","}, {"?", "."}, {"!", "."}};
  }
  auto lookup = [&](const char* s, int64_t* id) {
    if (!s) return false;
    auto it = t->ids.find(s);
    if (it == t->ids.end()) return false;
    *id = it->second;
    return true;
  };
  int64_t blank_id = 0, blank_word_id = 0, bos_id = 0, eos_id = 0;
  const bool has_blank = lookup(blank, &blank_id);
  bool has_blank_word = lookup(blank_word, &blank_word_id);
  if (!has_blank_word && has_blank) {
    has_blank_word = true;
    blank_word_id = blank_id;
  }
  std::vector<std::vector<int64_t>> words;
  if (o.auto_bos_eos && lookup(bos, &bos_id)) words.push_back({bos_id});
  size_t k = 0;
  std::vector<std::string> subs;
  for (int w = 0; w < n_words; ++w) {
    if (word_lengths[w] < 0) return ph_fail(M3_ERR_INVALID, "m3_phonemes_to_ids: negative word length");
    std::vector<int64_t> ids;
    for (int j = 0; j < word_lengths[w]; ++j, ++k) {
      const char* ph = phonemes[k];
      if (!ph || !*ph) continue;
      const std::string phoneme(ph);
      const std::vector<std::string>* mapped = nullptr;
      auto mit = t->map.find(phoneme);
      std::vector<std::string> self;
      if (mit != t->map.end()) mapped = &mit->second;
      else {
        self.push_back(phoneme);
        mapped = &self;
      }
      for (const std::string& m : *mapped) {
        const std::string* p = &m;
        if (o.simple_punct) {
          auto pit = o.punct.find(m);
          if (pit != o.punct.end()) p = &pit->second;
        }
        subs.clear();
        split_phoneme(*p, o, subs);
        for (const std::string& s : subs) {
          auto it = t->ids.find(s);
          if (it != t->ids.end()) ids.push_back(it->second);  // unknown phonemes are dropped (fail_on_missing=False)
        }
      }
    }
    if (!ids.empty()) words.push_back(std::move(ids));
  }
  if (o.auto_bos_eos && lookup(eos, &eos_id)) words.push_back({eos_id});
  std::vector<int64_t> out;
  if (!words.empty()) {
    if (o.blank_at_start && has_blank) out.push_back(blank_id);
    const bool between_tokens = (o.blank_between == 0 || o.blank_between == 2) && has_blank;
    const bool between_words = (o.blank_between == 1 || o.blank_between == 2) && has_blank_word;
    for (size_t wi = 0; wi < words.size(); ++wi) {
      const bool last_word = wi + 1 == words.size();
      for (size_t ti = 0; ti < words[wi].size(); ++ti) {
        out.push_back(words[wi][ti]);
        const bool last_token = ti + 1 == words[wi].size();
        if (between_tokens && !(last_token && (last_word || between_words))) out.push_back(blank_id);
      }
      if (between_words && !last_word) out.push_back(blank_word_id);
    }
    if (o.blank_at_end && has_blank) out.push_back(blank_id);
  }
  *n_out = int64_t(out.size());
  if (int64_t(out.size()) > out_cap || (!out_ids && !out.empty()))
    return ph_fail(M3_ERR_INVALID, "m3_phonemes_to_ids: output buffer too small (n_out holds the size needed)");
  if (!out.empty()) memcpy(out_ids, out.data(), out.size() * sizeof(int64_t));
  return M3_OK;
}

}  // extern "C"

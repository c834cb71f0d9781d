/* libm3b200 -- C ABI of the B200-native VITS engine that replaces the
 * onnxruntime.InferenceSession inside Mimic 3's Mimic3Voice.ids_to_audio.
 *
 * Plain C, plain pointers and sizes; no CUDA / torch types in any signature.
 * Every entry point names the reference interface it replaces
 * (paths relative to the MycroftAI/mimic3 checkout).
 *
 * Threading: one m3_voice may be shared by any number of host threads
 * (like the shared ORT session, mimic3_tts/voice.py:277-292); m3_infer is
 * re-entrant, weights are immutable, each call takes a private stream+workspace
 * from a pool.  Errors never abort the process: a non-zero code is returned and
 * m3_last_error() (thread-local) holds the message, which the Python wrapper turns
 * into an exception the way ORT does (mimic3_http/synthesis.py:129-133).
 */
#ifndef M3B200_H_
#define M3B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3_OK 0
#define M3_ERR_INVALID 1 /* bad argument: id / sid out of range, bad shape (-> ValueError) */
#define M3_ERR_IO 2      /* voice directory / file unreadable                               */
#define M3_ERR_MODEL 3   /* generator.onnx / config.json inconsistent                       */
#define M3_ERR_CUDA 4    /* CUDA runtime error                                              */
#define M3_ERR_NOGPU 5   /* no usable sm_100 device: the engine has NO CPU fallback         */

/* flags for m3_infer */
#define M3_FLAG_KEEP_FLOAT 1u     /* also return the float32 waveform (ORT's "output")     */
#define M3_FLAG_DEBUG_TENSORS 2u  /* keep named intermediates (tests only; slow)           */
#define M3_FLAG_DEVICE_IDS 4u     /* `ids` is device memory already (benchmark "resident") */
#define M3_FLAG_NO_HOST_COPY 8u   /* leave PCM on the device; host pointers are NULL       */
#define M3_FLAG_STAGE_TIMING 16u  /* CUDA-event time per stage, read back as tensors "ms:<stage>" */

typedef struct m3_voice m3_voice;   /* loaded voice == the ORT session object (voice.py:77,83) */
typedef struct m3_result m3_result; /* one run's outputs == ORT's returned arrays             */

typedef struct m3_voice_info {
  int32_t num_symbols;     /* ModelConfig.num_symbols, mimic3_tts/config.py:116            */
  int32_t n_speakers;      /* ModelConfig.n_speakers,  config.py:117                       */
  int32_t is_multispeaker; /* TrainingConfig.is_multispeaker, config.py:316-318            */
  int32_t has_speaker_embedding; /* emb_g present: "sid" is a graph input                  */
  int32_t sample_rate;     /* AudioConfig.sample_rate, config.py:38                        */
  int32_t hop_length;      /* product of upsample_rates (== AudioConfig.hop_length)        */
  int32_t hidden_channels, inter_channels;
  float noise_scale, length_scale, noise_w; /* InferenceConfig defaults, config.py:260-262 */
  int64_t n_params;        /* fp32 parameters bound from generator.onnx                    */
  int32_t device;          /* CUDA device ordinal the weights live on                      */
  int32_t reserved;
} m3_voice_info;

const char* m3_version(void);
/* Message of the last failing call on this thread ("" if none). */
const char* m3_last_error(void);
/* Number of visible CUDA devices with compute capability 10.x (0 => M3_ERR_NOGPU on load). */
int32_t m3_device_count(void);

/* Replaces Mimic3Voice._load_model: onnxruntime.InferenceSession(str(generator_path), ...)
 * (mimic3_tts/voice.py:378-407).  `path` is the voice directory (config.json,
 * generator.onnx -- voice.py:261,273) or the generator.onnx inside it. */
int32_t m3_voice_load(const char* path, int32_t device, m3_voice** out);

/* Packed-weight cache.  The reference parses + optimises generator.onnx once per path per
 * process (voice.py:277-299, 378-407) and checks voice files against the sha256_sum entries of
 * mimic3_tts/voices.json (mimic3_tts/_resources.py:35-51, download.py:108-117).  Here the one-time conversion of
 * generator.onnx into the engine's packed operand slabs is kept as <cache_dir>/<key>.m3w and re-read on later loads;
 * the blob records the sha256 of the generator.onnx it was made from, so the manifest check is a string compare.
 * Zero-initialise, then set struct_size = sizeof(m3_load_opts). */
#define M3_LOAD_VERIFY_SHA256 1u  /* re-hash generator.onnx on a cache hit too (otherwise the recorded digest is used) */
#define M3_LOAD_NO_CACHE_WRITE 2u /* read an existing blob but never write one                                        */
typedef struct m3_load_opts {
  uint32_t struct_size;
  uint32_t flags;               /* M3_LOAD_*                                                                          */
  const char* cache_dir;        /* directory of *.m3w blobs; NULL: $M3B200_WEIGHT_CACHE, unset/empty = no cache       */
  const char* expected_sha256;  /* voices.json "sha256_sum" of generator.onnx (64 hex digits) or NULL: a mismatch is
                                   M3_ERR_MODEL, like a failed download check                                          */
} m3_load_opts;
typedef struct m3_load_stats {
  int32_t from_cache;     /* 1: slabs came from a cache blob                          */
  int32_t cache_written;  /* 1: this load wrote the blob                              */
  double parse_ms;        /* config.json + generator.onnx -> named fp32 parameters    */
  double pack_ms;         /* bind + pack into GEMM-ready slabs                        */
  double cache_read_ms;   /* read + validate the blob                                 */
  double hash_ms;         /* sha256 passes over generator.onnx                        */
  double upload_ms;       /* cudaMalloc + host->device copies                         */
  double total_ms;
  char onnx_sha256[65];   /* digest of generator.onnx ("" if never computed / recorded) */
  char cache_file[512];   /* blob path ("" without a cache directory)                 */
} m3_load_stats;
int32_t m3_voice_load_ex(const char* path, int32_t device, const m3_load_opts* opts, m3_voice** out);
int32_t m3_voice_load_stats(const m3_voice* voice, m3_load_stats* stats);
/* GPU-free halves of the same thing: convert a voice into its blob (returns the blob path in out_file), validate a
 * blob completely (header, checksum, every offset) and report the generator.onnx digest it records, hash a file the
 * way mimic3_tts/utils.py file_sha256_sum does. */
int32_t m3_weight_cache_build(const char* path, const char* cache_dir, const char* expected_sha256, char* out_file,
                              int32_t out_cap);
int32_t m3_weight_cache_check(const char* cache_file, char onnx_sha256_out[65]);
int32_t m3_sha256_file(const char* path, char out[65]);

void m3_voice_free(m3_voice* voice);
int32_t m3_voice_get_info(const m3_voice* voice, m3_voice_info* info);

/* Replaces `self.onnx_model.run(None, inputs)` + `audio_float_to_int16`
 * (mimic3_tts/voice.py:230-231, mimic3_tts/utils.py:237-244).
 *   ids      int64 [batch][t_stride]  == inputs["input"]         (voice.py:180)
 *   lengths  int64 [batch]            == inputs["input_lengths"] (voice.py:181)
 *   scales   float [3] = {noise_scale, length_scale, noise_w}    (voice.py:182-189)
 *   sid      int64 [batch] or NULL    == inputs["sid"]           (voice.py:217-218)
 *   seed     selects the Philox noise stream (noise scales > 0 only)
 * Every utterance is computed with batch-1 edge semantics (the reference never
 * batches, voice.py:180-181) and is peak-normalised on its own. */
int32_t m3_infer(m3_voice* voice, const int64_t* ids, const int64_t* lengths, int32_t batch, int32_t t_stride,
                 const float* scales, const int64_t* sid, uint64_t seed, uint32_t flags, m3_result** out);

/* Per-utterance settings and the PCM post chain of one call (all members optional; zero-initialise,
 * then set struct_size = sizeof(m3_infer_opts)).  This is what lets ONE engine call serve all sentences
 * that Mimic3TextToSpeechSystem.end_utterance() speaks one by one (mimic3_tts/tts.py:470-515): each
 * sentence carries its own Mimic3Settings (length_scale / noise_scale / noise_w / rate / volume,
 * tts.py:519-543) and is separated from its neighbours by the silence of add_break (tts.py:452-465). */
typedef struct m3_infer_opts {
  uint32_t struct_size;          /* sizeof(m3_infer_opts) of the caller's header (ABI versioning)           */
  uint32_t flags;                /* M3_FLAG_*                                                                */
  uint64_t seed;                 /* Philox noise stream                                                      */
  const float* row_scales;       /* [batch][3] {noise_scale, length_scale, noise_w} per utterance, or NULL:
                                    the `scales` argument applies to every row (voice.py:182-189)           */
  const double* volume;          /* [batch] audioop.mul(audio, 2, volume/100) factor (tts.py:540-543) or NULL */
  const int64_t* lead_silence;   /* [batch] zero samples inserted before utterance b (tts.py:452-465) or NULL */
  const int64_t* trail_silence;  /* [batch] zero samples appended after utterance b, or NULL                  */
  int32_t wav_header;            /* != 0: m3_result_stream() starts with the 44-byte RIFF/WAVE header that
                                    AudioResult.to_wav_bytes writes (opentts_abc/__init__.py:117-127)        */
  int32_t reserved;
} m3_infer_opts;

/* m3_infer with per-utterance settings and the on-device PCM post chain.  `scales` may be NULL when
 * opts->row_scales is given.  With silence, m3_result_sample_offsets()[b] is where utterance b's own
 * samples start inside the output stream (its length is num_frames[b] * hop_length) and
 * offsets[batch] is the total number of stream samples.  M3_FLAG_KEEP_FLOAT is rejected together with
 * volume / silence / wav_header. */
int32_t m3_infer_ex(m3_voice* voice, const int64_t* ids, const int64_t* lengths, int32_t batch, int32_t t_stride,
                    const float* scales, const int64_t* sid, const m3_infer_opts* opts, m3_result** out);
/* The output stream as bytes: [WAV header if asked][int16 LE samples, silences included]; host memory. */
const uint8_t* m3_result_stream(const m3_result* r, int64_t* n_bytes);
/* Host helper: the 44-byte header of 16-bit mono PCM at `sample_rate` with `n_samples` samples. */
int32_t m3_wav_header(int32_t sample_rate, int64_t n_samples, uint8_t out[44]);

int32_t m3_result_batch(const m3_result* r);
/* sample_offsets[batch+1]: utterance b occupies [off[b], off[b+1]) of the packed buffers. */
const int64_t* m3_result_sample_offsets(const m3_result* r);
const int64_t* m3_result_num_frames(const m3_result* r);      /* [batch] mel frames (sum of durations) */
const int16_t* m3_result_pcm(const m3_result* r);             /* packed int16, host (NULL with NO_HOST_COPY) */
const float* m3_result_audio(const m3_result* r);             /* packed float32, host (KEEP_FLOAT only)      */
const float* m3_result_peaks(const m3_result* r);             /* [batch] max|audio| per utterance            */
const void* m3_result_device_pcm(const m3_result* r);         /* packed int16, device memory                 */
double m3_result_device_ms(const m3_result* r);               /* GPU time of the call, CUDA events           */
int64_t m3_result_kernel_launches(const m3_result* r);        /* kernels this call launched                  */
/* Debug intermediates (M3_FLAG_DEBUG_TENSORS): row-major float [rows][cols]; returns M3_ERR_INVALID if absent. */
int32_t m3_result_tensor(const m3_result* r, const char* name, const float** data, int64_t* rows, int64_t* cols);
void m3_result_free(m3_result* r);

/* ---- phonemes -> ids: the step immediately before the engine --------------------------
 * Replaces phonemes2ids.load_phoneme_ids / load_phoneme_map (mimic3_tts/voice.py:268-271, 302-307) and
 * phonemes2ids.phonemes2ids as called by Mimic3Voice.phonemes_to_ids (voice.py:126-152) with the voice's
 * PhonemesConfig (mimic3_tts/config.py:147-176).  Strings are NUL-terminated UTF-8; NULL = Python's None. */
typedef struct m3_phoneme_table m3_phoneme_table; /* phoneme -> id and phoneme -> [phoneme...] hash tables */
#define M3_PH_AUTO_BOS_EOS 1u
#define M3_PH_BLANK_AT_START 2u
#define M3_PH_BLANK_AT_END 4u
#define M3_PH_SIMPLE_PUNCTUATION 8u
#define M3_PH_SEPARATE_GRAPHEMES 16u
#define M3_PH_SEPARATE_TONES 32u
#define M3_PH_TONE_BEFORE 64u
#define M3_PH_BLANK_BETWEEN_TOKENS 0
#define M3_PH_BLANK_BETWEEN_WORDS 1
#define M3_PH_BLANK_BETWEEN_TOKENS_AND_WORDS 2
typedef struct m3_phoneme_opts {          /* PhonemesConfig fields, config.py:147-176 */
  uint32_t struct_size;
  uint32_t flags;                         /* M3_PH_*                                                        */
  int32_t blank_between;                  /* M3_PH_BLANK_BETWEEN_*                                          */
  int32_t n_punctuation;                  /* entries of punctuation_from/to; < 0: the default ; : -> , ? ! -> . */
  const char* bos;
  const char* eos;
  const char* blank;
  const char* blank_word;
  const char* const* punctuation_from;
  const char* const* punctuation_to;
  const char* const* separate;            /* symbols split off as phonemes of their own (stress marks ...)  */
  int32_t n_separate;
  int32_t reserved;
} m3_phoneme_opts;
int32_t m3_phoneme_table_create(m3_phoneme_table** out);
void m3_phoneme_table_free(m3_phoneme_table* table);
int32_t m3_phoneme_table_load_ids(m3_phoneme_table* table, const char* phonemes_txt);      /* voice.py:268-271 */
int32_t m3_phoneme_table_load_map(m3_phoneme_table* table, const char* phoneme_map_txt);   /* voice.py:302-307 */
int32_t m3_phoneme_table_add(m3_phoneme_table* table, const char* phoneme, int64_t id);
int32_t m3_phoneme_table_add_map(m3_phoneme_table* table, const char* from, const char* const* to, int32_t n_to);
int64_t m3_phoneme_table_size(const m3_phoneme_table* table);
int32_t m3_phoneme_table_lookup(const m3_phoneme_table* table, const char* phoneme, int64_t* id);
/* `phonemes` holds the phonemes of all words back to back, word w has word_lengths[w] of them.  Unknown phonemes are
 * dropped (fail_on_missing=False, voice.py:151).  *n_out = ids produced; M3_ERR_INVALID if out_cap is smaller (then
 * *n_out is the size needed).  opts == NULL: phonemes2ids' keyword defaults (no blank symbol, no bos/eos, default punctuation map unused). */
int32_t m3_phonemes_to_ids(const m3_phoneme_table* table, const m3_phoneme_opts* opts, const char* const* phonemes,
                           const int32_t* word_lengths, int32_t n_words, int64_t* out_ids, int64_t out_cap,
                           int64_t* n_out);

/* Diagnostics: runs test #which of the tcgen05/TMEM building blocks on the current device and
 * stores the max abs error against a host reference (negative = CUDA error code). */
int32_t m3_selftest(int32_t which, double* result);

#ifdef __cplusplus
}
#endif
#endif /* M3B200_H_ */

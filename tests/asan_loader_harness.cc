// AddressSanitizer / UBSan harness for the host-side voice loader (csrc/onnx_reader.cc, csrc/voice_model.cc,
// csrc/json_min.h -- plain C++, no CUDA): loads every voice directory given on the command line and counts
// successes / clean failures.  Built and run by tests/test_cabi_and_host.py::test_loader_is_clean_under_asan.
#include <cstdio>
#include <stdexcept>
#include "voice_model.h"
int main(int argc, char** argv) {
  int ok = 0, bad = 0;
  for (int i = 1; i < argc; ++i) {
    try {
      m3::HostVoice hv = m3::load_host_voice(argv[i]);
      ++ok;
    } catch (const std::exception& e) {
      ++bad;
    }
  }
  printf("ok %d bad %d\n", ok, bad);
  return 0;
}

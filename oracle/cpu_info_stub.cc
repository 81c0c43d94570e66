// Test-harness code (NOT product code). Stands in for the reference's
// webrtc/rtc_base/cpu_info.cc when building oracle/_ref so the ISA-specific
// code path of the reference can be pinned: the
// canonical oracle is the AVX2 path; WAP_REF_ISA=sse2|scalar selects the
// others for diagnosing divergence.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "rtc_base/cpu_info.h"

namespace webrtc {
namespace cpu_info {

uint32_t DetectNumberOfCores() {
  unsigned n = std::thread::hardware_concurrency();
  return n ? n : 1;
}

bool Supports(ISA isa) {
  const char* e = std::getenv("WAP_REF_ISA");
  int level = 2;  // 0 scalar, 1 sse2, 2 avx2
  if (e) {
    if (!std::strcmp(e, "scalar")) level = 0;
    else if (!std::strcmp(e, "sse2")) level = 1;
  }
  switch (isa) {
    case ISA::kSSE2:
    case ISA::kSSE3:
      return level >= 1;
    case ISA::kAVX2:
    case ISA::kFMA3:
      return level >= 2;
    default:
      return false;
  }
}

}  // namespace cpu_info
}  // namespace webrtc

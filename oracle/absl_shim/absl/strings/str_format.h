// Test-harness shim (NOT product code).
#pragma once
#include <cstdio>
#include <string>
#include "absl/strings/has_absl_stringify.h"
namespace absl {
template <class... Args>
std::string StrFormat(const char* fmt, Args... args) {
  char buf[512];
  std::snprintf(buf, sizeof(buf), fmt, args...);
  return buf;
}
template <class Sink, class... Args>
void Format(Sink* sink, const char* fmt, Args... args) {
  char buf[512];
  std::snprintf(buf, sizeof(buf), fmt, args...);
  sink->Append(buf);
}
}  // namespace absl

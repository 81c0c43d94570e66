// Test-harness shim (NOT product code): maps the abseil vocabulary types the
// reference uses onto the C++ standard library so the unmodified reference
// sources under /root/reference compile without abseil-cpp (absent offline).
#pragma once
#include <cstring>
#include <functional>
#include <string>
#include <string_view>
namespace absl {
using string_view = std::string_view;
}

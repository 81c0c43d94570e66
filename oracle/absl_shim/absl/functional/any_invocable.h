// Test-harness shim (NOT product code).
#pragma once
#include <functional>
namespace absl {
template <class Sig>
using AnyInvocable = std::move_only_function<Sig>;
}

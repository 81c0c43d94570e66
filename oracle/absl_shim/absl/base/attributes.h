// Test-harness shim (NOT product code).
#pragma once
#define ABSL_CONST_INIT constinit
#define ABSL_MUST_USE_RESULT [[nodiscard]]
#define ABSL_DEPRECATED(msg) [[deprecated(msg)]]
#define ABSL_ATTRIBUTE_LIFETIME_BOUND
#define ABSL_ATTRIBUTE_UNUSED [[maybe_unused]]

// Test-harness shim (NOT product code).
#pragma once
#define absl_nonnull
#define absl_nullable
#define absl_nullability_unknown
#define ABSL_NULLABILITY_COMPATIBLE

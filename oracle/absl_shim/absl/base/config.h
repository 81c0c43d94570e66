// Test-harness shim (NOT product code).
#pragma once
#define ABSL_HAVE_THREAD_LOCAL 1

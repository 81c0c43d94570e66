// Kernel-launch shim: real <<<>>> under nvcc, coroutine emulator under the
// test-only g++ build (tests/emu).
#pragma once
#if defined(WAP_EMU)
#define WAP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#else
#define WAP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<dim3(grid), dim3(block), (size_t)(smem), stream>>>(__VA_ARGS__)
#endif

// Test-only CUDA execution emulator runtime (see cuda_emu.h).  x86-64 only.
#include <dlfcn.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include "cuda_emu.h"

namespace emu {

Block* g_block = nullptr;
Thread* g_cur = nullptr;
uint3 g_blockIdx{0, 0, 0};
dim3 g_blockDim;
dim3 g_gridDim;

// Saves the callee-saved registers on the current stack, stores the stack
// pointer in *save_sp and resumes the context whose stack pointer is new_sp.
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace {

constexpr size_t kStackBytes = 512 * 1024;
unsigned char* g_stacks = nullptr;
size_t g_stack_threads = 0;

void thread_finish() {
  Block* b = g_block;
  Thread* me = g_cur;
  me->done = true;
  b->alive--;
  if (b->alive == 0) {
    void* dummy;
    emu_switch(&dummy, b->main_sp);
  }
  yield_next();
  std::abort();  // never resumed
}

void thread_entry() {
  g_block->body();
  thread_finish();
}

void ensure_stacks(size_t n) {
  if (n <= g_stack_threads) return;
  if (g_stacks) munmap(g_stacks, g_stack_threads * kStackBytes);
  g_stacks = static_cast<unsigned char*>(mmap(nullptr, n * kStackBytes, PROT_READ | PROT_WRITE,
                                              MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
  if (g_stacks == MAP_FAILED) {
    std::perror("emu mmap");
    std::abort();
  }
  g_stack_threads = n;
}

}  // namespace

void yield_next() {
  Block* b = g_block;
  Thread* me = g_cur;
  int n = (int)b->threads.size();
  int i = b->cur;
  for (int step = 0; step < n; ++step) {
    i = (i + 1 == n) ? 0 : i + 1;
    if (!b->threads[i].done) break;
  }
  Thread* next = &b->threads[i];
  if (next == me && !me->done) return;
  b->cur = i;
  g_cur = next;
  emu_switch(&me->sp, next->sp);
}

namespace {
constexpr int kTraceDepth = 6;
void report_divergence(void* const* a, int na, void* const* b, int nb) {
  std::fprintf(stderr, "emu: lanes of one warp met at a barrier from different call sites\n first lane:\n");
  backtrace_symbols_fd(a, na, 2);
  std::fprintf(stderr, " this lane (%d):\n", (int)(g_cur->tid.x & 31));
  backtrace_symbols_fd(b, nb, 2);
  std::abort();
}
}  // namespace

void warp_barrier() {
  Warp& w = *g_cur->warp;
  unsigned g = w.gen;
  static const bool check = getenv("WAP_EMU_CHECK_DIVERGENCE") != nullptr;
  if (check) {
    void* pcs[kTraceDepth];
    int n = backtrace(pcs, kTraceDepth);
    if (w.arrived == 0) {
      w.first_n = n;
      for (int i = 0; i < n; ++i) w.first_trace[i] = pcs[i];
    } else if (n != w.first_n || std::memcmp(pcs, w.first_trace, n * sizeof(void*)) != 0) {
      report_divergence(w.first_trace, w.first_n, pcs, n);
    }
  }
  if (++w.arrived == 32) {
    w.arrived = 0;
    w.gen++;
    return;
  }
  g_cur->wait_pc = __builtin_return_address(0);
  unsigned long spins = 0;
  while (w.gen == g) {
    yield_next();
    if (++spins == 2000000ul) {
      // Some lanes of this warp will never arrive (exited, or a barrier inside
      // lane-divergent control flow): report where every lane is waiting.
      std::fprintf(stderr, "emu: warp barrier deadlock (arrived %d/32). lane: waiting-at / done\n", w.arrived);
      Block* b = g_block;
      for (size_t t = 0; t < b->threads.size(); ++t)
        if (b->threads[t].warp == &w)
        {
          Dl_info di{};
          void* pc = b->threads[t].wait_pc;
          if (pc) dladdr(pc, &di);
          std::fprintf(stderr, "  lane %2zu: +0x%lx %s\n", t & 31,
                       pc ? (unsigned long)((char*)pc - (char*)di.dli_fbase) : 0ul, b->threads[t].done ? "done" : "");
        }
      std::abort();
    }
  }
  g_cur->wait_pc = nullptr;
}

void block_barrier() {
  Block* b = g_block;
  unsigned g = b->barrier_gen;
  if (++b->barrier_arrived == (int)b->threads.size()) {
    b->barrier_arrived = 0;
    b->barrier_gen++;
    return;
  }
  while (b->barrier_gen == g) yield_next();
}

namespace {
void emu_segv(int sig) {
  void* frames[48];
  int n = backtrace(frames, 48);
  const char msg[] = "emu: fatal signal in emulated kernel, backtrace:\n";
  (void)!write(2, msg, sizeof(msg) - 1);
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}
void install_segv_handler() {
  static bool done = false;
  if (done || !getenv("WAP_EMU_BACKTRACE")) return;
  done = true;
  static unsigned char alt[1 << 16];
  stack_t ss{};
  ss.ss_sp = alt;
  ss.ss_size = sizeof(alt);
  sigaltstack(&ss, nullptr);
  struct sigaction sa{};
  sa.sa_handler = emu_segv;
  sa.sa_flags = SA_ONSTACK;
  sigaction(SIGSEGV, &sa, nullptr);
  sigaction(SIGBUS, &sa, nullptr);
  sigaction(SIGFPE, &sa, nullptr);
}
}  // namespace

void run_block(const std::function<void()>& body, dim3 grid, dim3 block, size_t smem_bytes) {
  install_segv_handler();
  unsigned old_csr = _mm_getcsr();
  _mm_setcsr(old_csr | 0x8040);  // FTZ | DAZ, as the device build's -ftz=true
  size_t nthreads = (size_t)block.x * block.y * block.z;
  if (nthreads % 32 != 0) {
    std::fprintf(stderr, "emu: block size must be a multiple of 32\n");
    std::abort();
  }
  ensure_stacks(nthreads);
  g_gridDim = grid;
  g_blockDim = block;
  std::vector<unsigned char> smem(smem_bytes + 64);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block b;
        b.body = body;
        b.smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 63) & ~uintptr_t(63));
        std::memset(smem.data(), 0xCD, smem.size());
        b.threads.resize(nthreads);
        b.warps.resize(nthreads / 32);
        b.alive = (int)nthreads;
        for (size_t t = 0; t < nthreads; ++t) {
          Thread& th = b.threads[t];
          th.tid.x = (unsigned)(t % block.x);
          th.tid.y = (unsigned)((t / block.x) % block.y);
          th.tid.z = (unsigned)(t / ((size_t)block.x * block.y));
          th.warp = &b.warps[t / 32];
          // Initial frame: six zeroed callee-saved registers, then the return
          // address; after `ret` rsp must be 8 mod 16 as at a normal call.
          uintptr_t top = reinterpret_cast<uintptr_t>(g_stacks + (t + 1) * kStackBytes);
          top &= ~uintptr_t(15);
          void** sp = reinterpret_cast<void**>(top);
          *--sp = nullptr;                                   // alignment pad / fake return
          *--sp = reinterpret_cast<void*>(&thread_entry);    // ret target
          for (int r = 0; r < 6; ++r) *--sp = nullptr;
          th.sp = sp;
        }
        g_blockIdx = uint3{bx, by, bz};
        g_block = &b;
        b.cur = 0;
        g_cur = &b.threads[0];
        emu_switch(&b.main_sp, g_cur->sp);
        g_block = nullptr;
        g_cur = nullptr;
      }
  _mm_setcsr(old_csr);
}

}  // namespace emu

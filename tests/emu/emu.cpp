// TEST INFRASTRUCTURE ONLY — fiber scheduler and warp collectives of the CPU emulation (see cuda_runtime.h).
#include "cuda_runtime.h"

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

// ---- context switch
#if defined(__x86_64__)
extern "C" void emu_switch_sp(void** from_sp, void* const* to_sp);
asm(R"(
.text
.globl emu_switch_sp
.type emu_switch_sp,@function
emu_switch_sp:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq (%rsi), %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch_sp,.-emu_switch_sp
)");
static inline void ctx_switch(emu::Context* from, emu::Context* to) { emu_switch_sp(&from->sp, &to->sp); }
static inline void ctx_make(emu::Context* c, char* stack, size_t size, void (*fn)()) {
  // after the six pops and the ret of emu_switch_sp the fiber starts in fn with a call-like stack alignment
  uintptr_t top = ((uintptr_t)stack + size) & ~(uintptr_t)15;
  void** sp = (void**)top;
  *--sp = nullptr;             // fake return address of fn (it never returns)
  *--sp = (void*)fn;
  for (int i = 0; i < 6; ++i) *--sp = nullptr;
  c->sp = sp;
}
#else
static inline void ctx_switch(emu::Context* from, emu::Context* to) { swapcontext(&from->uc, &to->uc); }
static inline void ctx_make(emu::Context* c, char* stack, size_t size, void (*fn)()) {
  getcontext(&c->uc); c->uc.uc_stack.ss_sp = stack; c->uc.uc_stack.ss_size = size; c->uc.uc_link = nullptr; makecontext(&c->uc, fn, 0);
}
#endif

namespace emu {
thread_local Block* g_blk = nullptr;
thread_local unsigned long long g_rng = 12345;
thread_local std::function<void()> g_body;

namespace {
struct Wait { const unsigned* ptr = nullptr; unsigned val = 0; };
thread_local std::vector<Wait> g_wait;
thread_local std::vector<int> g_exited;          // per warp: lanes that have returned

unsigned alive_mask(Block* b, int w) { return b->warps[w].alive; }

void complete(Block* b, int w, Block::Slot& S) {
  const unsigned g = S.gen & 1u;
  unsigned bal = 0;
  for (int l = 0; l < 32; ++l) {
    S.res[g][l] = S.val[l];
    if ((S.mask >> l & 1u) && !b->f[w * 32 + l].done && (S.val[l] & 1ull)) bal |= 1u << l;
  }
  S.ballot[g] = bal; S.arrived = 0; ++S.gen;
}

void trampoline() {
  Block* b = g_blk;
  const int me = b->cur;
  g_body();
  b->f[me].done = true;
  --b->alive;
  // a lane that returns no longer takes part in collectives: complete those that were waiting for it
  const int w = me / 32;
  b->warps[w].alive &= ~(1u << (me % 32));
  const unsigned alive = alive_mask(b, w);
  for (auto& S : b->warps[w].slots)
    if (S.arrived > 0 && S.arrived == __builtin_popcount(S.mask & alive)) complete(b, w, S);
  ctx_switch(&b->f[me].ctx, &b->sched);
}

void wait_on(const unsigned* ptr, unsigned val) {
  Block* b = g_blk;
  const int me = b->cur;
  while (*ptr == val) {
    g_wait[me].ptr = ptr; g_wait[me].val = val;
    ctx_switch(&b->f[me].ctx, &b->sched);
  }
  g_wait[me].ptr = nullptr;
}

// deposits v among the lanes of `mask`; returns the slot and the generation parity that holds the result
Block::Slot& collective(unsigned mask, unsigned long long v, unsigned* parity) {
  Block* b = g_blk;
  const int me = b->cur, w = me / 32, lane = me % 32;
  Block::Warp& W = b->warps[w];
  Block::Slot* S = nullptr;
  for (auto& s : W.slots) if (s.mask == mask) { S = &s; break; }
  if (!S) { W.slots.emplace_back(); S = &W.slots.back(); S->mask = mask; }
  const unsigned gen = S->gen;
  *parity = gen & 1u;
  S->val[lane] = v;
  ++S->arrived;
  if (S->arrived == __builtin_popcount(mask & alive_mask(b, w))) complete(b, w, *S);
  else wait_on(&S->gen, gen);
  return *S;
}
}  // namespace

void yield() {
  Block* b = g_blk;
  if (!b || b->cur < 0) return;
  const int me = b->cur;
  ctx_switch(&b->f[me].ctx, &b->sched);
}

void run_block(dim3 grid, dim3 block, dim3 bid, const std::function<void()>& body) {
  constexpr size_t STACK = 256 * 1024;
  const int n = (int)(block.x * block.y * block.z);
  if (n % 32) { fprintf(stderr, "emu: block size must be a multiple of 32\n"); abort(); }
  Block blk;
  blk.f.resize(n);
  blk.warps.resize(n / 32);
  blk.alive = n;
  g_wait.assign(n, Wait());
  g_exited.assign(n / 32, 0);
  g_blk = &blk;
  g_body = body;
  gridDim = grid; blockDim = block; blockIdx = bid;
  for (int i = 0; i < n; ++i) {
    Fiber& f = blk.f[i];
    static thread_local std::vector<char*> pool;                 // fiber stacks are reused across blocks and launches
    while ((int)pool.size() <= i) pool.push_back((char*)malloc(STACK));
    f.tid = dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
    ctx_make(&f.ctx, pool[i], STACK, trampoline);
  }
  while (blk.alive > 0) {
    // a random runnable fiber: random start, linear probe (most fibers are runnable most of the time)
    int pick = -1;
    const int start = (int)(rnd() % (unsigned)n);
    for (int k = 0; k < n; ++k) {
      const int i = start + k < n ? start + k : start + k - n;
      if (!blk.f[i].done && (g_wait[i].ptr == nullptr || *g_wait[i].ptr != g_wait[i].val)) { pick = i; break; }
    }
    if (pick < 0) { fprintf(stderr, "emu: deadlock (%d threads alive, none runnable)\n", blk.alive); abort(); }
    blk.cur = pick;
    threadIdx = blk.f[pick].tid;
    ctx_switch(&blk.sched, &blk.f[pick].ctx);
  }
  blk.cur = -1;
  g_blk = nullptr;
}
}  // namespace emu

void __syncthreads() {
  emu::Block* b = emu::g_blk;
  const unsigned gen = b->bar_gen;
  int live = 0;
  for (auto& f : b->f) live += f.done ? 0 : 1;
  if (++b->bar_arrived == live) { b->bar_arrived = 0; ++b->bar_gen; }
  else emu::wait_on(&b->bar_gen, gen);
}
// barrier + OR of the predicates: accumulate, barrier, read, barrier, reset, barrier
int __syncthreads_or(int pred) {
  static thread_local int acc = 0;
  if (pred) acc = 1;
  __syncthreads();
  const int r = acc;
  __syncthreads();
  acc = 0;
  __syncthreads();
  return r;
}
void __syncwarp(unsigned mask) { unsigned g; emu::collective(mask, 0, &g); }
unsigned __activemask() {            // lanes of the warp that have not returned (kernels here do not diverge around collectives)
  emu::Block* b = emu::g_blk;
  return emu::alive_mask(b, b->cur / 32);
}
unsigned __ballot_sync(unsigned mask, int pred) {
  unsigned g;
  emu::Block::Slot& S = emu::collective(mask, pred ? 1ull : 0ull, &g);
  return S.ballot[g];
}
unsigned long long emu_shfl(unsigned mask, unsigned long long v, int x, int mode) {
  emu::Block* b = emu::g_blk;
  const int lane = b->cur % 32;
  unsigned g;
  emu::Block::Slot& S = emu::collective(mask, v, &g);
  int src = mode == 0 ? (x & 31) : mode == 2 ? (lane ^ x) & 31 : lane - x;
  if (src < 0) src = lane;
  return S.res[g][src];
}

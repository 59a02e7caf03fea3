// TEST INFRASTRUCTURE ONLY — fiber scheduler and warp collectives of the CPU emulation (see cuda_runtime.h).
#include "cuda_runtime.h"

dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
Block* g_blk = nullptr;
unsigned long long g_rng = 12345;
std::function<void()> g_body;

namespace {
struct Wait { const unsigned* ptr = nullptr; unsigned val = 0; };
std::vector<Wait> g_wait;
std::vector<int> g_exited;          // per warp: lanes that have returned

void trampoline() {
  Block* b = g_blk;
  const int me = b->cur;
  g_body();
  b->f[me].done = true;
  --b->alive;
  // a lane that returns no longer takes part in collectives: complete one that was waiting for it
  const int w = me / 32;
  ++g_exited[w];
  Block::Warp& W = b->warps[w];
  if (W.arrived > 0 && W.arrived == 32 - g_exited[w]) {
    const unsigned g = W.gen & 1u;
    unsigned bal = 0;
    for (int l = 0; l < 32; ++l) W.res[g][l] = W.val[l];
    for (int l = 0; l < 32; ++l) if (!b->f[w * 32 + l].done && (W.val[l] & 1ull)) bal |= 1u << l;
    W.ballot[g] = bal; W.arrived = 0; ++W.gen;
  }
  swapcontext(&b->f[me].ctx, &b->sched);
}

void wait_on(const unsigned* ptr, unsigned val) {
  Block* b = g_blk;
  const int me = b->cur;
  while (*ptr == val) {
    g_wait[me].ptr = ptr; g_wait[me].val = val;
    swapcontext(&b->f[me].ctx, &b->sched);
  }
  g_wait[me].ptr = nullptr;
}

// deposits v, returns the generation parity whose res[] / ballot[] hold the result
unsigned collective(unsigned long long v) {
  Block* b = g_blk;
  const int me = b->cur, w = me / 32, lane = me % 32;
  Block::Warp& W = b->warps[w];
  const unsigned gen = W.gen, g = gen & 1u;
  W.val[lane] = v;
  ++W.arrived;
  if (W.arrived == 32 - g_exited[w]) {
    unsigned bal = 0;
    for (int l = 0; l < 32; ++l) {
      W.res[g][l] = W.val[l];
      if (!b->f[w * 32 + l].done && (W.val[l] & 1ull)) bal |= 1u << l;
    }
    W.ballot[g] = bal; W.arrived = 0; ++W.gen;
  } else {
    wait_on(&W.gen, gen);
  }
  return g;
}
}  // namespace

void yield() {
  Block* b = g_blk;
  if (!b || b->cur < 0) return;
  const int me = b->cur;
  swapcontext(&b->f[me].ctx, &b->sched);
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
    static std::vector<char*> pool;                 // fiber stacks are reused across blocks and launches
    while ((int)pool.size() <= i) pool.push_back((char*)malloc(STACK));
    f.tid = dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = pool[i];
    f.ctx.uc_stack.ss_size = STACK;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, trampoline, 0);
  }
  std::vector<int> runnable;
  while (blk.alive > 0) {
    runnable.clear();
    for (int i = 0; i < n; ++i)
      if (!blk.f[i].done && (g_wait[i].ptr == nullptr || *g_wait[i].ptr != g_wait[i].val)) runnable.push_back(i);
    if (runnable.empty()) { fprintf(stderr, "emu: deadlock (%d threads alive, none runnable)\n", blk.alive); abort(); }
    // run a random runnable fiber; stay on it for a random number of scheduling points to get long and short interleavings
    const int pick = runnable[rnd() % runnable.size()];
    blk.cur = pick;
    threadIdx = blk.f[pick].tid;
    swapcontext(&blk.sched, &blk.f[pick].ctx);
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
void __syncwarp(unsigned) { emu::collective(0); }
unsigned __ballot_sync(unsigned, int pred) {
  emu::Block* b = emu::g_blk;
  const int w = b->cur / 32;
  const unsigned g = emu::collective(pred ? 1ull : 0ull);
  return b->warps[w].ballot[g];
}
unsigned long long emu_shfl(unsigned long long v, int x, int mode) {
  emu::Block* b = emu::g_blk;
  const int me = b->cur, w = me / 32, lane = me % 32;
  const unsigned g = emu::collective(v);
  int src = mode == 0 ? (x & 31) : lane - x;
  if (src < 0) src = lane;
  return b->warps[w].res[g][src];
}

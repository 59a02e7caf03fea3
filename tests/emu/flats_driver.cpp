// TEST INFRASTRUCTURE ONLY — runs the Garbrecht-Martz flat resolution of taudem_b200/csrc/flats.cu on the CPU
// emulation, on one strip or on several row strips (one host thread per strip, td_strip_comm callbacks that
// rendezvous on a barrier: what taudem_b200/dist.py does with torch.distributed).  tests/test_emu.py
// compares the resolved directions with the oracle.
#include <pthread.h>

#include <thread>

#include "flats_emu.inc"   // the transformed kernel source (written by tests/test_emu.py)

using td::Strip;

namespace {
struct World;
struct Rank {
  World* w = nullptr;
  int id = 0;
  Strip s;
  std::vector<float> elev;
  std::vector<short> p;
  std::vector<float> ang;
  std::vector<double> dxc, dyc, thA, thB;
  int row0 = 0;
  int rc = 0;
  long long left = 0;
};
struct World {
  int n = 0;
  pthread_barrier_t bar;
  std::vector<Rank> r;
  // published by every rank before the barrier
  std::vector<const void*> ptr;
  std::vector<unsigned long long> val;
  long long collectives = 0;
};

int cb_share(void* user, void* arr, int elem) {
  Rank* R = (Rank*)user; World* W = R->w;
  W->ptr[R->id] = arr;
  pthread_barrier_wait(&W->bar);
  const size_t rowb = (size_t)R->s.pitch * elem;
  char* mine = (char*)arr;
  if (R->id > 0) {                                 // the last owned row of the strip above -> my halo row 0
    const Rank& A = W->r[R->id - 1];
    memcpy(mine, (const char*)W->ptr[A.id] + (size_t)A.s.ny * rowb, rowb);
  }
  if (R->id + 1 < W->n) {                          // the first owned row of the strip below -> my halo row ny+1
    const Rank& B = W->r[R->id + 1];
    memcpy(mine + (size_t)(R->s.ny + 1) * rowb, (const char*)W->ptr[B.id] + rowb, rowb);
  }
  pthread_barrier_wait(&W->bar);
  if (R->id == 0) ++W->collectives;
  return 0;
}
int cb_collect(void* user, const void* arr, int elem, void* recv_top, void* recv_bot) {
  Rank* R = (Rank*)user; World* W = R->w;
  W->ptr[R->id] = arr;
  pthread_barrier_wait(&W->bar);
  const size_t rowb = (size_t)R->s.pitch * elem;
  if (R->id > 0 && recv_top) {                     // what the strip above holds in ITS bottom halo row (= my row 1)
    const Rank& A = W->r[R->id - 1];
    memcpy(recv_top, (const char*)W->ptr[A.id] + (size_t)(A.s.ny + 1) * rowb, rowb);
  }
  if (R->id + 1 < W->n && recv_bot) memcpy(recv_bot, (const char*)W->ptr[R->id + 1], rowb);   // the strip below's halo row 0 (= my row ny)
  pthread_barrier_wait(&W->bar);
  if (R->id == 0) ++W->collectives;
  return 0;
}
int cb_allreduce(void* user, unsigned long long* v, int n) {
  Rank* R = (Rank*)user; World* W = R->w;
  if (n != 1) return 1;
  W->val[R->id] = v[0];
  pthread_barrier_wait(&W->bar);
  unsigned long long s = 0;
  for (int i = 0; i < W->n; ++i) s += W->val[i];
  pthread_barrier_wait(&W->bar);
  v[0] = s;
  if (R->id == 0) ++W->collectives;
  return 0;
}
}  // namespace

// dir: int16 p (D8, flat cells 0) or float ang (D-infinity, flat cells -1) after the slope stencil, resolved in place;
// fel is not modified (the strips work on copies, like the reference works on its elevDEM partition).
extern "C" int emu_flats(int dinf, const float* fel, void* dir, int nx, int ny, double dx, double dy, int nstrips, unsigned long long seed,
                         long long* left_out, long long* collectives_out) {
  if (nstrips < 1 || ny / nstrips < 1) return 1;
  World W;
  W.n = nstrips; W.r.resize(nstrips); W.ptr.assign(nstrips, nullptr); W.val.assign(nstrips, 0);
  pthread_barrier_init(&W.bar, nullptr, nstrips);
  const int per = ny / nstrips;
  for (int i = 0; i < nstrips; ++i) {
    Rank& R = W.r[i];
    R.w = &W; R.id = i; R.row0 = i * per;
    const int sny = i == nstrips - 1 ? ny - i * per : per;
    td_strip ts; ts.nx = nx; ts.ny = sny; ts.pitch = (nx + 31) / 32 * 32; ts.has_top = i > 0; ts.has_bot = i + 1 < nstrips;
    R.s = Strip(ts);
    const size_t n = (size_t)R.s.cells();
    R.elev.assign(n, 0.f); R.p.assign(n, 0); R.ang.assign(n, 0.f);
    R.dxc.assign(sny, dx); R.dyc.assign(sny, dy); R.thA.assign(sny, atan2(dy, dx)); R.thB.assign(sny, atan2(dx, dy));
    for (int r = 0; r <= sny + 1; ++r) {            // halo rows current on entry (DistTools.share)
      const int gr = R.row0 + r - 1;
      if (gr < 0 || gr >= ny) continue;
      for (int c = 0; c < nx; ++c) {
        const size_t o = (size_t)R.s.idx(r, c), src = (size_t)gr * nx + c;
        R.elev[o] = fel[src];
        if (dinf) R.ang[o] = ((const float*)dir)[src]; else R.p[o] = ((const short*)dir)[src];
      }
    }
  }
  auto work = [&](int i) {
    Rank& R = W.r[i];
    emu::g_rng = (seed + 17 * i) * 2654435761ull + 1;
    td_ctx ctx;
    td_strip_comm comm; comm.user = &R; comm.share = cb_share; comm.collect = cb_collect; comm.allreduce_sum = cb_allreduce;
    if (dinf) R.rc = td::resolve_flats_dinf(&ctx, R.elev.data(), R.ang.data(), R.s, R.dxc.data(), R.dyc.data(), R.thA.data(), R.thB.data(),
                                            &R.left, nstrips > 1 ? &comm : nullptr, nullptr);
    else R.rc = td::resolve_flats_d8(&ctx, R.elev.data(), R.p.data(), R.s, R.dxc.data(), R.dyc.data(), &R.left, nstrips > 1 ? &comm : nullptr, nullptr);
  };
  std::vector<std::thread> th;
  for (int i = 0; i < nstrips; ++i) th.emplace_back(work, i);
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&W.bar);
  for (auto& R : W.r) {
    if (R.rc) return R.rc;
    for (int r = 1; r <= R.s.ny; ++r)
      for (int c = 0; c < nx; ++c) {
        const size_t dst = (size_t)(R.row0 + r - 1) * nx + c;
        if (dinf) ((float*)dir)[dst] = R.ang[R.s.idx(r, c)]; else ((short*)dir)[dst] = R.p[R.s.idx(r, c)];
      }
  }
  if (left_out) *left_out = W.r[0].left;
  if (collectives_out) *collectives_out = W.collectives;
  return 0;
}

// TEST INFRASTRUCTURE ONLY — the shared-memory tile dataflow sweep (taudem_b200/csrc/sweep_tiles.cu) on the CPU
// emulation: persistent CTAs (run one after the other: the first one drains the queue), the ticket queue, the
// four-state tile protocol, the in-tile wavefront with its spin loops.  mode 0 = the whole sweep with the tile
// kernel; mode 1 = "hybrid": every tile once, then k_ready + k_walk.
#include <string>

#include "sweep_tiles_emu.inc"
#include "kernels.h"
#include "fill_emu.inc"

using td::Strip;
extern "C" int emu_ref_deps(int dinf, const void* dir, unsigned short* node, unsigned char* cnt, int nx, int ny, float dir_nodata, double dx, double dy);

extern "C" int emu_tiles(int dinf, int hybrid, const void* dir, float* out, const float* wgt, int nx, int ny, float dir_nodata, int usew,
                         int contcheck, float w_nodata, double dx, double dy, unsigned long long seed, unsigned long long* visits) {
  emu::g_rng = seed * 2654435761ull + 1;
  td_strip ts; ts.nx = nx; ts.ny = ny; ts.pitch = (nx + 31) / 32 * 32; ts.has_top = 0; ts.has_bot = 0;
  const Strip s(ts);
  const size_t n = (size_t)s.cells();
  std::vector<unsigned short> node(n, 0), nd0((size_t)nx * ny);
  std::vector<unsigned char> cnt((n + 3) / 4 * 4, 0xff), c0((size_t)nx * ny);
  std::vector<float> area(n, -1.0f), w(n, 0.f), ang(n, 0.f);
  std::vector<double> theta(2 * (size_t)ny), dxc(ny, dx);
  for (int j = 0; j < ny; ++j) { theta[j] = atan2(dy, dx); theta[ny + j] = atan2(dx, dy); }
  if (emu_ref_deps(dinf, dir, nd0.data(), c0.data(), nx, ny, dir_nodata, dx, dy)) return 1;
  for (int r = 1; r <= ny; ++r)
    for (int c = 0; c < nx; ++c) {
      const size_t o = (size_t)s.idx(r, c), src = (size_t)(r - 1) * nx + c;
      node[o] = nd0[src]; cnt[o] = c0[src];
      if (dinf) ang[o] = ((const float*)dir)[src];
      if (wgt) w[o] = wgt[src];
    }
  td_ctx ctx;
  ctx.node.p = node.data(); ctx.node.cap = node.size() * 2;
  ctx.cnt.p = cnt.data(); ctx.cnt.cap = cnt.size();
  ctx.sweep_dinf = dinf;
  std::vector<int> halo(2 * (size_t)s.pitch, 0);
  int rc = td::sweep_begin(&ctx, s, nullptr);
  ctx.sweep_once = hybrid;
  if (!rc) rc = td::sweep_run(&ctx, dinf != 0, area.data(), usew ? w.data() : nullptr, ang.data(), s, w_nodata, usew, contcheck, theta.data(),
                              dxc.data(), halo.data(), nullptr);
  ctx.sweep_once = 0;
  if (visits) *visits = ctx.d_ctr[24 + 3];
  if (!rc && hybrid)
    rc = td::sweep_walk(&ctx, dinf != 0, area.data(), usew ? w.data() : nullptr, ang.data(), s, w_nodata, usew, contcheck, theta.data(),
                        dxc.data(), halo.data(), nullptr);
  for (int r = 1; r <= ny; ++r)
    for (int c = 0; c < nx; ++c) out[(size_t)(r - 1) * nx + c] = area[s.idx(r, c)];
  ctx.tileflags.release();
  return rc;
}

// pitremove (taudem_b200/csrc/fill.cu): initialisation + tile-local Planchon-Darboux relaxation over an active-tile list
extern "C" int emu_fill(const float* dem, float* fel, const short* mask, int nx, int ny, float nodata, int four, unsigned long long seed) {
  emu::g_rng = seed * 2654435761ull + 1;
  td_strip ts; ts.nx = nx; ts.ny = ny; ts.pitch = (nx + 31) / 32 * 32; ts.has_top = 0; ts.has_bot = 0;
  const Strip s(ts);
  const size_t n = (size_t)s.cells();
  std::vector<float> z(n, 0.f), W(n, 0.f);
  std::vector<short> m(n, 0);
  for (int r = 1; r <= ny; ++r)
    for (int c = 0; c < nx; ++c) { z[s.idx(r, c)] = dem[(size_t)(r - 1) * nx + c]; if (mask) m[s.idx(r, c)] = mask[(size_t)(r - 1) * nx + c]; }
  td_ctx ctx;
  int rc = td::fill_init(z.data(), mask ? m.data() : nullptr, W.data(), s, nodata, four, nullptr);
  int changed = 0;
  if (!rc) rc = td::fill_relax(&ctx, z.data(), W.data(), s, four, &changed, nullptr);
  for (int r = 1; r <= ny; ++r)
    for (int c = 0; c < nx; ++c) fel[(size_t)(r - 1) * nx + c] = W[s.idx(r, c)];
  ctx.tileflags.release();
  return rc;
}

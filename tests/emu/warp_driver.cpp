// TEST INFRASTRUCTURE ONLY — the warp-per-tile dataflow sweep (taudem_b200/csrc/sweep_warp.cu) on the CPU emulation:
// the warps of a persistent CTA are independent workers (ticket queue, four-state tile protocol, warp-local wavefront
// with shared-memory counts), interleaved at random at every atomic / volatile load / fence.
#include <string>

#include "sweep_warp_emu.inc"
#include "kernels.h"

using td::Strip;
extern "C" int emu_ref_deps(int dinf, const void* dir, unsigned short* node, unsigned char* cnt, int nx, int ny, float dir_nodata, double dx, double dy);

extern "C" int emu_wtiles(int dinf, const void* dir, float* out, const float* wgt, int nx, int ny, float dir_nodata, int usew,
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
  td::make_prop_row(theta[0], true, &ctx.prop);
  std::vector<int> halo(2 * (size_t)s.pitch, 0);
  int rc = td::wsweep_begin(&ctx, s, nullptr);
  if (!rc) rc = td::wsweep_run(&ctx, dinf != 0, area.data(), usew ? w.data() : nullptr, ang.data(), s, w_nodata, usew, contcheck, theta.data(),
                               dxc.data(), halo.data(), nullptr);
  if (visits) *visits = ctx.d_ctr[24 + 3];
  for (int r = 1; r <= ny; ++r)
    for (int c = 0; c < nx; ++c) out[(size_t)(r - 1) * nx + c] = area[s.idx(r, c)];
  // every cell of the flow field must have been evaluated (count byte 0xFE)
  for (int r = 1; r <= ny && !rc; ++r)
    for (int c = 0; c < nx; ++c) { const unsigned char b = cnt[s.idx(r, c)]; if (b <= 8) { rc = 77; break; } }
  ctx.tileflags.release();
  return rc;
}

// TEST INFRASTRUCTURE ONLY — pit removal (taudem_b200/csrc/fill.cu) on the CPU emulation.
#include <string>

#include "kernels.h"
#include "fill_emu.inc"

using td::Strip;

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

// TEST INFRASTRUCTURE ONLY — runs one of the stencil kernels (selected with -DEMU_WHICH=...) on the CPU emulation:
// 1 = k_d8_stencil, 2 = k_dinf_stencil, 3 = k_deps_d8, 4 = k_deps_dinf (taudem_b200/csrc).  tests/test_emu.py compares
// the rasters with the oracle and the dependency state with the plain-loop restatement in driver.cpp.
#if EMU_WHICH == 1
#include "d8_stencil_emu.inc"
#include "rowfact_emu.inc"
#elif EMU_WHICH == 2
#include "dinf_stencil_emu.inc"
#elif EMU_WHICH == 3
#include "area_d8_emu.inc"
#else
#include "area_dinf_emu.inc"
#endif
#include "ctx.h"

using td::Strip;

namespace {
struct Grid {
  Strip s;
  std::vector<double> dxc, dyc, th;
  Grid(int nx, int ny, double dx, double dy) {
    td_strip ts; ts.nx = nx; ts.ny = ny; ts.pitch = (nx + 31) / 32 * 32; ts.has_top = 0; ts.has_bot = 0;
    s = Strip(ts);
    dxc.assign(ny, dx); dyc.assign(ny, dy); th.assign(2 * (size_t)ny + 2, atan2(dy, dx));     // + the rows above / below the strip (theta_of_row)
    for (int j = 0; j < ny; ++j) { th[j] = atan2(dy, dx); th[ny + j] = atan2(dx, dy); }
  }
  template <typename T> std::vector<T> in(const T* src) const {
    std::vector<T> v((size_t)s.cells(), T(0));
    for (int r = 1; r <= s.ny; ++r) for (int c = 0; c < s.nx; ++c) v[s.idx(r, c)] = src[(size_t)(r - 1) * s.nx + c];
    return v;
  }
  template <typename T> void out(const std::vector<T>& v, T* dst) const {
    for (int r = 1; r <= s.ny; ++r) for (int c = 0; c < s.nx; ++c) dst[(size_t)(r - 1) * s.nx + c] = v[s.idx(r, c)];
  }
};
}  // namespace

#if EMU_WHICH == 1
extern "C" int emu_d8_stencil(const float* fel, short* p, float* sd8, int nx, int ny, float nodata, double dx, double dy, unsigned long long* nflat) {
  Grid g(nx, ny, dx, dy);
  auto e = g.in(fel);
  std::vector<short> d((size_t)g.s.cells(), 0); std::vector<float> sl((size_t)g.s.cells(), 0.f);
  *nflat = 0;
  std::vector<td::RowFact> rf(ny);
  td::launch_row_factors(g.dxc.data(), g.dyc.data(), nullptr, nullptr, rf.data(), ny, nullptr);
  if (td::launch_d8_stencil(e.data(), d.data(), sl.data(), rf.data(), g.s, nodata, nflat, nullptr)) return 1;
  g.out(d, p); g.out(sl, sd8);
  return 0;
}
#elif EMU_WHICH == 2
extern "C" int emu_dinf_stencil(const float* fel, float* ang, float* slp, int nx, int ny, float nodata, double dx, double dy, unsigned long long* nflat) {
  Grid g(nx, ny, dx, dy);
  auto e = g.in(fel);
  std::vector<float> a((size_t)g.s.cells(), 0.f), sl((size_t)g.s.cells(), 0.f);
  *nflat = 0;
  std::vector<td::RowFact> rf(ny);
  td::launch_row_factors(g.dxc.data(), g.dyc.data(), g.th.data(), g.th.data() + ny, rf.data(), ny, nullptr);
  if (td::launch_dinf_stencil(e.data(), a.data(), sl.data(), rf.data(), g.s, nodata, nflat, nullptr)) return 1;
  g.out(a, ang); g.out(sl, slp);
  return 0;
}
#elif EMU_WHICH == 3
extern "C" int emu_deps_d8(const short* p, unsigned short* node, unsigned char* cnt, float* area, int nx, int ny, short nodata) {
  Grid g(nx, ny, 30., 30.);
  auto d = g.in(p);
  std::vector<unsigned short> nd((size_t)g.s.cells(), 0); std::vector<unsigned char> cn((size_t)g.s.cells() + 4, 0); std::vector<float> ar((size_t)g.s.cells(), 0.f);
  td::launch_deps_d8(d.data(), nd.data(), cn.data(), ar.data(), g.s, nodata, nullptr, -1.0f);
  g.out(nd, node); g.out(cn, cnt); g.out(ar, area);
  return 0;
}
#else
extern "C" int emu_deps_dinf(const float* ang, unsigned short* node, unsigned char* cnt, float* area, int nx, int ny, float nodata, double dx, double dy) {
  Grid g(nx, ny, dx, dy);
  auto a = g.in(ang);
  std::vector<unsigned short> nd((size_t)g.s.cells(), 0); std::vector<unsigned char> cn((size_t)g.s.cells() + 4, 0); std::vector<float> ar((size_t)g.s.cells(), 0.f);
  td::launch_deps_dinf(a.data(), nd.data(), cn.data(), ar.data(), g.s, nodata, g.th.data(), nullptr, -1.0f);
  g.out(nd, node); g.out(cn, cnt); g.out(ar, area);
  return 0;
}
#endif

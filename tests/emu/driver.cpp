// TEST INFRASTRUCTURE ONLY — runs the level / ready / walk kernels of taudem_b200/csrc/sweep_walk.cu on the
// CPU emulation (cuda_runtime.h, emu.cpp) from a flow-direction grid; tests/test_emu.py compares the result
// with the oracle.  The dependency state (node words, counts) is rebuilt here in plain loops following the
// description of k_deps_d8 / k_deps_dinf (taudem_b200/csrc/area_d8.cu, area_dinf.cu).
#include <string>

#include "sweep_walk_emu.inc"   // the transformed kernel source (written by tests/test_emu.py)

namespace td {
unsigned long long g_launches = 0;
static std::string g_err;
void set_error(const std::string& m) { g_err = m; }
int cuda_fail(cudaError_t, const char* what) { g_err = what; return 90; }
}  // namespace td

td_ctx::td_ctx() { d_ctr = (unsigned long long*)calloc(32, 8); h_ctr = (unsigned long long*)calloc(32, 8); }
td_ctx::~td_ctx() {
  free(d_ctr); free(h_ctr);
  node.p = cnt.p = nullptr;       // owned by the caller below
  listA.release(); listB.release(); listC.release();
}

using td::Strip;
using td::dcol;
using td::drow;

extern "C" int emu_sweep(int dinf, int mode, int passes, const void* dir, float* out, const float* wgt, int nx, int ny, float dir_nodata,
                         int usew, int contcheck, float w_nodata, double dx, double dy, unsigned long long seed) {
  emu::g_rng = seed * 2654435761ull + 1;
  td_strip ts; ts.nx = nx; ts.ny = ny; ts.pitch = (nx + 31) / 32 * 32; ts.has_top = 0; ts.has_bot = 0;
  const Strip s(ts);
  const size_t n = (size_t)s.cells();
  std::vector<unsigned short> node(n, 0);
  std::vector<unsigned char> cnt((n + 3) / 4 * 4, 0xff);
  std::vector<float> area(n, -1.0f), w(n, 0.f), ang(n, 0.f);
  std::vector<short> p(n, 0);
  std::vector<double> theta(2 * (size_t)ny), dxc(ny, dx);
  for (int j = 0; j < ny; ++j) { theta[j] = atan2(dy, dx); theta[ny + j] = atan2(dx, dy); }
  for (int r = 1; r <= ny; ++r)
    for (int c = 0; c < nx; ++c) {
      const size_t o = (size_t)s.idx(r, c), src = (size_t)(r - 1) * nx + c;
      if (dinf) ang[o] = ((const float*)dir)[src]; else p[o] = ((const short*)dir)[src];
      if (wgt) w[o] = wgt[src];
    }
  const unsigned VALID = 0x8000u, CON = 0x1000u;
  if (!dinf) {
    const short nd = (short)dir_nodata;
    for (int r = 1; r <= ny; ++r)
      for (int c = 0; c < nx; ++c) {
        const int d = p[s.idx(r, c)];
        if (d == nd || d < 0 || d > 8) continue;
        unsigned mask = 0; bool con = false;
        for (int k = 1; k <= 8; ++k) {
          const int rn = r + drow(k), cn = c + dcol(k);
          const bool on = s.on_grid(rn, cn);
          const int dn = on ? p[s.idx(rn, cn)] : nd;
          const bool miss = !on || dn == nd;
          const bool toward = (dn - k == 4) || (dn - k == -4);
          const bool inrange = dn >= 0 && dn <= 8;
          if (!miss && toward && inrange) mask |= 1u << (k - 1);
          if (miss || (toward && !inrange)) con = true;
        }
        node[s.idx(r, c)] = (unsigned short)(VALID | (con ? CON : 0u) | ((unsigned)d << 8) | mask);
        cnt[s.idx(r, c)] = (unsigned char)__builtin_popcount(mask);
      }
  } else {
    std::vector<unsigned char> code(n, 0);
    for (int r = 1; r <= ny; ++r)
      for (int c = 0; c < nx; ++c) {
        const float av = ang[s.idx(r, c)];
        if (fabsf(av - dir_nodata) < 1e-5f) continue;
        const td::Outflow o = td::dinf_outflow(av, theta[r - 1]);
        code[s.idx(r, c)] = (unsigned char)(o.k1 | (o.k2 << 4));
      }
    for (int r = 1; r <= ny; ++r)
      for (int c = 0; c < nx; ++c) {
        if (fabsf(ang[s.idx(r, c)] - dir_nodata) < 1e-5f) continue;
        unsigned mask = 0; bool con = false;
        for (int k = 1; k <= 8; ++k) {
          const int rn = r + drow(k), cn = c + dcol(k);
          if (!s.on_grid(rn, cn) || fabsf(ang[s.idx(rn, cn)] - dir_nodata) < 1e-5f) { con = true; continue; }
          const int kk = k > 4 ? k - 4 : k + 4;
          const unsigned cd = code[s.idx(rn, cn)];
          if ((int)(cd & 15u) == kk || (int)(cd >> 4) == kk) mask |= 1u << (k - 1);
        }
        node[s.idx(r, c)] = (unsigned short)(VALID | (con ? CON : 0u) | mask);
        cnt[s.idx(r, c)] = (unsigned char)__builtin_popcount(mask);
      }
  }
  td_ctx ctx;
  ctx.node.p = node.data(); ctx.node.cap = node.size() * 2;
  ctx.cnt.p = cnt.data(); ctx.cnt.cap = cnt.size();
  std::vector<int> halo(2 * (size_t)s.pitch, 0);
  int rc = 0;
  if (mode == 1)
    rc = td::sweep_levels(&ctx, dinf != 0, passes, area.data(), usew ? w.data() : nullptr, ang.data(), s, w_nodata, usew, contcheck,
                          theta.data(), dxc.data(), halo.data(), nullptr);
  if (!rc)
    rc = td::sweep_walk(&ctx, dinf != 0, area.data(), usew ? w.data() : nullptr, ang.data(), s, w_nodata, usew, contcheck, theta.data(),
                        dxc.data(), halo.data(), nullptr);
  for (int r = 1; r <= ny; ++r)
    for (int c = 0; c < nx; ++c) out[(size_t)(r - 1) * nx + c] = area[s.idx(r, c)];
  return rc;
}

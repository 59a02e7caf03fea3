// Shared device/host helpers for the taudem_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/taudem_b200.h"

namespace td {

// Neighbour offsets, k = 1..8 = E,NE,N,NW,W,SW,S,SE (reference src/commonLib.h:83-84;
// row index grows southward).
__host__ __device__ __forceinline__ constexpr int dcol(int k) { return (k == 1 || k == 2 || k == 8) ? 1 : (k >= 4 && k <= 6) ? -1 : 0; }
__host__ __device__ __forceinline__ constexpr int drow(int k) { return (k >= 2 && k <= 4) ? -1 : (k >= 6 && k <= 8) ? 1 : 0; }

// Reference constants (src/commonLib.h:76-81).
#define TD_PI 3.14159265359
#define TD_MISSINGSHORT ((short)-32768)
#define TD_MISSINGFLOAT (-FLT_MAX)
#define TD_MINEPS 1e-5f

// isNodata of linearpart<float/short> (src/linearpart.h:471-483):
//   abs((float)(v - nodata)) < MINEPS
__device__ __forceinline__ bool nd_f(float v, float nd) { return fabsf(v - nd) < TD_MINEPS; }
__device__ __forceinline__ bool nd_s(short v, short nd) { return fabsf((float)((int)v - (int)nd)) < TD_MINEPS; }

struct Strip {
  int nx, ny, pitch, has_top, has_bot;
  __host__ __device__ Strip() {}
  __host__ __device__ Strip(const td_strip& s) : nx(s.nx), ny(s.ny), pitch(s.pitch), has_top(s.has_top), has_bot(s.has_bot) {}
  // strip rows are 0..ny+1 (0 and ny+1 are halo rows)
  __host__ __device__ __forceinline__ bool on_grid(int r, int c) const {   // hasAccess
    return c >= 0 && c < nx && ((r >= 1 && r <= ny) || (r == 0 && has_top) || (r == ny + 1 && has_bot));
  }
  __host__ __device__ __forceinline__ bool owned(int r, int c) const { return c >= 0 && c < nx && r >= 1 && r <= ny; }
  // cell on the edge of the whole grid (one of its 4-neighbours is off-grid)
  __host__ __device__ __forceinline__ bool global_edge(int r, int c) const {
    return c == 0 || c == nx - 1 || (r == 1 && !has_top) || (r == ny && !has_bot);
  }
  __host__ __device__ __forceinline__ long long idx(int r, int c) const { return (long long)r * pitch + c; }
  __host__ __device__ __forceinline__ long long cells() const { return (long long)(ny + 2) * pitch; }
};

// ----------------------------------------------------------------------------
// TMA (bulk async copy) row-strip tile loader.
//
// A tile covers TH output rows x TW output columns; shared memory receives
// TH+2 rows (one halo row above/below) of TW + 2*HP elements, HP = 16 bytes of
// halo padding each side so that every row copy is a 16-byte aligned,
// 16-byte-multiple 1-D bulk copy (cp.async.bulk.shared::cluster.global with
// mbarrier complete_tx).  Smem row t <-> strip row r0-1+t, smem column s <->
// grid column c0-HP+s.  Rows outside [0, ny+1] and columns outside [0, pitch)
// are not copied (their smem contents are never used for decisions: callers
// test coordinates with Strip::on_grid first).
// ----------------------------------------------------------------------------
#ifndef TD_EMU   // (tests/emu compiles the queue / count protocols for the CPU; no TMA there)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

#endif  // TD_EMU

template <typename T, int TW, int TH>
struct TileGeom {
  static constexpr int HP = 16 / (int)sizeof(T);
  static constexpr int SW = TW + 2 * HP;          // smem row stride (elements)
  static constexpr int ROWS = TH + 2;
  static constexpr int ELEMS = SW * ROWS;
};

#ifndef TD_EMU
// Must be called by every thread of the CTA (contains __syncthreads()).  `bar` must be
// a fresh (never used) mbarrier word in shared memory; the tile is loaded once per CTA.
template <typename T, int TW, int TH>
__device__ __forceinline__ void load_tile_tma(T* tile, uint64_t* bar, const T* __restrict__ g, const Strip& s, int r0, int c0) {
  using G = TileGeom<T, TW, TH>;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (tid < 32) {
    const int cs = max(c0 - G::HP, 0), ce = min(c0 + TW + G::HP, s.pitch);
    const uint32_t row_bytes = (uint32_t)(ce - cs) * (uint32_t)sizeof(T);
    const int gr_lo = max(r0 - 1, 0), gr_hi = min(r0 + TH, s.ny + 1);   // inclusive strip rows
    if (tid == 0) mbar_expect_tx(bar, row_bytes * (uint32_t)(gr_hi - gr_lo + 1));
    __syncwarp();
    for (int gr = gr_lo + tid; gr <= gr_hi; gr += 32) {
      const int t = gr - (r0 - 1);
      tma_load_1d(tile + t * G::SW + (cs - (c0 - G::HP)), g + (long long)gr * s.pitch + cs, row_bytes, bar);
    }
  }
  mbar_wait(bar, 0);
}

#else   // TD_EMU: the same staging contract with plain copies
template <typename T, int TW, int TH>
__device__ __forceinline__ void load_tile_tma(T* tile, uint64_t*, const T* __restrict__ g, const Strip& s, int r0, int c0) {
  using G = TileGeom<T, TW, TH>;
  const int cs = max(c0 - G::HP, 0), ce = min(c0 + TW + G::HP, s.pitch);
  const int gr_lo = max(r0 - 1, 0), gr_hi = min(r0 + TH, s.ny + 1);
  __syncthreads();
  for (int gr = gr_lo + (int)threadIdx.x; gr <= gr_hi; gr += (int)blockDim.x) {
    const int t = gr - (r0 - 1);
    for (int c = cs; c < ce; ++c) tile[t * G::SW + (c - (c0 - G::HP))] = g[(long long)gr * s.pitch + c];
  }
  __syncthreads();
}
#endif  // TD_EMU

// launch accounting (bench gpu_launches)
extern unsigned long long g_launches;
#define TD_LAUNCHED() (++::td::g_launches)

}  // namespace td

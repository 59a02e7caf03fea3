// Persistent tile pipeline of the 3x3 stencil kernels (sm_100a): 2-D TMA tiles through a ring of shared-memory stages.
//
// A stencil kernel is a grid of persistent CTAs (SMs x resident CTAs).  CTA b handles the tiles b, b + grid, b + 2 grid, ...
// of the strip in raster order (neighbouring CTAs work on neighbouring tiles at the same time, so the halo rows / columns
// two tiles share are L2 hits).  One tile = TH x TW cells; shared memory receives the tile with a one-row / HP-column halo
// ((TH + 2) x (TW + 2 HP) elements, HP = 16 bytes of elements so that the box is a multiple of 16 bytes) with ONE
// cp.async.bulk.tensor.2d per stage (a tensor map over the (ny + 2) x pitch strip; what lies outside the strip is
// zero-filled by the TMA unit and never used for decisions: callers test coordinates first).  STAGES tiles are in flight
// per CTA: while the threads compute tile i, the copies of tiles i+1 .. i+STAGES-1 are under way; a stage is re-armed by
// one thread after the __syncthreads() that ends the tile's computation (full barrier = mbarrier with complete_tx).
#pragma once
#include "common.cuh"

#ifndef TD_EMU
#include <cuda.h>
#endif

namespace td {

struct TileMap {
#ifndef TD_EMU
  CUtensorMap m;
#else
  const void* base; int pitch, rows, elem, box_w, box_h;
#endif
};

// host: tensor map over a strip of `rows` rows x `pitch` elements of `elem` bytes, box = box_w x box_h elements
// (tile_pipe.cpp; the driver entry point is looked up at run time — the library has no link-time dependency on libcuda)
#ifndef TD_EMU
int make_tile_map(TileMap* tm, const void* base, int elem, int pitch, int rows, int box_w, int box_h);
// persistent grid of a stencil kernel: SMs x resident CTAs, at most one CTA per tile (sets the kernel's dynamic
// shared-memory limit on first use)
int stencil_grid(const void* kernel, int threads, size_t smem, long long ntiles, int* grid);
#else
inline int make_tile_map(TileMap* tm, const void* base, int elem, int pitch, int rows, int box_w, int box_h) {
  tm->base = base; tm->elem = elem; tm->pitch = pitch; tm->rows = rows; tm->box_w = box_w; tm->box_h = box_h;
  return 0;
}
inline int stencil_grid(const void*, int, size_t, long long ntiles, int* grid) { *grid = (int)(ntiles < 3 ? ntiles : 3); return 0; }
#endif

#ifndef TD_EMU
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const TileMap* tm, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}
#define TD_GRID_CONSTANT __grid_constant__
#else
#define TD_GRID_CONSTANT
#endif

template <typename T, int TW, int TH, int STAGES>
struct TileRing {
  using G = TileGeom<T, TW, TH>;
  static constexpr int ELEMS = G::ELEMS;
  static constexpr uint32_t BYTES = (uint32_t)(G::ELEMS * sizeof(T));
  static constexpr size_t STAGE_BYTES = (BYTES + 127u) / 128u * 128u;          // every stage starts 128-byte aligned
  static constexpr size_t SMEM = STAGES * STAGE_BYTES + STAGES * sizeof(uint64_t);

  unsigned char* base;
  uint64_t* full;
  int ntx;
  long long ntiles;
  int stage;
  uint32_t phase;

  __device__ __forceinline__ T* buf(int st) const { return reinterpret_cast<T*>(base + (size_t)st * STAGE_BYTES); }
  __device__ __forceinline__ void coords(long long t, int& r0, int& c0) const {
    const int ty = (int)(t / ntx), tx = (int)(t - (long long)ty * ntx);
    r0 = 1 + ty * TH; c0 = tx * TW;
  }
  // one thread
  __device__ __forceinline__ void issue(const TileMap* tm, int st, long long t) {
    int r0, c0;
    coords(t, r0, c0);
#ifndef TD_EMU
    mbar_expect_tx(full + st, BYTES);
    tma_load_2d(buf(st), tm, c0 - G::HP, r0 - 1, full + st);
#else
    T* dst = buf(st);
    for (int j = 0; j < G::ROWS; ++j)
      for (int i = 0; i < G::SW; ++i) {
        const int gr = r0 - 1 + j, gc = c0 - G::HP + i;
        dst[j * G::SW + i] = (gr >= 0 && gr < tm->rows && gc >= 0 && gc < tm->pitch) ? reinterpret_cast<const T*>(tm->base)[(long long)gr * tm->pitch + gc] : T(0);
      }
#endif
  }
  // all threads of the CTA; smem = dynamic shared memory (128-byte aligned)
  __device__ __forceinline__ void init(unsigned char* smem, const TileMap* tm, const Strip& s) {
    base = smem;
    full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    ntx = (s.pitch + TW - 1) / TW;
    ntiles = (long long)ntx * ((s.ny + TH - 1) / TH);
    stage = 0; phase = 0;
#ifndef TD_EMU
    if (threadIdx.x == 0) {
      for (int i = 0; i < STAGES; ++i) mbar_init(full + i, 1);
      mbar_fence_init();
    }
#endif
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < STAGES; ++i) {
        const long long t = (long long)blockIdx.x + (long long)i * gridDim.x;
        if (t < ntiles) issue(tm, i, t);
      }
#ifdef TD_EMU
    __syncthreads();
#endif
  }
  // all threads: the staged tile t (and its coordinates) once its bytes have landed
  __device__ __forceinline__ const T* acquire(long long t, int& r0, int& c0) {
    coords(t, r0, c0);
#ifndef TD_EMU
    mbar_wait(full + stage, phase);
#endif
    return buf(stage);
  }
  // all threads, after the last read of the stage: re-arm it with the tile STAGES grid strides ahead
  __device__ __forceinline__ void release(const TileMap* tm, long long t) {
    __syncthreads();
    const long long nt = t + (long long)STAGES * gridDim.x;
    if (threadIdx.x == 0 && nt < ntiles) issue(tm, stage, nt);
#ifdef TD_EMU
    __syncthreads();
#endif
    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
  }
};

}  // namespace td

// Multi-GPU contributing area behind the file-level entry points (mgpu.cu): one forked process per GPU.
#pragma once
#include <stddef.h>

namespace td {
struct MgpuJob {
  int dinf = 0;                 // 0 = aread8 (int16 directions), 1 = areadinf (float angles)
  const char* dirfile = nullptr;
  const char* wfile = nullptr;  // weight grid (usew)
  int usew = 0, contcheck = 1;
  int nx = 0, ny = 0;
  float* out = nullptr;         // nx * ny floats in a mapping from mgpu_alloc_shared: every rank stores its rows
};
int mgpu_world();                                  // TAUDEM_B200_GPUS (1 = the single-GPU path)
void* mgpu_alloc_shared(size_t bytes);             // anonymous shared mapping (visible to the forked ranks)
void mgpu_free_shared(void* p, size_t bytes);
// runs the job on `world` ranks; compute_seconds = the slowest rank's time from the dependency stencil to the end of the sweep
int mgpu_area(const MgpuJob& job, int world, double* compute_seconds, int* rounds);
}  // namespace td

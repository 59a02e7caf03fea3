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
// the other three tools of the path on row strips: tool 0 = pitremove (out0 = fel), 1 = d8flowdir (out0 = p int16, out1 = sd8),
// 2 = dinfflowdir (out0 = ang, out1 = slp); the rasters live in mappings from mgpu_alloc_shared
struct MgpuFlowJob {
  int tool = 0;
  const char* demfile = nullptr;
  const char* maskfile = nullptr;   // pitremove: depression mask (use_mask)
  int use_mask = 0, four = 0;
  int nx = 0, ny = 0;
  void* out0 = nullptr;
  float* out1 = nullptr;
};
int mgpu_world();                                  // TAUDEM_B200_GPUS (1 = the single-GPU path)
void* mgpu_alloc_shared(size_t bytes);             // anonymous shared mapping (visible to the forked ranks)
void mgpu_free_shared(void* p, size_t bytes);
// runs the job on `world` ranks; compute_seconds = the slowest rank's time from the dependency stencil to the end of the sweep
int mgpu_area(const MgpuJob& job, int world, double* compute_seconds, int* rounds);
// rounds = relaxation / exchange rounds of pitremove, 0 for the flow directions; flats_left = unresolved flat cells of the whole grid
int mgpu_flow(const MgpuFlowJob& job, int world, double* compute_seconds, int* rounds, long long* flats_left);
}  // namespace td

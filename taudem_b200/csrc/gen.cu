// Synthetic fractal DEM / weight-grid generators (bench and test inputs):
// multi-octave lattice value noise with an integer hash (amplitude 2^(-H*octave)),
// scaled to 1000 m relief + 100 m, plus the plane tilt*1000*(y + 0.37 x)/n.
// Written straight into device strips so that the 65536^2 bench input never touches the host.
#include "common.cuh"

namespace td {
namespace {
__host__ __device__ __forceinline__ unsigned hash3(unsigned x, unsigned y, unsigned z) {
  unsigned h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u + 0xC2B2AE3Du) ^ (z * 0x27D4EB2Fu + 0x165667B1u);
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
__host__ __device__ __forceinline__ float u01(unsigned h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

__global__ void k_gen_dem(float* __restrict__ dem, Strip s, int row0, int total_ny, unsigned seed, float hurst, float tilt) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x, r = 1 + blockIdx.x;   // rows on grid.x (no 65535 limit)
  if (c >= s.nx) return;
  const int y = row0 + r - 1;
  const int n = max(s.nx, total_ny);
  int noct = 0;
  while ((n >> (noct + 1)) >= 2) ++noct;      // lattice spacing from n/2 down to 2 cells
  float sum = 0.f, norm = 0.f;
  for (int o = 0; o < noct; ++o) {
    const float L = (float)n / (float)(2 << o);
    const float gx = (float)c / L, gy = (float)y / L;
    const float fx0 = floorf(gx), fy0 = floorf(gy);
    const unsigned ix = (unsigned)fx0, iy = (unsigned)fy0;
    float fx = gx - fx0, fy = gy - fy0;
    fx = fx * fx * (3.f - 2.f * fx); fy = fy * fy * (3.f - 2.f * fy);
    const unsigned so = seed + 7919u * (unsigned)o;
    const float v00 = u01(hash3(ix, iy, so)), v10 = u01(hash3(ix + 1, iy, so));
    const float v01 = u01(hash3(ix, iy + 1, so)), v11 = u01(hash3(ix + 1, iy + 1, so));
    const float a = v00 + (v10 - v00) * fx, b = v01 + (v11 - v01) * fx;
    const float amp = exp2f(-hurst * (float)o);
    sum += amp * (a + (b - a) * fy);
    norm += amp;
  }
  const float v = sum / norm;
  dem[s.idx(r, c)] = 100.f + 1000.f * v + tilt * 1000.f * ((float)y + 0.37f * (float)c) / (float)n;
}

__global__ void k_gen_w(float* __restrict__ w, Strip s, int row0, unsigned seed) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x, r = 1 + blockIdx.x;   // rows on grid.x (no 65535 limit)
  if (c >= s.nx) return;
  w[s.idx(r, c)] = u01(hash3((unsigned)c, (unsigned)(row0 + r - 1), seed));
}
}  // namespace

cudaError_t launch_gen_dem(float* dem, const Strip& s, int row0, int total_ny, unsigned seed, float hurst, float tilt, cudaStream_t st) {
  dim3 grid(s.ny, (s.nx + 255) / 256);
  k_gen_dem<<<grid, 256, 0, st>>>(dem, s, row0, total_ny, seed, hurst, tilt);
  TD_LAUNCHED();
  return cudaGetLastError();
}
cudaError_t launch_gen_w(float* w, const Strip& s, int row0, unsigned seed, cudaStream_t st) {
  dim3 grid(s.ny, (s.nx + 255) / 256);
  k_gen_w<<<grid, 256, 0, st>>>(w, s, row0, seed);
  TD_LAUNCHED();
  return cudaGetLastError();
}
}  // namespace td

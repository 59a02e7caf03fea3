// Multi-GPU contributing area behind the executables: `TAUDEM_B200_GPUS=N aread8 ...` / `areadinf ...`.
//
// reference: the callers' contract is `mpiexec -n N aread8` (src/aread8.cpp:57,100: MPI_Init, one row strip per rank,
// src/linearpart.h:160-200 the partition, src/aread8.cpp:280-304 the border exchange + ringTerm loop).  Here the
// executable itself forks one process per GPU; every rank reads its own rows (plus one halo row either side) of the
// input rasters, runs the dependency stencil and the sweep on its device strip through the device-strip level of the
// C ABI, and stores its rows of the result into a shared mapping that the parent writes as one GeoTIFF.
//
// Two ways to get across the strip boundary:
//   peer   : the sweep kernels deliver into the neighbour GPU themselves (CUDA IPC + NVLink system-scope atomics,
//            sweep_warp.cu); the processes only exchange IPC handles through the shared mapping and meet at barriers.
//   rounds : the reference's scheme (evaluate until no cell of the strip is ready, hand the decrements and the edge rows
//            to the neighbours, repeat until nobody handed anything over), staged through the shared mapping.  Used when
//            two ranks share a device or the devices cannot reach each other; TAUDEM_B200_PEER=0/1 overrides.
// The parent never touches CUDA (a CUDA context does not survive fork()).
#include <cuda_runtime.h>
#include <sched.h>
#include <signal.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/taudem_b200.h"
#include "mgpu.h"
#include "tiff_io.h"

namespace td {
void set_error(const std::string& msg);

namespace {
constexpr int MAXR = 64;
struct Shared {
  std::atomic<int> err;
  std::atomic<unsigned> bar_count, bar_gen;
  std::atomic<long long> handed[2];            // rounds mode: decrements handed over in this round (by round parity)
  double secs[MAXR];
  int rounds;
  long long flats_left;
  unsigned long long red[MAXR][8];             // allreduce_sum staging
  unsigned char handles[MAXR][320];
  int meta[MAXR][8];
  int device[MAXR], can_peer[MAXR];
  char msg[MAXR][256];
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// meets the other ranks; false if any rank failed (the caller bails out: nobody waits for a dead rank)
bool barrier(Shared* S, int world) {
  const unsigned gen = S->bar_gen.load();
  if (S->bar_count.fetch_add(1) + 1 == (unsigned)world) { S->bar_count.store(0); S->bar_gen.fetch_add(1); }
  else {
    for (long spins = 0; S->bar_gen.load() == gen; ++spins) {
      if (S->err.load()) return false;
      if (spins < 20000) sched_yield(); else usleep(100);
    }
  }
  return S->err.load() == 0;
}

struct Fail { std::string what; };
#define MG_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw Fail{std::string(#x) + ": " + cudaGetErrorString(e_)}; } while (0)
#define MG_TD(x) do { int rc_ = (x); if (rc_ != TD_OK) throw Fail{std::string(#x) + ": " + td_last_error()}; } while (0)
#define MG_BAR() do { if (!barrier(S, world)) throw Fail{"another rank failed"}; } while (0)

void partition(int total_ny, int world, int rank, int* row0, int* ny) {       // linearpart::init, src/linearpart.h:160-200
  const int n = total_ny / world;
  *row0 = rank * n;
  *ny = n + (rank == world - 1 ? total_ny % world : 0);
}

// rows [row0 - 1, row0 + ny] of a raster -> the device strip (rows 0 .. ny + 1), through a pinned buffer
void load_strip(tdio::Raster& r, tdio::DType t, void* d_strip, int nx, int pitch, int row0, int ny, int total_ny, cudaStream_t st) {
  const int eb = tdio::dtype_bytes(t);
  MG_CUDA(cudaMemsetAsync(d_strip, 0, (size_t)(ny + 2) * pitch * eb, st));
  const long first = row0 > 0 ? row0 - 1 : 0, last = std::min<long>(total_ny, (long)row0 + ny + 1);     // [first, last)
  const long blk = std::max<long>(1, (64l << 20) / ((long)nx * eb));
  void* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2];
  for (int i = 0; i < 2; ++i) { MG_CUDA(cudaMallocHost(&pin[i], (size_t)blk * nx * eb)); MG_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); }
  int b = 0;
  for (long y = first; y < last; y += blk, b ^= 1) {
    const long n = std::min<long>(blk, last - y);
    MG_CUDA(cudaEventSynchronize(ev[b]));                       // the copy that used this buffer two blocks ago
    std::string err;
    if (!r.read(0, y, n, nx, pin[b], t, &err)) throw Fail{"read: " + err};
    char* dst = (char*)d_strip + (size_t)(y - row0 + 1) * pitch * eb;
    MG_CUDA(cudaMemcpy2DAsync(dst, (size_t)pitch * eb, pin[b], (size_t)nx * eb, (size_t)nx * eb, (size_t)n, cudaMemcpyHostToDevice, st));
    MG_CUDA(cudaEventRecord(ev[b], st));
  }
  MG_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 2; ++i) { cudaFreeHost(pin[i]); cudaEventDestroy(ev[i]); }
}

struct RoundBuf {                // rounds mode: what a rank shows its neighbours (in the shared mapping, after Shared)
  static size_t bytes(int pitch) { return (size_t)pitch * (2 * sizeof(int) + 2 * sizeof(float)); }
  char* base; int pitch;
  int* halo(int rank) const { return (int*)(base + bytes(pitch) * rank); }                                  // [0,pitch): sent up, [pitch,2 pitch): sent down
  float* row(int rank, int which) const { return (float*)(halo(rank) + 2 * pitch) + (size_t)which * pitch; }   // 0 = first owned row, 1 = last
};

void worker(const MgpuJob& J, Shared* S, char* extra, int rank, int world) {
  int ndev = 0;
  MG_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev < 1) throw Fail{"no CUDA device"};
  const int dev = rank % ndev;
  MG_CUDA(cudaSetDevice(dev));
  cudaStream_t st;
  MG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));

  tdio::Raster in, wr;
  std::string err;
  if (!in.open(J.dirfile, &err)) throw Fail{"open " + std::string(J.dirfile) + ": " + err};
  const int nx = (int)in.width(), total_ny = (int)in.height();
  if (J.usew && !wr.open(J.wfile, &err)) throw Fail{"open " + std::string(J.wfile) + ": " + err};
  int row0, ny;
  partition(total_ny, world, rank, &row0, &ny);
  td_strip s;
  s.nx = nx; s.ny = ny; s.pitch = td_pitch_for(nx); s.has_top = rank > 0; s.has_bot = rank < world - 1;
  const size_t cells = (size_t)(ny + 2) * s.pitch;
  const tdio::DType dt = J.dinf ? tdio::DT_F32 : tdio::DT_I16;
  void* d_dir = nullptr; float *d_out = nullptr, *d_w = nullptr; int* d_halo = nullptr; double* d_dx = nullptr;
  MG_CUDA(cudaMalloc(&d_dir, cells * tdio::dtype_bytes(dt)));
  MG_CUDA(cudaMalloc(&d_out, cells * 4));
  MG_CUDA(cudaMalloc(&d_halo, sizeof(int) * 4 * (size_t)s.pitch));        // halo_out (2 pitch) + the received decrements (2 pitch)
  load_strip(in, dt, d_dir, nx, s.pitch, row0, ny, total_ny, st);
  if (J.usew) { MG_CUDA(cudaMalloc(&d_w, cells * 4)); load_strip(wr, tdio::DT_F32, d_w, nx, s.pitch, row0, ny, total_ny, st); }
  if (J.dinf) {
    std::vector<double> dxc, dyc;
    in.cell_sizes(&dxc, &dyc);
    MG_CUDA(cudaMalloc(&d_dx, sizeof(double) * 2 * (size_t)ny));
    MG_CUDA(cudaMemcpyAsync(d_dx, dxc.data() + row0, sizeof(double) * ny, cudaMemcpyHostToDevice, st));
    MG_CUDA(cudaMemcpyAsync(d_dx + ny, dyc.data() + row0, sizeof(double) * ny, cudaMemcpyHostToDevice, st));
    MG_CUDA(cudaStreamSynchronize(st));
  }
  td_ctx* ctx = td_ctx_create();
  if (!ctx) throw Fail{"td_ctx_create failed"};

  // ---- which way across the boundary: peer mode needs every pair of neighbours on two devices that reach each other
  S->device[rank] = dev;
  int can = world <= ndev ? 1 : 0;
  for (int nb = rank - 1; nb <= rank + 1 && can; nb += 2) {
    if (nb < 0 || nb >= world) continue;
    int ok = 0;
    MG_CUDA(cudaDeviceCanAccessPeer(&ok, dev, nb % ndev));
    if (!ok) can = 0;
  }
  S->can_peer[rank] = can;
  MG_BAR();
  bool peer = true;
  for (int r = 0; r < world; ++r) peer = peer && S->can_peer[r];
  if (const char* pe = getenv("TAUDEM_B200_PEER")) peer = atoi(pe) == 1;

  const double t0 = now();
  if (peer) {
    MG_TD(td_sweep_peer_export_dev(ctx, s, J.dinf, S->handles[rank], S->meta[rank], st));
    MG_BAR();
    if (rank > 0) MG_TD(td_sweep_peer_connect_dev(ctx, 0, S->handles[rank - 1], S->meta[rank - 1]));
    if (rank < world - 1) MG_TD(td_sweep_peer_connect_dev(ctx, 1, S->handles[rank + 1], S->meta[rank + 1]));
    MG_TD(td_sweep_peer_connect_dev(ctx, 2, rank == 0 ? nullptr : S->handles[0], nullptr));
  }
  if (J.dinf) {
    // the neighbour strips' edge rows keep their own cell sizes (geographic rasters: src/areadinf.cpp:199-201 getdxdyc(jn))
    std::vector<double> dxc, dyc;
    in.cell_sizes(&dxc, &dyc);
    td_set_halo_cell_sizes_dev(ctx, row0 > 0 ? dxc[row0 - 1] : 0., row0 > 0 ? dyc[row0 - 1] : 0., row0 + ny < total_ny ? dxc[row0 + ny] : 0.,
                               row0 + ny < total_ny ? dyc[row0 + ny] : 0.);
    MG_TD(td_area_deps_dev(ctx, (const float*)d_dir, d_out, s, (float)in.nodata(), d_dx, d_dx + ny, st));
  }
  else MG_TD(td_aread8_deps_dev(ctx, (const int16_t*)d_dir, d_out, s, (int16_t)in.nodata(), st));
  auto run = [&]() {
    MG_CUDA(cudaMemsetAsync(d_halo, 0, sizeof(int) * 2 * (size_t)s.pitch, st));
    if (J.dinf) MG_TD(td_area_sweep_run_dev(ctx, (const float*)d_dir, d_w, d_out, s, J.usew, J.contcheck, d_dx, d_halo, st));
    else MG_TD(td_aread8_sweep_run_dev(ctx, d_w, d_out, s, J.usew ? (float)wr.nodata() : 0.f, J.usew, J.contcheck, d_halo, st));
    MG_CUDA(cudaStreamSynchronize(st));
  };
  int rounds = 0;
  if (peer) {
    MG_TD(td_sweep_peer_begin_dev(ctx, s, st));
    MG_CUDA(cudaStreamSynchronize(st));
    MG_BAR();                                     // every strip is counted in the global counter before anybody can see it at zero
    run();
    MG_BAR();
    td_sweep_peer_off_dev(ctx);
    rounds = 1;
  } else {
    const RoundBuf R{extra, s.pitch};
    std::vector<int> hal(2 * (size_t)s.pitch), dec(2 * (size_t)s.pitch);
    MG_TD(td_sweep_begin_dev(ctx, s, st));
    for (;;) {
      run();
      ++rounds;
      // what I hand over: the decrement counts and my edge rows (DistTools.share + exchange_counts, src/aread8.cpp:283-297)
      MG_CUDA(cudaMemcpy(hal.data(), d_halo, sizeof(int) * 2 * (size_t)s.pitch, cudaMemcpyDeviceToHost));
      memcpy(R.halo(rank), hal.data(), sizeof(int) * 2 * (size_t)s.pitch);
      MG_CUDA(cudaMemcpy(R.row(rank, 0), d_out + (size_t)1 * s.pitch, sizeof(float) * s.pitch, cudaMemcpyDeviceToHost));
      MG_CUDA(cudaMemcpy(R.row(rank, 1), d_out + (size_t)ny * s.pitch, sizeof(float) * s.pitch, cudaMemcpyDeviceToHost));
      long long mine = 0;
      for (int v : hal) mine += v;
      std::atomic<long long>& total = S->handed[rounds & 1];
      total.fetch_add(mine);
      MG_BAR();
      const long long all = total.load();
      std::fill(dec.begin(), dec.end(), 0);
      if (rank > 0) {                                           // what the strip above sent down, and its last row
        memcpy(dec.data(), R.halo(rank - 1) + s.pitch, sizeof(int) * s.pitch);
        MG_CUDA(cudaMemcpy(d_out, R.row(rank - 1, 1), sizeof(float) * s.pitch, cudaMemcpyHostToDevice));
      }
      if (rank < world - 1) {                                   // what the strip below sent up, and its first row
        memcpy(dec.data() + s.pitch, R.halo(rank + 1), sizeof(int) * s.pitch);
        MG_CUDA(cudaMemcpy(d_out + (size_t)(ny + 1) * s.pitch, R.row(rank + 1, 0), sizeof(float) * s.pitch, cudaMemcpyHostToDevice));
      }
      S->handed[(rounds + 1) & 1].store(0);                      // the next round's total (nobody adds to it before the barrier below)
      MG_BAR();
      if (all == 0) break;                                      // ringTerm: nobody handed anything over
      MG_CUDA(cudaMemcpy(d_halo + 2 * (size_t)s.pitch, dec.data(), sizeof(int) * 2 * (size_t)s.pitch, cudaMemcpyHostToDevice));
      MG_TD(td_sweep_apply_halo_dev(ctx, s, rank > 0 ? d_halo + 2 * (size_t)s.pitch : nullptr,
                                    rank < world - 1 ? d_halo + 3 * (size_t)s.pitch : nullptr, st));
    }
  }
  S->secs[rank] = now() - t0;
  if (rank == 0) S->rounds = rounds;
  MG_CUDA(cudaMemcpy2D(J.out + (size_t)row0 * nx, (size_t)nx * 4, d_out + s.pitch, (size_t)s.pitch * 4, (size_t)nx * 4, (size_t)ny, cudaMemcpyDeviceToHost));
  td_ctx_destroy(ctx);
  cudaFree(d_dir); cudaFree(d_out); cudaFree(d_halo); cudaFree(d_w); cudaFree(d_dx);
}

// ---- pitremove / d8flowdir / dinfflowdir on row strips.  What the reference does with linearpart::share() and MPI_Allreduce
// (src/flood.cpp:344,401,468; src/d8.cpp:549-668) goes through the shared mapping: every rank has four row slots of 8 bytes per
// cell (its first / last owned row, its two halo rows) and a line of eight words for the sums.
struct RowSlots {
  static size_t bytes(int pitch) { return (size_t)pitch * 8 * 4; }
  char* base; int pitch;
  char* slot(int rank, int which) const { return base + bytes(pitch) * rank + (size_t)which * pitch * 8; }
};
struct StripComm {
  Shared* S; RowSlots R; int rank, world; td_strip s; cudaStream_t st;
  bool bar() const { return barrier(S, world); }
  // first / last owned row -> the halo rows of the strips above / below
  int share(void* arr, int eb) const {
    if (eb > 8) return 1;
    const size_t rb = (size_t)s.pitch * eb;
    char* a = (char*)arr;
    if (cudaMemcpyAsync(R.slot(rank, 0), a + rb, rb, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    if (cudaMemcpyAsync(R.slot(rank, 1), a + rb * (size_t)s.ny, rb, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    if (!bar()) return 1;
    if (rank > 0 && cudaMemcpyAsync(a, R.slot(rank - 1, 1), rb, cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    if (rank < world - 1 && cudaMemcpyAsync(a + rb * (size_t)(s.ny + 1), R.slot(rank + 1, 0), rb, cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    return bar() ? 0 : 1;
  }
  // the reverse: what the neighbours hold in their halo rows for my first / last row
  int collect(const void* arr, int eb, void* recv_top, void* recv_bot) const {
    if (eb > 8) return 1;
    const size_t rb = (size_t)s.pitch * eb;
    const char* a = (const char*)arr;
    if (cudaMemcpyAsync(R.slot(rank, 2), a, rb, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    if (cudaMemcpyAsync(R.slot(rank, 3), a + rb * (size_t)(s.ny + 1), rb, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    if (!bar()) return 1;
    if (rank > 0 && recv_top && cudaMemcpyAsync(recv_top, R.slot(rank - 1, 3), rb, cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    if (rank < world - 1 && recv_bot && cudaMemcpyAsync(recv_bot, R.slot(rank + 1, 2), rb, cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    return bar() ? 0 : 1;
  }
  int allreduce_sum(unsigned long long* v, int n) const {
    if (n > 8) return 1;
    for (int i = 0; i < n; ++i) S->red[rank][i] = v[i];
    if (!bar()) return 1;
    for (int i = 0; i < n; ++i) { unsigned long long t = 0; for (int r = 0; r < world; ++r) t += S->red[r][i]; v[i] = t; }
    return bar() ? 0 : 1;
  }
};
int cb_share(void* u, void* arr, int eb) { return ((const StripComm*)u)->share(arr, eb); }
int cb_collect(void* u, const void* arr, int eb, void* rt, void* rb) { return ((const StripComm*)u)->collect(arr, eb, rt, rb); }
int cb_allreduce(void* u, unsigned long long* v, int n) { return ((const StripComm*)u)->allreduce_sum(v, n); }

void flow_worker(const MgpuFlowJob& J, Shared* S, char* extra, int rank, int world) {
  int ndev = 0;
  MG_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev < 1) throw Fail{"no CUDA device"};
  MG_CUDA(cudaSetDevice(rank % ndev));
  cudaStream_t st;
  MG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  tdio::Raster in, mk;
  std::string err;
  if (!in.open(J.demfile, &err)) throw Fail{"open " + std::string(J.demfile) + ": " + err};
  const int nx = (int)in.width(), total_ny = (int)in.height();
  if (J.tool == 0 && J.use_mask && !mk.open(J.maskfile, &err)) throw Fail{"open " + std::string(J.maskfile) + ": " + err};
  int row0, ny;
  partition(total_ny, world, rank, &row0, &ny);
  td_strip s;
  s.nx = nx; s.ny = ny; s.pitch = td_pitch_for(nx); s.has_top = rank > 0; s.has_bot = rank < world - 1;
  const size_t cells = (size_t)(ny + 2) * s.pitch;
  float *d_z = nullptr, *d_f = nullptr, *d_slp = nullptr; void* d_dir = nullptr; int16_t* d_mask = nullptr; double* d_dx = nullptr;
  MG_CUDA(cudaMalloc(&d_z, cells * 4));
  load_strip(in, tdio::DT_F32, d_z, nx, s.pitch, row0, ny, total_ny, st);
  td_ctx* ctx = td_ctx_create();
  if (!ctx) throw Fail{"td_ctx_create failed"};
  const StripComm C{S, RowSlots{extra, s.pitch}, rank, world, s, st};
  td_strip_comm comm;
  comm.user = (void*)&C; comm.share = cb_share; comm.collect = cb_collect; comm.allreduce_sum = cb_allreduce;
  MG_BAR();
  const double t0 = now();
  int rounds = 0;
  long long left = 0;
  if (J.tool == 0) {
    // flood(): local relaxation to convergence, fresh halo rows, repeat until no strip moved (src/flood.cpp:344-479)
    if (J.use_mask) { MG_CUDA(cudaMalloc(&d_mask, cells * 2)); load_strip(mk, tdio::DT_I16, d_mask, nx, s.pitch, row0, ny, total_ny, st); }
    MG_CUDA(cudaMalloc(&d_f, cells * 4));
    MG_CUDA(cudaMemsetAsync(d_f, 0, cells * 4, st));
    MG_TD(td_flood_init_dev(ctx, d_z, d_mask, d_f, s, (float)in.nodata(), J.four, st));
    for (bool first = true;; first = false) {
      MG_CUDA(cudaStreamSynchronize(st));
      if (C.share(d_f, 4)) throw Fail{"row exchange failed"};
      int moved = 0;
      if (first) MG_TD(td_flood_relax_dev(ctx, d_z, d_f, s, J.four, &moved, st));
      else MG_TD(td_flood_relax_edges_dev(ctx, d_z, d_f, s, J.four, &moved, st));
      ++rounds;
      unsigned long long any = moved ? 1ull : 0ull;
      if (C.allreduce_sum(&any, 1)) throw Fail{"all-reduce failed"};
      if (any == 0) break;
    }
    MG_CUDA(cudaMemcpy2D((float*)J.out0 + (size_t)row0 * nx, (size_t)nx * 4, d_f + s.pitch, (size_t)s.pitch * 4, (size_t)nx * 4, (size_t)ny, cudaMemcpyDeviceToHost));
  } else {
    const bool dinf = J.tool == 2;
    std::vector<double> dxc, dyc;
    in.cell_sizes(&dxc, &dyc);
    MG_CUDA(cudaMalloc(&d_dx, sizeof(double) * 2 * (size_t)ny));
    MG_CUDA(cudaMemcpyAsync(d_dx, dxc.data() + row0, sizeof(double) * ny, cudaMemcpyHostToDevice, st));
    MG_CUDA(cudaMemcpyAsync(d_dx + ny, dyc.data() + row0, sizeof(double) * ny, cudaMemcpyHostToDevice, st));
    const size_t eb = dinf ? 4 : 2;
    MG_CUDA(cudaMalloc(&d_dir, cells * eb));
    MG_CUDA(cudaMalloc(&d_slp, cells * 4));
    MG_CUDA(cudaMemsetAsync(d_dir, 0, cells * eb, st));
    MG_CUDA(cudaMemsetAsync(d_slp, 0, cells * 4, st));
    MG_CUDA(cudaStreamSynchronize(st));
    long long nflat = 0;
    if (dinf) MG_TD(td_dinf_slopes_dev(ctx, d_z, (float*)d_dir, d_slp, s, (float)in.nodata(), d_dx, d_dx + ny, &nflat, st));
    else MG_TD(td_d8_slopes_dev(ctx, d_z, (int16_t*)d_dir, d_slp, s, (float)in.nodata(), d_dx, d_dx + ny, &nflat, st));
    MG_CUDA(cudaStreamSynchronize(st));
    unsigned long long total = (unsigned long long)nflat;
    if (C.allreduce_sum(&total, 1)) throw Fail{"all-reduce failed"};
    if (total) {
      // Garbrecht-Martz on the strips: the halo rows of the directions first, then the BFS passes with their exchanges
      if (C.share(d_dir, (int)eb)) throw Fail{"row exchange failed"};
      if (dinf) MG_TD(td_dinf_flats_strip_dev(ctx, d_z, (float*)d_dir, s, d_dx, d_dx + ny, &left, &comm, st));
      else MG_TD(td_d8_flats_strip_dev(ctx, d_z, (int16_t*)d_dir, s, d_dx, d_dx + ny, &left, &comm, st));
      MG_CUDA(cudaStreamSynchronize(st));
    }
    MG_CUDA(cudaMemcpy2D((char*)J.out0 + (size_t)row0 * nx * eb, (size_t)nx * eb, (char*)d_dir + (size_t)s.pitch * eb, (size_t)s.pitch * eb, (size_t)nx * eb, (size_t)ny,
                         cudaMemcpyDeviceToHost));
    MG_CUDA(cudaMemcpy2D(J.out1 + (size_t)row0 * nx, (size_t)nx * 4, d_slp + s.pitch, (size_t)s.pitch * 4, (size_t)nx * 4, (size_t)ny, cudaMemcpyDeviceToHost));
  }
  S->secs[rank] = now() - t0;
  if (rank == 0) { S->rounds = rounds; S->flats_left = left; }
  td_ctx_destroy(ctx);
  cudaFree(d_z); cudaFree(d_f); cudaFree(d_slp); cudaFree(d_dir); cudaFree(d_mask); cudaFree(d_dx);
}
}  // namespace

void* mgpu_alloc_shared(size_t bytes) {
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  return p == MAP_FAILED ? nullptr : p;
}
void mgpu_free_shared(void* p, size_t bytes) { if (p) munmap(p, bytes); }

int mgpu_world() {
  const char* e = getenv("TAUDEM_B200_GPUS");
  const int n = e ? atoi(e) : 1;
  return n < 1 ? 1 : (n > MAXR ? MAXR : n);
}

namespace {
// forks `world` ranks over a shared control block (+ extra bytes), waits for them, collects the timings
template <class Fn>
int run_ranks(const char* who, int world, size_t extra_bytes, Fn&& rank_fn, double* compute_seconds, int* rounds, long long* flats_left) {
  const size_t bytes = sizeof(Shared) + extra_bytes;
  char* mem = (char*)mgpu_alloc_shared(bytes);
  if (!mem) { set_error(std::string(who) + ": cannot map the shared control block"); return TD_ERR_IO; }
  Shared* S = new (mem) Shared();
  S->err.store(0); S->bar_count.store(0); S->bar_gen.store(0); S->handed[0].store(0); S->handed[1].store(0);
  fflush(stdout); fflush(stderr);
  std::vector<pid_t> pids(world, (pid_t)-1);
  for (int r = 0; r < world; ++r) {
    const pid_t pid = fork();
    if (pid < 0) { S->err.store(1); break; }
    if (pid == 0) {
      int code = 0;
      try { rank_fn(S, mem + sizeof(Shared), r); }
      catch (const Fail& f) { snprintf(S->msg[r], sizeof(S->msg[r]), "%s", f.what.c_str()); code = 1; }
      catch (const std::exception& e) { snprintf(S->msg[r], sizeof(S->msg[r]), "exception: %s", e.what()); code = 1; }
      if (code) S->err.store(1);
      fflush(stdout); fflush(stderr);
      _exit(code);                                  // no atexit handlers of the parent's image in the child
    }
    pids[r] = pid;
  }
  // the parent only waits: a rank that dies takes the others with it (they see err at their next barrier; ranks that are
  // stuck in a kernel waiting for the dead one are killed after a grace period)
  int left = 0, bad = 0;
  for (pid_t p : pids) if (p > 0) ++left;
  double t_err = 0.;
  while (left > 0) {
    bool any = false;
    for (int r = 0; r < world; ++r) {
      if (pids[r] <= 0) continue;
      int status = 0;
      const pid_t w = waitpid(pids[r], &status, WNOHANG);
      if (w == pids[r]) {
        any = true; pids[r] = -1; --left;
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) {
          ++bad; S->err.store(1);
          if (!S->msg[r][0]) snprintf(S->msg[r], sizeof(S->msg[r]), "rank ended abnormally (status 0x%x)", status);
        }
      }
    }
    if (S->err.load()) {
      if (t_err == 0.) t_err = now();
      else if (now() - t_err > 20.) { for (pid_t p : pids) if (p > 0) kill(p, SIGKILL); }
    }
    if (!any) usleep(2000);
  }
  int rc = TD_OK;
  if (bad || S->err.load()) {
    std::string m = "multi-GPU run failed:";
    for (int r = 0; r < world; ++r) if (S->msg[r][0] && strcmp(S->msg[r], "another rank failed") != 0) m += " [rank " + std::to_string(r) + "] " + S->msg[r];
    set_error(m);
    rc = TD_ERR_CUDA;
  } else {
    double mx = 0.;
    for (int r = 0; r < world; ++r) mx = std::max(mx, S->secs[r]);
    if (compute_seconds) *compute_seconds = mx;
    if (rounds) *rounds = S->rounds;
    if (flats_left) *flats_left = S->flats_left;
  }
  mgpu_free_shared(mem, bytes);
  return rc;
}
}  // namespace

int mgpu_area(const MgpuJob& J, int world, double* compute_seconds, int* rounds) {
  if (world < 2 || world > MAXR) { set_error("mgpu_area: between 2 and 64 ranks"); return TD_ERR_ARG; }
  if (J.ny < world) { set_error("mgpu_area: fewer rows than ranks"); return TD_ERR_ARG; }
  const int pitch = td_pitch_for(J.nx);
  return run_ranks("mgpu_area", world, RoundBuf::bytes(pitch) * (size_t)world,
                   [&](Shared* S, char* extra, int r) { worker(J, S, extra, r, world); }, compute_seconds, rounds, nullptr);
}

int mgpu_flow(const MgpuFlowJob& J, int world, double* compute_seconds, int* rounds, long long* flats_left) {
  if (world < 2 || world > MAXR) { set_error("mgpu_flow: between 2 and 64 ranks"); return TD_ERR_ARG; }
  if (J.ny < world) { set_error("mgpu_flow: fewer rows than ranks"); return TD_ERR_ARG; }
  if (J.tool < 0 || J.tool > 2 || !J.out0 || (J.tool > 0 && !J.out1)) { set_error("mgpu_flow: bad job"); return TD_ERR_ARG; }
  const int pitch = td_pitch_for(J.nx);
  return run_ranks("mgpu_flow", world, RowSlots::bytes(pitch) * (size_t)world,
                   [&](Shared* S, char* extra, int r) { flow_worker(J, S, extra, r, world); }, compute_seconds, rounds, flats_left);
}

}  // namespace td

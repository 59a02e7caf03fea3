// TEST INFRASTRUCTURE ONLY.  fork()+socketpair implementation of oracle/shim/mpi.h.
#include "mpi.h"
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <errno.h>
#include <sched.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include <vector>

namespace {
int g_rank = 0, g_size = 1;
std::vector<int> g_fd;          // g_fd[peer] = socket to peer
std::vector<pid_t> g_children;
struct Hdr { int tag; int bytes; };

int tsize(MPI_Datatype t) {
  switch (t) { case MPI_BYTE: return 1; case MPI_SHORT: case MPI_INT16_T: return 2;
    case MPI_INT: case MPI_FLOAT: case MPI_INT32_T: return 4; case MPI_LONG: case MPI_DOUBLE: return 8; }
  return 1;
}
void die(const char* m) { fprintf(stderr, "minimpi rank %d: %s (%s)\n", g_rank, m, strerror(errno)); _exit(99); }
void wr(int fd, const void* p, size_t n) {
  const char* c = (const char*)p;
  while (n) { ssize_t k = write(fd, c, n); if (k < 0) { if (errno == EINTR) continue; die("write"); } c += k; n -= k; }
}
void rd(int fd, void* p, size_t n) {
  char* c = (char*)p;
  while (n) { ssize_t k = read(fd, c, n); if (k < 0) { if (errno == EINTR) continue; die("read"); } if (k == 0) die("peer closed"); c += k; n -= k; }
}
void send_raw(int dest, int tag, const void* buf, int bytes) { Hdr h{tag, bytes}; wr(g_fd[dest], &h, sizeof h); if (bytes) wr(g_fd[dest], buf, bytes); }
void recv_raw(int src, int tag, void* buf, int maxbytes, int* got) {
  Hdr h; rd(g_fd[src], &h, sizeof h);
  if (h.tag != tag) { fprintf(stderr, "minimpi rank %d: tag mismatch from %d: got %d want %d\n", g_rank, src, h.tag, tag); _exit(98); }
  if (h.bytes > maxbytes) die("message too long");
  if (h.bytes) rd(g_fd[src], buf, h.bytes);
  if (got) *got = h.bytes;
}
}  // namespace

extern "C" {
int MPI_Init(int*, char***) {
  const char* e = getenv("MINIMPI_NP");
  g_size = e ? atoi(e) : 1;
  if (g_size < 1) g_size = 1;
  g_rank = 0;
  if (g_size == 1) return 0;
  fflush(stdout); fflush(stderr);
  // one socketpair per unordered rank pair
  std::vector<std::vector<int>> sv(g_size, std::vector<int>(g_size, -1));
  for (int a = 0; a < g_size; a++)
    for (int b = a + 1; b < g_size; b++) {
      int p[2];
      if (socketpair(AF_UNIX, SOCK_STREAM, 0, p) != 0) die("socketpair");
      int sz = 8 << 20;
      for (int k = 0; k < 2; k++) { setsockopt(p[k], SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz); setsockopt(p[k], SOL_SOCKET, SO_RCVBUF, &sz, sizeof sz); }
      sv[a][b] = p[0]; sv[b][a] = p[1];
    }
  for (int r = 1; r < g_size; r++) {
    pid_t pid = fork();
    if (pid < 0) die("fork");
    if (pid == 0) { g_rank = r; g_children.clear(); break; }
    g_children.push_back(pid);
  }
  // MINIMPI_PIN=1: rank r runs on the r-th CPU of the process's affinity mask (stable timings between hosts: no rank migration)
  if (const char* pin = getenv("MINIMPI_PIN")) {
    if (atoi(pin) > 0) {
      cpu_set_t have, want;
      if (sched_getaffinity(0, sizeof have, &have) == 0) {
        const int n = CPU_COUNT(&have);
        int k = n > 0 ? g_rank % n : 0;
        for (int c = 0; c < CPU_SETSIZE; c++)
          if (CPU_ISSET(c, &have) && k-- == 0) { CPU_ZERO(&want); CPU_SET(c, &want); sched_setaffinity(0, sizeof want, &want); break; }
      }
    }
  }
  g_fd.assign(g_size, -1);
  for (int a = 0; a < g_size; a++)
    for (int b = 0; b < g_size; b++) {
      if (a == b || sv[a][b] < 0) continue;
      if (a == g_rank) g_fd[b] = sv[a][b]; else close(sv[a][b]);
    }
  return 0;
}
int MPI_Finalize(void) {
  fflush(stdout); fflush(stderr);
  if (g_size > 1) {
    if (g_rank == 0) { for (pid_t p : g_children) { int st; waitpid(p, &st, 0); } }
    else _exit(0);     // children end here; rank 0 carries the process exit status
  }
  return 0;
}
int MPI_Comm_rank(MPI_Comm, int* r) { *r = g_rank; return 0; }
int MPI_Comm_size(MPI_Comm, int* s) { *s = g_size; return 0; }
double MPI_Wtime(void) { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int MPI_Abort(MPI_Comm, int code) {
  fflush(stdout); fflush(stderr);
  if (g_size > 1) kill(0, SIGTERM);
  _exit(code & 0xff ? code & 0xff : 1);
}
int MPI_Send(const void* b, int n, MPI_Datatype t, int dest, int tag, MPI_Comm) { send_raw(dest, tag, b, n * tsize(t)); return 0; }
int MPI_Bsend(const void* b, int n, MPI_Datatype t, int dest, int tag, MPI_Comm) { send_raw(dest, tag, b, n * tsize(t)); return 0; }
int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Status* st) {
  int got = 0; recv_raw(src, tag, b, n * tsize(t), &got);
  if (st) { st->MPI_SOURCE = src; st->MPI_TAG = tag; st->MPI_ERROR = 0; st->count_bytes = got; }
  return 0;
}
int MPI_Probe(int src, int tag, MPI_Comm, MPI_Status* st) {
  Hdr h; size_t have = 0;
  while (have < sizeof h) {
    ssize_t k = recv(g_fd[src], &h, sizeof h, MSG_PEEK | MSG_WAITALL);
    if (k < 0) { if (errno == EINTR) continue; die("probe"); }
    if (k == 0) die("peer closed in probe");
    have = (size_t)k;
  }
  if (h.tag != tag) die("probe tag mismatch");
  if (st) { st->MPI_SOURCE = src; st->MPI_TAG = tag; st->MPI_ERROR = 0; st->count_bytes = h.bytes; }
  return 0;
}
int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* c) { *c = st->count_bytes / tsize(t); return 0; }
int MPI_Buffer_attach(void*, int) { return 0; }
int MPI_Buffer_detach(void* p, int* n) { if (p) *(void**)p = NULL; if (n) *n = 0; return 0; }
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm) {
  if (g_size == 1) return 0;
  const int bytes = n * tsize(t);
  if (g_rank == root) { for (int r = 0; r < g_size; r++) if (r != root) send_raw(r, 888, b, bytes); }
  else recv_raw(root, 888, b, bytes, NULL);
  return 0;
}
int MPI_Allreduce(const void* in, void* out, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  const int bytes = n * tsize(t);
  if (in != out) memcpy(out, in, bytes);
  if (g_size == 1) return 0;
  if (g_rank == 0) {
    std::vector<char> tmp(bytes);
    for (int r = 1; r < g_size; r++) {
      recv_raw(r, 777, tmp.data(), bytes, NULL);
      for (int i = 0; i < n; i++) {
        switch (t) {
          case MPI_LONG: ((long*)out)[i] += ((long*)tmp.data())[i]; break;
          case MPI_INT: case MPI_INT32_T: ((int*)out)[i] += ((int*)tmp.data())[i]; break;
          case MPI_DOUBLE: ((double*)out)[i] += ((double*)tmp.data())[i]; break;
          case MPI_FLOAT: ((float*)out)[i] += ((float*)tmp.data())[i]; break;
          case MPI_SHORT: case MPI_INT16_T: ((short*)out)[i] += ((short*)tmp.data())[i]; break;
          default: die("allreduce type");
        }
      }
    }
    for (int r = 1; r < g_size; r++) send_raw(r, 778, out, bytes);
  } else {
    send_raw(0, 777, out, bytes);
    recv_raw(0, 778, out, bytes, NULL);
  }
  return 0;
}
int MPI_Barrier(MPI_Comm c) { int a = 0, b = 0; return MPI_Allreduce(&a, &b, 1, MPI_INT, MPI_SUM, c); }
}

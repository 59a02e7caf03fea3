/* Discrete-event model of the tile-dataflow sweep (taudem_b200/csrc/sweep_tiles.cu) on a real flow
 * field: how many tile visits a scheduling policy / tile shape needs and how long the schedule takes
 * under a simple per-visit cost model.  Research tooling for choosing the kernel's tile geometry and
 * queue policy without a GPU; nothing in the product or the tests depends on it.
 *
 *   gcc -O2 -shared -fPIC -o libsweepsim.so sim.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int nx, ny, twx, twy, ntx, nty;
  const int32_t* d1;   /* downslope cell 1 (or -1) */
  const int32_t* d2;   /* downslope cell 2 (or -1), D-infinity only */
  uint8_t* cnt;        /* remaining arrivals; 0xFF = done */
  int32_t* depth;      /* scratch: in-visit chain depth */
} Grid;

static inline int tile_of(const Grid* g, int c) {
  const int r = c / g->nx, x = c - r * g->nx;
  return (r / g->twy) * g->ntx + x / g->twx;
}

/* one visit: evaluate everything that is ready inside the tile; crossings are appended to out[] */
static int visit(Grid* g, int t, int32_t* stack, int32_t* out, int* nout, int* maxdepth) {
  const int tx = t % g->ntx, ty = t / g->ntx;
  const int c0 = tx * g->twx, r0 = ty * g->twy;
  const int c1 = c0 + g->twx < g->nx ? c0 + g->twx : g->nx, r1 = r0 + g->twy < g->ny ? r0 + g->twy : g->ny;
  int sp = 0, done = 0, md = 0;
  *nout = 0;
  for (int r = r0; r < r1; ++r)
    for (int x = c0; x < c1; ++x) {
      const int c = r * g->nx + x;
      if (g->cnt[c] == 0) { stack[sp++] = c; g->depth[c] = 1; }
    }
  while (sp) {
    const int c = stack[--sp];
    g->cnt[c] = 0xFF;
    ++done;
    const int dc = g->depth[c];
    if (dc > md) md = dc;
    for (int j = 0; j < 2; ++j) {
      const int d = j == 0 ? g->d1[c] : (g->d2 ? g->d2[c] : -1);
      if (d < 0 || g->cnt[d] == 0xFF) continue;
      const int dr = d / g->nx, dx = d - dr * g->nx;
      if (dr >= r0 && dr < r1 && dx >= c0 && dx < c1) {
        if (--g->cnt[d] == 0) { g->depth[d] = dc + 1; stack[sp++] = d; }   /* chain depth through the last arrival */
      } else out[(*nout)++] = d;
    }
  }
  *maxdepth = md;
  return done;
}

/* binary heap of (key, seq, tile) */
typedef struct { double key; long seq; int tile; } HItem;
typedef struct { HItem* a; int n, cap; } Heap;
static int hless(const HItem* x, const HItem* y) { return x->key < y->key || (x->key == y->key && x->seq < y->seq); }
static void hpush(Heap* h, HItem it) {
  if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 1024; h->a = (HItem*)realloc(h->a, sizeof(HItem) * h->cap); }
  int i = h->n++;
  while (i > 0) { int p = (i - 1) / 2; if (!hless(&it, &h->a[p])) break; h->a[i] = h->a[p]; i = p; }
  h->a[i] = it;
}
static HItem hpop(Heap* h) {
  HItem top = h->a[0], it = h->a[--h->n];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    const HItem* best = &it;
    if (l < h->n && hless(&h->a[l], best)) { m = l; best = &h->a[l]; }
    if (r < h->n && hless(&h->a[r], best)) { m = r; }
    if (m == i) break;
    h->a[i] = h->a[m]; i = m;
  }
  h->a[i] = it;
  return top;
}

/* policy: 0 = FIFO (all tiles queued in raster order first: the round-1 kernel),
 *         1 = priority by tilekey[] (smaller first; re-activations keep the tile's key),
 *         2 = FIFO for the first pass, re-activations to a priority heap served first,
 *         3 = LIFO for re-activations (newest first), first pass FIFO behind them,
 *         4 = every tile exactly once in raster order, no re-activation (phase A of a two-phase sweep);
 *             cnt_out then receives the remaining counts (0xFF = evaluated).
 * lazy: a re-activated tile is only queued once `lazy` arrivals are pending or nothing else is queued (0 = off)
 * cost of a visit (us) = c_fixed + c_cell * cells evaluated + c_hop * longest in-visit chain
 * result[0] visits, [1] makespan us, [2] sum of visit costs us, [3] cells evaluated, [4] max visits of one tile,
 *       [5] visits that evaluated nothing */
int sim_run(int nx, int ny, const int32_t* d1, const int32_t* d2, const uint8_t* cnt0, int twx, int twy, int workers,
            int policy, const double* tilekey, double c_fixed, double c_cell, double c_hop, double* result,
            int32_t* visits_per_tile, uint8_t* cnt_out) {
  Grid g;
  g.nx = nx; g.ny = ny; g.twx = twx; g.twy = twy; g.ntx = (nx + twx - 1) / twx; g.nty = (ny + twy - 1) / twy;
  g.d1 = d1; g.d2 = d2;
  const long n = (long)nx * ny;
  const int nt = g.ntx * g.nty;
  g.cnt = (uint8_t*)malloc(n); memcpy(g.cnt, cnt0, n);
  g.depth = (int32_t*)calloc(n, 4);
  int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * twx * twy);
  int32_t* out = (int32_t*)malloc(sizeof(int32_t) * twx * twy * 2);
  uint8_t* state = (uint8_t*)calloc(nt, 1);   /* 0 idle 1 queued 2 running 3 running+dirty */
  Heap q = {0, 0, 0}, ev = {0, 0, 0};         /* q: ready tiles; ev: running visits keyed by finish time */
  long seq = 0;
  /* deferred deliveries of running visits: applied when the visit finishes */
  int32_t** pend = (int32_t**)calloc(nt, sizeof(int32_t*));
  int* npend = (int*)calloc(nt, sizeof(int));
  memset(visits_per_tile, 0, sizeof(int32_t) * nt);
  for (int t = 0; t < nt; ++t) {
    HItem it; it.tile = t; it.seq = seq++;
    it.key = policy == 1 ? tilekey[t] : (policy == 0 || policy == 4 ? 0.0 : 1.0);   /* policies 2,3: first pass behind re-activations */
    hpush(&q, it); state[t] = 1;
  }
  double now = 0, busy = 0;
  long visits = 0, cells = 0, empty = 0;
  int running = 0;
  for (;;) {
    while (running < workers && q.n > 0) {
      HItem it = hpop(&q);
      const int t = it.tile;
      state[t] = 2;
      int no, md;
      const int done = visit(&g, t, stack, out, &no, &md);
      pend[t] = (int32_t*)malloc(sizeof(int32_t) * (no + 1));
      memcpy(pend[t], out, sizeof(int32_t) * no); npend[t] = no;
      const double cost = c_fixed + c_cell * done + c_hop * md;
      busy += cost; ++visits; cells += done; ++visits_per_tile[t];
      if (done == 0) ++empty;
      HItem e; e.key = now + cost; e.seq = seq++; e.tile = t;
      hpush(&ev, e); ++running;
    }
    if (ev.n == 0) break;
    HItem e = hpop(&ev);
    now = e.key; --running;
    const int t = e.tile;
    int self = 0;
    for (int i = 0; i < npend[t]; ++i) {
      const int d = pend[t][i];
      if (g.cnt[d] == 0xFF || g.cnt[d] == 0) continue;
      if (--g.cnt[d] == 0) {
        const int u = tile_of(&g, d);
        if (u == t) { self = 1; continue; }
        if (policy == 4) continue;
        if (state[u] == 0) {
          state[u] = 1;
          HItem it; it.tile = u; it.seq = policy == 3 ? -(seq++) : seq++;
          it.key = policy == 1 ? tilekey[u] : 0.0;
          hpush(&q, it);
        } else if (state[u] == 2) state[u] = 3;
      }
    }
    free(pend[t]); pend[t] = 0; npend[t] = 0;
    if (policy != 4 && (state[t] == 3 || self)) {
      state[t] = 1;
      HItem it; it.tile = t; it.seq = policy == 3 ? -(seq++) : seq++;
      it.key = policy == 1 ? tilekey[t] : 0.0;
      hpush(&q, it);
    } else state[t] = 0;
  }
  int mv = 0;
  for (int t = 0; t < nt; ++t) if (visits_per_tile[t] > mv) mv = visits_per_tile[t];
  long left = 0;
  for (long c = 0; c < n; ++c) if (g.cnt[c] != 0xFF && cnt0[c] != 0xFE) ++left;
  result[0] = (double)visits; result[1] = now; result[2] = busy; result[3] = (double)cells; result[4] = mv; result[5] = (double)empty;
  result[6] = (double)left;
  if (cnt_out) memcpy(cnt_out, g.cnt, n);
  free(g.cnt); free(g.depth); free(stack); free(out); free(state); free(q.a); free(ev.a); free(pend); free(npend);
  return 0;
}

/* Longest dependency chain (in cells).  weight[c] = 0 excludes a cell from the count (e.g. cells a first,
 * purely tile-local pass has already evaluated); returns the maximum and fills depth[]. */
long sim_longest(int nx, int ny, const int32_t* d1, const int32_t* d2, const uint8_t* cnt0, const uint8_t* weight, int32_t* depth) {
  const long n = (long)nx * ny;
  uint8_t* cnt = (uint8_t*)malloc(n); memcpy(cnt, cnt0, n);
  int32_t* q = (int32_t*)malloc(sizeof(int32_t) * n);
  long head = 0, tail = 0, best = 0;
  for (long c = 0; c < n; ++c) { depth[c] = 0; if (cnt[c] == 0) q[tail++] = (int32_t)c; }
  while (head < tail) {
    const int c = q[head++];
    depth[c] += weight ? weight[c] : 1;
    if (depth[c] > best) best = depth[c];
    for (int j = 0; j < 2; ++j) {
      const int d = j == 0 ? d1[c] : (d2 ? d2[c] : -1);
      if (d < 0) continue;
      if (depth[d] < depth[c]) depth[d] = depth[c];
      if (--cnt[d] == 0) q[tail++] = d;
    }
  }
  free(cnt); free(q);
  return best;
}

/* marks the cells a purely tile-local first pass evaluates (every contributor, transitively, lies in the same tile) */
long sim_local(int nx, int ny, const int32_t* d1, const int32_t* d2, const uint8_t* cnt0, int twx, int twy, uint8_t* local) {
  const long n = (long)nx * ny;
  const int ntx = (nx + twx - 1) / twx;
  uint8_t* cnt = (uint8_t*)malloc(n); memcpy(cnt, cnt0, n);
  int32_t* q = (int32_t*)malloc(sizeof(int32_t) * n);
  long head = 0, tail = 0;
  for (long c = 0; c < n; ++c) { local[c] = 0; if (cnt[c] == 0) q[tail++] = (int32_t)c; }
  while (head < tail) {
    const int c = q[head++];
    local[c] = 1;
    const int tc = ((c / nx) / twy) * ntx + (c % nx) / twx;
    for (int j = 0; j < 2; ++j) {
      const int d = j == 0 ? d1[c] : (d2 ? d2[c] : -1);
      if (d < 0) continue;
      const int td = ((d / nx) / twy) * ntx + (d % nx) / twx;
      if (td != tc) { cnt[d] = 0x7f; continue; }           /* a remote contributor: never local */
      if (cnt[d] < 0x40 && --cnt[d] == 0) q[tail++] = d;
    }
  }
  free(cnt); free(q);
  return tail;
}

/* Model of the streaming level passes (k_level): a pass visits the rows in bands of `band` rows (what runs
 * concurrently on the GPU); a cell is evaluated in a pass if it is ready when its band starts.  dir = +1 forward
 * raster order, -1 reverse, 0 alternate per pass.  done_after[p] = cells evaluated after pass p+1. */
int sim_levelpasses(int nx, int ny, const int32_t* d1, const int32_t* d2, const uint8_t* cnt0, int band, int passes, int dir,
                    int64_t* done_after, uint8_t* done_out) {
  const long n = (long)nx * ny;
  uint8_t* cnt = (uint8_t*)malloc(n); memcpy(cnt, cnt0, n);
  uint8_t* done = (uint8_t*)calloc(n, 1);
  int32_t* ready = (int32_t*)malloc(sizeof(int32_t) * (size_t)nx * band);
  long total = 0;
  const int nb = (ny + band - 1) / band;
  for (int p = 0; p < passes; ++p) {
    const int rev = dir < 0 || (dir == 0 && (p & 1));
    for (int bi = 0; bi < nb; ++bi) {
      const int b = rev ? nb - 1 - bi : bi;
      const int r0 = b * band, r1 = r0 + band < ny ? r0 + band : ny;
      long nr = 0;
      for (long c = (long)r0 * nx; c < (long)r1 * nx; ++c) if (!done[c] && cnt[c] == 0) ready[nr++] = (int32_t)c;
      for (long i = 0; i < nr; ++i) {
        const long c = ready[i];
        done[c] = 1; ++total;
        if (d1[c] >= 0) --cnt[d1[c]];
        if (d2 && d2[c] >= 0) --cnt[d2[c]];
      }
    }
    done_after[p] = total;
  }
  if (done_out) memcpy(done_out, done, n);
  free(cnt); free(done); free(ready);
  return 0;
}

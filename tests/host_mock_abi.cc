// Recording stand-in for the C ABI (include/hemelb_b200.h), for the CPU test of the C++ host-side
// drop-in's call sequence (tests/test_host_lbm.py).  Every entry point the host headers use writes
// one line to $HLB_MOCK_LOG and succeeds; nothing is computed.  Test infrastructure only -- the
// product library has no CPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "hemelb_b200.h"

struct hlb_gpu_handle { hlb_gpu_config cfg; };

namespace {
  // one log per process; or, for harnesses whose ranks are threads (tests/host_lbm_real.cc), one per thread:
  // $HLB_MOCK_LOG.rank<r> once the thread has said which rank it is
  thread_local int mock_rank = -1;
  thread_local FILE* mock_fh = nullptr;
  FILE* out() {
    if (mock_rank >= 0) {
      if (!mock_fh) {
        const char* p = getenv("HLB_MOCK_LOG");
        mock_fh = p ? fopen((std::string(p) + ".rank" + std::to_string(mock_rank)).c_str(), "w") : stderr;
      }
      return mock_fh;
    }
    static FILE* fh = nullptr;
    if (!fh) {
      const char* p = getenv("HLB_MOCK_LOG");
      fh = p ? fopen(p, "w") : stderr;
    }
    return fh;
  }
  long long ll(int64_t v) { return (long long)v; }
}

extern "C" {
void hlb_mock_set_rank(int r) {
  if (mock_fh && mock_fh != stderr) fclose(mock_fh);
  mock_fh = nullptr;
  mock_rank = r;
}
const char* hlb_gpu_last_error(void) { return "mock"; }
int hlb_gpu_device_count(int* n) { *n = 1; return 0; }
int hlb_gpu_create(const hlb_gpu_config* c, hlb_gpu_t* h) {
  *h = new hlb_gpu_handle{*c};
  fprintf(out(), "create lattice=%d kernel=%d wall=%d inlet=%d outlet=%d tau=%.17g rank=%d nranks=%d n_sites=%lld "
          "mid=%lld,%lld,%lld,%lld,%lld,%lld edge=%lld,%lld,%lld,%lld,%lld,%lld shared=%lld neighbours=%d inlets=%d outlets=%d reorder=%d\n",
          c->lattice, c->kernel, c->wall, c->inlet, c->outlet, c->tau, c->rank, c->nranks, ll(c->n_sites),
          ll(c->mid_count[0]), ll(c->mid_count[1]), ll(c->mid_count[2]), ll(c->mid_count[3]), ll(c->mid_count[4]), ll(c->mid_count[5]),
          ll(c->edge_count[0]), ll(c->edge_count[1]), ll(c->edge_count[2]), ll(c->edge_count[3]), ll(c->edge_count[4]), ll(c->edge_count[5]),
          ll(c->total_shared_fs), c->n_neighbours, c->n_inlets, c->n_outlets, c->reorder);
  return 0;
}
int hlb_gpu_destroy(hlb_gpu_t h) { fprintf(out(), "destroy\n"); fflush(out()); delete h; return 0; }
int hlb_gpu_set_neighbour_indices(hlb_gpu_t, int64_t a, int64_t n, const int64_t* idx) {
  fprintf(out(), "set_neighbour_indices %lld %lld first=%lld\n", ll(a), ll(n), ll(idx[0])); return 0; }
int hlb_gpu_set_site_data(hlb_gpu_t, int64_t a, int64_t n, const uint32_t* w, const uint32_t* i, const int32_t*) {
  fprintf(out(), "set_site_data %lld %lld wall0=%u iolet0=%u\n", ll(a), ll(n), w[0], i[0]); return 0; }
int hlb_gpu_set_wall_distances(hlb_gpu_t, int64_t a, int64_t n, const double*) { fprintf(out(), "set_wall_distances %lld %lld\n", ll(a), ll(n)); return 0; }
int hlb_gpu_set_wall_normals(hlb_gpu_t, int64_t a, int64_t n, const double*) { fprintf(out(), "set_wall_normals %lld %lld\n", ll(a), ll(n)); return 0; }
int hlb_gpu_set_site_coords(hlb_gpu_t, int64_t a, int64_t n, const int64_t*) { fprintf(out(), "set_site_coords %lld %lld\n", ll(a), ll(n)); return 0; }
int hlb_gpu_set_neighbours(hlb_gpu_t h, const int* r, const int64_t* c, const int64_t* f) {
  fprintf(out(), "set_neighbours");
  for (int i = 0; i < h->cfg.n_neighbours; ++i) fprintf(out(), " %d:%lld:%lld", r[i], ll(c[i]), ll(f[i]));
  fprintf(out(), "\n");
  return 0;
}
int hlb_gpu_set_streaming_indices(hlb_gpu_t h, const int64_t* idx) {
  long long sum = 0;
  for (int64_t i = 0; i < h->cfg.total_shared_fs; ++i) sum += (i + 1) * idx[i];
  fprintf(out(), "set_streaming_indices weighted_sum=%lld\n", sum);
  return 0;
}
int hlb_gpu_set_iolets(hlb_gpu_t, int which, int n, const double* r) {
  fprintf(out(), "set_iolets %d %d kind0=%d min_density=%.17g warmup0=%g\n", which, n, (int)r[0], r[14], r[13]); return 0; }
int hlb_gpu_set_gzs_remote(hlb_gpu_t, int64_t n, const int64_t* site, const int32_t* dir, const int32_t* owner, const int64_t* key) {
  fprintf(out(), "set_gzs_remote %lld", ll(n));
  for (int64_t k = 0; k < n; ++k) fprintf(out(), " %lld:%d:%d:%lld", ll(site[k]), dir[k], owner[k], ll(key[k]));
  fprintf(out(), "\n");
  return 0;
}
int hlb_gpu_set_gzs_serve(hlb_gpu_t, int64_t n, const int32_t* rank, const int64_t* site) {
  fprintf(out(), "set_gzs_serve %lld", ll(n));
  for (int64_t k = 0; k < n; ++k) fprintf(out(), " %d:%lld", rank[k], ll(site[k]));
  fprintf(out(), "\n");
  return 0;
}
int hlb_gpu_exchange_site_halo(hlb_gpu_t) { fprintf(out(), "exchange_site_halo\n"); return 0; }
int hlb_gpu_finalise(hlb_gpu_t) { fprintf(out(), "finalise\n"); return 0; }
int hlb_gpu_comm_unique_id(void* id) { memset(id, 0, 128); return 0; }
int hlb_gpu_comm_init(hlb_gpu_t, const void*) { fprintf(out(), "comm_init\n"); return 0; }
int hlb_gpu_set_f(hlb_gpu_t, int which, const double* f) { fprintf(out(), "set_f %d f0=%.17g\n", which, f[0]); return 0; }
int hlb_gpu_get_f(hlb_gpu_t h, int which, double* f) {
  fprintf(out(), "get_f %d\n", which);
  memset(f, 0, sizeof(double) * (h->cfg.n_sites * h->cfg.lattice + 1 + h->cfg.total_shared_fs));
  return 0;
}
int hlb_gpu_request_comms(hlb_gpu_t) { fprintf(out(), "request_comms\n"); return 0; }
int hlb_gpu_copy_received(hlb_gpu_t) { fprintf(out(), "copy_received\n"); return 0; }
int hlb_gpu_swap(hlb_gpu_t) { fprintf(out(), "swap\n"); return 0; }
int hlb_gpu_sync(hlb_gpu_t) { fprintf(out(), "sync\n"); return 0; }
int hlb_gpu_set_step_scalars(hlb_gpu_t h, uint64_t t, const double* in, const double* o, uint32_t mask) {
  fprintf(out(), "set_step_scalars t=%llu mask=%u", (unsigned long long)t, mask);
  for (int i = 0; i < h->cfg.n_inlets; ++i) fprintf(out(), " in%d=%.17g", i, in[i]);
  for (int i = 0; i < h->cfg.n_outlets; ++i) fprintf(out(), " out%d=%.17g", i, o[i]);
  fprintf(out(), "\n");
  return 0;
}
int hlb_gpu_stream_and_collide(hlb_gpu_t, int slot, int64_t a, int64_t n) { fprintf(out(), "stream_and_collide %d %lld %lld\n", slot, ll(a), ll(n)); return 0; }
int hlb_gpu_post_step(hlb_gpu_t, int slot, int64_t a, int64_t n) { fprintf(out(), "post_step %d %lld %lld\n", slot, ll(a), ll(n)); return 0; }
int hlb_gpu_stability(hlb_gpu_t, int conv, double* out2) {
  fprintf(out(), "stability convergence=%d\n", conv);
  out2[0] = 0.0;
  out2[1] = conv ? 1e-3 : 0.0;
  return 0;
}
int hlb_gpu_monitor(hlb_gpu_t, double* out4) {
  static int calls = 0;  // one process per rank: a density range that widens with every call, as a run's extrema would
  ++calls;
  fprintf(out(), "monitor\n");
  // ranks as threads: extrema that tell the ranks apart, so that a test can see whose reached the root
  const int k = mock_rank >= 0 ? mock_rank + 1 : calls;
  out4[0] = 0.01;
  out4[1] = 1.0 - 0.001 * k;
  out4[2] = 1.0 + 0.002 * k;
  out4[3] = 0.003 * k;
  return 0;
}
int hlb_gpu_edge_done(hlb_gpu_t) { fprintf(out(), "edge_done\n"); return 0; }
int hlb_gpu_get_cache(hlb_gpu_t h, uint32_t which, double* o) {
  fprintf(out(), "get_cache %u\n", which);
  const int w = which == HLB_CACHE_STRESS_TENSOR ? 9 : (which & (HLB_CACHE_VELOCITY | HLB_CACHE_TRACTION | HLB_CACHE_TANGENTIAL_TRACTION) ? 3 : 1);
  for (int64_t i = 0; i < h->cfg.n_sites * w; ++i) o[i] = (double)which;
  return 0;
}
}

// ---- extraction entry points used by extraction/GpuPropertyEncoder.h
struct hlb_xtr_handle { int n_fields; };
extern "C" {
int hlb_xtr_create(hlb_gpu_t, const hlb_xtr_spec* s, const int64_t* c, hlb_xtr_t* x) {
  fprintf(out(), "xtr_create selector=%d params=%.9g,%.9g,%.9g,%.9g,%.9g,%.9g,%.9g units=%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g coords0=%lld,%lld,%lld",
          s->selector, s->selector_params[0], s->selector_params[1], s->selector_params[2], s->selector_params[3],
          s->selector_params[4], s->selector_params[5], s->selector_params[6], s->time_step, s->voxel_size, s->origin[0],
          s->origin[1], s->origin[2], s->fluid_density, s->reference_pressure, ll(c[0]), ll(c[1]), ll(c[2]));
  for (int i = 0; i < s->n_fields; ++i) {
    const hlb_xtr_field& f = s->fields[i];
    fprintf(out(), " field=%s:%d:%d:%u", f.name, f.source, f.typecode, f.n_offsets);
    for (uint32_t k = 0; k < f.n_offsets; ++k) fprintf(out(), ":%.17g", f.offsets[k]);
  }
  fprintf(out(), "\n");
  *x = new hlb_xtr_handle{s->n_fields};
  return 0;
}
int hlb_xtr_destroy(hlb_xtr_t x) { fprintf(out(), "xtr_destroy\n"); fflush(out()); delete x; return 0; }
int hlb_xtr_sizes(hlb_xtr_t x, uint64_t* n, uint64_t* len, uint64_t* head) { *n = 3; *len = 8 * x->n_fields; *head = 60 + 4 * x->n_fields; return 0; }
int hlb_xtr_required_caches(hlb_xtr_t x, uint32_t* m) { *m = 100 + x->n_fields; return 0; }
int hlb_xtr_header(hlb_xtr_t x, uint64_t global, void* buf, uint64_t cap) {
  fprintf(out(), "xtr_header global=%llu capacity=%llu\n", (unsigned long long)global, (unsigned long long)cap);
  memset(buf, 'H', cap);
  return 0;
}
int hlb_xtr_encode(hlb_xtr_t, uint64_t first, uint64_t n, void* buf, uint64_t cap) {
  fprintf(out(), "xtr_encode first=%llu n=%llu capacity=%llu\n", (unsigned long long)first, (unsigned long long)n, (unsigned long long)cap);
  memset(buf, 'R', cap);
  return 0;
}
}

// METIS-free k-way partition of the fluid-site graph -- host code of the C ABI (hlb_part_*).
//
// Stands where the reference calls ParMETIS_V3_PartKway (Code/geometry/decomposition/
// OptimisedDecomposition.cc:138-154): same inputs -- the CSR site graph of PopulateAdjacencyData
// (:311-379), vertex weights by collision type (DecompositionWeights.h.in:25-62), the number of
// parts, the balance tolerance ubvec (:133) -- and the same output, a part per vertex, refined from
// the partition the vertices arrive with (the reference's BasicDecomposition, or hlb_part_bisect).
// ParMETIS (4.0.2, dependencies/ParMETIS/build.cmake:7; a third-party dependency whose source is not
// under the reference tree) is not in this image and no reference test pins a partition: parity is
// unpinned by design.  What is kept of its published scheme (Karypis & Kumar's parallel multilevel
// k-way) is the refinement half: balance first, by diffusion over the part graph, then greedy
// boundary moves to the best-connected part in alternating directions of part index; the numpy statement of the same algorithm (hemelb_b200/partition.py) is what the
// tests compare with, move for move.
//
// No device code: partitions are made once, before the tables are built.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

#include "engine_internal.h"

namespace {

int fail(const std::string& m) { return hlb_internal_fail(m.c_str()); }

struct Entry {  // boundary vertex u touches part q with `cnt` links; `internal` links stay inside its own part
  int64_t u;
  int32_t q;
  int64_t cnt, internal;
};

struct Graph {
  int64_t n;
  const int64_t* xadj;
  const int64_t* adjncy;
};

// every (boundary vertex, other part) pair, sorted by vertex then part, looking only at the sorted
// candidate vertices `watch` (a superset of the boundary), which shrinks to the boundary itself
void boundary(const Graph& g, const int32_t* part, int nparts, std::vector<int64_t>& watch, std::vector<Entry>& out,
              std::vector<int64_t>& scratch) {
  out.clear();
  scratch.assign(nparts, 0);
  std::vector<int32_t> touched;
  size_t kept = 0;
  for (int64_t u : watch) {
    const int32_t p = part[u];
    int64_t outside = 0;
    touched.clear();
    for (int64_t e = g.xadj[u]; e < g.xadj[u + 1]; ++e) {
      const int32_t q = part[g.adjncy[e]];
      if (q == p) continue;
      if (scratch[q]++ == 0) touched.push_back(q);
      ++outside;
    }
    if (!outside) continue;
    watch[kept++] = u;
    std::sort(touched.begin(), touched.end());
    const int64_t internal = (g.xadj[u + 1] - g.xadj[u]) - outside;
    for (int32_t q : touched) {
      out.push_back({u, q, scratch[q], internal});
      scratch[q] = 0;
    }
  }
  watch.resize(kept);
}

// cut links seen from the vertices in `some` (all vertices when null)
int64_t directed_cut(const Graph& g, const int32_t* part, const std::vector<int64_t>* some = nullptr) {
  int64_t c = 0;
  auto one = [&](int64_t u) {
    for (int64_t e = g.xadj[u]; e < g.xadj[u + 1]; ++e) c += part[g.adjncy[e]] != part[u];
  };
  if (some)
    for (int64_t u : *some) one(u);
  else
    for (int64_t u = 0; u < g.n; ++u) one(u);
  return c;
}

// the moved vertices and their neighbours, sorted: every vertex whose cut links can have changed
void touched_by(const Graph& g, const std::vector<int64_t>& moved, std::vector<int64_t>& out) {
  out = moved;
  for (int64_t u : moved) out.insert(out.end(), g.adjncy + g.xadj[u], g.adjncy + g.xadj[u + 1]);
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
}

void merge_into(std::vector<int64_t>& watch, const std::vector<int64_t>& extra) {
  std::vector<int64_t> all(watch.size() + extra.size());
  std::merge(watch.begin(), watch.end(), extra.begin(), extra.end(), all.begin());
  all.erase(std::unique(all.begin(), all.end()), all.end());
  watch.swap(all);
}

struct Cand {
  int64_t u;
  int32_t q;
  int64_t gain, group;
  double w;
};

// of the admissible entries of one vertex keep the one with the largest gain (ties: lowest part)
template <class Admit>
void best_per_vertex(const std::vector<Entry>& b, Admit admit, std::vector<Cand>& out) {
  out.clear();
  for (size_t i = 0; i < b.size();) {
    size_t j = i;
    bool have = false;
    Cand best{};
    for (; j < b.size() && b[j].u == b[i].u; ++j) {
      if (!admit(b[j])) continue;
      const int64_t gain = b[j].cnt - b[j].internal;
      if (!have || gain > best.gain) {
        best = {b[j].u, b[j].q, gain, 0, 0.0};
        have = true;
      }
    }
    if (have) out.push_back(best);
    i = j;
  }
}

// each group may take budget[group] weight, largest gain first (ties: lowest vertex); a candidate is
// accepted when the weight asked for so far, itself included (less half its own when `half`), fits
void take_within(std::vector<Cand>& c, const std::vector<double>& budget, bool half, std::vector<char>& keep) {
  std::stable_sort(c.begin(), c.end(), [](const Cand& a, const Cand& b) {
    if (a.group != b.group) return a.group < b.group;
    return a.gain > b.gain;
  });
  keep.assign(c.size(), 0);
  double cum = 0;
  for (size_t i = 0; i < c.size(); ++i) {
    if (i == 0 || c[i].group != c[i - 1].group) cum = 0;
    cum += c[i].w;
    keep[i] = cum - (half ? c[i].w / 2 : 0.0) <= budget[c[i].group];
  }
}

void part_loads(int64_t n, const int32_t* part, const double* w, int nparts, std::vector<double>& pl) {
  pl.assign(nparts, 0.0);
  for (int64_t i = 0; i < n; ++i) pl[part[i]] += w[i];
}

// potentials x with L x = rhs on the part graph (L = D - A), least-squares on every connected
// component; only differences x_p - x_q along edges are used
void potentials(const std::vector<char>& A, int k, std::vector<double> rhs, std::vector<double>& x) {
  x.assign(k, 0.0);
  std::vector<int> comp(k, -1);
  for (int s = 0; s < k; ++s) {
    if (comp[s] >= 0) continue;
    std::vector<int> nodes{s};
    comp[s] = s;
    for (size_t h = 0; h < nodes.size(); ++h)
      for (int q = 0; q < k; ++q)
        if (A[nodes[h] * k + q] && comp[q] < 0) {
          comp[q] = s;
          nodes.push_back(q);
        }
    const int m = (int)nodes.size();
    if (m == 1) continue;
    double mean = 0;
    for (int v : nodes) mean += rhs[v];
    mean /= m;
    // pin nodes[0] at 0 and solve the remaining (m-1) x (m-1) SPD system by elimination
    const int r = m - 1;
    std::vector<double> M((size_t)r * r, 0.0), b(r);
    for (int i = 0; i < r; ++i) {
      const int vi = nodes[i + 1];
      b[i] = rhs[vi] - mean;
      double deg = 0;
      for (int q = 0; q < k; ++q) deg += A[vi * k + q];
      M[(size_t)i * r + i] = deg;
      for (int j = 0; j < r; ++j)
        if (j != i && A[vi * k + nodes[j + 1]]) M[(size_t)i * r + j] = -1.0;
    }
    for (int c = 0; c < r; ++c) {
      int piv = c;
      for (int i = c + 1; i < r; ++i)
        if (std::fabs(M[(size_t)i * r + c]) > std::fabs(M[(size_t)piv * r + c])) piv = i;
      if (piv != c) {
        for (int j = 0; j < r; ++j) std::swap(M[(size_t)c * r + j], M[(size_t)piv * r + j]);
        std::swap(b[c], b[piv]);
      }
      const double d = M[(size_t)c * r + c];
      if (d == 0.0) continue;
      for (int i = c + 1; i < r; ++i) {
        const double f = M[(size_t)i * r + c] / d;
        if (f == 0.0) continue;
        for (int j = c; j < r; ++j) M[(size_t)i * r + j] -= f * M[(size_t)c * r + j];
        b[i] -= f * b[c];
      }
    }
    for (int i = r - 1; i >= 0; --i) {
      double sum = b[i];
      for (int j = i + 1; j < r; ++j) sum -= M[(size_t)i * r + j] * x[nodes[j + 1]];
      const double d = M[(size_t)i * r + i];
      x[nodes[i + 1]] = d != 0.0 ? sum / d : 0.0;
    }
  }
}

// eigenvector of the largest eigenvalue of a symmetric 3x3 matrix (cyclic Jacobi)
void principal_axis(double a[3][3], double axis[3]) {
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    if (off <= 1e-300 || off <= 1e-18 * (std::fabs(a[0][0]) + std::fabs(a[1][1]) + std::fabs(a[2][2]))) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int k = 1; k < 3; ++k)
    if (a[k][k] > a[best][best]) best = k;
  int big = 0;
  for (int k = 1; k < 3; ++k)
    if (std::fabs(v[k][best]) > std::fabs(v[big][best])) big = k;
  const double sign = v[big][best] > 0 ? 1.0 : -1.0;  // eigenvectors carry no sign
  for (int k = 0; k < 3; ++k) axis[k] = sign * v[k][best];
}

}  // namespace

extern "C" {

int hlb_part_bisect(int64_t n, const int64_t* coords, const double* weights, int nparts, int inertial, int32_t* part) {
  if (n < 0 || nparts < 1 || (n > 0 && (!coords || !weights || !part))) return fail("hlb_part_bisect: bad argument");
  struct Job { std::vector<int64_t> idx; int parts, first; };
  std::vector<Job> todo;
  {
    Job all;
    all.idx.resize(n);
    std::iota(all.idx.begin(), all.idx.end(), (int64_t)0);
    all.parts = nparts;
    all.first = 0;
    todo.push_back(std::move(all));
  }
  std::vector<double> key;
  while (!todo.empty()) {
    Job job = std::move(todo.back());
    todo.pop_back();
    const int64_t m = (int64_t)job.idx.size();
    if (job.parts == 1 || m == 0) {
      for (int64_t i : job.idx) part[i] = job.first;
      continue;
    }
    std::vector<int64_t> order(m);
    std::iota(order.begin(), order.end(), (int64_t)0);
    if (inertial) {
      double wsum = 0, mu[3] = {0, 0, 0};
      for (int64_t i : job.idx) {
        wsum += weights[i];
        for (int k = 0; k < 3; ++k) mu[k] += coords[3 * i + k] * weights[i];
      }
      for (int k = 0; k < 3; ++k) mu[k] /= wsum;
      double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int64_t i : job.idx) {
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = coords[3 * i + k] - mu[k];
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) cov[a][b] += weights[i] * d[a] * d[b];
      }
      double axis[3];
      principal_axis(cov, axis);
      key.resize(m);
      for (int64_t j = 0; j < m; ++j) {
        const int64_t i = job.idx[j];
        key[j] = (coords[3 * i] - mu[0]) * axis[0] + (coords[3 * i + 1] - mu[1]) * axis[1] + (coords[3 * i + 2] - mu[2]) * axis[2];
      }
      std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key[a] < key[b]; });
    } else {
      int64_t lo[3], hi[3];
      for (int k = 0; k < 3; ++k) lo[k] = hi[k] = coords[3 * job.idx[0] + k];
      for (int64_t i : job.idx)
        for (int k = 0; k < 3; ++k) {
          lo[k] = std::min(lo[k], coords[3 * i + k]);
          hi[k] = std::max(hi[k], coords[3 * i + k]);
        }
      int ax = 0;
      for (int k = 1; k < 3; ++k)
        if (hi[k] - lo[k] > hi[ax] - lo[ax]) ax = k;
      const int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
      std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        const int64_t* p = coords + 3 * job.idx[a];
        const int64_t* q = coords + 3 * job.idx[b];
        if (p[ax] != q[ax]) return p[ax] < q[ax];
        if (p[a1] != q[a1]) return p[a1] < q[a1];
        return p[a2] < q[a2];
      });
    }
    std::vector<double> cum(m);
    double run = 0;
    for (int64_t j = 0; j < m; ++j) cum[j] = run += weights[job.idx[order[j]]];
    const int lo_parts = job.parts / 2;
    int64_t k = std::lower_bound(cum.begin(), cum.end(), cum[m - 1] * lo_parts / job.parts) - cum.begin();
    k = std::min(std::max(k + 1, (int64_t)lo_parts), m - (job.parts - lo_parts));
    Job a, b;
    a.parts = lo_parts;
    a.first = job.first;
    b.parts = job.parts - lo_parts;
    b.first = job.first + lo_parts;
    for (int64_t j = 0; j < m; ++j) (j < k ? a : b).idx.push_back(job.idx[order[j]]);
    todo.push_back(std::move(a));
    todo.push_back(std::move(b));
  }
  return 0;
}

int hlb_part_refine_kway(int64_t n, const int64_t* xadj, const int64_t* adjncy, const double* vwgt, int nparts,
                         double ubvec, int passes, int32_t* part, int64_t* edgecut) {
  if (n < 0 || nparts < 1 || (n > 0 && (!xadj || !vwgt || !part)) || (n > 0 && xadj[n] > 0 && !adjncy))
    return fail("hlb_part_refine_kway: bad argument");
  for (int64_t i = 0; i < n; ++i)
    if (part[i] < 0 || part[i] >= nparts) return fail("hlb_part_refine_kway: initial part outside [0, nparts)");
  const Graph g{n, xadj, adjncy};
  if (n == 0 || nparts < 2) {
    if (edgecut) *edgecut = 0;
    return 0;
  }
  double total = 0, heaviest = 0;
  for (int64_t i = 0; i < n; ++i) {
    total += vwgt[i];
    heaviest = std::max(heaviest, vwgt[i]);
  }
  const double mean = total / nparts;
  const double cap = std::max(ubvec * mean, mean + heaviest);  // a part can always take one more site than the mean
  std::vector<double> pl;
  part_loads(n, part, vwgt, nparts, pl);
  for (int e = 0; e < nparts; ++e) {
    if (pl[e] != 0.0) continue;
    // an empty part is seeded with the first site of the heaviest one and grows by diffusion
    const int heavy = (int)(std::max_element(pl.begin(), pl.end()) - pl.begin());
    for (int64_t i = 0; i < n; ++i)
      if (part[i] == heavy) {
        part[i] = e;
        break;
      }
    part_loads(n, part, vwgt, nparts, pl);
  }
  std::vector<Entry> b;
  std::vector<int64_t> scratch, watch(n), moved_now, near;
  std::iota(watch.begin(), watch.end(), (int64_t)0);
  std::vector<Cand> cand;
  std::vector<char> keep;
  std::vector<double> budget, x;
  for (int it = 0; it < passes; ++it) {
    if (*std::max_element(pl.begin(), pl.end()) > cap) {
      // (i) balance: loads diffuse over the part graph, one layer of boundary sites per pass
      boundary(g, part, nparts, watch, b, scratch);
      std::vector<char> A((size_t)nparts * nparts, 0);
      for (const Entry& e : b) A[(size_t)part[e.u] * nparts + e.q] = A[(size_t)e.q * nparts + part[e.u]] = 1;
      std::vector<double> rhs(nparts);
      for (int p = 0; p < nparts; ++p) rhs[p] = pl[p] - mean;
      potentials(A, nparts, rhs, x);
      budget.assign((size_t)nparts * nparts, 0.0);
      for (int p = 0; p < nparts; ++p)
        for (int q = 0; q < nparts; ++q)
          if (A[(size_t)p * nparts + q])  // to 1/1024 of a weight unit: moves do not hang on the solver's last bits
            budget[(size_t)p * nparts + q] = std::nearbyint((x[p] - x[q]) * 1024.0) / 1024.0;
      best_per_vertex(b, [&](const Entry& e) { return budget[(size_t)part[e.u] * nparts + e.q] > 0.5 * vwgt[e.u]; }, cand);
      if (cand.empty()) break;
      for (Cand& c : cand) {
        c.group = (int64_t)part[c.u] * nparts + c.q;
        c.w = vwgt[c.u];
      }
      take_within(cand, budget, true, keep);
      moved_now.clear();
      for (size_t i = 0; i < cand.size(); ++i)
        if (keep[i]) {
          part[cand[i].u] = cand[i].q;
          moved_now.push_back(cand[i].u);
        }
      if (moved_now.empty()) break;
      touched_by(g, moved_now, near);
      merge_into(watch, near);
      part_loads(n, part, vwgt, nparts, pl);
      continue;
    }
    // (ii) cut: one sweep to higher part indices, one to lower; a sweep that does not pay is undone
    int64_t moved = 0;
    for (int up = 1; up >= 0; --up) {
      boundary(g, part, nparts, watch, b, scratch);
      best_per_vertex(b, [&](const Entry& e) {
        const int32_t p = part[e.u];
        if (up ? !(e.q > p) : !(e.q < p)) return false;
        const int64_t gain = e.cnt - e.internal;
        return gain > 0 || (gain == 0 && pl[p] - pl[e.q] > 2 * vwgt[e.u]);
      }, cand);
      if (cand.empty()) continue;
      budget.assign(nparts, 0.0);
      for (int q = 0; q < nparts; ++q) budget[q] = cap - pl[q];
      for (Cand& c : cand) {
        c.group = c.q;
        c.w = vwgt[c.u];
      }
      take_within(cand, budget, false, keep);
      moved_now.clear();
      for (size_t i = 0; i < cand.size(); ++i)
        if (keep[i]) moved_now.push_back(cand[i].u);
      if (moved_now.empty()) continue;
      touched_by(g, moved_now, near);
      const int64_t before = directed_cut(g, part, &near);
      std::vector<int32_t> old;
      for (size_t i = 0; i < cand.size(); ++i)
        if (keep[i]) {
          old.push_back(part[cand[i].u]);
          part[cand[i].u] = cand[i].q;
        }
      std::vector<double> new_pl;
      part_loads(n, part, vwgt, nparts, new_pl);
      // stale gains of neighbours moving together; and no part may empty
      if (directed_cut(g, part, &near) > before || *std::min_element(new_pl.begin(), new_pl.end()) <= 0) {
        size_t k = 0;
        for (size_t i = 0; i < cand.size(); ++i)
          if (keep[i]) part[cand[i].u] = old[k++];
        continue;
      }
      merge_into(watch, near);
      pl.swap(new_pl);
      moved += (int64_t)moved_now.size();
    }
    if (!moved) break;
  }
  if (edgecut) *edgecut = directed_cut(g, part) / 2;
  return 0;
}

}  // extern "C"

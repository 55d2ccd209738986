// pp_host_lb.cpp -- the weight-diffusion plan of the particle load balancer.
//
// Replaces, for ParticleBalancer::balance (src/pumipic_lb.cpp:478-511):
//   the N-graph of ParticleBalancer::buildNgraph (src/pumipic_lb.cpp:395-462): one graph vertex per
//   (sbar, part) pair -- vertex id = sbar id + index of the part in the sbar's sorted part list
//   (:410, :443) -- one hyperedge per sbar, plus one extra vertex per part that carries the weight
//   of particles already forced onto it (:417, pumipic_lb.hpp:196-200);
//   engpar::balanceWeights (EnGPar >= 1.1.0, third party, NOT in the reference tree; CMakeLists.txt:61)
//   and the WeightPartitionMap it returns (:484-507).
// EnGPar's sources are absent and no reference test fixes its output (test/test_lb.cpp only bounds
// the resulting imbalance): PARITY UNPINNED.  EnGPar diffuses weight between neighbouring parts over
// many iterations, each exchanging weights with its neighbours.  Here every rank holds the global
// weight vector after ONE all-reduce of <= max_sbar + nranks doubles, so the plan is computed
// directly, and identically on every rank, as the transportation problem that diffusion
// approximates: overloaded parts give up their surplus over the average, through the regions
// (sbars) they share with underloaded parts, up to what each region holds -- a maximum flow on a
// graph of tens of nodes.  Consequences:
//   * the plan is the best any one-hop selection can reach (particles can only be handed to a part
//     that holds their element safely, i.e. a part of the same sbar);
//   * a vertex never plans to send more than the particles it holds when the plan is made (the
//     reference forwards weight it has only been promised and then selects what it can);
//   * step_factor (EnGPar's diffusion rate) has nothing to control and is only validated.
// No CUDA.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "pumipic_b200.h"

void pp_set_error(const char* fmt, ...);

namespace {

template <class T>
T* dup_out(const std::vector<T>& v) {
  T* p = (T*)malloc(sizeof(T) * (v.size() ? v.size() : 1));
  if (p && !v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
  return p;
}

double imbalance_of(const std::vector<double>& W) {
  double tot = 0, mx = 0;
  for (double w : W) { tot += w; mx = std::max(mx, w); }
  return tot > 0 ? mx / (tot / (double)W.size()) : 1.0;
}

}  // namespace

extern "C" pp_status pp_host_lb_plan(int32_t nranks, int32_t nsbars, const int32_t* sbar_ids,
                                     const int32_t* parts_off, const int32_t* parts, int32_t nverts,
                                     const double* vert_weight, const double* forced, double tol,
                                     double step_factor, int32_t* nsends,
                                     int32_t** send_vert, int32_t** send_part, double** send_weight,
                                     double* imbalance) {
  if (nranks < 1 || nsbars < 0 || nverts < 0 || !nsends || !send_vert || !send_part || !send_weight ||
      (nsbars > 0 && (!sbar_ids || !parts_off || !parts || !vert_weight))) {
    pp_set_error("pp_host_lb_plan: bad argument");
    return PP_ERR_INVALID;
  }
  if (!(step_factor > 0) || step_factor > 1) {
    pp_set_error("pp_host_lb_plan: step_factor must be in (0, 1]");
    return PP_ERR_INVALID;
  }
  // sbars in ascending id order: the plan must not depend on the order the caller lists them in
  std::vector<int> order((size_t)nsbars);
  for (int i = 0; i < nsbars; ++i) order[(size_t)i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return sbar_ids[a] < sbar_ids[b]; });
  for (int i = 0; i < nsbars; ++i) {
    const int s = order[(size_t)i];
    const int k = parts_off[s + 1] - parts_off[s];
    bool ok = k >= 1 && sbar_ids[s] >= 0 && sbar_ids[s] + k <= nverts;
    for (int j = 0; ok && j < k; ++j) {
      const int p = parts[parts_off[s] + j];
      ok = p >= 0 && p < nranks && (j == 0 || parts[parts_off[s] + j - 1] < p);
    }
    if (ok && i + 1 < nsbars) ok = sbar_ids[s] + k <= sbar_ids[order[(size_t)i + 1]];
    if (!ok) {
      pp_set_error("pp_host_lb_plan: sbar %d is malformed (parts must be sorted, ids spaced by size)",
                   sbar_ids[s]);
      return PP_ERR_INVALID;
    }
  }
  std::vector<double> W((size_t)nranks, 0.0), avail((size_t)nverts, 0.0);
  for (int p = 0; p < nranks; ++p) W[(size_t)p] = forced ? forced[p] : 0.0;
  for (int i = 0; i < nsbars; ++i) {
    const int s = order[(size_t)i];
    for (int j = parts_off[s]; j < parts_off[s + 1]; ++j) {
      const int v = sbar_ids[s] + (j - parts_off[s]);
      const double w = vert_weight[v] > 0 ? vert_weight[v] : 0.0;
      avail[(size_t)v] = w;
      W[(size_t)parts[j]] += w;
    }
  }
  if (imbalance) imbalance[0] = imbalance_of(W);
  double avg = 0;
  for (double x : W) avg += x;
  avg /= (double)nranks;

  // Transportation problem: source -> overloaded part p (capacity: its surplus W[p] - avg)
  //   -> p's vertex in sbar s (capacity: the particles it holds) -> every underloaded part q of s
  //   -> sink (capacity: q's deficit avg - W[q]); solved by shortest augmenting paths.
  std::map<std::pair<int, int>, double> flow;   // (vertex, target part) -> weight
  if (avg > 0 && imbalance_of(W) > tol) {
    // node numbering: 0 source, 1 sink, 2 + p parts as senders, 2 + R + q parts as receivers,
    // 2 + 2R + v vertices
    const int R = nranks, NN = 2 + 2 * R + nverts;
    struct Edge { int to; double cap; };
    std::vector<Edge> edges;
    std::vector<std::vector<int>> adj((size_t)NN);
    auto add_edge = [&](int u, int v, double c) {
      adj[(size_t)u].push_back((int)edges.size()); edges.push_back({v, c});
      adj[(size_t)v].push_back((int)edges.size()); edges.push_back({u, 0.0});
    };
    for (int p = 0; p < R; ++p) {
      if (W[(size_t)p] > avg) add_edge(0, 2 + p, W[(size_t)p] - avg);
      if (W[(size_t)p] < avg) add_edge(2 + R + p, 1, avg - W[(size_t)p]);
    }
    std::vector<std::pair<int, std::pair<int, int>>> send_edges;   // edge index -> (vertex, target)
    for (int i = 0; i < nsbars; ++i) {
      const int s = order[(size_t)i];
      for (int a = parts_off[s]; a < parts_off[s + 1]; ++a) {
        const int p = parts[a], v = sbar_ids[s] + (a - parts_off[s]);
        if (!(W[(size_t)p] > avg) || !(avail[(size_t)v] > 0)) continue;
        add_edge(2 + p, 2 + 2 * R + v, avail[(size_t)v]);
        for (int b = parts_off[s]; b < parts_off[s + 1]; ++b) {
          const int q = parts[b];
          if (q == p || !(W[(size_t)q] < avg)) continue;
          send_edges.push_back(std::make_pair((int)edges.size(), std::make_pair(v, q)));
          add_edge(2 + 2 * R + v, 2 + R + q, avail[(size_t)v]);
        }
      }
    }
    const double eps = 1e-9;
    std::vector<int> prev_edge((size_t)NN), queue;
    for (long guard = 0; guard < 4L * NN * (long)edges.size() + 64; ++guard) {
      std::fill(prev_edge.begin(), prev_edge.end(), -1);
      queue.assign(1, 0);
      prev_edge[0] = -2;
      for (size_t h = 0; h < queue.size() && prev_edge[1] == -1; ++h)
        for (int e : adj[(size_t)queue[h]])
          if (edges[(size_t)e].cap > eps && prev_edge[(size_t)edges[(size_t)e].to] == -1) {
            prev_edge[(size_t)edges[(size_t)e].to] = e;
            queue.push_back(edges[(size_t)e].to);
          }
      if (prev_edge[1] == -1) break;
      double f = 1e300;
      for (int v = 1; v != 0; v = edges[(size_t)(prev_edge[(size_t)v] ^ 1)].to)
        f = std::min(f, edges[(size_t)prev_edge[(size_t)v]].cap);
      for (int v = 1; v != 0; v = edges[(size_t)(prev_edge[(size_t)v] ^ 1)].to) {
        edges[(size_t)prev_edge[(size_t)v]].cap -= f;
        edges[(size_t)(prev_edge[(size_t)v] ^ 1)].cap += f;
      }
    }
    for (const auto& se : send_edges) {
      const double f = edges[(size_t)(se.first ^ 1)].cap;   // flow = residual of the reverse edge
      if (!(f > 0)) continue;
      flow[se.second] += f;
      const int v = se.second.first, q = se.second.second;
      int p = -1;
      for (int i = 0; i < nsbars && p < 0; ++i)
        if (v >= sbar_ids[i] && v < sbar_ids[i] + (parts_off[i + 1] - parts_off[i])) p = parts[parts_off[i] + v - sbar_ids[i]];
      W[(size_t)p] -= f;
      W[(size_t)q] += f;
    }
  }
  std::vector<int32_t> ov, op;
  std::vector<double> ow;
  for (const auto& kv : flow) {
    if (!(kv.second > 1e-6)) continue;
    ov.push_back(kv.first.first);
    op.push_back(kv.first.second);
    ow.push_back(kv.second);
  }
  if (imbalance) imbalance[1] = imbalance_of(W);
  *nsends = (int32_t)ov.size();
  *send_vert = dup_out(ov);
  *send_part = dup_out(op);
  *send_weight = dup_out(ow);
  if (!*send_vert || !*send_part || !*send_weight) {
    free(*send_vert); free(*send_part); free(*send_weight);
    *send_vert = *send_part = nullptr; *send_weight = nullptr; *nsends = 0;
    pp_set_error("pp_host_lb_plan: out of memory");
    return PP_ERR_INVALID;
  }
  return PP_OK;
}

// sharding.hpp — which GPU scores which swarm (swarms are independent, no collective on the data path).
//
// Round 1 used swarm s -> GPU s mod G.  The swarms of a run do not cost the same: a swarm whose centre sits close to
// the receptor has 3-5x the in-reach atom pairs of one further out, and at 8 GPUs the slowest rank of the static map
// ran 10 % above the mean.  The map below is cost-aware and deterministic:
//   cost(swarm) = sum over receptor atoms of F(|atom - swarm centre|), F(d) = expected number of ligand atoms within the
//                 15 A DFIRE reach of a point at distance d from the ligand's centroid, averaged over orientations
//                 (closed form per ligand atom: the fraction of the sphere of radius |l_j| inside the reach) -- the
//                 expectation of what the ligand-frame cell lists hand to the pair loop;
//   assignment  = longest-processing-time first: swarms by decreasing cost, each to the GPU with the least load so far
//                 (ties: lowest GPU index; equal costs: lowest swarm index first).
// The kernels are batch-invariant, so the map cannot change a bit of any result.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <numeric>
#include <vector>

namespace lightdock {

struct SwarmCostModel {
  std::vector<double> rec;   // [n][3] receptor coordinates (lab frame)
  std::vector<double> f;     // F(d) sampled every `step` A
  double step = 0.25, reach = 15.0;
  double centroid[3] = {0, 0, 0};

  SwarmCostModel(const std::vector<double> &rec_xyz, const std::vector<double> &lig_xyz, double reach_ = 15.0)
      : rec(rec_xyz), reach(reach_) {
    const size_t nl = lig_xyz.size() / 3;
    for (size_t j = 0; j < nl; ++j)
      for (int d = 0; d < 3; ++d) centroid[d] += lig_xyz[3 * j + d] / (double)std::max<size_t>(nl, 1);
    std::vector<double> a(nl);
    double amax = 0.0;
    for (size_t j = 0; j < nl; ++j) {
      double s = 0.0;
      for (int d = 0; d < 3; ++d) s += (lig_xyz[3 * j + d] - centroid[d]) * (lig_xyz[3 * j + d] - centroid[d]);
      a[j] = std::sqrt(s);
      amax = std::max(amax, a[j]);
    }
    const size_t nb = (size_t)std::ceil((amax + reach) / step) + 2;
    f.assign(nb, 0.0);
    for (size_t b = 0; b < nb; ++b) {
      const double d = b * step;
      double sum = 0.0;
      for (size_t j = 0; j < nl; ++j) {
        const double aj = a[j];
        if (d + aj <= reach) sum += 1.0;
        else if (std::fabs(d - aj) >= reach) continue;
        else if (aj < 1e-9 || d < 1e-9) sum += std::max(aj, d) <= reach ? 1.0 : 0.0;
        else sum += std::min(1.0, std::max(0.0, 0.5 * (1.0 - (aj * aj + d * d - reach * reach) / (2.0 * aj * d))));
      }
      f[b] = sum;
    }
  }

  // centre: where the ligand's centroid sits for a typical pose of the swarm (mean translation of its glowworms; the
  // reference's ligands are centred at the origin, so the rotation of the centroid is ignored)
  double cost(const double centre[3]) const {
    double total = 0.0;
    const size_t n = rec.size() / 3;
    for (size_t i = 0; i < n; ++i) {
      const double dx = rec[3 * i] - centre[0] - centroid[0], dy = rec[3 * i + 1] - centre[1] - centroid[1],
                   dz = rec[3 * i + 2] - centre[2] - centroid[2];
      const double x = std::sqrt(dx * dx + dy * dy + dz * dz) / step;
      const size_t b = (size_t)x;
      if (b + 1 >= f.size()) continue;
      total += f[b] + (f[b + 1] - f[b]) * (x - (double)b);
    }
    return total + 1.0;  // + a constant per swarm: a swarm out of reach still costs its per-pose overhead
  }
};

// Longest-processing-time-first assignment: gpu_of[s] for every swarm.
inline std::vector<int> assign_swarms_lpt(const std::vector<double> &cost, int n_gpus) {
  const size_t n = cost.size();
  std::vector<size_t> order(n);
  std::iota(order.begin(), order.end(), (size_t)0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost[a] > cost[b]; });
  std::vector<double> load((size_t)std::max(1, n_gpus), 0.0);
  std::vector<int> gpu_of(n, 0);
  for (size_t s : order) {
    size_t best = 0;
    for (size_t g = 1; g < load.size(); ++g)
      if (load[g] < load[best]) best = g;
    gpu_of[s] = (int)best;
    load[best] += cost[s];
  }
  return gpu_of;
}

}  // namespace lightdock

// ld_capi.cu — C-ABI implementation (include/lightdock_b200.h): builds the device-resident complex,
// owns device/pinned buffers and the stream, and launches the three kernels per batch.
// There is no CPU fallback anywhere in this file: every entry point either runs the CUDA kernels
// or returns an error.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "ld_cells.cuh"
#include "ld_gso.cuh"
#include "ld_kernels.cuh"
#include "ld_rigid.cuh"

// NVTX ranges around the phases of a host-buffer call (staging + H2D, kernel launches, D2H + wait): visible in an
// Nsight Systems timeline, free otherwise.
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
namespace { struct Nvtx { explicit Nvtx(const char *n) { nvtxRangePushA(n); } ~Nvtx() { nvtxRangePop(); } }; }
#else
namespace { struct Nvtx { explicit Nvtx(const char *) {} }; }
#endif

using namespace ldb200;

// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                                    \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return fail(LD_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                std::to_string(__LINE__) + ")");                                    \
  } while (0)

extern "C" const char *ld_last_error(void) { return g_err.c_str(); }
extern "C" int ld_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
extern "C" const char *ld_version(void) { return "lightdock_b200 0.2 (sm_100a)"; }

// Process-wide tuning defaults read by ld_create / the launch code (ld_set_option); the library itself never looks
// at the environment.
namespace {
struct Options {
  int rigid_rows = RG_MAX_ROWS;   // table rows a receptor group may span (1..RG_MAX_ROWS)
  double cell_size = 0.0;         // ligand-frame cell size in A; 0 = chosen per complex (build_cells)
  int units_per_sm = 16;          // rigid kernel: work units per SM
  int flex_min_warps = 18;        // FLEX: warps per CTA wanted before a receptor group may span one more table row
  int default_path = LD_PATH_AUTO;
  int flex = 1;                   // 0: ligands with ANM modes stay on the generic kernel (no FLEX instance of the ligand-frame path)
  int cells_on_host = 0;          // 1: build the ligand-frame cell lists with host threads (the round-1 builder; cross-check)
  int compact_tiles = 1;          // 0: keep the plain bisection order of the atoms (tiles less compact; A/B aid)
  int dna_fused = 1;              // 0: DNA/pyDock poses go through transform_kernel and per-pose coordinate blocks (cross-check)
};
Options g_opt;
}  // namespace

extern "C" int ld_set_option(const char *key, double value) {
  if (!key) return fail(LD_EINVAL, "ld_set_option: NULL key");
  const std::string k(key);
  if (k == "rigid_rows") g_opt.rigid_rows = std::max(1, std::min(RG_MAX_ROWS, (int)value));
  else if (k == "cell_size") g_opt.cell_size = value <= 0.0 ? 0.0 : std::max(0.5, std::min(8.0, value));
  else if (k == "units_per_sm") g_opt.units_per_sm = std::max(1, (int)value);
  else if (k == "flex_min_warps") g_opt.flex_min_warps = std::max(8, std::min(RG_WARPS, (int)value));
  else if (k == "cells_on_host") g_opt.cells_on_host = value != 0.0;
  else if (k == "flex") g_opt.flex = value != 0.0;
  else if (k == "compact_tiles") g_opt.compact_tiles = value != 0.0;
  else if (k == "dna_fused") g_opt.dna_fused = value != 0.0;
  else if (k == "default_path") {
    if (value != LD_PATH_AUTO && value != LD_PATH_GENERIC) return fail(LD_EINVAL, "ld_set_option: default_path is AUTO or GENERIC");
    g_opt.default_path = (int)value;
  } else return fail(LD_EINVAL, "ld_set_option: unknown key " + k);
  return LD_OK;
}

// ---------------------------------------------------------------------------------------------
// Everything one batch in flight needs: a stream, pinned staging, device pose/energy buffers and the kernels'
// work buffers.  A handle owns LD_SLOTS of them so ld_score_batch_begin/_end can keep two batches in flight (the
// host prepares the next batch while the device scores the current one); the synchronous calls use slot 0.
struct Workspace {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t cap_poses = 0;   // capacity of poses/energies/detail buffers
  int64_t cap_chunk = 0;   // capacity (poses) of partial/bitmap buffers
  int64_t cap_blocks = 0;  // capacity (poses) of the per-pose coordinate blocks (generic path only)
  int cap_splits = 0;
  double *d_poses = nullptr, *d_energies = nullptr;
  ld_pose_detail *d_detail = nullptr;
  unsigned char *d_lig_blocks = nullptr, *d_rec_blocks = nullptr;
  double *d_partials = nullptr;
  unsigned *d_iface_rec = nullptr, *d_iface_lig = nullptr;
  double *h_poses = nullptr, *h_energies = nullptr;  // pinned
  int64_t cap_pinned = 0;
  unsigned *d_unit_counter = nullptr;   // rigid path: work-unit counter
  double *d_prep = nullptr;             // rigid path: [cap_prep][RG_PREP] per-pose rotation data
  int64_t cap_prep = 0;
  float4 *d_lig4p = nullptr;            // FLEX: [cap_flex][n_lig_pad] per-pose ligand blocks (ligand frame, f32)
  float *d_flag = nullptr;              // FLEX: [cap_flex] 0, or the displacement that sends the pose to the brute-force route
  int64_t cap_flex = 0;
  ld_batch_stats stats{};
  // profiling: events bracketing every kernel of the last call (4 per chunk)
  std::vector<cudaEvent_t> prof_events;
  size_t prof_used = 0;
  cudaStream_t last_stream = nullptr;
  // Orders the work buffers of this workspace across streams: recorded after the last kernel that touches them,
  // waited on by the next user whatever stream it launches on (ld_score_batch_device takes the caller's stream).
  cudaEvent_t done = nullptr;
  bool done_recorded = false;
  int64_t pending = -1;                 // poses of the batch begun on this slot and not yet ended (-1 = none)
};

struct ld_handle {
  int device = 0;
  Workspace ws[LD_SLOTS];
  Workspace *w = &ws[0];                // the slot the current call works on (one host thread per handle)
  Workspace *last_w = &ws[0];           // the slot of the last scoring call (ld_get_stats)
  DeviceComplex cx{};
  std::vector<void *> owned;  // device allocations of the complex
  std::vector<int> rec_perm, lig_perm;  // sorted position -> original atom index
  std::vector<double> rec_xyz_orig;     // original receptor coordinates (for ld_transform_batch)
  int use_anm = 0;
  int forced_splits = 0;
  size_t lig_block = 0, rec_block = 0;
  // rigid-ligand DFIRE path (ld_rigid.cuh): type-grouped receptor + ligand-frame cell lists
  bool rigid_ok = false;
  int path_mode = LD_PATH_AUTO;
  RigidComplex rc{};
  DeviceComplex cxr{};                  // what finalize_kernel sees on the rigid path (groups as tiles)
  std::vector<int> rec_perm_r;          // grouped position -> original atom index, -1 = pad lane
  RigidComplex *d_rc = nullptr;         // device copy of rc for the rare exact path
  std::string rigid_info;
  uint2 *d_cells = nullptr;             // ligand-frame cell lists (replaced when the FLEX slacks grow)
  unsigned short *d_cell_tiles = nullptr;
  std::vector<double> lig_sx, lig_sy, lig_sz;  // sorted ligand coordinates (cell-list rebuilds)
  // FLEX: ligand with ANM modes on the ligand-frame path
  bool flex = false;
  int flex_warps = 0, flex_rebuilds = 0;
  std::vector<float> tile_slack;        // per ligand tile: slack the current lists were built with
  float *d_tile_slack = nullptr;
  int *d_need = nullptr, *h_need = nullptr;  // running max of the tile displacements seen (float bits), device / pinned
  int max_smem_optin = 0;
  bool dna_fused = false;               // DNA/pyDock: the pair kernel transforms its pose itself (option at ld_create time)
  unsigned char *gso_slab = nullptr;    // device memory of the handle's ld_gso, kept across ld_gso_create/_destroy (an
  size_t gso_slab_bytes = 0;            // allocation per optimisation run costs up to 100+ ms now and then)
  bool gso_live = false;                // an ld_gso exists on this handle
  int sm_count = 0;
  double create_ms[4] = {0, 0, 0, 0};   // ld_create: CUDA context, complex (sort + upload), rigid groups, cell lists
  bool profiling = false;
  bool prof_continue = false;  // second part of a two-part call: keep the first part's profiling events
};

static double ms_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

template <typename T>
static int upload(ld_handle *h, const std::vector<T> &v, const T **out) {
  *out = nullptr;
  if (v.empty()) return LD_OK;
  void *d = nullptr;
  CU(cudaMalloc(&d, v.size() * sizeof(T)));
  h->owned.push_back(d);
  CU(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<const T *>(d);
  return LD_OK;
}

// Spatial tiling: recursive bisection along the longest axis, cutting at a multiple of `tile`, so
// that consecutive runs of `tile` atoms are compact.  Returns perm[sorted position] = original index.
static void bisect(std::vector<int> &idx, int lo, int hi, const double *xyz, int tile) {
  const int n = hi - lo;
  if (n <= tile) return;
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = lo; i < hi; ++i)
    for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], xyz[3 * idx[i] + d]);
      mx[d] = std::max(mx[d], xyz[3 * idx[i] + d]);
    }
  int ax = 0;
  for (int d = 1; d < 3; ++d)
    if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
  const int tiles = (n + tile - 1) / tile;
  const int mid = lo + (tiles / 2) * tile;
  std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
    const double va = xyz[3 * a + ax], vb = xyz[3 * b + ax];
    return va < vb || (va == vb && a < b);
  });
  bisect(idx, lo, mid, xyz, tile);
  bisect(idx, mid, hi, xyz, tile);
}
// Makes the tiles of a bisection order more COMPACT.  A tile enters a cell's list (ligand-frame path) or survives the
// sphere culling (generic paths) as soon as ONE of its atoms is in reach, and then all of its atoms are tested: the
// executed pair tests follow the tiles' radii.  Bisection leaves 1k4c's 8-atom ligand tiles with a mean radius of
// 4.2 A; a capacity-constrained k-means started from those tiles (atoms claim their nearest centroid in order of
// decreasing regret) followed by pairwise swaps that lower the sum of squared distances brings it to 3.3 A, which is
// 6 % fewer executed pair tests on the bench workload.  Deterministic (no random numbers, ties by index); tile k
// stays where bisection put it, so consecutive tile ids remain neighbours.
static void compact_tiles(std::vector<int> &idx, const double *xyz, int n, int tile) {
  const int K = (n + tile - 1) / tile;
  if (K < 2) return;
  if ((size_t)n * (size_t)K > (size_t)40000000) return;  // the one full atom x tile search below is quadratic: beyond
                                                          // ~18,000 atoms (seconds) the bisection order is kept
  std::vector<int> cap(K, tile), assign(n);
  cap[K - 1] = n - (K - 1) * tile;
  for (int i = 0; i < n; ++i) assign[idx[i]] = i / tile;
  std::vector<double> cent(3 * (size_t)K);
  auto centroids = [&] {
    std::fill(cent.begin(), cent.end(), 0.0);
    std::vector<int> cnt(K, 0);
    for (int a = 0; a < n; ++a) {
      for (int d = 0; d < 3; ++d) cent[3 * (size_t)assign[a] + d] += xyz[3 * (size_t)a + d];
      ++cnt[assign[a]];
    }
    for (int k = 0; k < K; ++k)
      for (int d = 0; d < 3; ++d) cent[3 * (size_t)k + d] /= std::max(1, cnt[k]);
  };
  auto dist2 = [&](int a, int k) {
    double v = 0.0;
    for (int d = 0; d < 3; ++d) {
      const double e = xyz[3 * (size_t)a + d] - cent[3 * (size_t)k + d];
      v += e * e;
    }
    return v;
  };
  // candidate tiles of an atom: the MC tiles nearest to it at the start (centroids move by a fraction of a tile, so the
  // set is searched once, in full, and only re-ranked afterwards)
  const int MC = std::min(16, K), M = std::min(12, K);
  std::vector<int> candk((size_t)n * MC), nb((size_t)n * M);
  std::vector<double> nd((size_t)n * M);
  {
    centroids();
    std::vector<std::pair<double, int>> all(K);
    for (int a = 0; a < n; ++a) {
      for (int k = 0; k < K; ++k) all[k] = {dist2(a, k), k};
      std::partial_sort(all.begin(), all.begin() + MC, all.end());
      for (int j = 0; j < MC; ++j) candk[(size_t)a * MC + j] = all[j].second;
    }
  }
  auto nearest = [&] {  // the M nearest of the candidate centroids of every atom, ascending
    std::pair<double, int> c[16];
    for (int a = 0; a < n; ++a) {
      for (int j = 0; j < MC; ++j) c[j] = {dist2(a, candk[(size_t)a * MC + j]), candk[(size_t)a * MC + j]};
      std::sort(c, c + MC);
      for (int j = 0; j < M; ++j) { nd[(size_t)a * M + j] = c[j].first; nb[(size_t)a * M + j] = c[j].second; }
    }
  };
  std::vector<int> order(n), load(K), next(n);
  for (int it = 0; it < 12; ++it) {  // capacity-constrained k-means
    centroids();
    nearest();
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {  // largest regret (second best - best) first
      const double ra = M > 1 ? nd[(size_t)a * M + 1] - nd[(size_t)a * M] : 0.0;
      const double rb = M > 1 ? nd[(size_t)b * M + 1] - nd[(size_t)b * M] : 0.0;
      return ra > rb;
    });
    std::fill(load.begin(), load.end(), 0);
    std::fill(next.begin(), next.end(), -1);
    for (int a : order)
      for (int j = 0; j < M; ++j) {
        const int k = nb[(size_t)a * M + j];
        if (load[k] < cap[k]) { next[a] = k; ++load[k]; break; }
      }
    for (int a : order)
      if (next[a] < 0) {  // its M nearest tiles were full: the nearest one with room
        int best = -1;
        double bd = 0.0;
        for (int k = 0; k < K; ++k)
          if (load[k] < cap[k]) {
            const double d2 = dist2(a, k);
            if (best < 0 || d2 < bd) { best = k; bd = d2; }
          }
        next[a] = best;
        ++load[best];
      }
    if (next == assign) break;
    assign = next;
  }
  std::vector<std::vector<int>> members(K);
  for (int a = 0; a < n; ++a) members[assign[a]].push_back(a);
  std::vector<double> own(n), rad2(K);
  for (int pass = 0; pass < 20; ++pass) {  // pairwise swaps between neighbouring tiles
    centroids();
    nearest();
    std::fill(rad2.begin(), rad2.end(), 0.0);
    for (int a = 0; a < n; ++a) {
      own[a] = dist2(a, assign[a]);
      rad2[assign[a]] = std::max(rad2[assign[a]], own[a]);
    }
    int moved = 0;
    for (int a = 0; a < n; ++a) {
      const int ka = assign[a];
      const double da = dist2(a, ka);
      for (int j = 0; j < std::min(4, M); ++j) {
        const int kb = nb[(size_t)a * M + j];
        if (kb == ka) continue;
        // gain of swapping a with b of tile kb = [d(a,ka) - d(a,kb)] + [d(b,kb) - d(b,ka)]; the second bracket is at
        // most the squared radius of tile kb
        const double ga = da - nd[(size_t)a * M + j];
        if (ga + rad2[kb] <= 1e-9) continue;
        int bj = -1;
        double bg = 1e-9;
        for (size_t t = 0; t < members[kb].size(); ++t) {
          const int b = members[kb][t];
          const double gain = ga + dist2(b, kb) - dist2(b, ka);
          if (gain > bg) { bg = gain; bj = (int)t; }
        }
        if (bj >= 0) {
          const int b = members[kb][bj];
          members[kb][bj] = a;
          *std::find(members[ka].begin(), members[ka].end(), a) = b;
          assign[a] = kb;
          assign[b] = ka;
          ++moved;
          break;
        }
      }
    }
    if (moved * 200 < n) break;  // converged: fewer than 0.5 % of the atoms still move
  }
  int pos = 0;
  for (int k = 0; k < K; ++k) {
    std::sort(members[k].begin(), members[k].end());
    for (int a : members[k]) idx[pos++] = a;
  }
}
static std::vector<int> spatial_order(const double *xyz, int n, int tile, bool compact) {
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  bisect(idx, 0, n, xyz, tile);
  if (compact && g_opt.compact_tiles) compact_tiles(idx, xyz, n, tile);
  return idx;
}

static int check_molecule(const ld_molecule_desc &m, int method, int use_anm, const char *who) {
  const std::string w(who);
  if (m.n_atoms < 0) return fail(LD_EINVAL, w + ": negative atom count");
  if (m.n_atoms > 0 && !m.coords) return fail(LD_EINVAL, w + ": coords is NULL");
  if (method == LD_METHOD_DFIRE) {
    if (m.n_atoms > 0 && !m.dfire_type) return fail(LD_EINVAL, w + ": dfire_type is NULL");
    for (int i = 0; i < m.n_atoms; ++i)
      if (m.dfire_type[i] < 0 || m.dfire_type[i] > 167)
        return fail(LD_EINVAL, w + ": DFIRE atom type out of range (the reference would index past the table)");
  } else if (m.n_atoms > 0 && (!m.ele_charge || !m.vdw_energy || !m.vdw_radius)) {
    return fail(LD_EINVAL, w + ": DNA/pyDock parameters are NULL");
  }
  if (use_anm && m.n_modes > 0 && m.n_atoms > 0 && !m.modes) return fail(LD_EINVAL, w + ": modes is NULL");
  if (m.n_modes < 0 || m.n_modes > 64) return fail(LD_EINVAL, w + ": unsupported number of ANM modes");
  if (m.n_restraints < 0 || m.n_membrane < 0) return fail(LD_EINVAL, w + ": negative count");
  if (m.n_restraints > 0) {
    if (!m.rst_offsets || !m.rst_atoms) return fail(LD_EINVAL, w + ": restraint CSR is NULL");
    for (int r = 0; r < m.n_restraints; ++r)
      if (m.rst_offsets[r] > m.rst_offsets[r + 1]) return fail(LD_EINVAL, w + ": restraint offsets not monotone");
    for (int k = m.rst_offsets[0]; k < m.rst_offsets[m.n_restraints]; ++k)
      if (m.rst_atoms[k] < 0 || m.rst_atoms[k] >= m.n_atoms)
        return fail(LD_EINVAL, w + ": restraint atom index out of range");
  }
  for (int k = 0; k < m.n_membrane; ++k)
    if (!m.membrane || m.membrane[k] < 0 || m.membrane[k] >= m.n_atoms)
      return fail(LD_EINVAL, w + ": membrane index out of range");
  return LD_OK;
}

struct SortedMol {
  std::vector<int> perm, inv;
  std::vector<double> x, y, z, q, eps, rad, modes;
  std::vector<int> toff;
  std::vector<unsigned short> tb20;
  std::vector<int> rst_off, rst_idx, mem_idx;
  int n = 0, n_pad = 0, n_tiles = 0;
};
static SortedMol sort_molecule(const ld_molecule_desc &m, int method, int tile, double pad, int n_modes_eff,
                               bool compact) {
  SortedMol s;
  s.n = m.n_atoms;
  s.n_tiles = (m.n_atoms + tile - 1) / tile;
  s.n_pad = s.n_tiles * tile;
  s.perm = spatial_order(m.coords, m.n_atoms, tile, compact);
  s.inv.assign(m.n_atoms, 0);
  for (int i = 0; i < m.n_atoms; ++i) s.inv[s.perm[i]] = i;
  s.x.assign(s.n_pad, pad); s.y.assign(s.n_pad, pad); s.z.assign(s.n_pad, pad);
  for (int i = 0; i < s.n; ++i) {
    const int o = s.perm[i];
    s.x[i] = m.coords[3 * o]; s.y[i] = m.coords[3 * o + 1]; s.z[i] = m.coords[3 * o + 2];
  }
  if (method == LD_METHOD_DFIRE) {
    s.toff.assign(s.n_pad, 0);
    s.tb20.assign(s.n_pad, 0);
    for (int i = 0; i < s.n; ++i) {
      s.toff[i] = m.dfire_type[s.perm[i]] * DFIRE_ROW;
      s.tb20[i] = (unsigned short)(m.dfire_type[s.perm[i]] * 20);
    }
  } else {
    s.q.assign(s.n_pad, 0.0); s.eps.assign(s.n_pad, 0.0); s.rad.assign(s.n_pad, 0.0);
    for (int i = 0; i < s.n; ++i) {
      const int o = s.perm[i];
      s.q[i] = m.ele_charge[o]; s.eps[i] = m.vdw_energy[o]; s.rad[i] = m.vdw_radius[o];
    }
  }
  if (n_modes_eff > 0) {  // [k][atom][3] -> [k][3][sorted atom]
    s.modes.assign((size_t)n_modes_eff * 3 * s.n_pad, 0.0);
    for (int k = 0; k < n_modes_eff; ++k)
      for (int i = 0; i < s.n; ++i)
        for (int d = 0; d < 3; ++d)
          s.modes[((size_t)k * 3 + d) * s.n_pad + i] = m.modes[((size_t)k * m.n_atoms + s.perm[i]) * 3 + d];
  }
  s.rst_off.assign(1, 0);
  for (int r = 0; r < m.n_restraints; ++r) {
    for (int k = m.rst_offsets[r]; k < m.rst_offsets[r + 1]; ++k) s.rst_idx.push_back(s.inv[m.rst_atoms[k]]);
    s.rst_off.push_back((int)s.rst_idx.size());
  }
  for (int k = 0; k < m.n_membrane; ++k) s.mem_idx.push_back(s.inv[m.membrane[k]]);
  return s;
}

extern "C" int ld_destroy(ld_handle *h) {
  if (!h) return LD_OK;
  cudaSetDevice(h->device);
  for (Workspace &w : h->ws) {
    if (w.stream) cudaStreamSynchronize(w.stream);
    cudaFree(w.d_poses); cudaFree(w.d_energies); cudaFree(w.d_detail);
    cudaFree(w.d_lig_blocks); cudaFree(w.d_rec_blocks); cudaFree(w.d_partials);
    cudaFree(w.d_iface_rec); cudaFree(w.d_iface_lig); cudaFree(w.d_unit_counter); cudaFree(w.d_prep);
    cudaFree(w.d_lig4p); cudaFree(w.d_flag);
    cudaFreeHost(w.h_poses); cudaFreeHost(w.h_energies);
    for (cudaEvent_t e : w.prof_events) cudaEventDestroy(e);
    if (w.ev0) cudaEventDestroy(w.ev0);
    if (w.ev1) cudaEventDestroy(w.ev1);
    if (w.done) cudaEventDestroy(w.done);
    if (w.stream) cudaStreamDestroy(w.stream);
  }
  for (void *p : h->owned) cudaFree(p);
  cudaFree(h->gso_slab);
  cudaFree(h->d_rc);
  cudaFree(h->d_cells); cudaFree(h->d_cell_tiles); cudaFree(h->d_tile_slack); cudaFree(h->d_need);
  cudaFreeHost(h->h_need);
  delete h;
  return LD_OK;
}

// ---------------------------------------------------------------------------------------------
// Rigid-ligand DFIRE path (ld_rigid.cuh): everything below is built once per complex.
//
// Receptor atoms are packed into groups of <= 32 atoms spanning <= rows_max DFIRE types: types with
// >= 32 atoms fill whole groups on their own, the remainders are packed first-fit-decreasing.
struct RigidGroup {
  std::vector<int> atoms, types;
};
static std::vector<RigidGroup> pack_groups(const ld_molecule_desc &m, int rows_max) {
  std::vector<std::vector<int>> by_type(169);
  for (int i = 0; i < m.n_atoms; ++i) by_type[m.dfire_type[i]].push_back(i);
  std::vector<RigidGroup> groups;
  std::vector<std::pair<int, std::vector<int>>> pieces;  // (type, atoms) with < 32 atoms
  for (int t = 0; t < 169; ++t) {
    const std::vector<int> &a = by_type[t];
    size_t pos = 0;
    for (; a.size() - pos >= 32; pos += 32) {
      RigidGroup g;
      g.atoms.assign(a.begin() + pos, a.begin() + pos + 32);
      g.types.push_back(t);
      groups.push_back(std::move(g));
    }
    if (pos < a.size()) pieces.emplace_back(t, std::vector<int>(a.begin() + pos, a.end()));
  }
  std::stable_sort(pieces.begin(), pieces.end(),
                   [](const auto &x, const auto &y) { return x.second.size() > y.second.size(); });
  const size_t first_open = groups.size();
  for (auto &pc : pieces) {
    size_t k = first_open;
    for (; k < groups.size(); ++k)
      if (groups[k].atoms.size() + pc.second.size() <= 32 && (int)groups[k].types.size() < rows_max) break;
    if (k == groups.size()) groups.emplace_back();
    groups[k].atoms.insert(groups[k].atoms.end(), pc.second.begin(), pc.second.end());
    groups[k].types.push_back(pc.first);
  }
  return groups;
}

// Ligand-frame cell lists: uniform grid over the ligand's bounding box grown by the cut-off (+ the largest tile slack);
// a cell lists every ligand tile with an atom within 15 A + the tile's slack + 0.01 of the cell's box (0.01: f32 cell
// assignment + the classification margin delta).  Slack is 0 for a rigid ligand; for a ligand with ANM modes it is how
// far the tile's atoms may move in the ligand frame before a pose must be scored by brute force (ld_rigid.cuh, FLEX),
// learnt from the poses the handle has seen.  Called by ld_create and again whenever the slacks grow.
static int build_cells(ld_handle *h) {
  const auto t_cells = std::chrono::steady_clock::now();
  struct Stamp { ld_handle *h; std::chrono::steady_clock::time_point t; ~Stamp() { h->create_ms[3] = ms_since(t); } } stamp_{h, t_cells};
  const DeviceComplex &cx = h->cx;
  RigidComplex &rc = h->rc;
  const std::vector<double> &LX = h->lig_sx, &LY = h->lig_sy, &LZ = h->lig_sz;
  double cell = g_opt.cell_size;
  if (cell <= 0.0) {
    // Finer cells = tighter lists (fewer pair tests that cannot be in range), as long as the grid and its lists stay
    // well inside L2: ~27,000 A^3 of cells list a tile (its atoms' 15 A spheres), 2 bytes per entry, 8 bytes per cell of
    // the ligand's box grown by the reach.  The finest of 0.6 / 0.75 / 0.85 / 1 A whose estimate stays below 50 MB:
    // 1k4c (409 tiles) 0.85 A, 1ppe (28 tiles) 0.6 A.  Measured (profiles/r2_rigid_ab_run5_cell.txt): 1k4c 28.06 ms per
    // 80,000 poses at 1 A, 27.81 at 0.85, 27.78 at 0.75 -- but with 59 MB of lists at 0.75 A the L2 hit rate of the list
    // reads falls from 94 % to 85 % and the launch moves 1.4 GB through DRAM instead of 0.45 -- 28.14 at 0.6; 1ppe 2.13 /
    // 2.09 / 2.07 ms at 1 / 0.75 / 0.6 A.  The FLEX instance keeps 1 A (2uuy: 3.95 / 3.98 / 4.07 ms).
    cell = 1.0;
    if (!h->flex) {
      double vol = 1.0;
      {
        double lo3[3] = {1e300, 1e300, 1e300}, hi3[3] = {-1e300, -1e300, -1e300};
        const std::vector<double> *C3[3] = {&LX, &LY, &LZ};
        for (int j = 0; j < cx.n_lig; ++j)
          for (int d = 0; d < 3; ++d) {
            lo3[d] = std::min(lo3[d], (*C3[d])[j]);
            hi3[d] = std::max(hi3[d], (*C3[d])[j]);
          }
        for (int d = 0; d < 3; ++d) vol *= hi3[d] - lo3[d] + 30.1;
      }
      for (double c : {0.6, 0.75, 0.85}) {
        const double est = ((double)cx.n_lig_tiles * 27000.0 * 2.0 + vol * 8.0) / (c * c * c);
        if (est < 50.0e6) { cell = c; break; }
      }
    }
  }
  double max_slack = 0.0;
  for (float v : h->tile_slack) max_slack = std::max(max_slack, (double)v);
  const double base_reach = 15.0 + 0.01, reach_max = base_reach + max_slack;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  const std::vector<double> *LC[3] = {&LX, &LY, &LZ};
  for (int j = 0; j < cx.n_lig; ++j)
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], (*LC[d])[j]);
      hi[d] = std::max(hi[d], (*LC[d])[j]);
    }
  float g0[3];
  int nc[3];
  float inv_h = 0.f;
  double hh = 0.0, maxabs = 0.0;
  size_t ncell = 0;
  // 1 A cells by default; a ligand so large that the grid would pass 2^24 cells (~270 MB of host scratch, ~130 MB
  // of offsets on the device) gets proportionally coarser cells instead of losing the fast path
  for (;; cell *= 1.25) {
    inv_h = (float)(1.0 / cell);
    hh = 1.0 / (double)inv_h;  // the cell size the device's (f - g0) * inv_h implies
    maxabs = 0.0;
    for (int d = 0; d < 3; ++d) {
      g0[d] = (float)(lo[d] - reach_max - 0.05);
      nc[d] = (int)std::floor((hi[d] + reach_max + 0.05 - (double)g0[d]) / hh) + 1;
      maxabs = std::max(maxabs, std::max(std::fabs((double)g0[d]), std::fabs((double)g0[d] + nc[d] * hh)));
    }
    ncell = (size_t)nc[0] * nc[1] * nc[2];
    if (ncell <= ((size_t)1 << 24) || cell > 16.0) break;
  }
  if (ncell > ((size_t)1 << 24)) { h->rigid_info = "rigid path off: cell grid too large"; h->rigid_ok = false; return LD_OK; }
  const double delta = 2.0e-4 + 1.3e-5 * maxabs;  // 2x the |d2f - dist_ref| bound derived at rigid_row()
  if (!(delta < 0.01)) {
    h->rigid_info = "rigid path off: ligand extent makes the FP32 margin too wide";
    h->rigid_ok = false;
    return LD_OK;
  }
  size_t total = 0, longest = 0, nonempty = 0;
  std::vector<uint2> cells;
  std::vector<unsigned short> flat;
  uint2 *d_cells_new = nullptr;
  unsigned short *d_flat_new = nullptr;
  if (!g_opt.cells_on_host) {
    // ---- device builder (ld_cells.cuh): count, scan, fill ----
    CellGrid grid{};
    for (int d = 0; d < 3; ++d) { grid.g0[d] = g0[d]; grid.nc[d] = nc[d]; }
    grid.hh = hh;
    grid.base_reach = base_reach;
    std::vector<TileBox> boxes(cx.n_lig_tiles);
    for (int t = 0; t < cx.n_lig_tiles; ++t) {
      TileBox &b = boxes[t];
      for (int d = 0; d < 3; ++d) { b.lo[d] = 3.0e38f; b.hi[d] = -3.0e38f; }
      for (int j = t * LIG_TILE; j < std::min((t + 1) * LIG_TILE, cx.n_lig); ++j)
        for (int d = 0; d < 3; ++d) {
          const double v = (*LC[d])[j];
          b.lo[d] = std::min(b.lo[d], std::nextafter((float)v, -3.0e38f));
          b.hi[d] = std::max(b.hi[d], std::nextafter((float)v, 3.0e38f));
        }
    }
    struct Tmp {  // scratch freed on every exit path
      TileBox *boxes = nullptr; float *slack = nullptr; unsigned *counts = nullptr, *stats = nullptr;
      unsigned long long *sums = nullptr;
      ~Tmp() { cudaFree(boxes); cudaFree(slack); cudaFree(counts); cudaFree(stats); cudaFree(sums); }
    } tmp;
    const int n_blocks = (int)((ncell + (size_t)SCAN_THREADS * SCAN_ITEMS - 1) / ((size_t)SCAN_THREADS * SCAN_ITEMS));
    CU(cudaMalloc(reinterpret_cast<void **>(&tmp.boxes), boxes.size() * sizeof(TileBox)));
    CU(cudaMalloc(reinterpret_cast<void **>(&tmp.slack), h->tile_slack.size() * sizeof(float)));
    CU(cudaMalloc(reinterpret_cast<void **>(&tmp.counts), ncell * sizeof(unsigned)));
    CU(cudaMalloc(reinterpret_cast<void **>(&tmp.stats), 2 * sizeof(unsigned)));
    CU(cudaMalloc(reinterpret_cast<void **>(&tmp.sums), ((size_t)n_blocks + 1) * sizeof(unsigned long long)));
    CU(cudaMalloc(reinterpret_cast<void **>(&d_cells_new), ncell * sizeof(uint2)));
    CU(cudaMemcpy(tmp.boxes, boxes.data(), boxes.size() * sizeof(TileBox), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(tmp.slack, h->tile_slack.data(), h->tile_slack.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemset(tmp.stats, 0, 2 * sizeof(unsigned)));
    const unsigned wgrid = (unsigned)((ncell + 127) / 128);
    cells_walk_kernel<false><<<wgrid, 128>>>(grid, cx.lig_x, cx.lig_y, cx.lig_z, cx.n_lig, cx.n_lig_tiles, tmp.boxes,
                                             tmp.slack, tmp.counts, nullptr, nullptr);
    cells_scan_blocks_kernel<<<n_blocks, SCAN_THREADS>>>(tmp.counts, ncell, tmp.sums);
    cells_scan_sums_kernel<<<1, 1024>>>(tmp.sums, n_blocks);
    cells_scan_write_kernel<<<n_blocks, SCAN_THREADS>>>(tmp.counts, ncell, tmp.sums, d_cells_new, tmp.stats);
    unsigned long long tot = 0;
    unsigned st[2] = {0, 0};
    CU(cudaMemcpy(&tot, tmp.sums + n_blocks, sizeof tot, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(st, tmp.stats, sizeof st, cudaMemcpyDeviceToHost));
    total = (size_t)tot; nonempty = st[0]; longest = st[1];
    if (total >= ((size_t)1 << 31)) {
      cudaFree(d_cells_new);
      h->rigid_info = "rigid path off: cell lists too large"; h->rigid_ok = false; return LD_OK;
    }
    CU(cudaMalloc(reinterpret_cast<void **>(&d_flat_new), std::max<size_t>(total, 1) * sizeof(unsigned short)));
    cells_walk_kernel<true><<<wgrid, 128>>>(grid, cx.lig_x, cx.lig_y, cx.lig_z, cx.n_lig, cx.n_lig_tiles, tmp.boxes,
                                            tmp.slack, nullptr, d_cells_new, d_flat_new);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
  } else {
  // Two passes (count, fill) over z-slabs of the grid, one host thread per slab: a slab owns its cells, so
  // the passes need no synchronisation, and tiles are visited in ascending order, so every list is sorted.
  std::vector<unsigned> count(ncell, 0u);
  std::vector<int> stamp(ncell, -1);
  cells.assign(ncell, make_uint2(0u, 0u));
  auto sweep = [&](int z_lo, int z_hi, bool fill) {
    for (int t = 0; t < cx.n_lig_tiles; ++t) {
      const double reach = base_reach + (double)h->tile_slack[t], reach2 = reach * reach;
      for (int j = t * LIG_TILE; j < std::min((t + 1) * LIG_TILE, cx.n_lig); ++j) {
        // the f32 value the kernel uses and the f64 one differ by < 1e-5: inside the slack
        const double a[3] = {LX[j], LY[j], LZ[j]};
        int c0[3], c1[3];
        for (int d = 0; d < 3; ++d) {
          c0[d] = std::max(0, (int)std::floor((a[d] - reach - (double)g0[d]) / hh));
          c1[d] = std::min(nc[d] - 1, (int)std::floor((a[d] + reach - (double)g0[d]) / hh));
        }
        for (int cz = std::max(c0[2], z_lo); cz <= std::min(c1[2], z_hi - 1); ++cz) {
          const double bz0 = (double)g0[2] + cz * hh, ez = std::max(0.0, std::max(bz0 - a[2], a[2] - (bz0 + hh)));
          for (int cy = c0[1]; cy <= c1[1]; ++cy) {
            const double by0 = (double)g0[1] + cy * hh, ey = std::max(0.0, std::max(by0 - a[1], a[1] - (by0 + hh)));
            const double eyz = ez * ez + ey * ey;
            if (eyz > reach2) continue;
            size_t c = ((size_t)cz * nc[1] + cy) * nc[0] + c0[0];
            for (int cxx = c0[0]; cxx <= c1[0]; ++cxx, ++c) {
              const double bx0 = (double)g0[0] + cxx * hh, ex = std::max(0.0, std::max(bx0 - a[0], a[0] - (bx0 + hh)));
              if (ex * ex + eyz > reach2 || stamp[c] == t) continue;
              stamp[c] = t;
              if (fill) flat[cells[c].x + count[c]] = (unsigned short)t;
              ++count[c];
            }
          }
        }
      }
    }
  };
  const int n_thr = std::max(1, std::min(std::min(16, nc[2]), (int)std::thread::hardware_concurrency()));
  auto run_pass = [&](bool fill) {
    std::vector<std::thread> pool;
    for (int w = 0; w < n_thr; ++w)
      pool.emplace_back(sweep, (int)((long)nc[2] * w / n_thr), (int)((long)nc[2] * (w + 1) / n_thr), fill);
    for (auto &th : pool) th.join();
  };
  run_pass(false);
  for (size_t c = 0; c < ncell; ++c) {
    cells[c] = make_uint2((unsigned)total, count[c]);
    total += count[c];
    longest = std::max<size_t>(longest, count[c]);
    nonempty += count[c] != 0;
  }
  if (total >= ((size_t)1 << 31)) { h->rigid_info = "rigid path off: cell lists too large"; h->rigid_ok = false; return LD_OK; }
  flat.assign(std::max<size_t>(total, 1), 0);
  std::fill(count.begin(), count.end(), 0u);
  std::fill(stamp.begin(), stamp.end(), -1);
  run_pass(true);
    CU(cudaMalloc(reinterpret_cast<void **>(&d_cells_new), cells.size() * sizeof(uint2)));
    CU(cudaMalloc(reinterpret_cast<void **>(&d_flat_new), flat.size() * sizeof(unsigned short)));
    CU(cudaMemcpy(d_cells_new, cells.data(), cells.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_flat_new, flat.data(), flat.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
  }

  // replace the device copies (a rebuild happens only with every stream of the handle idle: grow_flex_slack)
  cudaFree(h->d_cells); cudaFree(h->d_cell_tiles);
  h->d_cells = d_cells_new; h->d_cell_tiles = d_flat_new;
  if (h->d_tile_slack)
    CU(cudaMemcpy(h->d_tile_slack, h->tile_slack.data(), h->tile_slack.size() * sizeof(float), cudaMemcpyHostToDevice));
  rc.cells = h->d_cells; rc.cell_tiles = h->d_cell_tiles;
  rc.gx0 = g0[0]; rc.gy0 = g0[1]; rc.gz0 = g0[2]; rc.inv_h = inv_h;
  rc.nx = nc[0]; rc.ny = nc[1]; rc.nz = nc[2];
  rc.thr_out = (float)(225.0 + delta);
  rc.half_minus_eps = (float)(0.5 - 2.5e-5);
  rc.delta = (float)(1.02 * delta);
  rc.grid_maxabs = (float)(maxabs * 1.000001);
  if (h->d_rc) CU(cudaMemcpy(h->d_rc, &rc, sizeof(RigidComplex), cudaMemcpyHostToDevice));
  char buf[640];
  snprintf(buf, sizeof buf,
           "rigid path on%s: %d receptor groups (%.1f atoms/group, <=%d table rows each), cell %.2f A, grid %dx%dx%d, "
           "%zu non-empty cells, %zu list entries (longest %zu), delta %.2e, smem %zu B",
           h->flex ? " (flexible ligand: per-pose ligand blocks, slack lists, fixed-point sums)" : "", rc.n_groups,
           (double)cx.n_rec / rc.n_groups, rc.rows_max, hh, nc[0], nc[1], nc[2], nonempty, total, longest, delta,
           rigid_smem_bytes(cx.n_lig_pad, rc.rows_max, rc.row_bytes, h->flex ? h->flex_warps : 1));
  h->rigid_info = buf;
  if (h->flex) {
    snprintf(buf, sizeof buf, ", %d warps per CTA, tile slack max %.2f A (rebuilt %d times)", h->flex_warps, max_slack,
             h->flex_rebuilds);
    h->rigid_info += buf;
  }
  return LD_OK;
}

static int build_rigid(const ld_complex_desc *desc, ld_handle *h, const SortedMol &L) {
  const DeviceComplex &cx = h->cx;
  const ld_molecule_desc &R = desc->receptor;
  h->rigid_ok = false;
  h->flex = false;
  if (cx.method != 0) { h->rigid_info = "rigid path off: not DFIRE"; return LD_OK; }
  if (cx.n_rec == 0 || cx.n_lig == 0) { h->rigid_info = "rigid path off: empty partner"; return LD_OK; }
  if (cx.n_lig_tiles > 65535) { h->rigid_info = "rigid path off: ligand tile ids exceed 16 bits"; return LD_OK; }
  // FLEX: a table row holds only the DFIRE types the ligand has (1czy: 36 of 169, 2uuy 140): smaller rows leave room for
  // more rows per receptor group (fuller groups, fewer (group, pose) tasks: 1czy 48 -> 42 groups) or for more warps next to
  // them (2uuy 16 -> 18), and there is less to copy on a group switch.  The rigid instance keeps full rows of a
  // compile-time size: its ligands are large (1k4c has 167 of the types) and a run-time row size costs its code
  // generation 1 % (profiles/r2_flex_latency_ab.txt, run 52).
  std::vector<int> ctype(169, -1);
  int n_ct = 0;
  if (cx.n_lig_modes > 0) {
    std::vector<char> present(169, 0);
    for (int j = 0; j < cx.n_lig; ++j) present[L.tb20[j] / 20] = 1;
    for (int t = 0; t < 169; ++t)
      if (present[t]) ctype[t] = n_ct++;
  } else {
    for (int t = 0; t < 169; ++t) ctype[t] = t;
    n_ct = 169;
  }
  const int row_bytes = cx.n_lig_modes > 0 ? n_ct * RG_TB_BYTES : RG_ROW_BYTES;  // multiples of 16 (RG_TB_BYTES = 240)
  int rows_max = 0;
  if (cx.n_lig_modes == 0) {
    const long avail = (long)h->max_smem_optin - (long)rigid_smem_bytes(cx.n_lig_pad, 0, row_bytes);
    rows_max = (int)std::min<long>(4, avail / row_bytes);  // rigid: as measured in round 2 (4 x 40,560 B next to the ligand)
    rows_max = std::min(rows_max, g_opt.rigid_rows);
    if (rows_max < 1) { h->rigid_info = "rigid path off: ligand + one table row exceed shared memory"; return LD_OK; }
  } else if (!g_opt.flex) {
    h->rigid_info = "rigid path off: the ligand has ANM modes (FLEX disabled by ld_set_option)";
    return LD_OK;
  } else {
    // FLEX: one ligand block per warp next to the table rows; prefer many rows (fewer, fuller receptor groups) as
    // long as enough warps fit to keep the SM busy -- the FLEX instance is latency bound, so warps come first.  Measured
    // (20,000 poses, profiles/r2_flex_latency_ab.txt run 53): 2uuy 4 rows / 14 warps 4.18 ms, 3 / 19 3.57 ms, 2 / 20 3.96 ms;
    // ab_icode 6 / 17 4.34 ms, 5 / 20 3.93 ms; 1czy fits 8 rows and 20 warps (1.55 ms; 4 rows: 1.63 ms)
    const long lig_bytes = (long)cx.n_lig_pad * 16;
    int best_w = 0;
    for (int r = std::min(RG_MAX_ROWS, g_opt.rigid_rows); r >= 1 && rows_max == 0; --r) {
      const long avail = (long)h->max_smem_optin - 128 - (long)r * row_bytes;
      const int w = (int)std::min<long>(RG_WARPS, avail / lig_bytes);
      if (w >= g_opt.flex_min_warps || (r <= 2 && w >= 8)) { rows_max = r; best_w = w; }
    }
    if (rows_max == 0) {
      h->rigid_info = "rigid path off: the ligand has ANM modes and its per-warp blocks do not fit in shared memory";
      return LD_OK;
    }
    h->flex = true;
    h->flex_warps = best_w;
  }

  RigidComplex &rc = h->rc;
  rc = RigidComplex{};
  std::vector<RigidGroup> groups = pack_groups(R, rows_max);
  // most expensive first: protein atoms meet the ligand, membrane beads (type 167) rarely do
  std::vector<int> order(groups.size());
  std::iota(order.begin(), order.end(), 0);
  auto weight = [&](int g) {
    int w = 0;
    for (int a : groups[g].atoms) w += R.dfire_type[a] == 167 ? 1 : 8;
    return w;
  };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight(a) > weight(b); });

  const int ng = (int)groups.size(), npos = ng * 32;
  const int nrm = cx.n_rec_modes;
  std::vector<double> x(npos, REC_PAD), y(npos, REC_PAD), z(npos, REC_PAD);
  std::vector<int> slot(npos, -1), toff(npos, 0), gtypes((size_t)ng * RG_MAX_ROWS, -1), inv(R.n_atoms, -1);
  std::vector<double> modes((size_t)nrm * 3 * npos, 0.0);
  h->rec_perm_r.assign(npos, -1);
  for (int g = 0; g < ng; ++g) {
    for (size_t r = 0; r < groups[g].types.size(); ++r) gtypes[(size_t)g * RG_MAX_ROWS + r] = groups[g].types[r];
    for (size_t k = 0; k < groups[g].atoms.size(); ++k) {
      const int a = groups[g].atoms[k], pos = g * 32 + (int)k;
      x[pos] = R.coords[3 * a]; y[pos] = R.coords[3 * a + 1]; z[pos] = R.coords[3 * a + 2];
      const int ty = R.dfire_type[a];
      slot[pos] = (int)(std::find(groups[g].types.begin(), groups[g].types.end(), ty) - groups[g].types.begin());
      toff[pos] = ty * DFIRE_ROW;
      inv[a] = pos;
      h->rec_perm_r[pos] = a;
      for (int m = 0; m < nrm; ++m)
        for (int d = 0; d < 3; ++d)
          modes[((size_t)m * 3 + d) * npos + pos] = R.modes[((size_t)m * R.n_atoms + a) * 3 + d];
    }
  }
  std::vector<int> rst_idx, mem_idx;
  for (int r = 0; r < R.n_restraints; ++r)
    for (int k = R.rst_offsets[r]; k < R.rst_offsets[r + 1]; ++k) rst_idx.push_back(inv[R.rst_atoms[k]]);
  for (int k = 0; k < R.n_membrane; ++k) mem_idx.push_back(inv[R.membrane[k]]);
  // the receptor positions whose interface flag finalize_kernel reads (active restraints, membrane beads)
  std::vector<unsigned> gneed(ng, 0u);
  for (int pos : rst_idx) gneed[pos >> 5] |= 1u << (pos & 31);
  for (int pos : mem_idx) gneed[pos >> 5] |= 1u << (pos & 31);

  // ligand, local frame, f32 + column offset
  std::vector<float4> l4(cx.n_lig_pad);
  for (int j = 0; j < cx.n_lig_pad; ++j) {
    if (j < cx.n_lig) {
      l4[j] = make_float4((float)L.x[j], (float)L.y[j], (float)L.z[j], (float)(ctype[L.tb20[j] / 20] * RG_SLOTS));
    } else {
      l4[j] = make_float4(1.0e6f, 1.0e6f, 1.0e6f, 0.f);
    }
  }
  h->lig_sx = L.x; h->lig_sy = L.y; h->lig_sz = L.z;
  h->tile_slack.assign(cx.n_lig_tiles, 0.f);

  int rcode;
#define UPR(vec, field) \
  if ((rcode = upload(h, vec, &rc.field)) != LD_OK) return rcode
  UPR(x, rec_x); UPR(y, rec_y); UPR(z, rec_z); UPR(slot, rec_slot); UPR(toff, rec_toff);
  UPR(gtypes, group_types); UPR(order, group_order); UPR(modes, rec_modes);
  UPR(l4, lig4); UPR(gneed, group_need);
#undef UPR
  rc.lig_need = desc->ligand.n_restraints > 0 ? 1 : 0;
  // the table re-indexed by the truncated bin-space value idx = -1..28 (as cx.potx, create_impl) for the ligand's types only
  const size_t row8 = (size_t)row_bytes / 8;
  std::vector<double> potx(169 * row8, 0.0);
  for (int ta = 0; ta < 169; ++ta)
    for (int tb = 0; tb < 169; ++tb) {
      if (ctype[tb] < 0) continue;
      for (int sidx = 0; sidx < RG_SLOTS; ++sidx) {
        const int idx = sidx + RG_SLOT0;
        if (idx > 28) continue;
        const int bin = idx <= 2 ? 0 : (idx <= 15 ? idx - 2 : 13 + ((idx - 15) >> 1));
        potx[ta * row8 + (size_t)ctype[tb] * RG_SLOTS + sidx] = desc->dfire_potential[(size_t)ta * DFIRE_ROW + tb * 20 + bin];
      }
    }
  if ((rcode = upload(h, potx, &rc.potx)) != LD_OK) return rcode;
  rc.n_groups = ng; rc.n_rec_pos = npos;
  rc.n_lig = cx.n_lig; rc.n_lig_pad = cx.n_lig_pad; rc.n_lig_tiles = cx.n_lig_tiles;
  rc.n_rec_modes = nrm; rc.pose_len = cx.pose_len; rc.rows_max = rows_max; rc.row_bytes = row_bytes;
  rc.lig_x = cx.lig_x; rc.lig_y = cx.lig_y; rc.lig_z = cx.lig_z; rc.lig_tb20 = cx.lig_tb20; rc.pot = cx.pot;
  rc.flex = h->flex ? 1 : 0;
  rc.n_lig_modes = cx.n_lig_modes;
  rc.lig_modes = cx.lig_modes;

  // what finalize_kernel sees: groups play the role of receptor tiles
  h->cxr = cx;
  h->cxr.n_rec_tiles = ng;
  h->cxr.n_rec_pad = npos;
  if ((rcode = upload(h, rst_idx, &h->cxr.rec_rst_idx)) != LD_OK) return rcode;
  if ((rcode = upload(h, mem_idx, &h->cxr.membrane_idx)) != LD_OK) return rcode;
  if (h->flex) {
    // fixed-point copy of the re-indexed table: scale 2^k with 32 atoms x n_lig_pad pairs x max|value| x 2^k < 2^61, so
    // that the sum over one (receptor group, pose) cannot overflow; finalize_kernel converts each group's sum back
    // and adds the groups in order
    const size_t n_potx = potx.size();
    double vmax = 1.0;
    for (double v : potx) vmax = std::max(vmax, std::fabs(v));
    const int k = 61 - (int)std::ceil(std::log2(32.0 * cx.n_lig_pad * vmax));
    if (k < 20) { h->rigid_info = "rigid path off: table values too large for the fixed-point sums"; h->flex = false; return LD_OK; }
    rc.fx_scale = std::ldexp(1.0, k);
    std::vector<long long> fx(n_potx);
    for (size_t i = 0; i < n_potx; ++i) fx[i] = std::llrint(potx[i] * rc.fx_scale);
    if ((rcode = upload(h, fx, &rc.potx_fx)) != LD_OK) return rcode;
    h->cxr.fx_inv_scale = 1.0 / rc.fx_scale;
    CU(cudaMalloc(reinterpret_cast<void **>(&h->d_tile_slack), cx.n_lig_tiles * sizeof(float)));
    CU(cudaMalloc(reinterpret_cast<void **>(&h->d_need), cx.n_lig_tiles * sizeof(int)));
    CU(cudaMemset(h->d_need, 0, cx.n_lig_tiles * sizeof(int)));
    CU(cudaHostAlloc(reinterpret_cast<void **>(&h->h_need), cx.n_lig_tiles * sizeof(int), cudaHostAllocDefault));
    rc.tile_slack = h->d_tile_slack;
    CU(cudaFuncSetAttribute(dfire_rigid_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
    CU(cudaFuncSetAttribute(dfire_rigid_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  } else {
    CU(cudaFuncSetAttribute(dfire_rigid_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
    CU(cudaFuncSetAttribute(dfire_rigid_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  }
  CU(cudaMalloc(reinterpret_cast<void **>(&h->d_rc), sizeof(RigidComplex)));
  h->rigid_ok = true;
  return build_cells(h);  // fills the grid fields of rc, uploads d_rc, clears rigid_ok if the grid cannot be built
}

// FLEX: after a call has completed, grow the tile slacks to what the poses seen so far need (with head-room, so a
// drifting swarm does not trigger a rebuild per step) and rebuild the cell lists.  Every stream of the handle must be
// idle: the lists are replaced in place.  Results never depend on this: a pose whose tiles exceed the slacks is scored
// by brute force in the same launch, and FLEX sums are exact integers.
static int grow_flex_slack(ld_handle *h) {
  if (!h->flex || !h->rigid_ok) return LD_OK;
  bool grow = false;
  for (int t = 0; t < h->cx.n_lig_tiles; ++t) {
    float need;
    std::memcpy(&need, &h->h_need[t], sizeof need);
    if (need > h->tile_slack[t]) grow = true;
  }
  if (!grow) return LD_OK;
  for (Workspace &w : h->ws)
    if (w.stream) CU(cudaStreamSynchronize(w.stream));
  CU(cudaDeviceSynchronize());  // device-API calls may sit on caller streams
  for (int t = 0; t < h->cx.n_lig_tiles; ++t) {
    float need;
    std::memcpy(&need, &h->h_need[t], sizeof need);
    if (need > h->tile_slack[t]) h->tile_slack[t] = need * 1.25f + 0.25f;
  }
  ++h->flex_rebuilds;
  return build_cells(h);
}

extern "C" int ld_init_device(int32_t device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(LD_ECUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                              (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= ndev) return fail(LD_EINVAL, "device ordinal out of range");
  CU(cudaSetDevice(device));
  CU(cudaFree(nullptr));  // forces the primary context into existence
  return LD_OK;
}

extern "C" int ld_get_create_ms(const ld_handle *h, double *out4) {
  if (!h || !out4) return fail(LD_EINVAL, "ld_get_create_ms: NULL argument");
  for (int i = 0; i < 4; ++i) out4[i] = h->create_ms[i];
  return LD_OK;
}

static int create_impl(const ld_complex_desc *desc, ld_handle *h) {
  const auto t_sort = std::chrono::steady_clock::now();
  const int method = desc->method == LD_METHOD_DFIRE ? 0 : 1;
  h->device = desc->device;
  h->use_anm = desc->use_anm ? 1 : 0;
  // Host-only work first: a driver that creates the CUDA context on a helper thread (ld_init_device) is still waiting
  // for it at this point, so sorting the atoms into compact tiles costs it nothing.  The ligand's tiles of 8 are always
  // compacted (they carry the culling of every path); the receptor's tiles of 32 only where a kernel uses them as
  // culling units (DNA/pyDock; the DFIRE ligand-frame path groups the receptor by type instead).
  const int nrm = h->use_anm ? desc->receptor.n_modes : 0, nlm = h->use_anm ? desc->ligand.n_modes : 0;
  SortedMol R = sort_molecule(desc->receptor, desc->method == LD_METHOD_DFIRE ? LD_METHOD_DFIRE : LD_METHOD_DNA,
                              REC_TILE, REC_PAD, nrm, method != 0);
  SortedMol L = sort_molecule(desc->ligand, desc->method == LD_METHOD_DFIRE ? LD_METHOD_DFIRE : LD_METHOD_DNA,
                              LIG_TILE, LIG_PAD, nlm, true);
  const double sort_ms = ms_since(t_sort);
  const auto t_start = std::chrono::steady_clock::now();
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(LD_ECUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                              (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (desc->device < 0 || desc->device >= ndev) return fail(LD_EINVAL, "device ordinal out of range");
  CU(cudaSetDevice(desc->device));
  CU(cudaFree(nullptr));
  h->create_ms[0] = ms_since(t_start);
  const auto t_complex = std::chrono::steady_clock::now();
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, desc->device));
  if (prop.major < 10)
    return fail(LD_ECUDA, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                              "; this library is built for sm_100a only");
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  h->sm_count = prop.multiProcessorCount;
  for (Workspace &w : h->ws) {
    CU(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&w.ev0));
    CU(cudaEventCreate(&w.ev1));
    CU(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
  }

  h->rec_perm = R.perm;
  h->lig_perm = L.perm;
  // DNA/pyDock: sqrt(eps_r * eps_l) (src/dna.rs:496) as sqrt(eps_r) * sqrt(eps_l) with the roots taken once here;
  // a negative vdw energy keeps the reference's form (the product of two negatives has a real root)
  // DNA/pyDock van der Waals term by type pair: ve (p6^2 - 2 p6) = x6 (A x6 - B) with x6 = 1/d^6,
  // A = sqrt(e_r e_l) (r_r + r_l)^12, B = 2 sqrt(e_r e_l) (r_r + r_l)^6 (src/dna.rs:495-499), tabulated over the distinct
  // (energy, radius) pairs of each partner; built from the caller's values before the roots are taken below
  std::vector<double2> vdw_tab;
  std::vector<int> rec_vt(R.n_pad, 0), lig_vt(L.n_pad, 0);
  int vdw_nr = 0, vdw_nl = 0;
  if (method != 0) {
    std::vector<std::pair<double, double>> tr, tl;
    auto type_of = [](std::vector<std::pair<double, double>> &types, double e, double r) {
      for (size_t k = 0; k < types.size(); ++k)
        if (types[k].first == e && types[k].second == r) return (int)k;
      types.emplace_back(e, r);
      return (int)types.size() - 1;
    };
    bool ok = true;
    for (int i = 0; i < R.n && ok; ++i) { rec_vt[i] = type_of(tr, R.eps[i], R.rad[i]); ok = tr.size() <= 1024; }
    for (int j = 0; j < L.n && ok; ++j) { lig_vt[j] = type_of(tl, L.eps[j], L.rad[j]); ok = tl.size() <= 1024; }
    if (ok && !tr.empty() && !tl.empty() && tr.size() * tl.size() <= 1024) {
      vdw_nr = (int)tr.size(); vdw_nl = (int)tl.size();
      vdw_tab.resize((size_t)vdw_nr * vdw_nl);
      for (int b = 0; b < vdw_nl; ++b)
        for (int a = 0; a < vdw_nr; ++a) {
          const double ve = std::sqrt(tr[a].first * tl[b].first), vr = tr[a].second + tl[b].second;
          const double vr2 = vr * vr, vr6 = vr2 * (vr2 * vr2);
          vdw_tab[(size_t)b * vdw_nr + a] = make_double2(ve * vr6 * vr6, 2.0 * ve * vr6);
        }
      for (int &v : rec_vt) v *= 16;
      for (int &v : lig_vt) v *= vdw_nr * 16;
    }
  }
  bool eps_nonneg = true;
  for (double v : R.eps) eps_nonneg = eps_nonneg && v >= 0.0;
  for (double v : L.eps) eps_nonneg = eps_nonneg && v >= 0.0;
  if (eps_nonneg) {
    for (double &v : R.eps) v = std::sqrt(v);
    for (double &v : L.eps) v = std::sqrt(v);
  }
  h->cx.vdw_sqrt_hoisted = eps_nonneg ? 1 : 0;
  {
    // a tile pair further apart than this holds no vdW / interface pair (10 A) and no Coulomb term that reaches the
    // +-4/332 clamp: |q_r q_l| / d2 > C needs d2 < 83 |q_r| |q_l| (the 1.001 covers the 1e-12 error of the kernel's quotient)
    double qr = 0.0, ql = 0.0;
    for (double v : R.q) qr = std::max(qr, std::fabs(v));
    for (double v : L.q) ql = std::max(ql, std::fabs(v));
    h->cx.dna_close_reach = (float)(std::max(10.0, std::sqrt(83.0 * 1.001 * qr * ql)) * 1.0001 + 1e-3);
  }
  h->rec_xyz_orig.assign(desc->receptor.coords, desc->receptor.coords + (size_t)3 * desc->receptor.n_atoms);

  DeviceComplex &cx = h->cx;
  cx.method = method;
  cx.n_rec = R.n; cx.n_lig = L.n;
  cx.n_rec_pad = R.n_pad; cx.n_lig_pad = L.n_pad;
  cx.n_rec_tiles = R.n_tiles; cx.n_lig_tiles = L.n_tiles;
  cx.n_rec_modes = nrm; cx.n_lig_modes = nlm;
  cx.pose_len = 7 + nrm + nlm;
  int rc;
#define UP(vec, field) \
  if ((rc = upload(h, vec, &cx.field)) != LD_OK) return rc
  UP(R.x, rec_x); UP(R.y, rec_y); UP(R.z, rec_z);
  UP(R.toff, rec_toff); UP(R.q, rec_q); UP(R.eps, rec_seps); UP(R.rad, rec_rad); UP(R.modes, rec_modes);
  UP(L.x, lig_x); UP(L.y, lig_y); UP(L.z, lig_z);
  UP(L.tb20, lig_tb20); UP(L.q, lig_q); UP(L.eps, lig_seps); UP(L.rad, lig_rad); UP(L.modes, lig_modes);
  UP(R.rst_off, rec_rst_off); UP(R.rst_idx, rec_rst_idx); UP(L.rst_off, lig_rst_off); UP(L.rst_idx, lig_rst_idx);
  UP(R.mem_idx, membrane_idx);
  cx.vdw_nr = vdw_nr; cx.vdw_nl = vdw_nl;
  if (!vdw_tab.empty()) { UP(vdw_tab, vdw_tab); UP(rec_vt, rec_vt); UP(lig_vt, lig_vt); }
  cx.n_rec_rst = desc->receptor.n_restraints;
  cx.n_lig_rst = desc->ligand.n_restraints;
  cx.n_membrane = desc->receptor.n_membrane;
  std::vector<float4> sph(R.n_tiles);
  for (int t = 0; t < R.n_tiles; ++t)
    sph[t] = tile_sphere(R.x.data(), R.y.data(), R.z.data(), t * REC_TILE, std::min((t + 1) * REC_TILE, R.n));
  UP(sph, rec_sphere);
  {
    double mx = 0.0;
    for (int i = 0; i < R.n; ++i) mx = std::max(mx, std::max(std::fabs(R.x[i]), std::max(std::fabs(R.y[i]), std::fabs(R.z[i]))));
    cx.rec_maxabs = (float)(mx * 1.0001 + 1.0);
  }
  if (method == 0) {
    std::vector<double> pot(desc->dfire_potential, desc->dfire_potential + LD_DFIRE_TABLE_LEN);
    UP(pot, pot);
    // the table re-indexed by the truncated bin-space value idx = -1..28 (DIST_TO_BINS applied here once,
    // src/dfire.rs:49-53,337; idx -1 = the saturated `d as usize` of a negative d = index 0)
    const size_t row8 = RG_ROW_BYTES / 8;
    std::vector<double> potx(169 * row8, 0.0);
    for (int ta = 0; ta < 169; ++ta)
      for (int tb = 0; tb < 169; ++tb)
        for (int sidx = 0; sidx < RG_SLOTS; ++sidx) {
          const int idx = sidx + RG_SLOT0;
          if (idx > 28) continue;  // padding slot (RG_SLOTS = 31): never indexed, stays 0
          const int bin = idx <= 2 ? 0 : (idx <= 15 ? idx - 2 : 13 + ((idx - 15) >> 1));
          potx[ta * row8 + (size_t)tb * RG_SLOTS + sidx] = pot[(size_t)ta * DFIRE_ROW + tb * 20 + bin];
        }
    UP(potx, potx);
    std::vector<unsigned> rowx(R.n_pad, 0u);
    for (int i = 0; i < R.n; ++i)
      rowx[i] = (unsigned)(R.toff[i] / DFIRE_ROW) * (unsigned)row8 - (unsigned)RG_SLOT0 - RG_MAGIC_BITS;
    UP(rowx, rec_rowx);
  }
#undef UP
  h->lig_block = lig_block_bytes(cx.n_lig_pad, cx.n_lig_tiles, cx.method);
  h->rec_block = nrm > 0 ? rec_block_bytes(cx.n_rec_pad, cx.n_rec_tiles) : 0;
  if (h->lig_block >= (1u << 20)) return fail(LD_ELIMIT, "ligand too large for one bulk copy");

  // the pair kernels need opt-in dynamic shared memory; check the worst case (one split) fits
  const int lig_words = (cx.n_lig_pad + 31) / 32;
  const size_t need = pair_smem_bytes(method, cx.n_lig_pad, cx.n_lig_tiles, lig_words, cx.n_rec_tiles, cx.vdw_nr * cx.vdw_nl);
  if (need > (size_t)h->max_smem_optin)
    return fail(LD_ELIMIT, "ligand of " + std::to_string(cx.n_lig) + " atoms needs " + std::to_string(need) +
                               " B of shared memory per CTA; limit is " + std::to_string(h->max_smem_optin));
  CU(cudaFuncSetAttribute(dfire_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CU(cudaFuncSetAttribute(dfire_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CU(cudaFuncSetAttribute(dna_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CU(cudaFuncSetAttribute(dna_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CU(cudaFuncSetAttribute(dna_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CU(cudaFuncSetAttribute(dna_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  h->path_mode = g_opt.default_path;
  h->dna_fused = method != 0 && g_opt.dna_fused && h->cx.n_rec_modes + h->cx.n_lig_modes <= 64;
  h->create_ms[1] = ms_since(t_complex) + sort_ms;
  const auto t_rigid = std::chrono::steady_clock::now();
  const int rrc = build_rigid(desc, h, L);
  h->create_ms[2] = ms_since(t_rigid) - h->create_ms[3];
  return rrc;
}

extern "C" int ld_create(const ld_complex_desc *desc, ld_handle **out) {
  if (!desc || !out) return fail(LD_EINVAL, "ld_create: NULL argument");
  *out = nullptr;
  if (desc->method != LD_METHOD_DFIRE && desc->method != LD_METHOD_DNA && desc->method != LD_METHOD_PYDOCK)
    return fail(LD_EINVAL, "method not supported");  // src/bin/lightdock-rust.rs:111-114
  if (desc->method == LD_METHOD_DFIRE && !desc->dfire_potential)
    return fail(LD_EINVAL, "DFIRE needs the DCparams table (src/dfire.rs:236-257)");
  int rc;
  if ((rc = check_molecule(desc->receptor, desc->method, desc->use_anm, "receptor")) != LD_OK) return rc;
  if ((rc = check_molecule(desc->ligand, desc->method, desc->use_anm, "ligand")) != LD_OK) return rc;
  ld_handle *h = new (std::nothrow) ld_handle();
  if (!h) return fail(LD_ENOMEM, "out of host memory");
  rc = create_impl(desc, h);
  if (rc != LD_OK) {
    std::string keep = g_err;
    ld_destroy(h);
    g_err = keep;
    return rc;
  }
  *out = h;
  return LD_OK;
}

extern "C" int ld_pose_len(const ld_handle *h) { return h ? h->cx.pose_len : LD_EINVAL; }

extern "C" int ld_set_rec_splits(ld_handle *h, int32_t splits) {
  if (!h || splits < 0) return fail(LD_EINVAL, "ld_set_rec_splits: bad argument");
  h->forced_splits = splits;
  return LD_OK;
}

extern "C" int ld_set_path(ld_handle *h, int32_t path) {
  if (!h || path < LD_PATH_AUTO || path > LD_PATH_RIGID) return fail(LD_EINVAL, "ld_set_path: bad argument");
  if (path == LD_PATH_RIGID && !h->rigid_ok) return fail(LD_EINVAL, "ld_set_path: " + h->rigid_info);
  h->path_mode = path;
  return LD_OK;
}

extern "C" const char *ld_path_info(const ld_handle *h) { return h ? h->rigid_info.c_str() : ""; }

extern "C" int ld_set_profiling(ld_handle *h, int32_t on) {
  if (!h) return fail(LD_EINVAL, "ld_set_profiling: NULL handle");
  h->profiling = on != 0;
  return LD_OK;
}

static int get_stats(ld_handle *h, Workspace *w, ld_batch_stats *out) {
  if (h->profiling && w->prof_used >= 4) {
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(w->last_stream));
    double t[3] = {0, 0, 0};
    for (size_t c = 0; c + 4 <= w->prof_used; c += 4)
      for (int k = 0; k < 3; ++k) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, w->prof_events[c + k], w->prof_events[c + k + 1]));
        t[k] += ms;
      }
    w->stats.transform_ms = t[0];
    w->stats.pair_ms = t[1];
    w->stats.finalize_ms = t[2];
  }
  *out = w->stats;
  return LD_OK;
}

extern "C" int ld_get_stats(ld_handle *h, ld_batch_stats *out) {
  if (!h || !out) return fail(LD_EINVAL, "ld_get_stats: NULL argument");
  return get_stats(h, h->last_w, out);
}

extern "C" int ld_get_stats_slot(ld_handle *h, int32_t slot, ld_batch_stats *out) {
  if (!h || !out || slot < 0 || slot >= LD_SLOTS) return fail(LD_EINVAL, "ld_get_stats_slot: bad argument");
  return get_stats(h, &h->ws[slot], out);
}

static int prof_mark(ld_handle *h, cudaStream_t st) {
  if (!h->profiling) return LD_OK;
  if (h->w->prof_used == h->w->prof_events.size()) {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    h->w->prof_events.push_back(e);
  }
  CU(cudaEventRecord(h->w->prof_events[h->w->prof_used++], st));
  return LD_OK;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static int regrow(T **p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  if (count == 0) return LD_OK;
  CU(cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T)));
  return LD_OK;
}

// Receptor tile ranges per pose (= CTAs per pose) for the generic kernels when a batch is too small to fill the
// GPU with one CTA per pose (one swarm = 200 poses).  A CTA's warps take tiles from a queue, so its time is
// ceil(tiles per CTA / warps) tile-times plus a fixed prologue (ligand staging, barriers: about half a tile-time);
// the launch takes ceil(CTAs / resident CTAs) such waves.  Pick the split count that minimises that product —
// e.g. 51 tiles over 3 CTAs of 16 warps is 17 tiles each, i.e. TWO rounds; 4 CTAs of 13 tiles is one.
static int choose_splits(const ld_handle *h, int64_t n) {
  const int tiles = std::max(1, h->cx.n_rec_tiles);
  if (h->forced_splits > 0) return std::min(h->forced_splits, tiles);
  const bool dna = h->cx.method != 0;
  const int warps = (dna ? DNA_THREADS : PAIR_THREADS) / 32;
  const int64_t resident = (int64_t)h->sm_count * (dna ? DNA_CTAS_PER_SM : 2);
  if (n >= 2 * resident) return 1;
  double best = 1e300;
  int best_s = 1;
  for (int s = 1; s <= tiles; ++s) {
    const int tps = (tiles + s - 1) / s;
    const double rounds = (double)((tps + warps - 1) / warps) + 0.5;
    const double waves = (double)((n * s + resident - 1) / resident);
    const double cost = rounds * waves;
    if (cost < best - 1e-9) {
      best = cost;
      best_s = s;
    }
  }
  return best_s;
}

static bool use_rigid(const ld_handle *h) { return h->rigid_ok && h->path_mode != LD_PATH_GENERIC; }

static int64_t chunk_limit(const ld_handle *h, bool rigid) {
  if (rigid && h->flex)  // per-pose f32 ligand blocks, read once per receptor group: keep a chunk's worth L2-resident
    return std::max<int64_t>(256, std::min<int64_t>((int64_t)1 << 20, ((int64_t)64 << 20) / ((int64_t)h->cx.n_lig_pad * 16)));
  if (rigid) return (int64_t)1 << 20;  // no per-pose coordinate blocks on this path (~2 KB per pose)
  if (h->dna_fused) return (int64_t)1 << 18;  // DNA/pyDock, pose transform inside the pair kernel: no coordinate blocks either
  const size_t per_pose = h->lig_block + h->rec_block + 64;
  int64_t c = (int64_t)((size_t)1 << 30) / (int64_t)per_pose;  // <= 1 GiB of coordinate blocks in flight
  return std::max<int64_t>(1, std::min<int64_t>(c, 16384));
}

static int ensure_chunk(ld_handle *h, int64_t chunk, int splits, bool need_blocks) {
  int rc;
  if (need_blocks && chunk > h->w->cap_blocks) {
    if ((rc = regrow(&h->w->d_lig_blocks, (size_t)chunk * h->lig_block)) != LD_OK) return rc;
    if ((rc = regrow(&h->w->d_rec_blocks, (size_t)chunk * h->rec_block)) != LD_OK) return rc;
    h->w->cap_blocks = chunk;
  }
  if (chunk <= h->w->cap_chunk && splits <= h->w->cap_splits) return LD_OK;
  chunk = std::max(chunk, h->w->cap_chunk);
  splits = std::max(splits, h->w->cap_splits);
  const int lig_words = (h->cx.n_lig_pad + 31) / 32;
  const int tiles = std::max(1, std::max(h->cx.n_rec_tiles, h->rigid_ok ? h->rc.n_groups : 0));
  if ((rc = regrow(&h->w->d_partials, (size_t)chunk * tiles * 2)) != LD_OK) return rc;
  if ((rc = regrow(&h->w->d_iface_rec, (size_t)chunk * tiles)) != LD_OK) return rc;
  if ((rc = regrow(&h->w->d_iface_lig, (size_t)chunk * splits * std::max(1, lig_words))) != LD_OK) return rc;
  h->w->cap_chunk = chunk;
  h->w->cap_splits = splits;
  return LD_OK;
}

static int ensure_poses(ld_handle *h, int64_t n, bool detail) {
  if (n > h->w->cap_poses) {
    int rc;
    if ((rc = regrow(&h->w->d_poses, (size_t)n * h->cx.pose_len)) != LD_OK) return rc;
    if ((rc = regrow(&h->w->d_energies, (size_t)n)) != LD_OK) return rc;
    if (h->w->d_detail) { cudaFree(h->w->d_detail); h->w->d_detail = nullptr; }
    h->w->cap_poses = n;
  }
  if (detail && !h->w->d_detail) {
    int rc;
    if ((rc = regrow(&h->w->d_detail, (size_t)h->w->cap_poses)) != LD_OK) return rc;
  }
  if (n > h->w->cap_pinned) {
    cudaFreeHost(h->w->h_poses); cudaFreeHost(h->w->h_energies);
    h->w->h_poses = h->w->h_energies = nullptr;
    CU(cudaHostAlloc(reinterpret_cast<void **>(&h->w->h_poses), (size_t)n * h->cx.pose_len * sizeof(double),
                     cudaHostAllocDefault));
    CU(cudaHostAlloc(reinterpret_cast<void **>(&h->w->h_energies), (size_t)n * sizeof(double), cudaHostAllocDefault));
    h->w->cap_pinned = n;
  }
  return LD_OK;
}

// Launches transform -> pair -> finalize for poses [0, n) living on the device.
// If host_iface_* are given (detail mode) the bitmaps of each chunk are copied back as they are produced.
// d_n_live (device-resident callers): the number of real rows, known only on the device; the launches are sized for
// n and every kernel skips the rows beyond *d_n_live.
static int run_device(ld_handle *h, int64_t n, const double *d_poses, double *d_energies, cudaStream_t st,
                      ld_pose_detail *d_detail, std::vector<unsigned> *host_ifr, std::vector<unsigned> *host_ifl,
                      const int *d_n_live = nullptr) {
  Nvtx range_launch("ld: kernel launches (prep/transform, pair, finalize)");
  const DeviceComplex &cx = h->cx;
  h->w->stats = ld_batch_stats{};
  h->w->stats.n_poses = n;
  h->w->stats.pair_evals_bruteforce = n * (int64_t)cx.n_rec * (int64_t)cx.n_lig;
  if (!h->prof_continue) h->w->prof_used = 0;
  // the previous user of this workspace's buffers may have launched on another stream
  if (h->w->done_recorded && h->w->last_stream != st) CU(cudaStreamWaitEvent(st, h->w->done, 0));
  h->w->last_stream = st;
  h->last_w = h->w;
  if (n == 0) return LD_OK;
  const int64_t climit = chunk_limit(h, use_rigid(h));
  const int lig_words = (cx.n_lig_pad + 31) / 32;
  const bool detail = d_detail != nullptr;
  int launches = 0, pair_launches = 0;
  for (int64_t p0 = 0; p0 < n; p0 += climit) {
    const int64_t nc = std::min(climit, n - p0);
    const int splits = choose_splits(h, nc);
    int rc;
    const bool rigid = use_rigid(h);
    // DNA/pyDock: the pair kernel transforms its pose itself (pair_prologue, dna_pair_kernel): no transform kernel, no
    // per-pose coordinate blocks
    const bool dna_fused = !rigid && h->dna_fused;
    if ((rc = ensure_chunk(h, nc, rigid ? 1 : splits, !rigid && !dna_fused)) != LD_OK) return rc;
    if (rigid) {
      const RigidComplex &rg = h->rc;
      BatchBuffers bb{};
      bb.poses = d_poses + (size_t)p0 * cx.pose_len;
      bb.partials = h->w->d_partials;
      bb.iface_rec = h->w->d_iface_rec;
      bb.iface_lig = h->w->d_iface_lig;
      bb.energies = d_energies + p0;
      bb.detail = detail ? (void *)(d_detail + p0) : nullptr;
      bb.rec_splits = 1;
      bb.tiles_per_split = rg.n_groups;
      bb.lig_words = lig_words;
      bb.n_live = d_n_live;
      bb.n_live_off = (int)p0;
      h->w->stats.rec_splits = 1;
      h->w->stats.path = LD_PATH_RIGID;
      if (!h->w->d_unit_counter) CU(cudaMalloc(reinterpret_cast<void **>(&h->w->d_unit_counter), sizeof(unsigned)));
      if (nc > h->w->cap_prep) {
        if ((rc = regrow(&h->w->d_prep, (size_t)nc * RG_PREP)) != LD_OK) return rc;
        h->w->cap_prep = nc;
      }
      if (h->flex && nc > h->w->cap_flex) {
        if ((rc = regrow(&h->w->d_lig4p, (size_t)nc * cx.n_lig_pad)) != LD_OK) return rc;
        if ((rc = regrow(&h->w->d_flag, (size_t)nc)) != LD_OK) return rc;
        h->w->cap_flex = nc;
      }
      if ((rc = prof_mark(h, st)) != LD_OK) return rc;
      // per-pose rotation data (the only "transform" on this path for a rigid ligand: nothing is moved per atom);
      // FLEX adds the per-pose f32 ligand block and the slack test
      if (h->flex)
        flex_prep_kernel<<<(unsigned)((nc + FLEX_PP - 1) / FLEX_PP), FLEX_THREADS,
                           flex_prep_smem(cx.n_lig_modes, cx.n_lig_tiles), st>>>(rg, bb, bb.poses, (int)nc, h->w->d_prep,
                                                                                 h->w->d_lig4p, h->w->d_flag, h->d_need);
      else
        rigid_prep_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(bb, bb.poses, (int)nc, cx.pose_len, h->w->d_prep);
      ++launches;
      if ((rc = prof_mark(h, st)) != LD_OK) return rc;
      CU(cudaMemsetAsync(h->w->d_iface_lig, 0, (size_t)nc * lig_words * sizeof(unsigned), st));
      CU(cudaMemsetAsync(h->w->d_unit_counter, 0, sizeof(unsigned), st));
      // work units: (group, range of poses); ~16 units per SM and group changes kept rare
      const int units_per_sm = g_opt.units_per_sm;
      const int cta_warps = h->flex ? h->flex_warps : RG_WARPS;
      int64_t ppu = (nc * rg.n_groups + (int64_t)h->sm_count * units_per_sm - 1) / ((int64_t)h->sm_count * units_per_sm);
      ppu = std::max<int64_t>(cta_warps, std::min<int64_t>(ppu, 1024));
      const int n_chunks = (int)((nc + ppu - 1) / ppu);
      const int64_t n_units = (int64_t)n_chunks * rg.n_groups;
      const unsigned grid = (unsigned)std::min<int64_t>(h->sm_count, n_units);
      const size_t smem = rigid_smem_bytes(rg.n_lig_pad, rg.rows_max, rg.row_bytes, h->flex ? cta_warps : 1);
      const unsigned threads = (unsigned)cta_warps * 32u;
      if (h->flex) {
        if (detail)
          dfire_rigid_kernel<true, true><<<grid, threads, smem, st>>>(rg, bb, (int)nc, (int)ppu, n_chunks, h->w->d_unit_counter,
                                                                      h->d_rc, h->w->d_prep, h->w->d_lig4p, h->w->d_flag);
        else
          dfire_rigid_kernel<false, true><<<grid, threads, smem, st>>>(rg, bb, (int)nc, (int)ppu, n_chunks, h->w->d_unit_counter,
                                                                       h->d_rc, h->w->d_prep, h->w->d_lig4p, h->w->d_flag);
      } else if (detail) {
        dfire_rigid_kernel<true, false><<<grid, threads, smem, st>>>(rg, bb, (int)nc, (int)ppu, n_chunks, h->w->d_unit_counter,
                                                                     h->d_rc, h->w->d_prep, nullptr, nullptr);
      } else {
        dfire_rigid_kernel<false, false><<<grid, threads, smem, st>>>(rg, bb, (int)nc, (int)ppu, n_chunks, h->w->d_unit_counter,
                                                                      h->d_rc, h->w->d_prep, nullptr, nullptr);
      }
      ++launches;
      ++pair_launches;
      if ((rc = prof_mark(h, st)) != LD_OK) return rc;
      const unsigned fgrid = (unsigned)((nc + 3) / 4);
      if (detail) finalize_kernel<true><<<fgrid, 128, 0, st>>>(h->cxr, bb, (int)nc);
      else finalize_kernel<false><<<fgrid, 128, 0, st>>>(h->cxr, bb, (int)nc);
      ++launches;
      if ((rc = prof_mark(h, st)) != LD_OK) return rc;
      CU(cudaGetLastError());
      if (host_ifr) {
        CU(cudaMemcpyAsync(host_ifr->data() + (size_t)p0 * rg.n_groups, h->w->d_iface_rec,
                           (size_t)nc * rg.n_groups * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(host_ifl->data() + (size_t)p0 * lig_words, h->w->d_iface_lig,
                           (size_t)nc * lig_words * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
      }
      continue;
    }
    h->w->stats.path = LD_PATH_GENERIC;
    BatchBuffers bb{};
    bb.poses = d_poses + (size_t)p0 * cx.pose_len;
    bb.lig_blocks = dna_fused ? nullptr : h->w->d_lig_blocks;
    bb.rec_blocks = dna_fused ? nullptr : h->w->d_rec_blocks;
    bb.partials = h->w->d_partials;
    bb.iface_rec = h->w->d_iface_rec;
    bb.iface_lig = h->w->d_iface_lig;
    bb.energies = d_energies + p0;
    bb.detail = detail ? (void *)(d_detail + p0) : nullptr;
    bb.rec_splits = splits;
    bb.tiles_per_split = (std::max(1, cx.n_rec_tiles) + splits - 1) / splits;
    bb.lig_words = lig_words;
    bb.n_live = d_n_live;
    bb.n_live_off = (int)p0;
    h->w->stats.rec_splits = splits;
    if ((rc = prof_mark(h, st)) != LD_OK) return rc;
    if (!dna_fused) {
      transform_kernel<<<(unsigned)((nc + TRANSFORM_PP - 1) / TRANSFORM_PP), 256,
                         (size_t)TRANSFORM_PP * (cx.n_rec_modes + cx.n_lig_modes) * sizeof(double), st>>>(cx, bb, (int)nc);
      ++launches;
    }
    if ((rc = prof_mark(h, st)) != LD_OK) return rc;
    if (cx.n_rec_tiles > 0) {
      const size_t smem = pair_smem_bytes(cx.method, cx.n_lig_pad, cx.n_lig_tiles, lig_words, bb.tiles_per_split, cx.vdw_nr * cx.vdw_nl);
      const unsigned grid = (unsigned)(nc * splits);
      if (cx.method == 0) {
        if (detail) dfire_pair_kernel<true><<<grid, PAIR_THREADS, smem, st>>>(cx, bb, (int)nc);
        else dfire_pair_kernel<false><<<grid, PAIR_THREADS, smem, st>>>(cx, bb, (int)nc);
      } else {
        const bool tab = cx.vdw_tab != nullptr;
        if (detail && tab) dna_pair_kernel<true, true><<<grid, DNA_THREADS, smem, st>>>(cx, bb, (int)nc);
        else if (detail) dna_pair_kernel<true, false><<<grid, DNA_THREADS, smem, st>>>(cx, bb, (int)nc);
        else if (tab) dna_pair_kernel<false, true><<<grid, DNA_THREADS, smem, st>>>(cx, bb, (int)nc);
        else dna_pair_kernel<false, false><<<grid, DNA_THREADS, smem, st>>>(cx, bb, (int)nc);
      }
      ++launches;
      ++pair_launches;
    } else {
      
      CU(cudaMemsetAsync(h->w->d_iface_lig, 0, (size_t)nc * splits * std::max(1, lig_words) * sizeof(unsigned), st));
    }
    if ((rc = prof_mark(h, st)) != LD_OK) return rc;
    const unsigned fgrid = (unsigned)((nc + 3) / 4);
    if (detail) finalize_kernel<true><<<fgrid, 128, 0, st>>>(cx, bb, (int)nc);
    else finalize_kernel<false><<<fgrid, 128, 0, st>>>(cx, bb, (int)nc);
    ++launches;
    if ((rc = prof_mark(h, st)) != LD_OK) return rc;
    CU(cudaGetLastError());
    if (host_ifr) {
      // detail mode: fetch this chunk's bitmaps (OR over splits for the ligand) before they are overwritten
      std::vector<unsigned> lig_tmp((size_t)nc * splits * lig_words);
      CU(cudaMemcpyAsync(host_ifr->data() + (size_t)p0 * cx.n_rec_tiles, h->w->d_iface_rec,
                         (size_t)nc * cx.n_rec_tiles * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(lig_tmp.data(), h->w->d_iface_lig, lig_tmp.size() * sizeof(unsigned), cudaMemcpyDeviceToHost,
                         st));
      CU(cudaStreamSynchronize(st));
      for (int64_t p = 0; p < nc; ++p)
        for (int c = 0; c < splits; ++c)
          for (int w = 0; w < lig_words; ++w)
            (*host_ifl)[(size_t)(p0 + p) * lig_words + w] |= lig_tmp[((size_t)p * splits + c) * lig_words + w];
    }
  }
  h->w->stats.kernel_launches = launches;
  h->w->stats.pair_launches = pair_launches;
  CU(cudaEventRecord(h->w->done, st));
  h->w->done_recorded = true;
  return LD_OK;
}

extern "C" int ld_score_batch_device(ld_handle *h, int64_t n_poses, const double *d_poses, double *d_energies,
                                     void *stream) {
  if (!h || n_poses < 0 || (n_poses > 0 && (!d_poses || !d_energies)))
    return fail(LD_EINVAL, "ld_score_batch_device: bad argument");
  h->w = &h->ws[0];
  if (h->w->pending >= 0) return fail(LD_EINVAL, "ld_score_batch_device: slot 0 has a batch pending");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->w->stream;
  return run_device(h, n_poses, d_poses, d_energies, st, nullptr, nullptr, nullptr);
}

static int score_host(ld_handle *h, int64_t n, const double *poses, double *energies, ld_pose_detail *detail,
                      uint8_t *iface_rec, uint8_t *iface_lig) {
  if (!h || n < 0 || (n > 0 && (!poses || !energies))) return fail(LD_EINVAL, "ld_score_batch: bad argument");
  h->w = &h->ws[0];
  if (h->w->pending >= 0) return fail(LD_EINVAL, "ld_score_batch: slot 0 has a batch pending (ld_score_batch_end it first)");
  CU(cudaSetDevice(h->device));
  if (n == 0) {
    h->w->stats = ld_batch_stats{};
    return LD_OK;
  }
  const DeviceComplex &cx = h->cx;
  const bool want_detail = detail != nullptr;
  int rc;
  if ((rc = ensure_poses(h, n, want_detail)) != LD_OK) return rc;
  const size_t pose_bytes = (size_t)n * cx.pose_len * sizeof(double);
  // Large plain batches go in two parts so that staging the bulk of the pose rows into pinned memory overlaps the
  // kernels of a small first part instead of preceding all device work (the two parts are ordinary sub-batches:
  // results do not depend on the split).
  const int64_t n_head = (!want_detail && n >= 65536) ? n / 8 : 0;  // smaller batches: a second launch costs more
  ld_batch_stats head_stats{};
  CU(cudaEventRecord(h->w->ev0, h->w->stream));
  Nvtx range_call("ld_score_batch");
  if (n_head > 0) {
    const size_t head_bytes = (size_t)n_head * cx.pose_len * sizeof(double);
    std::memcpy(h->w->h_poses, poses, head_bytes);
    CU(cudaMemcpyAsync(h->w->d_poses, h->w->h_poses, head_bytes, cudaMemcpyHostToDevice, h->w->stream));
    if ((rc = run_device(h, n_head, h->w->d_poses, h->w->d_energies, h->w->stream, nullptr, nullptr, nullptr)) != LD_OK)
      return rc;
    head_stats = h->w->stats;
    std::memcpy(reinterpret_cast<char *>(h->w->h_poses) + head_bytes, reinterpret_cast<const char *>(poses) + head_bytes,
                pose_bytes - head_bytes);
    CU(cudaMemcpyAsync(reinterpret_cast<char *>(h->w->d_poses) + head_bytes,
                       reinterpret_cast<char *>(h->w->h_poses) + head_bytes, pose_bytes - head_bytes,
                       cudaMemcpyHostToDevice, h->w->stream));
  } else {
    std::memcpy(h->w->h_poses, poses, pose_bytes);
    CU(cudaMemcpyAsync(h->w->d_poses, h->w->h_poses, pose_bytes, cudaMemcpyHostToDevice, h->w->stream));
  }
  if (want_detail) CU(cudaMemsetAsync(h->w->d_detail, 0, (size_t)n * sizeof(ld_pose_detail), h->w->stream));
  const int lig_words = (cx.n_lig_pad + 31) / 32;
  std::vector<unsigned> ifr, ifl;
  const bool want_iface = want_detail && (iface_rec || iface_lig);
  const bool rigid = use_rigid(h);
  const int rec_tiles = rigid ? h->rc.n_groups : cx.n_rec_tiles;
  if (want_iface) {
    ifr.assign((size_t)n * std::max(1, rec_tiles), 0u);
    ifl.assign((size_t)n * std::max(1, lig_words), 0u);
  }
  h->prof_continue = n_head > 0;
  rc = run_device(h, n - n_head, h->w->d_poses + (size_t)n_head * cx.pose_len, h->w->d_energies + n_head, h->w->stream,
                  want_detail ? h->w->d_detail : nullptr, want_iface ? &ifr : nullptr, want_iface ? &ifl : nullptr);
  h->prof_continue = false;
  if (rc != LD_OK) return rc;
  if (n_head > 0) {  // the call's counters cover both parts
    h->w->stats.n_poses += head_stats.n_poses;
    h->w->stats.pair_evals_bruteforce += head_stats.pair_evals_bruteforce;
    h->w->stats.kernel_launches += head_stats.kernel_launches;
    h->w->stats.pair_launches += head_stats.pair_launches;
  }
  CU(cudaMemcpyAsync(h->w->h_energies, h->w->d_energies, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->w->stream));
  if (want_detail)
    CU(cudaMemcpyAsync(detail, h->w->d_detail, (size_t)n * sizeof(ld_pose_detail), cudaMemcpyDeviceToHost, h->w->stream));
  CU(cudaEventRecord(h->w->ev1, h->w->stream));
  if (h->flex && use_rigid(h))
    CU(cudaMemcpyAsync(h->h_need, h->d_need, (size_t)cx.n_lig_tiles * sizeof(int), cudaMemcpyDeviceToHost, h->w->stream));
  {
    Nvtx range_wait("ld_score_batch: wait for the device + D2H");
    CU(cudaStreamSynchronize(h->w->stream));
  }
  if (h->flex && use_rigid(h) && (rc = grow_flex_slack(h)) != LD_OK) return rc;
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, h->w->ev0, h->w->ev1));
  h->w->stats.device_ms = ms;
  std::memcpy(energies, h->w->h_energies, (size_t)n * sizeof(double));
  if (want_iface) {
    for (int64_t p = 0; p < n; ++p) {
      if (iface_rec) {
        const std::vector<int> &perm = rigid ? h->rec_perm_r : h->rec_perm;
        const int npos = rigid ? h->rc.n_rec_pos : cx.n_rec;
        for (int i = 0; i < npos; ++i)
          if (perm[i] >= 0)
            iface_rec[(size_t)p * cx.n_rec + perm[i]] = (ifr[(size_t)p * rec_tiles + (i >> 5)] >> (i & 31)) & 1u;
      }
      if (iface_lig)
        for (int j = 0; j < cx.n_lig; ++j)
          iface_lig[(size_t)p * cx.n_lig + h->lig_perm[j]] = (ifl[(size_t)p * lig_words + (j >> 5)] >> (j & 31)) & 1u;
    }
  }
  return LD_OK;
}

extern "C" int ld_score_batch(ld_handle *h, int64_t n_poses, const double *poses, double *energies) {
  return score_host(h, n_poses, poses, energies, nullptr, nullptr, nullptr);
}

extern "C" int ld_score_batch_begin(ld_handle *h, int32_t slot, int64_t n, const double *poses) {
  if (!h || slot < 0 || slot >= LD_SLOTS || n < 0 || (n > 0 && !poses))
    return fail(LD_EINVAL, "ld_score_batch_begin: bad argument");
  Workspace *w = &h->ws[slot];
  if (w->pending >= 0) return fail(LD_EINVAL, "ld_score_batch_begin: the slot already has a batch pending");
  h->w = w;
  CU(cudaSetDevice(h->device));
  if (n > 0) {
    int rc;
    if ((rc = ensure_poses(h, n, false)) != LD_OK) return rc;
    const size_t pose_bytes = (size_t)n * h->cx.pose_len * sizeof(double);
    std::memcpy(w->h_poses, poses, pose_bytes);
    CU(cudaEventRecord(w->ev0, w->stream));
    CU(cudaMemcpyAsync(w->d_poses, w->h_poses, pose_bytes, cudaMemcpyHostToDevice, w->stream));
    if ((rc = run_device(h, n, w->d_poses, w->d_energies, w->stream, nullptr, nullptr, nullptr)) != LD_OK) return rc;
    CU(cudaMemcpyAsync(w->h_energies, w->d_energies, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaEventRecord(w->ev1, w->stream));
    if (h->flex && use_rigid(h))
      CU(cudaMemcpyAsync(h->h_need, h->d_need, (size_t)h->cx.n_lig_tiles * sizeof(int), cudaMemcpyDeviceToHost, w->stream));
  } else {
    w->stats = ld_batch_stats{};
  }
  w->pending = n;
  return LD_OK;
}

extern "C" int ld_score_batch_end(ld_handle *h, int32_t slot, double *energies) {
  if (!h || slot < 0 || slot >= LD_SLOTS) return fail(LD_EINVAL, "ld_score_batch_end: bad argument");
  Workspace *w = &h->ws[slot];
  if (w->pending < 0) return fail(LD_EINVAL, "ld_score_batch_end: no batch pending on the slot");
  const int64_t n = w->pending;
  w->pending = -1;
  if (n == 0) return LD_OK;
  if (!energies) return fail(LD_EINVAL, "ld_score_batch_end: energies is NULL");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(w->stream));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, w->ev0, w->ev1));
  w->stats.device_ms = ms;
  std::memcpy(energies, w->h_energies, (size_t)n * sizeof(double));
  // FLEX: the other slot may still be in flight; grow_flex_slack waits for it before it replaces the lists
  if (h->flex && use_rigid(h)) return grow_flex_slack(h);
  return LD_OK;
}

// ---------------------------------------------------------------------------------------------
// Device-resident GSO (ld_gso.cuh)
struct ld_gso {
  ld_handle *h = nullptr;
  GsoState st{};
  int cur = 0;                 // buffer of st.poses holding the current poses
  int step = 0;                // GSO steps run so far
  int cap_steps = 0;           // entries of st.n_packed
  int64_t energy_calls = 0;
  std::vector<void *> owned;
  std::vector<int> h_counts;   // n_packed of the steps of one ld_gso_run call
};

extern "C" int ld_gso_destroy(ld_gso *g) {
  if (!g) return LD_OK;
  cudaSetDevice(g->h->device);
  if (g->h->ws[0].stream) cudaStreamSynchronize(g->h->ws[0].stream);
  for (void *p : g->owned) cudaFree(p);
  g->h->gso_live = false;  // the slab stays with the handle for its next ld_gso
  delete g;
  return LD_OK;
}

static int gso_create_impl(ld_gso *g, const double *positions, const uint64_t *seeds) {
  GsoState &st = g->st;
  const size_t G = (size_t)st.n_swarms * st.n_glow, pl = (size_t)st.pose_len;
  cudaStream_t stream = g->h->ws[0].stream;
  double *p0 = nullptr, *p1 = nullptr, *lum = nullptr, *vis = nullptr, *sc = nullptr, *packed = nullptr, *en = nullptr;
  int *nn = nullptr, *slot = nullptr, *np = nullptr, *failed = nullptr;
  uint32_t *keys = nullptr;
  g->cap_steps = 1 << 16;
  // one allocation, carved (256-byte aligned pieces): creating and destroying a ld_gso costs one cudaMalloc / cudaFree
  size_t total = 0;
  auto reserve = [&](size_t bytes) { const size_t at = total; total += (std::max<size_t>(bytes, 1) + 255) / 256 * 256; return at; };
  const size_t o_p0 = reserve(G * pl * 8), o_p1 = reserve(G * pl * 8), o_packed = reserve(G * pl * 8), o_lum = reserve(G * 8),
               o_vis = reserve(G * 8), o_sc = reserve(G * 8), o_en = reserve(G * 8), o_nn = reserve(G * 4), o_slot = reserve(G * 4),
               o_np = reserve((size_t)g->cap_steps * 4), o_failed = reserve((size_t)st.n_swarms * 4),
               o_keys = reserve((size_t)st.n_swarms * 32);
  ld_handle *h = g->h;
  if (total > h->gso_slab_bytes) {
    cudaFree(h->gso_slab);
    h->gso_slab = nullptr;
    h->gso_slab_bytes = 0;
    CU(cudaMalloc(reinterpret_cast<void **>(&h->gso_slab), total));
    h->gso_slab_bytes = total;
  }
  unsigned char *slab = h->gso_slab;
  p0 = reinterpret_cast<double *>(slab + o_p0); p1 = reinterpret_cast<double *>(slab + o_p1);
  packed = reinterpret_cast<double *>(slab + o_packed); lum = reinterpret_cast<double *>(slab + o_lum);
  vis = reinterpret_cast<double *>(slab + o_vis); sc = reinterpret_cast<double *>(slab + o_sc);
  en = reinterpret_cast<double *>(slab + o_en); nn = reinterpret_cast<int *>(slab + o_nn);
  slot = reinterpret_cast<int *>(slab + o_slot); np = reinterpret_cast<int *>(slab + o_np);
  failed = reinterpret_cast<int *>(slab + o_failed); keys = reinterpret_cast<uint32_t *>(slab + o_keys);
  st.poses[0] = p0; st.poses[1] = p1; st.luciferin = lum; st.vision = vis; st.scoring = sc; st.packed = packed;
  st.energies = en; st.n_neighbors = nn; st.slot = slot; st.n_packed = np; st.failed = failed; st.keys = keys;
  // Glowworm::new (src/glowworm.rs:29-59): luciferin 5, vision range 0.2, scoring 0, no neighbours; step 0 = every
  // glowworm is scored by the first update_luciferin, so the first packed batch is every pose, in order
  std::vector<double> h_lum(G, 5.0), h_vis(G, 0.2);
  std::vector<int> h_slot(G);
  std::iota(h_slot.begin(), h_slot.end(), 0);
  // StdRng::seed_from_u64 (rand_core 0.5): the u64 expanded to the 256-bit ChaCha key by PCG32 output steps
  std::vector<uint32_t> h_keys((size_t)st.n_swarms * 8);
  for (int s = 0; s < st.n_swarms; ++s) {
    uint64_t state = seeds[s];
    for (int w = 0; w < 8; ++w) {
      state = state * 6364136223846793005ULL + 11634580027462260723ULL;
      const uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
      const uint32_t rot = (uint32_t)(state >> 59);
      h_keys[(size_t)s * 8 + w] = (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
  }
  const int first = (int)G;
  CU(cudaMemcpyAsync(p0, positions, G * pl * sizeof(double), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(packed, positions, G * pl * sizeof(double), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(lum, h_lum.data(), G * sizeof(double), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(vis, h_vis.data(), G * sizeof(double), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(slot, h_slot.data(), G * sizeof(int), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(keys, h_keys.data(), h_keys.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  CU(cudaMemsetAsync(sc, 0, G * sizeof(double), stream));
  CU(cudaMemsetAsync(nn, 0, G * sizeof(int), stream));
  CU(cudaMemsetAsync(failed, 0, (size_t)st.n_swarms * sizeof(int), stream));
  CU(cudaMemsetAsync(np, 0, (size_t)g->cap_steps * sizeof(int), stream));
  CU(cudaMemcpyAsync(np, &first, sizeof(int), cudaMemcpyHostToDevice, stream));
  CU(cudaStreamSynchronize(stream));  // the host vectors above go out of scope
  return LD_OK;
}

extern "C" int ld_gso_create(ld_handle *h, int32_t n_swarms, int32_t n_glowworms, const double *positions,
                             const uint64_t *seeds, ld_gso **out) {
  if (out) *out = nullptr;
  if (!h || !out || n_swarms <= 0 || n_glowworms <= 0 || !positions || !seeds)
    return fail(LD_EINVAL, "ld_gso_create: bad argument");
  if (n_glowworms > LD_GSO_MAX_GLOWWORMS)
    return fail(LD_ELIMIT, "ld_gso_create: more than LD_GSO_MAX_GLOWWORMS glowworms per swarm");
  if ((int64_t)n_swarms * n_glowworms > (int64_t)1 << 30) return fail(LD_ELIMIT, "ld_gso_create: too many glowworms");
  if (h->ws[0].pending >= 0) return fail(LD_EINVAL, "ld_gso_create: slot 0 has a batch pending");
  if (h->gso_live) return fail(LD_EINVAL, "ld_gso_create: the handle already has an ld_gso (one at a time)");
  CU(cudaSetDevice(h->device));
  ld_gso *g = new ld_gso();
  g->h = h;
  g->st.n_swarms = n_swarms;
  g->st.n_glow = n_glowworms;
  g->st.pose_len = h->cx.pose_len;
  g->st.n_rec_ext = h->cx.n_rec_modes;
  g->st.n_lig_ext = h->cx.n_lig_modes;
  h->gso_live = true;
  const int rc = gso_create_impl(g, positions, seeds);
  if (rc != LD_OK) {
    const std::string keep = g_err;
    ld_gso_destroy(g);
    g_err = keep;
    return rc;
  }
  *out = g;
  return LD_OK;
}

extern "C" int ld_gso_run(ld_gso *g, int32_t n_steps) {
  if (!g || n_steps < 0) return fail(LD_EINVAL, "ld_gso_run: bad argument");
  if (n_steps == 0) return LD_OK;
  ld_handle *h = g->h;
  if (g->step + n_steps >= g->cap_steps) return fail(LD_ELIMIT, "ld_gso_run: step counter capacity exceeded");
  h->w = &h->ws[0];
  if (h->w->pending >= 0) return fail(LD_EINVAL, "ld_gso_run: slot 0 has a batch pending");
  CU(cudaSetDevice(h->device));
  cudaStream_t stream = h->w->stream;
  const GsoState &st = g->st;
  const int64_t G = (int64_t)st.n_swarms * st.n_glow;
  const unsigned threads = (unsigned)((st.n_glow + 31) / 32 * 32);
  const size_t smem = (size_t)st.n_glow * 4 * sizeof(double);
  const int first = g->step;
  Nvtx range("ld_gso_run");
  for (int k = 0; k < n_steps; ++k) {
    const int step = g->step + 1;  // 1-based, src/lib.rs:47
    // update_luciferin's scoring pass: the rows packed by the previous step (all poses before step 1)
    int rc = run_device(h, G, st.packed, st.energies, stream, nullptr, nullptr, nullptr, st.n_packed + (step - 1));
    if (rc != LD_OK) return rc;
    gso_step_kernel<<<(unsigned)st.n_swarms, threads, smem, stream>>>(st, step, g->cur);
    CU(cudaGetLastError());
    g->cur ^= 1;
    g->step = step;
  }
  g->h_counts.resize((size_t)n_steps);  // pageable on purpose: no pinned allocation per ld_gso (they cost up to 100 ms now and then)
  CU(cudaMemcpyAsync(g->h_counts.data(), st.n_packed + first, (size_t)n_steps * sizeof(int), cudaMemcpyDeviceToHost, stream));
  const bool flex = h->flex && use_rigid(h);
  if (flex) CU(cudaMemcpyAsync(h->h_need, h->d_need, (size_t)h->cx.n_lig_tiles * sizeof(int), cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  for (int k = 0; k < n_steps; ++k) g->energy_calls += g->h_counts[k];
  if (flex) return grow_flex_slack(h);  // poses beyond the slacks were scored by brute force (exact); lists grow for the next call
  return LD_OK;
}

extern "C" int ld_gso_state(ld_gso *g, double *poses, double *luciferin, double *vision_range, double *scoring,
                            int32_t *n_neighbors, int32_t *failed_step) {
  if (!g) return fail(LD_EINVAL, "ld_gso_state: bad argument");
  CU(cudaSetDevice(g->h->device));
  cudaStream_t stream = g->h->ws[0].stream;
  const GsoState &st = g->st;
  const size_t G = (size_t)st.n_swarms * st.n_glow;
  if (poses) CU(cudaMemcpyAsync(poses, st.poses[g->cur], G * st.pose_len * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (luciferin) CU(cudaMemcpyAsync(luciferin, st.luciferin, G * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (vision_range) CU(cudaMemcpyAsync(vision_range, st.vision, G * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (scoring) CU(cudaMemcpyAsync(scoring, st.scoring, G * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (n_neighbors) CU(cudaMemcpyAsync(n_neighbors, st.n_neighbors, G * sizeof(int), cudaMemcpyDeviceToHost, stream));
  if (failed_step) CU(cudaMemcpyAsync(failed_step, st.failed, (size_t)st.n_swarms * sizeof(int), cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  return LD_OK;
}

extern "C" int32_t ld_gso_steps(const ld_gso *g) { return g ? g->step : LD_EINVAL; }
extern "C" int64_t ld_gso_energy_calls(const ld_gso *g) { return g ? g->energy_calls : LD_EINVAL; }

extern "C" int ld_score_batch_detail(ld_handle *h, int64_t n_poses, const double *poses, double *energies,
                                     ld_pose_detail *detail, uint8_t *iface_rec, uint8_t *iface_lig) {
  if (!detail) return fail(LD_EINVAL, "ld_score_batch_detail: detail is NULL");
  return score_host(h, n_poses, poses, energies, detail, iface_rec, iface_lig);
}

extern "C" int ld_transform_batch(ld_handle *h, int64_t n, const double *poses, double *rec_coords,
                                  double *lig_coords) {
  if (!h || n < 0 || (n > 0 && !poses)) return fail(LD_EINVAL, "ld_transform_batch: bad argument");
  h->w = &h->ws[0];
  if (h->w->pending >= 0) return fail(LD_EINVAL, "ld_transform_batch: slot 0 has a batch pending");
  CU(cudaSetDevice(h->device));
  if (n == 0) return LD_OK;
  const DeviceComplex &cx = h->cx;
  int rc;
  if ((rc = ensure_poses(h, n, false)) != LD_OK) return rc;
  if (h->w->done_recorded && h->w->last_stream != h->w->stream) CU(cudaStreamWaitEvent(h->w->stream, h->w->done, 0));
  CU(cudaMemcpyAsync(h->w->d_poses, poses, (size_t)n * cx.pose_len * sizeof(double), cudaMemcpyHostToDevice, h->w->stream));
  const int64_t climit = chunk_limit(h, false);
  std::vector<unsigned char> lb, rb;
  for (int64_t p0 = 0; p0 < n; p0 += climit) {
    const int64_t nc = std::min(climit, n - p0);
    if ((rc = ensure_chunk(h, nc, 1, true)) != LD_OK) return rc;
    BatchBuffers bb{};
    bb.poses = h->w->d_poses + (size_t)p0 * cx.pose_len;
    bb.lig_blocks = h->w->d_lig_blocks;
    bb.rec_blocks = h->w->d_rec_blocks;
    transform_kernel<<<(unsigned)((nc + TRANSFORM_PP - 1) / TRANSFORM_PP), 256,
                       (size_t)TRANSFORM_PP * (cx.n_rec_modes + cx.n_lig_modes) * sizeof(double), h->w->stream>>>(cx, bb, (int)nc);
    CU(cudaGetLastError());
    lb.resize((size_t)nc * h->lig_block);
    CU(cudaMemcpyAsync(lb.data(), h->w->d_lig_blocks, lb.size(), cudaMemcpyDeviceToHost, h->w->stream));
    if (h->rec_block) {
      rb.resize((size_t)nc * h->rec_block);
      CU(cudaMemcpyAsync(rb.data(), h->w->d_rec_blocks, rb.size(), cudaMemcpyDeviceToHost, h->w->stream));
    }
    CU(cudaStreamSynchronize(h->w->stream));
    for (int64_t p = 0; p < nc; ++p) {
      if (lig_coords) {
        const double *x = reinterpret_cast<const double *>(lb.data() + (size_t)p * h->lig_block);
        const double *y = x + cx.n_lig_pad, *z = y + cx.n_lig_pad;
        double *o = lig_coords + (size_t)(p0 + p) * cx.n_lig * 3;
        for (int j = 0; j < cx.n_lig; ++j) {
          const int a = h->lig_perm[j];
          o[3 * a] = x[j]; o[3 * a + 1] = y[j]; o[3 * a + 2] = z[j];
        }
      }
      if (rec_coords) {
        double *o = rec_coords + (size_t)(p0 + p) * cx.n_rec * 3;
        if (h->rec_block) {
          const double *x = reinterpret_cast<const double *>(rb.data() + (size_t)p * h->rec_block);
          const double *y = x + cx.n_rec_pad, *z = y + cx.n_rec_pad;
          for (int i = 0; i < cx.n_rec; ++i) {
            const int a = h->rec_perm[i];
            o[3 * a] = x[i]; o[3 * a + 1] = y[i]; o[3 * a + 2] = z[i];
          }
        } else {
          std::memcpy(o, h->rec_xyz_orig.data(), (size_t)cx.n_rec * 3 * sizeof(double));
        }
      }
    }
  }
  return LD_OK;
}

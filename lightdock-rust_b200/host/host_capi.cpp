// host_capi.cpp — small C entry points over the C++ host layer so the Python tests / bench can drive
// exactly the code the CLI runs (model building, batched scoring, GSO).  Not part of the drop-in ABI.
#include <cstring>
#include <stdexcept>
#include <string>

#include "gso.hpp"
#include "sharding.hpp"
#include "simulate.hpp"

using namespace lightdock;

static thread_local std::string g_err;
#define LDH_TRY try {
#define LDH_CATCH(ret)             \
  }                                \
  catch (const std::exception &e) { \
    g_err = e.what();              \
    return ret;                    \
  }

struct ldh_case {
  LoadedCase lc;
  const CudaScore *cuda() const { return static_cast<const CudaScore *>(lc.scoring.get()); }
};

extern "C" {

const char *ldh_last_error(void) { return g_err.c_str(); }

static bool parse_method(const char *m, Method *out) {
  const std::string s(m);
  if (s == "dfire") *out = Method::DFIRE;
  else if (s == "dna") *out = Method::DNA;
  else if (s == "pydock") *out = Method::PYDOCK;
  else return false;
  return true;
}

ldh_case *ldh_open_case(const char *setup_json, const char *method, const char *anm_dir, int device) {
  LDH_TRY
  Method m;
  if (!parse_method(method, &m)) { g_err = "Error: method not supported"; return nullptr; }
  const std::string path(setup_json);
  const SetupFile setup = read_setup_from_file(path);
  const size_t slash = path.find_last_of('/');
  auto *c = new ldh_case();
  try {
    c->lc = load_case(slash == std::string::npos ? "" : path.substr(0, slash), setup, m, anm_dir ? anm_dir : "",
                      device, false);
  } catch (...) {
    delete c;
    throw;
  }
  return c;
  LDH_CATCH(nullptr)
}

void ldh_close_case(ldh_case *c) { delete c; }

ld_handle *ldh_case_handle(ldh_case *c) { return c->cuda()->handle(); }

int ldh_case_info(ldh_case *c, int *n_rec, int *n_lig, int *pose_len, int *use_anm, unsigned long long *seed) {
  *n_rec = (int)c->cuda()->receptor().num_atoms();
  *n_lig = (int)c->cuda()->ligand().num_atoms();
  *pose_len = (int)c->lc.scoring->pose_len();
  *use_anm = c->lc.setup.use_anm ? 1 : 0;
  *seed = c->lc.seed;
  return 0;
}

// which: 0 receptor, 1 ligand.  Any output pointer may be NULL.  Arrays sized n_atoms (coords 3n).
int ldh_case_model(ldh_case *c, int which, int *dfire_types, double *coords, double *ele, double *vdw_e,
                   double *vdw_r) {
  const DockingModel &m = which == 0 ? c->cuda()->receptor() : c->cuda()->ligand();
  if (dfire_types && !m.atoms.empty()) std::memcpy(dfire_types, m.atoms.data(), m.atoms.size() * sizeof(int));
  if (coords) std::memcpy(coords, m.coordinates.data(), m.coordinates.size() * sizeof(double));
  if (ele && !m.ele_charges.empty()) std::memcpy(ele, m.ele_charges.data(), m.ele_charges.size() * sizeof(double));
  if (vdw_e && !m.vdw_charges.empty()) std::memcpy(vdw_e, m.vdw_charges.data(), m.vdw_charges.size() * sizeof(double));
  if (vdw_r && !m.vdw_radii.empty()) std::memcpy(vdw_r, m.vdw_radii.data(), m.vdw_radii.size() * sizeof(double));
  return 0;
}

// Sizes first (outputs NULL), then the data.  Restraint groups are the ACTIVE ones, ordered by residue id.
int ldh_case_restraints(ldh_case *c, int which, int *n_groups, int *n_idx, int *n_membrane, int *offsets, int *idx,
                        int *membrane) {
  const DockingModel &m = which == 0 ? c->cuda()->receptor() : c->cuda()->ligand();
  int total = 0;
  for (const auto &kv : m.active_restraints) total += (int)kv.second.size();
  if (n_groups) *n_groups = (int)m.active_restraints.size();
  if (n_idx) *n_idx = total;
  if (n_membrane) *n_membrane = (int)m.membrane.size();
  if (offsets && idx) {
    int k = 0, g = 0;
    offsets[0] = 0;
    for (const auto &kv : m.active_restraints) {
      for (int a : kv.second) idx[k++] = a;
      offsets[++g] = k;
    }
  }
  if (membrane && !m.membrane.empty()) std::memcpy(membrane, m.membrane.data(), m.membrane.size() * sizeof(int));
  return 0;
}

int ldh_case_energy_batch(ldh_case *c, long long n, const double *poses, double *energies) {
  LDH_TRY
  c->lc.scoring->energy_batch((size_t)n, poses, energies);
  return 0;
  LDH_CATCH(-1)
}

// Score::energy, single pose through the trait-shaped entry point.
int ldh_case_energy(ldh_case *c, const double *translation, const double *quat, const double *rec_nm, int n_rec_nm,
                    const double *lig_nm, int n_lig_nm, double *out) {
  LDH_TRY
  *out = c->lc.scoring->energy(std::vector<double>(translation, translation + 3),
                               Quaternion(quat[0], quat[1], quat[2], quat[3]),
                               std::vector<double>(rec_nm, rec_nm + n_rec_nm),
                               std::vector<double>(lig_nm, lig_nm + n_lig_nm));
  return 0;
  LDH_CATCH(-1)
}

// One swarm, exactly what the CLI does after loading: GSO::new + run.  out_dir may be NULL/"" (no files).
// final_state (optional): per glowworm [luciferin, scoring, n_neighbors, vision_range] then the pose row.
int ldh_case_gso(ldh_case *c, const char *positions_file, unsigned steps, const char *out_dir, double *final_state,
                 unsigned long long *energy_calls) {
  LDH_TRY
  const auto positions = parse_input_coordinates(positions_file);
  const SetupFile &s = c->lc.setup;
  GSO gso(positions, c->lc.seed, c->lc.scoring.get(), s.use_anm, s.anm_rec, s.anm_lig, out_dir ? out_dir : "");
  gso.run(steps);
  if (energy_calls) *energy_calls = gso.swarm.energy_calls;
  if (final_state) {
    const size_t pl = c->lc.scoring->pose_len();
    for (size_t i = 0; i < gso.swarm.glowworms.size(); ++i) {
      const Glowworm &g = gso.swarm.glowworms[i];
      double *r = final_state + i * (4 + pl);
      r[0] = g.luciferin; r[1] = g.scoring; r[2] = (double)g.neighbors.size(); r[3] = g.vision_range;
      g.write_pose(r + 4);
    }
  }
  return 0;
  LDH_CATCH(-1)
}

// Many swarms in lock-step, one batched launch per step.  positions: [n_swarms][n_glowworms][pose_len].
// seeds: one per swarm.  final_state as above, [n_swarms][n_glowworms][4+pose_len].  out_dirs may be NULL.
int ldh_case_multi_gso(ldh_case *c, int n_swarms, int n_glowworms, const double *positions,
                       const unsigned long long *seeds, unsigned steps, int host_threads, const char *const *out_dirs,
                       double *final_state, unsigned long long *energy_calls) {
  LDH_TRY
  const SetupFile &s = c->lc.setup;
  const size_t pl = c->lc.scoring->pose_len();
  MultiGSO multi(c->lc.scoring.get());
  for (int w = 0; w < n_swarms; ++w) {
    std::vector<std::vector<double>> pos(n_glowworms);
    for (int g = 0; g < n_glowworms; ++g) {
      const double *row = positions + ((size_t)w * n_glowworms + g) * pl;
      pos[g].assign(row, row + pl);
    }
    multi.add(pos, seeds[w], s.use_anm, s.anm_rec, s.anm_lig, out_dirs && out_dirs[w] ? out_dirs[w] : "");
  }
  multi.run(steps, host_threads);
  if (energy_calls) *energy_calls = multi.energy_calls();
  const auto failures = multi.failures();
  if (final_state)
    for (int w = 0; w < n_swarms; ++w)
      for (int i = 0; i < n_glowworms; ++i) {
        const Glowworm &g = multi.runs[w].swarm.glowworms[i];
        double *r = final_state + ((size_t)w * n_glowworms + i) * (4 + pl);
        r[0] = g.luciferin; r[1] = g.scoring; r[2] = (double)g.neighbors.size(); r[3] = g.vision_range;
        g.write_pose(r + 4);
      }
  if (!failures.empty()) {  // the other swarms ran to the end; their state is in final_state
    std::string msg = std::to_string(failures.size()) + " swarm(s) stopped early:";
    for (const auto &f : failures) msg += " [" + std::to_string(f.first) + "] " + f.second + ";";
    throw std::runtime_error(msg);
  }
  return 0;
  LDH_CATCH(-1)
}

// The same run with the whole GSO step on the device (DeviceGSO -> ld_gso_*).  Arguments as ldh_case_multi_gso.
int ldh_case_device_gso(ldh_case *c, int n_swarms, int n_glowworms, const double *positions,
                        const unsigned long long *seeds, unsigned steps, int host_threads, const char *const *out_dirs,
                        double *final_state, unsigned long long *energy_calls) {
  LDH_TRY
  const SetupFile &s = c->lc.setup;
  const size_t pl = c->lc.scoring->pose_len();
  DeviceGSO multi(c->lc.scoring.get());
  for (int w = 0; w < n_swarms; ++w) {
    std::vector<std::vector<double>> pos(n_glowworms);
    for (int g = 0; g < n_glowworms; ++g) {
      const double *row = positions + ((size_t)w * n_glowworms + g) * pl;
      pos[g].assign(row, row + pl);
    }
    multi.add(pos, seeds[w], s.use_anm, s.anm_rec, s.anm_lig, out_dirs && out_dirs[w] ? out_dirs[w] : "");
  }
  multi.run(steps, host_threads);
  if (energy_calls) *energy_calls = multi.energy_calls();
  const auto failures = multi.failures();
  if (final_state) {
    const DeviceGSO::State &st = multi.state;
    for (size_t k = 0; k < (size_t)n_swarms * n_glowworms; ++k) {
      double *r = final_state + k * (4 + pl);
      r[0] = st.luciferin[k]; r[1] = st.scoring[k]; r[2] = (double)st.n_neighbors[k]; r[3] = st.vision[k];
      std::copy(st.poses.begin() + k * pl, st.poses.begin() + (k + 1) * pl, r + 4);
    }
  }
  if (!failures.empty()) {
    std::string msg = std::to_string(failures.size()) + " swarm(s) stopped early:";
    for (const auto &f : failures) msg += " [" + std::to_string(f.first) + "] " + f.second + ";";
    throw std::runtime_error(msg);
  }
  return 0;
  LDH_CATCH(-1)
}

// Host-only pieces, testable without a GPU ------------------------------------------------------
// The neighbour search of Swarm::movement_phase on a given swarm state.  xyz [n][3]; out_offsets [n+1];
// out_idx must hold n*(n-1) entries.  Returns the total number of neighbours.
int ldh_find_neighbors(int n, const double *xyz, const double *luciferin, const double *vision_range, int *out_offsets,
                       int *out_idx) {
  LDH_TRY
  Swarm sw;
  for (int i = 0; i < n; ++i) {
    sw.glowworms.emplace_back((uint32_t)i, std::vector<double>{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]},
                              Quaternion(1.0, 0.0, 0.0, 0.0), std::vector<double>(), std::vector<double>(), nullptr, false);
    sw.glowworms.back().luciferin = luciferin[i];
    sw.glowworms.back().vision_range = vision_range[i];
  }
  sw.find_neighbors();
  int k = 0;
  for (int i = 0; i < n; ++i) {
    out_offsets[i] = k;
    for (uint32_t id : sw.glowworms[i].neighbors) out_idx[k++] = (int)id;
  }
  out_offsets[n] = k;
  return k;
  LDH_CATCH(-1)
}

// Cost-aware swarm -> GPU map (host/sharding.hpp).  centres [n_swarms][3]: mean translation of each swarm's glowworms.
// out_cost [n_swarms] (may be NULL), out_gpu [n_swarms].  Host-only: needs no device.
int ldh_shard_swarms(int n_rec, const double *rec_xyz, int n_lig, const double *lig_xyz, int n_swarms,
                     const double *centres, int n_gpus, double *out_cost, int *out_gpu) {
  LDH_TRY
  const SwarmCostModel model(std::vector<double>(rec_xyz, rec_xyz + (size_t)3 * n_rec),
                             std::vector<double>(lig_xyz, lig_xyz + (size_t)3 * n_lig));
  std::vector<double> cost(n_swarms);
  for (int s = 0; s < n_swarms; ++s) cost[s] = model.cost(centres + (size_t)3 * s);
  const std::vector<int> gpu = assign_swarms_lpt(cost, n_gpus);
  for (int s = 0; s < n_swarms; ++s) {
    if (out_cost) out_cost[s] = cost[s];
    out_gpu[s] = gpu[s];
  }
  return 0;
  LDH_CATCH(-1)
}

// Host-only test hooks for the text I/O either side of the path -----------------------------------
// 1 if `token.parse::<f64>()` succeeds in Rust (value in *out), 0 if it is a ParseFloatError there.
int ldh_parse_f64(const char *token, double *out) {
  double v = 0.0;
  const bool ok = parse_f64_like_rust(token, v);
  if (ok && out) *out = v;
  return ok ? 1 : 0;
}
// Writes a swarm state with Swarm::save (src/swarm.rs:128-167): rows [n][4 + pose_len] = luciferin, scoring,
// vision range, n_neighbors, pose...
int ldh_save_swarm(int n, int n_rec_anm, int n_lig_anm, const double *rows, unsigned step, const char *dir) {
  LDH_TRY
  const int pl = 7 + n_rec_anm + n_lig_anm;
  Swarm sw;
  std::vector<std::vector<double>> pos(n);
  for (int i = 0; i < n; ++i) pos[i].assign(rows + (size_t)i * (4 + pl) + 4, rows + (size_t)(i + 1) * (4 + pl));
  sw.add_glowworms(pos, nullptr, n_rec_anm + n_lig_anm > 0, n_rec_anm, n_lig_anm);
  for (int i = 0; i < n; ++i) {
    const double *r = rows + (size_t)i * (4 + pl);
    sw.glowworms[i].luciferin = r[0];
    sw.glowworms[i].scoring = r[1];
    sw.glowworms[i].vision_range = r[2];
    sw.glowworms[i].neighbors.assign((size_t)r[3], 0u);
  }
  sw.save(step, dir);
  return 0;
  LDH_CATCH(-1)
}

double ldh_rng_draws(unsigned long long seed, int n, double *out) {
  StdRng r = StdRng::seed_from_u64(seed);
  double last = 0;
  for (int i = 0; i < n; ++i) {
    last = r.gen_f64();
    if (out) out[i] = last;
  }
  return last;
}
void ldh_slerp(const double *a, const double *b, double t, double *out) {
  const Quaternion q = Quaternion(a[0], a[1], a[2], a[3]).slerp(Quaternion(b[0], b[1], b[2], b[3]), t);
  out[0] = q.w; out[1] = q.x; out[2] = q.y; out[3] = q.z;
}
void ldh_rotate(const double *q, const double *v, double *out) {
  Quaternion(q[0], q[1], q[2], q[3]).rotate(v, out);
}

// Builds the numeric model of one structure without touching the GPU (typing / restraint parity tests).
// method: "dfire" | "dna" | "pydock"; restraints: '\n'-separated residue ids.  Two-call pattern: with
// all outputs NULL returns n_atoms; negative on error.
int ldh_build_model(const char *pdb_path, const char *method, const char *active_restraints, int *dfire_types,
                    double *coords, double *ele, double *vdw_e, double *vdw_r, int *membrane_count, int *membrane,
                    int *rst_groups, int *rst_offsets, int *rst_idx) {
  LDH_TRY
  Method m;
  if (!parse_method(method, &m)) { g_err = "Error: method not supported"; return -1; }
  std::vector<std::string> active;
  std::string cur;
  for (const char *p = active_restraints ? active_restraints : ""; ; ++p) {
    if (*p == '\n' || *p == '\0') {
      if (!cur.empty()) active.push_back(cur);
      cur.clear();
      if (*p == '\0') break;
    } else cur += *p;
  }
  const PDB pdb = open_pdb(pdb_path);
  const DockingModel dm = DockingModel::build(m, pdb, active, {}, {}, 0);
  if (dfire_types && !dm.atoms.empty()) std::memcpy(dfire_types, dm.atoms.data(), dm.atoms.size() * sizeof(int));
  if (coords) std::memcpy(coords, dm.coordinates.data(), dm.coordinates.size() * sizeof(double));
  if (ele && !dm.ele_charges.empty()) std::memcpy(ele, dm.ele_charges.data(), dm.ele_charges.size() * sizeof(double));
  if (vdw_e && !dm.vdw_charges.empty()) std::memcpy(vdw_e, dm.vdw_charges.data(), dm.vdw_charges.size() * sizeof(double));
  if (vdw_r && !dm.vdw_radii.empty()) std::memcpy(vdw_r, dm.vdw_radii.data(), dm.vdw_radii.size() * sizeof(double));
  if (membrane_count) *membrane_count = (int)dm.membrane.size();
  if (membrane && !dm.membrane.empty()) std::memcpy(membrane, dm.membrane.data(), dm.membrane.size() * sizeof(int));
  if (rst_groups) *rst_groups = (int)dm.active_restraints.size();
  if (rst_offsets && rst_idx) {
    int k = 0, g = 0;
    rst_offsets[0] = 0;
    for (const auto &kv : dm.active_restraints) {
      for (int a : kv.second) rst_idx[k++] = a;
      rst_offsets[++g] = k;
    }
  }
  return (int)dm.num_atoms();
  LDH_CATCH(-1)
}

}  // extern "C"

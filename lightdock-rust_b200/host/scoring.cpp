#include "scoring.hpp"
#include "log.hpp"
#include "setup.hpp"

#include <cctype>
#include <cstring>
#include <iterator>
#include <thread>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <stdexcept>

#include "forcefield.hpp"

namespace lightdock {

const char *method_name(Method m) {
  switch (m) {
    case Method::DFIRE: return "DFIRE";
    case Method::DNA: return "DNA";
    default: return "PYDOCK";
  }
}

static bool contains(const std::vector<std::string> &v, const std::string &s) {
  return std::find(v.begin(), v.end(), s) != v.end();
}

DockingModel DockingModel::build(Method method, const PDB &structure, const std::vector<std::string> &active,
                                 const std::vector<std::string> &passive, const std::vector<double> &nmodes,
                                 size_t num_anm) {
  const ForceField &ff = ForceField::instance();
  DockingModel model;
  model.nmodes = nmodes;
  model.num_anm = num_anm;
  const char *tag = method == Method::DNA ? "DNA" : "PYDOCK";
  int atom_index = 0;
  for (const Atom &atom : structure.atoms) {
    const std::string res_id = residue_id(atom);
    // membrane beads MMB.BJ (src/dfire.rs:145-149)
    if (atom.res_name + atom.name == "MMBBJ") model.membrane.push_back(atom_index);
    if (contains(active, res_id)) model.active_restraints[res_id].push_back(atom_index);
    if (contains(passive, res_id)) model.passive_restraints[res_id].push_back(atom_index);

    if (method == Method::DFIRE) {
      // r3_to_numerical + ATOMNUMBER + ATOMRES (src/dfire.rs:18-46,56-101,177-183) folded into one table
      auto it = ff.dfire_type.find(atom.res_name + "\t" + atom.name);
      if (it == ff.dfire_type.end()) {
        bool known_residue = false;
        for (const auto &kv : ff.dfire_type)
          if (kv.first.compare(0, atom.res_name.size() + 1, atom.res_name + "\t") == 0) { known_residue = true; break; }
        if (!known_residue) throw std::runtime_error("Residue name not supported in DFIRE scoring function");
        throw std::runtime_error("Not supported atom type \"" + atom.res_name + atom.name + "\"");
      }
      model.atoms.push_back(it->second);
    } else {
      // src/dna.rs:314-356 ; src/pydock.rs:318-372
      const auto &amber = method == Method::PYDOCK ? ff.amber_type_pydock : ff.amber_type;
      const auto &ele = method == Method::PYDOCK ? ff.ele_charge_pydock : ff.ele_charge;
      std::string atom_id = atom.res_name + "-" + atom.name;
      auto at = amber.find(atom_id);
      if (at == amber.end()) {
        if (atom.name == "H1" || atom.name == "H2" || atom.name == "H3") {
          atom_id = atom.res_name + "-H";
          at = amber.find(atom_id);
          if (at == amber.end())
            throw std::runtime_error(std::string(tag) + " Error: Atom [\"" + atom_id + "\"] not supported");
        } else if (method == Method::PYDOCK) {
          // warn!(..) of src/pydock.rs:332-335: shown at RUST_LOG=warn and above, silent by default (env_logger)
          log_line(LogLevel::Warn, "lightdock::pydock", "PYDOCK Warning: Atom [\"" + atom_id + "\"] not supported, trying generic");
          if (atom.name.empty())
            throw std::runtime_error("PYDOCK Error: Atom element could not be guessed from [\"\"]");
          atom_id = std::string("*-") + atom.name[0];
          at = amber.find(atom_id);
          if (at == amber.end())
            throw std::runtime_error("PYDOCK Error: Atom [\"" + atom_id + "\"] not supported");
        } else {
          throw std::runtime_error("DNA Error: Atom [\"" + atom_id + "\"] not supported");
        }
      }
      auto q = ele.find(atom_id);
      if (q == ele.end()) {
        q = ff.nt_ele_charge.find(atom_id);
        if (q == ff.nt_ele_charge.end())
          throw std::runtime_error(std::string(tag) + " Error: Atom [\"" + atom_id +
                                   "\"] electrostatics charge not found");
      }
      model.ele_charges.push_back(q->second);
      auto e = ff.vdw_energy.find(at->second);
      if (e == ff.vdw_energy.end())
        throw std::runtime_error(std::string(tag) + " Error: Atom [\"" + atom_id + "\"] VDW charge not found");
      model.vdw_charges.push_back(e->second);
      auto r = ff.vdw_radius.find(at->second);
      if (r == ff.vdw_radius.end())
        throw std::runtime_error(std::string(tag) + " Error: Atom [\"" + atom_id + "\"] VDW radius not found");
      model.vdw_radii.push_back(r->second);
    }
    model.coordinates.push_back(atom.x);
    model.coordinates.push_back(atom.y);
    model.coordinates.push_back(atom.z);
    ++atom_index;
  }
  if (method == Method::PYDOCK)  // info!("Atoms read: {}", atom_index), src/pydock.rs:379
    log_line(LogLevel::Info, "lightdock::pydock", "Atoms read: " + std::to_string(atom_index));
  return model;
}

std::vector<double> load_potentials() {
  // src/dfire.rs:236-257: every line of $LIGHTDOCK_DATA/DCparams (default data/DCparams), the first 169*169*20 of
  // them parsed with `line.trim().parse::<f64>().unwrap()`.  13 MB of text: read at once and parsed on a few threads
  // (the single-threaded getline + strtod loop was 60 ms of a 0.5 s single-swarm run).
  const char *env = std::getenv("LIGHTDOCK_DATA");
  const std::string folder = env ? env : "data";
  const std::string path = folder + "/DCparams";
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("Unable to open DFIRE parameters: " + path);
  std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  std::vector<std::pair<size_t, size_t>> lines;  // [begin, end) of the first LD_DFIRE_TABLE_LEN lines
  lines.reserve(LD_DFIRE_TABLE_LEN);
  for (size_t st = 0; st < text.size() && (int)lines.size() < LD_DFIRE_TABLE_LEN;) {
    const char *nl = static_cast<const char *>(std::memchr(text.data() + st, '\n', text.size() - st));
    const size_t en = nl ? (size_t)(nl - text.data()) : text.size();
    lines.emplace_back(st, en);
    st = en + 1;
  }
  if ((int)lines.size() < LD_DFIRE_TABLE_LEN)
    throw std::runtime_error("DFIRE parameters: " + path + " has fewer than 169*169*20 lines");
  std::vector<double> potential(LD_DFIRE_TABLE_LEN);
  const int n_thr = std::max(1, std::min(8, (int)std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  std::vector<int> bad(n_thr, -1);
  for (int t = 0; t < n_thr; ++t)
    pool.emplace_back([&, t] {
      const size_t lo = lines.size() * t / n_thr, hi = lines.size() * (t + 1) / n_thr;
      std::string tok;
      for (size_t i = lo; i < hi; ++i) {
        size_t a = lines[i].first, b = lines[i].second;
        while (a < b && std::isspace((unsigned char)text[a])) ++a;       // trim()
        while (b > a && std::isspace((unsigned char)text[b - 1])) --b;
        tok.assign(text, a, b - a);
        if (!parse_f64_like_rust(tok, potential[i])) { bad[t] = (int)i; return; }
      }
    });
  for (auto &th : pool) th.join();
  for (int b : bad)
    if (b >= 0) throw std::runtime_error("Unable to read DFIRE parameters: bad line " + std::to_string(b + 1) + " in " + path);
  return potential;
}

namespace {
struct Csr {
  std::vector<int> off{0}, idx;
};
Csr to_csr(const std::map<std::string, std::vector<int>> &groups) {
  Csr c;
  for (const auto &kv : groups) {
    c.idx.insert(c.idx.end(), kv.second.begin(), kv.second.end());
    c.off.push_back((int)c.idx.size());
  }
  return c;
}
void fill_desc(ld_molecule_desc &d, const DockingModel &m, const Csr &rst, bool use_anm) {
  d.n_atoms = (int32_t)m.num_atoms();
  d.coords = m.coordinates.data();
  d.dfire_type = m.atoms.empty() ? nullptr : m.atoms.data();
  d.ele_charge = m.ele_charges.empty() ? nullptr : m.ele_charges.data();
  d.vdw_energy = m.vdw_charges.empty() ? nullptr : m.vdw_charges.data();
  d.vdw_radius = m.vdw_radii.empty() ? nullptr : m.vdw_radii.data();
  // energy() applies ANM only when use_anm && num_anm > 0 (src/dfire.rs:290,306)
  d.n_modes = use_anm ? (int32_t)m.num_anm : 0;
  d.modes = m.nmodes.empty() ? nullptr : m.nmodes.data();
  d.n_restraints = (int32_t)rst.off.size() - 1;
  d.rst_offsets = rst.off.data();
  d.rst_atoms = rst.idx.empty() ? nullptr : rst.idx.data();
  d.n_membrane = (int32_t)m.membrane.size();
  d.membrane = m.membrane.empty() ? nullptr : m.membrane.data();
}
}  // namespace

CudaScore::CudaScore(Method method, DockingModel receptor, DockingModel ligand, bool use_anm,
                     std::vector<double> potential, int device)
    : method_(method), receptor_(std::move(receptor)), ligand_(std::move(ligand)), use_anm_(use_anm),
      potential_(std::move(potential)) {
  for (const DockingModel *m : {&receptor_, &ligand_})
    if (use_anm_ && m->num_anm > 0 && m->nmodes.size() != m->num_atoms() * 3 * m->num_anm)
      throw std::runtime_error("ANM data does not correspond to the number of atoms");
  ld_complex_desc desc{};
  desc.method = method == Method::DFIRE ? LD_METHOD_DFIRE : (method == Method::DNA ? LD_METHOD_DNA : LD_METHOD_PYDOCK);
  desc.use_anm = use_anm_ ? 1 : 0;
  const Csr rr = to_csr(receptor_.active_restraints), lr = to_csr(ligand_.active_restraints);
  fill_desc(desc.receptor, receptor_, rr, use_anm_);
  fill_desc(desc.ligand, ligand_, lr, use_anm_);
  desc.dfire_potential = potential_.empty() ? nullptr : potential_.data();
  desc.device = device;
  if (ld_create(&desc, &handle_) != LD_OK)
    throw std::runtime_error(std::string("lightdock_b200: ") + ld_last_error());
  pose_len_ = (size_t)ld_pose_len(handle_);
}

CudaScore::~CudaScore() { ld_destroy(handle_); }

void CudaScore::energy_batch(size_t n, const double *poses, double *energies) const {
  if (ld_score_batch(handle_, (int64_t)n, poses, energies) != LD_OK)
    throw std::runtime_error(std::string("lightdock_b200: ") + ld_last_error());
}

void CudaScore::energy_batch_begin(int slot, size_t n, const double *poses) const {
  if (ld_score_batch_begin(handle_, slot, (int64_t)n, poses) != LD_OK)
    throw std::runtime_error(std::string("lightdock_b200: ") + ld_last_error());
}

void CudaScore::energy_batch_end(int slot, double *energies) const {
  if (ld_score_batch_end(handle_, slot, energies) != LD_OK)
    throw std::runtime_error(std::string("lightdock_b200: ") + ld_last_error());
}

double CudaScore::energy(const std::vector<double> &translation, const Quaternion &rotation,
                         const std::vector<double> &rec_nmodes, const std::vector<double> &lig_nmodes) const {
  std::vector<double> row(pose_len_, 0.0);
  row[0] = translation[0]; row[1] = translation[1]; row[2] = translation[2];
  row[3] = rotation.w; row[4] = rotation.x; row[5] = rotation.y; row[6] = rotation.z;
  size_t k = 7;
  const size_t nr = use_anm_ ? receptor_.num_anm : 0, nl = use_anm_ ? ligand_.num_anm : 0;
  for (size_t i = 0; i < nr; ++i) row[k++] = rec_nmodes.at(i);  // indexing past the slice panics in the reference
  for (size_t i = 0; i < nl; ++i) row[k++] = lig_nmodes.at(i);
  double e = 0.0;
  energy_batch(1, row.data(), &e);
  return e;
}

#define LD_DEFINE_CREATE(NAME, METHOD, NEEDS_TABLE)                                                              \
  std::unique_ptr<Score> NAME::create(                                                                           \
      const PDB &receptor, const std::vector<std::string> &rec_active, const std::vector<std::string> &rec_passive, \
      const std::vector<double> &rec_nmodes, size_t rec_num_anm, const PDB &ligand,                                \
      const std::vector<std::string> &lig_active, const std::vector<std::string> &lig_passive,                     \
      const std::vector<double> &lig_nmodes, size_t lig_num_anm, bool use_anm, int device) {                       \
    DockingModel r = DockingModel::build(METHOD, receptor, rec_active, rec_passive, rec_nmodes, rec_num_anm);     \
    DockingModel l = DockingModel::build(METHOD, ligand, lig_active, lig_passive, lig_nmodes, lig_num_anm);       \
    std::vector<double> pot;                                                                                     \
    if (NEEDS_TABLE) pot = load_potentials();                                                                    \
    return std::unique_ptr<Score>(new CudaScore(METHOD, std::move(r), std::move(l), use_anm, std::move(pot), device)); \
  }
LD_DEFINE_CREATE(DFIRE, Method::DFIRE, true)
LD_DEFINE_CREATE(DNA, Method::DNA, false)
LD_DEFINE_CREATE(PYDOCK, Method::PYDOCK, false)

}  // namespace lightdock

#include "simulate.hpp"

#include <cstdio>
#include <stdexcept>

#include "gso.hpp"

namespace lightdock {

std::string rust_debug_str(const std::string &s) {
  std::string o = "\"";
  for (char c : s) {
    if (c == '"' || c == '\\') { o += '\\'; o += c; }
    else if (c == '\n') o += "\\n";
    else if (c == '\t') o += "\\t";
    else if (c == '\r') o += "\\r";
    else o += c;
  }
  return o + "\"";
}

LoadedCase load_case(const std::string &simulation_path, const SetupFile &setup, Method method,
                     const std::string &anm_dir, int device, bool verbose) {
  LoadedCase lc;
  lc.setup = setup;
  lc.method = method;
  lc.seed = setup.seed ? *setup.seed : DEFAULT_SEED;
  const std::string prefix = "lightdock_";  // DEFAULT_LIGHTDOCK_PREFIX, src/constants.rs:18
  const std::string receptor_filename =
      simulation_path.empty() ? prefix + setup.receptor_pdb : simulation_path + "/" + prefix + setup.receptor_pdb;
  if (verbose) std::printf("Reading receptor input structure: %s\n", receptor_filename.c_str());
  const PDB receptor = open_pdb(receptor_filename);
  const std::string ligand_filename =
      simulation_path.empty() ? prefix + setup.ligand_pdb : simulation_path + "/" + prefix + setup.ligand_pdb;
  if (verbose) std::printf("Reading ligand input structure: %s\n", ligand_filename.c_str());
  const PDB ligand = open_pdb(ligand_filename);

  std::vector<double> rec_nm, lig_nm;
  if (setup.use_anm) {
    const std::string dir = anm_dir.empty() ? "" : anm_dir + "/";
    if (setup.anm_rec > 0) {
      try {
        rec_nm = read_npy_f64(dir + "rec_nm.npy");  // DEFAULT_REC_NM_FILE
      } catch (const std::exception &e) {
        throw std::runtime_error("Error reading receptor ANM file [\"rec_nm.npy\"]: " + rust_debug_str(e.what()));
      }
      if (rec_nm.size() != receptor.atom_count() * 3 * setup.anm_rec)
        throw std::runtime_error("Number of read ANM in receptor does not correspond to the number of atoms");
    }
    if (setup.anm_lig > 0) {
      try {
        lig_nm = read_npy_f64(dir + "lig_nm.npy");  // DEFAULT_LIG_NM_FILE
      } catch (const std::exception &e) {
        throw std::runtime_error("Error reading ligand ANM file [\"lig_nm.npy\"]: " + rust_debug_str(e.what()));
      }
      if (lig_nm.size() != ligand.atom_count() * 3 * setup.anm_lig)
        throw std::runtime_error("Number of read ANM in ligand does not correspond to the number of atoms");
    }
  }
  auto pick = [](const std::optional<std::map<std::string, std::vector<std::string>>> &r, const char *key) {
    if (!r) return std::vector<std::string>();
    auto it = r->find(key);
    if (it == r->end())  // restraints["active"] on a missing key panics in the reference
      throw std::runtime_error(std::string("restraints map has no key ") + key);
    return it->second;
  };
  const auto rec_active = pick(setup.receptor_restraints, "active");
  const auto rec_passive = pick(setup.receptor_restraints, "passive");
  const auto lig_active = pick(setup.ligand_restraints, "active");
  const auto lig_passive = pick(setup.ligand_restraints, "passive");

  if (verbose) std::printf("Loading %s scoring function\n", method_name(method));
  switch (method) {
    case Method::DFIRE:
      lc.scoring = DFIRE::create(receptor, rec_active, rec_passive, rec_nm, setup.anm_rec, ligand, lig_active,
                                 lig_passive, lig_nm, setup.anm_lig, setup.use_anm, device);
      break;
    case Method::DNA:
      lc.scoring = DNA::create(receptor, rec_active, rec_passive, rec_nm, setup.anm_rec, ligand, lig_active,
                               lig_passive, lig_nm, setup.anm_lig, setup.use_anm, device);
      break;
    case Method::PYDOCK:
      lc.scoring = PYDOCK::create(receptor, rec_active, rec_passive, rec_nm, setup.anm_rec, ligand, lig_active,
                                  lig_passive, lig_nm, setup.anm_lig, setup.use_anm, device);
      break;
  }
  return lc;
}

}  // namespace lightdock

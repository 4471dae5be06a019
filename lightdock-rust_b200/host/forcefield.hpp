// forcefield.hpp — parameter tables of the scoring functions, loaded from the TSV data files
// (lightdock-rust_b200/data/, extracted from src/dfire.rs:18-101 and src/dna.rs:64-233 /
// src/pydock.rs:147-148,209-210 by tools/extract_forcefield_tables.py).
#pragma once
#include <string>
#include <unordered_map>

namespace lightdock {

struct ForceField {
  std::unordered_map<std::string, int> dfire_type;          // "RES\tATOM" -> DFIRE atom type
  std::unordered_map<std::string, std::string> amber_type;  // "RES-ATOM" -> AMBER type (DNA)
  std::unordered_map<std::string, std::string> amber_type_pydock;
  std::unordered_map<std::string, double> ele_charge, ele_charge_pydock, nt_ele_charge;
  std::unordered_map<std::string, double> vdw_energy, vdw_radius;  // by AMBER type

  // Directory resolution: $LIGHTDOCK_B200_DATA, else <directory of this shared object>/data.
  static const ForceField &instance();
  static std::string data_dir();
};

}  // namespace lightdock

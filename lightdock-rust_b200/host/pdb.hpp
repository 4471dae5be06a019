// pdb.hpp — minimal fixed-column PDB reader standing in for pdbtbx 0.11 (third-party crate used at
// src/bin/lightdock-rust.rs:200-214).  Only what the scoring models consume is kept: atoms in the
// order pdbtbx iterates them (chain -> residue(serial, insertion code) -> conformer(name, alt loc)
// -> atom; src/dfire.rs:133-186), with names trimmed.
#pragma once
#include <string>
#include <vector>

namespace lightdock {

struct Atom {
  std::string name;     // trimmed, columns 13-16
  std::string res_name; // residue name as pdbtbx's Residue::name() reports it
  std::string chain;    // column 22
  long res_seq = 0;     // columns 23-26
  std::string icode;    // column 27, empty if blank
  double x = 0, y = 0, z = 0;
  bool hetero = false;
  std::string record;   // the ATOM/HETATM line as read (for writing the atom back with new coordinates)
};

struct PDB {
  std::vector<Atom> atoms;  // iteration order of structure.chains().residues().atoms()
  size_t atom_count() const { return atoms.size(); }
};

// Throws std::runtime_error on I/O or parse errors (the reference `unwrap()`s the pdbtbx result).
PDB open_pdb(const std::string &path);

// "{chain}.{res_name}.{serial}{icode}", src/dfire.rs:138-141
std::string residue_id(const Atom &a);

// The atom's record with columns 31-54 replaced by the given coordinates (%8.3f each), padded to 54 columns at least.
std::string atom_record_at(const Atom &a, double x, double y, double z);

}  // namespace lightdock

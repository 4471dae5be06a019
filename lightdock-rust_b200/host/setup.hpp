// setup.hpp — input files of the driver: setup.json (`SetupFile`, src/bin/lightdock-rust.rs:27-58),
// start positions (`parse_input_coordinates`, :60-75) and the flattened ANM .npy files (:216-254).
// serde_json / npyz are third-party crates in the reference; tiny stand-ins live here.
#pragma once
#include <cstdint>
#include <map>
#include <optional>
#include <string>
#include <vector>

namespace lightdock {

struct SetupFile {
  std::optional<uint64_t> seed;
  uint64_t anm_seed = 0;
  bool noh = false;
  size_t anm_rec = 0, anm_lig = 0;
  uint32_t swarms = 0, starting_points_seed = 0, glowworms = 0;
  bool verbose_parser = false, noxt = false, now = false, use_anm = false, membrane = false;
  std::string receptor_pdb, ligand_pdb;
  std::optional<std::map<std::string, std::vector<std::string>>> receptor_restraints, ligand_restraints;
};

// Throws std::runtime_error with a serde-like message on malformed input or a missing required key.
SetupFile read_setup_from_file(const std::string &path);
std::vector<std::vector<double>> parse_input_coordinates(const std::string &swarm_filename);
// `token.parse::<f64>()` as Rust accepts it (no hex floats, no nan(...), no white space); false = ParseFloatError.
bool parse_f64_like_rust(const std::string &token, double &out);
// 1-D (or any C-order) little-endian f64 .npy -> flat vector
std::vector<double> read_npy_f64(const std::string &path);
std::optional<int> parse_swarm_id(const std::string &path);  // :150-156

}  // namespace lightdock

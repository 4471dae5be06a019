// simulate.hpp — what `simulate` does before the optimisation starts (src/bin/lightdock-rust.rs:158-316):
// resolve paths, read structures / ANM / restraints and build the scoring object.
#pragma once
#include <memory>
#include <string>

#include "scoring.hpp"
#include "setup.hpp"

namespace lightdock {

struct LoadedCase {
  SetupFile setup;
  uint64_t seed = 0;
  std::unique_ptr<Score> scoring;
  Method method = Method::DFIRE;
};

// simulation_path: directory of setup.json (PDBs resolve against it); anm_dir: where rec_nm.npy /
// lig_nm.npy live (the reference reads them from the current directory).  `verbose` reproduces the
// reference's progress lines on stdout.
LoadedCase load_case(const std::string &simulation_path, const SetupFile &setup, Method method,
                     const std::string &anm_dir, int device, bool verbose);

std::string rust_debug_str(const std::string &s);  // {:?} formatting of a string

}  // namespace lightdock

// lightdock-rust-conformations — the predicted complexes of a swarm as PDB files (what the Python
// tool-chain's lgd_generate_conformations.py does downstream of the reference, example/1czy/analysis.sh:12), computed by
// the library's own pose transform (ld_transform_batch: src/dfire.rs:282-320 on the device, bit-identical to the
// coordinates the pair loop sees).
//
//   lightdock-rust-conformations <setup.json> <swarm_N/gso_<step>.out> <dfire|dna|pydock> [glowworm ids, default all]
//
// Inputs resolve like the driver's: PDBs (prefix lightdock_) next to setup.json, rec_nm.npy / lig_nm.npy and
// data/DCparams relative to the CWD.  Output: lightdock_<glowworm>.pdb next to the gso file -- the receptor (deformed by
// the pose's receptor ANM extents when use_anm) followed by the ligand at the pose, every ATOM/HETATM record as read,
// columns 31-54 replaced.
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lightdock_b200.h"
#include "pdb.hpp"
#include "scoring.hpp"
#include "setup.hpp"
#include "simulate.hpp"

using namespace lightdock;

// "(v, v, ...)    0    0   lum  nn vis score" (src/swarm.rs:131-164) -> the values between the parentheses
static std::vector<std::vector<double>> read_gso_poses(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open " + path);
  std::vector<std::vector<double>> out;
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#') continue;
    const size_t a = line.find('('), b = line.find(')');
    if (a == std::string::npos || b == std::string::npos || b < a) throw std::runtime_error("malformed line in " + path);
    std::vector<double> row;
    size_t i = a + 1;
    while (i < b) {
      size_t j = line.find(',', i);
      if (j == std::string::npos || j > b) j = b;
      std::string tok = line.substr(i, j - i);
      tok.erase(0, tok.find_first_not_of(' '));
      double v;
      if (!parse_f64_like_rust(tok, v)) throw std::runtime_error("malformed number '" + tok + "' in " + path);
      row.push_back(v);
      i = j + 1;
    }
    out.push_back(std::move(row));
  }
  return out;
}

int main(int argc, char **argv) {
  try {
    if (argc < 4) {
      std::fprintf(stderr, "Usage: %s setup.json swarm_N/gso_<step>.out <dfire|dna|pydock> [glowworm ids]\n", argv[0]);
      return 2;
    }
    const std::string setup_filename = argv[1], gso_filename = argv[2];
    std::string m = argv[3];
    std::transform(m.begin(), m.end(), m.begin(), [](unsigned char c) { return std::tolower(c); });
    Method method;
    if (m == "dfire") method = Method::DFIRE;
    else if (m == "dna") method = Method::DNA;
    else if (m == "pydock") method = Method::PYDOCK;
    else throw std::runtime_error("method not supported");
    const SetupFile setup = read_setup_from_file(setup_filename);
    const size_t slash = setup_filename.find_last_of('/');
    const std::string simulation_path = slash == std::string::npos ? "" : setup_filename.substr(0, slash);
    const auto rows = read_gso_poses(gso_filename);
    std::vector<size_t> ids;
    for (int i = 4; i < argc; ++i) {
      char *end = nullptr;
      const unsigned long long v = std::strtoull(argv[i], &end, 10);
      if (*end != '\0' || v >= rows.size()) throw std::runtime_error(std::string("no glowworm ") + argv[i] + " in " + gso_filename);
      ids.push_back((size_t)v);
    }
    if (ids.empty())
      for (size_t i = 0; i < rows.size(); ++i) ids.push_back(i);

    int device = 0;
    if (const char *dev = std::getenv("LIGHTDOCK_B200_DEVICE")) device = std::atoi(dev);
    LoadedCase lc = load_case(simulation_path, setup, method, "", device, false);
    const auto *cs = dynamic_cast<const CudaScore *>(lc.scoring.get());
    const size_t pl = lc.scoring->pose_len();
    const std::string prefix = simulation_path.empty() ? std::string("lightdock_") : simulation_path + "/lightdock_";
    const PDB receptor = open_pdb(prefix + setup.receptor_pdb), ligand = open_pdb(prefix + setup.ligand_pdb);
    const size_t nr = receptor.atom_count(), nl = ligand.atom_count();

    std::vector<double> poses(ids.size() * pl);
    for (size_t k = 0; k < ids.size(); ++k) {
      const auto &row = rows[ids[k]];
      if (row.size() < pl) throw std::runtime_error("glowworm row has fewer columns than the pose of this set-up");
      std::copy(row.begin(), row.begin() + pl, poses.begin() + k * pl);
    }
    std::vector<double> rec(ids.size() * nr * 3), lig(ids.size() * nl * 3);
    if (ld_transform_batch(cs->handle(), (int64_t)ids.size(), poses.data(), rec.data(), lig.data()) != LD_OK)
      throw std::runtime_error(std::string("ld_transform_batch: ") + ld_last_error());

    const size_t dslash = gso_filename.find_last_of('/');
    const std::string out_dir = dslash == std::string::npos ? "" : gso_filename.substr(0, dslash + 1);
    for (size_t k = 0; k < ids.size(); ++k) {
      const std::string path = out_dir + "lightdock_" + std::to_string(ids[k]) + ".pdb";
      FILE *f = std::fopen(path.c_str(), "w");
      if (!f) throw std::runtime_error("cannot create " + path);
      const double *r = rec.data() + k * nr * 3, *l = lig.data() + k * nl * 3;
      for (size_t i = 0; i < nr; ++i)
        std::fprintf(f, "%s\n", atom_record_at(receptor.atoms[i], r[3 * i], r[3 * i + 1], r[3 * i + 2]).c_str());
      std::fprintf(f, "TER\n");
      for (size_t j = 0; j < nl; ++j)
        std::fprintf(f, "%s\n", atom_record_at(ligand.atoms[j], l[3 * j], l[3 * j + 1], l[3 * j + 2]).c_str());
      std::fprintf(f, "TER\nEND\n");
      std::fclose(f);
    }
    std::printf("Wrote %zu structures next to %s\n", ids.size(), gso_filename.c_str());
    std::fflush(nullptr);
    _exit(0);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "lightdock-rust-conformations: %s\n", e.what());
    return 1;
  }
}

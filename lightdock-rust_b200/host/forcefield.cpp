#include "forcefield.hpp"

#include <dlfcn.h>

#include <cstdlib>
#include <fstream>
#include <stdexcept>

namespace lightdock {

std::string ForceField::data_dir() {
  if (const char *e = std::getenv("LIGHTDOCK_B200_DATA")) return e;
  Dl_info info;
  if (dladdr(reinterpret_cast<void *>(&ForceField::data_dir), &info) && info.dli_fname) {
    std::string p(info.dli_fname);
    const size_t slash = p.find_last_of('/');
    const std::string dir = slash == std::string::npos ? "." : p.substr(0, slash);
    for (const char *rel : {"/data", "/../data"}) {
      std::ifstream probe(dir + rel + "/dfire_atom_types.tsv");
      if (probe) return dir + rel;
    }
  }
  return "data";
}

template <typename F>
static void read_tsv(const std::string &path, F &&row) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("Unable to open parameter table " + path);
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#') continue;
    row(line);
  }
}

static ForceField load() {
  ForceField ff;
  const std::string d = ForceField::data_dir();
  read_tsv(d + "/dfire_atom_types.tsv", [&](const std::string &l) {
    const size_t t = l.rfind('\t');
    ff.dfire_type[l.substr(0, t)] = std::atoi(l.c_str() + t + 1);
  });
  auto str_table = [&](const std::string &f, std::unordered_map<std::string, std::string> &m) {
    read_tsv(d + "/" + f, [&](const std::string &l) {
      const size_t t = l.find('\t');
      m[l.substr(0, t)] = l.substr(t + 1);
    });
  };
  auto num_table = [&](const std::string &f, std::unordered_map<std::string, double> &m) {
    read_tsv(d + "/" + f, [&](const std::string &l) {
      const size_t t = l.find('\t');
      m[l.substr(0, t)] = std::strtod(l.c_str() + t + 1, nullptr);
    });
  };
  str_table("amber_types.tsv", ff.amber_type);
  ff.amber_type_pydock = ff.amber_type;
  str_table("amber_types_pydock_extra.tsv", ff.amber_type_pydock);
  num_table("ele_charges.tsv", ff.ele_charge);
  ff.ele_charge_pydock = ff.ele_charge;
  num_table("ele_charges_pydock_extra.tsv", ff.ele_charge_pydock);
  num_table("nt_ele_charges.tsv", ff.nt_ele_charge);
  num_table("vdw_energy.tsv", ff.vdw_energy);
  num_table("vdw_radius.tsv", ff.vdw_radius);
  return ff;
}

const ForceField &ForceField::instance() {
  static const ForceField ff = load();
  return ff;
}

}  // namespace lightdock

#include "pdb.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <stdexcept>
#include <utility>

namespace lightdock {

static std::string trim(const std::string &s) {
  size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}
static std::string cols(const std::string &line, size_t from, size_t to) {  // 0-based [from, to)
  if (line.size() <= from) return "";
  return line.substr(from, std::min(to, line.size()) - from);
}
static double parse_f64(const std::string &s, const std::string &what, size_t lineno) {
  const std::string t = trim(s);
  char *end = nullptr;
  const double v = std::strtod(t.c_str(), &end);
  if (t.empty() || end == t.c_str() || *end != '\0')
    throw std::runtime_error("PDB parse error: bad " + what + " at line " + std::to_string(lineno));
  return v;
}

namespace {
struct Conformer {
  std::string name, alt;
  std::vector<Atom> atoms;
};
struct Residue {
  long serial;
  std::string icode;
  std::vector<Conformer> conformers;
};
struct Chain {
  std::string id;
  std::vector<Residue> residues;
};
}  // namespace

PDB open_pdb(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("Could not open PDB file " + path);
  std::vector<Chain> chains;
  std::string line;
  size_t lineno = 0;
  while (std::getline(in, line)) {
    ++lineno;
    const std::string rec = cols(line, 0, 6);
    if (rec == "ENDMDL") break;  // first model only
    const bool het = rec == "HETATM";
    if (rec != "ATOM  " && !het) continue;
    Atom a;
    a.hetero = het;
    a.name = trim(cols(line, 12, 16));
    const std::string alt = trim(cols(line, 16, 17));
    const std::string res_name = trim(cols(line, 17, 20));
    a.chain = cols(line, 21, 22);
    const std::string serial = trim(cols(line, 22, 26));
    char *end = nullptr;
    a.res_seq = std::strtol(serial.c_str(), &end, 10);
    if (serial.empty() || *end != '\0')
      throw std::runtime_error("PDB parse error: bad residue number at line " + std::to_string(lineno));
    a.icode = trim(cols(line, 26, 27));
    a.x = parse_f64(cols(line, 30, 38), "x", lineno);
    a.y = parse_f64(cols(line, 38, 46), "y", lineno);
    a.z = parse_f64(cols(line, 46, 54), "z", lineno);
    a.res_name = res_name;
    a.record = line;
    while (!a.record.empty() && (a.record.back() == '\r' || a.record.back() == '\n')) a.record.pop_back();
    // chain: first with the same id, else a new one
    Chain *ch = nullptr;
    for (auto &c : chains)
      if (c.id == a.chain) { ch = &c; break; }
    if (!ch) { chains.push_back(Chain{a.chain, {}}); ch = &chains.back(); }
    // residue: search from the back (atoms of a residue are normally contiguous)
    Residue *rs = nullptr;
    for (auto it = ch->residues.rbegin(); it != ch->residues.rend(); ++it)
      if (it->serial == a.res_seq && it->icode == a.icode) { rs = &*it; break; }
    if (!rs) { ch->residues.push_back(Residue{a.res_seq, a.icode, {}}); rs = &ch->residues.back(); }
    Conformer *cf = nullptr;
    for (auto &c : rs->conformers)
      if (c.name == res_name && c.alt == alt) { cf = &c; break; }
    if (!cf) { rs->conformers.push_back(Conformer{res_name, alt, {}}); cf = &rs->conformers.back(); }
    cf->atoms.push_back(std::move(a));
  }
  PDB pdb;
  for (auto &c : chains)
    for (auto &r : c.residues) {
      // Residue::name() is only defined when all conformers agree; the reference panics otherwise
      // ("PDB Parsing Error: Residue name error", src/dfire.rs:134-137)
      for (auto &cf : r.conformers)
        if (cf.name != r.conformers.front().name)
          throw std::runtime_error("PDB Parsing Error: Residue name error");
      for (auto &cf : r.conformers)
        for (auto &a : cf.atoms) pdb.atoms.push_back(a);
    }
  return pdb;
}

std::string atom_record_at(const Atom &a, double x, double y, double z) {
  std::string r = a.record;
  if (r.size() < 54) r.resize(54, ' ');
  char buf[32];
  std::snprintf(buf, sizeof buf, "%8.3f%8.3f%8.3f", x, y, z);
  r.replace(30, 24, std::string(buf).substr(0, 24));
  return r;
}

std::string residue_id(const Atom &a) {
  return a.chain + "." + a.res_name + "." + std::to_string(a.res_seq) + a.icode;
}

}  // namespace lightdock

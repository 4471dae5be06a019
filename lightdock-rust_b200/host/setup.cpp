#include "setup.hpp"

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace lightdock {

// ---- a small JSON value + recursive-descent parser (objects, arrays, strings, numbers, literals)
namespace {
struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string text;  // string value, or the raw number token
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;
  const Json *get(const std::string &k) const {
    const Json *hit = nullptr;
    for (const auto &kv : obj)
      if (kv.first == k) hit = &kv.second;  // last duplicate wins, as serde does for structs? (it errors) — keep last
    return hit;
  }
};
struct Parser {
  const std::string &s;
  size_t i = 0;
  explicit Parser(const std::string &src) : s(src) {}
  [[noreturn]] void err(const std::string &m) const {
    throw std::runtime_error(m + " at offset " + std::to_string(i));
  }
  void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
  Json value() {
    ws();
    if (i >= s.size()) err("EOF while parsing a value");
    const char c = s[i];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') { Json j; j.kind = Json::String; j.text = string(); return j; }
    if (!s.compare(i, 4, "true")) { i += 4; Json j; j.kind = Json::Bool; j.b = true; return j; }
    if (!s.compare(i, 5, "false")) { i += 5; Json j; j.kind = Json::Bool; j.b = false; return j; }
    if (!s.compare(i, 4, "null")) { i += 4; return Json{}; }
    return number();
  }
  Json number() {
    const size_t st = i;
    while (i < s.size() && (std::isdigit((unsigned char)s[i]) || std::strchr("+-.eE", s[i]))) ++i;
    if (st == i) err("expected value");
    Json j; j.kind = Json::Number; j.text = s.substr(st, i - st); j.num = std::strtod(j.text.c_str(), nullptr);
    return j;
  }
  std::string string() {
    std::string out;
    ++i;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < s.size()) {
        const char e = s[++i];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            unsigned cp = std::strtoul(s.substr(i + 1, 4).c_str(), nullptr, 16);
            i += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += e;
        }
        ++i;
      } else {
        out += s[i++];
      }
    }
    if (i >= s.size()) err("EOF while parsing a string");
    ++i;
    return out;
  }
  Json array() {
    Json j; j.kind = Json::Array;
    ++i; ws();
    if (i < s.size() && s[i] == ']') { ++i; return j; }
    for (;;) {
      j.arr.push_back(value());
      ws();
      if (i < s.size() && s[i] == ',') { ++i; continue; }
      if (i < s.size() && s[i] == ']') { ++i; return j; }
      err("expected `,` or `]`");
    }
  }
  Json object() {
    Json j; j.kind = Json::Object;
    ++i; ws();
    if (i < s.size() && s[i] == '}') { ++i; return j; }
    for (;;) {
      ws();
      if (i >= s.size() || s[i] != '"') err("key must be a string");
      std::string k = string();
      ws();
      if (i >= s.size() || s[i] != ':') err("expected `:`");
      ++i;
      j.obj.emplace_back(std::move(k), value());
      ws();
      if (i < s.size() && s[i] == ',') { ++i; continue; }
      if (i < s.size() && s[i] == '}') { ++i; return j; }
      err("expected `,` or `}`");
    }
  }
};

const Json &required(const Json &root, const char *key) {
  const Json *v = root.get(key);
  if (!v) throw std::runtime_error(std::string("missing field `") + key + "`");
  return *v;
}
bool as_bool(const Json &root, const char *key) {
  const Json &v = required(root, key);
  if (v.kind != Json::Bool) throw std::runtime_error(std::string("invalid type for `") + key + "`, expected a boolean");
  return v.b;
}
uint64_t as_u64(const Json &v, const char *key) {
  if (v.kind != Json::Number || v.text.find_first_of(".eE-") != std::string::npos)
    throw std::runtime_error(std::string("invalid type for `") + key + "`, expected an unsigned integer");
  return std::strtoull(v.text.c_str(), nullptr, 10);
}
std::string as_string(const Json &root, const char *key) {
  const Json &v = required(root, key);
  if (v.kind != Json::String) throw std::runtime_error(std::string("invalid type for `") + key + "`, expected a string");
  return v.text;
}
std::optional<std::map<std::string, std::vector<std::string>>> as_restraints(const Json &root, const char *key) {
  const Json *v = root.get(key);
  if (!v || v->kind == Json::Null) return std::nullopt;
  if (v->kind != Json::Object) throw std::runtime_error(std::string("invalid type for `") + key + "`, expected a map");
  std::map<std::string, std::vector<std::string>> out;
  for (const auto &kv : v->obj) {
    if (kv.second.kind != Json::Array) throw std::runtime_error(std::string("invalid type in `") + key + "`");
    std::vector<std::string> list;
    for (const Json &e : kv.second.arr) {
      if (e.kind != Json::String) throw std::runtime_error(std::string("invalid type in `") + key + "`");
      list.push_back(e.text);
    }
    out[kv.first] = std::move(list);
  }
  return out;
}
}  // namespace

SetupFile read_setup_from_file(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("No such file or directory (os error 2)");
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string text = ss.str();
  Parser p(text);
  const Json root = p.value();
  p.ws();
  if (p.i != text.size()) p.err("trailing characters");
  if (root.kind != Json::Object) throw std::runtime_error("invalid type: expected struct SetupFile");
  SetupFile s;
  if (const Json *v = root.get("seed"); v && v->kind != Json::Null) s.seed = as_u64(*v, "seed");
  s.anm_seed = as_u64(required(root, "anm_seed"), "anm_seed");
  s.noh = as_bool(root, "noh");
  s.anm_rec = as_u64(required(root, "anm_rec"), "anm_rec");
  s.anm_lig = as_u64(required(root, "anm_lig"), "anm_lig");
  s.swarms = (uint32_t)as_u64(required(root, "swarms"), "swarms");
  s.starting_points_seed = (uint32_t)as_u64(required(root, "starting_points_seed"), "starting_points_seed");
  s.verbose_parser = as_bool(root, "verbose_parser");
  s.noxt = as_bool(root, "noxt");
  s.now = as_bool(root, "now");
  s.use_anm = as_bool(root, "use_anm");
  s.glowworms = (uint32_t)as_u64(required(root, "glowworms"), "glowworms");
  s.membrane = as_bool(root, "membrane");
  s.receptor_pdb = as_string(root, "receptor_pdb");
  s.ligand_pdb = as_string(root, "ligand_pdb");
  s.receptor_restraints = as_restraints(root, "receptor_restraints");
  s.ligand_restraints = as_restraints(root, "ligand_restraints");
  return s;
}

// `str::parse::<f64>()` (Rust core, dec2flt): an optional sign, then either decimal digits with an optional fraction and
// an optional exponent (at least one digit overall), or "inf" / "infinity" / "nan" in any case.  No surrounding white
// space, no hexadecimal floats, no "nan(...)", no digit separators: everything strtod accepts beyond that is refused
// here so that a start file the reference would panic on is not silently read.
bool parse_f64_like_rust(const std::string &tok, double &out) {
  size_t i = 0;
  const size_t n = tok.size();
  if (i < n && (tok[i] == '+' || tok[i] == '-')) ++i;
  if (i == n) return false;
  auto ieq = [&](const char *w) {
    size_t k = 0;
    for (; w[k]; ++k)
      if (i + k >= n || std::tolower((unsigned char)tok[i + k]) != w[k]) return false;
    return i + k == n;
  };
  if (!(ieq("inf") || ieq("infinity") || ieq("nan"))) {
    size_t digits = 0, j = i;
    while (j < n && std::isdigit((unsigned char)tok[j])) { ++j; ++digits; }
    if (j < n && tok[j] == '.') {
      ++j;
      while (j < n && std::isdigit((unsigned char)tok[j])) { ++j; ++digits; }
    }
    if (digits == 0) return false;
    if (j < n && (tok[j] == 'e' || tok[j] == 'E')) {
      ++j;
      if (j < n && (tok[j] == '+' || tok[j] == '-')) ++j;
      size_t ed = 0;
      while (j < n && std::isdigit((unsigned char)tok[j])) { ++j; ++ed; }
      if (ed == 0) return false;
    }
    if (j != n) return false;
  }
  char *end = nullptr;
  out = std::strtod(tok.c_str(), &end);  // correctly rounded, like dec2flt
  return *end == '\0';
}

std::vector<std::vector<double>> parse_input_coordinates(const std::string &swarm_filename) {
  std::ifstream in(swarm_filename);
  if (!in) throw std::runtime_error("Error reading the input file");
  std::vector<std::vector<double>> positions;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::vector<double> position;
    size_t st = 0;
    for (;;) {  // split on single spaces like `split(' ')`: an empty token fails to parse, as in the reference
      const size_t sp = line.find(' ', st);
      std::string tok = line.substr(st, sp == std::string::npos ? std::string::npos : sp - st);
      size_t a = tok.find_first_not_of(" \t");
      size_t b = tok.find_last_not_of(" \t");
      tok = a == std::string::npos ? "" : tok.substr(a, b - a + 1);
      double v = 0.0;
      if (!parse_f64_like_rust(tok, v))
        throw std::runtime_error("called `Result::unwrap()` on an `Err` value: ParseFloatError (start positions)");
      position.push_back(v);
      if (sp == std::string::npos) break;
      st = sp + 1;
    }
    positions.push_back(std::move(position));
  }
  return positions;
}

std::vector<double> read_npy_f64(const std::string &path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("No such file or directory (os error 2)");
  std::vector<char> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  if (bytes.size() < 10 || std::memcmp(bytes.data(), "\x93NUMPY", 6) != 0) throw std::runtime_error("not an NPY file");
  const int major = (unsigned char)bytes[6];
  size_t hlen, hoff;
  if (major == 1) { hlen = (unsigned char)bytes[8] | ((unsigned char)bytes[9] << 8); hoff = 10; }
  else {
    if (bytes.size() < 12) throw std::runtime_error("truncated NPY header");
    hlen = (unsigned char)bytes[8] | ((unsigned char)bytes[9] << 8) | ((unsigned char)bytes[10] << 16) |
           ((size_t)(unsigned char)bytes[11] << 24);
    hoff = 12;
  }
  if (hoff + hlen > bytes.size()) throw std::runtime_error("truncated NPY header");
  const std::string header(bytes.data() + hoff, hlen);
  if (header.find("'<f8'") == std::string::npos && header.find("'|f8'") == std::string::npos &&
      header.find("'=f8'") == std::string::npos)
    throw std::runtime_error("NPY file is not little-endian float64");
  if (header.find("'fortran_order': True") != std::string::npos) throw std::runtime_error("NPY file is Fortran-ordered");
  const size_t sp = header.find("'shape':");
  const size_t lp = header.find('(', sp), rp = header.find(')', lp);
  if (sp == std::string::npos || lp == std::string::npos || rp == std::string::npos)
    throw std::runtime_error("NPY header has no shape");
  size_t count = 1;
  std::string dims = header.substr(lp + 1, rp - lp - 1);
  std::stringstream ds(dims);
  std::string tok;
  while (std::getline(ds, tok, ',')) {
    const size_t a = tok.find_first_not_of(' ');
    if (a == std::string::npos) continue;
    count *= std::strtoull(tok.c_str() + a, nullptr, 10);
  }
  const size_t data = hoff + hlen;
  if (data + count * 8 > bytes.size()) throw std::runtime_error("NPY data shorter than its shape");
  std::vector<double> out(count);
  std::memcpy(out.data(), bytes.data() + data, count * 8);
  return out;
}

std::optional<int> parse_swarm_id(const std::string &path) {
  const size_t slash = path.find_last_of('/');
  std::string name = slash == std::string::npos ? path : path.substr(slash + 1);
  const std::string prefix = "initial_positions_", suffix = ".dat";
  if (name.size() < prefix.size() + suffix.size() || name.compare(0, prefix.size(), prefix) != 0 ||
      name.compare(name.size() - suffix.size(), suffix.size(), suffix) != 0)
    return std::nullopt;
  const std::string num = name.substr(prefix.size(), name.size() - prefix.size() - suffix.size());
  if (num.empty()) return std::nullopt;
  size_t k = (num[0] == '-' || num[0] == '+') ? 1 : 0;
  if (k == num.size()) return std::nullopt;
  for (size_t i = k; i < num.size(); ++i)
    if (!std::isdigit((unsigned char)num[i])) return std::nullopt;
  return std::atoi(num.c_str());
}

}  // namespace lightdock

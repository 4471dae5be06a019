// lightdock-rust-multi — all swarms of a docking run in ONE process, sharded over the GPUs of the box.
//
// The reference runs one `lightdock-rust` process per swarm and leaves the fan-out to an external scheduler
// (`ant_thony.py --cores N task.list`, example/1czy/execution.sh:21-25).  On a GPU that wastes the device: 200 poses
// per launch.  This driver takes the same work description and advances every swarm in lock-step with one batched
// launch per step and device (MultiGSO, host/gso.hpp):
//
//   lightdock-rust-multi task.list                                        # the ant_thony task file, as is
//   lightdock-rust-multi <setup.json> <steps> <dfire|dna|pydock> <initial_positions_N.dat>...
//
// Each swarm keeps its own StdRng seeded like a stand-alone process, writes the same swarm_N/gso_<step>.out files
// (relative to the CWD), and its trajectory is bit-identical to `lightdock-rust setup.json initial_positions_N.dat
// steps method` (tests/test_gpu_trajectory.py).  Swarm i of the list goes to device i mod G; devices are
// $LIGHTDOCK_B200_DEVICES (comma-separated ordinals) or every visible one.  No data crosses GPUs.
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "gso.hpp"
#include "pdb.hpp"
#include "sharding.hpp"
#include "simulate.hpp"

using namespace lightdock;

namespace {
struct Job {
  std::string setup, method;
  unsigned long long steps = 0;
  std::vector<std::string> swarm_files;
};

[[noreturn]] void usage(const char *argv0) {
  std::fprintf(stderr,
               "Usage: %s task.list\n       %s setup_filename steps method swarm_filename [swarm_filename ...]\n", argv0,
               argv0);
  std::exit(2);
}

// One ant_thony line: "<binary> setup.json init/initial_positions_7.dat 100 dfire;"
Job parse_task_list(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open task list " + path);
  Job job;
  std::string line;
  while (std::getline(in, line)) {
    const size_t semi = line.find(';');
    if (semi != std::string::npos) line.resize(semi);
    std::istringstream ss(line);
    std::vector<std::string> tok;
    for (std::string t; ss >> t;) tok.push_back(t);
    if (tok.empty()) continue;
    if (tok.size() != 5) throw std::runtime_error("task list line is not `<binary> setup swarm_file steps method`: " + line);
    char *end = nullptr;
    const unsigned long long steps = std::strtoull(tok[3].c_str(), &end, 10);
    if (*end != '\0') throw std::runtime_error("steps argument must be a number: " + line);
    if (job.swarm_files.empty()) {
      job.setup = tok[1]; job.steps = steps; job.method = tok[4];
    } else if (job.setup != tok[1] || job.steps != steps || job.method != tok[4]) {
      throw std::runtime_error("all tasks must share setup file, steps and method: " + line);
    }
    job.swarm_files.push_back(tok[2]);
  }
  if (job.swarm_files.empty()) throw std::runtime_error("task list is empty");
  return job;
}

std::vector<int> pick_devices() {
  std::vector<int> dev;
  if (const char *e = std::getenv("LIGHTDOCK_B200_DEVICES")) {
    std::stringstream ss(e);
    for (std::string t; std::getline(ss, t, ',');)
      if (!t.empty()) dev.push_back(std::atoi(t.c_str()));
  } else {
    for (int d = 0; d < ld_device_count(); ++d) dev.push_back(d);
  }
  if (dev.empty()) throw std::runtime_error("no CUDA device available (there is no CPU fallback)");
  return dev;
}
}  // namespace

int main(int argc, char **argv) {
  try {
    Job job;
    if (argc == 2) {
      job = parse_task_list(argv[1]);
    } else if (argc >= 5) {
      job.setup = argv[1];
      char *end = nullptr;
      job.steps = std::strtoull(argv[2], &end, 10);
      if (*end != '\0') {
        std::fprintf(stderr, "Error: steps argument must be a number\n");
        return 2;
      }
      job.method = argv[3];
      for (int i = 4; i < argc; ++i) job.swarm_files.push_back(argv[i]);
    } else {
      usage(argv[0]);
    }
    if (job.steps > 0xffffffffULL) throw std::runtime_error("steps argument must be a number");
    std::string m = job.method;
    std::transform(m.begin(), m.end(), m.begin(), [](unsigned char c) { return std::tolower(c); });
    Method method;
    if (m == "dfire") method = Method::DFIRE;
    else if (m == "dna") method = Method::DNA;
    else if (m == "pydock") method = Method::PYDOCK;
    else throw std::runtime_error("method not supported");
    const SetupFile setup = read_setup_from_file(job.setup);
    const size_t slash = job.setup.find_last_of('/');
    const std::string simulation_path = slash == std::string::npos ? "" : job.setup.substr(0, slash);

    // swarm ids, output directories and start positions (same rules as the single-swarm driver)
    const size_t ns = job.swarm_files.size();
    std::vector<std::string> dirs(ns);
    std::vector<std::vector<std::vector<double>>> positions(ns);
    for (size_t s = 0; s < ns; ++s) {
      const std::optional<int> id = parse_swarm_id(job.swarm_files[s]);
      if (!id) throw std::runtime_error("Could not parse swarm from swarm filename " + job.swarm_files[s]);
      dirs[s] = "swarm_" + std::to_string(*id);
      struct stat st;
      if (stat(dirs[s].c_str(), &st) != 0 || !S_ISDIR(st.st_mode))
        if (mkdir(dirs[s].c_str(), 0777) != 0) throw std::runtime_error("Error creating directory " + dirs[s]);
      positions[s] = parse_input_coordinates(job.swarm_files[s]);
    }

    const std::vector<int> devices = pick_devices();
    const size_t G = std::min(devices.size(), ns);
    const int host_threads = std::max(1, (int)std::thread::hardware_concurrency() / (int)G);
    std::printf("%zu swarms, %llu steps, %s scoring on %zu GPU(s), %d host threads each\n", ns, job.steps,
                method_name(method), G, host_threads);
    std::fflush(stdout);
    // which GPU scores which swarm: cost-aware and deterministic (host/sharding.hpp); results do not depend on it
    std::vector<std::vector<size_t>> mine(G);
    {
      const std::string dir = simulation_path.empty() ? std::string("lightdock_") : simulation_path + "/lightdock_";
      const PDB rec_pdb = open_pdb(dir + setup.receptor_pdb);  // the files load_case reads (src/constants.rs:18)
      const PDB lig_pdb = open_pdb(dir + setup.ligand_pdb);
      std::vector<double> rxyz, lxyz;
      for (const auto &a : rec_pdb.atoms) { rxyz.push_back(a.x); rxyz.push_back(a.y); rxyz.push_back(a.z); }
      for (const auto &a : lig_pdb.atoms) { lxyz.push_back(a.x); lxyz.push_back(a.y); lxyz.push_back(a.z); }
      const SwarmCostModel model(rxyz, lxyz);
      std::vector<double> cost(ns);
      for (size_t s = 0; s < ns; ++s) {
        double c[3] = {0, 0, 0};
        for (const auto &p : positions[s])
          for (int d = 0; d < 3 && p.size() >= 3; ++d) c[d] += p[d] / (double)positions[s].size();
        cost[s] = model.cost(c);
      }
      const std::vector<int> gpu_of = assign_swarms_lpt(cost, (int)G);
      for (size_t s = 0; s < ns; ++s) mine[(size_t)gpu_of[s]].push_back(s);
    }
    std::vector<std::exception_ptr> errors(G);
    std::vector<uint64_t> calls(G, 0);
    std::vector<std::vector<std::pair<std::string, std::string>>> failed(G);  // (swarm dir, message) per GPU
    std::vector<std::thread> drivers;
    for (size_t g = 0; g < G; ++g)
      drivers.emplace_back([&, g] {
        try {
          LoadedCase lc = load_case(simulation_path, setup, method, "", devices[g], false);
          auto drive = [&](auto &multi) {
            for (size_t s : mine[g])
              multi.add(positions[s], lc.seed, setup.use_anm, setup.anm_rec, setup.anm_lig, dirs[s]);
            multi.run((uint32_t)job.steps, host_threads);
            calls[g] = multi.energy_calls();
            for (const auto &f : multi.failures()) failed[g].emplace_back(dirs[mine[g][f.first]], f.second);
          };
          const char *gso_mode = std::getenv("LIGHTDOCK_GSO");  // "device": the GSO step itself runs on the GPU
          if (gso_mode && std::string(gso_mode) == "device") {
            DeviceGSO multi(lc.scoring.get());
            drive(multi);
          } else {
            MultiGSO multi(lc.scoring.get());
            drive(multi);
          }
        } catch (...) {
          errors[g] = std::current_exception();
        }
      });
    for (auto &d : drivers) d.join();
    for (auto &e : errors)
      if (e) std::rethrow_exception(e);
    uint64_t total = 0;
    for (uint64_t c : calls) total += c;
    std::printf("Done: %llu poses scored\n", (unsigned long long)total);
    // a swarm that hit what is a panic in the reference stopped alone (one process per swarm there); report and fail
    size_t n_failed = 0;
    for (const auto &fg : failed)
      for (const auto &f : fg) {
        std::fprintf(stderr, "lightdock-rust-multi: %s failed: %s\n", f.first.c_str(), f.second.c_str());
        ++n_failed;
      }
    std::fflush(nullptr);
    _exit(n_failed ? 1 : 0);  // skip the tear-down of up to eight CUDA contexts: every file is written
  } catch (const std::exception &e) {
    std::fprintf(stderr, "lightdock-rust-multi: %s\n", e.what());
    return 1;
  }
}

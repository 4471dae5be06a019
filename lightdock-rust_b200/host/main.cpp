// lightdock-rust (B200 build) — drop-in for the reference driver src/bin/lightdock-rust.rs:
//   lightdock-rust <setup.json> <initial_positions_N.dat> <steps> <dfire|dna|pydock>
// Same argv, same progress lines, same path-resolution quirks (PDBs relative to the setup.json
// directory; rec_nm.npy / lig_nm.npy, data/DCparams and swarm_N/ relative to the CWD), same
// swarm_N/gso_<step>.out files.  Scoring runs on CUDA device $LIGHTDOCK_B200_DEVICE (default 0).
// LIGHTDOCK_GSO=device moves the whole GSO step onto the GPU as well (host/gso.hpp: DeviceGSO).
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <stdexcept>
#include <string>
#include <thread>

#include "../../include/lightdock_b200.h"
#include "gso.hpp"
#include "scoring.hpp"
#include "simulate.hpp"

using namespace lightdock;

// LDB200_TIMING=1: one line on stderr with where the wall clock of the run went (bench.py's single_swarm_runs reads
// it).  stdout stays byte-identical to the reference's.
static const auto g_t0 = std::chrono::steady_clock::now();
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_t0).count(); }

static int simulate(const std::string &simulation_path, const SetupFile &setup, const std::string &swarm_filename,
                    uint32_t steps, Method method, int device) {
  std::printf("Reading starting positions from %s\n", rust_debug_str(swarm_filename).c_str());
  const std::optional<int> swarm_id = parse_swarm_id(swarm_filename);
  if (!swarm_id) throw std::runtime_error("Could not parse swarm from swarm filename");
  std::printf("Swarm ID %d\n", *swarm_id);
  const std::string swarm_directory = "swarm_" + std::to_string(*swarm_id);
  struct stat st;
  if (stat(swarm_directory.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) {
    std::fprintf(stderr, "Output directory does not exist for swarm %d, creating it\n", *swarm_id);
    if (mkdir(swarm_directory.c_str(), 0777) != 0) throw std::runtime_error("Error creating directory");
  }
  std::printf("Writing to swarm dir %s\n", rust_debug_str(swarm_directory).c_str());
  const auto positions = parse_input_coordinates(swarm_filename);
  const double t_inputs = now_ms();
  // never destroyed: the process exits right after the run, and freeing the device buffers one by one is wasted time
  LoadedCase &lc = *new LoadedCase(load_case(simulation_path, setup, method, "", device, true));
  const double t_loaded = now_ms();
  std::printf("Creating GSO with %zu glowworms\n", positions.size());
  // LIGHTDOCK_GSO=device: the whole GSO step runs on the GPU (DeviceGSO, ld_gso_*); default: the host loop, whose
  // trajectories are byte-identical to the reference's
  const char *gso_mode = std::getenv("LIGHTDOCK_GSO");
  const bool on_device = gso_mode && std::string(gso_mode) == "device";
  uint64_t energy_calls = 0;
  if (on_device) {
    DeviceGSO gso(lc.scoring.get());
    gso.add(positions, lc.seed, setup.use_anm, setup.anm_rec, setup.anm_lig, swarm_directory);
    std::printf("Starting optimization (%u steps)\n", steps);
    std::fflush(stdout);
    gso.run(steps);
    energy_calls = gso.energy_calls();
    if (!gso.failures().empty()) throw std::runtime_error(gso.failures()[0].second);
  } else {
    GSO gso(positions, lc.seed, lc.scoring.get(), setup.use_anm, setup.anm_rec, setup.anm_lig, swarm_directory);
    std::printf("Starting optimization (%u steps)\n", steps);
    std::fflush(stdout);
    gso.run(steps);
    energy_calls = gso.swarm.energy_calls;
  }
  const double t_done = now_ms();
  if (std::getenv("LDB200_TIMING")) {
    double c[4] = {0, 0, 0, 0};
    if (const auto *cs = dynamic_cast<const CudaScore *>(lc.scoring.get())) ld_get_create_ms(cs->handle(), c);
    std::fprintf(stderr,
                 "[ldb200 timing] until_main_inputs_ms=%.1f load_case_ms=%.1f (ld_create: context_wait=%.1f complex=%.1f "
                 "groups=%.1f cells=%.1f) gso_ms=%.1f energy_calls=%llu total_in_main_ms=%.1f\n",
                 t_inputs, t_loaded - t_inputs, c[0], c[1], c[2], c[3], t_done - t_loaded,
                 (unsigned long long)energy_calls, t_done);
  }
  return 0;
}

int main(int argc, char **argv) {
  if (argc != 5) {
    std::fprintf(stderr, "Wrong command line. Usage: %s setup_filename swarm_filename steps method\n", argv[0]);
    return 0;  // the reference returns normally on CLI errors
  }
  const std::string setup_filename = argv[1], swarm_filename = argv[2], num_steps = argv[3];
  char *end = nullptr;
  const unsigned long long steps64 = std::strtoull(num_steps.c_str(), &end, 10);
  if (num_steps.empty() || *end != '\0' || num_steps[0] == '-' || steps64 > 0xffffffffULL) {
    std::fprintf(stderr, "Error: steps argument must be a number\n");
    return 0;
  }
  std::string method_type = argv[4];
  std::transform(method_type.begin(), method_type.end(), method_type.begin(), [](unsigned char c) { return std::tolower(c); });
  Method method;
  if (method_type == "dfire") method = Method::DFIRE;
  else if (method_type == "dna") method = Method::DNA;
  else if (method_type == "pydock") method = Method::PYDOCK;
  else {
    std::fprintf(stderr, "Error: method not supported\n");
    return 0;
  }
  SetupFile setup;
  try {
    setup = read_setup_from_file(setup_filename);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "Error reading setup file [%s]: %s\n", rust_debug_str(setup_filename).c_str(),
                 rust_debug_str(e.what()).c_str());
    return 0;
  }
  const size_t slash = setup_filename.find_last_of('/');
  const std::string simulation_path = slash == std::string::npos ? "" : setup_filename.substr(0, slash);
  // Scoring runs on CUDA device $LIGHTDOCK_B200_DEVICE (default 0).  Unless the user has narrowed the visible devices
  // already, narrow them to that one before the first CUDA call: on an 8-GPU box the runtime otherwise initialises all
  // eight (seconds, for a run whose 100 steps take 40 ms).  The context is then created on a helper thread while this
  // thread parses the structures, the ANM files and the 13 MB DCparams.
  int device = 0;
  if (const char *dev = std::getenv("LIGHTDOCK_B200_DEVICE")) device = std::atoi(dev);
  if (!std::getenv("CUDA_VISIBLE_DEVICES") && device >= 0) {
    setenv("CUDA_VISIBLE_DEVICES", std::to_string(device).c_str(), 1);
    device = 0;
  }
  std::thread warm([device] { ld_init_device(device); });  // errors resurface, with their message, in ld_create
  struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{warm};
  try {
    const int rc = simulate(simulation_path, setup, swarm_filename, (uint32_t)steps64, method, device);
    // Every output file is closed and the scoring object is gone; what is left is tearing down the CUDA context and the
    // runtime's worker threads (50-700 ms, measured as the noisiest part of a 0.45 s run).  The process is done: leave.
    std::fflush(nullptr);
    if (warm.joinable()) warm.join();
    _exit(rc);
  } catch (const std::exception &e) {
    // the reference panics here; mirror the message and the panic exit status
    std::fprintf(stderr, "thread '<unnamed>' panicked: %s\n", e.what());
    return 101;
  }
}

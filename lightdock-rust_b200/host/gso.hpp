// gso.hpp — host side of the Glowworm Swarm Optimisation loop, mirroring src/glowworm.rs,
// src/swarm.rs and src/lib.rs.  Control flow, neighbour selection and the seeded RNG stream stay on
// the host exactly as in the reference; the ONE body that changes is Swarm::update_luciferin, which
// gathers every glowworm that must be rescored (`moved || step == 0`) and scores them with a single
// batched Score::energy_batch call instead of one Score::energy call per glowworm.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "quaternion.hpp"
#include "scoring.hpp"
#include "stdrng.hpp"

namespace lightdock {

constexpr uint64_t DEFAULT_SEED = 324324;          // src/constants.rs:2
constexpr double DEFAULT_TRANSLATION_STEP = 0.5;   // src/constants.rs:5
constexpr double DEFAULT_ROTATION_STEP = 0.5;      // src/constants.rs:8
constexpr double DEFAULT_NMODES_STEP = 0.5;        // src/constants.rs:24

struct Glowworm {  // src/glowworm.rs:6-26
  uint32_t id = 0;
  std::vector<double> translation;
  Quaternion rotation;
  std::vector<double> rec_nmodes, lig_nmodes;
  const Score *scoring_function = nullptr;
  double rho = 0.5, gamma = 0.4, beta = 0.08;
  double luciferin = 5.0, vision_range = 0.2, max_vision_range = 5.0;
  uint32_t max_neighbors = 5;
  std::vector<uint32_t> neighbors;
  std::vector<double> probabilities;
  double scoring = 0.0;
  bool moved = false;
  uint32_t step = 0;
  bool use_anm = false;

  Glowworm(uint32_t id, std::vector<double> translation, Quaternion rotation, std::vector<double> rec_nmodes,
           std::vector<double> lig_nmodes, const Score *scoring_function, bool use_anm);

  bool needs_scoring() const { return moved || step == 0; }  // src/glowworm.rs:62
  void write_pose(double *row) const;
  void compute_luciferin();                         // src/glowworm.rs:61-72 (single-pose form)
  void apply_luciferin(double new_scoring, bool scored);  // the arithmetic of :70-71 after a batched score
  void update_vision_range();                       // :91-96
  void compute_probability_moving_toward_neighbor(const std::vector<double> &luciferins);  // :98-112
  uint32_t select_random_neighbor(double random_number);  // :114-126
  void move_towards(uint32_t other_id, const std::vector<double> &other_position, const Quaternion &other_rotation,
                    const std::vector<double> &other_anm_rec, const std::vector<double> &other_anm_lig);  // :128-190
};

double distance(const Glowworm &one, const Glowworm &two);  // src/glowworm.rs:193-202

struct Swarm {  // src/swarm.rs
  std::vector<Glowworm> glowworms;
  uint64_t energy_calls = 0;  // poses actually scored so far

  void add_glowworms(const std::vector<std::vector<double>> &positions, const Score *scoring, bool use_anm,
                     size_t rec_num_anm, size_t lig_num_anm);  // :26-64
  void update_luciferin();                                    // :66-70, batched
  // the two halves of update_luciferin, exposed so several swarms can share one launch
  size_t gather_poses(std::vector<double> &rows, std::vector<uint32_t> &who) const;
  void scatter_scores(const std::vector<uint32_t> &who, const double *scores);
  void movement_phase(StdRng &rng);                           // :72-126
  void find_neighbors();                                      // :85-103, fills Glowworm::neighbors
  void save(uint32_t step, const std::string &output_directory) const;  // :128-167

 private:  // scratch of movement_phase
  std::vector<std::vector<double>> snap_positions, snap_anm_recs, snap_anm_ligs;
  std::vector<Quaternion> snap_rotations;
  std::vector<double> snap_luciferins, flat_xyz, scratch_d2;
  std::vector<unsigned char> scratch_flag;
};

struct GSO {  // src/lib.rs:20-59
  Swarm swarm;
  StdRng rng;
  std::string output_directory;
  // MultiGSO only: this swarm hit what is a panic in the reference (roulette overrun, unwritable output file, ...);
  // it stops there, like the reference's one-process-per-swarm model, and the other swarms go on
  bool failed = false;
  std::string error;

  GSO(const std::vector<std::vector<double>> &positions, uint64_t seed, const Score *scoring, bool use_anm,
      size_t rec_num_anm, size_t lig_num_anm, std::string output_directory);
  void run(uint32_t steps);
};

// Many independent swarms of the SAME complex advanced in lock-step: each step gathers the poses of
// all swarms into one Score::energy_batch call (SURVEY.md §8 e/f: swarms are independent, so a GPU
// owns a set of swarms and scores them together).  Each swarm keeps its own StdRng seeded exactly as
// a stand-alone reference process would be, so every swarm's trajectory equals the single-swarm run.
struct MultiGSO {
  std::vector<GSO> runs;
  const Score *scoring;
  explicit MultiGSO(const Score *s) : scoring(s) {}
  void add(const std::vector<std::vector<double>> &positions, uint64_t seed, bool use_anm, size_t rec_num_anm,
           size_t lig_num_anm, std::string output_directory);
  // host_threads: persistent workers for the per-swarm host phases (gather, luciferin update, movement, save).
  // Swarms never interact and the kernels are batch-invariant, so trajectories do not depend on it.
  // The swarms leapfrog in two sets over Score::energy_batch_begin/_end, so one set's host phases overlap the
  // other set's device scoring.
  void run(uint32_t steps, int host_threads = 1);
  uint64_t energy_calls() const;
  // (index into `runs`, message) of every swarm that stopped early
  std::vector<std::pair<size_t, std::string>> failures() const;

 private:
  void run_lane(const std::vector<size_t> &mine, const Score *sc, uint32_t steps, int host_threads);
};

// The same optimisation with the WHOLE step on the device (include/lightdock_b200.h: ld_gso_*):
// luciferin update, neighbour search, roulette (the swarm's own ChaCha20 stream, evaluated on the device), move_towards
// and the rescoring of the glowworms that moved run without a host round trip per step; the host only fetches the
// swarm state at the steps Swarm::save writes (1 and every 10th, src/lib.rs:51) and writes the files while the device
// runs on.  Same interface as MultiGSO.  Every decision is taken in the reference's f64 operation order; poses can
// differ from a host run in the last bits (CUDA's acos/sin inside slerp), so this path is opt-in (LIGHTDOCK_GSO=device
// in the drivers) and the byte-identical host loop stays the default.  All swarms must have the same glowworm count.
struct DeviceGSO {
  const Score *scoring;
  // state after run(), flat, [swarm][glowworm]: pose rows (pose_len each), luciferin, vision range, scoring, neighbour COUNT
  struct State {
    size_t n_swarms = 0, n_glowworms = 0, pose_len = 0;
    std::vector<double> poses, luciferin, vision, scoring;
    std::vector<int32_t> n_neighbors, failed;
  } state;
  Swarm swarm(size_t s) const;  // swarm s of `state` as a Swarm (what Swarm::save prints)
  explicit DeviceGSO(const Score *s);  // throws unless `s` scores on the GPU (CudaScore)
  void add(const std::vector<std::vector<double>> &positions, uint64_t seed, bool use_anm, size_t rec_num_anm,
           size_t lig_num_anm, std::string output_directory);
  void run(uint32_t steps, int host_threads = 1);
  uint64_t energy_calls() const { return energy_calls_; }
  std::vector<std::pair<size_t, std::string>> failures() const { return failures_; }

 private:
  struct Pending {
    std::vector<std::vector<double>> positions;
    uint64_t seed;
    bool use_anm;
    size_t rec_num_anm, lig_num_anm;
    std::string output_directory;
  };
  std::vector<Pending> pending_;
  uint64_t energy_calls_ = 0;
  std::vector<std::pair<size_t, std::string>> failures_;
};

}  // namespace lightdock

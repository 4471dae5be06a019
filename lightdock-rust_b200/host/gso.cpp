#include "gso.hpp"
#include "log.hpp"

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace lightdock {

Glowworm::Glowworm(uint32_t id_, std::vector<double> translation_, Quaternion rotation_,
                   std::vector<double> rec_nmodes_, std::vector<double> lig_nmodes_, const Score *scoring_function_,
                   bool use_anm_)
    : id(id_), translation(std::move(translation_)), rotation(rotation_), rec_nmodes(std::move(rec_nmodes_)),
      lig_nmodes(std::move(lig_nmodes_)), scoring_function(scoring_function_), use_anm(use_anm_) {}

void Glowworm::write_pose(double *row) const {
  row[0] = translation[0]; row[1] = translation[1]; row[2] = translation[2];
  row[3] = rotation.w; row[4] = rotation.x; row[5] = rotation.y; row[6] = rotation.z;
  size_t k = 7;
  for (double v : rec_nmodes) row[k++] = v;
  for (double v : lig_nmodes) row[k++] = v;
}

void Glowworm::compute_luciferin() {
  const bool scored = needs_scoring();
  double s = scoring;
  if (scored) s = scoring_function->energy(translation, rotation, rec_nmodes, lig_nmodes);
  apply_luciferin(s, scored);
}

void Glowworm::apply_luciferin(double new_scoring, bool scored) {
  if (scored) scoring = new_scoring;
  luciferin = (1.0 - rho) * luciferin + gamma * scoring;
  step += 1;
}

void Glowworm::update_vision_range() {
  vision_range = std::fmin(
      max_vision_range,
      std::fmax(0.0, vision_range + beta * static_cast<double>(static_cast<int32_t>(max_neighbors) -
                                                               static_cast<int32_t>(neighbors.size()))));
}

void Glowworm::compute_probability_moving_toward_neighbor(const std::vector<double> &luciferins) {
  probabilities.clear();
  double total_sum = 0.0;
  for (uint32_t neighbor_id : neighbors) {
    const double difference = luciferins[neighbor_id] - luciferin;
    probabilities.push_back(difference);
    total_sum += difference;
  }
  for (double &p : probabilities) p /= total_sum;
}

uint32_t Glowworm::select_random_neighbor(double random_number) {
  if (neighbors.empty()) return id;
  double sum_probabilities = 0.0;
  size_t i = 0;
  while (sum_probabilities < random_number) {
    // the reference indexes probabilities[i] unchecked-by-logic and panics on overrun
    if (i >= probabilities.size()) throw std::runtime_error("index out of bounds in select_random_neighbor");
    sum_probabilities += probabilities[i];
    i += 1;
  }
  if (i == 0) throw std::runtime_error("index out of bounds in select_random_neighbor");
  return neighbors[i - 1];
}

static void step_towards(std::vector<double> &mine, const std::vector<double> &other, double step_size) {
  std::vector<double> delta;
  delta.reserve(mine.size());
  double cum_norm = 0.0;
  for (size_t i = 0; i < mine.size(); ++i) {
    const double diff = other[i] - mine[i];
    delta.push_back(diff);
    cum_norm += diff * diff;
  }
  const double coef = step_size / std::sqrt(cum_norm);
  for (size_t i = 0; i < mine.size(); ++i) {
    delta[i] *= coef;
    mine[i] += delta[i];
  }
}

void Glowworm::move_towards(uint32_t other_id, const std::vector<double> &other_position,
                            const Quaternion &other_rotation, const std::vector<double> &other_anm_rec,
                            const std::vector<double> &other_anm_lig) {
  moved = id != other_id;
  if (id == other_id) return;
  double delta_x[3] = {other_position[0] - translation[0], other_position[1] - translation[1],
                       other_position[2] - translation[2]};
  const double norm = std::sqrt(delta_x[0] * delta_x[0] + delta_x[1] * delta_x[1] + delta_x[2] * delta_x[2]);
  const double coef = DEFAULT_TRANSLATION_STEP / norm;
  for (int d = 0; d < 3; ++d) {
    delta_x[d] *= coef;
    translation[d] += delta_x[d];
  }
  rotation = rotation.slerp(other_rotation, DEFAULT_ROTATION_STEP);
  if (use_anm && !rec_nmodes.empty()) step_towards(rec_nmodes, other_anm_rec, DEFAULT_NMODES_STEP);
  if (use_anm && !lig_nmodes.empty()) step_towards(lig_nmodes, other_anm_lig, DEFAULT_NMODES_STEP);
}

double distance(const Glowworm &one, const Glowworm &two) {
  const double x1 = one.translation[0], x2 = two.translation[0];
  const double y1 = one.translation[1], y2 = two.translation[1];
  const double z1 = one.translation[2], z2 = two.translation[2];
  return std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
}

void Swarm::add_glowworms(const std::vector<std::vector<double>> &positions, const Score *scoring, bool use_anm,
                          size_t rec_num_anm, size_t lig_num_anm) {
  for (size_t i = 0; i < positions.size(); ++i) {
    const std::vector<double> &position = positions[i];
    if (position.size() < 7) throw std::runtime_error("index out of bounds: start position has fewer than 7 values");
    std::vector<double> translation{position[0], position[1], position[2]};
    const Quaternion rotation(position[3], position[4], position[5], position[6]);
    std::vector<double> rec_nmodes, lig_nmodes;
    if (use_anm && rec_num_anm > 0) {
      if (position.size() < 7 + rec_num_anm) throw std::runtime_error("index out of bounds: receptor ANM extents");
      rec_nmodes.assign(position.begin() + 7, position.begin() + 7 + rec_num_anm);
    }
    if (use_anm && lig_num_anm > 0)
      for (size_t j = 7 + rec_num_anm; j < position.size(); ++j) lig_nmodes.push_back(position[j]);
    glowworms.emplace_back(static_cast<uint32_t>(i), std::move(translation), rotation, std::move(rec_nmodes),
                           std::move(lig_nmodes), scoring, use_anm);
  }
}

size_t Swarm::gather_poses(std::vector<double> &rows, std::vector<uint32_t> &who) const {
  if (glowworms.empty()) return 0;
  const size_t pl = glowworms[0].scoring_function->pose_len();
  size_t n = 0;
  for (const Glowworm &g : glowworms)
    if (g.needs_scoring()) {
      // The reference's energy() reads exactly num_anm extents per partner (src/dfire.rs:290-320) and ignores trailing
      // columns a start file may carry (they end up at the tail of lig_nmodes, src/swarm.rs:47-50); too few is an
      // out-of-bounds panic there.
      const size_t have = 7 + g.rec_nmodes.size() + g.lig_nmodes.size();
      if (have < pl) throw std::runtime_error("index out of bounds: start position has fewer columns than the pose");
      rows.resize(rows.size() + pl);
      if (have == pl) {
        g.write_pose(rows.data() + rows.size() - pl);
      } else {
        std::vector<double> full(have);
        g.write_pose(full.data());
        std::copy(full.begin(), full.begin() + pl, rows.end() - pl);
      }
      who.push_back(g.id);
      ++n;
    }
  return n;
}

void Swarm::scatter_scores(const std::vector<uint32_t> &who, const double *scores) {
  size_t k = 0;
  for (Glowworm &g : glowworms) {
    const bool scored = k < who.size() && who[k] == g.id;
    g.apply_luciferin(scored ? scores[k] : 0.0, scored);
    if (scored) ++k;
  }
  energy_calls += who.size();
}

void Swarm::update_luciferin() {
  if (glowworms.empty()) return;
  std::vector<double> rows, scores;
  std::vector<uint32_t> who;
  const size_t n = gather_poses(rows, who);
  scores.resize(n);
  if (n) glowworms[0].scoring_function->energy_batch(n, rows.data(), scores.data());
  scatter_scores(who, scores.data());
}

// Squared distances from (x1, y1, z1) to every glowworm and the candidate flags `luciferin test && d2 <= hi`,
// element-wise and in the reference's operation order ((x1-x2)*(x1-x2) + (y1-y2)*(y1-y2) + (z1-z2)*(z1-z2), never
// fused: the file is built with -ffp-contract=off), so SIMD lanes compute exactly what the scalar loop would.
// Compiled for AVX2 and baseline x86-64; the loader picks at run time.
__attribute__((target_clones("avx2", "default"), optimize("O3"))) static unsigned neighbour_candidates(
    const double *__restrict__ px, const double *__restrict__ py, const double *__restrict__ pz,
    const double *__restrict__ lum, size_t n, double x1, double y1, double z1, double l1, double hi,
    double *__restrict__ d2s, unsigned char *__restrict__ flag) {
  unsigned any = 0;
  for (size_t j = 0; j < n; ++j) {
    const double dx = x1 - px[j], dy = y1 - py[j], dz = z1 - pz[j];
    const double d2 = dx * dx + dy * dy + dz * dz;
    d2s[j] = d2;
    const unsigned char f = (unsigned char)((l1 < lum[j]) & (d2 <= hi));
    flag[j] = f;
    any |= f;
  }
  return any;
}

void Swarm::find_neighbors() {
  const size_t n = glowworms.size();
  snap_luciferins.resize(n);
  for (size_t i = 0; i < n; ++i) snap_luciferins[i] = glowworms[i].luciferin;
  // Neighbour search, src/swarm.rs:88-103: j is a neighbour of i iff luciferin_i < luciferin_j and
  // distance(i, j) < vision_range_i, distance = sqrt(dx*dx + dy*dy + dz*dz) (src/glowworm.rs:193-202).
  // O(n^2) per swarm and step and the bulk of the host time, so it runs over flat copies of the positions and
  // without the square root: with v2 = v*v and eps = 2^-53, d2 < v2*(1 - 4 eps) implies sqrt(d2) < v and
  // d2 > v2*(1 + 4 eps) implies the opposite whatever the roundings of sqrt and of v*v; only inside that sliver
  // (relative width 1e-15) is the reference's expression evaluated.  The decisions are exactly the reference's.
  flat_xyz.resize(3 * n);
  for (size_t i = 0; i < n; ++i) {
    flat_xyz[i] = glowworms[i].translation[0];
    flat_xyz[n + i] = glowworms[i].translation[1];
    flat_xyz[2 * n + i] = glowworms[i].translation[2];
  }
  const double *px = flat_xyz.data(), *py = px + n, *pz = py + n, *lum = snap_luciferins.data();
  constexpr double kSliver = 4.0 * 1.1102230246251565e-16;
  scratch_d2.resize(n);
  scratch_flag.assign((n + 7) / 8 * 8, 0);  // padded to whole 8-byte words
  double *d2s = scratch_d2.data();
  unsigned char *flag = scratch_flag.data();
  for (size_t i = 0; i < n; ++i) {
    Glowworm &g1 = glowworms[i];
    g1.neighbors.clear();
    const double x1 = px[i], y1 = py[i], z1 = pz[i], l1 = lum[i], v = g1.vision_range;
    const double v2 = v * v, lo = v2 * (1.0 - kSliver), hi = v2 * (1.0 + kSliver);
    // pass 1, branch-free (the luciferin test is a coin flip: a branch on it mispredicts half the time)
    const unsigned any = neighbour_candidates(px, py, pz, lum, n, x1, y1, z1, l1, hi, d2s, flag);
    if (!any) continue;
    // pass 2: the few candidates, in index order (flags scanned eight at a time)
    for (size_t j0 = 0; j0 < n; j0 += 8) {
      uint64_t w;
      std::memcpy(&w, flag + j0, 8);
      if (!w) continue;
      for (size_t j = j0; j < std::min(n, j0 + 8); ++j) {
        if (!flag[j] || j == i) continue;
        if (d2s[j] < lo || std::sqrt(d2s[j]) < v) g1.neighbors.push_back(glowworms[j].id);
      }
    }
  }
}

void Swarm::movement_phase(StdRng &rng) {
  const size_t n = glowworms.size();
  // snapshot of every glowworm's pose before anybody moves (src/swarm.rs:74-86): a glowworm moves towards where
  // its neighbour WAS at the start of the step.  Flat scratch buffers, reused from step to step.
  const size_t nr = n ? glowworms[0].rec_nmodes.size() : 0, nl = n ? glowworms[0].lig_nmodes.size() : 0;
  snap_positions.resize(n);
  snap_rotations.resize(n);
  snap_anm_recs.resize(n);
  snap_anm_ligs.resize(n);
  snap_luciferins.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const Glowworm &g = glowworms[i];
    snap_positions[i] = g.translation;  // element-wise assignment into already-sized vectors: no allocation
    snap_rotations[i] = g.rotation;
    if (nr || !g.rec_nmodes.empty()) snap_anm_recs[i] = g.rec_nmodes;
    if (nl || !g.lig_nmodes.empty()) snap_anm_ligs[i] = g.lig_nmodes;
    snap_luciferins[i] = g.luciferin;
  }
  find_neighbors();
  for (size_t i = 0; i < n; ++i) glowworms[i].compute_probability_moving_toward_neighbor(snap_luciferins);
  for (size_t i = 0; i < n; ++i) {
    Glowworm &g = glowworms[i];
    const uint32_t nid = g.select_random_neighbor(rng.gen_f64());  // always one draw per glowworm (:118)
    g.move_towards(nid, snap_positions[nid], snap_rotations[nid], snap_anm_recs[nid], snap_anm_ligs[nid]);
    g.update_vision_range();
  }
}

// `{:.N}` of an f64 as Rust prints it: NaN, inf, -inf for the non-finite values (C prints nan / -nan).
static void print_f(FILE *f, const char *prefix, double v, int prec) {
  if (std::isnan(v)) std::fprintf(f, "%sNaN", prefix);
  else if (std::isinf(v)) std::fprintf(f, "%s%s", prefix, v > 0 ? "inf" : "-inf");
  else std::fprintf(f, "%s%.*f", prefix, prec, v);
}

void Swarm::save(uint32_t step, const std::string &output_directory) const {
  const std::string path = output_directory + "/gso_" + std::to_string(step) + ".out";
  FILE *f = std::fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("Error saving GSO output: cannot create " + path);
  std::fprintf(f, "#Coordinates  RecID  LigID  Luciferin  Neighbor's number  Vision Range  Scoring\n");
  for (const Glowworm &g : glowworms) {
    const double head[7] = {g.translation[0], g.translation[1], g.translation[2], g.rotation.w, g.rotation.x,
                            g.rotation.y, g.rotation.z};
    for (int k = 0; k < 7; ++k) print_f(f, k == 0 ? "(" : ", ", head[k], 7);
    if (g.use_anm && !g.rec_nmodes.empty())
      for (double v : g.rec_nmodes) print_f(f, ", ", v, 7);
    if (g.use_anm && !g.lig_nmodes.empty())
      for (double v : g.lig_nmodes) print_f(f, ", ", v, 7);
    print_f(f, ")    0    0   ", g.luciferin, 8);
    std::fprintf(f, "  %zu ", g.neighbors.size());
    print_f(f, "", g.vision_range, 3);
    print_f(f, " ", g.scoring, 8);
    std::fputc('\n', f);
  }
  std::fclose(f);
}

GSO::GSO(const std::vector<std::vector<double>> &positions, uint64_t seed, const Score *scoring, bool use_anm,
         size_t rec_num_anm, size_t lig_num_anm, std::string output_directory_)
    : rng(StdRng::seed_from_u64(seed)), output_directory(std::move(output_directory_)) {
  swarm.add_glowworms(positions, scoring, use_anm, rec_num_anm, lig_num_anm);
}

void GSO::run(uint32_t steps) {
  for (uint32_t step = 1; step <= steps; ++step) {
    log_line(LogLevel::Info, "lightdock", "Step " + std::to_string(step));  // src/lib.rs:48
    {
      NvtxRange r("update_luciferin (gather + ld_score_batch + scatter)");
      swarm.update_luciferin();
    }
    {
      NvtxRange r("movement_phase");
      swarm.movement_phase(rng);
    }
    if ((step % 10 == 0 || step == 1) && !output_directory.empty()) {
      NvtxRange r("save");
      swarm.save(step, output_directory);
    }
  }
}

void MultiGSO::add(const std::vector<std::vector<double>> &positions, uint64_t seed, bool use_anm,
                   size_t rec_num_anm, size_t lig_num_anm, std::string output_directory) {
  runs.emplace_back(positions, seed, scoring, use_anm, rec_num_anm, lig_num_anm, std::move(output_directory));
}

uint64_t MultiGSO::energy_calls() const {
  uint64_t n = 0;
  for (const GSO &g : runs) n += g.swarm.energy_calls;
  return n;
}

namespace {
// A fixed set of workers that run fn(i) for i in [0, n) (striped) and wait for each other; created once per
// lane so a GSO step costs two wake-ups instead of two rounds of thread creation.
class Workers {
 public:
  explicit Workers(int n) : n_(std::max(1, n)) {
    for (int t = 1; t < n_; ++t) pool_.emplace_back([this, t] { loop(t); });
  }
  ~Workers() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto &th : pool_) th.join();
  }
  template <typename F>
  void for_each(size_t n, F &&fn) {
    if (n_ == 1 || n < 2) {
      for (size_t i = 0; i < n; ++i) fn(i);
      return;
    }
    std::function<void(size_t)> f = std::ref(fn);
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &f;
      count_ = n;
      pending_ = n_ - 1;
      error_ = nullptr;
      ++epoch_;
    }
    cv_.notify_all();
    work(0);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
    job_ = nullptr;
    if (error_) std::rethrow_exception(error_);
  }

 private:
  void work(int t) {
    try {
      for (size_t i = (size_t)t; i < count_; i += (size_t)n_) (*job_)(i);
    } catch (...) {
      std::lock_guard<std::mutex> lk(m_);
      if (!error_) error_ = std::current_exception();
    }
  }
  void loop(int t) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
      }
      work(t);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  int n_;
  std::vector<std::thread> pool_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(size_t)> *job_ = nullptr;
  size_t count_ = 0;
  int pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
  std::exception_ptr error_;
};
}  // namespace

void MultiGSO::run_lane(const std::vector<size_t> &mine, const Score *sc, uint32_t steps, int host_threads) {
  const size_t ns = mine.size();
  if (ns == 0 || steps == 0) return;
  Workers workers(std::min<int>(std::max(1, host_threads), (int)ns));
  const size_t pl = sc->pose_len();
  // The swarms are split into `nh` sets that leapfrog: while the device scores one set's batch the host runs the
  // other set's luciferin update, movement phase and gather.  Every swarm still sees gather -> score -> update ->
  // move once per step, in that order, so its trajectory is what the one-set loop (and a stand-alone run) gives.
  const int nh = (sc->slots() >= 2 && ns >= 2) ? 2 : 1;
  struct Half {
    size_t lo = 0, hi = 0;  // range of `mine`
    std::vector<double> rows, scores;
    std::vector<size_t> first;
  } half[2];
  for (int k = 0; k < nh; ++k) {
    half[k].lo = ns * k / nh;
    half[k].hi = ns * (k + 1) / nh;
    half[k].first.assign(half[k].hi - half[k].lo + 1, 0);
  }
  std::vector<std::vector<uint32_t>> who(ns);
  std::vector<std::vector<double>> swarm_rows(ns);
  auto gather_and_begin = [&](int k) {
    NvtxRange range(k == 0 ? "gather + begin (set 0)" : "gather + begin (set 1)");
    Half &h = half[k];
    workers.for_each(h.hi - h.lo, [&](size_t i) {  // which glowworms must be rescored, and their pose rows
      const size_t s = h.lo + i;
      who[s].clear();
      swarm_rows[s].clear();
      GSO &r = runs[mine[s]];
      if (r.failed) return;
      try {
        r.swarm.gather_poses(swarm_rows[s], who[s]);
      } catch (const std::exception &e) {  // what would be a panic of that swarm's own process in the reference
        r.failed = true;
        r.error = e.what();
        who[s].clear();
        swarm_rows[s].clear();
      }
    });
    for (size_t i = 0; i < h.hi - h.lo; ++i) h.first[i + 1] = h.first[i] + swarm_rows[h.lo + i].size();
    h.rows.resize(h.first.back());
    workers.for_each(h.hi - h.lo, [&](size_t i) {
      std::copy(swarm_rows[h.lo + i].begin(), swarm_rows[h.lo + i].end(), h.rows.begin() + h.first[i]);
    });
    h.scores.resize(h.rows.size() / pl);
    sc->energy_batch_begin(k, h.scores.size(), h.rows.data());  // ONE batched launch for all swarms of the set
  };
  auto end_and_move = [&](int k, uint32_t step) {
    NvtxRange range(k == 0 ? "end + scatter + movement (set 0)" : "end + scatter + movement (set 1)");
    Half &h = half[k];
    {
      NvtxRange wait("ld_score_batch_end (device wait)");
      sc->energy_batch_end(k, h.scores.data());
    }
    workers.for_each(h.hi - h.lo, [&](size_t i) {
      GSO &r = runs[mine[h.lo + i]];
      if (r.failed) return;
      // One process per swarm in the reference (example/1czy/execution.sh:21-25): a panic there ends that swarm
      // only.  Same here: the swarm is marked failed and dropped from the following steps, the others go on.
      try {
        r.swarm.scatter_scores(who[h.lo + i], h.scores.data() + h.first[i] / pl);
        r.swarm.movement_phase(r.rng);
        if ((step % 10 == 0 || step == 1) && !r.output_directory.empty()) r.swarm.save(step, r.output_directory);
      } catch (const std::exception &e) {
        r.failed = true;
        r.error = std::string("step ") + std::to_string(step) + ": " + e.what();
      }
    });
  };
  // A failure of the scoring call itself (CUDA error, bad argument) is not a per-swarm event: it ends the lane, but
  // not before every slot with a batch in flight has been collected, so the handle stays usable.
  bool in_flight[2] = {false, false};
  try {
    for (int k = 0; k < nh; ++k) { gather_and_begin(k); in_flight[k] = true; }
    for (uint32_t step = 1; step <= steps; ++step)
      for (int k = 0; k < nh; ++k) {
        if (k == 0) log_line(LogLevel::Info, "lightdock", "Step " + std::to_string(step));  // src/lib.rs:48
        in_flight[k] = false;
        end_and_move(k, step);
        if (step < steps) { gather_and_begin(k); in_flight[k] = true; }
      }
  } catch (...) {
    for (int k = 0; k < nh; ++k)
      if (in_flight[k]) {
        try {
          half[k].scores.resize(half[k].rows.size() / pl);
          sc->energy_batch_end(k, half[k].scores.data());
        } catch (...) {
        }
      }
    throw;
  }
}

std::vector<std::pair<size_t, std::string>> MultiGSO::failures() const {
  std::vector<std::pair<size_t, std::string>> out;
  for (size_t s = 0; s < runs.size(); ++s)
    if (runs[s].failed) out.emplace_back(s, runs[s].error);
  return out;
}

void MultiGSO::run(uint32_t steps, int host_threads) {
  std::vector<size_t> all(runs.size());
  for (size_t s = 0; s < all.size(); ++s) all[s] = s;
  run_lane(all, scoring, steps, std::max(1, host_threads));
}

// ---------------------------------------------------------------------------------------------------------------
DeviceGSO::DeviceGSO(const Score *s) : scoring(s) {
  if (!dynamic_cast<const CudaScore *>(s))
    throw std::runtime_error("DeviceGSO needs a scoring object that lives on the GPU (CudaScore)");
}

void DeviceGSO::add(const std::vector<std::vector<double>> &positions, uint64_t seed, bool use_anm, size_t rec_num_anm,
                    size_t lig_num_anm, std::string output_directory) {
  pending_.push_back(Pending{positions, seed, use_anm, rec_num_anm, lig_num_anm, std::move(output_directory)});
}

namespace {
struct GsoGuard {  // ld_gso_destroy on every exit path
  ld_gso *g = nullptr;
  ~GsoGuard() { ld_gso_destroy(g); }
};
}  // namespace

// One swarm of a flat state as Swarm objects (neighbour lists hold the right COUNT of placeholder ids: the count is
// what Swarm::save prints)
template <typename P>
static Swarm swarm_of(const DeviceGSO::State &b, size_t s, const P &p, const Score *scoring) {
  const size_t n = b.n_glowworms, pl = b.pose_len;
  Swarm sw;
  std::vector<std::vector<double>> pos(n);
  for (size_t i = 0; i < n; ++i) pos[i].assign(b.poses.begin() + (s * n + i) * pl, b.poses.begin() + (s * n + i + 1) * pl);
  sw.add_glowworms(pos, scoring, p.use_anm, p.rec_num_anm, p.lig_num_anm);
  for (size_t i = 0; i < n; ++i) {
    Glowworm &g = sw.glowworms[i];
    g.luciferin = b.luciferin[s * n + i];
    g.vision_range = b.vision[s * n + i];
    g.scoring = b.scoring[s * n + i];
    g.neighbors.assign((size_t)b.n_neighbors[s * n + i], 0u);
  }
  return sw;
}

Swarm DeviceGSO::swarm(size_t s) const { return swarm_of(state, s, pending_.at(s), scoring); }

void DeviceGSO::run(uint32_t steps, int host_threads) {
  const size_t S = pending_.size();
  state = State{};
  failures_.clear();
  energy_calls_ = 0;
  if (S == 0) return;
  const CudaScore *cs = dynamic_cast<const CudaScore *>(scoring);
  const size_t pl = scoring->pose_len();
  const size_t n = pending_[0].positions.size();
  if (n == 0) return;
  // pose rows as Swarm::add_glowworms + gather_poses make them (src/swarm.rs:26-64): translation, rotation, then the
  // first pose_len - 7 extents; a start row with fewer columns is an out-of-bounds panic in the reference
  std::vector<double> rows(S * n * pl);
  std::vector<uint64_t> seeds(S);
  for (size_t s = 0; s < S; ++s) {
    const Pending &p = pending_[s];
    if (p.positions.size() != n) throw std::runtime_error("DeviceGSO: every swarm must have the same number of glowworms");
    const size_t want = 7 + (p.use_anm ? p.rec_num_anm + p.lig_num_anm : 0);
    if (want != pl) throw std::runtime_error("DeviceGSO: swarm set-up does not match the scoring object's pose length");
    for (size_t i = 0; i < n; ++i) {
      if (p.positions[i].size() < pl)
        throw std::runtime_error("index out of bounds: start position has fewer columns than the pose");
      std::copy(p.positions[i].begin(), p.positions[i].begin() + pl, rows.begin() + (s * n + i) * pl);
    }
    seeds[s] = p.seed;
  }
  GsoGuard guard;
  if (ld_gso_create(cs->handle(), (int32_t)S, (int32_t)n, rows.data(), seeds.data(), &guard.g) != LD_OK)
    throw std::runtime_error(std::string("ld_gso_create: ") + ld_last_error());

  // swarm state on the host, double-buffered: one copy is being written to disk while the next is fetched
  State snap[2];
  for (State &b : snap) {
    b.n_swarms = S; b.n_glowworms = n; b.pose_len = pl;
    b.poses.resize(S * n * pl); b.luciferin.resize(S * n); b.vision.resize(S * n); b.scoring.resize(S * n);
    b.n_neighbors.resize(S * n); b.failed.resize(S);
  }
  auto materialise = [&](const State &b, size_t s) { return swarm_of(b, s, pending_[s], scoring); };
  bool any_output = false;
  for (const Pending &p : pending_) any_output = any_output || !p.output_directory.empty();
  std::vector<std::string> save_error(S);
  std::vector<int32_t> failed_at(S, 0);
  std::thread writer;
  auto join_writer = [&] { if (writer.joinable()) writer.join(); };
  struct Joiner { std::function<void()> f; ~Joiner() { f(); } } joiner{join_writer};
  uint32_t done = 0;
  int which = 0;
  const int nt = std::max(1, std::min<int>(host_threads, (int)S));
  while (done < steps) {
    // run up to the next step whose state is saved: 1, 10, 20, ... (src/lib.rs:51)
    const uint32_t next = done == 0 ? 1 : std::min<uint32_t>(steps, (done / 10 + 1) * 10);
    {
      NvtxRange r("ld_gso_run (device-resident steps)");
      for (uint32_t k = done; k < next; ++k) log_line(LogLevel::Info, "lightdock", "Step " + std::to_string(k + 1));
      if (ld_gso_run(guard.g, (int32_t)(next - done)) != LD_OK)
        throw std::runtime_error(std::string("ld_gso_run: ") + ld_last_error());
    }
    done = next;
    const bool save_step = done % 10 == 0 || done == 1;
    if ((!save_step || !any_output) && done < steps) continue;  // nothing to write: the state is fetched at the end only
    join_writer();  // the previous snapshot's files are on disk; its buffer is two saves old after the flip
    State &b = snap[which];
    which ^= 1;
    if (ld_gso_state(guard.g, b.poses.data(), b.luciferin.data(), b.vision.data(), b.scoring.data(), b.n_neighbors.data(),
                     b.failed.data()) != LD_OK)
      throw std::runtime_error(std::string("ld_gso_state: ") + ld_last_error());
    for (size_t s = 0; s < S; ++s)
      if (b.failed[s] && !failed_at[s]) failed_at[s] = b.failed[s];
    if (!save_step || !any_output) break;  // (only reached at the last step)
    const uint32_t step = done;
    writer = std::thread([&, step, nt, bp = &b] {
      NvtxRange r("save (overlaps the next device steps)");
      std::vector<std::thread> pool;
      for (int t = 0; t < nt; ++t)
        pool.emplace_back([&, t] {
          for (size_t s = (size_t)t; s < S; s += (size_t)nt) {
            if (bp->failed[s] || pending_[s].output_directory.empty() || !save_error[s].empty()) continue;
            try {
              materialise(*bp, s).save(step, pending_[s].output_directory);
            } catch (const std::exception &e) {
              save_error[s] = std::string("step ") + std::to_string(step) + ": " + e.what();
            }
          }
        });
      for (auto &th : pool) th.join();
    });
  }
  join_writer();
  if (steps == 0) {  // nothing ran: the state is the start state
    if (ld_gso_state(guard.g, snap[0].poses.data(), snap[0].luciferin.data(), snap[0].vision.data(), snap[0].scoring.data(),
                     snap[0].n_neighbors.data(), snap[0].failed.data()) != LD_OK)
      throw std::runtime_error(std::string("ld_gso_state: ") + ld_last_error());
    which = 1;
  }
  state = std::move(snap[which ^ 1]);
  energy_calls_ = (uint64_t)ld_gso_energy_calls(guard.g);
  for (size_t s = 0; s < S; ++s) {
    if (failed_at[s])  // src/glowworm.rs:114-126: the reference indexes past its probabilities and panics
      failures_.emplace_back(s, "step " + std::to_string(failed_at[s]) + ": index out of bounds in select_random_neighbor");
    else if (!save_error[s].empty())
      failures_.emplace_back(s, save_error[s]);
  }
}

}  // namespace lightdock

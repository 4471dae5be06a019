// scoring.hpp — host-side mirror of the reference's scoring plug-in interface for the hot path:
// `trait Score` (src/scoring.rs:11-19) and its three implementors DFIRE (src/dfire.rs:194-262),
// DNA (src/dna.rs:367-408) and PYDOCK (src/pydock.rs:384-423).  Model building (atom typing,
// restraint / membrane indexing) happens here on the host exactly as in the reference constructors;
// `energy` forwards to the CUDA library through the C ABI (include/lightdock_b200.h).  There is no
// CPU implementation of the pair loop in this layer.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/lightdock_b200.h"
#include "pdb.hpp"
#include "quaternion.hpp"

namespace lightdock {

constexpr double INTERFACE_CUTOFF = 3.9;          // src/constants.rs:14
constexpr double MEMBRANE_PENALTY_SCORE = 999.0;  // src/constants.rs:21

enum class Method { DFIRE, DNA, PYDOCK };  // src/scoring.rs:5-9
const char *method_name(Method m);           // Debug formatting of the enum

// The plug-in boundary.  `energy` keeps the reference signature; `energy_batch` is the one addition:
// all poses of a step in one call (what Swarm::update_luciferin uses).
class Score {
 public:
  virtual ~Score() = default;
  virtual double energy(const std::vector<double> &translation, const Quaternion &rotation,
                        const std::vector<double> &rec_nmodes, const std::vector<double> &lig_nmodes) const = 0;
  // poses: n rows of pose_len() doubles (tx,ty,tz,qw,qx,qy,qz, rec extents, lig extents)
  virtual void energy_batch(size_t n, const double *poses, double *energies) const = 0;
  virtual size_t pose_len() const = 0;
  // energy_batch in two halves, so a caller can keep `slots()` batches in flight: begin enqueues the batch and
  // returns (the pose buffer may be reused at once), end waits for it and writes the energies.  The default is
  // the synchronous call at `end`; CudaScore maps them to ld_score_batch_begin / _end.
  virtual int slots() const { return 1; }
  virtual void energy_batch_begin(int slot, size_t n, const double *poses) const {
    (void)slot;
    pending_.assign(poses, poses + n * pose_len());
    pending_n_ = n;
  }
  virtual void energy_batch_end(int slot, double *energies) const {
    (void)slot;
    if (pending_n_) energy_batch(pending_n_, pending_.data(), energies);
    pending_n_ = 0;
  }

 private:
  mutable std::vector<double> pending_;
  mutable size_t pending_n_ = 0;
};

// Numeric content of DFIREDockingModel / DNADockingModel / PYDOCKDockingModel.
struct DockingModel {
  std::vector<int> atoms;  // DFIRE atom types (src/dfire.rs:104); empty for DNA/PYDOCK
  std::vector<double> coordinates;  // [n][3]
  std::vector<int> membrane;
  std::map<std::string, std::vector<int>> active_restraints, passive_restraints;
  size_t num_anm = 0;
  std::vector<double> nmodes;
  std::vector<double> vdw_radii, vdw_charges, ele_charges;  // DNA/PYDOCK
  size_t num_atoms() const { return coordinates.size() / 3; }

  // DFIREDockingModel::new (src/dfire.rs:115-190), DNADockingModel::new (src/dna.rs:249-364),
  // PYDOCKDockingModel::new (src/pydock.rs:253-382).  Throws std::runtime_error where the reference panics.
  static DockingModel build(Method method, const PDB &structure, const std::vector<std::string> &active_restraints,
                            const std::vector<std::string> &passive_restraints, const std::vector<double> &nmodes,
                            size_t num_anm);
};

// DFIRE::load_potentials (src/dfire.rs:236-257): $LIGHTDOCK_DATA or "data", file DCparams, first 169*169*20 lines.
std::vector<double> load_potentials();

// A Score whose energy() runs on the GPU.  DFIRE, DNA and PYDOCK below are thin named constructors,
// like the reference's `DFIRE::new(..) -> Box<dyn Score>`.
class CudaScore : public Score {
 public:
  CudaScore(Method method, DockingModel receptor, DockingModel ligand, bool use_anm, std::vector<double> potential,
            int device);
  ~CudaScore() override;
  CudaScore(const CudaScore &) = delete;
  CudaScore &operator=(const CudaScore &) = delete;

  double energy(const std::vector<double> &translation, const Quaternion &rotation,
                const std::vector<double> &rec_nmodes, const std::vector<double> &lig_nmodes) const override;
  void energy_batch(size_t n, const double *poses, double *energies) const override;
  size_t pose_len() const override { return pose_len_; }
  int slots() const override { return LD_SLOTS; }
  void energy_batch_begin(int slot, size_t n, const double *poses) const override;
  void energy_batch_end(int slot, double *energies) const override;

  const DockingModel &receptor() const { return receptor_; }
  const DockingModel &ligand() const { return ligand_; }
  ld_handle *handle() const { return handle_; }
  Method method() const { return method_; }

 private:
  Method method_;
  DockingModel receptor_, ligand_;
  bool use_anm_;
  std::vector<double> potential_;
  ld_handle *handle_ = nullptr;
  size_t pose_len_ = 7;
};

struct DFIRE {
  static std::unique_ptr<Score> create(const PDB &receptor, const std::vector<std::string> &rec_active_restraints,
                                       const std::vector<std::string> &rec_passive_restraints,
                                       const std::vector<double> &rec_nmodes, size_t rec_num_anm, const PDB &ligand,
                                       const std::vector<std::string> &lig_active_restraints,
                                       const std::vector<std::string> &lig_passive_restraints,
                                       const std::vector<double> &lig_nmodes, size_t lig_num_anm, bool use_anm,
                                       int device = 0);
};
struct DNA {
  static std::unique_ptr<Score> create(const PDB &receptor, const std::vector<std::string> &rec_active_restraints,
                                       const std::vector<std::string> &rec_passive_restraints,
                                       const std::vector<double> &rec_nmodes, size_t rec_num_anm, const PDB &ligand,
                                       const std::vector<std::string> &lig_active_restraints,
                                       const std::vector<std::string> &lig_passive_restraints,
                                       const std::vector<double> &lig_nmodes, size_t lig_num_anm, bool use_anm,
                                       int device = 0);
};
struct PYDOCK {
  static std::unique_ptr<Score> create(const PDB &receptor, const std::vector<std::string> &rec_active_restraints,
                                       const std::vector<std::string> &rec_passive_restraints,
                                       const std::vector<double> &rec_nmodes, size_t rec_num_anm, const PDB &ligand,
                                       const std::vector<std::string> &lig_active_restraints,
                                       const std::vector<std::string> &lig_passive_restraints,
                                       const std::vector<double> &lig_nmodes, size_t lig_num_anm, bool use_anm,
                                       int device = 0);
};

}  // namespace lightdock

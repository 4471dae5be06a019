/*
 * lightdock_b200.h — C ABI of the B200-native (sm_100a) LightDock scoring library.
 *
 * This is the drop-in boundary for ONE path of lightdock-rust v0.3.2: evaluating the DFIRE and
 * DNA/pyDock energy of every glowworm pose of a swarm at each GSO step.  The reference has no FFI
 * today; its plug-in boundary is the Rust trait
 *
 *     trait Score { fn energy(&self, translation:&[f64], rotation:&Quaternion,
 *                             rec_nmodes:&[f64], lig_nmodes:&[f64]) -> f64 }      src/scoring.rs:11-19
 *
 * implemented by DFIRE (src/dfire.rs:265-362), DNA (src/dna.rs:411-529) and PYDOCK
 * (src/pydock.rs:426-544) and called once per glowworm per step from
 * Glowworm::compute_luciferin (src/glowworm.rs:61-72) <- Swarm::update_luciferin
 * (src/swarm.rs:66-70).  The entry points below are what a Rust `extern "C"` block (or cgo/ctypes)
 * binds to replace those three `energy` bodies; INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * LD_E* code on failure, with a thread-local message in ld_last_error().  Nothing unwinds across
 * the ABI.  There is NO CPU fallback: without a CUDA device ld_create fails.
 * A handle is used by one host thread at a time; different handles may be used concurrently.
 */
#ifndef LIGHTDOCK_B200_H
#define LIGHTDOCK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LD_OK 0
#define LD_EINVAL (-1)   /* bad argument / malformed descriptor (the reference would panic) */
#define LD_ECUDA (-2)    /* CUDA runtime error (message carries cudaGetErrorString)          */
#define LD_ENOMEM (-3)   /* host or device allocation failed                                */
#define LD_ELIMIT (-4)   /* complex too large for the shared-memory staging of this build   */

/* enum Method, src/scoring.rs:5-9.  PYDOCK shares the DNA kernel: only the host-side atom
 * parameterisation differs (src/pydock.rs:332-345). */
#define LD_METHOD_DFIRE 0
#define LD_METHOD_DNA 1
#define LD_METHOD_PYDOCK 2

/* 169*169*20 entries are read from DCparams (src/dfire.rs:254). */
#define LD_DFIRE_TABLE_LEN 571220

/* One docking partner: the numeric content of DFIREDockingModel (src/dfire.rs:103-111) /
 * DNADockingModel (src/dna.rs:235-246).  All arrays are caller-owned and copied by ld_create. */
typedef struct ld_molecule_desc {
  int32_t n_atoms;
  const double *coords;        /* [n_atoms][3] AoS f64, atom order of the model                 */
  const int32_t *dfire_type;   /* [n_atoms] DFIRE atom type 0..168 (src/dfire.rs:177-183); DFIRE */
  const double *ele_charge;    /* [n_atoms] src/dna.rs:245 ; DNA/PYDOCK                          */
  const double *vdw_energy;    /* [n_atoms] "vdw_charges" src/dna.rs:244 ; DNA/PYDOCK            */
  const double *vdw_radius;    /* [n_atoms] src/dna.rs:243 ; DNA/PYDOCK                          */
  int32_t n_modes;             /* num_anm (setup.anm_rec / anm_lig)                              */
  const double *modes;         /* [n_modes][n_atoms][3] flat f64 (src/dfire.rs:292-293) or NULL  */
  int32_t n_restraints;        /* ACTIVE restraint residues found in the structure               */
  const int32_t *rst_offsets;  /* [n_restraints+1] CSR offsets into rst_atoms                    */
  const int32_t *rst_atoms;    /* atom indices of each restraint residue (src/dfire.rs:151-162)  */
  int32_t n_membrane;          /* MMB.BJ bead count (src/dfire.rs:146-149)                       */
  const int32_t *membrane;     /* [n_membrane] atom indices                                      */
} ld_molecule_desc;

typedef struct ld_complex_desc {
  int32_t method;              /* LD_METHOD_*                                                     */
  int32_t use_anm;             /* setup.use_anm: pose rows carry n_modes extents per partner      */
  ld_molecule_desc receptor;
  ld_molecule_desc ligand;
  const double *dfire_potential; /* [LD_DFIRE_TABLE_LEN] f64, DFIRE only                          */
  int32_t device;              /* CUDA device ordinal                                             */
  int32_t reserved;
} ld_complex_desc;

/* Quantities the reference computes inside energy() but does not return; the parity tests compare
 * the integer ones bit-exact against the oracle. */
typedef struct ld_pose_detail {
  double raw_sum;              /* DFIRE: sum of table values (src/dfire.rs:338); DNA: total_elec  */
  double raw_sum2;             /* DNA: total_vdw (src/dna.rs:503)                                 */
  int64_t n_in_cutoff;         /* DFIRE: dist<=225 (src/dfire.rs:334); DNA: d2<=900               */
  int64_t n_in_cutoff2;        /* DNA: d2<=100 (src/dna.rs:494)                                   */
  int64_t n_interface_pairs;   /* pairs passing the interface test                                */
  int64_t bin_hist[21];        /* DFIRE: histogram of dfire_bin (src/dfire.rs:337)                */
  int32_t rec_rst_hit;         /* satisfied receptor restraint residues (src/scoring.rs:21-36)    */
  int32_t lig_rst_hit;
  int32_t membrane_hit;        /* beads at the interface (src/scoring.rs:38-47)                   */
  int32_t reserved;
  int64_t n_pairs_tested;      /* atom-pair distance tests the GPU executed (after tile culling)   */
  int64_t n_exact_fallback;    /* DFIRE: pairs too close to a decision threshold for the FP32
                                  classification, re-evaluated in exact FP64                       */
} ld_pose_detail;

/* Work counters of the last ld_score_batch* call (what the GPU actually executed). */
typedef struct ld_batch_stats {
  int64_t n_poses;
  int64_t pair_evals_bruteforce; /* n_poses * n_rec * n_lig: the reference's loop count           */
  int32_t kernel_launches;       /* CUDA kernels launched by the call                             */
  int32_t rec_splits;            /* receptor tile ranges per pose (CTAs per pose)                 */
  double device_ms;              /* device time of the call (CUDA events); host-buffer calls only */
  /* per-kernel device time of the last call, filled only while profiling is on (ld_set_profiling)
   * and read back by ld_get_stats, which synchronises the stream the call used */
  double transform_ms, pair_ms, finalize_ms;
  int32_t path;                  /* LD_PATH_GENERIC or LD_PATH_RIGID: the pair kernel the call used */
  int32_t pair_launches;         /* launches of that pair kernel                                  */
} ld_batch_stats;

typedef struct ld_handle ld_handle;

/* Builds the device-resident scoring object.  Replaces DFIRE::new / DNA::new / PYDOCK::new
 * (src/dfire.rs:201-234, src/dna.rs:375-407, src/pydock.rs:391-422) after the host has typed the
 * atoms. */
int ld_create(const ld_complex_desc *desc, ld_handle **out);
int ld_destroy(ld_handle *h);

/* Doubles per pose row: 7 (+ receptor.n_modes + ligand.n_modes when use_anm).
 * Row = tx,ty,tz, qw,qx,qy,qz, receptor extents..., ligand extents...  (src/swarm.rs:33-51). */
int ld_pose_len(const ld_handle *h);

/* Score::energy for n_poses poses in ONE batched launch sequence; host buffers, synchronous.
 * Replaces the loop of Swarm::update_luciferin over Score::energy (src/swarm.rs:66-70). */
int ld_score_batch(ld_handle *h, int64_t n_poses, const double *poses, double *energies);

/* The same call split in two so a caller can keep LD_SLOTS batches in flight: _begin copies the pose rows into
 * the slot's pinned staging buffer and enqueues the copies and kernels on the slot's own stream, then returns;
 * _end waits for that slot and writes the energies.  The pose buffer may be reused as soon as _begin returns.
 * Typical use (host/gso.cpp, MultiGSO): while the device scores one half of the swarms the host runs the GSO
 * movement phase of the other half.  One batch per slot at a time; a slot with a batch pending must be ended before
 * it is begun again, and before any other scoring call on the handle uses slot 0. */
#define LD_SLOTS 2
int ld_score_batch_begin(ld_handle *h, int32_t slot, int64_t n_poses, const double *poses);
int ld_score_batch_end(ld_handle *h, int32_t slot, double *energies);

/* Same, with device-resident poses/energies on a caller-provided CUDA stream (cudaStream_t passed
 * as void*; NULL = the handle's own stream).  Asynchronous with respect to the host.  The call works in the
 * handle's slot-0 work buffers; the library orders successive users of those buffers itself (an event recorded
 * after the call's last kernel is waited on by the next call, on whatever stream that one launches), so calls on
 * different streams serialise on the device instead of racing.  The caller still owns the ordering of ITS buffers
 * (d_poses must be ready on `stream`; d_energies is complete when `stream` reaches the end of the call). */
int ld_score_batch_device(ld_handle *h, int64_t n_poses, const double *d_poses, double *d_energies,
                          void *stream);

/* ---- Device-resident GSO -------------------------------------------------------------------
 * The whole optimisation loop of GSO::run (src/lib.rs:46-58) on the device, for n_swarms independent swarms of the
 * handle's complex advanced in lock-step: Swarm::update_luciferin (src/swarm.rs:66-70, src/glowworm.rs:61-72),
 * Swarm::movement_phase (src/swarm.rs:72-126: snapshot, neighbour search, probabilities, one StdRng draw per glowworm,
 * Glowworm::move_towards incl. Quaternion::slerp src/qt.rs:67-91, update_vision_range) and the scoring pass of the
 * glowworms that moved -- no host round trip per step.  Each swarm consumes the ChaCha20 stream of
 * `StdRng::seed_from_u64(seeds[s])` exactly as a stand-alone reference process would (the stream is evaluated on the
 * device).  All arithmetic is f64 in the reference's operation order; the only place where the device can round
 * differently from a host run is acos/sin inside slerp (CUDA's libm vs the host's), i.e. poses agree to ~1e-15
 * relative per step, far inside the 1e-6 contract, instead of bit for bit -- which is why the drop-in CLI keeps the
 * host loop by default and takes this one on request (LIGHTDOCK_GSO=device).
 * A swarm that hits what is a panic in the reference (roulette overrun / draw of exactly 0, src/glowworm.rs:114-126)
 * stops at that step, as its own process would; the other swarms go on (failed_step reports it).
 *
 * positions: [n_swarms][n_glowworms][ld_pose_len(h)] start poses (src/swarm.rs:26-64); seeds: [n_swarms].
 * One ld_gso per handle at a time; between ld_gso_create and ld_gso_destroy the handle's other scoring calls may be
 * used only while no ld_gso_run is executing (they share the handle's slot-0 work buffers and stream).  Destroy the
 * ld_gso before its handle. */
typedef struct ld_gso ld_gso;
#define LD_GSO_MAX_GLOWWORMS 1024
int ld_gso_create(ld_handle *h, int32_t n_swarms, int32_t n_glowworms, const double *positions, const uint64_t *seeds,
                  ld_gso **out);
/* Advances every swarm by n_steps GSO steps (step numbering continues from the previous call) and waits. */
int ld_gso_run(ld_gso *g, int32_t n_steps);
/* The swarm state as Swarm::save prints it (src/swarm.rs:128-167), after the steps run so far: poses
 * [S][n][pose_len], luciferin / vision_range / scoring [S][n], n_neighbors [S][n], failed_step [S] (0 = running).
 * Any pointer may be NULL. */
int ld_gso_state(ld_gso *g, double *poses, double *luciferin, double *vision_range, double *scoring,
                 int32_t *n_neighbors, int32_t *failed_step);
/* Steps run so far / Score::energy evaluations so far (glowworms rescored: `moved || step == 0`, src/glowworm.rs:62). */
int32_t ld_gso_steps(const ld_gso *g);
int64_t ld_gso_energy_calls(const ld_gso *g);
int ld_gso_destroy(ld_gso *g);

/* Same as ld_score_batch plus the per-pose diagnostics.  iface_rec [n_poses][n_rec] and iface_lig
 * [n_poses][n_lig] (0/1 bytes, ORIGINAL atom order; src/dfire.rs:322-323) may be NULL. */
int ld_score_batch_detail(ld_handle *h, int64_t n_poses, const double *poses, double *energies,
                          ld_pose_detail *detail, uint8_t *iface_rec, uint8_t *iface_lig);

/* The pose transform alone (src/dfire.rs:282-320): coordinates of both partners as the pair loop
 * sees them, ORIGINAL atom order, [n_poses][n_atoms][3].  Either output may be NULL. */
int ld_transform_batch(ld_handle *h, int64_t n_poses, const double *poses, double *rec_coords,
                       double *lig_coords);

/* Counters of the last scoring call on the handle (whatever slot it used) / of the last call on one slot. */
int ld_get_stats(ld_handle *h, ld_batch_stats *out);
int ld_get_stats_slot(ld_handle *h, int32_t slot, ld_batch_stats *out);

/* Tuning knob for benchmarks/tests: force the number of receptor splits (0 = automatic). */
int ld_set_rec_splits(ld_handle *h, int32_t splits);

/* Pair-kernel selection.  AUTO = RIGID whenever it applies (DFIRE, no ligand ANM modes, ligand and
 * table rows fit in shared memory), else GENERIC.  RIGID moves each receptor atom into the ligand's
 * frame and looks candidates up in ligand-frame cell lists built once by ld_create; GENERIC moves the
 * ligand per pose and culls with bounding spheres.  Both give the same discrete outputs; energies may
 * differ in the last bits (summation order).  Forcing RIGID on a complex it cannot take is LD_EINVAL. */
#define LD_PATH_AUTO 0
#define LD_PATH_GENERIC 1
#define LD_PATH_RIGID 2
int ld_set_path(ld_handle *h, int32_t path);
/* One-line description of the rigid-path structures (groups, cells, list sizes) or why it is off. */
const char *ld_path_info(const ld_handle *h);

/* Brackets every kernel launch with CUDA events on the launching stream (bench.py's roofline leg). */
int ld_set_profiling(ld_handle *h, int32_t on);

/* On-box micro-benchmarks for the roofline denominators SURVEY.md §8(d) asks for (MEASURED_PEAKS.json
 * has no FP64 / L2-gather figure): sustained non-fused FP64 add+mul rate in TFLOP/s, FP32 FMA-free
 * rate in TFLOP/s, and random 8-byte gather rate from a DFIRE-table-sized (4.57 MB) L2-resident
 * window in G loads/s. */
int ld_probe_peaks(int32_t device, double *fp64_nonfma_tflops, double *fp32_nonfma_tflops,
                   double *l2_gather_gloads);

/* Process-wide tuning defaults for handles created afterwards (benchmark / experiment aid; the library never reads
 * the environment): "rigid_rows" 1..8 table rows per receptor group at most (the rigid instance takes up to 4, the FLEX instance up to 8), "cell_size" ligand-frame cell edge in A
 * (0.5..8; 0, the default: chosen per complex so that grid + lists stay well inside L2), "units_per_sm" rigid-kernel work units per SM, "flex_min_warps" warps per CTA the FLEX instance wants before a receptor
 * group may span one more table row, "default_path" LD_PATH_AUTO | LD_PATH_GENERIC, "flex" 0 keeps
 * ligands with ANM modes on the generic kernel, "cells_on_host" 1 builds the ligand-frame cell lists with host threads
 * (the cross-check of the device builder), "compact_tiles" 0 keeps the plain bisection order of the atoms (the tiles of 8 / 32
 * atoms are otherwise made more compact by a capacity-constrained k-means: fewer executed pair tests), "dna_fused" 0 sends
 * DNA/pyDock poses through the separate transform kernel and per-pose coordinate blocks (the cross-check of the pair kernel's
 * own pose transform). */
int ld_set_option(const char *key, double value);

/* Creates the CUDA context of `device` (cudaSetDevice + first runtime call).  Optional: ld_create does it too; a
 * driver can call this from a helper thread at start-up so the 0.2-0.4 s of context creation overlap its file parsing. */
int ld_init_device(int32_t device);
/* Milliseconds ld_create spent: [0] CUDA context, [1] sorting + uploading the complex, [2] receptor groups of the
 * ligand-frame path, [3] ligand-frame cell lists ([3] is refreshed by every FLEX rebuild). */
int ld_get_create_ms(const ld_handle *h, double *out4);

/* Number of CUDA devices visible to the process (0 if none / no driver): the multi-swarm driver shards swarms
 * over them (swarm s -> device s mod count). */
int ld_device_count(void);

const char *ld_last_error(void);
const char *ld_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTDOCK_B200_H */

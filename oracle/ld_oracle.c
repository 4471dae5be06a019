/*
 * ld_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the scoring hot path of lightdock-rust v0.3.2 and of the host-side GSO
 * loop that calls it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library; the product (lightdock-rust_b200/) never does.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Operation order is the reference's (Rust evaluates a+b+c as (a+b)+c and never fuses mul+add);
 * build with -O2 -ffp-contract=off (oracle/Makefile).
 *
 * Pinning (tests/test_oracle_golden.py): DNA/pyDock known answer -364.88126358158974
 * (src/dna.rs:571, src/pydock.rs:586), 200 pose->energy pairs of example/1azp/swarm_0/gso_1.out,
 * the 100-step 1azp trajectory gso_{1,10..100}.out, Quaternion::rotate (src/qt.rs:360-369) and the
 * StdRng known answer (src/qt.rs:451-462).  DFIRE: the algorithm is restated but the reference's
 * DFIRE numbers (src/dfire.rs:415 and the DFIRE gso files) need data/DCparams, which is absent
 * from this container => DFIRE numeric parity is UNPINNED here (pinned automatically when a real
 * DCparams is supplied through LIGHTDOCK_DATA).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* constants: src/constants.rs:14-15,21 ; src/dna.rs:15-25                                     */
static const double INTERFACE_CUTOFF = 3.9;
static const double MEMBRANE_PENALTY_SCORE = 999.0;
static const double DNA_EPSILON = 4.0, DNA_FACTOR = 332.0;
static const double MAX_ES_CUTOFF = 1.0, MIN_ES_CUTOFF = -1.0, VDW_CUTOFF = 1.0;
static const double ELEC_DIST_CUTOFF = 30.0, VDW_DIST_CUTOFF = 10.0;

/* src/dfire.rs:49-53 */
static const int DIST_TO_BINS[51] = {1,  1,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 14,
                                     15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22, 23,
                                     23, 24, 24, 25, 25, 26, 26, 27, 27, 28, 28, 29, 29, 30, 30, 31, 32};

/* ------------------------------------------------------------------------------------------ */
/* One docking partner + the scoring parameters; filled by the Python side (oracle/oracle.py).  */
typedef struct {
  int32_t n_atoms;
  const double *coords;       /* [n][3] AoS, as Vec<[f64;3]> (src/dfire.rs:104, src/dna.rs:237) */
  const int32_t *dfire_type;  /* [n] DFIRE atom type 0..167 (src/dfire.rs:177-183) or NULL      */
  const double *ele_charge;   /* [n] (src/dna.rs:245) or NULL                                    */
  const double *vdw_energy;   /* [n] "vdw_charges" (src/dna.rs:244)                              */
  const double *vdw_radius;   /* [n] (src/dna.rs:243)                                            */
  int32_t n_modes;            /* num_anm */
  const double *modes;        /* [n_modes][n][3] flat (src/dfire.rs:292-293)                     */
  int32_t n_restraints;       /* number of active-restraint residues FOUND in the structure      */
  const int32_t *rst_offsets; /* [n_restraints+1] CSR                                            */
  const int32_t *rst_atoms;   /* atom indices                                                    */
  int32_t n_membrane;
  const int32_t *membrane;    /* atom indices of MMB.BJ beads (src/dfire.rs:146-149)             */
} oracle_molecule_t;

typedef struct {
  int32_t method; /* 0 = DFIRE, 1 = DNA/pyDock */
  int32_t use_anm;
  oracle_molecule_t rec, lig;
  const double *dfire_potential; /* 169*169*20 (src/dfire.rs:254) */
} oracle_complex_t;

/* Diagnostics the reference computes internally but does not expose; compared bit-exact. */
typedef struct {
  double raw_sum;            /* DFIRE: sum of table values; DNA: total_elec before *FACTOR/EPSILON */
  double raw_sum2;           /* DNA: total_vdw                                                     */
  int64_t n_in_cutoff;       /* DFIRE: dist<=225 ; DNA: d2<=900                                    */
  int64_t n_in_cutoff2;      /* DNA: d2<=100                                                       */
  int64_t n_interface_pairs; /* pairs satisfying the interface test                                */
  int64_t bin_hist[21];      /* DFIRE: histogram of dfire_bin (0..20)                              */
  int32_t rec_rst_hit, lig_rst_hit, membrane_hit;
} oracle_diag_t;

/* ------------------------------------------------------------------------------------------ */
/* Quaternion algebra: src/qt.rs                                                               */
typedef struct {
  double w, x, y, z;
} oq_t;

/* src/qt.rs:174-185 */
static inline oq_t oq_mul(oq_t a, oq_t b) {
  oq_t r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
/* src/qt.rs:24-26,32-34,48-50,187-198 */
static inline oq_t oq_inverse(oq_t q) {
  double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  oq_t r = {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
  return r;
}
/* src/qt.rs:57-61 */
static inline void oq_rotate(oq_t q, const double v[3], double out[3]) {
  oq_t qv = {0., v[0], v[1], v[2]};
  oq_t r = oq_mul(oq_mul(q, qv), oq_inverse(q));
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
}

ORACLE_API void oracle_rotate(const double q[4], const double v[3], double out[3]) {
  oq_t qq = {q[0], q[1], q[2], q[3]};
  oq_rotate(qq, v, out);
}

static inline double oq_dot(oq_t a, oq_t b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }
/* src/qt.rs:36-46 */
static inline oq_t oq_normalize(oq_t q) {
  double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  oq_t r = {q.w / n, q.x / n, q.y / n, q.z / n};
  return r;
}
static inline oq_t oq_scale(oq_t q, double s) { /* src/qt.rs:161-172: scalar * component */
  oq_t r = {s * q.w, s * q.x, s * q.y, s * q.z};
  return r;
}
static inline oq_t oq_add(oq_t a, oq_t b) {
  oq_t r = {a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z};
  return r;
}
static inline oq_t oq_sub(oq_t a, oq_t b) {
  oq_t r = {a.w - b.w, a.x - b.x, a.y - b.y, a.z - b.z};
  return r;
}
/* src/qt.rs:67-91 */
static oq_t oq_slerp(oq_t self, oq_t other, double t) {
  oq_t q1 = oq_normalize(self), q2 = oq_normalize(other);
  double q_dot = oq_dot(q1, q2);
  if (q_dot < 0.0) {
    q1.w = -q1.w; q1.x = -q1.x; q1.y = -q1.y; q1.z = -q1.z;
    q_dot *= -1.0;
  }
  if (q_dot > 0.9995) { /* LINEAR_THRESHOLD src/constants.rs:11 */
    oq_t result = oq_add(q1, oq_scale(oq_sub(q2, q1), t));
    return oq_normalize(result);
  } else {
    q_dot = fmax(fmin(q_dot, 1.0), -1.0);
    double omega = acos(q_dot);
    double so = sin(omega);
    return oq_add(oq_scale(q1, sin((1.0 - t) * omega) / so), oq_scale(q2, sin(t * omega) / so));
  }
}
ORACLE_API void oracle_slerp(const double a[4], const double b[4], double t, double out[4]) {
  oq_t qa = {a[0], a[1], a[2], a[3]}, qb = {b[0], b[1], b[2], b[3]};
  oq_t r = oq_slerp(qa, qb, t);
  out[0] = r.w; out[1] = r.x; out[2] = r.y; out[3] = r.z;
}

/* ------------------------------------------------------------------------------------------ */
/* Pose transform: src/dfire.rs:275-320 (= src/dna.rs:419-464, src/pydock.rs:434-479)          */
/* pose row = tx,ty,tz,qw,qx,qy,qz, rec extents[n_rec_modes], lig extents[n_lig_modes]          */
static void transform(const oracle_complex_t *c, const double *pose, double *rec, double *lig) {
  const int nr = c->rec.n_atoms, nl = c->lig.n_atoms;
  const double *t = pose;
  oq_t q = {pose[3], pose[4], pose[5], pose[6]};
  const double *rec_ext = pose + 7;
  const double *lig_ext = pose + 7 + (c->use_anm ? c->rec.n_modes : 0);
  memcpy(rec, c->rec.coords, sizeof(double) * 3 * nr);
  memcpy(lig, c->lig.coords, sizeof(double) * 3 * nl);
  for (int i = 0; i < nl; ++i) {
    double r[3];
    oq_rotate(q, &lig[3 * i], r);
    lig[3 * i + 0] = r[0] + t[0];
    lig[3 * i + 1] = r[1] + t[1];
    lig[3 * i + 2] = r[2] + t[2];
    if (c->use_anm && c->lig.n_modes > 0) {
      for (int k = 0; k < c->lig.n_modes; ++k) {
        lig[3 * i + 0] += c->lig.modes[(size_t)k * nl * 3 + i * 3 + 0] * lig_ext[k];
        lig[3 * i + 1] += c->lig.modes[(size_t)k * nl * 3 + i * 3 + 1] * lig_ext[k];
        lig[3 * i + 2] += c->lig.modes[(size_t)k * nl * 3 + i * 3 + 2] * lig_ext[k];
      }
    }
  }
  for (int i = 0; i < nr; ++i) {
    if (c->use_anm && c->rec.n_modes > 0) {
      for (int k = 0; k < c->rec.n_modes; ++k) {
        rec[3 * i + 0] += c->rec.modes[(size_t)k * nr * 3 + i * 3 + 0] * rec_ext[k];
        rec[3 * i + 1] += c->rec.modes[(size_t)k * nr * 3 + i * 3 + 1] * rec_ext[k];
        rec[3 * i + 2] += c->rec.modes[(size_t)k * nr * 3 + i * 3 + 2] * rec_ext[k];
      }
    }
  }
}

/* src/scoring.rs:21-36 — returns the integer count; fraction = count / n_restraints */
static int restraints_hit(const uint8_t *iface, const oracle_molecule_t *m) {
  int num = 0;
  for (int r = 0; r < m->n_restraints; ++r)
    for (int k = m->rst_offsets[r]; k < m->rst_offsets[r + 1]; ++k)
      if (iface[m->rst_atoms[k]] == 1) {
        ++num;
        break;
      }
  return num;
}
/* src/scoring.rs:38-47 */
static int membrane_hit(const uint8_t *iface, const oracle_molecule_t *m) {
  int num = 0;
  for (int k = 0; k < m->n_membrane; ++k) num += iface[m->membrane[k]];
  return num;
}

/* shared epilogue: src/dfire.rs:349-361 = src/dna.rs:516-528 */
static double epilogue(const oracle_complex_t *c, double score, const uint8_t *irec, const uint8_t *ilig,
                       oracle_diag_t *dg) {
  int hr = restraints_hit(irec, &c->rec), hl = restraints_hit(ilig, &c->lig);
  int hm = membrane_hit(irec, &c->rec);
  double pr = c->rec.n_restraints ? (double)hr / (double)c->rec.n_restraints : 0.0;
  double pl = c->lig.n_restraints ? (double)hl / (double)c->lig.n_restraints : 0.0;
  double membrane_penalty = 0.0;
  double intersection = c->rec.n_membrane ? (double)hm / (double)c->rec.n_membrane : 0.0;
  if (intersection > 0.0) membrane_penalty = MEMBRANE_PENALTY_SCORE * intersection;
  if (dg) {
    dg->rec_rst_hit = hr;
    dg->lig_rst_hit = hl;
    dg->membrane_hit = hm;
  }
  return score + pr * score + pl * score - membrane_penalty;
}

/* Rust `d as usize` (saturating, NaN -> 0) for the range that can occur here */
static inline size_t as_usize(double d) {
  if (!(d > 0.0)) return 0;
  return (size_t)d;
}

/* src/dfire.rs:265-362 */
static double dfire_energy(const oracle_complex_t *c, const double *pose, double *rec, double *lig,
                           uint8_t *irec, uint8_t *ilig, oracle_diag_t *dg) {
  const int nr = c->rec.n_atoms, nl = c->lig.n_atoms;
  double score = 0.0;
  transform(c, pose, rec, lig);
  memset(irec, 0, nr);
  memset(ilig, 0, nl);
  if (dg) memset(dg, 0, sizeof(*dg));
  for (int i = 0; i < nr; ++i) {
    const double x1 = rec[3 * i], y1 = rec[3 * i + 1], z1 = rec[3 * i + 2];
    const int atoma = c->rec.dfire_type[i];
    for (int j = 0; j < nl; ++j) {
      const double *la = &lig[3 * j];
      double dist = (x1 - la[0]) * (x1 - la[0]) + (y1 - la[1]) * (y1 - la[1]) + (z1 - la[2]) * (z1 - la[2]);
      if (dist <= 225.) {
        const int atomb = c->lig.dfire_type[j];
        double d = sqrt(dist) * 2.0 - 1.0;
        int dfire_bin = DIST_TO_BINS[as_usize(d)] - 1;
        score += c->dfire_potential[atoma * 169 * 20 + atomb * 20 + dfire_bin];
        if (dg) {
          dg->n_in_cutoff++;
          dg->bin_hist[dfire_bin]++;
        }
        if (d <= INTERFACE_CUTOFF) {
          irec[i] = 1;
          ilig[j] = 1;
          if (dg) dg->n_interface_pairs++;
        }
      }
    }
  }
  if (dg) dg->raw_sum = score;
  score = (score * 0.0157 - 4.7) * -1.0;
  return epilogue(c, score, irec, ilig, dg);
}

/* Rust f64::powi lowers to llvm.powi = repeated squaring (compiler-rt __powidf2):
 * r=1; loop { if (b&1) r*=a; b/=2; if(!b) break; a*=a; }                                       */
static inline double powi(double a, int b) {
  double r = 1.0;
  for (;;) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return r;
}

/* src/dna.rs:411-529 = src/pydock.rs:426-544 */
static double dna_energy(const oracle_complex_t *c, const double *pose, double *rec, double *lig, uint8_t *irec,
                         uint8_t *ilig, oracle_diag_t *dg) {
  const int nr = c->rec.n_atoms, nl = c->lig.n_atoms;
  const double ELEC_DIST_CUTOFF2 = ELEC_DIST_CUTOFF * ELEC_DIST_CUTOFF;
  const double VDW_DIST_CUTOFF2 = VDW_DIST_CUTOFF * VDW_DIST_CUTOFF;
  const double ELEC_MAX_CUTOFF = MAX_ES_CUTOFF * DNA_EPSILON / DNA_FACTOR;
  const double ELEC_MIN_CUTOFF = MIN_ES_CUTOFF * DNA_EPSILON / DNA_FACTOR;
  const double INTERFACE_CUTOFF2 = INTERFACE_CUTOFF * INTERFACE_CUTOFF;
  transform(c, pose, rec, lig);
  memset(irec, 0, nr);
  memset(ilig, 0, nl);
  if (dg) memset(dg, 0, sizeof(*dg));
  double total_elec = 0.0, total_vdw = 0.0;
  for (int i = 0; i < nr; ++i) {
    const double x1 = rec[3 * i], y1 = rec[3 * i + 1], z1 = rec[3 * i + 2];
    for (int j = 0; j < nl; ++j) {
      const double *la = &lig[3 * j];
      double distance2 =
          (x1 - la[0]) * (x1 - la[0]) + (y1 - la[1]) * (y1 - la[1]) + (z1 - la[2]) * (z1 - la[2]);
      if (distance2 <= ELEC_DIST_CUTOFF2) {
        double atom_elec = c->rec.ele_charge[i] * c->lig.ele_charge[j] / distance2;
        if (atom_elec > ELEC_MAX_CUTOFF) atom_elec = ELEC_MAX_CUTOFF;
        if (atom_elec < ELEC_MIN_CUTOFF) atom_elec = ELEC_MIN_CUTOFF;
        total_elec += atom_elec;
        if (dg) dg->n_in_cutoff++;
      }
      if (distance2 <= VDW_DIST_CUTOFF2) {
        double vdw_energy = sqrt(c->rec.vdw_energy[i] * c->lig.vdw_energy[j]);
        double vdw_radius = c->rec.vdw_radius[i] + c->lig.vdw_radius[j];
        double p6 = powi(vdw_radius, 6) / powi(distance2, 3);
        double k = vdw_energy * (p6 * p6 - 2.0 * p6);
        if (k > VDW_CUTOFF) k = VDW_CUTOFF;
        total_vdw += k;
        if (dg) dg->n_in_cutoff2++;
      }
      if (distance2 <= INTERFACE_CUTOFF2) {
        irec[i] = 1;
        ilig[j] = 1;
        if (dg) dg->n_interface_pairs++;
      }
    }
  }
  if (dg) {
    dg->raw_sum = total_elec;
    dg->raw_sum2 = total_vdw;
  }
  total_elec = total_elec * DNA_FACTOR / DNA_EPSILON;
  double score = (total_elec + total_vdw) * -1.0;
  return epilogue(c, score, irec, ilig, dg);
}

/* ------------------------------------------------------------------------------------------ */
/* Public oracle entry points                                                                  */
typedef struct {
  double *rec, *lig;
  uint8_t *irec, *ilig;
} scratch_t;

static int scratch_init(scratch_t *s, const oracle_complex_t *c) {
  s->rec = (double *)malloc(sizeof(double) * 3 * (c->rec.n_atoms + 1));
  s->lig = (double *)malloc(sizeof(double) * 3 * (c->lig.n_atoms + 1));
  s->irec = (uint8_t *)malloc(c->rec.n_atoms + 1);
  s->ilig = (uint8_t *)malloc(c->lig.n_atoms + 1);
  return s->rec && s->lig && s->irec && s->ilig;
}
static void scratch_free(scratch_t *s) {
  free(s->rec); free(s->lig); free(s->irec); free(s->ilig);
}
static double energy_one(const oracle_complex_t *c, const double *pose, scratch_t *s, oracle_diag_t *dg) {
  return c->method == 0 ? dfire_energy(c, pose, s->rec, s->lig, s->irec, s->ilig, dg)
                        : dna_energy(c, pose, s->rec, s->lig, s->irec, s->ilig, dg);
}

static int pose_len(const oracle_complex_t *c) {
  return 7 + (c->use_anm ? c->rec.n_modes + c->lig.n_modes : 0);
}

/* Score::energy for n poses, sequentially (src/swarm.rs:66-70 calls it once per glowworm).
 * diag / iface_rec / iface_lig / coords_* may be NULL. */
ORACLE_API int oracle_score_batch(const oracle_complex_t *c, int n_poses, const double *poses, double *energies,
                                  oracle_diag_t *diag, uint8_t *iface_rec, uint8_t *iface_lig,
                                  double *coords_rec, double *coords_lig) {
  scratch_t s;
  if (!scratch_init(&s, c)) return -1;
  const int pl = pose_len(c);
  for (int p = 0; p < n_poses; ++p) {
    energies[p] = energy_one(c, poses + (size_t)p * pl, &s, diag ? &diag[p] : NULL);
    if (iface_rec) memcpy(iface_rec + (size_t)p * c->rec.n_atoms, s.irec, c->rec.n_atoms);
    if (iface_lig) memcpy(iface_lig + (size_t)p * c->lig.n_atoms, s.ilig, c->lig.n_atoms);
    if (coords_rec) memcpy(coords_rec + (size_t)p * 3 * c->rec.n_atoms, s.rec, sizeof(double) * 3 * c->rec.n_atoms);
    if (coords_lig) memcpy(coords_lig + (size_t)p * 3 * c->lig.n_atoms, s.lig, sizeof(double) * 3 * c->lig.n_atoms);
  }
  scratch_free(&s);
  return 0;
}

/* CPU baseline: the same scalar loop, poses split over n_threads host threads (the reference's
 * scale-out model is one process per core over distinct swarms, example/1czy/execution.sh:21-25;
 * poses are independent, so threads over poses measure the same thing). */
typedef struct {
  const oracle_complex_t *c;
  const double *poses;
  double *energies;
  int begin, end, ok;
} mt_job_t;
static void *mt_worker(void *arg) {
  mt_job_t *j = (mt_job_t *)arg;
  scratch_t s;
  j->ok = scratch_init(&s, j->c);
  if (!j->ok) return NULL;
  const int pl = pose_len(j->c);
  for (int p = j->begin; p < j->end; ++p) j->energies[p] = energy_one(j->c, j->poses + (size_t)p * pl, &s, NULL);
  scratch_free(&s);
  return NULL;
}
ORACLE_API int oracle_score_batch_mt(const oracle_complex_t *c, int n_poses, const double *poses, double *energies,
                                     int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_poses) n_threads = n_poses > 0 ? n_poses : 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
  mt_job_t *jobs = (mt_job_t *)malloc(sizeof(mt_job_t) * n_threads);
  for (int t = 0; t < n_threads; ++t) {
    jobs[t].c = c; jobs[t].poses = poses; jobs[t].energies = energies;
    jobs[t].begin = (int)((int64_t)n_poses * t / n_threads);
    jobs[t].end = (int)((int64_t)n_poses * (t + 1) / n_threads);
    jobs[t].ok = 0;
    pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
  }
  int ok = 1;
  for (int t = 0; t < n_threads; ++t) {
    pthread_join(th[t], NULL);
    ok &= jobs[t].ok;
  }
  free(th); free(jobs);
  return ok ? 0 : -1;
}

/* ------------------------------------------------------------------------------------------ */
/* rand 0.7.3 StdRng = rand_chacha ChaCha20Rng (third-party crate, not vendored in the          */
/* reference; semver range Cargo.toml:12).  Published algorithm restated:                        */
/*  - SeedableRng::seed_from_u64: PCG32 (MUL 6364136223846793005, INC 11634580027462260723)      */
/*    expands the u64 into the 32-byte key, 4 bytes per PCG step, little endian;                  */
/*  - ChaCha with 20 rounds, 64-bit block counter (words 12,13) from 0, 64-bit stream id 0;       */
/*  - BlockRng::next_u64 = lo | hi<<32 from consecutive u32 words; gen::<f64>() = (u64>>11)*2^-53 */
/* Call sites: src/lib.rs:38, src/swarm.rs:118.  Pinned by src/qt.rs:451-462.                     */
typedef struct {
  uint32_t key[8];
  uint64_t counter;
  uint32_t buf[16];
  int idx;
} oracle_rng_t;

#define ROTL32(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define QR(a, b, c, d) \
  a += b; d ^= a; d = ROTL32(d, 16); c += d; b ^= c; b = ROTL32(b, 12); \
  a += b; d ^= a; d = ROTL32(d, 8);  c += d; b ^= c; b = ROTL32(b, 7);

static void chacha20_block(oracle_rng_t *r) {
  uint32_t s[16], x[16];
  s[0] = 0x61707865; s[1] = 0x3320646e; s[2] = 0x79622d32; s[3] = 0x6b206574;
  for (int i = 0; i < 8; ++i) s[4 + i] = r->key[i];
  s[12] = (uint32_t)r->counter; s[13] = (uint32_t)(r->counter >> 32);
  s[14] = 0; s[15] = 0;
  memcpy(x, s, sizeof(s));
  for (int i = 0; i < 10; ++i) {
    QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
    QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 16; ++i) r->buf[i] = x[i] + s[i];
  r->counter++;
  r->idx = 0;
}
ORACLE_API void oracle_rng_seed(oracle_rng_t *r, uint64_t state) {
  const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
  for (int i = 0; i < 8; ++i) {
    state = state * MUL + INC;
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    r->key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
  }
  r->counter = 0;
  r->idx = 16;
}
static uint32_t rng_u32(oracle_rng_t *r) {
  if (r->idx >= 16) chacha20_block(r);
  return r->buf[r->idx++];
}
ORACLE_API double oracle_rng_f64(oracle_rng_t *r) {
  uint64_t lo = rng_u32(r), hi = rng_u32(r);
  uint64_t v = lo | (hi << 32);
  return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}
ORACLE_API int oracle_rng_sizeof(void) { return (int)sizeof(oracle_rng_t); }

/* ------------------------------------------------------------------------------------------ */
/* GSO host loop: src/lib.rs:46-58, src/swarm.rs:26-167, src/glowworm.rs:29-202                */
typedef struct {
  double *pose; /* [pose_len] translation, quaternion, rec extents, lig extents */
  double luciferin, vision_range, scoring;
  int moved, step;
  int n_neighbors;
  int *neighbors;
  double *probabilities;
} gw_t;

typedef struct {
  int64_t n_energy_calls;
} oracle_gso_stats_t;

static int write_swarm(const gw_t *g, int n, int plen, int use_anm, int step, const char *out_dir) {
  char path[4096];
  snprintf(path, sizeof(path), "%s/gso_%d.out", out_dir, step);
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  fprintf(f, "#Coordinates  RecID  LigID  Luciferin  Neighbor's number  Vision Range  Scoring\n");
  for (int i = 0; i < n; ++i) {
    fprintf(f, "(%.7f, %.7f, %.7f, %.7f, %.7f, %.7f, %.7f", g[i].pose[0], g[i].pose[1], g[i].pose[2], g[i].pose[3],
            g[i].pose[4], g[i].pose[5], g[i].pose[6]);
    if (use_anm)
      for (int k = 7; k < plen; ++k) fprintf(f, ", %.7f", g[i].pose[k]);
    fprintf(f, ")    0    0   %.8f  %d %.3f %.8f\n", g[i].luciferin, g[i].n_neighbors, g[i].vision_range,
            g[i].scoring);
  }
  fclose(f);
  return 0;
}

/* Runs `steps` GSO steps with the oracle energy.  out_dir may be NULL (no files).  If `trace` is
 * non-NULL it receives, per step and glowworm, [luciferin, scoring, n_neighbors, vision, moved]
 * (5 doubles) followed by the pose row, i.e. steps*n*(5+pose_len) doubles. */
/* n_threads > 1 only changes WHO evaluates a step's energies (the step's batch of moved glowworms is scored by
 * oracle_score_batch_mt, each pose by the same energy_one): every number is bit-identical to the scalar run. */
ORACLE_API int oracle_gso_run_mt(const oracle_complex_t *c, int n, const double *positions, uint64_t seed, int steps,
                                 const char *out_dir, double *final_poses, double *trace, oracle_gso_stats_t *stats,
                                 int n_threads) {
  const int plen = pose_len(c);
  const int nrm = c->use_anm ? c->rec.n_modes : 0, nlm = c->use_anm ? c->lig.n_modes : 0;
  scratch_t s;
  if (!scratch_init(&s, c)) return -1;
  gw_t *g = (gw_t *)calloc(n, sizeof(gw_t));
  double *snap = (double *)malloc(sizeof(double) * n * plen);
  double *lucs = (double *)malloc(sizeof(double) * n);
  double *batch = (double *)malloc(sizeof(double) * n * plen), *batch_e = (double *)malloc(sizeof(double) * n);
  int *batch_i = (int *)malloc(sizeof(int) * n);
  for (int i = 0; i < n; ++i) { /* src/glowworm.rs:29-59 */
    g[i].pose = (double *)malloc(sizeof(double) * plen);
    memcpy(g[i].pose, positions + (size_t)i * plen, sizeof(double) * plen);
    g[i].luciferin = 5.0;
    g[i].vision_range = 0.2;
    g[i].neighbors = (int *)malloc(sizeof(int) * n);
    g[i].probabilities = (double *)malloc(sizeof(double) * n);
  }
  oracle_rng_t rng;
  oracle_rng_seed(&rng, seed);
  int64_t calls = 0;
  const double rho = 0.5, gamma = 0.4, beta = 0.08, max_vision_range = 5.0;
  const int max_neighbors = 5;
  for (int step = 1; step <= steps; ++step) {
    /* update_luciferin: src/swarm.rs:66-70, src/glowworm.rs:61-72 */
    if (n_threads > 1) {
      int nb = 0;
      for (int i = 0; i < n; ++i)
        if (g[i].moved || g[i].step == 0) {
          memcpy(batch + (size_t)nb * plen, g[i].pose, sizeof(double) * plen);
          batch_i[nb++] = i;
        }
      if (nb > 0 && oracle_score_batch_mt(c, nb, batch, batch_e, n_threads) != 0) return -1;
      for (int k = 0; k < nb; ++k) g[batch_i[k]].scoring = batch_e[k];
      calls += nb;
    }
    for (int i = 0; i < n; ++i) {
      if (n_threads <= 1 && (g[i].moved || g[i].step == 0)) {
        g[i].scoring = energy_one(c, g[i].pose, &s, NULL);
        ++calls;
      }
      g[i].luciferin = (1.0 - rho) * g[i].luciferin + gamma * g[i].scoring;
      g[i].step += 1;
    }
    /* movement_phase: src/swarm.rs:72-126 */
    for (int i = 0; i < n; ++i) memcpy(snap + (size_t)i * plen, g[i].pose, sizeof(double) * plen);
    for (int i = 0; i < n; ++i) {
      g[i].n_neighbors = 0;
      for (int j = 0; j < n; ++j) {
        if (i == j) continue;
        if (g[i].luciferin < g[j].luciferin) {
          double x1 = g[i].pose[0], x2 = g[j].pose[0], y1 = g[i].pose[1], y2 = g[j].pose[1];
          double z1 = g[i].pose[2], z2 = g[j].pose[2];
          double distance = sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
          if (distance < g[i].vision_range) g[i].neighbors[g[i].n_neighbors++] = j;
        }
      }
    }
    for (int i = 0; i < n; ++i) lucs[i] = g[i].luciferin;
    for (int i = 0; i < n; ++i) { /* src/glowworm.rs:98-112 */
      double total_sum = 0.0;
      for (int k = 0; k < g[i].n_neighbors; ++k) {
        double difference = lucs[g[i].neighbors[k]] - g[i].luciferin;
        g[i].probabilities[k] = difference;
        total_sum += difference;
      }
      for (int k = 0; k < g[i].n_neighbors; ++k) g[i].probabilities[k] /= total_sum;
    }
    for (int i = 0; i < n; ++i) {
      double rnd = oracle_rng_f64(&rng); /* src/swarm.rs:118: always one draw per glowworm */
      int nid = i;                       /* src/glowworm.rs:114-126 */
      if (g[i].n_neighbors > 0) {
        double sum_probabilities = 0.0;
        int k = 0;
        while (sum_probabilities < rnd) {
          if (k >= g[i].n_neighbors) { /* the reference would panic on the out-of-bounds index */
            fprintf(stderr, "oracle_gso_run: roulette ran past the neighbour list (reference panics)\n");
            return -2;
          }
          sum_probabilities += g[i].probabilities[k];
          ++k;
        }
        /* rnd == 0.0 with k == 0 makes the reference index neighbors[usize::MAX] -> panic */
        if (k == 0) return -2;
        nid = g[i].neighbors[k - 1];
      }
      /* move_towards: src/glowworm.rs:128-190 */
      g[i].moved = (nid != i);
      if (nid != i) {
        const double *o = snap + (size_t)nid * plen;
        double *p = g[i].pose;
        double dx[3] = {o[0] - p[0], o[1] - p[1], o[2] - p[2]};
        double norm = sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
        double coef = 0.5 / norm;
        dx[0] *= coef; dx[1] *= coef; dx[2] *= coef;
        p[0] += dx[0]; p[1] += dx[1]; p[2] += dx[2];
        oq_t qs = {p[3], p[4], p[5], p[6]}, qo = {o[3], o[4], o[5], o[6]};
        oq_t qr = oq_slerp(qs, qo, 0.5);
        p[3] = qr.w; p[4] = qr.x; p[5] = qr.y; p[6] = qr.z;
        for (int part = 0; part < 2; ++part) {
          int off = part == 0 ? 7 : 7 + nrm, cnt = part == 0 ? nrm : nlm;
          if (cnt == 0) continue;
          double cum_norm = 0.0;
          double delta[64];
          for (int k = 0; k < cnt; ++k) {
            double diff = o[off + k] - p[off + k];
            delta[k] = diff;
            cum_norm += diff * diff;
          }
          double anm_coef = 0.5 / sqrt(cum_norm);
          for (int k = 0; k < cnt; ++k) {
            delta[k] *= anm_coef;
            p[off + k] += delta[k];
          }
        }
      }
      /* update_vision_range: src/glowworm.rs:91-96 */
      g[i].vision_range =
          fmin(max_vision_range, fmax(0.0, g[i].vision_range + beta * (double)(max_neighbors - g[i].n_neighbors)));
    }
    if (trace) {
      double *t = trace + (size_t)(step - 1) * n * (5 + plen);
      for (int i = 0; i < n; ++i) {
        double *r = t + (size_t)i * (5 + plen);
        r[0] = g[i].luciferin; r[1] = g[i].scoring; r[2] = g[i].n_neighbors; r[3] = g[i].vision_range;
        r[4] = g[i].moved;
        memcpy(r + 5, g[i].pose, sizeof(double) * plen);
      }
    }
    if (out_dir && (step % 10 == 0 || step == 1)) /* src/lib.rs:51 */
      if (write_swarm(g, n, plen, c->use_anm && (nrm + nlm) > 0, step, out_dir)) return -3;
  }
  if (final_poses)
    for (int i = 0; i < n; ++i) memcpy(final_poses + (size_t)i * plen, g[i].pose, sizeof(double) * plen);
  if (stats) stats->n_energy_calls = calls;
  for (int i = 0; i < n; ++i) { free(g[i].pose); free(g[i].neighbors); free(g[i].probabilities); }
  free(g); free(snap); free(lucs); free(batch); free(batch_e); free(batch_i);
  scratch_free(&s);
  return 0;
}

ORACLE_API int oracle_gso_run(const oracle_complex_t *c, int n, const double *positions, uint64_t seed, int steps,
                              const char *out_dir, double *final_poses, double *trace,
                              oracle_gso_stats_t *stats) {
  return oracle_gso_run_mt(c, n, positions, seed, steps, out_dir, final_poses, trace, stats, 1);
}

ORACLE_API int oracle_sizeof_complex(void) { return (int)sizeof(oracle_complex_t); }
ORACLE_API int oracle_sizeof_diag(void) { return (int)sizeof(oracle_diag_t); }

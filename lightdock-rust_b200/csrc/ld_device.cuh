// ld_device.cuh — device-side data layout shared by the kernels and the C-ABI host code.
//
// HBM layout (all immutable per complex, built once in ld_create):
//   receptor, sorted into spatial tiles of REC_TILE=32 atoms (one atom per lane of a warp):
//     SoA f64 x[], y[], z[] padded to a tile multiple (pads at +1e30), per-atom DFIRE table row offset
//     (type*3380) or DNA charge/eps/radius, one float4 bounding sphere per tile.
//   ligand, sorted into spatial tiles of LIG_TILE=8 atoms: local-frame SoA f64 coordinates, per-atom
//     DFIRE column offset (type*20, u16) or DNA parameters.
//   ANM modes re-ordered to [mode][xyz][sorted atom] so a warp reads them coalesced.
//   DFIRE potential: 571,220 f64 (4.57 MB) — stays resident in the 126 MB L2.
// Per batch: one "ligand block" per pose written by the transform kernel (16-byte aligned, contiguous):
//     [x f64][y f64][z f64]  exact transformed coordinates (n_lig_pad each)
//     [float4 xyzt]          DFIRE: the same coordinates rounded to f32 + the table column offset (type * RG_SLOTS) as a float
//     [double4 xyzq]         DNA/pyDock instead: the exact coordinates + the atom's charge, interleaved per atom
//     [float4 sphere]        one conservative bounding sphere per ligand tile
//     [float4 meta]          meta.x = max |coordinate| of the pose's ligand (sets the rigorous f32 margins)
//   The pair kernels pull the part they need into shared memory with cp.async.bulk (TMA) copies:
//   DFIRE takes [xyzt|sphere|meta] (its f64 coordinates are touched only by the rare exact fallback, from
//   L2), DNA takes [xyzq|sphere|meta].  A "receptor block" ([x][y][z][sphere][meta]) exists per
//   pose only when receptor ANM is active.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldb200 {

constexpr int REC_TILE = 32;  // receptor atoms per tile = lanes of a warp
constexpr int LIG_TILE = 8;   // ligand atoms per tile (unit of sphere culling)
constexpr int PAIR_THREADS = 512;  // generic DFIRE pair kernel
#ifndef LDB200_DNA_THREADS
#define LDB200_DNA_THREADS 192
#endif
#ifndef LDB200_DNA_CTAS
#define LDB200_DNA_CTAS 6
#endif
// DNA/pyDock pair kernel: small complexes (1azp: 35 receptor tiles), so small CTAs, many per SM.  Measured on 1azp (20,000
// poses, pair kernel): 6 x 192 threads (56 registers; 35 tiles on 6 warps = 6 rounds, 97 % of the warp slots used) 11.31 ms;
// 4 x 256 (64 registers; 8 warps = 5 rounds, 87.5 %) 11.57; 5 x 256 (48) 11.51; 8 x 128 (64) 11.75; 3 x 256 (80) 12.6;
// 2 x 256 (126) 12.0.  The landscape is flat: the kernel sits at the dispatch rate of its instruction mix (an FP64
// instruction holds the dispatch port for two cycles), not at an occupancy or latency limit.
constexpr int DNA_CTAS_PER_SM = LDB200_DNA_CTAS;
constexpr int DNA_THREADS = LDB200_DNA_THREADS;
constexpr int DFIRE_ROW = 169 * 20;  // src/dfire.rs:338  atoma*169*20
constexpr double REC_PAD = 1.0e30;   // coordinates of padding atoms: never within any cut-off
constexpr double LIG_PAD = -1.0e30;

// DFIRE table re-indexed by the truncated bin-space value idx = floor(2*sqrt(dist) - 1) instead of the bin
// (`potx`, built by ld_create): both DFIRE kernels classify a pair by rounding t - 0.5 with the magic-number add
// x + 1.5*2^23, which leaves MAGIC_BITS + rint(x) in the float's bit pattern.
constexpr int RG_SLOT0 = -1;                        // first bin-space index held in a shared-memory row: rint(t - 0.5)
                                                    // is -1 for t < 0 (dist < 0.25), which the reference's saturating
                                                    // `d as usize` sends to index 0, so slot -1 repeats slot 0
constexpr int RG_SLOTS = 30;                        // indices -1..28 (29, the cut-off itself, is never decided in FP32)
constexpr int RG_PREP = 16;                         // doubles per pose written by rigid_prep_kernel
constexpr int RG_TB_BYTES = RG_SLOTS * 8;           // one ligand type inside a row
constexpr int RG_ROW_BYTES = (169 * RG_TB_BYTES + 15) / 16 * 16;  // 40,560
constexpr float RG_MAGIC = 12582912.0f;             // 1.5 * 2^23: x + MAGIC rounds x to the nearest integer
constexpr unsigned RG_MAGIC_BITS = 0x4B400000u;

struct DeviceComplex {
  int method;  // 0 DFIRE, 1 DNA/pyDock
  int n_rec, n_lig;
  int n_rec_pad, n_lig_pad;
  int n_rec_tiles, n_lig_tiles;
  int n_rec_modes, n_lig_modes;  // effective (0 when use_anm is false)
  int pose_len;
  // receptor (sorted order)
  const double *rec_x, *rec_y, *rec_z;
  const int *rec_toff;                      // DFIRE: type * 3380 (exact path)
  const unsigned *rec_rowx;                 // DFIRE: type * (RG_ROW_BYTES/8) - RG_SLOT0 - RG_MAGIC_BITS (mod 2^32): element
                                            // index into potx = bits(m + w) + rec_rowx, see dfire_items()
  const double *rec_q, *rec_seps, *rec_rad;  // DNA: charge, sqrt(vdw energy) (the energy itself if !vdw_sqrt_hoisted), radius
  // DNA: van der Waals (A, B) = (sqrt(e_r e_l) (r_r + r_l)^12, 2 sqrt(e_r e_l) (r_r + r_l)^6) per pair of distinct
  // (energy, radius) types, [vdw_nl][vdw_nr]; nullptr when there are more than 1024 pairs (the kernel then evaluates
  // the term from the per-atom parameters); rec_vt = type * 16, lig_vt = type * vdw_nr * 16 (byte offsets)
  const double2 *vdw_tab;
  const int *rec_vt, *lig_vt;
  int vdw_nr, vdw_nl;
  float dna_close_reach;                    // DNA: max(10 A, largest distance at which q_r q_l / d2 can reach the +-4/332 clamp)
  int vdw_sqrt_hoisted;                     // DNA: 1 = *_seps hold square roots (every vdw energy >= 0)
  const float4 *rec_sphere;                 // static tile spheres (used when n_rec_modes == 0)
  float rec_maxabs;                         // max |coordinate| of the static receptor
  const double *rec_modes;                  // [k][3][n_rec_pad]
  // ligand (sorted order, local frame)
  const double *lig_x, *lig_y, *lig_z;
  const unsigned short *lig_tb20;           // DFIRE: type * 20
  const double *lig_q, *lig_seps, *lig_rad;  // DNA
  const double *lig_modes;                  // [k][3][n_lig_pad]
  const double *pot;                        // DFIRE table as the reference indexes it
  const double *potx;                       // [169][RG_ROW_BYTES/8]: row ta = [tb][RG_SLOTS], slot s <-> idx s + RG_SLOT0
  double fx_inv_scale;                      // DFIRE ligand-frame FLEX path: per-group sums are 64-bit fixed point, value = sum * this; 0 = f64 sums
  // restraints (sorted atom positions) and membrane beads
  int n_rec_rst, n_lig_rst, n_membrane;
  const int *rec_rst_off, *rec_rst_idx, *lig_rst_off, *lig_rst_idx, *membrane_idx;
};

// byte offsets inside one ligand block; `wide` = bytes per atom of the staged part: 16 for DFIRE (float4 x,y,z,type
// offset), 32 for DNA/pyDock (double x,y | z,charge: the pair loop reads an atom with two broadcast LDS.128)
__host__ __device__ inline int lig_wide(int method) { return method == 0 ? 16 : 32; }
__host__ __device__ inline size_t lig_off_f4(int n_pad) { return (size_t)n_pad * 24; }
__host__ __device__ inline size_t lig_off_sph(int n_pad, int method) { return (size_t)n_pad * (24 + lig_wide(method)); }
__host__ __device__ inline size_t lig_off_meta(int n_pad, int n_tiles, int method) {
  return lig_off_sph(n_pad, method) + (size_t)n_tiles * 16;
}
__host__ __device__ inline size_t lig_block_bytes(int n_pad, int n_tiles, int method) {
  return lig_off_meta(n_pad, n_tiles, method) + 16;
}
// receptor block (ANM only): x,y,z f64 [n_pad], float4 spheres [n_tiles], float4 meta
__host__ __device__ inline size_t rec_block_bytes(int n_pad, int n_tiles) {
  return (size_t)n_pad * 24 + (size_t)n_tiles * 16 + 16;
}

struct BatchBuffers {
  const double *poses;      // [n][pose_len]
  unsigned char *lig_blocks;  // [n][lig_block_bytes]
  unsigned char *rec_blocks;  // [n][rec_block_bytes] (receptor ANM only)
  double *partials;         // per-tile sums: DFIRE [n][n_rec_tiles], DNA [n][n_rec_tiles][2]
  unsigned *iface_rec;      // [n][n_rec_tiles]
  unsigned *iface_lig;      // [n][rec_splits][lig_words]
  double *energies;         // [n]
  void *detail;             // ld_pose_detail[n] or nullptr
  int rec_splits;
  int tiles_per_split;
  int lig_words;
  // Device-resident callers (ld_gso.cuh) only know on the DEVICE how many of the rows are real: when n_live is set,
  // rows at and beyond *n_live - n_live_off are skipped by every kernel (the launch is sized for the capacity).
  const int *n_live;
  int n_live_off;
};
__device__ __forceinline__ int live_poses(const BatchBuffers &bb, int n_poses) {
  if (bb.n_live == nullptr) return n_poses;
  return max(0, min(n_poses, *bb.n_live - bb.n_live_off));
}

}  // namespace ldb200

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
//     [float4 xyzt]          the same coordinates rounded to f32 + the DFIRE column offset (type*20) as int bits
//     [float4 sphere]        one conservative bounding sphere per ligand tile
//     [float4 meta]          meta.x = max |coordinate| of the pose's ligand (sets the rigorous f32 margins)
//   The pair kernels pull the part they need into shared memory with cp.async.bulk (TMA) copies:
//   DFIRE takes [xyzt|sphere|meta] (its f64 coordinates are touched only by the rare exact fallback, from
//   L2), DNA takes [x|y|z] and [sphere|meta].  A "receptor block" ([x][y][z][sphere][meta]) exists per
//   pose only when receptor ANM is active.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldb200 {

constexpr int REC_TILE = 32;  // receptor atoms per tile = lanes of a warp
constexpr int LIG_TILE = 8;   // ligand atoms per tile (unit of sphere culling)
constexpr int PAIR_THREADS = 512;  // generic DFIRE pair kernel
#ifndef LDB200_DNA_THREADS
#define LDB200_DNA_THREADS 256
#endif
constexpr int DNA_THREADS = LDB200_DNA_THREADS;  // DNA/pyDock pair kernel: small complexes (1azp: 35 receptor tiles), so
                                                 // smaller CTAs (more per SM) balance the tiles over the warps better
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
  const double *rec_q, *rec_eps, *rec_rad;  // DNA
  const float4 *rec_sphere;                 // static tile spheres (used when n_rec_modes == 0)
  float rec_maxabs;                         // max |coordinate| of the static receptor
  const double *rec_modes;                  // [k][3][n_rec_pad]
  // ligand (sorted order, local frame)
  const double *lig_x, *lig_y, *lig_z;
  const unsigned short *lig_tb20;           // DFIRE: type * 20
  const double *lig_q, *lig_eps, *lig_rad;  // DNA
  const double *lig_modes;                  // [k][3][n_lig_pad]
  const double *pot;                        // DFIRE table as the reference indexes it
  const double *potx;                       // [169][RG_ROW_BYTES/8]: row ta = [tb][RG_SLOTS], slot s <-> idx s + RG_SLOT0
  // restraints (sorted atom positions) and membrane beads
  int n_rec_rst, n_lig_rst, n_membrane;
  const int *rec_rst_off, *rec_rst_idx, *lig_rst_off, *lig_rst_idx, *membrane_idx;
};

// byte offsets inside one ligand block
__host__ __device__ inline size_t lig_off_f4(int n_pad) { return (size_t)n_pad * 24; }
__host__ __device__ inline size_t lig_off_sph(int n_pad) { return (size_t)n_pad * 40; }
__host__ __device__ inline size_t lig_off_meta(int n_pad, int n_tiles) { return (size_t)n_pad * 40 + (size_t)n_tiles * 16; }
__host__ __device__ inline size_t lig_block_bytes(int n_pad, int n_tiles) { return lig_off_meta(n_pad, n_tiles) + 16; }
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
};

}  // namespace ldb200

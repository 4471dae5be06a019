// ld_rigid.cuh — DFIRE pair loop (src/dfire.rs:325-345) in the LIGAND's frame: the fast path of the library.
// Written for a rigid ligand (no ligand ANM modes: the 1k4c-class workloads); a ligand WITH ANM modes takes the
// same kernel in its FLEX instantiation (end of this comment).
//
// Idea.  |r - (R l + t)| = |R^-1 (r - t) - l|: instead of moving the 3,268 ligand atoms of every pose
// into the lab frame, each receptor atom is moved into the ligand's LOCAL frame (one 3x3 f64 product
// per atom and pose).  There the ligand never changes, so everything that depends on it is built once
// in ld_create and shared by all poses:
//   * lig4[]      f32 local coordinates + table column offset, resident in shared memory for the whole
//                 life of a CTA (no per-pose staging, no transform kernel on this path);
//   * a uniform cell grid over the ligand's bounding box grown by the 15 A cut-off; each cell lists the
//                 ligand tiles (8 atoms) holding at least one atom within 15 A (+slack) of the cell box.
//                 The per-pose culling of the generic kernel (sphere tests, levels A and B) becomes one
//                 cell lookup per receptor atom.
// Because the cell lookup is per ATOM, receptor atoms need no spatial coherence, so they are grouped
// by DFIRE TYPE instead: a group is <= 32 atoms (one per lane) of <= 4 types, and the <= 4 table rows
// those types index (169 ligand types x 25 reachable distance slots x f64 = 33.8 KB each) are pulled
// into shared memory by TMA when a CTA switches group.  The table gather of the hot loop is therefore a
// shared-memory load; the 4.57 MB table is read from L2 only on group switches and by the exact path.
//
// Work decomposition: persistent CTAs (one per SM, 32 warps).  A work unit is (receptor group, range of
// poses); CTAs pull units from a global counter, group-major, so a CTA changes rows rarely.  Inside a
// unit every warp pulls poses from a shared-memory counter and scores (its group x that pose) alone:
// no CTA-wide barrier inside a unit, and a pose's result never depends on the batch it came in.
//
// Exactness.  Same contract as the generic kernel: the distance is classified in FP32 and the decision
// is accepted only when it is PROVABLY the reference's FP64 decision; everything else (and every pair
// in the thin 2.4-2.5 A shell around the 2.45 A interface edge) is re-evaluated with the reference's own
// lab-frame FP64 arithmetic (rigid_exact_pair).  See the bound next to rigid_row().
//
// FLEX (ligand with ANM modes, src/dfire.rs:290-301).  In the ligand's frame atom j of pose p sits at
// l_j + R^T D_j(p), D_j = sum_k mode_k[j] * extent_k(p): the ligand is no longer the same for every pose, but it moves
// little.  flex_prep_kernel writes, per pose, the f32 ligand block (same layout as lig4) and the largest displacement
// of every ligand tile; the cell lists are built with a per-tile SLACK added to the 15 A reach, so a list still holds
// every tile that can be in range as long as the pose's tile displacements stay below the slacks.  A pose that exceeds
// one (the first batch of a fresh handle, before the slacks have been learnt) is scored in the same launch by brute
// force over all ligand tiles -- still exact -- and the host then grows the slacks and rebuilds the lists for the
// following calls.  Each warp stages the block of the pose it is working on into its own slice of shared memory.
// Because the lists now depend on the history of the handle, the per-lane summation order would too; the FLEX
// instance therefore accumulates the table values in 64-bit FIXED POINT (table * 2^k rounded once, at ld_create):
// integer sums are exact, so a pose's energy is bit-identical whatever the lists, the batch or the warp that took it.
#pragma once
#include "ld_kernels.cuh"

namespace ldb200 {

#ifndef LDB200_RG_THREADS
#define LDB200_RG_THREADS 640
#endif
constexpr int RG_THREADS = LDB200_RG_THREADS;
constexpr int RG_WARPS = RG_THREADS / 32;
constexpr int RG_MAX_ROWS = 8;

struct RigidComplex {
  int n_groups, n_rec_pos;  // n_rec_pos = n_groups * 32 (type-grouped receptor positions, pads interspersed)
  int n_lig, n_lig_pad, n_lig_tiles;
  int n_rec_modes, pose_len;
  int rows_max;
  int row_bytes;                        // one table row: FLEX n_lig_types x RG_TB_BYTES (only the ligand's own types), rigid RG_ROW_BYTES
  const double *rec_x, *rec_y, *rec_z;  // [n_rec_pos] lab frame, pads at REC_PAD
  const int *rec_slot;                  // [n_rec_pos] which of the group's rows this atom indexes (0..3)
  const int *rec_toff;                  // [n_rec_pos] type * 3380 (exact path)
  const int *group_types;               // [n_groups][RG_MAX_ROWS] DFIRE type of each row, -1 = unused
  const int *group_order;               // groups, most expensive first
  const double *rec_modes;              // [k][3][n_rec_pos]
  const float4 *lig4;                   // [n_lig_pad] local f32 x,y,z + (float)(type * RG_SLOTS)
  const double *lig_x, *lig_y, *lig_z;  // [n_lig_pad] local f64 (exact path)
  const unsigned short *lig_tb20;       // type * 20 (exact path)
  const double *potx;                   // [169][row_bytes/8]: row ta = [compact ligand type][slot], slot s <-> index s-1
  const double *pot;                    // the reference's table (exact path)
  float gx0, gy0, gz0, inv_h;           // cell grid in the ligand frame
  int nx, ny, nz;
  const uint2 *cells;                   // [nz][ny][nx] {offset into cell_tiles, count}
  const unsigned short *cell_tiles;
  float thr_out;                        // 225 + delta
  float half_minus_eps;                 // 0.5 - 2.5e-5 (the f32 evaluation error of t)
  float delta;                          // 1.02 * delta: eps_t = 1.02 * delta / sqrt(d2f) + 2.5e-5
  // FLEX (ligand with ANM modes)
  int flex;                             // 1: per-pose ligand blocks + slack lists + fixed-point sums
  int n_lig_modes;
  const double *lig_modes;              // [k][3][n_lig_pad] (exact path)
  const long long *potx_fx;             // potx * fx_scale rounded to nearest: same layout
  double fx_scale;                      // 2^k
  const float *tile_slack;              // [n_lig_tiles] slack the current lists were built with
  float grid_maxabs;                    // largest |coordinate| inside the cell grid (the M of the FP32 error bound)
  // Who reads the interface flags (src/dfire.rs:339-342): finalize_kernel looks at the receptor atoms of the ACTIVE
  // restraints and the membrane beads, and at the ligand atoms of the active ligand restraints (src/dfire.rs:351-357,
  // src/scoring.rs:21-47); nothing else of the flag arrays can reach the energy.
  const unsigned *group_need;           // [n_groups] lanes whose receptor atom is in an active restraint or a bead
  int lig_need;                         // the ligand has active restraints: every contact's ligand flag matters
};

// ligand copies: 1 (rigid: shared by the CTA) or one per warp (FLEX: each warp works on its own pose)
__host__ __device__ inline size_t rigid_smem_bytes(int n_lig_pad, int rows, int row_bytes, int lig_copies = 1) {
  return 128 + (size_t)lig_copies * n_lig_pad * 16 + (size_t)rows * row_bytes;
}

extern __shared__ __align__(128) unsigned char smem_rigid[];
constexpr uint32_t RG_NEED_OFF = 16;  // header: mbarrier (8 B), unit, next pose, then the current group's interface-need mask
// Loads at byte offsets of the kernel's dynamic shared memory (plain C++ so the scheduler may interleave the
// eight pairs of an item; the array is known to live in shared memory, so these are LDS with 32-bit addresses).
__device__ __forceinline__ double lds_f64(uint32_t off) { return *reinterpret_cast<const double *>(smem_rigid + off); }
__device__ __forceinline__ long long lds_i64(uint32_t off) { return *reinterpret_cast<const long long *>(smem_rigid + off); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte asynchronous global -> shared copy (LDGSTS) to a byte offset of the kernel's dynamic shared memory
__device__ __forceinline__ void cp_async16(uint32_t smem_off, const void *gmem) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_rigid + smem_off);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Per-pose quantities shared by the 100+ (group, pose) work items of a pose, computed once per batch:
//   prep[0..8]  M = R^T (row major), R the rotation of q: q v q^-1 is a rotation for any q != 0 (src/qt.rs:57-61)
//   prep[9..11] M t          -> ligand-frame position of a lab-frame point r is  M r - M t
//   prep[12..15] q^-1 = conj(q)/norm2(q) exactly as Quaternion::inverse computes it (src/qt.rs:24-34,48-50),
//               for the exact path.
__global__ void __launch_bounds__(256)
    rigid_prep_kernel(const BatchBuffers bb, const double *poses, int n_poses, int pose_len, double *prep) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= live_poses(bb, n_poses)) return;
  const double *pose = poses + (size_t)p * pose_len;
  const double tx = pose[0], ty = pose[1], tz = pose[2];
  const double qw = pose[3], qx = pose[4], qy = pose[5], qz = pose[6];
  const double n2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(qw, qw), __dmul_rn(qx, qx)), __dmul_rn(qy, qy)),
                              __dmul_rn(qz, qz));
  const double s2 = 2.0 / n2;
  const double xx = qx * qx * s2, yy = qy * qy * s2, zz = qz * qz * s2;
  const double xy = qx * qy * s2, xz = qx * qz * s2, yz = qy * qz * s2;
  const double wx = qw * qx * s2, wy = qw * qy * s2, wz = qw * qz * s2;
  // columns of R = rows of R^T
  const double m[9] = {1.0 - (yy + zz), xy + wz, xz - wy, xy - wz, 1.0 - (xx + zz), yz + wx, xz + wy, yz - wx,
                       1.0 - (xx + yy)};
  double *o = prep + (size_t)p * RG_PREP;
#pragma unroll
  for (int i = 0; i < 9; ++i) o[i] = m[i];
  o[9] = fma(m[0], tx, fma(m[1], ty, m[2] * tz));
  o[10] = fma(m[3], tx, fma(m[4], ty, m[5] * tz));
  o[11] = fma(m[6], tx, fma(m[7], ty, m[8] * tz));
  o[12] = __ddiv_rn(qw, n2); o[13] = __ddiv_rn(-qx, n2); o[14] = __ddiv_rn(-qy, n2); o[15] = __ddiv_rn(-qz, n2);
}

// FLEX per-pose preparation: FLEX_PP poses per CTA so that a ligand atom's mode vectors are read once for all of them
// (one pose per CTA re-read every mode of the ligand per pose from L2: that is what bounds transform_kernel).
//   prep[p]        as rigid_prep_kernel
//   lig4p[p][j]    ligand-frame position l_j + R^T D_j, D_j = sum_k mode_k[j] * extent_k (src/dfire.rs:290-301 gives
//                  the lab-frame R l_j + t + D_j; the same point seen from the ligand's frame), rounded to f32, with
//                  the table column offset in .w (copied from the static lig4)
//   pose_flag[p]   0 if every ligand tile stayed within the slack its cell lists were built with; otherwise the
//                  pose's largest tile displacement (> 0): the pair kernel then scores the pose against every ligand
//                  tile and widens its FP32 margins to the larger coordinates that can be in range
//   need[t]        running maximum (over every pose ever prepared) of tile t's displacement, for the host to grow
//                  the slacks; non-negative floats compared as integers
// The displacements only feed the f32 coordinates and the (conservative, inflated) slack test; every decision that
// could depend on their last bits is re-derived by rigid_exact_pair with the reference's own arithmetic.
constexpr int FLEX_PP = 8;
constexpr int FLEX_THREADS = 128;
__global__ void __launch_bounds__(FLEX_THREADS)
    flex_prep_kernel(const RigidComplex rc, const BatchBuffers bb, const double *__restrict__ poses, int n_poses,
                     double *__restrict__ prep, float4 *__restrict__ lig4p, float *__restrict__ pose_flag,
                     int *__restrict__ need) {
  n_poses = live_poses(bb, n_poses);
  extern __shared__ __align__(16) unsigned char smem_flex[];
  double *sM = reinterpret_cast<double *>(smem_flex);                       // [FLEX_PP][16]
  double *sE = sM + FLEX_PP * RG_PREP;                                      // [FLEX_PP][n_lig_modes]
  int *sD = reinterpret_cast<int *>(sE + FLEX_PP * max(rc.n_lig_modes, 1)); // [FLEX_PP][n_lig_tiles] max |D|^2 (float bits)
  __shared__ int s_flag[FLEX_PP], s_dmax[FLEX_PP];
  const int p0 = blockIdx.x * FLEX_PP, np = min(FLEX_PP, n_poses - p0);
  if (np <= 0) return;  // CTA-uniform
  const int tid = threadIdx.x;
  if (tid < np) {
    const double *pose = poses + (size_t)(p0 + tid) * rc.pose_len;
    const double tx = pose[0], ty = pose[1], tz = pose[2];
    const double qw = pose[3], qx = pose[4], qy = pose[5], qz = pose[6];
    const double n2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(qw, qw), __dmul_rn(qx, qx)), __dmul_rn(qy, qy)),
                                __dmul_rn(qz, qz));
    const double s2 = 2.0 / n2;
    const double xx = qx * qx * s2, yy = qy * qy * s2, zz = qz * qz * s2;
    const double xy = qx * qy * s2, xz = qx * qz * s2, yz = qy * qz * s2;
    const double wx = qw * qx * s2, wy = qw * qy * s2, wz = qw * qz * s2;
    const double m[9] = {1.0 - (yy + zz), xy + wz, xz - wy, xy - wz, 1.0 - (xx + zz), yz + wx, xz + wy, yz - wx,
                         1.0 - (xx + yy)};
    double *o = sM + tid * RG_PREP;
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i] = m[i];
    o[9] = fma(m[0], tx, fma(m[1], ty, m[2] * tz));
    o[10] = fma(m[3], tx, fma(m[4], ty, m[5] * tz));
    o[11] = fma(m[6], tx, fma(m[7], ty, m[8] * tz));
    o[12] = __ddiv_rn(qw, n2); o[13] = __ddiv_rn(-qx, n2); o[14] = __ddiv_rn(-qy, n2); o[15] = __ddiv_rn(-qz, n2);
    double *g = prep + (size_t)(p0 + tid) * RG_PREP;
#pragma unroll
    for (int i = 0; i < RG_PREP; ++i) g[i] = o[i];
    s_flag[tid] = 0;
    s_dmax[tid] = 0;
  }
  for (int i = tid; i < np * rc.n_lig_modes; i += FLEX_THREADS) {
    const int p = i / rc.n_lig_modes, k = i % rc.n_lig_modes;
    sE[p * rc.n_lig_modes + k] = poses[(size_t)(p0 + p) * rc.pose_len + 7 + rc.n_rec_modes + k];
  }
  for (int i = tid; i < FLEX_PP * rc.n_lig_tiles; i += FLEX_THREADS) sD[i] = 0;
  __syncthreads();
  for (int j = tid; j < rc.n_lig_pad; j += FLEX_THREADS) {
    const float4 st = rc.lig4[j];  // static block: pads at 1e6, .w = (float)(type * RG_SLOTS)
    if (j >= rc.n_lig) {
      for (int p = 0; p < np; ++p) lig4p[(size_t)(p0 + p) * rc.n_lig_pad + j] = st;
      continue;
    }
    double dx[FLEX_PP], dy[FLEX_PP], dz[FLEX_PP];
#pragma unroll
    for (int p = 0; p < FLEX_PP; ++p) dx[p] = dy[p] = dz[p] = 0.0;
    for (int k = 0; k < rc.n_lig_modes; ++k) {
      const double *m = rc.lig_modes + (size_t)k * 3 * rc.n_lig_pad;
      const double mx = m[j], my = m[rc.n_lig_pad + j], mz = m[2 * rc.n_lig_pad + j];
#pragma unroll
      for (int p = 0; p < FLEX_PP; ++p) {
        const double e = sE[p * rc.n_lig_modes + k];
        dx[p] = fma(mx, e, dx[p]); dy[p] = fma(my, e, dy[p]); dz[p] = fma(mz, e, dz[p]);
      }
    }
    const double lx = rc.lig_x[j], ly = rc.lig_y[j], lz = rc.lig_z[j];
#pragma unroll
    for (int p = 0; p < FLEX_PP; ++p) {
      if (p >= np) break;
      const double *M = sM + p * RG_PREP;
      const double x = lx + fma(M[0], dx[p], fma(M[1], dy[p], M[2] * dz[p]));
      const double y = ly + fma(M[3], dx[p], fma(M[4], dy[p], M[5] * dz[p]));
      const double z = lz + fma(M[6], dx[p], fma(M[7], dy[p], M[8] * dz[p]));
      lig4p[(size_t)(p0 + p) * rc.n_lig_pad + j] = make_float4((float)x, (float)y, (float)z, st.w);
      // |R^T D| = |D|; rounded up so the f32 value never understates it
      const float d2 = __double2float_ru(fma(dz[p], dz[p], fma(dy[p], dy[p], dx[p] * dx[p])));
      atomicMax(&sD[p * rc.n_lig_tiles + j / LIG_TILE], __float_as_int(d2));
    }
  }
  __syncthreads();
  for (int i = tid; i < np * rc.n_lig_tiles; i += FLEX_THREADS) {
    const int p = i / rc.n_lig_tiles, t = i % rc.n_lig_tiles;
    // displacement of the tile, inflated (f32 square root rounded up, plus the f32 rounding of the coordinates)
    const float d = __fsqrt_ru(__int_as_float(sD[p * rc.n_lig_tiles + t])) * 1.000001f + 1.0e-5f;
    if (d > rc.tile_slack[t]) s_flag[p] = 1;
    atomicMax(&s_dmax[p], __float_as_int(d));
    if (__float_as_int(d) > need[t]) atomicMax(&need[t], __float_as_int(d));
  }
  __syncthreads();
  if (tid < np) pose_flag[p0 + tid] = s_flag[tid] ? __int_as_float(s_dmax[tid]) : 0.f;
}
__host__ __device__ inline size_t flex_prep_smem(int n_lig_modes, int n_lig_tiles) {
  return (size_t)FLEX_PP * RG_PREP * 8 + (size_t)FLEX_PP * (n_lig_modes > 1 ? n_lig_modes : 1) * 8 +
         (size_t)FLEX_PP * n_lig_tiles * 4;
}

// The reference's arithmetic for ONE pair, lab frame, never fused: src/qt.rs:57-61,174-185 (rotate),
// src/dfire.rs:286-288 (translate), :304-320 (receptor ANM), :331-342 (distance, bin, interface).
// Returns -1 if the pair is outside the cut-off, else bin | (interface ? 32 : 0).
__device__ __noinline__ int rigid_exact_pair(const RigidComplex *__restrict__ rcp, const double *pose,
                                             const double *prep, int ia, int j) {
  const RigidComplex &rc = *rcp;  // the copy in global memory: this path is rare and must not pin registers
  const double tx = pose[0], ty = pose[1], tz = pose[2];
  const Quat q = {pose[3], pose[4], pose[5], pose[6]};
  const Quat qi = {prep[12], prep[13], prep[14], prep[15]};
  const Quat v = {0.0, rc.lig_x[j], rc.lig_y[j], rc.lig_z[j]};
  const Quat r = qmul(qmul(q, v), qi);
  double lx = __dadd_rn(r.x, tx), ly = __dadd_rn(r.y, ty), lz = __dadd_rn(r.z, tz);
  for (int k = 0; k < rc.n_lig_modes; ++k) {  // src/dfire.rs:290-301
    const double e = pose[7 + rc.n_rec_modes + k];
    const double *m = rc.lig_modes + (size_t)k * 3 * rc.n_lig_pad;
    lx = __dadd_rn(lx, __dmul_rn(m[j], e));
    ly = __dadd_rn(ly, __dmul_rn(m[rc.n_lig_pad + j], e));
    lz = __dadd_rn(lz, __dmul_rn(m[2 * rc.n_lig_pad + j], e));
  }
  double x = rc.rec_x[ia], y = rc.rec_y[ia], z = rc.rec_z[ia];
  for (int k = 0; k < rc.n_rec_modes; ++k) {
    const double e = pose[7 + k];
    const double *m = rc.rec_modes + (size_t)k * 3 * rc.n_rec_pos;
    x = __dadd_rn(x, __dmul_rn(m[ia], e));
    y = __dadd_rn(y, __dmul_rn(m[rc.n_rec_pos + ia], e));
    z = __dadd_rn(z, __dmul_rn(m[2 * rc.n_rec_pos + ia], e));
  }
  const double ex = __dsub_rn(x, lx), ey = __dsub_rn(y, ly), ez = __dsub_rn(z, lz);
  const double dist = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
  if (!(dist <= 225.0)) return -1;
  const double d = __dsub_rn(__dmul_rn(__dsqrt_rn(dist), 2.0), 1.0);
  const int bin = dfire_bin_of((int)d);
  return bin | (d <= 3.9 ? 32 : 0);
}

// One row of <= 32 work items; an item is (receptor atom = owner lane, ligand tile of 8 atoms).
//
// FP32 classification.  With a = local receptor atom (f64 product rounded to f32), l = local ligand atom
// (f64 rounded to f32), M = largest |coordinate| inside the cell grid:
//   |d2f - dist_ref| <= 2*sqrt(3)*15.1*(2^-23 M + 2^-24*15.1) + 3*2^-24*228 + f64 noise  <  6.3e-6 M + 9e-5,
// delta = 1.3e-5 M + 2e-4 (2x that bound).  Bin-space value t = 2 sqrt(dist) - 1 (src/dfire.rs:336):
//   u = 2*d2f*rsqrt.approx(d2f) - 1.5 = t_f32 - 0.5,
//   |t_f32 - t_ref| <= 2|sqrt(d2f) - sqrt(dist_ref)| + 6.4e-6 <= delta/sqrt(min(d2f, dist_ref)) + 6.4e-6
//                   <= 1.001*delta*rsqrt(d2f) + 6.4e-6 =: E_t            (d2f >= 6.25, delta < 0.01),
//   m = u + MAGIC  -> rint(u) in the low mantissa bits,  g = u - rint(u) = frac(t_f32) - 0.5 (exact).
// If |g| <= 0.5 - eps_t (eps_t = 1.02*delta*rsqrt.approx(d2f) + 2.5e-5 > E_t) then floor(t_ref) = rint(u): the bin
// is the reference's.
// d2f <= 225 + delta together with that leaves indices 0..28 only (index 29 needs t >= 29 + eps_t, i.e.
// d2f > 225 + delta); such a pair's table value is added in the hot loop, every other pair with
// d2f <= 225 + delta goes to rigid_exact_pair.  The interface test t <= 3.9 (dist <= 6.0025) is kept out of
// the hot loop: a per-item min(d2f) sends the rare items with a contact below 2.45 A + to a second pass
// that decides it in FP32 outside 6.0025 +- delta and with rigid_exact_pair inside.
__device__ __forceinline__ float4 lds_f4(uint32_t off) { return *reinterpret_cast<const float4 *>(smem_rigid + off); }
// The hot loop's loads take ABSOLUTE shared-window addresses (base of smem_rigid folded into the per-lane tile / row
// address once): with `smem_rigid + off` the compiler adds the window base to every address (one IADD per LDS, two
// LDS per pair).  Deliberately not volatile and without a memory clobber, so that the eight pairs of an item are
// still interleaved freely; what orders them after the copies that fill the memory they read is rg_after_copy() on
// the address they are computed from.
__device__ __forceinline__ uint32_t rg_smem_base() { return (uint32_t)__cvta_generic_to_shared(smem_rigid); }
__device__ __forceinline__ float4 lds_f4_abs(uint32_t a) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds_f64_abs(uint32_t a) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
// an address "produced" after the wait that made the data behind it visible: no load computed from it can be moved
// above that wait by the compiler
__device__ __forceinline__ void rg_after_copy(uint32_t &addr) { asm volatile("" : "+r"(addr) : : "memory"); }

// Accumulator of the table values: f64 (rigid ligand: lists are static, so the order of a pose's additions is a
// function of the pose alone) or 64-bit fixed point (FLEX: exact whatever the order).
template <bool FLEX> struct RgAcc { typedef double type; };
template <> struct RgAcc<true> { typedef long long type; };
__device__ __forceinline__ void rg_add(double &acc, uint32_t addr) { acc = __dadd_rn(acc, lds_f64_abs(addr)); }
__device__ __forceinline__ void rg_add(long long &acc, uint32_t addr) { acc += lds_i64(addr); }  // FLEX: relative
__device__ __forceinline__ double rg_join(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ long long rg_join(long long a, long long b) { return a + b; }
__device__ __forceinline__ void rg_add_exact(double &acc, double v, double) { acc = __dadd_rn(acc, v); }
__device__ __forceinline__ void rg_add_exact(long long &acc, double v, double scale) { acc += __double2ll_rn(v * scale); }

template <bool DETAIL, bool FLEX>
__device__ __forceinline__ void rigid_row(const RigidComplex &rc, const BatchBuffers &bb, uint32_t l4_addr,
                                          uint32_t lane_sw, bool active, int o, int lt, float rxf, float ryf,
                                          float rzf, unsigned rowoff, int p, int pos_base,
                                          typename RgAcc<FLEX>::type &acc0, typename RgAcc<FLEX>::type &acc1,
                                          unsigned &ifr_mask, const RigidComplex *rc_dev, const double *prep,
                                          float thr_out, float hme, float delta) {
  ld_pose_detail *dt = DETAIL ? reinterpret_cast<ld_pose_detail *>(bb.detail) + p : nullptr;
  const int lane = threadIdx.x & 31;
  const float ax = __shfl_sync(0xffffffffu, rxf, o), ay = __shfl_sync(0xffffffffu, ryf, o),
              az = __shfl_sync(0xffffffffu, rzf, o);
  const unsigned rb = __shfl_sync(0xffffffffu, rowoff, o);
  if (!active) return;
  const int jbase = lt * LIG_TILE;
  // per-lane slot: atom (k ^ lane) & 7 of the tile, so that any 8 consecutive lanes read 8 different 16-byte slots
  // -> conflict-free LDS.128 whatever tiles the lanes hold
  // tile base (128-byte aligned) | the lane's slot bits: slot k of this lane is one XOR away
  const uint32_t tile_addr = (l4_addr + (uint32_t)lt * (LIG_TILE * 16)) | lane_sw;
  unsigned slow_bits = 0u;
  unsigned n_fast = 0;
  float mind2 = 3.0e38f;
#pragma unroll
  for (int k = 0; k < LIG_TILE; ++k) {
    const float4 a = FLEX ? lds_f4(tile_addr ^ (uint32_t)(k << 4)) : lds_f4_abs(tile_addr ^ (uint32_t)(k << 4));
    const float dx = ax - a.x, dy = ay - a.y, dz = az - a.z;
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    mind2 = fminf(mind2, d2);
    const float rs = rsqrt_approx(d2);
    const float d = d2 * rs;
    const float u = fmaf(d, 2.0f, -1.5f);  // t - 0.5 in [-1.5, 28.5]: rint(u) = -1 for t < 0, see RG_SLOT0
    const float m = __fadd_rn(u, RG_MAGIC);
    const float g = __fsub_rn(u, __fsub_rn(m, RG_MAGIC));
    const float marg = fmaf(delta, rs, fabsf(g));  // |g| + delta/d
    const bool inr = d2 <= thr_out;
    const bool fast = inr & (marg <= hme);  // |g| + delta/d <= 0.5 - 2.5e-5
    if (inr & !fast) slow_bits |= 1u << k;
    // a.w = ligand type * RG_SLOTS as a float: adding it to m (both integers < 2^24) is exact and leaves
    // MAGIC_BITS + type*RG_SLOTS + index in the mantissa -> one shift-add gives the byte address
    const uint32_t addr = ((uint32_t)__float_as_int(__fadd_rn(m, a.w)) << 3) + rb;
    if (fast) {
      if (k & 1) rg_add(acc1, addr);
      else rg_add(acc0, addr);
      if (DETAIL) {
        ++n_fast;
        atomicAdd(reinterpret_cast<unsigned long long *>(&dt->bin_hist[dfire_bin_fast(__float_as_int(m) - (int)RG_MAGIC_BITS)]),
                  1ull);
      }
    }
  }
  // rare: a contact near or below the 2.45 A interface edge (src/dfire.rs:339-342) -- looked at only where a flag it
  // could set is read by someone (RG_NEED_OFF: the group's mask of such lanes; all ones with DETAIL or ligand restraints)
  if (mind2 <= 6.0025f + delta && (FLEX || ((*reinterpret_cast<const unsigned *>(smem_rigid + RG_NEED_OFF) >> o) & 1u))) {
    const double *pose = bb.poses + (size_t)p * rc_dev->pose_len;
    for (int k = 0; k < LIG_TILE; ++k) {
      if ((slow_bits >> k) & 1u) continue;  // the exact path below owns this pair entirely
      // (a plain load at the relative offset: an asm load of the same address would be merged with the one above and
      // keep the eight atoms of the item in registers across the hot loop)
      const float4 a = lds_f4((tile_addr - (FLEX ? 0u : rg_smem_base())) ^ (uint32_t)(k << 4));
      const float dx = ax - a.x, dy = ay - a.y, dz = az - a.z;
      const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));  // same operations as above: same bits
      if (d2 > 6.0025f + delta) continue;
      const int j = jbase + ((k ^ lane) & (LIG_TILE - 1));
      bool ifc = d2 < 6.0025f - delta;  // t <= 3.9 <=> dist <= 6.0025, decided in FP32 outside +-delta
      if (!ifc) ifc = (rigid_exact_pair(rc_dev, pose, prep, pos_base + o, j) & ~31) == 32;  // -1 -> all bits set -> false
      if (ifc) {
        ifr_mask |= 1u << o;
        atomicOr(&bb.iface_lig[(size_t)p * bb.lig_words + (j >> 5)], 1u << (j & 31));
        if (DETAIL) atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_interface_pairs), 1ull);
      }
    }
  }
  if (DETAIL) {
    atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_pairs_tested),
              (unsigned long long)min(LIG_TILE, rc.n_lig - jbase));
    if (n_fast) atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_in_cutoff), (unsigned long long)n_fast);
  }
  if (slow_bits) {  // rare: near a decision threshold, or closer than 2.5 A
    typename RgAcc<FLEX>::type extra = 0;
    const double fx_scale = FLEX ? rc_dev->fx_scale : 0.0;
    const int toff = rc_dev->rec_toff[pos_base + o];
    const double *pose = bb.poses + (size_t)p * rc_dev->pose_len;
    for (unsigned b = slow_bits; b; b &= b - 1) {
      const int j = jbase + (((__ffs(b) - 1) ^ lane) & (LIG_TILE - 1));
      const int r = rigid_exact_pair(rc_dev, pose, prep, pos_base + o, j);
      if (DETAIL) atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_exact_fallback), 1ull);
      if (r >= 0) {
        rg_add_exact(extra, __ldg(rc_dev->pot + toff + rc_dev->lig_tb20[j] + (r & 31)), fx_scale);
        if (r & 32) {
          ifr_mask |= 1u << o;
          atomicOr(&bb.iface_lig[(size_t)p * bb.lig_words + (j >> 5)], 1u << (j & 31));
          if (DETAIL) atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_interface_pairs), 1ull);
        }
        if (DETAIL) {
          atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_in_cutoff), 1ull);
          atomicAdd(reinterpret_cast<unsigned long long *>(&dt->bin_hist[r & 31]), 1ull);
        }
      }
    }
    if (FLEX) acc0 += extra;
    else rg_add_exact(acc0, (double)extra, 0.0);
  }
}

// Sum of a lane value over the warp in a fixed order (f64) / exactly (fixed point).
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// One (receptor group, pose): the group's 32 atoms (one per lane) against the ligand of pose p.  Adds the table values
// to acc0/acc1 (per-lane partial sums, fixed order) and ORs the lanes' interface flags into ifr_mask.
//   rowoff   per lane: shared-memory byte address of the lane's table row minus what the magic-number index carries;
//            0xffffffff = pad lane
//   prep     the pose's rotation data (rigid_prep_kernel)
//   brute..reach_abs  FLEX only: the pose left its slacks -> every ligand tile is a candidate, with margins widened to
//            the coordinates that can then be in range
template <bool DETAIL, bool FLEX>
__device__ __forceinline__ void rg_score_group(const RigidComplex &rc, const BatchBuffers &bb, uint32_t l4_addr,
                                               uint32_t lane_sw, unsigned rowoff, int p, int pos_base,
                                               const double *prep, bool brute, float thr_out, float hme,
                                               float delta102, float reach_abs, typename RgAcc<FLEX>::type &acc0,
                                               typename RgAcc<FLEX>::type &acc1, unsigned &ifr_mask,
                                               const RigidComplex *rc_dev) {
  const int lane = threadIdx.x & 31;
  float fx, fy, fz;
  unsigned my_off;
  int my_n;
  {
    // receptor atom -> ligand frame: M r - M t (rigid_prep_kernel)
    double ax = rc.rec_x[pos_base + lane], ay = rc.rec_y[pos_base + lane], az = rc.rec_z[pos_base + lane];
    if (rc.n_rec_modes > 0) {  // src/dfire.rs:304-320 (lab frame)
      const double *pose = bb.poses + (size_t)p * rc.pose_len;
      if (FLEX) {
        // a FLEX task is latency bound (ncu: long_scoreboard 2.5 per issued instruction): keep the loads of five modes
        // in flight; the additions stay in k order.  Only in this instantiation: the same pragma in the rigid one
        // changes nothing it executes on 1k4c / 1ppe (no receptor modes) and still costs it 2.4 % (code generation).
#pragma unroll 5
        for (int k = 0; k < rc.n_rec_modes; ++k) {
          const double e = pose[7 + k];
          const double *m = rc.rec_modes + (size_t)k * 3 * rc.n_rec_pos + pos_base + lane;
          ax = __dadd_rn(ax, __dmul_rn(m[0], e));
          ay = __dadd_rn(ay, __dmul_rn(m[rc.n_rec_pos], e));
          az = __dadd_rn(az, __dmul_rn(m[2 * rc.n_rec_pos], e));
        }
      } else {
        for (int k = 0; k < rc.n_rec_modes; ++k) {
          const double e = pose[7 + k];
          const double *m = rc.rec_modes + (size_t)k * 3 * rc.n_rec_pos + pos_base + lane;
          ax = __dadd_rn(ax, __dmul_rn(m[0], e));
          ay = __dadd_rn(ay, __dmul_rn(m[rc.n_rec_pos], e));
          az = __dadd_rn(az, __dmul_rn(m[2 * rc.n_rec_pos], e));
        }
      }
    }
    fx = (float)(fma(prep[0], ax, fma(prep[1], ay, fma(prep[2], az, -prep[9]))));
    fy = (float)(fma(prep[3], ax, fma(prep[4], ay, fma(prep[5], az, -prep[10]))));
    fz = (float)(fma(prep[6], ax, fma(prep[7], ay, fma(prep[8], az, -prep[11]))));
    const int cxi = __float2int_rd((fx - rc.gx0) * rc.inv_h), cyi = __float2int_rd((fy - rc.gy0) * rc.inv_h),
              czi = __float2int_rd((fz - rc.gz0) * rc.inv_h);
    uint2 ce = make_uint2(0u, 0u);
    if ((unsigned)cxi < (unsigned)rc.nx && (unsigned)cyi < (unsigned)rc.ny && (unsigned)czi < (unsigned)rc.nz)
      ce = __ldg(rc.cells + ((size_t)czi * rc.ny + cyi) * rc.nx + cxi);
    my_off = ce.x;
    my_n = (int)rowoff == -1 ? 0 : (int)ce.y;  // pad lanes own nothing
    if (FLEX && brute) {  // every ligand tile, whatever the lists say: entry k of the "list" is tile k
      my_off = 0u;
      // an atom further out than the grid grown by dmax has no partner within 15 A
      const bool far = fmaxf(fabsf(fx), fmaxf(fabsf(fy), fabsf(fz))) > reach_abs;
      my_n = ((int)rowoff == -1 || far || !(fx == fx)) ? 0 : rc.n_lig_tiles;
    }
  }
  auto tile_at = [&](unsigned idx) -> unsigned {
    return (FLEX && brute) ? idx : (unsigned)__ldg(rc.cell_tiles + idx);
  };

  // Work items = (owner lane, ligand tile) for every entry of the 32 lanes' cell lists, scored 32 at a time.
  //  (1) an owner with >= 32 entries fills whole rows on its own: lane i takes entry r*32 + i of its list;
  //  (2) the tails (n mod 32 entries per owner) are pooled: an inclusive scan of the tail lengths numbers
  //      them, item k belongs to the lane o = #{l : end_l <= k} (5-step binary search over the ends with
  //      register shuffles) and is entry k - end_{o-1} of o's tail.
  // The next row's (owner, tile) is produced before the current row is scored, so its list read is in flight.
  unsigned big = __ballot_sync(0xffffffffu, my_n >= 32);
  int big_o = 0, big_left = 0;       // current whole-row owner, entries left in whole rows
  unsigned big_off = 0u;
  const int tail_n = my_n & 31;
  const unsigned tail_off = my_off + (unsigned)(my_n & ~31);
  int end = tail_n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, end, d);
    if (lane >= d) end += t;
  }
  const int total = __shfl_sync(0xffffffffu, end, 31);
  int k0 = 0;
  bool act = false, act_nxt = false;
  int o = 0, o_nxt = 0;
  unsigned lt = 0u, lt_nxt = 0u;
  auto produce = [&]() -> bool {  // warp-uniform control flow; fills (act_nxt, o_nxt, lt_nxt)
    if (big_left == 0 && big != 0u) {
      big_o = __ffs(big) - 1;
      big &= big - 1;
      big_left = __shfl_sync(0xffffffffu, my_n, big_o) & ~31;
      big_off = __shfl_sync(0xffffffffu, my_off, big_o);
    }
    if (big_left > 0) {
      o_nxt = big_o;
      lt_nxt = tile_at(big_off + lane);
      act_nxt = true;
      big_off += 32u;
      big_left -= 32;
      return true;
    }
    if (k0 < total) {
      const int k = k0 + lane;
      int oo = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int v = __shfl_sync(0xffffffffu, end, oo + step - 1);
        if (v <= k) oo += step;
      }
      const int prev = __shfl_sync(0xffffffffu, end, (oo - 1) & 31);
      const unsigned off = __shfl_sync(0xffffffffu, tail_off, oo);
      o_nxt = oo;
      act_nxt = k < total;
      lt_nxt = 0u;
      if (act_nxt) lt_nxt = tile_at(off + (unsigned)(k - (oo > 0 ? prev : 0)));
      k0 += 32;
      return true;
    }
    return false;
  };
  bool have = produce();
  if (FLEX) {  // the pose's ligand block (cp.async copies issued by the caller) has to be in this warp's slice now
    cp_async_wait_all();
    __syncwarp();
  }
  while (have) {
    act = act_nxt; o = o_nxt; lt = lt_nxt;
    have = produce();
    rigid_row<DETAIL, FLEX>(rc, bb, l4_addr, lane_sw, act, o, (int)lt, fx, fy, fz, rowoff, p, pos_base, acc0, acc1,
                            ifr_mask, rc_dev, prep, thr_out, hme, delta102);
  }}

// lig4p / pose_flag: FLEX only -- per-pose ligand blocks [n_poses][n_lig_pad] and 1 = "a tile moved further than its
// slack: score this pose against every ligand tile" (both written by flex_prep_kernel).
template <bool DETAIL, bool FLEX>
__global__ void __launch_bounds__(RG_THREADS, 1)
    dfire_rigid_kernel(const RigidComplex rc, const BatchBuffers bb, int n_poses, int poses_per_unit, int n_chunks,
                       unsigned *unit_counter, const RigidComplex *rc_dev, const double *prep_all,
                       const float4 *__restrict__ lig4p, const float *__restrict__ pose_flag) {
  typedef typename RgAcc<FLEX>::type acc_t;
  n_poses = live_poses(bb, n_poses);
  unsigned char *smem_raw = smem_rigid;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  int *s_unit = reinterpret_cast<int *>(smem_raw + 8);
  int *s_pose_next = reinterpret_cast<int *>(smem_raw + 12);
  const int lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5;
  // rigid: one ligand block for the CTA; FLEX: one per warp (the pose it is working on)
  const uint32_t l4_off = 128u + (FLEX ? (uint32_t)(threadIdx.x >> 5) * (uint32_t)rc.n_lig_pad * 16u : 0u);
  // rigid: absolute (shared window) addresses, see lds_f4_abs; the FLEX instance keeps offsets relative to smem_rigid
  // (with absolute ones ptxas spills in its row loop: 2uuy 3.94 -> 4.45 ms)
  const uint32_t abs_base = FLEX ? 0u : rg_smem_base();
  uint32_t l4_addr = abs_base + l4_off;
  float4 *l4 = reinterpret_cast<float4 *>(smem_raw + l4_off);
  unsigned char *rows = smem_raw + 128 + (size_t)(FLEX ? n_warps : 1) * rc.n_lig_pad * 16;
  const uint32_t lane_sw = (uint32_t)(lane & 7) << 4;
  const int n_units = rc.n_groups * n_chunks;
  // rigid: full rows, a compile-time size (a run-time one costs the rigid instance 1 % in code generation); FLEX: rows of the
  // ligand's own types only
  const uint32_t row_bytes = FLEX ? (uint32_t)rc.row_bytes : (uint32_t)RG_ROW_BYTES;
  const unsigned char *rows_src = FLEX ? reinterpret_cast<const unsigned char *>(rc.potx_fx)
                                       : reinterpret_cast<const unsigned char *>(rc.potx);

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  bool lig_loaded = FLEX;  // FLEX: nothing static to load, every warp stages its pose's block itself
  int cur_g = -1;
  unsigned rowoff = 0u;

  for (;;) {
    __syncthreads();  // every warp is done with the previous unit: rows and counters may be rewritten
    if (threadIdx.x == 0) {
      const unsigned un = atomicAdd(unit_counter, 1u);
      *s_unit = (int)un;
      *s_pose_next = 0;
      if (!FLEX)  // (the FLEX instance looks at every contact: its code generation does not take the extra test well)
        *reinterpret_cast<unsigned *>(smem_raw + RG_NEED_OFF) =
            (DETAIL || rc.lig_need || un >= (unsigned)n_units) ? 0xffffffffu : rc.group_need[rc.group_order[un / (unsigned)n_chunks]];
    }
    __syncthreads();
    const int u = *s_unit;
    if (u >= n_units) break;
    const int g = rc.group_order[u / n_chunks];
    // rigid: the unit's poses are p0, p0 + n_chunks, p0 + 2 n_chunks, ...  A batch is a sequence of swarms (200 similar
    // poses each) whose cost differs by up to 5x, and a strided unit holds the batch's mix of them instead of two or
    // three swarms: a rank's 50 swarms of the 1k4c bench 3.73 -> 3.65 ms (profiles/r2_rigid_ab_run4_units.txt).  FLEX
    // keeps contiguous ranges: its per-pose ligand blocks stream better that way (2uuy 3.94 against 3.99 ms).
    const int pstep = FLEX ? 1 : n_chunks;
    const int p0 = FLEX ? (u % n_chunks) * poses_per_unit : u % n_chunks;
    const int p1 = FLEX ? min(p0 + poses_per_unit, n_poses) : n_poses;
    if (p0 >= p1) continue;  // CTA-uniform: a unit beyond the live rows (device-resident callers)
    if (g != cur_g) {  // CTA-uniform
      if (threadIdx.x == 0) {
        fence_proxy_async();
        uint32_t bytes = lig_loaded ? 0u : (uint32_t)rc.n_lig_pad * 16u;
        int n_rows = 0;
        for (int r = 0; r < rc.rows_max; ++r) n_rows += rc.group_types[g * RG_MAX_ROWS + r] >= 0;
        bytes += (uint32_t)n_rows * row_bytes;
        mbar_expect_tx(bar, bytes);
        if (!lig_loaded) bulk_g2s(l4, rc.lig4, (uint32_t)rc.n_lig_pad * 16u, bar);
        for (int r = 0; r < rc.rows_max; ++r) {
          const int ty = rc.group_types[g * RG_MAX_ROWS + r];
          if (ty >= 0)
            bulk_g2s(rows + (size_t)r * row_bytes, rows_src + (size_t)ty * row_bytes, row_bytes, bar);
        }
      }
      const int ia = g * 32 + lane;
      // byte address of (row, ligand type 0, index 4) minus what the magic-number index carries
      const int slot = rc.rec_slot[ia];  // -1 = pad lane
      rowoff = slot < 0 ? 0xffffffffu
                        : abs_base + (unsigned)(rows - smem_rigid) + (unsigned)slot * row_bytes -
                              ((RG_MAGIC_BITS + (unsigned)RG_SLOT0) << 3);
      mbar_wait(bar, phase);
      if (!FLEX) {
        rg_after_copy(rowoff);
        rg_after_copy(l4_addr);
      }
      phase ^= 1u;
      lig_loaded = true;
      cur_g = g;
    }
    const int pos_base = g * 32;

    for (;;) {
      int p = 0;
      if (lane == 0) p = atomicAdd(s_pose_next, 1);
      p = __shfl_sync(0xffffffffu, p, 0) * pstep + p0;
      if (p >= p1) break;
      const double *prep = prep_all + (size_t)p * RG_PREP;
      bool brute = false;
      // FP32 classification margins (see rigid_row): the handle's, valid for coordinates inside the cell grid
      float thr_out = rc.thr_out, hme = rc.half_minus_eps, delta102 = rc.delta, reach_abs = 3.0e38f;
      if (FLEX) {
        // this pose's ligand block (ligand frame, f32) into the warp's slice of shared memory
        // asynchronous copies (LDGSTS: no registers, all in flight at once), waited for right before the first row, so
        // they land while the warp moves its receptor atoms into the pose's frame and looks their cells up
        const float4 *src = lig4p + (size_t)p * rc.n_lig_pad;
        for (int i = lane; i < rc.n_lig_pad; i += 32) cp_async16(l4_off + (uint32_t)i * 16u, src + i);
        cp_async_commit();
        const float dmax = pose_flag[p];
        brute = dmax != 0.f;
        if (brute) {
          // A tile moved further than its slack: every ligand tile is a candidate, and an atom can be in range up to
          // dmax outside the grid, so the bound M on the coordinates (hence delta) grows by dmax; beyond the range
          // the bin-space test was proven for (delta < 0.01) nothing is decided in FP32 (hme < 0: all exact).
          reach_abs = rc.grid_maxabs + dmax + 0.1f;
          const float d = 2.0e-4f + 1.3e-5f * reach_abs;
          thr_out = 225.0f + d;
          delta102 = 1.02f * d;
          hme = d < 0.01f ? rc.half_minus_eps : -1.0f;
        }
      }
      acc_t acc0 = 0, acc1 = 0;
      unsigned ifr_mask = 0u;
      rg_score_group<DETAIL, FLEX>(rc, bb, l4_addr, lane_sw, rowoff, p, pos_base, prep, brute, thr_out, hme, delta102,
                                   reach_abs, acc0, acc1, ifr_mask, rc_dev);
      __syncwarp();  // FLEX: also "every lane is done with this pose's ligand block"
      const acc_t tsum = warp_sum(rg_join(acc0, acc1));
      const unsigned rbits = __reduce_or_sync(0xffffffffu, ifr_mask);
      if (lane == 0) {
        reinterpret_cast<acc_t *>(bb.partials)[(size_t)p * rc.n_groups + g] = tsum;  // both are 8 bytes
        bb.iface_rec[(size_t)p * rc.n_groups + g] = rbits;
      }
    }
  }
}

}  // namespace ldb200

// ld_kernels.cuh — hand-written sm_100a kernels for the LightDock scoring hot path.
//
// Three kernels per batch of poses (all on one stream):
//   1. transform_kernel   — Quaternion::rotate + translation + ANM displacement for every ligand
//                           atom (and ANM for the receptor), bit-exact operation order of
//                           src/qt.rs:57-61,174-185 and src/dfire.rs:282-320; writes one SoA block per
//                           pose plus conservative float4 bounding spheres per spatial tile.
//   2. dfire_pair_kernel / dna_pair_kernel
//                         — the cut-off pair loop (src/dfire.rs:325-345, src/dna.rs:471-512).  One CTA per
//                           (pose, receptor tile range).  The pose's ligand block is staged into shared
//                           memory with one cp.async.bulk (TMA) copy completing on an mbarrier; each warp
//                           owns receptor tiles (one atom per lane, coordinates in registers), culls
//                           ligand tiles 32-at-a-time with an FP32 sphere test + ballot, and evaluates the
//                           surviving pairs in exact, never-fused FP64.  Energies: per-lane FP64 accumulators
//                           -> fixed-order warp-shuffle tree -> per-tile slot in shared memory -> ordered CTA
//                           sum (bit-reproducible run to run).  Interface flags: ballot words / shared-memory
//                           bitmaps.
//   3. finalize_kernel    — restraint fractions, membrane fraction and the score algebra
//                           (src/dfire.rs:347-361, src/dna.rs:513-528, src/scoring.rs:21-47).
//
// Everything that feeds a discrete decision (cut-off tests, bin index, interface test) is computed
// with __dadd_rn/__dsub_rn/__dmul_rn/__dsqrt_rn so the compiler can never contract it into an FMA.
#pragma once
#include "ld_device.cuh"
#include "../../include/lightdock_b200.h"

namespace ldb200 {

// ---------------------------------------------------------------------------------------------
// small PTX wrappers (mbarrier + bulk async copy = TMA 1-D)
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LD_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LD_DONE_%=;\n"
      "bra LD_WAIT_%=;\n"
      "LD_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// Quaternion product, src/qt.rs:174-185 — four terms per component, summed left to right, no FMA.
struct Quat {
  double w, x, y, z;
};
__device__ __forceinline__ Quat qmul(const Quat &a, const Quat &b) {
  Quat r;
  r.w = __dsub_rn(__dsub_rn(__dsub_rn(__dmul_rn(a.w, b.w), __dmul_rn(a.x, b.x)), __dmul_rn(a.y, b.y)),
                  __dmul_rn(a.z, b.z));
  r.x = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.w, b.x), __dmul_rn(a.x, b.w)), __dmul_rn(a.y, b.z)),
                  __dmul_rn(a.z, b.y));
  r.y = __dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn(a.w, b.y), __dmul_rn(a.x, b.z)), __dmul_rn(a.y, b.w)),
                  __dmul_rn(a.z, b.x));
  r.z = __dadd_rn(__dsub_rn(__dadd_rn(__dmul_rn(a.w, b.z), __dmul_rn(a.x, b.y)), __dmul_rn(a.y, b.x)),
                  __dmul_rn(a.z, b.w));
  return r;
}

// Conservative bounding sphere of atoms [a, b) of an SoA block: centre = box centre rounded to f32,
// radius = max f64 distance to that f32 centre, inflated so the f32 value is never too small.
__host__ __device__ inline float4 tile_sphere(const double *x, const double *y, const double *z, int a, int b) {
  double lox = x[a], hix = x[a], loy = y[a], hiy = y[a], loz = z[a], hiz = z[a];
  for (int i = a + 1; i < b; ++i) {
    lox = x[i] < lox ? x[i] : lox; hix = x[i] > hix ? x[i] : hix;
    loy = y[i] < loy ? y[i] : loy; hiy = y[i] > hiy ? y[i] : hiy;
    loz = z[i] < loz ? z[i] : loz; hiz = z[i] > hiz ? z[i] : hiz;
  }
  float4 s;
  s.x = (float)(0.5 * (lox + hix));
  s.y = (float)(0.5 * (loy + hiy));
  s.z = (float)(0.5 * (loz + hiz));
  double r2 = 0.0;
  for (int i = a; i < b; ++i) {
    const double dx = x[i] - (double)s.x, dy = y[i] - (double)s.y, dz = z[i] - (double)s.z;
    const double d2 = dx * dx + dy * dy + dz * dz;
    r2 = d2 > r2 ? d2 : r2;
  }
  s.w = (float)(sqrt(r2) * 1.000001 + 1.0e-6);
  return s;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Kernel 1: pose transform.  grid = ceil(poses / TRANSFORM_PP), block = 256.
// A CTA transforms TRANSFORM_PP poses at once: thread = atom, and the atom's ANM mode vectors are read ONCE for all of
// them (one pose per CTA re-read every mode of both partners per pose from L2: 300 KB per pose for 1czy, which made
// this kernel L2-bandwidth bound and 30-45 % of the device time of the ANM configurations).  Per pose the operations
// and their order are the reference's: rotate, + translation, then mode k = 0, 1, ... each as multiply-then-add.
constexpr int TRANSFORM_PP = 4;
__global__ void __launch_bounds__(256) transform_kernel(const DeviceComplex cx, const BatchBuffers bb, int n_poses) {
  n_poses = live_poses(bb, n_poses);
  __shared__ float s_max[TRANSFORM_PP][2][8];
  __shared__ double s_pose[TRANSFORM_PP][11];  // tx ty tz | q | q^-1
  extern __shared__ double s_ext[];            // [TRANSFORM_PP][n_rec_modes + n_lig_modes]
  const int p0 = blockIdx.x * TRANSFORM_PP, np = min(TRANSFORM_PP, n_poses - p0);
  if (np <= 0) return;  // CTA-uniform (device-resident callers: rows beyond the live count)
  const int n_ext = cx.n_rec_modes + cx.n_lig_modes;
  if (threadIdx.x < np) {
    const double *pose = bb.poses + (size_t)(p0 + threadIdx.x) * cx.pose_len;
    const Quat q = {pose[3], pose[4], pose[5], pose[6]};
    // inverse(): conjugate / norm2, src/qt.rs:24-34,48-50,187-198
    const double n2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q.w, q.w), __dmul_rn(q.x, q.x)), __dmul_rn(q.y, q.y)),
                                __dmul_rn(q.z, q.z));
    double *o = s_pose[threadIdx.x];
    o[0] = pose[0]; o[1] = pose[1]; o[2] = pose[2];
    o[3] = q.w; o[4] = q.x; o[5] = q.y; o[6] = q.z;
    o[7] = __ddiv_rn(q.w, n2); o[8] = __ddiv_rn(-q.x, n2); o[9] = __ddiv_rn(-q.y, n2); o[10] = __ddiv_rn(-q.z, n2);
  }
  for (int i = threadIdx.x; i < np * n_ext; i += blockDim.x)
    s_ext[i] = bb.poses[(size_t)(p0 + i / n_ext) * cx.pose_len + 7 + i % n_ext];
  __syncthreads();

  const size_t lbs = lig_block_bytes(cx.n_lig_pad, cx.n_lig_tiles, cx.method);
  const size_t rbs = rec_block_bytes(cx.n_rec_pad, cx.n_rec_tiles);
  float lmax[TRANSFORM_PP], rmax[TRANSFORM_PP];
#pragma unroll
  for (int p = 0; p < TRANSFORM_PP; ++p) lmax[p] = rmax[p] = 0.f;
  for (int i = threadIdx.x; i < cx.n_lig_pad; i += blockDim.x) {
    double x[TRANSFORM_PP], y[TRANSFORM_PP], z[TRANSFORM_PP];
    float tw = 0.f;  // DFIRE: (float)(type * RG_SLOTS), see dfire_items()
    const bool real = i < cx.n_lig;
    if (real) {
      const Quat v = {0.0, cx.lig_x[i], cx.lig_y[i], cx.lig_z[i]};
#pragma unroll
      for (int p = 0; p < TRANSFORM_PP; ++p) {
        if (p >= np) break;
        const double *o = s_pose[p];
        const Quat q = {o[3], o[4], o[5], o[6]}, qi = {o[7], o[8], o[9], o[10]};
        // rotate(): self * (0, v) * self.inverse(), src/qt.rs:57-61
        const Quat r = qmul(qmul(q, v), qi);
        x[p] = __dadd_rn(r.x, o[0]);  // src/dfire.rs:286-288
        y[p] = __dadd_rn(r.y, o[1]);
        z[p] = __dadd_rn(r.z, o[2]);
      }
      for (int k = 0; k < cx.n_lig_modes; ++k) {  // src/dfire.rs:290-301
        const double *m = cx.lig_modes + (size_t)k * 3 * cx.n_lig_pad;
        const double mx = m[i], my = m[cx.n_lig_pad + i], mz = m[2 * cx.n_lig_pad + i];
#pragma unroll
        for (int p = 0; p < TRANSFORM_PP; ++p) {
          if (p >= np) break;
          const double e = s_ext[p * n_ext + cx.n_rec_modes + k];
          x[p] = __dadd_rn(x[p], __dmul_rn(mx, e));
          y[p] = __dadd_rn(y[p], __dmul_rn(my, e));
          z[p] = __dadd_rn(z[p], __dmul_rn(mz, e));
        }
      }
      if (cx.method == 0) tw = (float)((cx.lig_tb20[i] / 20) * RG_SLOTS);
    }
    const double lq = (cx.method != 0 && real) ? cx.lig_q[i] : 0.0;  // pads carry charge 0
#pragma unroll
    for (int p = 0; p < TRANSFORM_PP; ++p) {
      if (p >= np) break;
      if (!real) x[p] = y[p] = z[p] = LIG_PAD;
      else lmax[p] = fmaxf(lmax[p], fmaxf(fabsf((float)x[p]), fmaxf(fabsf((float)y[p]), fabsf((float)z[p]))));
      unsigned char *lb = bb.lig_blocks + (size_t)(p0 + p) * lbs;
      double *ox = reinterpret_cast<double *>(lb);
      ox[i] = x[p]; ox[cx.n_lig_pad + i] = y[p]; ox[2 * cx.n_lig_pad + i] = z[p];
      if (cx.method == 0) {
        reinterpret_cast<float4 *>(lb + lig_off_f4(cx.n_lig_pad))[i] = make_float4((float)x[p], (float)y[p], (float)z[p], tw);
      } else {
        double2 *od4 = reinterpret_cast<double2 *>(lb + lig_off_f4(cx.n_lig_pad));
        od4[2 * i] = make_double2(x[p], y[p]);
        od4[2 * i + 1] = make_double2(z[p], lq);
      }
    }
  }
  if (cx.n_rec_modes > 0) {  // src/dfire.rs:304-320
    for (int i = threadIdx.x; i < cx.n_rec_pad; i += blockDim.x) {
      double x[TRANSFORM_PP], y[TRANSFORM_PP], z[TRANSFORM_PP];
      const double x0 = cx.rec_x[i], y0 = cx.rec_y[i], z0 = cx.rec_z[i];
#pragma unroll
      for (int p = 0; p < TRANSFORM_PP; ++p) { x[p] = x0; y[p] = y0; z[p] = z0; }
      if (i < cx.n_rec) {
        for (int k = 0; k < cx.n_rec_modes; ++k) {
          const double *m = cx.rec_modes + (size_t)k * 3 * cx.n_rec_pad;
          const double mx = m[i], my = m[cx.n_rec_pad + i], mz = m[2 * cx.n_rec_pad + i];
#pragma unroll
          for (int p = 0; p < TRANSFORM_PP; ++p) {
            if (p >= np) break;
            const double e = s_ext[p * n_ext + k];
            x[p] = __dadd_rn(x[p], __dmul_rn(mx, e));
            y[p] = __dadd_rn(y[p], __dmul_rn(my, e));
            z[p] = __dadd_rn(z[p], __dmul_rn(mz, e));
          }
        }
      }
#pragma unroll
      for (int p = 0; p < TRANSFORM_PP; ++p) {
        if (p >= np) break;
        if (i < cx.n_rec)
          rmax[p] = fmaxf(rmax[p], fmaxf(fabsf((float)x[p]), fmaxf(fabsf((float)y[p]), fabsf((float)z[p]))));
        double *rx = reinterpret_cast<double *>(bb.rec_blocks + (size_t)(p0 + p) * rbs);
        rx[i] = x[p]; rx[cx.n_rec_pad + i] = y[p]; rx[2 * cx.n_rec_pad + i] = z[p];
      }
    }
  }
#pragma unroll
  for (int p = 0; p < TRANSFORM_PP; ++p) {
    const float a = warp_max_f(lmax[p]), b = warp_max_f(rmax[p]);
    if ((threadIdx.x & 31) == 0) { s_max[p][0][threadIdx.x >> 5] = a; s_max[p][1][threadIdx.x >> 5] = b; }
  }
  __syncthreads();  // block-scope visibility of the coordinates just written + the per-warp maxima
  if (threadIdx.x < np) {
    const int p = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a = fmaxf(a, s_max[p][0][w]); b = fmaxf(b, s_max[p][1][w]); }
    // inflate: the f32 conversions above round to nearest
    unsigned char *lb = bb.lig_blocks + (size_t)(p0 + p) * lbs;
    *reinterpret_cast<float4 *>(lb + lig_off_meta(cx.n_lig_pad, cx.n_lig_tiles, cx.method)) =
        make_float4(a * 1.0001f + 1.0f, 0.f, 0.f, 0.f);
    if (cx.n_rec_modes > 0) {
      unsigned char *rb = bb.rec_blocks + (size_t)(p0 + p) * rbs;
      reinterpret_cast<float4 *>(rb + (size_t)cx.n_rec_pad * 24)[cx.n_rec_tiles] = make_float4(b * 1.0001f + 1.0f, 0.f, 0.f, 0.f);
    }
  }
  for (int w = threadIdx.x; w < np * cx.n_lig_tiles; w += blockDim.x) {
    const int p = w / cx.n_lig_tiles, t = w % cx.n_lig_tiles;
    unsigned char *lb = bb.lig_blocks + (size_t)(p0 + p) * lbs;
    const double *ox = reinterpret_cast<const double *>(lb);
    const int a = t * LIG_TILE, b = min(a + LIG_TILE, cx.n_lig);
    reinterpret_cast<float4 *>(lb + lig_off_sph(cx.n_lig_pad, cx.method))[t] =
        tile_sphere(ox, ox + cx.n_lig_pad, ox + 2 * cx.n_lig_pad, a, b);
  }
  if (cx.n_rec_modes > 0)
    for (int w = threadIdx.x; w < np * cx.n_rec_tiles; w += blockDim.x) {
      const int p = w / cx.n_rec_tiles, t = w % cx.n_rec_tiles;
      unsigned char *rb = bb.rec_blocks + (size_t)(p0 + p) * rbs;
      const double *rx = reinterpret_cast<const double *>(rb);
      const int a = t * REC_TILE, b = min(a + REC_TILE, cx.n_rec);
      reinterpret_cast<float4 *>(rb + (size_t)cx.n_rec_pad * 24)[t] =
          tile_sphere(rx, rx + cx.n_rec_pad, rx + 2 * cx.n_rec_pad, a, b);
    }
}

// ---------------------------------------------------------------------------------------------
// shared-memory carve-up of the pair kernels
//   [0,8) mbarrier | [8,12) next-tile counter | [16,48) 4 u64 detail counters | [48,144) 24 u32 histogram
//   [144, ...)   TMA destination: DFIRE float4 xyzt[n_lig_pad] + spheres + meta ; DNA double4 xyzq[n_lig_pad] + same
//   DNA only:    ligand sqrt(eps) / radius (f64 each)
//   iface_lig bitmap | per-tile sums | DFIRE: per-warp work-item rings (64 u32 each)
//   DNA only:    van der Waals (A, B) table [lig types][rec types] (double2) + per ligand atom row offset (bytes)
struct PairSmem {
  uint64_t *bar;
  int *next_tile;
  unsigned long long *counters;  // detail: [0]=n_in_cutoff [1]=n_in_cutoff2 / ambiguous [2]=n_iface_pairs [3]=pairs tested
  unsigned *hist;                // detail: 21 bins (+pad)
  unsigned char *tma;            // start of the TMA-filled region
  const float4 *l4;              // DFIRE: f32 coordinates + type*20 bits
  const double *lxyzq;           // DNA: exact coordinates + charge, {x, y, z, q} per atom
  const float4 *lsph;
  const float4 *lmeta;
  double *lig_static;            // DNA: sqrt(eps) (or eps), rad
  unsigned *iface_lig;           // [lig_words]
  double *tile_sum;              // [tiles_per_split] (x2 for DNA)
  unsigned *rings;               // DFIRE: [warps][64]
  const unsigned char *vtab;     // DNA: (A, B) pairs, row = ligand type, column = receptor type
  const int *lig_vt;             // DNA: [n_lig_pad] byte offset of the atom's row in vtab
  double *pose;                  // DNA: tx ty tz | q | q^-1 | receptor extents | ligand extents of the CTA's pose
};
__host__ __device__ inline size_t dna_pose_doubles(int n_rec_modes, int n_lig_modes) { return 11 + n_rec_modes + n_lig_modes; }
__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
constexpr int RING = 64;
__host__ __device__ inline size_t pair_smem_bytes(int method, int n_lig_pad, int n_lig_tiles, int lig_words,
                                                  int tiles_per_split, int vdw_pairs) {
  size_t o = 144;
  if (method == 0) {
    o += (size_t)n_lig_pad * 16 + (size_t)n_lig_tiles * 16 + 16;
  } else {
    o += (size_t)n_lig_pad * 32 + (size_t)n_lig_tiles * 16 + 16;
    o += (size_t)n_lig_pad * 16;
  }
  o += align16((size_t)lig_words * 4);
  o += align16((size_t)tiles_per_split * 8 * (method == 0 ? 1 : 2));
  if (method == 0) o += (size_t)(PAIR_THREADS / 32) * RING * 4;  // work-item rings
  else o += (size_t)vdw_pairs * 16 + align16((size_t)n_lig_pad * 4) + 11 * 8 + 64 * 8;  // + pose scratch (<= 64 extents)
  return o;
}
__device__ __forceinline__ PairSmem carve(unsigned char *base, const DeviceComplex &cx, const BatchBuffers &bb) {
  PairSmem s;
  s.bar = reinterpret_cast<uint64_t *>(base);
  s.next_tile = reinterpret_cast<int *>(base + 8);
  s.counters = reinterpret_cast<unsigned long long *>(base + 16);
  s.hist = reinterpret_cast<unsigned *>(base + 48);
  unsigned char *o = base + 144;
  s.tma = o;
  s.l4 = nullptr; s.lxyzq = nullptr; s.lig_static = nullptr; s.rings = nullptr; s.vtab = nullptr; s.lig_vt = nullptr;
  s.pose = nullptr;
  if (cx.method == 0) {
    s.l4 = reinterpret_cast<const float4 *>(o);
    o += (size_t)cx.n_lig_pad * 16;
  } else {
    s.lxyzq = reinterpret_cast<const double *>(o);
    o += (size_t)cx.n_lig_pad * 32;
  }
  s.lsph = reinterpret_cast<const float4 *>(o);
  o += (size_t)cx.n_lig_tiles * 16;
  s.lmeta = reinterpret_cast<const float4 *>(o);
  o += 16;
  if (cx.method != 0) {
    s.lig_static = reinterpret_cast<double *>(o);
    o += (size_t)cx.n_lig_pad * 16;
  }
  s.iface_lig = reinterpret_cast<unsigned *>(o);
  o += align16((size_t)bb.lig_words * 4);
  s.tile_sum = reinterpret_cast<double *>(o);
  o += align16((size_t)bb.tiles_per_split * 8 * (cx.method == 0 ? 1 : 2));
  if (cx.method == 0) {
    s.rings = reinterpret_cast<unsigned *>(o);
  } else {
    s.vtab = o;
    o += (size_t)cx.vdw_nr * cx.vdw_nl * 16;
    s.lig_vt = reinterpret_cast<const int *>(o);
    o += align16((size_t)cx.n_lig_pad * 4);
    s.pose = reinterpret_cast<double *>(o);
  }
  return s;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// DIST_TO_BINS[idx] - 1 for idx 0..29 (src/dfire.rs:49-53,337).
// idx: 0 1 2 3 4 ... 15 | 16 17 18 19 20 21 22 23 24 25 26 27 28 29
// bin: 0 0 0 1 2 ... 13 | 13 14 14 15 15 16 16 17 17 18 18 19 19 20
__device__ __forceinline__ int dfire_bin_of(int idx) {
  return idx <= 2 ? 0 : (idx <= 15 ? idx - 2 : 13 + ((idx - 15) >> 1));
}

// Common prologue: stage the pose's ligand data (TMA) and the static ligand data, zero the bitmaps.
template <int METHOD>
__device__ __forceinline__ void pair_prologue(const DeviceComplex &cx, const BatchBuffers &bb, const PairSmem &s,
                                              int pose, int n_tiles_here) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(s.bar, 1);
    fence_mbar_init();
    *s.next_tile = 0;
  }
  __syncthreads();
  const bool fused = METHOD != 0 && bb.lig_blocks == nullptr;  // DNA/pyDock: the CTA transforms its pose's ligand itself
  if (fused) {
    // The pose transform of src/dna.rs:426-446 for this CTA's pose, straight into the staging area the pair loop reads
    // -- the same operations in the same order as transform_kernel (rotate = q (0,v) q^-1 with the reference's term order,
    // + translation, then mode k = 0, 1, ... each as multiply-then-add), so the coordinates are the same bits; what is saved
    // is a kernel, 56 bytes per atom and pose of HBM writes and the read back.
    const double *row = bb.poses + (size_t)pose * cx.pose_len;
    const int n_ext = cx.n_rec_modes + cx.n_lig_modes;
    if (tid == 0) {
      const Quat q = {row[3], row[4], row[5], row[6]};
      const double n2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q.w, q.w), __dmul_rn(q.x, q.x)), __dmul_rn(q.y, q.y)),
                                  __dmul_rn(q.z, q.z));
      double *o = s.pose;
      o[0] = row[0]; o[1] = row[1]; o[2] = row[2];
      o[3] = q.w; o[4] = q.x; o[5] = q.y; o[6] = q.z;
      o[7] = __ddiv_rn(q.w, n2); o[8] = __ddiv_rn(-q.x, n2); o[9] = __ddiv_rn(-q.y, n2); o[10] = __ddiv_rn(-q.z, n2);
    }
    for (int i = tid; i < n_ext; i += blockDim.x) s.pose[11 + i] = row[7 + i];
    __syncthreads();
    const double *o = s.pose;
    const Quat q = {o[3], o[4], o[5], o[6]}, qi = {o[7], o[8], o[9], o[10]};
    double2 *dst = reinterpret_cast<double2 *>(const_cast<double *>(s.lxyzq));
    for (int i = tid; i < cx.n_lig_pad; i += blockDim.x) {
      double x = LIG_PAD, y = LIG_PAD, z = LIG_PAD, lq = 0.0;  // pads: never within a cut-off, charge 0
      if (i < cx.n_lig) {
        const Quat v = {0.0, cx.lig_x[i], cx.lig_y[i], cx.lig_z[i]};
        const Quat r = qmul(qmul(q, v), qi);
        x = __dadd_rn(r.x, o[0]); y = __dadd_rn(r.y, o[1]); z = __dadd_rn(r.z, o[2]);
        for (int k = 0; k < cx.n_lig_modes; ++k) {
          const double *m = cx.lig_modes + (size_t)k * 3 * cx.n_lig_pad;
          const double e = o[11 + cx.n_rec_modes + k];
          x = __dadd_rn(x, __dmul_rn(m[i], e));
          y = __dadd_rn(y, __dmul_rn(m[cx.n_lig_pad + i], e));
          z = __dadd_rn(z, __dmul_rn(m[2 * cx.n_lig_pad + i], e));
        }
        lq = cx.lig_q[i];
      }
      dst[2 * i] = make_double2(x, y);
      dst[2 * i + 1] = make_double2(z, lq);
    }
    __syncthreads();
    float4 *sph = const_cast<float4 *>(s.lsph);
    for (int t = tid; t < cx.n_lig_tiles; t += blockDim.x) {  // tile_sphere() on the interleaved layout
      const int a = t * LIG_TILE, b = min(a + LIG_TILE, cx.n_lig);
      double lo[3] = {dst[2 * a].x, dst[2 * a].y, dst[2 * a + 1].x}, hi[3] = {lo[0], lo[1], lo[2]};
      for (int i = a + 1; i < b; ++i) {
        const double c[3] = {dst[2 * i].x, dst[2 * i].y, dst[2 * i + 1].x};
        for (int d = 0; d < 3; ++d) { lo[d] = c[d] < lo[d] ? c[d] : lo[d]; hi[d] = c[d] > hi[d] ? c[d] : hi[d]; }
      }
      float4 sp;
      sp.x = (float)(0.5 * (lo[0] + hi[0])); sp.y = (float)(0.5 * (lo[1] + hi[1])); sp.z = (float)(0.5 * (lo[2] + hi[2]));
      double r2 = 0.0;
      for (int i = a; i < b; ++i) {
        const double dx = dst[2 * i].x - (double)sp.x, dy = dst[2 * i].y - (double)sp.y, dz = dst[2 * i + 1].x - (double)sp.z;
        const double d2 = dx * dx + dy * dy + dz * dz;
        r2 = d2 > r2 ? d2 : r2;
      }
      sp.w = (float)(sqrt(r2) * 1.000001 + 1.0e-6);
      sph[t] = sp;
    }
  }
  if (!fused && tid == 0) {
    const unsigned char *lb = bb.lig_blocks + (size_t)pose * lig_block_bytes(cx.n_lig_pad, cx.n_lig_tiles, cx.method);
    // xyzt (DFIRE) or xyzq (DNA) | spheres | meta are contiguous: one bulk copy
    const uint32_t bytes = (uint32_t)cx.n_lig_pad * (uint32_t)lig_wide(METHOD) + (uint32_t)(cx.n_lig_tiles * 16 + 16);
    mbar_expect_tx(s.bar, bytes);
    bulk_g2s(s.tma, lb + lig_off_f4(cx.n_lig_pad), bytes, s.bar);
  }
  if (METHOD != 0) {
    double *le = s.lig_static, *lr = le + cx.n_lig_pad;
    for (int i = tid; i < cx.n_lig_pad; i += blockDim.x) {
      le[i] = cx.lig_seps[i]; lr[i] = cx.lig_rad[i];
    }
    if (cx.vdw_tab) {
      double2 *vt = reinterpret_cast<double2 *>(const_cast<unsigned char *>(s.vtab));
      for (int i = tid; i < cx.vdw_nr * cx.vdw_nl; i += blockDim.x) vt[i] = cx.vdw_tab[i];
      int *lv = const_cast<int *>(s.lig_vt);
      for (int i = tid; i < cx.n_lig_pad; i += blockDim.x) lv[i] = cx.lig_vt[i];
    }
  }
  for (int i = tid; i < bb.lig_words; i += blockDim.x) s.iface_lig[i] = 0u;
  for (int i = tid; i < n_tiles_here * (METHOD == 0 ? 1 : 2); i += blockDim.x) s.tile_sum[i] = 0.0;
  if (tid < 4) s.counters[tid] = 0ull;
  if (tid < 24) s.hist[tid] = 0u;
  __syncthreads();
  if (!fused) mbar_wait(s.bar, 0);
}

// ---------------------------------------------------------------------------------------------
// Kernel 2a: DFIRE pair loop, src/dfire.rs:325-345 — three culling levels, then an FP32 classification
// that is PROVABLY equal to the reference's FP64 decisions, with an exact FP64 fallback for the rest:
//   A. warp tile (32 receptor atoms) x 32 ligand tiles at a time: sphere-sphere test, one ligand tile per lane;
//   B. every lane tests ITS atom against each surviving ligand-tile sphere; (atom, tile) hits are compacted
//      (ballot + popc) into a warp-private ring of work items in shared memory;
//   C. whenever 32 items are queued each lane takes one: the atom's f32 coordinates come from the owning
//      lane by shuffle, the tile's 8 ligand atoms from shared memory (rotated start so lanes hit different
//      banks), distance^2 in FP32.  |d2f - dist_f64| <= delta (bound below), so if d2f is further than
//      delta from every decision threshold (the 29 bin edges ((k+1)/2)^2, the 225 cut-off, the 6.0025
//      interface edge) the bin / cut-off / interface decisions taken on d2f are exactly the reference's.
//      Otherwise (~0.1 % of in-range pairs) the pair is re-evaluated in exact never-fused FP64 from the
//      coordinates in global memory.  The table value is gathered as f64 and accumulated in f64.
// Error bound: coordinates are rounded to f32 (<= 2^-24*M each, M = max |coordinate|), the difference and
// the 3-term sum add <= 4 ulp_f32 of d2f:  |d2f - dist| <= 2*sqrt(3*240)*2^-23*M + 5*2^-24*240 + fp64 noise
// < 6.4e-6*M + 8e-5.  delta = 2e-5*M + 5e-4 leaves a 3x margin.
// Exact fallback for one pair: the reference's arithmetic, src/dfire.rs:331-342, from the f64 coordinates in
// global memory.  Returns -1 if the pair is outside the cut-off, else bin | (interface ? 32 : 0).
// Deliberately not inlined: it runs for ~0.1 % of the in-range pairs and must not bloat the hot loop.
__device__ __noinline__ int dfire_exact_pair(const double *gx, const double *gy, const double *gz, const double *glx,
                                             const double *gly, const double *glz, int ia, int j) {
  const double ex = __dsub_rn(gx[ia], glx[j]), ey = __dsub_rn(gy[ia], gly[j]), ez = __dsub_rn(gz[ia], glz[j]);
  const double dist = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
  if (!(dist <= 225.0)) return -1;
  const double d = __dsub_rn(__dmul_rn(__dsqrt_rn(dist), 2.0), 1.0);
  const int bin = dfire_bin_of((int)d);  // `d as usize`: truncation, d in (-1, 29]
  return bin | (d <= 3.9 ? 32 : 0);      // INTERFACE_CUTOFF on the bin-space value, src/dfire.rs:339
}

__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// DIST_TO_BINS[idx]-1 without branches: idx-2 below the knee at 15, (idx+11)/2 above it, clamped at 0.
__device__ __forceinline__ int dfire_bin_fast(int idx) { return min(max(idx - 2, 0), (idx + 11) >> 1); }

// One row of <= 32 work items (receptor atom = owner lane, ligand tile of 8 atoms), one item per lane.
// Classification in bin space, t = 2*sqrt(dist) - 1 (src/dfire.rs:336):
//   u = 2*d2f*rsqrt.approx(d2f) - 1.5 = t_f32 - 0.5;  m = u + 1.5*2^23 holds rint(u) in its low mantissa bits;
//   g = u - rint(u) = frac(t_f32) - 0.5 exactly.
//   |t_f32 - t_ref| <= 2|sqrt(d2f) - sqrt(dist_ref)| + 6.4e-6 <= 1.02*delta/sqrt(d2f) + 6.4e-6 for d2f >= 3.9 and
//   delta <= 0.3, so |g| + 1.02*delta*rsqrt(d2f) <= 0.5 - 2.5e-5 proves floor(t_ref) = rint(u) there; below d2f = 3.9
//   every candidate index (t < 3) is bin 0 anyway (DIST_TO_BINS, src/dfire.rs:49-53), and rint(u) = -1 for t < 0,
//   where the reference's `d as usize` saturates to index 0, is slot -1 = slot 0 of the re-indexed table potx.
// A pair with d2f <= 225 + delta that fails the test is re-evaluated exactly (dfire_exact_pair); the interface test
// (d <= 3.9 <=> dist <= 6.0025) is decided outside the hot loop: a per-item min(d2f) sends the rare items with a
// contact near or below 2.45 A to a second pass that decides it in FP32 outside 6.0025 +- delta, exactly inside.
// The table value is one 8-byte gather from potx: element bits(m + w) + rowx, w = (float)(type_lig * RG_SLOTS).
template <bool DETAIL>
__device__ __forceinline__ void dfire_items(const PairSmem &s, const unsigned *ring, int head, int n_active, float rxf,
                                            float ryf, float rzf, unsigned rowx, int toff, int tile_base,
                                            const double *gx, const double *gy, const double *gz, const double *glx,
                                            const double *gly, const double *glz, const double *__restrict__ pot,
                                            const double *__restrict__ potx, const unsigned short *lig_tb20,
                                            float delta, float hme, int n_lig, double &acc0, double &acc1,
                                            unsigned &ifr_mask, unsigned &n_in, unsigned &n_if, unsigned &n_tested,
                                            unsigned &n_amb) {
  const int lane = threadIdx.x & 31;
  const bool active = lane < n_active;
  unsigned item = active ? ring[(head + lane) & (RING - 1)] : (unsigned)(lane << 16);
  const int i = item >> 16, lt = item & 0xffffu;
  const float ax = __shfl_sync(0xffffffffu, rxf, i), ay = __shfl_sync(0xffffffffu, ryf, i),
              az = __shfl_sync(0xffffffffu, rzf, i);
  const unsigned rx = __shfl_sync(0xffffffffu, rowx, i);
  const int at = __shfl_sync(0xffffffffu, toff, i);
  if (!active) return;
  const float thr_out = 225.0f + delta, delta102 = 1.02f * delta;
  if (DETAIL) n_tested += min(LIG_TILE, n_lig - lt * LIG_TILE);
  const int jbase = lt * LIG_TILE;
  unsigned slow_bits = 0u;
  float mind2 = 3.0e38f;
#pragma unroll
  for (int k = 0; k < LIG_TILE; ++k) {
    const float4 a = s.l4[jbase + ((k + lane) & (LIG_TILE - 1))];  // rotated start: conflict-free LDS.128
    const float dx = ax - a.x, dy = ay - a.y, dz = az - a.z;
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    mind2 = fminf(mind2, d2);
    const float rs = rsqrt_approx(d2);
    const float u = fmaf(d2 * rs, 2.0f, -1.5f);
    const float m = __fadd_rn(u, RG_MAGIC);
    const float g = __fsub_rn(u, __fsub_rn(m, RG_MAGIC));
    const bool inr = d2 <= thr_out;
    const bool fast = inr & (fmaf(delta102, rs, fabsf(g)) <= hme);
    if (inr & !fast) slow_bits |= 1u << k;
    if (fast) {
      const double v = __ldg(potx + (unsigned)((unsigned)__float_as_int(__fadd_rn(m, a.w)) + rx));
      if (k & 1) acc1 = __dadd_rn(acc1, v);
      else acc0 = __dadd_rn(acc0, v);
      if (DETAIL) {
        ++n_in;
        atomicAdd(&s.hist[dfire_bin_fast(__float_as_int(m) - (int)RG_MAGIC_BITS)], 1u);
      }
    }
  }
  if (mind2 <= 6.0025f + delta) {  // rare: a contact near or below the 2.45 A interface edge (src/dfire.rs:339-342)
    for (int k = 0; k < LIG_TILE; ++k) {
      if ((slow_bits >> k) & 1u) continue;  // the exact path below owns this pair entirely
      const int j = jbase + ((k + lane) & (LIG_TILE - 1));
      const float4 a = s.l4[j];
      const float dx = ax - a.x, dy = ay - a.y, dz = az - a.z;
      const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));  // same operations as above: same bits
      if (d2 > 6.0025f + delta) continue;
      bool ifc = d2 < 6.0025f - delta;
      if (!ifc) ifc = (dfire_exact_pair(gx, gy, gz, glx, gly, glz, tile_base + i, j) & ~31) == 32;  // -1 -> false
      if (ifc) {
        ifr_mask |= 1u << i;
        atomicOr(&s.iface_lig[j >> 5], 1u << (j & 31));
        if (DETAIL) ++n_if;
      }
    }
  }
  if (slow_bits) {  // rare: too close to a decision threshold -> the reference's FP64 arithmetic for that pair
    double extra = 0.0;
    for (unsigned b = slow_bits; b; b &= b - 1) {
      const int j = jbase + ((__ffs(b) - 1 + lane) & (LIG_TILE - 1));
      if (DETAIL) ++n_amb;
      const int r = dfire_exact_pair(gx, gy, gz, glx, gly, glz, tile_base + i, j);
      if (r >= 0) {
        extra = __dadd_rn(extra, __ldg(pot + at + lig_tb20[j] + (r & 31)));
        if (r & 32) {
          ifr_mask |= 1u << i;
          atomicOr(&s.iface_lig[j >> 5], 1u << (j & 31));
          if (DETAIL) ++n_if;
        }
        if (DETAIL) {
          ++n_in;
          atomicAdd(&s.hist[r & 31], 1u);
        }
      }
    }
    acc0 = __dadd_rn(acc0, extra);
  }
}

template <bool DETAIL>
__global__ void __launch_bounds__(PAIR_THREADS, 2)
    dfire_pair_kernel(const DeviceComplex cx, const BatchBuffers bb, int n_poses) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pose = blockIdx.x / bb.rec_splits, split = blockIdx.x % bb.rec_splits;
  if (pose >= live_poses(bb, n_poses)) return;
  const int t0 = split * bb.tiles_per_split;
  const int t1 = min(t0 + bb.tiles_per_split, cx.n_rec_tiles);
  const PairSmem s = carve(smem_raw, cx, bb);
  pair_prologue<0>(cx, bb, s, pose, t1 - t0);

  const int lane = threadIdx.x & 31;
  unsigned *ring = s.rings + (threadIdx.x >> 5) * RING;
  const unsigned char *lb = bb.lig_blocks + (size_t)pose * lig_block_bytes(cx.n_lig_pad, cx.n_lig_tiles, cx.method);
  const double *glx = reinterpret_cast<const double *>(lb), *gly = glx + cx.n_lig_pad, *glz = gly + cx.n_lig_pad;
  const double *gx = cx.rec_x, *gy = cx.rec_y, *gz = cx.rec_z;
  const float4 *gsph = cx.rec_sphere;
  float maxabs = fmaxf(cx.rec_maxabs, s.lmeta->x);
  if (cx.n_rec_modes > 0) {
    const unsigned char *rb = bb.rec_blocks + (size_t)pose * rec_block_bytes(cx.n_rec_pad, cx.n_rec_tiles);
    gx = reinterpret_cast<const double *>(rb); gy = gx + cx.n_rec_pad; gz = gy + cx.n_rec_pad;
    gsph = reinterpret_cast<const float4 *>(gz + cx.n_rec_pad);
    maxabs = fmaxf(gsph[cx.n_rec_tiles].x, s.lmeta->x);
  }
  const float delta = 5.0e-4f + 2.0e-5f * maxabs;  // |d2f - dist| bound (see above)
  const float lin = 1.0e-4f + 2.4e-7f * maxabs;    // error of an f32 atom-to-sphere-centre distance
  // the bin-space test of dfire_items needs delta <= 0.3 (coordinates below ~15,000 A); beyond that nothing is
  // decided in FP32 (hme < 0 sends every in-range pair to the exact path)
  const float hme = delta <= 0.3f ? 0.5f - 2.5e-5f : -1.0f;
  const double *__restrict__ pot = cx.pot;
  const double *__restrict__ potx = cx.potx;
  unsigned *iface_rec_out = bb.iface_rec + (size_t)pose * cx.n_rec_tiles;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(s.next_tile, 1);
    t = __shfl_sync(0xffffffffu, t, 0) + t0;
    if (t >= t1) break;
    const int ia = t * REC_TILE + lane;
    const float rxf = (float)gx[ia], ryf = (float)gy[ia], rzf = (float)gz[ia];
    const int toff = cx.rec_toff[ia];
    const unsigned rowx = cx.rec_rowx[ia];
    const float4 rs = gsph[t];
    double acc0 = 0.0, acc1 = 0.0;
    unsigned ifr_mask = 0u, n_in = 0, n_if = 0, n_tested = 0, n_amb = 0;
    int q_head = 0, q_count = 0;  // warp-uniform

    for (int lt0 = 0; lt0 < cx.n_lig_tiles; lt0 += 32) {
      // A: this warp's tile sphere against 32 ligand-tile spheres, one per lane
      const int ltl = lt0 + lane;
      bool hit = false;
      if (ltl < cx.n_lig_tiles) {
        const float4 ls = s.lsph[ltl];
        const float dx = rs.x - ls.x, dy = rs.y - ls.y, dz = rs.z - ls.z;
        const float d2 = dx * dx + dy * dy + dz * dz;
        const float reach = 15.0f + rs.w + ls.w;  // sqrt(225): src/dfire.rs:334
        hit = d2 <= reach * reach * 1.00001f;
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      while (m) {
        const int lt = lt0 + __ffs(m) - 1;
        m &= m - 1;
        // B: my atom against this ligand tile's sphere
        const float4 ls = s.lsph[lt];
        const float dx = rxf - ls.x, dy = ryf - ls.y, dz = rzf - ls.z;
        const float d2 = dx * dx + dy * dy + dz * dz;
        const float reach = 15.0f + ls.w + lin;
        const bool mine = d2 <= reach * reach * 1.00001f;
        const unsigned pm = __ballot_sync(0xffffffffu, mine);
        if (pm) {
          if (mine) ring[(q_head + q_count + __popc(pm & lt_mask)) & (RING - 1)] = ((unsigned)lane << 16) | (unsigned)lt;
          q_count += __popc(pm);
          __syncwarp();
          if (q_count >= 32) {  // C: a full row of work items
            dfire_items<DETAIL>(s, ring, q_head, 32, rxf, ryf, rzf, rowx, toff, t * REC_TILE, gx, gy, gz, glx, gly, glz,
                                pot, potx, cx.lig_tb20, delta, hme, cx.n_lig, acc0, acc1, ifr_mask, n_in, n_if, n_tested,
                                n_amb);
            q_head = (q_head + 32) & (RING - 1);
            q_count -= 32;
            __syncwarp();
          }
        }
      }
    }
    if (q_count > 0)
      dfire_items<DETAIL>(s, ring, q_head, q_count, rxf, ryf, rzf, rowx, toff, t * REC_TILE, gx, gy, gz, glx, gly, glz,
                          pot, potx, cx.lig_tb20, delta, hme, cx.n_lig, acc0, acc1, ifr_mask, n_in, n_if, n_tested, n_amb);
    __syncwarp();
    const double tsum = warp_sum(__dadd_rn(acc0, acc1));
    const unsigned rbits = __reduce_or_sync(0xffffffffu, ifr_mask);
    if (lane == 0) {
      s.tile_sum[t - t0] = tsum;
      iface_rec_out[t] = rbits;
    }
    if (DETAIL) {
      n_in = warp_sum_u32(n_in); n_if = warp_sum_u32(n_if);
      n_tested = warp_sum_u32(n_tested); n_amb = warp_sum_u32(n_amb);
      if (lane == 0) {
        atomicAdd(&s.counters[0], (unsigned long long)n_in);
        atomicAdd(&s.counters[1], (unsigned long long)n_amb);
        atomicAdd(&s.counters[2], (unsigned long long)n_if);
        atomicAdd(&s.counters[3], (unsigned long long)n_tested);
      }
    }
  }
  __syncthreads();
  // per-tile sums go out as they are: the finalize kernel adds them in tile order, so a pose's energy
  // does not depend on how many CTAs shared its receptor nor on the batch it was scored in
  for (int i = threadIdx.x; i < t1 - t0; i += blockDim.x)
    bb.partials[(size_t)pose * cx.n_rec_tiles + t0 + i] = s.tile_sum[i];
  unsigned *ifl = bb.iface_lig + ((size_t)pose * bb.rec_splits + split) * bb.lig_words;
  for (int i = threadIdx.x; i < bb.lig_words; i += blockDim.x) ifl[i] = s.iface_lig[i];
  if (DETAIL) {
    ld_pose_detail *dt = reinterpret_cast<ld_pose_detail *>(bb.detail) + pose;
    if (threadIdx.x == 0) {
      atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_in_cutoff), s.counters[0]);
      atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_exact_fallback), s.counters[1]);
      atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_interface_pairs), s.counters[2]);
      atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_pairs_tested), s.counters[3]);
    }
    if (threadIdx.x < 21)
      atomicAdd(reinterpret_cast<unsigned long long *>(&dt->bin_hist[threadIdx.x]),
                (unsigned long long)s.hist[threadIdx.x]);
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2b: DNA / pyDock pair loop, src/dna.rs:471-512 (= src/pydock.rs:486-527).
//
// The kernel is FP64-pipe bound (one warp instruction per 2 cycles per scheduler), so everything that does not have
// to be an FP64 instruction is not one, and the eight pairs of a ligand tile are evaluated branch-free so their
// dependency chains interleave:
//   * d2 is evaluated with two FMAs (6 FP64 instructions instead of 8).  It differs from the reference's never-fused
//     d2 by < 4 ulp, so a cut-off decision taken on it is the reference's unless d2 sits within a few ulp of the
//     threshold.  The decisions are taken on the HIGH WORD of d2 with integer compares (ALU pipe): 900.0 and 100.0 are
//     exactly representable in the high word (0x408C2000'00000000, 0x40590000'00000000), so hi < H-1 is surely
//     inside and hi > H surely outside (4 ulp is 2^-30 of a high-word step); if any pair of the tile has one of the
//     two high words around a threshold (|d2 - T| < 1e-3, about one pair in 10^6) the whole tile is redone by
//     dna_tile_exact with the reference's own never-fused arithmetic and IEEE divides.  Same for 3.9*3.9 =
//     0x402E6B85'1EB851EB with the single ambiguous high word 0x402E6B85.
//     tests/test_gpu_parity.py::test_dna_decision_thresholds_exact drives pairs whose fused and unfused d2 fall on
//     different sides of each threshold through this path.
//   * The two quotients (Coulomb q1*q2/d2, van der Waals (r/d)^6) use the MUFU.RCP64H seed (rel. error < 2^-20) and
//     ONE Newton step, x1 = x + x*(1 - d2*x): relative error < 1e-12.  They feed only continuous quantities (a sum,
//     a clamp and a min whose two branches agree at the switch point), never a cut-off decision, so this stays 10^6
//     below the 1e-6 tolerance.  The receptor charge is factored out of the tile loop: sum_j clamp(q_l/d2, +-C/|q_r|)
//     is multiplied by q_r once per receptor tile (the clamp of q_r*q_l/d2 to +-C is the same set of pairs).
//   * sqrt(eps_r*eps_l) (src/dna.rs:496) is sqrt(eps_r)*sqrt(eps_l) with the per-atom roots taken once by ld_create
//     (continuous; 2 ulp); a complex with a negative eps keeps the reference's form (cx.vdw_sqrt_hoisted == 0).
//   * the clamp is an integer compare on the bit patterns (order-preserving for same-sign doubles).
//   * ligand tiles whose sphere is further than 10 A + radii from the receptor tile's cannot hold a van der Waals or
//     interface pair: they run a loop instance without those tests (1azp: 9 of 10 tile pairs), so the divergent
//     vdW branch is confined to the tile pairs that can take it.
// powi(6)/powi(3) follow LLVM's repeated-squaring expansion (x^2, x^4, x^2*x^4 ; x*x^2).
__device__ __forceinline__ double rcp_seed(double d) {  // MUFU.RCP64H: relative error < 2^-20
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  return x;
}
#ifndef LDB200_DNA_UNROLL
#define LDB200_DNA_UNROLL 8
#endif
constexpr int DNA_UNROLL = LDB200_DNA_UNROLL;  // pairs of a ligand tile evaluated interleaved (8 or 4)
constexpr unsigned DNA_HI_ELEC = 0x408C2000u;   // high word of 900.0 (low word 0)
constexpr unsigned DNA_HI_VDW = 0x40590000u;    // high word of 100.0 (low word 0)
constexpr unsigned DNA_HI_IFACE = 0x402E6B85u;  // high word of 3.9*3.9 = 15.209999999999999 (src/constants.rs:15)
static_assert(3.9 * 3.9 == 15.209999999999999, "INTERFACE_CUTOFF2");

struct DnaLane {  // one receptor atom (one lane)
  double x, y, z, q, se, rad;  // se = sqrt(eps) when hoisted, else eps
  long long clamp_bits;        // bit pattern of ELEC_MAX_CUTOFF / |q|: the clamp in units of q_l / d2
  int vt;                      // byte offset of the atom's van der Waals type inside a row of the (A, B) table
};
struct DnaAcc {
  double e_l0, e_l1;  // sum of clamped q_l/d2 (to be multiplied by the lane's q_r)
  double e_x;         // Coulomb terms of tiles redone exactly (already multiplied by q_r)
  double v;           // van der Waals
  unsigned n_e, n_v, n_if;
  bool iface_r;
};

// The reference's arithmetic for the eight pairs of one tile: never fused, IEEE divide and square root
// (src/dna.rs:471-512).  Rare (a pair within 1e-3 A^2 of a cut-off), so deliberately out of line.
__device__ __noinline__ void dna_tile_exact(const double2 *__restrict__ a2, const double *__restrict__ s_se,
                                            const double *__restrict__ s_rad, unsigned *iface_lig, bool hoisted,
                                            const DnaLane *rp, int j0, DnaAcc *accp) {
  const DnaLane &r = *rp;
  DnaAcc &acc = *accp;
  const double ELEC_MAX_CUTOFF = 1.0 * 4.0 / 332.0, ELEC_MIN_CUTOFF = -1.0 * 4.0 / 332.0;  // src/dna.rs:21-22
  for (int jj = 0; jj < LIG_TILE; ++jj) {
    const int j = j0 + jj;
    const double2 xy = a2[2 * jj], zq = a2[2 * jj + 1];
    const double dx = __dsub_rn(r.x, xy.x), dy = __dsub_rn(r.y, xy.y), dz = __dsub_rn(r.z, zq.x);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (d2 <= 900.0) {  // ELEC_DIST_CUTOFF2
      double e = __ddiv_rn(__dmul_rn(r.q, zq.y), d2);
      e = e > ELEC_MAX_CUTOFF ? ELEC_MAX_CUTOFF : e;
      e = e < ELEC_MIN_CUTOFF ? ELEC_MIN_CUTOFF : e;
      acc.e_x = __dadd_rn(acc.e_x, e);
      ++acc.n_e;
      if (d2 <= 100.0) {  // VDW_DIST_CUTOFF2
        const double pe = __dmul_rn(r.se, s_se[j]);
        const double ve = hoisted ? pe : __dsqrt_rn(pe);
        const double vr = __dadd_rn(r.rad, s_rad[j]);
        const double vr2 = __dmul_rn(vr, vr), vr4 = __dmul_rn(vr2, vr2), vr6 = __dmul_rn(vr2, vr4);
        const double p6 = __ddiv_rn(vr6, __dmul_rn(d2, __dmul_rn(d2, d2)));
        double k = __dmul_rn(ve, __dsub_rn(__dmul_rn(p6, p6), __dmul_rn(2.0, p6)));
        k = k > 1.0 ? 1.0 : k;
        acc.v = __dadd_rn(acc.v, k);
        ++acc.n_v;
        if (d2 <= 3.9 * 3.9) {  // INTERFACE_CUTOFF2
          acc.iface_r = true;
          atomicOr(&iface_lig[j >> 5], 1u << (j & 31));
          ++acc.n_if;
        }
      }
    }
  }
}

// acc += t if hi < bound, as ONE predicated DADD (the compiler's own choice is an unconditional DADD + two selects)
__device__ __forceinline__ void dadd_if_below(double &acc, double t, unsigned hi, unsigned bound) {
  asm("{\n.reg .pred p;\nsetp.lt.u32 p, %2, %3;\n@p add.rn.f64 %0, %0, %1;\n}" : "+d"(acc) : "d"(t), "r"(hi), "r"(bound));
}
__device__ __forceinline__ void dfma_if_below(double &acc, double a, double b, unsigned hi, unsigned bound) {
  asm("{\n.reg .pred p;\nsetp.lt.u32 p, %3, %4;\n@p fma.rn.f64 %0, %1, %2, %0;\n}" : "+d"(acc) : "d"(a), "d"(b), "r"(hi), "r"(bound));
}

// One ligand tile (8 atoms) against the lane's receptor atom.
//   CLOSE == false: the tile pair is further apart than cx.dna_close_reach (>= 10 A) + radii: no van der Waals pair,
//                   no interface pair, and no Coulomb term can reach the clamp (|q_r q_l| / d2 <= C there): the loop
//                   is 9 FP64 instructions per pair and has no branch.
//   CLOSE == true : clamp applied per term, and the van der Waals term
//        TAB == true : evaluated branch-free for all eight pairs from a shared-memory table indexed by the pair's
//                      (receptor, ligand) van der Waals types: with A = sqrt(e_r e_l) (r_r + r_l)^12 and
//                      B = 2 sqrt(e_r e_l) (r_r + r_l)^6 (f64, ld_create), k = ve (p6^2 - 2 p6) = x6 (A x6 - B),
//                      x6 = (1/d2)^3 from the reciprocal the Coulomb term already has: 5 FP64 instructions, no second
//                      reciprocal, no divergence (the per-lane candidate loop ran max-over-lanes trips with a third of
//                      the lanes active and cost more than the Coulomb part);
//        TAB == false: (more than 1024 type pairs) candidates collected in a mask and evaluated by their own lane.
template <bool CLOSE, bool DETAIL, bool TAB>
__device__ __forceinline__ void dna_tile(const PairSmem &s, const double *__restrict__ s_se,
                                         const double *__restrict__ s_rad, bool hoisted, const DnaLane &r, int j0,
                                         DnaAcc &acc) {
  const double2 *a2 = reinterpret_cast<const double2 *>(s.lxyzq) + 2 * j0;  // {x, y} {z, q} per ligand atom
  double t0 = 0.0, t1 = 0.0, v0 = 0.0, v1 = 0.0;
  unsigned kmin = 0xffffffffu, cnt = 0, cnt_v = 0;
  unsigned vmask = 0u, imask = 0u, near_v = 0u;  // CLOSE: pairs inside 100.0 / inside 3.9^2 by their high word; any
                                                 // pair with an ambiguous high word
  const unsigned c_hi = (unsigned)(r.clamp_bits >> 32), c_lo = (unsigned)r.clamp_bits;
#pragma unroll 1
  for (int j4 = 0; j4 < LIG_TILE; j4 += DNA_UNROLL) {
#pragma unroll
  for (int ju = 0; ju < DNA_UNROLL; ++ju) {
    const int jj = j4 + ju;
    const double2 xy = a2[2 * jj], zq = a2[2 * jj + 1];
    const double dx = __dsub_rn(r.x, xy.x), dy = __dsub_rn(r.y, xy.y), dz = __dsub_rn(r.z, zq.x);
    const double d2 = fma(dz, dz, fma(dy, dy, __dmul_rn(dx, dx)));
    const unsigned hi = (unsigned)__double2hiint(d2);
    kmin = min(kmin, hi - (DNA_HI_ELEC - 1u));  // <= 1 iff hi is one of the two words around 900.0
    const double x = rcp_seed(d2);
    const double e1 = fma(-d2, x, 1.0);
    const double x1 = fma(x, e1, x);
    if (!CLOSE) {
      dfma_if_below((ju & 1) ? t1 : t0, zq.y, x1, hi, DNA_HI_ELEC - 1u);
    } else {
      double t = __dmul_rn(zq.y, x1);
      {  // clamp q_l/d2 to +-C/|q_r| (src/dna.rs:484-489 divided by |q_r|) on the bit patterns
        const unsigned th = (unsigned)__double2hiint(t), ta = th & 0x7fffffffu, tl = (unsigned)__double2loint(t);
        if (ta > c_hi || (ta == c_hi && tl > c_lo)) t = __hiloint2double((int)(c_hi | (th & 0x80000000u)), (int)c_lo);
      }
      dadd_if_below((ju & 1) ? t1 : t0, t, hi, DNA_HI_ELEC - 1u);
      // the vdW / interface decisions are ambiguous on the high words around 100.0 and on the one holding 3.9*3.9
      near_v |= (hi - (DNA_HI_VDW - 1u) <= 1u) | (hi == DNA_HI_IFACE);
      if (TAB) {
        const double2 ab = *reinterpret_cast<const double2 *>(s.vtab + r.vt + s.lig_vt[j0 + jj]);
        const double x6 = __dmul_rn(__dmul_rn(x1, x1), x1);
        double k = __dmul_rn(fma(ab.x, x6, -ab.y), x6);
        if (__double2hiint(k) >= 0x3FF00000) k = 1.0;  // min(k, 1.0) (src/dna.rs:500-502) on the high word: k >= 1
        dadd_if_below((ju & 1) ? v1 : v0, k, hi, DNA_HI_VDW - 1u);
        imask |= (hi < DNA_HI_IFACE ? 1u : 0u) << jj;
        if (DETAIL) cnt_v += hi < DNA_HI_VDW - 1u;
      } else {
        vmask |= (hi <= DNA_HI_VDW ? 1u : 0u) << jj;
      }
    }
    if (DETAIL) cnt += hi < DNA_HI_ELEC - 1u;
  }
  }
  if (kmin <= 1u || (CLOSE && near_v)) {  // ~1e-6 of the tiles: the reference's own arithmetic decides
    DnaLane rc = r;  // copies: only this rare branch takes addresses, so r and acc stay in registers
    DnaAcc ac = acc;
    dna_tile_exact(a2, s_se, s_rad, s.iface_lig, hoisted, &rc, j0, &ac);
    acc = ac;
    vmask = 0u;
    imask = 0u;
  } else {
    acc.e_l0 = __dadd_rn(acc.e_l0, t0);
    acc.e_l1 = __dadd_rn(acc.e_l1, t1);
    if (DETAIL) acc.n_e += cnt;
    if (CLOSE && TAB) {
      acc.v = __dadd_rn(acc.v, __dadd_rn(v0, v1));
      if (DETAIL) acc.n_v += cnt_v;
    }
  }
  if (CLOSE && TAB && imask) {  // rare: a contact inside 3.9 A (src/dna.rs:507-510)
    acc.iface_r = true;
    for (unsigned m = imask; m; m &= m - 1u) {
      const int j = j0 + __ffs(m) - 1;
      atomicOr(&s.iface_lig[j >> 5], 1u << (j & 31));
      if (DETAIL) ++acc.n_if;
    }
  }
  if (CLOSE && !TAB) {
    // the few pairs inside 10 A (no ambiguous high word among them: near_v == 0): src/dna.rs:494-510
    for (unsigned m = vmask; m; m &= m - 1u) {
      const int jj = __ffs(m) - 1, j = j0 + jj;
      const double2 xy = a2[2 * jj], zq = a2[2 * jj + 1];
      const double dx = __dsub_rn(r.x, xy.x), dy = __dsub_rn(r.y, xy.y), dz = __dsub_rn(r.z, zq.x);
      const double d2 = fma(dz, dz, fma(dy, dy, __dmul_rn(dx, dx)));  // the same operations as above: the same bits
      const double pe = __dmul_rn(r.se, s_se[j]);
      const double ve = hoisted ? pe : __dsqrt_rn(pe);
      const double vr = __dadd_rn(r.rad, s_rad[j]);
      const double vr2 = __dmul_rn(vr, vr), vr4 = __dmul_rn(vr2, vr2);
      const double vr6 = __dmul_rn(vr2, vr4);
      const double d6 = __dmul_rn(d2, __dmul_rn(d2, d2));
      const double x6 = rcp_seed(d6);
      const double e6 = fma(-d6, x6, 1.0);
      const double t6 = __dmul_rn(vr6, x6);
      const double p6 = fma(t6, e6, t6);
      double k = __dmul_rn(ve, __dsub_rn(__dmul_rn(p6, p6), __dmul_rn(2.0, p6)));
      k = k > 1.0 ? 1.0 : k;
      acc.v = __dadd_rn(acc.v, k);
      if (DETAIL) ++acc.n_v;
      if ((unsigned)__double2hiint(d2) < DNA_HI_IFACE) {
        acc.iface_r = true;
        atomicOr(&s.iface_lig[j >> 5], 1u << (j & 31));
        if (DETAIL) ++acc.n_if;
      }
    }
  }
  // Lanes leave the branches above after different trip counts; without an explicit reconvergence point the warp
  // stays split over the following tiles (measured: 12 active threads per instruction on average).
  __syncwarp();
}

template <bool DETAIL, bool TAB>
__global__ void __launch_bounds__(DNA_THREADS, DNA_CTAS_PER_SM)
    dna_pair_kernel(const DeviceComplex cx, const BatchBuffers bb, int n_poses) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pose = blockIdx.x / bb.rec_splits, split = blockIdx.x % bb.rec_splits;
  if (pose >= live_poses(bb, n_poses)) return;
  const int t0 = split * bb.tiles_per_split;
  const int t1 = min(t0 + bb.tiles_per_split, cx.n_rec_tiles);
  const PairSmem s = carve(smem_raw, cx, bb);
  pair_prologue<1>(cx, bb, s, pose, t1 - t0);
  const double *s_se = s.lig_static, *s_rad = s_se + cx.n_lig_pad;
  const bool hoisted = cx.vdw_sqrt_hoisted != 0;

  const int lane = threadIdx.x & 31;
  const double *gx = cx.rec_x, *gy = cx.rec_y, *gz = cx.rec_z;
  const float4 *gsph = cx.rec_sphere;
  const bool rec_anm_here = cx.n_rec_modes > 0 && bb.rec_blocks == nullptr;  // fused: the warp deforms its own tile
  if (cx.n_rec_modes > 0 && !rec_anm_here) {
    const unsigned char *rb = bb.rec_blocks + (size_t)pose * rec_block_bytes(cx.n_rec_pad, cx.n_rec_tiles);
    gx = reinterpret_cast<const double *>(rb); gy = gx + cx.n_rec_pad; gz = gy + cx.n_rec_pad;
    gsph = reinterpret_cast<const float4 *>(gz + cx.n_rec_pad);
  }
  unsigned *iface_rec_out = bb.iface_rec + (size_t)pose * cx.n_rec_tiles;

  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(s.next_tile, 1);
    t = __shfl_sync(0xffffffffu, t, 0) + t0;
    if (t >= t1) break;
    const int ia = t * REC_TILE + lane;
    DnaLane r = {gx[ia], gy[ia], gz[ia], cx.rec_q[ia], cx.rec_seps[ia], cx.rec_rad[ia], 0, TAB ? cx.rec_vt[ia] : 0};
    float4 rs_here = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rec_anm_here) {
      // receptor ANM (src/dna.rs:448-464) for the 32 atoms of this tile, k ascending, multiply-then-add: the same bits
      // as transform_kernel writes; then tile_sphere() of the deformed tile by warp reductions (pads excluded)
      const bool real = ia < cx.n_rec;
      if (real)
        for (int k = 0; k < cx.n_rec_modes; ++k) {
          const double *m = cx.rec_modes + (size_t)k * 3 * cx.n_rec_pad;
          const double e = s.pose[11 + k];
          r.x = __dadd_rn(r.x, __dmul_rn(m[ia], e));
          r.y = __dadd_rn(r.y, __dmul_rn(m[cx.n_rec_pad + ia], e));
          r.z = __dadd_rn(r.z, __dmul_rn(m[2 * cx.n_rec_pad + ia], e));
        }
      double lo[3] = {real ? r.x : 1.0e300, real ? r.y : 1.0e300, real ? r.z : 1.0e300};
      double hi[3] = {real ? r.x : -1.0e300, real ? r.y : -1.0e300, real ? r.z : -1.0e300};
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
          hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
      rs_here.x = (float)(0.5 * (lo[0] + hi[0])); rs_here.y = (float)(0.5 * (lo[1] + hi[1])); rs_here.z = (float)(0.5 * (lo[2] + hi[2]));
      double d2 = 0.0;
      if (real) {
        const double dx = r.x - (double)rs_here.x, dy = r.y - (double)rs_here.y, dz = r.z - (double)rs_here.z;
        d2 = dx * dx + dy * dy + dz * dz;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
      rs_here.w = (float)(sqrt(d2) * 1.000001 + 1.0e-6);
    }
    // ELEC_MAX_CUTOFF / |q_r| (src/dna.rs:21); q_r == 0 gives +inf: no clamp, and every term is 0 * q_l/d2 = 0
    r.clamp_bits = __double_as_longlong(__ddiv_rn(1.0 * 4.0 / 332.0, fabs(r.q)));
    const float4 rs = rec_anm_here ? rs_here : gsph[t];
    DnaAcc acc = {0.0, 0.0, 0.0, 0.0, 0u, 0u, 0u, false};
    unsigned n_tested = 0;

    for (int lt0 = 0; lt0 < cx.n_lig_tiles; lt0 += 32) {
      const int lt = lt0 + lane;
      bool pass = false, close = false;
      if (lt < cx.n_lig_tiles) {
        const float4 ls = s.lsph[lt];
        const float dx = rs.x - ls.x, dy = rs.y - ls.y, dz = rs.z - ls.z;
        const float d2 = dx * dx + dy * dy + dz * dz;
        const float reach = 30.0f + rs.w + ls.w;  // ELEC_DIST_CUTOFF, the widest of the three
        pass = d2 <= reach * reach * 1.00001f;
        const float reach_v = cx.dna_close_reach + rs.w + ls.w;  // >= VDW_DIST_CUTOFF and the widest clamp distance
        close = d2 <= reach_v * reach_v * 1.00001f;
      }
      unsigned m = __ballot_sync(0xffffffffu, pass);
      const unsigned mv = __ballot_sync(0xffffffffu, close);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int j0 = (lt0 + b) * LIG_TILE;
        if (DETAIL && ia < cx.n_rec) n_tested += min(LIG_TILE, cx.n_lig - j0);
        if ((mv >> b) & 1u) dna_tile<true, DETAIL, TAB>(s, s_se, s_rad, hoisted, r, j0, acc);
        else dna_tile<false, DETAIL, TAB>(s, s_se, s_rad, hoisted, r, j0, acc);
      }
    }
    const double e_lane = __dadd_rn(__dmul_rn(r.q, __dadd_rn(acc.e_l0, acc.e_l1)), acc.e_x);
    const double esum = warp_sum(e_lane), vsum = warp_sum(acc.v);
    const unsigned rbits = __ballot_sync(0xffffffffu, acc.iface_r);
    if (lane == 0) {
      s.tile_sum[2 * (t - t0)] = esum;
      s.tile_sum[2 * (t - t0) + 1] = vsum;
      iface_rec_out[t] = rbits;
    }
    if (DETAIL) {
      const unsigned n_e = warp_sum_u32(acc.n_e), n_v = warp_sum_u32(acc.n_v), n_if = warp_sum_u32(acc.n_if);
      n_tested = warp_sum_u32(n_tested);
      if (lane == 0) {
        atomicAdd(&s.counters[0], (unsigned long long)n_e);
        atomicAdd(&s.counters[1], (unsigned long long)n_v);
        atomicAdd(&s.counters[2], (unsigned long long)n_if);
        atomicAdd(&s.counters[3], (unsigned long long)n_tested);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * (t1 - t0); i += blockDim.x)
    bb.partials[((size_t)pose * cx.n_rec_tiles + t0) * 2 + i] = s.tile_sum[i];
  unsigned *ifl = bb.iface_lig + ((size_t)pose * bb.rec_splits + split) * bb.lig_words;
  for (int i = threadIdx.x; i < bb.lig_words; i += blockDim.x) ifl[i] = s.iface_lig[i];
  if (DETAIL && threadIdx.x == 0) {
    ld_pose_detail *dt = reinterpret_cast<ld_pose_detail *>(bb.detail) + pose;
    atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_in_cutoff), s.counters[0]);
    atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_in_cutoff2), s.counters[1]);
    atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_interface_pairs), s.counters[2]);
    atomicAdd(reinterpret_cast<unsigned long long *>(&dt->n_pairs_tested), s.counters[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel 3: epilogue.  One warp per pose.  src/dfire.rs:347-361, src/dna.rs:513-528, src/scoring.rs:21-47.
__device__ __forceinline__ bool lig_bit(const unsigned *ifl, int splits, int words, int j) {
  unsigned w = 0;
  for (int c = 0; c < splits; ++c) w |= ifl[(size_t)c * words + (j >> 5)];
  return (w >> (j & 31)) & 1u;
}
template <bool DETAIL>
__global__ void __launch_bounds__(128) finalize_kernel(const DeviceComplex cx, const BatchBuffers bb, int n_poses) {
  const int pose = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pose >= live_poses(bb, n_poses)) return;
  const unsigned *ifr = bb.iface_rec + (size_t)pose * cx.n_rec_tiles;
  const unsigned *ifl = bb.iface_lig + (size_t)pose * bb.rec_splits * bb.lig_words;
  unsigned hr = 0, hl = 0, hm = 0;
  for (int r = lane; r < cx.n_rec_rst; r += 32) {
    bool hit = false;
    for (int k = cx.rec_rst_off[r]; k < cx.rec_rst_off[r + 1] && !hit; ++k) {
      const int a = cx.rec_rst_idx[k];
      hit = (ifr[a >> 5] >> (a & 31)) & 1u;
    }
    hr += hit;
  }
  for (int r = lane; r < cx.n_lig_rst; r += 32) {
    bool hit = false;
    for (int k = cx.lig_rst_off[r]; k < cx.lig_rst_off[r + 1] && !hit; ++k)
      hit = lig_bit(ifl, bb.rec_splits, bb.lig_words, cx.lig_rst_idx[k]);
    hl += hit;
  }
  for (int k = lane; k < cx.n_membrane; k += 32) {
    const int a = cx.membrane_idx[k];
    hm += (ifr[a >> 5] >> (a & 31)) & 1u;
  }
  hr = warp_sum_u32(hr); hl = warp_sum_u32(hl); hm = warp_sum_u32(hm);
  if (lane != 0) return;
  double s0 = 0.0, s1 = 0.0;
  if (cx.method == 0 && cx.fx_inv_scale != 0.0) {
    // FLEX: each group's sum is an exact 64-bit fixed-point integer (ld_rigid.cuh); converted (one rounding, a power-
    // of-two scale) and added in group order
    const long long *part = reinterpret_cast<const long long *>(bb.partials) + (size_t)pose * cx.n_rec_tiles;
    for (int t = 0; t < cx.n_rec_tiles; ++t) s0 = __dadd_rn(s0, __dmul_rn((double)part[t], cx.fx_inv_scale));
  } else if (cx.method == 0) {
    const double *part = bb.partials + (size_t)pose * cx.n_rec_tiles;
    for (int t = 0; t < cx.n_rec_tiles; ++t) s0 = __dadd_rn(s0, part[t]);
  } else {
    const double *part = bb.partials + (size_t)pose * cx.n_rec_tiles * 2;
    for (int t = 0; t < cx.n_rec_tiles; ++t) {
      s0 = __dadd_rn(s0, part[2 * t]);
      s1 = __dadd_rn(s1, part[2 * t + 1]);
    }
  }
  double score;
  if (cx.method == 0) {
    score = __dmul_rn(__dsub_rn(__dmul_rn(s0, 0.0157), 4.7), -1.0);  // src/dfire.rs:347
  } else {
    const double te = __ddiv_rn(__dmul_rn(s0, 332.0), 4.0);  // src/dna.rs:513
    score = __dmul_rn(__dadd_rn(te, s1), -1.0);            // src/dna.rs:514
  }
  const double pr = cx.n_rec_rst ? __ddiv_rn((double)hr, (double)cx.n_rec_rst) : 0.0;
  const double pl = cx.n_lig_rst ? __ddiv_rn((double)hl, (double)cx.n_lig_rst) : 0.0;
  double pen = 0.0;
  const double inter = cx.n_membrane ? __ddiv_rn((double)hm, (double)cx.n_membrane) : 0.0;
  if (inter > 0.0) pen = __dmul_rn(999.0, inter);  // MEMBRANE_PENALTY_SCORE, src/constants.rs:21
  bb.energies[pose] =
      __dsub_rn(__dadd_rn(__dadd_rn(score, __dmul_rn(pr, score)), __dmul_rn(pl, score)), pen);  // src/dfire.rs:361
  if (DETAIL) {
    ld_pose_detail *dt = reinterpret_cast<ld_pose_detail *>(bb.detail) + pose;
    dt->raw_sum = s0;
    dt->raw_sum2 = s1;
    dt->rec_rst_hit = (int)hr;
    dt->lig_rst_hit = (int)hl;
    dt->membrane_hit = (int)hm;
  }
}

}  // namespace ldb200

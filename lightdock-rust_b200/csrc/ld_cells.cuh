// ld_cells.cuh — ligand-frame cell lists built ON THE DEVICE (ld_create, and every FLEX slack growth).
//
// The lists are the culling structure of the ligand-frame DFIRE path (ld_rigid.cuh): cell c of a uniform grid lists
// every ligand tile (8 atoms) that has an atom within reach_t = 15.01 A + slack_t of the cell's box.  Round 1 built them
// with 16 host threads (70 ms for 1k4c: 820,000 cells, 10 M entries) and uploaded 26 MB; a single-swarm run is
// start-up bound, and the FLEX path rebuilds them whenever the learnt slacks grow, so they are
// now built where they are used:
//   1. cells_count_kernel   one thread per cell: tiles whose bounding box (grown by the reach) misses the cell are
//                           skipped with six compares, the others are tested atom by atom with the SAME f64 box-distance
//                           expression the host builder uses (kept below as the cross-check), never fused;
//   2. a three-kernel exclusive scan of the counts (4096 cells per block) -> {offset, count} per cell;
//   3. cells_fill_kernel    the same walk, writing tile ids in ascending order (every list is sorted).
// Both builders produce identical lists (tests/test_gpu_rigid_path.py::test_device_cell_lists_equal_host_lists).
#pragma once
#include "ld_device.cuh"

namespace ldb200 {

struct CellGrid {
  float g0[3];
  int nc[3];
  double hh;           // cell edge as the kernels' (f - g0) * inv_h implies it
  double base_reach;   // 15.0 + 0.01
};

struct TileBox {  // conservative f32 bounding box of a ligand tile's atoms
  float lo[3], hi[3];
};

__device__ __forceinline__ bool cell_lists_tile(const CellGrid &g, const double *__restrict__ lx,
                                                const double *__restrict__ ly, const double *__restrict__ lz, int n_lig,
                                                int t, double reach2, double bx0, double by0, double bz0) {
  const int j1 = min((t + 1) * LIG_TILE, n_lig);
  for (int j = t * LIG_TILE; j < j1; ++j) {
    const double ax = lx[j], ay = ly[j], az = lz[j];
    const double ez = fmax(0.0, fmax(__dsub_rn(bz0, az), __dsub_rn(az, __dadd_rn(bz0, g.hh))));
    const double ey = fmax(0.0, fmax(__dsub_rn(by0, ay), __dsub_rn(ay, __dadd_rn(by0, g.hh))));
    const double ex = fmax(0.0, fmax(__dsub_rn(bx0, ax), __dsub_rn(ax, __dadd_rn(bx0, g.hh))));
    const double eyz = __dadd_rn(__dmul_rn(ez, ez), __dmul_rn(ey, ey));
    if (eyz > reach2) continue;
    if (!(__dadd_rn(__dmul_rn(ex, ex), eyz) > reach2)) return true;
  }
  return false;
}

// FILL == false: counts[c] = number of tiles listed by cell c.  FILL == true: writes them at cells[c].x.
template <bool FILL>
__global__ void __launch_bounds__(128)
    cells_walk_kernel(const CellGrid g, const double *__restrict__ lx, const double *__restrict__ ly,
                      const double *__restrict__ lz, int n_lig, int n_tiles, const TileBox *__restrict__ boxes,
                      const float *__restrict__ slack, unsigned *__restrict__ counts, const uint2 *__restrict__ cells,
                      unsigned short *__restrict__ flat) {
  const size_t ncell = (size_t)g.nc[0] * g.nc[1] * g.nc[2];
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const int cx = (int)(c % g.nc[0]), cy = (int)((c / g.nc[0]) % g.nc[1]), cz = (int)(c / ((size_t)g.nc[0] * g.nc[1]));
  const double bx0 = __dadd_rn((double)g.g0[0], __dmul_rn((double)cx, g.hh));
  const double by0 = __dadd_rn((double)g.g0[1], __dmul_rn((double)cy, g.hh));
  const double bz0 = __dadd_rn((double)g.g0[2], __dmul_rn((double)cz, g.hh));
  const float fx0 = (float)bx0, fy0 = (float)by0, fz0 = (float)bz0, fh = (float)g.hh;
  unsigned n = 0;
  const unsigned base = FILL ? cells[c].x : 0u;
  for (int t = 0; t < n_tiles; ++t) {
    const TileBox b = boxes[t];
    const float reach = (float)g.base_reach + slack[t];
    // quick reject (conservative: 1e-3 A of slack covers every f32 rounding here)
    const float gx = fmaxf(0.f, fmaxf(b.lo[0] - (fx0 + fh), fx0 - b.hi[0]));
    const float gy = fmaxf(0.f, fmaxf(b.lo[1] - (fy0 + fh), fy0 - b.hi[1]));
    const float gz = fmaxf(0.f, fmaxf(b.lo[2] - (fz0 + fh), fz0 - b.hi[2]));
    const float r = reach + 1.0e-3f;
    if (gx * gx + gy * gy + gz * gz > r * r) continue;
    const double reach_d = __dadd_rn(g.base_reach, (double)slack[t]);
    if (cell_lists_tile(g, lx, ly, lz, n_lig, t, __dmul_rn(reach_d, reach_d), bx0, by0, bz0)) {
      if (FILL) flat[base + n] = (unsigned short)t;
      ++n;
    }
  }
  if (!FILL) counts[c] = n;
}

// ---- exclusive scan of the counts: 4096 cells per block -------------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 16;
__global__ void __launch_bounds__(SCAN_THREADS)
    cells_scan_blocks_kernel(const unsigned *__restrict__ counts, size_t n, unsigned long long *__restrict__ block_sum) {
  __shared__ unsigned long long s[SCAN_THREADS];
  const size_t base = ((size_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
  unsigned long long v = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) v += counts[base + k];
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = SCAN_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sum[blockIdx.x] = s[0];
}
// one block: exclusive scan of the block sums in place; total in block_sum[n_blocks]
__global__ void __launch_bounds__(1024) cells_scan_sums_kernel(unsigned long long *block_sum, int n_blocks) {
  __shared__ unsigned long long s[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const unsigned long long v = i < n_blocks ? block_sum[i] : 0ull;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive Hillis-Steele
      const unsigned long long a = threadIdx.x >= o ? s[threadIdx.x - o] : 0ull;
      __syncthreads();
      s[threadIdx.x] += a;
      __syncthreads();
    }
    if (i < n_blocks) block_sum[i] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sum[n_blocks] = carry;
}
// stats: [0] non-empty cells, [1] longest list
__global__ void __launch_bounds__(SCAN_THREADS)
    cells_scan_write_kernel(const unsigned *__restrict__ counts, size_t n, const unsigned long long *__restrict__ block_sum,
                            uint2 *__restrict__ cells, unsigned *__restrict__ stats) {
  __shared__ unsigned long long s[SCAN_THREADS];
  const size_t base = ((size_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
  unsigned local[SCAN_ITEMS];
  unsigned long long v = 0;
  unsigned nonempty = 0, longest = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    local[k] = base + k < n ? counts[base + k] : 0u;
    v += local[k];
    nonempty += local[k] != 0u;
    longest = max(longest, local[k]);
  }
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < SCAN_THREADS; o <<= 1) {
    const unsigned long long a = threadIdx.x >= o ? s[threadIdx.x - o] : 0ull;
    __syncthreads();
    s[threadIdx.x] += a;
    __syncthreads();
  }
  unsigned long long off = block_sum[blockIdx.x] + s[threadIdx.x] - v;
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) {
      cells[base + k] = make_uint2((unsigned)off, local[k]);
      off += local[k];
    }
  if (nonempty) atomicAdd(&stats[0], nonempty);
  if (longest) atomicMax(&stats[1], longest);
}

}  // namespace ldb200

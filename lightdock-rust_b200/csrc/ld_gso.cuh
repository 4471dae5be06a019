// ld_gso.cuh — the GSO step on the device: everything GSO::run does between two scoring passes
// (src/lib.rs:46-58), for many independent swarms at once, so that a whole run needs no host round trip per step.
//
//   Swarm::update_luciferin   src/swarm.rs:66-70, src/glowworm.rs:61-72   luciferin = (1 - rho) luciferin + gamma scoring
//   Swarm::movement_phase     src/swarm.rs:72-126                         snapshot, neighbour search, probabilities,
//                                                                         roulette with ONE draw per glowworm, move
//   Glowworm::move_towards    src/glowworm.rs:128-190                     translation step, slerp (src/qt.rs:67-91), ANM
//   Glowworm::update_vision_range  src/glowworm.rs:91-96
//   StdRng (rand 0.7.3 = ChaCha20, seed_from_u64)  src/lib.rs:38, src/swarm.rs:118  -- evaluated on the device: the
//                             stream is counter based, so draw k of a swarm is words 2k, 2k+1 of its key stream
//
// One CTA per swarm, one thread per glowworm.  All arithmetic is f64 in the reference's operation order (the library is
// built with -fmad=false, divisions and square roots are IEEE), so every DISCRETE decision (neighbour sets, roulette
// choice, moved flags) is the host's unless a comparison falls within the last-bit difference between CUDA's and the
// host libm's acos/sin inside slerp -- the one place where the device and the host can round differently.
//
// The step kernel also prepares the next scoring pass: only glowworms that moved are rescored (src/glowworm.rs:62), so
// it packs their pose rows into a contiguous batch (row order is irrelevant: the pair kernels are batch-invariant) and
// leaves the row count in device memory, where the scoring kernels read it (BatchBuffers::n_live).
#pragma once
#include "ld_device.cuh"

namespace ldb200 {

constexpr int GSO_MAX_GLOWWORMS = 1024;  // one thread per glowworm

struct GsoState {
  int n_swarms, n_glow, pose_len;
  int n_rec_ext, n_lig_ext;      // ANM extents per pose row (0 unless use_anm)
  double *poses[2];              // [G][pose_len], double-buffered: a glowworm moves towards where its neighbour WAS
  double *luciferin, *vision, *scoring;  // [G]
  int *n_neighbors;              // [G]
  int *slot;                     // [G] row of the glowworm in the packed batch being scored, -1 = not rescored
  double *packed;                // [G][pose_len] rows to (re)score
  double *energies;              // [G] their energies
  int *n_packed;                 // [max_steps + 1]: rows packed for the scoring pass that FOLLOWS step s
  int *failed;                   // [n_swarms] 0, or the step at which the swarm hit what is a panic in the reference
  const uint32_t *keys;          // [n_swarms][8] ChaCha20 key of the swarm's StdRng
};

__device__ __forceinline__ uint32_t gso_rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
#define LD_QR(a, b, c, d)                                                                \
  a += b; d = gso_rotl(d ^ a, 16); c += d; b = gso_rotl(b ^ c, 12); a += b; d = gso_rotl(d ^ a, 8); c += d; \
  b = gso_rotl(b ^ c, 7);
// gen::<f64>() number `k` (0-based) of the StdRng with this key: (next_u64() >> 11) * 2^-53, next_u64 = words 2k (low)
// and 2k+1 (high) of the ChaCha20 key stream (64-bit block counter from 0, stream id 0, 16 words per block).
__device__ inline double gso_draw(const uint32_t *key, unsigned long long k) {
  const unsigned long long block = k >> 3;
  uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                     key[4], key[5], key[6], key[7], (uint32_t)block, (uint32_t)(block >> 32), 0u, 0u};
  uint32_t x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3], x4 = in[4], x5 = in[5], x6 = in[6], x7 = in[7], x8 = in[8],
           x9 = in[9], x10 = in[10], x11 = in[11], x12 = in[12], x13 = in[13], x14 = in[14], x15 = in[15];
#pragma unroll 1
  for (int round = 0; round < 20; round += 2) {
    LD_QR(x0, x4, x8, x12) LD_QR(x1, x5, x9, x13) LD_QR(x2, x6, x10, x14) LD_QR(x3, x7, x11, x15)
    LD_QR(x0, x5, x10, x15) LD_QR(x1, x6, x11, x12) LD_QR(x2, x7, x8, x13) LD_QR(x3, x4, x9, x14)
  }
  const uint32_t out[16] = {x0 + in[0], x1 + in[1], x2 + in[2], x3 + in[3], x4 + in[4], x5 + in[5], x6 + in[6],
                            x7 + in[7], x8 + in[8], x9 + in[9], x10 + in[10], x11 + in[11], x12 + in[12],
                            x13 + in[13], x14 + in[14], x15 + in[15]};
  const int w = (int)(k & 7ull) * 2;
  uint32_t lo = 0u, hi = 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (w == 2 * i) { lo = out[2 * i]; hi = out[2 * i + 1]; }
  const unsigned long long u = (unsigned long long)lo | ((unsigned long long)hi << 32);
  return (double)(u >> 11) * (1.0 / 9007199254740992.0);
}
#undef LD_QR

struct GsoQuat { double w, x, y, z; };
__device__ __forceinline__ void gso_normalize(GsoQuat &q) {  // src/qt.rs:36-46
  const double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
// Quaternion::slerp, src/qt.rs:67-91
__device__ inline GsoQuat gso_slerp(GsoQuat q1, GsoQuat q2, double t) {
  gso_normalize(q1);
  gso_normalize(q2);
  double q_dot = q1.w * q2.w + q1.x * q2.x + q1.y * q2.y + q1.z * q2.z;
  if (q_dot < 0.0) {  // avoid the long path
    q1.w = -q1.w; q1.x = -q1.x; q1.y = -q1.y; q1.z = -q1.z;
    q_dot *= -1.0;
  }
  if (q_dot > 0.9995) {  // LINEAR_THRESHOLD, src/constants.rs:11
    GsoQuat r = {q1.w + t * (q2.w - q1.w), q1.x + t * (q2.x - q1.x), q1.y + t * (q2.y - q1.y), q1.z + t * (q2.z - q1.z)};
    gso_normalize(r);
    return r;
  }
  q_dot = fmax(fmin(q_dot, 1.0), -1.0);
  const double omega = acos(q_dot);
  const double so = sin(omega);
  const double a = sin((1.0 - t) * omega) / so, b = sin(t * omega) / so;
  const GsoQuat r = {a * q1.w + b * q2.w, a * q1.x + b * q2.x, a * q1.y + b * q2.y, a * q1.z + b * q2.z};
  return r;
}
// the ANM part of move_towards, src/glowworm.rs:154-187
__device__ inline void gso_step_towards(double *mine_out, const double *mine, const double *other, int n, double step) {
  double cum_norm = 0.0;
  for (int i = 0; i < n; ++i) {
    const double diff = other[i] - mine[i];
    cum_norm += diff * diff;
  }
  const double coef = step / sqrt(cum_norm);
  for (int i = 0; i < n; ++i) {
    double diff = other[i] - mine[i];
    diff *= coef;
    mine_out[i] = mine[i] + diff;
  }
}

// One GSO step of every swarm: update_luciferin (with the energies of the scoring pass that just ran), movement
// phase, vision range, and the packed batch for the next scoring pass.  grid = swarms, block >= glowworms per swarm.
// `step` is 1-based as in GSO::run; `cur` selects the buffer holding the poses the scoring pass just saw.
__global__ void __launch_bounds__(GSO_MAX_GLOWWORMS) gso_step_kernel(const GsoState st, int step, int cur) {
  extern __shared__ double gso_smem[];  // x[n] y[n] z[n] luciferin[n]
  __shared__ int s_warp_count[32];
  __shared__ int s_base;
  const int n = st.n_glow, s = blockIdx.x, i = threadIdx.x;
  if (st.failed[s] != 0) return;  // a swarm that "panicked" stays as it was, the others go on
  double *sx = gso_smem, *sy = sx + n, *sz = sy + n, *sl = sz + n;
  const size_t gi = (size_t)s * n + i;
  const int pl = st.pose_len;
  const double *P = st.poses[cur];
  double *Q = st.poses[cur ^ 1];
  const bool live = i < n;
  double lum = 0.0, vis = 0.0;
  if (live) {
    // Glowworm::compute_luciferin, src/glowworm.rs:61-72
    const int sl_row = st.slot[gi];
    if (sl_row >= 0) st.scoring[gi] = st.energies[sl_row];
    const double rho = 0.5, gamma = 0.4;
    lum = (1.0 - rho) * st.luciferin[gi] + gamma * st.scoring[gi];
    st.luciferin[gi] = lum;
    vis = st.vision[gi];
    sx[i] = P[gi * pl]; sy[i] = P[gi * pl + 1]; sz[i] = P[gi * pl + 2];
    sl[i] = lum;
  }
  __syncthreads();
  int nid = i, cnt = 0;
  bool panic = false;
  if (live) {
    // neighbours (src/swarm.rs:88-103) and the sum of the luciferin differences (src/glowworm.rs:98-107), j ascending
    const double x1 = sx[i], y1 = sy[i], z1 = sz[i];
    double total_sum = 0.0;
    for (int j = 0; j < n; ++j) {
      if (j == i || !(lum < sl[j])) continue;
      const double x2 = sx[j], y2 = sy[j], z2 = sz[j];
      const double d = sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
      if (d < vis) {
        ++cnt;
        total_sum += sl[j] - lum;
      }
    }
    // one draw per glowworm and step, whether or not it has neighbours (src/swarm.rs:118)
    const double r = gso_draw(st.keys + (size_t)s * 8, (unsigned long long)(step - 1) * (unsigned long long)n + (unsigned)i);
    if (cnt > 0) {
      // select_random_neighbor, src/glowworm.rs:114-126: while sum < r { sum += p[k]; k += 1 } -> neighbors[k - 1]
      double sum = 0.0;
      int sel = -1, taken = 0;
      for (int j = 0; j < n && taken < cnt; ++j) {
        if (j == i || !(lum < sl[j])) continue;
        const double x2 = sx[j], y2 = sy[j], z2 = sz[j];
        const double d = sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
        if (!(d < vis)) continue;
        if (!(sum < r)) break;
        sum += (sl[j] - lum) / total_sum;
        sel = j;
        ++taken;
      }
      // k == 0 (r == 0) underflows the index, running off the end of the probabilities overruns it: both panic
      if (sel < 0 || (taken == cnt && sum < r)) panic = true;
      else nid = sel;
    }
  }
  // a panic ends this swarm at this step (one process per swarm in the reference); nothing of the step is kept
  if (__syncthreads_or(panic ? 1 : 0)) {
    if (i == 0) st.failed[s] = step;
    return;
  }
  const bool moved = live && nid != i;
  if (live) {
    const double *mine = P + gi * pl;
    double *out = Q + gi * pl;
    if (!moved) {
      for (int k = 0; k < pl; ++k) out[k] = mine[k];
    } else {
      // Glowworm::move_towards, src/glowworm.rs:128-190
      const double *other = P + ((size_t)s * n + nid) * pl;
      double dx = other[0] - mine[0], dy = other[1] - mine[1], dz = other[2] - mine[2];
      const double norm = sqrt(dx * dx + dy * dy + dz * dz);
      const double coef = 0.5 / norm;  // DEFAULT_TRANSLATION_STEP
      dx *= coef; dy *= coef; dz *= coef;
      out[0] = mine[0] + dx; out[1] = mine[1] + dy; out[2] = mine[2] + dz;
      const GsoQuat q1 = {mine[3], mine[4], mine[5], mine[6]}, q2 = {other[3], other[4], other[5], other[6]};
      const GsoQuat q = gso_slerp(q1, q2, 0.5);  // DEFAULT_ROTATION_STEP
      out[3] = q.w; out[4] = q.x; out[5] = q.y; out[6] = q.z;
      if (st.n_rec_ext > 0) gso_step_towards(out + 7, mine + 7, other + 7, st.n_rec_ext, 0.5);  // DEFAULT_NMODES_STEP
      if (st.n_lig_ext > 0)
        gso_step_towards(out + 7 + st.n_rec_ext, mine + 7 + st.n_rec_ext, other + 7 + st.n_rec_ext, st.n_lig_ext, 0.5);
    }
    // update_vision_range, src/glowworm.rs:91-96 (beta 0.08, max_neighbors 5, max_vision_range 5)
    st.vision[gi] = fmin(5.0, fmax(0.0, vis + 0.08 * (double)(5 - cnt)));
    st.n_neighbors[gi] = cnt;
  }
  // rows of the next scoring pass: the glowworms that moved
  const unsigned ballot = __ballot_sync(0xffffffffu, moved);
  const int lane = i & 31, warp = i >> 5;
  if (lane == 0) s_warp_count[warp] = __popc(ballot);
  __syncthreads();
  if (i == 0) {
    int total = 0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) {
      const int c = s_warp_count[w];
      s_warp_count[w] = total;
      total += c;
    }
    s_base = total ? atomicAdd(st.n_packed + step, total) : 0;
  }
  __syncthreads();
  if (live) {
    int row = -1;
    if (moved) {
      row = s_base + s_warp_count[warp] + __popc(ballot & ((1u << lane) - 1u));
      const double *src = Q + gi * pl;
      double *dst = st.packed + (size_t)row * pl;
      for (int k = 0; k < pl; ++k) dst[k] = src[k];
    }
    st.slot[gi] = row;
  }
}

}  // namespace ldb200

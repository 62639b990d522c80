// CUDA kernels of the sparse Levenberg-Marquardt inner loop (sm_100a).
//
//   K1 linearize_kernel<Kind>      factor residual/Jacobian (generated fp64 device functions), fused
//                                  J^T J / J^T r block scatter + 0.5|r|^2 partial sums
//                                  -> replaces Linearizer::Relinearize (symforce/opt/linearizer.cc:55-120, 359-433)
//   K2 damping_kernel              DampHessian (levenberg_marquardt_solver.tcc:23-54) as a vector; H is never
//                                  modified in place, so no save/restore of the diagonal is needed
//   K3 schur_*                     C^-1 per landmark in registers, S = B - E C^-1 E^T per block from
//                                  precomputed match lists, reduced rhs, back-substitution
//                                  -> replaces SparseSchurSolver::Factorize/Solve (sparse_schur_solver.tcc:101-162)
//   K4 front_factor_kernel         multifrontal supernodal Cholesky, one CTA per front, level scheduled
//                                  -> replaces SparseCholeskySolver::Factorize (sparse_cholesky_solver.tcc:109-219)
//   K5 front_solve_*               supernodal forward/backward substitution (…tcc:232-259)
//   K6 retract_kernel              Values::Retract (symforce/opt/values.cc:315-327)
//   K7 lm_* / *_reduce             gain ratio, accept/reject, lambda update, state-block bookkeeping on the
//                                  device (levenberg_marquardt_solver.tcc:139-343) -- no host round trip
#include <cstdio>

#include "gen/factors_gen.cuh"
#include "kernels.cuh"

namespace sfx {

// Streams that are far larger than L2 and not re-read soon (E blocks, per-observation point contributions) are
// written / read with the evict-first hint so they do not push the reused data (camera values, W, point blocks) out
#ifdef SFX_NO_STREAM_HINTS
#define SFX_ST_STREAM(p, v) (*(p) = (v))
#define SFX_LD_STREAM(p) (*(p))
#else
#define SFX_ST_STREAM(p, v) __stcs((p), (v))
#define SFX_LD_STREAM(p) __ldcs(p)
#endif

int64_t g_launches = 0;  // kernels launched by this library (bench.py reports it)

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block sum (result valid in thread 0)
template <int THREADS>
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[THREADS / 32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) r += sh[i];
  }
  return r;
}

// mode 0: the Init block (skipped when its linearization is valid), 1: the New block, 2: block kEagerBlock
// unconditionally and whatever the control block says (eager linearization of freshly uploaded values, sfx_set_values)
constexpr int kModeEager = 2;
__device__ __forceinline__ int sel_block(const Ctrl* c, int mode) {
  return mode == kModeEager ? kEagerBlock : (mode == 0 ? c->init_idx : c->new_idx);
}

// ------------------------------------------------------------------------------------------------
// K1: linearize
// ------------------------------------------------------------------------------------------------
template <int KIND>
struct Kind;

#define SFX_KIND_BEGIN(ID, R_, T_, NOPT_, NUSED_)   \
  template <>                                       \
  struct Kind<ID> {                                 \
    static constexpr int R = R_, T = T_, NOPT = NOPT_, NUSED = NUSED_;

// arg pointers are given for USED args only, in argument order
SFX_KIND_BEGIN(SFX_KIND_SNAVELY, 2, 12, 3, 4)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 6 : (k == 1 ? 3 : 3); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 6 : 9); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_snavely(a[0], a[1], a[2], a[3], nullptr, res, J);
  }
};
SFX_KIND_BEGIN(SFX_KIND_BETWEEN_POSE3, 6, 12, 2, 5)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 6 : (k == 1 ? 6 : 0); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 6 : 12); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_between_pose3(a[0], a[1], a[2], a[3], a[4], res, J);
  }
};
SFX_KIND_BEGIN(SFX_KIND_PRIOR_POSE3, 6, 6, 1, 4)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 6 : (k == 1 ? 0 : 0); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 6 : 6); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_prior_pose3(a[0], a[1], a[2], a[3], res, J);
  }
};
SFX_KIND_BEGIN(SFX_KIND_MATCHING, 3, 6, 1, 4)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 6 : (k == 1 ? 0 : 0); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 6 : 6); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_matching(a[0], a[1], a[2], a[3], res, J);
  }
};
SFX_KIND_BEGIN(SFX_KIND_ODOMETRY, 6, 12, 2, 5)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 6 : (k == 1 ? 6 : 0); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 6 : 12); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_odometry(a[0], a[1], a[2], a[3], a[4], res, J);
  }
};
SFX_KIND_BEGIN(SFX_KIND_IRL_LINEAR_GNC, 2, 13, 3, 11)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 6 : (k == 1 ? 6 : 1); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 6 : 12); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_irl_linear_gnc(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], res, J);
  }
};
SFX_KIND_BEGIN(SFX_KIND_IRL_PRIOR, 1, 1, 1, 5)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 1 : (k == 1 ? 0 : 0); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 1 : 1); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_irl_prior(a[0], a[1], a[2], a[3], a[4], res, J);
  }
};
SFX_KIND_BEGIN(SFX_KIND_BETWEEN_ROT3, 3, 6, 2, 5)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 3 : (k == 1 ? 3 : 0); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 3 : 6); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_between_rot3(a[0], a[1], a[2], a[3], a[4], res, J);
  }
};
SFX_KIND_BEGIN(SFX_KIND_PRIOR_ROT3, 3, 3, 1, 4)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 3 : (k == 1 ? 0 : 0); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 3 : 3); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_prior_rot3(a[0], a[1], a[2], a[3], res, J);
  }
};
SFX_KIND_BEGIN(SFX_KIND_BARRON, 5, 5, 1, 4)
  __host__ __device__ static constexpr int dim(int k) { return k == 0 ? 5 : (k == 1 ? 0 : 0); }
  __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : (k == 1 ? 5 : 5); }
  __device__ static void eval(const double* const* a, double* res, double* J) {
    sfx_factor_barron(a[0], a[1], a[2], a[3], res, J);
  }
};

constexpr int kLinThreads = 128;

template <int KIND>
__global__ void __launch_bounds__(kLinThreads) linearize_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp, LinBatch b,
                                                                 int mode, double* __restrict__ partials) {
  using K = Kind<KIND>;
  if (ctrl->done) return;
  const int blk = sel_block(ctrl, mode);
  if (mode == 0 && ctrl->lin_valid[blk]) return;
  const double* __restrict__ values = sp.values[blk];
  double* __restrict__ H = sp.H[blk];
  double* __restrict__ rhs = sp.rhs[blk];
  double* __restrict__ resid = sp.res[blk];

  const int s = blockIdx.x * kLinThreads + threadIdx.x;
  double err = 0.0;
  if (s < b.n) {
    const double* a[K::NUSED];
#pragma unroll
    for (int u = 0; u < K::NUSED; ++u) a[u] = values + __ldg(b.arg_off + (size_t)u * b.n + s);
    double res[K::R];
    double J[K::R * K::T];
    K::eval(a, res, J);
    const int ro = __ldg(b.res_off + s);
#pragma unroll
    for (int q = 0; q < K::R; ++q) {
      resid[ro + q] = res[q];
      err += res[q] * res[q];
    }
    int rhs_base[3], diag_base[3];
#pragma unroll
    for (int g = 0; g < 3; ++g)
      if (g < b.n_groups) {
        rhs_base[g] = __ldg(b.rhs_off + (size_t)g * b.n + s);
        diag_base[g] = __ldg(b.diag_off + (size_t)g * b.n + s);
      }
    uint32_t off_base[3];
    {
      const int npairs = b.n_groups * (b.n_groups - 1) / 2;
#pragma unroll
      for (int p = 0; p < 3; ++p)
        if (p < npairs) off_base[p] = __ldg(b.off_off + (size_t)p * b.n + s);
    }
#pragma unroll
    for (int ka = 0; ka < K::NOPT; ++ka) {
      const int g = b.key_group[ka];
      if (g < 0) continue;
      const int sa = b.key_sub[ka];
      // rhs = J^T r
#pragma unroll
      for (int r = 0; r < K::dim(ka); ++r) {
        double v = 0;
#pragma unroll
        for (int q = 0; q < K::R; ++q) v += J[q + (K::col(ka) + r) * K::R] * res[q];
        atomicAdd(rhs + rhs_base[g] + sa + r, v);
      }
#pragma unroll
      for (int kb = 0; kb <= ka; ++kb) {
        const int h = b.key_group[kb];
        if (h < 0) continue;
        const int sb = b.key_sub[kb];
        if (g == h) {
          const int ld = b.group_dim[g];
          double* dst = H + diag_base[g];
          if (ka == kb) {
#pragma unroll
            for (int c = 0; c < K::dim(ka); ++c)
#pragma unroll
              for (int r = c; r < K::dim(ka); ++r) {
                double v = 0;
#pragma unroll
                for (int q = 0; q < K::R; ++q) v += J[q + (K::col(ka) + r) * K::R] * J[q + (K::col(ka) + c) * K::R];
                atomicAdd(dst + (sa + r) + (size_t)(sa + c) * ld, v);
              }
          } else {
            const bool a_is_row = sa > sb;
#pragma unroll
            for (int c = 0; c < K::dim(kb); ++c)
#pragma unroll
              for (int r = 0; r < K::dim(ka); ++r) {
                double v = 0;
#pragma unroll
                for (int q = 0; q < K::R; ++q) v += J[q + (K::col(ka) + r) * K::R] * J[q + (K::col(kb) + c) * K::R];
                const size_t o = a_is_row ? (size_t)(sa + r) + (size_t)(sb + c) * ld
                                          : (size_t)(sb + c) + (size_t)(sa + r) * ld;
                atomicAdd(dst + o, v);
              }
          }
        } else {
          const int G = g > h ? g : h, Hh = g > h ? h : g;
          const uint32_t ob = off_base[G * (G - 1) / 2 + Hh];
          double* dst = H + (ob & 0x3fffffffu);
          const bool excl = (ob >> 31) != 0;
          const bool tr = ((ob >> 30) & 1u) != 0;
          // row side of the stored block: group G unless transposed
          const bool a_is_row = ((g == G) != tr);
          const int ld = a_is_row ? b.group_dim[g] : b.group_dim[h];
#pragma unroll
          for (int c = 0; c < K::dim(kb); ++c)
#pragma unroll
            for (int r = 0; r < K::dim(ka); ++r) {
              double v = 0;
#pragma unroll
              for (int q = 0; q < K::R; ++q) v += J[q + (K::col(ka) + r) * K::R] * J[q + (K::col(kb) + c) * K::R];
              const size_t o = a_is_row ? (size_t)(sa + r) + (size_t)(sb + c) * ld
                                        : (size_t)(sb + c) + (size_t)(sa + r) * ld;
              if (excl)
                dst[o] = v;
              else
                atomicAdd(dst + o, v);
            }
        }
      }
    }
  }
  const double tot = block_sum<kLinThreads>(err);
  if (threadIdx.x == 0) partials[b.partial_base + blockIdx.x] = tot;
}

// ---- K1, BAL fast path ---------------------------------------------------------------------------
// Snavely factors whose camera pose + intrinsics are one 9-dim node and whose point is a 3-dim
// node.  Slots are camera-major, so a warp almost always works on ONE camera: its 45 + 9 camera
// block/rhs values are summed across the warp with a transpose-reduce (62 shuffles per 32 values,
// every lane ends up owning one total) and leave the SM as 54 atomics per warp instead of 1728.
// The 3x9 point-camera block of every observation is owned by that observation: it is staged in
// shared memory and written as one contiguous 6912-byte run per warp.  Point blocks go out as
// fp64 RED atomics (9 per observation).
__device__ __forceinline__ double transpose_reduce32(double (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16, n = 16; o >= 1; o >>= 1, n >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < n) {
        const double lo = v[i], hi = v[i + n];
        const double send = up ? lo : hi;
        const double keep = up ? hi : lo;
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
  }
  return v[0];  // total of value index: bit-reversed-free mapping, see value_of_lane()
}
// which of the 32 input values a lane owns after transpose_reduce32
__device__ __forceinline__ int value_of_lane(int lane) {
  return ((lane & 16) ? 16 : 0) + ((lane & 8) ? 8 : 0) + ((lane & 4) ? 4 : 0) + ((lane & 2) ? 2 : 0) + (lane & 1);
}

constexpr int kBalThreads = 128;
// per-observation point record: the 2 x 3 point Jacobian and the residual, 8 doubles = two whole 32-byte sectors of a
// 64-byte aligned slot.  bal_point_finalize_kernel forms J_p^T J_p and J_p^T r from it (18 FMAs per observation) with two
// 32-byte loads; the 9-double record of the products it replaces made every gathered slot touch three or four sectors
// (0.70 GB of DRAM reads for 0.36 GB of records).
constexpr int kPbufStride = 8;
#ifndef SFX_BAL_MINB
#define SFX_BAL_MINB 5  // 96 registers, 68 B of spills: 0.83 -> 0.80 ms at final-shape (6: 80 registers, slower)
#endif
// SKIP (timing experiments only, results invalid when != 0): bit0 camera block, bit1 point block,
// bit2 E block, bit3 factor arithmetic
template <int SKIP>
__global__ void __launch_bounds__(kBalThreads, SFX_BAL_MINB) linearize_bal_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                                    LinBatch b, int mode,
                                                                    double* __restrict__ partials, int block0) {
  __shared__ double stage[kBalThreads / 32][32 * 27];
  if (mode != kModeEager && ctrl->done) return;
  const int blk = sel_block(ctrl, mode);
  if (mode == 0 && ctrl->lin_valid[blk]) return;
  const double* __restrict__ values = sp.values[blk];
  double* __restrict__ H = sp.H[blk];
  double* __restrict__ rhs = sp.rhs[blk];
  double* __restrict__ resid = sp.res[blk];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cta = blockIdx.x + block0;  // (block0 > 0: the launch covers a sub-range of the batch)
  const int s = cta * kBalThreads + threadIdx.x;
  const bool valid = s < b.n;
  const int sc = valid ? s : b.n - 1;
  double res[2], J[24];
  {
    const double* a0 = values + __ldg(b.arg_off + sc);
    const double* a1 = values + __ldg(b.arg_off + (size_t)b.n + sc);
    const double* a2 = values + __ldg(b.arg_off + (size_t)2 * b.n + sc);
    const double* a3 = values + __ldg(b.arg_off + (size_t)3 * b.n + sc);
    if (SKIP & 8) {
      res[0] = a0[0] + a1[0] + a2[0] + a3[0];
      res[1] = a0[6] + a1[2] + a2[2] + a3[1];
#pragma unroll
      for (int i = 0; i < 24; ++i) J[i] = res[i & 1] * (i + 1);
    } else {
      sfx_factor_snavely(a0, a1, a2, a3, nullptr, res, J);
    }
  }
  if (!valid) {
    res[0] = res[1] = 0.0;
#pragma unroll
    for (int i = 0; i < 24; ++i) J[i] = 0.0;
  }
  double err = res[0] * res[0] + res[1] * res[1];
  if (valid) {
    const int ro = __ldg(b.res_off + s);
    resid[ro] = res[0];
    resid[ro + 1] = res[1];
  }
  const int cam_rhs = __ldg(b.rhs_off + sc);
  const int pt_rhs = __ldg(b.rhs_off + (size_t)b.n + sc);
  const int cam_diag = __ldg(b.diag_off + sc);
  const int pt_diag = __ldg(b.diag_off + (size_t)b.n + sc);
  const uint32_t eo = __ldg(b.off_off + sc);
  // ---- camera block: 45 lower entries (column-major packed) + 9 rhs ------------------------------
  const int key0 = __shfl_sync(0xffffffffu, cam_diag, 0);
  const bool uniform = __all_sync(0xffffffffu, !valid || cam_diag == key0);
  if (SKIP & 1) {
  } else if (uniform) {
    double v[32];
    // batch 0: packed entries 0..31
    {
      int idx = 0;
#pragma unroll
      for (int c = 0; c < 9; ++c)
#pragma unroll
        for (int r = c; r < 9; ++r) {
          if (idx < 32) v[idx] = J[2 * r] * J[2 * c] + J[2 * r + 1] * J[2 * c + 1];
          ++idx;
        }
    }
    const double t0 = transpose_reduce32(v);
    // batch 1: packed entries 32..44, then rhs 0..8 (22 values), rest zero
    {
      int idx = 0;
#pragma unroll
      for (int c = 0; c < 9; ++c)
#pragma unroll
        for (int r = c; r < 9; ++r) {
          if (idx >= 32) v[idx - 32] = J[2 * r] * J[2 * c] + J[2 * r + 1] * J[2 * c + 1];
          ++idx;
        }
#pragma unroll
      for (int r = 0; r < 9; ++r) v[13 + r] = J[2 * r] * res[0] + J[2 * r + 1] * res[1];
#pragma unroll
      for (int i = 22; i < 32; ++i) v[i] = 0.0;
    }
    const double t1 = transpose_reduce32(v);
    const int vi = value_of_lane(lane);
    // packed index -> (r, c)
    {
      int idx = vi, c = 0;
      while (idx >= 9 - c) {
        idx -= 9 - c;
        ++c;
      }
      atomicAdd(H + key0 + (c + idx) + c * 9, t0);
    }
    if (vi < 13) {
      int idx = vi + 32, c = 0;
      while (idx >= 9 - c) {
        idx -= 9 - c;
        ++c;
      }
      atomicAdd(H + key0 + (c + idx) + c * 9, t1);
    }
    {
      const int rb = __shfl_sync(0xffffffffu, cam_rhs, 0);
      if (vi >= 13 && vi < 22) atomicAdd(rhs + rb + (vi - 13), t1);
    }
  } else if (valid) {
#pragma unroll
    for (int c = 0; c < 9; ++c)
#pragma unroll
      for (int r = c; r < 9; ++r)
        atomicAdd(H + cam_diag + r + c * 9, J[2 * r] * J[2 * c] + J[2 * r + 1] * J[2 * c + 1]);
#pragma unroll
    for (int r = 0; r < 9; ++r) atomicAdd(rhs + cam_rhs + r, J[2 * r] * res[0] + J[2 * r + 1] * res[1]);
  }
  // ---- point block ---------------------------------------------------------------------------------
  if (b.pbuf != nullptr && !(SKIP & 2)) {
    // per-observation contribution (6 lower entries + 3 rhs) written coalesced; summed per point by
    // bal_point_finalize_kernel (no atomics on the 12 scattered point values of every observation)
    double* st = stage[warp];
#pragma unroll
    for (int q = 0; q < 6; ++q) st[lane * kPbufStride + q] = J[18 + q];
    st[lane * kPbufStride + 6] = res[0];
    st[lane * kPbufStride + 7] = res[1];
    __syncwarp();
    double* dst = b.pbuf + (size_t)(cta * kBalThreads + warp * 32) * kPbufStride;
#pragma unroll
    for (int i = 0; i < kPbufStride; ++i) SFX_ST_STREAM(dst + i * 32 + lane, st[i * 32 + lane]);
    __syncwarp();
  } else if (valid && !(SKIP & 2)) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = c; r < 3; ++r)
        atomicAdd(H + pt_diag + r + c * 3,
                  J[2 * (9 + r)] * J[2 * (9 + c)] + J[2 * (9 + r) + 1] * J[2 * (9 + c) + 1]);
#pragma unroll
    for (int r = 0; r < 3; ++r) atomicAdd(rhs + pt_rhs + r, J[2 * (9 + r)] * res[0] + J[2 * (9 + r) + 1] * res[1]);
  }
  // ---- E block (point rows x camera cols), owned by the observation --------------------------------
  if (!(SKIP & 4)) {
    const uint32_t off = eo & 0x3fffffffu;
    const bool excl = (eo >> 31) != 0;
    const uint32_t off0 = __shfl_sync(0xffffffffu, off, 0);
    const bool contiguous = __all_sync(0xffffffffu, valid && excl && off == off0 + 27u * lane);
    if (contiguous) {
      double* st = stage[warp];
#pragma unroll
      for (int c = 0; c < 9; ++c)
#pragma unroll
        for (int a = 0; a < 3; ++a)
          st[lane * 27 + a + 3 * c] = J[2 * (9 + a)] * J[2 * c] + J[2 * (9 + a) + 1] * J[2 * c + 1];
      __syncwarp();
      double* dst = H + off0;
#pragma unroll
      for (int i = 0; i < 27; ++i) SFX_ST_STREAM(dst + i * 32 + lane, st[i * 32 + lane]);
    } else if (valid) {
      double* dst = H + off;
#pragma unroll
      for (int c = 0; c < 9; ++c)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double v = J[2 * (9 + a)] * J[2 * c] + J[2 * (9 + a) + 1] * J[2 * c + 1];
          if (excl)
            dst[a + 3 * c] = v;
          else
            atomicAdd(dst + a + 3 * c, v);
        }
    }
  }
  const double tot = block_sum<kBalThreads>(err);
  if (threadIdx.x == 0) partials[b.partial_base + cta] = tot;
}

// sums the per-observation point contributions of linearize_bal_kernel in slot order (deterministic)
// and adds the total to the point's diagonal block and rhs
__global__ void __launch_bounds__(128) bal_point_finalize_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp, LinBatch b,
                                                                 int mode) {
  if (mode != kModeEager && ctrl->done) return;
  const int blk = sel_block(ctrl, mode);
  if (mode == 0 && ctrl->lin_valid[blk]) return;
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= b.n_pf) return;
  double a[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) a[i] = 0.0;
  const int q1 = __ldg(b.pf_ptr + pt + 1);
  for (int q = __ldg(b.pf_ptr + pt); q < q1; ++q) {
    const double* __restrict__ src = b.pbuf + (size_t)__ldg(b.pf_slot + q) * kPbufStride;
    double j[8];  // J[2 c + i] = d res_i / d x_c for the point columns c = 0..2, then res
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(j[0]), "=d"(j[1]), "=d"(j[2]), "=d"(j[3]) : "l"(src));
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(j[4]), "=d"(j[5]), "=d"(j[6]), "=d"(j[7]) : "l"(src + 4));
    int o = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = c; r < 3; ++r) a[o++] += j[2 * r] * j[2 * c] + j[2 * r + 1] * j[2 * c + 1];
#pragma unroll
    for (int r = 0; r < 3; ++r) a[6 + r] += j[2 * r] * j[6] + j[2 * r + 1] * j[7];
  }
  double* Hd = sp.H[blk] + __ldg(b.pf_diag + pt);
  double* rh = sp.rhs[blk] + __ldg(b.pf_rhs + pt);
  if (b.pf_exclusive) {
    // no other batch touches these point blocks (zeroed before): plain stores
    Hd[0] = a[0];
    Hd[1] = a[1];
    Hd[2] = a[2];
    Hd[4] = a[3];
    Hd[5] = a[4];
    Hd[8] = a[5];
    rh[0] = a[6];
    rh[1] = a[7];
    rh[2] = a[8];
  } else {
    atomicAdd(Hd + 0, a[0]);
    atomicAdd(Hd + 1, a[1]);
    atomicAdd(Hd + 2, a[2]);
    atomicAdd(Hd + 4, a[3]);
    atomicAdd(Hd + 5, a[4]);
    atomicAdd(Hd + 8, a[5]);
    atomicAdd(rh + 0, a[6]);
    atomicAdd(rh + 1, a[7]);
    atomicAdd(rh + 2, a[8]);
  }
}

__global__ void zero_lin_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp, int mode, int64_t n_h, int n_rhs) {
  if (mode != kModeEager && ctrl->done) return;
  const int blk = sel_block(ctrl, mode);
  if (mode == 0 && ctrl->lin_valid[blk]) return;
  double* H = sp.H[blk];
  double* rhs = sp.rhs[blk];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_h; i += stride) H[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rhs; i += stride) rhs[i] = 0.0;
}

// sums the per-CTA partials deterministically; err[target] = 0.5 * sum
__global__ void finish_error_kernel(Ctrl* ctrl, const double* __restrict__ partials, int n, int mode,
                                    double* __restrict__ eager_err) {
  if (mode != kModeEager && ctrl->done) return;
  const int blk = sel_block(ctrl, mode);
  if (mode == 0 && ctrl->lin_valid[blk]) return;
  double v = 0;
  for (int i = threadIdx.x; i < n; i += 1024) v += partials[i];
  const double tot = block_sum<1024>(v);
  if (threadIdx.x == 0) {
    if (mode == kModeEager)
      *eager_err = 0.5 * tot;  // kept beside the control block (which the next Optimize resets) until it is adopted
    else
      ctrl->red[0] = 0.5 * tot;  // this rank's part; committed by commit_error_kernel
  }
}

// the Init block takes over the linearization sfx_set_values computed while the values were still arriving
__global__ void adopt_eager_kernel(Ctrl* ctrl, const double* __restrict__ eager_err) {
  if (ctrl->done) return;
  if (ctrl->init_idx != kEagerBlock) {  // cannot happen (see kEagerBlock); never solve with another block's leftovers
    ctrl->done = 3;  // FAILED
    ctrl->failure_reason = 2;
    return;
  }
  ctrl->err[kEagerBlock] = *eager_err;
  ctrl->lin_valid[kEagerBlock] = 1;
}
void launch_adopt_eager(cudaStream_t st, Ctrl* ctrl, const double* eager_err) {
  adopt_eager_kernel<<<1, 1, 0, st>>>(ctrl, eager_err); ++g_launches;
}

// err[target] = (all-reduced) red[0]; marks the linearization valid
__global__ void commit_error_kernel(Ctrl* ctrl, int mode) {
  if (ctrl->done) return;
  const int blk = sel_block(ctrl, mode);
  if (mode == 0 && ctrl->lin_valid[blk]) return;
  ctrl->err[blk] = ctrl->red[0];
  ctrl->lin_valid[blk] = 1;
}

// multi-GPU: pack [B blocks | reduced rhs | red[0]] of the target block into a staging buffer for
// one ncclAllReduce, and unpack afterwards
__global__ void pack_b_kernel(Ctrl* ctrl, StatePtrs sp, int mode, int64_t nb, int nr, double* __restrict__ stage,
                              int unpack) {
  if (ctrl->done) return;
  const int blk = sel_block(ctrl, mode);
  if (mode == 0 && ctrl->lin_valid[blk]) return;
  double* H = sp.H[blk];
  double* rhs = sp.rhs[blk];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < nb; i += stride) {
    if (unpack)
      H[i] = stage[i];
    else
      stage[i] = H[i];
  }
  for (int64_t i = tid; i < nr; i += stride) {
    if (unpack)
      rhs[i] = stage[nb + i];
    else
      stage[nb + i] = rhs[i];
  }
  if (tid == 0) {
    if (unpack)
      ctrl->red[0] = stage[nb + nr];
    else
      stage[nb + nr] = ctrl->red[0];
  }
}
void launch_pack_b(cudaStream_t st, Ctrl* ctrl, StatePtrs sp, int mode, int64_t nb, int nr, double* stage, int unpack) {
  int grid = (int)((nb + nr + 255) / 256);
  if (grid > 148 * 4) grid = 148 * 4;
  if (grid < 1) grid = 1;
  pack_b_kernel<<<grid, 256, 0, st>>>(ctrl, sp, mode, nb, nr, stage, unpack); ++g_launches;
}
void launch_commit_error(cudaStream_t st, Ctrl* ctrl, int mode) {
  commit_error_kernel<<<1, 1, 0, st>>>(ctrl, mode); ++g_launches;
}

// stage[i] = mask[i] ? v[i] : 0  (gather of the sharded landmark values at the end of Optimize)
__global__ void mask_values_kernel(const double* __restrict__ v, const unsigned char* __restrict__ mask, int64_t n,
                                   double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = mask[i] ? v[i] : 0.0;
}
void launch_mask_values(cudaStream_t st, const double* v, const unsigned char* mask, int64_t n, double* out) {
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  mask_values_kernel<<<grid, 256, 0, st>>>(v, mask, n, out); ++g_launches;
}

void launch_zero_lin(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int mode, int64_t n_h, int n_rhs) {
  int64_t work = n_h > n_rhs ? n_h : n_rhs;
  int grid = (int)((work + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  zero_lin_kernel<<<grid, 256, 0, st>>>(ctrl, sp, mode, n_h, n_rhs); ++g_launches;
}

// multiprocessors of the current device (grid-stride / persistent launches are sized in multiples of it)
static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

int g_lin_skip = 0;  // debug (sfx_debug_time_linearize): parts of linearize_bal_kernel to leave out
// BAL fast path over the CTAs [block0, block1) of the batch (128 observations each), without the per-point sums
void launch_linearize_bal_range(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int mode, const LinBatch& b, double* partials,
                                int block0, int block1) {
  if (block1 <= block0) return;
  linearize_bal_kernel<0><<<block1 - block0, kBalThreads, 0, st>>>(ctrl, sp, b, mode, partials, block0); ++g_launches;
}
void launch_bal_point_finalize(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int mode, const LinBatch& b) {
  if (b.pbuf == nullptr) return;
  bal_point_finalize_kernel<<<(b.n_pf + 127) / 128, 128, 0, st>>>(ctrl, sp, b, mode); ++g_launches;
}
int linearize_bal_blocks(const LinBatch& b) { return (b.n + kBalThreads - 1) / kBalThreads; }

void launch_linearize(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int mode, const LinBatch& b, double* partials) {
  const int grid = (b.n + kLinThreads - 1) / kLinThreads;
  if (b.bal_fast) {
    switch (g_lin_skip) {
#define SFX_SKIP_CASE(V) \
  case V: linearize_bal_kernel<V><<<grid, kBalThreads, 0, st>>>(ctrl, sp, b, mode, partials, 0); break;
      SFX_SKIP_CASE(1)
      SFX_SKIP_CASE(2)
      SFX_SKIP_CASE(4)
      SFX_SKIP_CASE(7)
      SFX_SKIP_CASE(8)
      SFX_SKIP_CASE(15)
#undef SFX_SKIP_CASE
      default: linearize_bal_kernel<0><<<grid, kBalThreads, 0, st>>>(ctrl, sp, b, mode, partials, 0); break;
    }
    ++g_launches;
    if (b.pbuf != nullptr && !(g_lin_skip & 2)) {
      bal_point_finalize_kernel<<<(b.n_pf + 127) / 128, 128, 0, st>>>(ctrl, sp, b, mode); ++g_launches;
    }
    return;
  }
  switch (b.kind) {
#define SFX_CASE(ID) \
  case ID: linearize_kernel<ID><<<grid, kLinThreads, 0, st>>>(ctrl, sp, b, mode, partials); ++g_launches; break;
    SFX_CASE(SFX_KIND_SNAVELY)
    SFX_CASE(SFX_KIND_BETWEEN_POSE3)
    SFX_CASE(SFX_KIND_PRIOR_POSE3)
    SFX_CASE(SFX_KIND_MATCHING)
    SFX_CASE(SFX_KIND_ODOMETRY)
    SFX_CASE(SFX_KIND_IRL_LINEAR_GNC)
    SFX_CASE(SFX_KIND_IRL_PRIOR)
    SFX_CASE(SFX_KIND_BETWEEN_ROT3)
    SFX_CASE(SFX_KIND_PRIOR_ROT3)
    SFX_CASE(SFX_KIND_BARRON)
#undef SFX_CASE
  }
}

// Linearization::jacobian export (include_jacobians): one thread per factor slot evaluates the factor and stores its
// R x dim(key) Jacobian blocks into the reference's CSC value array; positions come from build_jacobian_csc (analysis.cc).
// Debug / introspection path: never part of an LM iteration.
template <int KIND>
__global__ void __launch_bounds__(kLinThreads) jacobian_kernel(const double* __restrict__ values, LinBatch b,
                                                                const int32_t* __restrict__ jac_base,
                                                                const int32_t* __restrict__ jac_colnnz,
                                                                double* __restrict__ out) {
  using K = Kind<KIND>;
  const int s = blockIdx.x * kLinThreads + threadIdx.x;
  if (s >= b.n) return;
  const double* a[K::NUSED];
#pragma unroll
  for (int u = 0; u < K::NUSED; ++u) a[u] = values + __ldg(b.arg_off + (size_t)u * b.n + s);
  double res[K::R];
  double J[K::R * K::T];
  K::eval(a, res, J);
#pragma unroll
  for (int ka = 0; ka < K::NOPT; ++ka) {
    const int base = __ldg(jac_base + (size_t)ka * b.n + s);
    if (base < 0) continue;
    const int cn = __ldg(jac_colnnz + (size_t)ka * b.n + s);
#pragma unroll
    for (int c = 0; c < K::dim(ka); ++c)
#pragma unroll
      for (int q = 0; q < K::R; ++q) out[(size_t)base + (size_t)c * cn + q] = J[q + (K::col(ka) + c) * K::R];
  }
}

void launch_jacobian(cudaStream_t st, const double* values, const LinBatch& b, const int32_t* jac_base,
                     const int32_t* jac_colnnz, double* out) {
  if (b.n == 0) return;
  const int grid = (b.n + kLinThreads - 1) / kLinThreads;
  switch (b.kind) {
#define SFX_CASE(ID) \
  case ID: jacobian_kernel<ID><<<grid, kLinThreads, 0, st>>>(values, b, jac_base, jac_colnnz, out); ++g_launches; break;
    SFX_CASE(SFX_KIND_SNAVELY)
    SFX_CASE(SFX_KIND_BETWEEN_POSE3)
    SFX_CASE(SFX_KIND_PRIOR_POSE3)
    SFX_CASE(SFX_KIND_MATCHING)
    SFX_CASE(SFX_KIND_ODOMETRY)
    SFX_CASE(SFX_KIND_IRL_LINEAR_GNC)
    SFX_CASE(SFX_KIND_IRL_PRIOR)
    SFX_CASE(SFX_KIND_BETWEEN_ROT3)
    SFX_CASE(SFX_KIND_PRIOR_ROT3)
    SFX_CASE(SFX_KIND_BARRON)
#undef SFX_CASE
  }
}

void launch_finish_error(cudaStream_t st, Ctrl* ctrl, int mode, const double* partials, int n_partials, double* eager_err) {
  finish_error_kernel<<<1, 1024, 0, st>>>(ctrl, partials, n_partials, mode, eager_err); ++g_launches;
}

// ------------------------------------------------------------------------------------------------
// K2: damping vector (internal scalar order)
// ------------------------------------------------------------------------------------------------
__global__ void damping_kernel(Ctrl* ctrl, StatePtrs sp, const int32_t* __restrict__ diag_pos, int N,
                               double* __restrict__ dvec, double* __restrict__ max_diag) {
  if (ctrl->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double* H = sp.H[ctrl->init_idx];
  const sfx_params& p = ctrl->p;
  const double lam = ctrl->lambda;
  double d = 0.0;
  if (p.use_diagonal_damping) {
    const double diag = H[diag_pos[i]];
    if (p.keep_max_diagonal_damping) {
      double m = ctrl->have_max_diag ? fmax(max_diag[i], diag) : fmax(diag, p.diagonal_damping_min);
      max_diag[i] = m;
      d = m * lam;
    } else {
      d = fmax(diag, p.diagonal_damping_min) * lam;
    }
  }
  if (p.use_unit_damping) d += lam;
  dvec[i] = d;
  if (p.debug_checks) {  // CheckHessianDiagonal: zero on the diagonal after damping
    if (fabs(H[diag_pos[i]] + d) < ctrl->epsilon) {
      const int slot = atomicAdd(&ctrl->n_zero_diag, 1);
      if (slot < 15) ctrl->zero_diag_idx[slot] = i;
    }
  }
}

__global__ void damping_flag_kernel(Ctrl* ctrl) {
  if (ctrl->done) return;
  if (ctrl->p.use_diagonal_damping && ctrl->p.keep_max_diagonal_damping) ctrl->have_max_diag = 1;
}

void launch_damping(cudaStream_t st, Ctrl* ctrl, StatePtrs sp, const int32_t* diag_pos, int N, double* dvec,
                    double* max_diag) {
  damping_kernel<<<(N + 255) / 256, 256, 0, st>>>(ctrl, sp, diag_pos, N, dvec, max_diag); ++g_launches;
  damping_flag_kernel<<<1, 1, 0, st>>>(ctrl); ++g_launches;
}

// ------------------------------------------------------------------------------------------------
// K3: Schur complement
// ------------------------------------------------------------------------------------------------
// C^-1 per landmark via dense LLT solve of the identity (sparse_schur_solver.tcc:105-119), in registers.
__global__ void schur_cinv_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp, SchurDev sd,
                                  const double* __restrict__ dvec) {
  if (ctrl->done) return;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= sd.n_landmarks) return;
  const int blk = ctrl->init_idx;
  const double* H = sp.H[blk];
  const double* rhs = sp.rhs[blk];
  const int d = sd.lm_dim[l];
  const int to = sd.lm_toff[l];
  const double* C = H + sd.lm_cdiag_off[l];
  double A[3][3], L[3][3], Ci[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      A[r][c] = 0.0;
      L[r][c] = 0.0;
      Ci[r][c] = 0.0;
    }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = c; r < 3; ++r)
      if (r < d && c < d) A[r][c] = C[r + c * d] + (r == c ? dvec[to + r] : 0.0);
  // Cholesky
#pragma unroll
  for (int j = 0; j < 3; ++j)
    if (j < d) {
      double x = A[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) x -= L[j][k] * L[j][k];
      x = sqrt(x);
      L[j][j] = x;
#pragma unroll
      for (int i = j + 1; i < 3; ++i)
        if (i < d) {
          double y = A[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) y -= L[i][k] * L[j][k];
          L[i][j] = y / x;
        }
    }
#pragma unroll
  for (int e = 0; e < 3; ++e)
    if (e < d) {
      double y[3] = {0, 0, 0};
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (i < d) {
          double v = (i == e) ? 1.0 : 0.0;
#pragma unroll
          for (int k = 0; k < i; ++k) v -= L[i][k] * y[k];
          y[i] = v / L[i][i];
        }
#pragma unroll
      for (int i = 2; i >= 0; --i)
        if (i < d) {
          double v = y[i];
#pragma unroll
          for (int k = i + 1; k < 3; ++k)
            if (k < d) v -= L[k][i] * y[k];
          y[i] = v / L[i][i];
        }
#pragma unroll
      for (int i = 0; i < 3; ++i) Ci[i][e] = y[i];
    }
  // symmetrise from the lower part (the reference keeps C_inv_lower and reads it as selfadjoint)
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < c; ++r) Ci[r][c] = Ci[c][r];
  double* out = sd.cinv + (size_t)l * 9;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) out[r + c * 3] = Ci[r][c];
  double w[3] = {0, 0, 0};
#pragma unroll
  for (int r = 0; r < 3; ++r)
    if (r < d) w[r] = rhs[to + r];
#pragma unroll
  for (int r = 0; r < 3; ++r) sd.tl[(size_t)l * 3 + r] = Ci[r][0] * w[0] + Ci[r][1] * w[1] + Ci[r][2] * w[2];
  if (sd.wl != nullptr) {
    // fast path: whitening factor L^-1 (C + D = L L^T) and u = L^-1 w, so that
    // E^T C^-1 E = (L^-1 E)^T (L^-1 E) and E^T C^-1 w = (L^-1 E)^T u
    const double i00 = 1.0 / L[0][0], i11 = 1.0 / L[1][1], i22 = 1.0 / L[2][2];
    const double i10 = -L[1][0] * i00 * i11;
    const double i21 = -L[2][1] * i11 * i22;
    const double i20 = -(L[2][0] * i00 + L[2][1] * i10) * i22;
    double* o = sd.wl + (size_t)l * 9;
    o[0] = i00;
    o[1] = i10;
    o[2] = i11;
    o[3] = i20;
    o[4] = i21;
    o[5] = i22;
    o[6] = i00 * w[0];
    o[7] = i10 * w[0] + i11 * w[1];
    o[8] = i20 * w[0] + i21 * w[1] + i22 * w[2];
  }
}

// One warp per work item = (S block, chunk of <= 32 matches): partial = sum_matches E_I^T C^-1 E_J.
// Items are sorted by block column, so concurrently running warps touch a narrow band of cameras
// and the E blocks are served from L2.  Blocks with one chunk are stored directly
// (S_IJ = B_IJ + D - partial); longer match lists are split and combined with fp64 RED atomics on
// the zeroed S.
constexpr int kSchurWarps = 4;
__global__ void __launch_bounds__(kSchurWarps * 32) schur_s_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                                   SchurDev sd, const double* __restrict__ dvec) {
  // sd.add_b == 0 (ranks > 0 of a sharded run): B and the damping are contributed by rank 0 only
  if (ctrl->done) return;
  __shared__ double sh[kSchurWarps][3 * 16 * 2 + 3 * 16 + 9];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kSchurWarps + wid;
  if (item >= sd.n_items) return;
  const int b = sd.item_blk[item];
  const int m0 = sd.item_m0[item], m1 = m0 + sd.item_cnt[item];
  const int first = (sd.item_flags[item] & 1) && sd.add_b, single = (sd.item_flags[item] >> 1) & 1;
  const double* H = sp.H[ctrl->init_idx];
  const int I = sd.s_row[b], J = sd.s_col[b];
  const int dI = sd.node_dim[I], dJ = sd.node_dim[J];
  const int ne = dI * dJ;
  double* EI = sh[wid];
  double* EJ = EI + 48;
  double* G = EJ + 48;
  double* Cs = G + 48;
  double acc[8];
  int er[8], ec[8];  // (row, col) of the S-block elements this lane owns (hoisted out of the match loop)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    acc[k] = 0.0;
    const int e = lane + 32 * k;
    er[k] = e % dI;
    ec[k] = e / dI;
  }
  const int nslots = (ne + 31) >> 5;
  // landmark dims are uniform inside a problem in practice; the generic loop bounds stay runtime
  for (int m = m0; m < m1; ++m) {
    const int l = sd.m_lm[m];
    const int dl = sd.lm_dim[l];
    const double* __restrict__ ei = H + sd.m_eoff_i[m];
    const double* __restrict__ ej = H + sd.m_eoff_j[m];
    // issue all global loads of this match first
    double vi0 = 0, vi1 = 0, vj0 = 0, vj1 = 0, vc = 0;
    const int ni = dl * dI, nj = dl * dJ;
    if (lane < ni) vi0 = ei[lane];
    if (lane + 32 < ni) vi1 = ei[lane + 32];
    if (lane < nj) vj0 = ej[lane];
    if (lane + 32 < nj) vj1 = ej[lane + 32];
    if (lane < 9) vc = sd.cinv[(size_t)l * 9 + lane];
    __syncwarp();
    if (lane < ni) EI[lane] = vi0;
    if (lane + 32 < ni) EI[lane + 32] = vi1;
    if (lane < nj) EJ[lane] = vj0;
    if (lane + 32 < nj) EJ[lane + 32] = vj1;
    if (lane < 9) Cs[lane] = vc;
    __syncwarp();
    // G = Cinv * EJ  (dl x dJ), element t = a + dl*c
    if (dl == 3) {
      for (int t = lane; t < nj; t += 32) {
        const int c = t / 3, a = t - 3 * c;
        G[t] = Cs[a] * EJ[3 * c] + Cs[a + 3] * EJ[3 * c + 1] + Cs[a + 6] * EJ[3 * c + 2];
      }
    } else {
      for (int t = lane; t < nj; t += 32) {
        const int a = t % dl, c = t / dl;
        double v = 0;
        for (int q = 0; q < dl; ++q) v += Cs[a + q * 3] * EJ[q + c * dl];
        G[t] = v;
      }
    }
    __syncwarp();
    if (dl == 3) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < nslots && lane + 32 * k < ne) {
          const double* a = EI + 3 * er[k];
          const double* g = G + 3 * ec[k];
          acc[k] += a[0] * g[0] + a[1] * g[1] + a[2] * g[2];
        }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < nslots && lane + 32 * k < ne) {
          double v = 0;
          for (int q = 0; q < dl; ++q) v += EI[q + er[k] * dl] * G[q + ec[k] * dl];
          acc[k] += v;
        }
    }
  }
  double* out = sd.S + sd.s_off[b];
  const int bsrc = sd.s_b_src[b];
  const int toI = sd.node_toff[I];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = lane + 32 * k;
    if (e < ne) {
      const int r = er[k], c = ec[k];
      double v = -acc[k];
      if (first) {
        if (bsrc >= 0) v += H[bsrc + e];
        if (I == J && r == c) v += dvec[toI + r];
      }
      if (single)
        out[e] = v;
      else
        atomicAdd(out + e, v);
    }
  }
}

__global__ void zero_kernel(const Ctrl* __restrict__ ctrl, double* __restrict__ p, int64_t n) {
  if (ctrl->done) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0.0;
}

// reduced rhs: v_I - sum_l E_{l,I}^T (C_l^-1 w_l); one warp per reduced node
__global__ void __launch_bounds__(128) schur_rhs_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp, SchurDev sd) {
  if (ctrl->done) return;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = blockIdx.x * 4 + wid;
  if (I >= sd.n_reduced_nodes) return;
  const int blk = ctrl->init_idx;
  const double* H = sp.H[blk];
  const double* rhs = sp.rhs[blk];
  const int dI = sd.node_dim[I];
  // lanes: r = lane % 16 is the column of E (component of node I), half = lane / 16 splits entries
  const int r = lane & 15, half = lane >> 4;
  double acc = 0.0;
  for (int q = sd.r_ptr[I] + half; q < sd.r_ptr[I + 1]; q += 2) {
    const int l = sd.r_lm[q];
    const int dl = sd.lm_dim[l];
    if (r < dI) {
      const double* e = H + sd.r_eoff[q] + r * dl;
      const double* t = sd.tl + (size_t)l * 3;
      double v = 0;
      for (int a = 0; a < dl; ++a) v += e[a] * t[a];
      acc += v;
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  if (half == 0 && r < dI) {
    const int to = sd.node_toff[I];
    sd.rhs_red[to + r] = (sd.add_b ? rhs[to + r] : 0.0) - acc;
  }
}

// z_l = C^-1 (w_l - E_l y) = t_l - C^-1 (E_l y); update = -[y; z]
__global__ void schur_back_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp, SchurDev sd, const double* __restrict__ y,
                                  double* __restrict__ upd) {
  if (ctrl->done) return;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  for (int i = l; i < sd.reduced_dim; i += stride) upd[i] = -y[i];
  if (l >= sd.n_landmarks) return;
  const double* H = sp.H[ctrl->init_idx];
  const int dl = sd.lm_dim[l];
  double s[3] = {0, 0, 0};
  for (int q = sd.lm_e_ptr[l]; q < sd.lm_e_ptr[l + 1]; ++q) {
    const int J = sd.lm_e_node[q];
    const int dJ = sd.node_dim[J];
    const double* e = H + sd.lm_e_off[q];
    const double* yj = y + sd.node_toff[J];
    for (int c = 0; c < dJ; ++c) {
      const double yc = yj[c];
      for (int a = 0; a < dl; ++a) s[a] += e[a + c * dl] * yc;
    }
  }
  const double* Ci = sd.cinv + (size_t)l * 9;
  const double* t = sd.tl + (size_t)l * 3;
  const int to = sd.lm_toff[l];
  for (int a = 0; a < dl; ++a) {
    double z = t[a] - (Ci[a] * s[0] + Ci[a + 3] * s[1] + Ci[a + 6] * s[2]);
    upd[to + a] = -z;
  }
}

// ---- K3 fast path: landmarks of dim 3, reduced nodes of dim <= 16 -----------------------------------
// (1) schur_g_rhs_kernel: one pass over the E blocks in camera-major order: G = C^-1 E is written
//     next to E (same offsets, second buffer) and the reduced rhs  v_I - sum E^T (C^-1 w)  is
//     accumulated (warp-level sum when the 32 blocks of a warp share the camera);
// (2) schur_s_dmma_kernel: S_IJ -= sum_matches E_I^T G_J as FP64 tensor-core products
//     (mma.sync.m8n8k4.f64, k = landmark dim padded to 4): per match 4 scalar loads + <= 4 DMMA per lane;
// (3) schur_back_fast_kernel: per E block  s_l += E y_I  (RED atomics, 3 per block), then
//     z_l = t_l - C^-1 s_l per landmark.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void schur_rhs_init_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp, SchurDev sd) {
  if (ctrl->done) return;
  const double* rhs = sp.rhs[ctrl->init_idx];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < sd.reduced_dim) sd.rhs_red[i] = sd.add_b ? rhs[i] : 0.0;
  for (int q = i; q < sd.n_landmarks * 3; q += gridDim.x * blockDim.x) sd.sl[q] = 0.0;
}

constexpr int kGThreads = 128;
__global__ void __launch_bounds__(kGThreads) schur_g_rhs_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                                SchurDev sd) {
  if (ctrl->done) return;
  const int q = blockIdx.x * kGThreads + threadIdx.x;
  const bool valid = q < sd.n_entries;
  const int qc = valid ? q : sd.n_entries - 1;
  const double* __restrict__ H = sp.H[ctrl->init_idx];
  const int I = sd.r_node[qc];
  const int l = sd.r_lm[qc];
  const int eoff = sd.r_eoff[qc];
  const int dI = sd.node_dim[I];
  const double* ci = sd.cinv + (size_t)l * 9;
  const double c00 = ci[0], c10 = ci[1], c20 = ci[2], c11 = ci[4], c21 = ci[5], c22 = ci[8];
  const double t0 = sd.tl[(size_t)l * 3], t1 = sd.tl[(size_t)l * 3 + 1], t2 = sd.tl[(size_t)l * 3 + 2];
  const double* e = H + eoff;
  double* g = sd.G + eoff;
  double r[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    r[c] = 0.0;
    if (c < dI && valid) {
      const double e0 = e[3 * c], e1 = e[3 * c + 1], e2 = e[3 * c + 2];
      g[3 * c] = c00 * e0 + c10 * e1 + c20 * e2;
      g[3 * c + 1] = c10 * e0 + c11 * e1 + c21 * e2;
      g[3 * c + 2] = c20 * e0 + c21 * e1 + c22 * e2;
      r[c] = e0 * t0 + e1 * t1 + e2 * t2;
    }
  }
  // reduced rhs: warp-level sum when the whole warp works on one camera
  const int I0 = __shfl_sync(0xffffffffu, I, 0);
  const bool uniform = __all_sync(0xffffffffu, I == I0);
  const int lane = threadIdx.x & 31;
  const int to = sd.node_toff[I];
  if (uniform) {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < dI) {
        double v = r[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == c) atomicAdd(sd.rhs_red + to + c, -v);
      }
  } else if (valid) {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < dI) atomicAdd(sd.rhs_red + to + c, -r[c]);
  }
}

__global__ void __launch_bounds__(kSchurWarps * 32) schur_s_dmma_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                                        SchurDev sd, const double* __restrict__ dvec) {
  if (ctrl->done) return;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kSchurWarps + wid;
  if (item >= sd.n_items) return;
  const int b = sd.item_blk[item];
  const int m0 = sd.item_m0[item], m1 = m0 + sd.item_cnt[item];
  const int first = (sd.item_flags[item] & 1) && sd.add_b, single = (sd.item_flags[item] >> 1) & 1;
  const double* __restrict__ H = sp.H[ctrl->init_idx];
  const double* __restrict__ G = sd.G;
  const int I = sd.s_row[b], J = sd.s_col[b];
  const int dI = sd.node_dim[I], dJ = sd.node_dim[J];
  const int g = lane >> 2, tq = lane & 3;
  // fragment element offsets inside a 3 x d block (column-major, ld 3); -1: padding
  const int a0 = (tq < 3 && g < dI) ? tq + 3 * g : -1;
  const int a1 = (tq < 3 && g + 8 < dI) ? tq + 3 * (g + 8) : -1;
  const int b0 = (tq < 3 && g < dJ) ? tq + 3 * g : -1;
  const int b1 = (tq < 3 && g + 8 < dJ) ? tq + 3 * (g + 8) : -1;
  const bool two_r = dI > 8, two_c = dJ > 8;
  double acc[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
  // software pipeline: loads of match m+1 are in flight while the MMAs of match m issue
  double fa0 = 0, fa1 = 0, fb0 = 0, fb1 = 0;
  if (m0 < m1) {
    const double* ei = H + sd.m_eoff_i[m0];
    const double* gj = G + sd.m_eoff_j[m0];
    fa0 = a0 >= 0 ? ei[a0] : 0.0;
    fa1 = a1 >= 0 ? ei[a1] : 0.0;
    fb0 = b0 >= 0 ? gj[b0] : 0.0;
    fb1 = b1 >= 0 ? gj[b1] : 0.0;
  }
  for (int m = m0; m < m1; ++m) {
    const double ca0 = fa0, ca1 = fa1, cb0 = fb0, cb1 = fb1;
    if (m + 1 < m1) {
      const double* ei = H + sd.m_eoff_i[m + 1];
      const double* gj = G + sd.m_eoff_j[m + 1];
      fa0 = a0 >= 0 ? ei[a0] : 0.0;
      fa1 = a1 >= 0 ? ei[a1] : 0.0;
      fb0 = b0 >= 0 ? gj[b0] : 0.0;
      fb1 = b1 >= 0 ? gj[b1] : 0.0;
    }
    dmma884(acc[0][0][0], acc[0][0][1], ca0, cb0);
    if (two_c) dmma884(acc[0][1][0], acc[0][1][1], ca0, cb1);
    if (two_r) {
      dmma884(acc[1][0][0], acc[1][0][1], ca1, cb0);
      if (two_c) dmma884(acc[1][1][0], acc[1][1][1], ca1, cb1);
    }
  }
  double* out = sd.S + sd.s_off[b];
  const int bsrc = sd.s_b_src[b];
  const int toI = sd.node_toff[I];
#pragma unroll
  for (int rb = 0; rb < 2; ++rb)
#pragma unroll
    for (int cb = 0; cb < 2; ++cb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = rb * 8 + g, c = cb * 8 + tq * 2 + e;
        if (r < dI && c < dJ) {
          const int idx = r + c * dI;
          double v = -acc[rb][cb][e];
          if (first) {
            if (bsrc >= 0) v += H[bsrc + idx];
            if (I == J && r == c) v += dvec[toI + r];
          }
          if (single)
            out[idx] = v;
          else
            atomicAdd(out + idx, v);
        }
      }
}

// ---- TMA (cp.async.bulk, 1-D) + mbarrier helpers of the streamed kernels -----------------------------------------
constexpr int kTmaSlots = 128;   // slots per tile (27,648 B)
#ifndef SFX_TMA_STAGES
#define SFX_TMA_STAGES 2
#endif
#ifndef SFX_TMA_CTAS
#define SFX_TMA_CTAS 3
#endif
// measured at Final-shape (1.08 GB of E blocks): 4 stages x 2 CTAs per SM 0.243 ms, 3 x 2 0.243, 2 x 4 0.216, 2 x 3 0.210
// (5.7 TB/s); the per-entry kernel it replaces: 0.291 ms
constexpr int kTmaStages = SFX_TMA_STAGES;
constexpr int kTmaCtas = SFX_TMA_CTAS;
constexpr int kTmaTileBytes = kTmaSlots * 27 * 8;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- K3 fast path, v2 ------------------------------------------------------------------------------
// (1) schur_w_rhs_kernel: W = L^-1 E (whitened point-camera blocks, written to the second buffer at the
//     offsets of E) and the reduced rhs v_I - sum W^T u.  A warp's 32 blocks are one contiguous run in
//     camera-major order: they are read and written coalesced through shared memory.
// (2) schur_s2_kernel: S_IJ -= sum_matches W_I^T W_J with the matches stacked along k: one
//     mma.sync.m8n8k4.f64 step covers 4/3 matches (no k padding); a 9 x 9 block takes 3 DMMAs per step
//     (the 9th row x 9th column corner is one FMA per lane + a 4-lane reduction at the end).  Match
//     offsets are loaded once per 32 matches (one coalesced load) and broadcast with shuffles; the
//     loads of the next four k-steps (16 per lane) are in flight while the current four issue.
constexpr int kWThreads = 128;
#ifndef SFX_W_CHUNKS
#define SFX_W_CHUNKS 4
#endif
constexpr int kWChunks = SFX_W_CHUNKS;  // schur_w_rhs_kernel<true>: chunks of 32 blocks per warp
#ifndef SFX_W_MINB
#define SFX_W_MINB 5
#endif
// DIAG: the warp also accumulates S_II -= sum over its 32 blocks W^T W from the staged blocks: 24 DMMA k-steps over the
// 96 stacked point rows (A and B operand are the same value: W^T W), border row / corner as FMAs.
template <bool DIAG>
__global__ void __launch_bounds__(kWThreads, SFX_W_MINB) schur_w_rhs_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                                SchurDev sd) {
  __shared__ double stage[kWThreads / 32][32 * 27];
  if (ctrl->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* __restrict__ H = sp.H[ctrl->init_idx];
  // DIAG: a warp takes kWChunks consecutive chunks of 32 blocks and keeps the S_II accumulators in registers across
  // them (flushed with REDs when the camera changes): a camera's ~2,800 blocks otherwise send ~90 warps' worth of REDs
  // to the same 81 addresses
  constexpr int kIter = DIAG ? kWChunks : 1;
  const int g_ = lane >> 2, tq_ = lane & 3;
  double acc0 = 0.0, acc1 = 0.0, row8 = 0.0, corner = 0.0;
  int accI = -1;
  auto flush = [&]() {
    if (accI < 0) return;
    row8 += __shfl_xor_sync(0xffffffffu, row8, 1);
    row8 += __shfl_xor_sync(0xffffffffu, row8, 2);
    corner += __shfl_xor_sync(0xffffffffu, corner, 1);
    corner += __shfl_xor_sync(0xffffffffu, corner, 2);
    // lower triangle only: the front assembly reads S_II as a lower block (FrontPlan::Copy::lower_only)
    double* out = sd.S + sd.s_diag_off[accI];
    if (g_ >= 2 * tq_) atomicAdd(out + g_ + (2 * tq_) * 9, -acc0);
    if (g_ >= 2 * tq_ + 1) atomicAdd(out + g_ + (2 * tq_ + 1) * 9, -acc1);
    if (tq_ == 0) atomicAdd(out + 8 + g_ * 9, -row8);
    if (lane == 2) atomicAdd(out + 8 + 8 * 9, -corner);
    acc0 = acc1 = row8 = corner = 0.0;
    accI = -1;
  };
#pragma unroll 1
  for (int it = 0; it < kIter; ++it) {
  const int q = (blockIdx.x * (kWThreads / 32) + warp) * (32 * kIter) + it * 32 + lane;
  if (q - lane >= sd.n_entries) break;
  const bool valid = q < sd.n_entries;
  const int qc = valid ? q : sd.n_entries - 1;
  const int I = __ldg(sd.r_node + qc);
  const int l = __ldg(sd.r_lm + qc);
  const int eoff = __ldg(sd.r_eoff + qc);
  const int dI = __ldg(sd.node_dim + I);
  const double* wl = sd.wl + (size_t)l * 9;
  const double i00 = wl[0], i10 = wl[1], i11 = wl[2], i20 = wl[3], i21 = wl[4], i22 = wl[5];
  const double u0 = wl[6], u1 = wl[7], u2 = wl[8];
  const int I0 = __shfl_sync(0xffffffffu, I, 0);
  const bool uniform = __all_sync(0xffffffffu, I == I0);
  const int eoff0 = __shfl_sync(0xffffffffu, eoff, 0);
  const int n = 3 * dI;
  const bool staged = uniform && dI == 9 && __all_sync(0xffffffffu, valid && eoff == eoff0 + n * lane);
  double r[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) r[c] = 0.0;
  if (staged) {
    double* st = stage[warp];
    const double* src = H + eoff0;
    {
      double t[27];
#pragma unroll
      for (int i = 0; i < 27; ++i) t[i] = SFX_LD_STREAM(src + i * 32 + lane);
#pragma unroll
      for (int i = 0; i < 27; ++i) st[i * 32 + lane] = t[i];
    }
    __syncwarp();
    double* e = st + lane * 27;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      {
        const double e0 = e[3 * c], e1 = e[3 * c + 1], e2 = e[3 * c + 2];
        const double w0 = i00 * e0, w1 = i10 * e0 + i11 * e1, w2 = i20 * e0 + i21 * e1 + i22 * e2;
        e[3 * c] = w0;
        e[3 * c + 1] = w1;
        e[3 * c + 2] = w2;
        r[c] = w0 * u0 + w1 * u1 + w2 * u2;
      }
    }
    __syncwarp();
    double* dst = sd.G + eoff0;
#pragma unroll
    for (int i = 0; i < 27; ++i) dst[i * 32 + lane] = st[i * 32 + lane];
    if (DIAG) {
      // element (point row k = 3 e + a, camera column c) of the stacked 96 x 9 matrix sits at st[27 e + 3 c + a]
      if (accI != I0) {
        flush();
        accI = I0;
      }
#pragma unroll 8
      for (int s4 = 0; s4 < 24; ++s4) {
        const int kk = 4 * s4 + tq_;
        const int en = (kk * 171) >> 9;  // kk / 3
        const double* wp = st + 27 * en + (kk - 3 * en);
        const double a = wp[3 * g_];
        const double a8 = wp[24];
        dmma884(acc0, acc1, a, a);
        row8 += a8 * a;
        corner += a8 * a8;
      }
      __syncwarp();  // the staging area is rewritten by the next chunk
    }
  } else {
    const double* e = H + eoff;
    double* g = sd.G + eoff;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      r[c] = 0.0;
      if (c < dI && valid) {
        const double e0 = e[3 * c], e1 = e[3 * c + 1], e2 = e[3 * c + 2];
        const double w0 = i00 * e0, w1 = i10 * e0 + i11 * e1, w2 = i20 * e0 + i21 * e1 + i22 * e2;
        g[3 * c] = w0;
        g[3 * c + 1] = w1;
        g[3 * c + 2] = w2;
        r[c] = w0 * u0 + w1 * u1 + w2 * u2;
      }
    }
    if (DIAG && valid) {
      // ragged warp (camera boundary, tail): the lane's own block, read back from what it has just written
      double* out = sd.S + sd.s_diag_off[I];
#pragma unroll 1
      for (int c = 0; c < 9; ++c)
#pragma unroll 1
        for (int rr = c; rr < 9; ++rr)
          atomicAdd(out + rr + c * 9, -(g[3 * rr] * g[3 * c] + g[3 * rr + 1] * g[3 * c + 1] + g[3 * rr + 2] * g[3 * c + 2]));
    }
  }
  const int to = sd.node_toff[I];
  if (uniform) {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < dI && (c < 9 || !staged)) {
        double v = r[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == c) atomicAdd(sd.rhs_red + to + c, -v);
      }
  } else if (valid) {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < dI) atomicAdd(sd.rhs_red + to + c, -r[c]);
  }
  }  // chunks
  if (DIAG) flush();
}

// S_II <- B_II + damping (rank 0 / single GPU; zero elsewhere) for the diagonal blocks schur_w_rhs_kernel<true> accumulates
__global__ void schur_diag_init_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp, SchurDev sd,
                                       const double* __restrict__ dvec) {
  if (ctrl->done) return;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= sd.n_reduced_nodes * 81) return;
  const int I = q / 81, idx = q - 81 * I;
  const int r = idx % 9, c = idx / 9;
  double v = 0.0;
  if (sd.add_b) {
    const int bsrc = sd.s_diag_bsrc[I];
    if (bsrc >= 0) v = sp.H[ctrl->init_idx][bsrc + idx];
    if (r == c) v += dvec[sd.node_toff[I] + r];
  }
  sd.S[sd.s_diag_off[I] + idx] = v;
}

struct alignas(16) SItem2 {
  int32_t m0, cnt;
  int32_t flags;  // bit0 first chunk of the block, bit1 only chunk, bits 8..15 dI, bits 16..23 dJ, bit 24: I == J
  int32_t toI;
  int64_t s_off;
  int32_t bsrc, pad;
};
static_assert(sizeof(SItem2) == 32, "SItem2 layout");

__global__ void __launch_bounds__(kSchurWarps * 32, 4) schur_s2_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                                      SchurDev sd, const double* __restrict__ dvec) {
  if (ctrl->done) return;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kSchurWarps + wid;
  if (item >= sd.n_items) return;
  SItem2 it;
  {
    const int4* ip = reinterpret_cast<const int4*>(sd.items2) + (size_t)item * 2;
    const int4 a = __ldg(ip), b = __ldg(ip + 1);
    it.m0 = a.x;
    it.cnt = a.y;
    it.flags = a.z;
    it.toI = a.w;
    it.s_off = ((int64_t)(uint32_t)b.x) | ((int64_t)b.y << 32);
    it.bsrc = b.z;
  }
  const double* __restrict__ W = sd.G;
  const double* __restrict__ zp = sd.zeros;
  const int dI = (it.flags >> 8) & 0xff, dJ = (it.flags >> 16) & 0xff;
  const int g = lane >> 2, tq = lane & 3;
  const bool two_r = dI > 8, two_c = dJ > 8;
  const bool corner_fma = dI == 9 && dJ == 9;
  // element offsets of this lane's fragment rows inside a 3 x d block; rows/cols >= d are padding
  const bool va0 = g < dI, va1 = g + 8 < dI, vb0 = g < dJ, vb1 = g + 8 < dJ;
  const int ra0 = 3 * g, ra1 = 3 * (g + 8);
  double acc[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
  double corner = 0.0;
  for (int base = 0; base < it.cnt; base += 32) {
    const int nm = min(32, it.cnt - base);
    int oi = 0, oj = 0;
    if (lane < nm) {
      oi = __ldg(sd.m_eoff_i + it.m0 + base + lane);
      oj = __ldg(sd.m_eoff_j + it.m0 + base + lane);
    }
    const int K = 3 * nm;
    const int nsteps = (K + 3) >> 2;
    double fa0[2][4], fa1[2][4], fb0[2][4], fb1[2][4];
    auto load_group = [&](int s0, int buf) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = 4 * (s0 + u) + tq;
        const bool ok = kk < K;
        const int mm = (kk * 43) >> 7;  // kk / 3 for kk < 128
        const int a = kk - 3 * mm;
        const int offI = __shfl_sync(0xffffffffu, oi, mm & 31);
        const int offJ = __shfl_sync(0xffffffffu, oj, mm & 31);
        const double* pa = W + offI + a;
        const double* pb = W + offJ + a;
        // padding lanes read a zero from `zp`: the select is on the address, so nothing consumes the
        // loaded value before the MMA and the loads of a group stay in flight together
        fa0[buf][u] = __ldg((ok && va0) ? pa + ra0 : zp);
        fb0[buf][u] = __ldg((ok && vb0) ? pb + ra0 : zp);
        fa1[buf][u] = __ldg((ok && va1) ? pa + ra1 : zp);
        fb1[buf][u] = __ldg((ok && vb1) ? pb + ra1 : zp);
      }
    };
    auto mma_group = [&](int buf) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dmma884(acc[0][0][0], acc[0][0][1], fa0[buf][u], fb0[buf][u]);
        if (two_c) dmma884(acc[0][1][0], acc[0][1][1], fa0[buf][u], fb1[buf][u]);
        if (two_r) {
          dmma884(acc[1][0][0], acc[1][0][1], fa1[buf][u], fb0[buf][u]);
          if (corner_fma)
            corner += fa1[buf][u] * fb1[buf][u];
          else if (two_c)
            dmma884(acc[1][1][0], acc[1][1][1], fa1[buf][u], fb1[buf][u]);
        }
      }
    };
    load_group(0, 0);
    for (int s0 = 0; s0 < nsteps; s0 += 8) {
      if (s0 + 4 < nsteps) load_group(s0 + 4, 1);
      mma_group(0);
      if (s0 + 4 < nsteps) {
        if (s0 + 8 < nsteps) load_group(s0 + 8, 0);
        mma_group(1);
      }
    }
  }
  if (corner_fma) {
    // lanes g == 0 hold the k-partials of entry (8, 8)
    corner += __shfl_xor_sync(0xffffffffu, corner, 1);
    corner += __shfl_xor_sync(0xffffffffu, corner, 2);
    if (lane == 0) acc[1][1][0] = corner;
  }
  const double* __restrict__ H = sp.H[ctrl->init_idx];
  double* out = sd.S + it.s_off;
  const int first = (it.flags & 1) && sd.add_b, single = (it.flags >> 1) & 1;
  const bool diag = (it.flags >> 24) & 1;
#pragma unroll
  for (int rb = 0; rb < 2; ++rb)
#pragma unroll
    for (int cb = 0; cb < 2; ++cb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = rb * 8 + g, c = cb * 8 + tq * 2 + e;
        if (r < dI && c < dJ) {
          const int idx = r + c * dI;
          double v = -acc[rb][cb][e];
          if (first) {
            if (it.bsrc >= 0) v += H[it.bsrc + idx];
            if (diag && r == c) v += dvec[it.toI + r];
          }
          if (single)
            out[idx] = v;
          else
            atomicAdd(out + idx, v);
        }
      }
}

// (2b) schur_s9_kernel: persistent kernel for the BAL shape (every reduced node has dim 9).  Items are
//      chunks of <= 64 matches whose offsets are stored padded (64 slots per item; unused slots point
//      at a block of zeros behind W), so an item is addressed by its id alone: the 16-byte header and
//      the offsets of the NEXT item are prefetched while the current one runs.  The 9 x 9 product is
//      split into the 8 x 8 core (ONE mma.sync.m8n8k4.f64 per k-step of 4/3 matches) and the border
//      row 8 / column 8 / corner, which are three FMAs per lane on operands the lane already holds,
//      reduced over the four k-lanes at the end.  Offsets sit in a per-warp shared-memory table; the 16
//      loads of the next four k-steps are in flight while the current four issue.
constexpr int kS9Warps = 4;
#ifndef SFX_S9_MINB
#define SFX_S9_MINB 4
#endif
__global__ void __launch_bounds__(kS9Warps * 32, SFX_S9_MINB) schur_s9_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                                   SchurDev sd, const double* __restrict__ dvec) {
  __shared__ int32_t offs[kS9Warps][2][128];  // [buffer][0..63: row operand, 64..127: column operand]
  if (ctrl->done) return;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = gridDim.x * kS9Warps;
  int item = blockIdx.x * kS9Warps + wid;
  if (item >= sd.n_items3) return;
  const double* __restrict__ W = sd.G;
  const double* __restrict__ H = sp.H[ctrl->init_idx];
  const int4* __restrict__ hdr = reinterpret_cast<const int4*>(sd.items3);
  const int g = lane >> 2, tq = lane & 3;
  const double* __restrict__ Wg = W + 3 * g;
  double fa0[2][4], fb0[2][4], fa8[2][4], fb8[2][4];
  auto load_group = [&](const int32_t* tab, int s0, int buf) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = 4 * (s0 + u) + tq;
      const int mm = (kk * 171) >> 9;  // kk / 3 for kk < 512
      const int a = kk - 3 * mm;
      const int oi = tab[mm] + a, oj = tab[64 + mm] + a;
      fa0[buf][u] = __ldg(Wg + oi);
      fb0[buf][u] = __ldg(Wg + oj);
      fa8[buf][u] = __ldg(W + oi + 24);
      fb8[buf][u] = __ldg(W + oj + 24);
    }
  };
  int4 h = __ldg(hdr + item);
  int cb = 0;
  {
    const int32_t* pi = sd.pm_i + (size_t)item * 64;
    const int32_t* pj = sd.pm_j + (size_t)item * 64;
    int32_t* tab = offs[wid][0];
    tab[lane] = __ldg(pi + lane);
    tab[32 + lane] = __ldg(pi + 32 + lane);
    tab[64 + lane] = __ldg(pj + lane);
    tab[96 + lane] = __ldg(pj + 32 + lane);
    __syncwarp();
  }
  load_group(offs[wid][0], 0, 0);
  for (;;) {
    const int nitem = item + nw;
    const bool more = nitem < sd.n_items3;
    int4 hn = make_int4(0, 0, 0, 0);
    int na = 0, nb = 0, nc = 0, nd = 0;
    if (more) {
      hn = __ldg(hdr + nitem);
      const int32_t* pi = sd.pm_i + (size_t)nitem * 64;
      const int32_t* pj = sd.pm_j + (size_t)nitem * 64;
      na = __ldg(pi + lane);
      nb = __ldg(pi + 32 + lane);
      nc = __ldg(pj + lane);
      nd = __ldg(pj + 32 + lane);
    }
    const int32_t* tab = offs[wid][cb];
    const int cnt = (h.w >> 25) & 0x7f;
    const int nsteps = (3 * cnt + 3) >> 2;
    double acc0 = 0.0, acc1 = 0.0, row8 = 0.0, col8 = 0.0, corner = 0.0;
    auto mma_group = [&](int buf) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dmma884(acc0, acc1, fa0[buf][u], fb0[buf][u]);
        row8 += fa8[buf][u] * fb0[buf][u];
        col8 += fa0[buf][u] * fb8[buf][u];
        corner += fa8[buf][u] * fb8[buf][u];
      }
    };
    for (int s0 = 0; s0 < nsteps; s0 += 8) {
      if (s0 + 4 < nsteps) load_group(tab, s0 + 4, 1);
      mma_group(0);
      if (s0 + 4 < nsteps) {
        if (s0 + 8 < nsteps) load_group(tab, s0 + 8, 0);
        mma_group(1);
      }
    }
    if (more) {
      // offsets of the next item -> the other table; its first loads fly during the epilogue
      int32_t* nt = offs[wid][cb ^ 1];
      nt[lane] = na;
      nt[32 + lane] = nb;
      nt[64 + lane] = nc;
      nt[96 + lane] = nd;
      __syncwarp();
      load_group(nt, 0, 0);
    }
    row8 += __shfl_xor_sync(0xffffffffu, row8, 1);
    row8 += __shfl_xor_sync(0xffffffffu, row8, 2);
    col8 += __shfl_xor_sync(0xffffffffu, col8, 1);
    col8 += __shfl_xor_sync(0xffffffffu, col8, 2);
    corner += __shfl_xor_sync(0xffffffffu, corner, 1);
    corner += __shfl_xor_sync(0xffffffffu, corner, 2);
    {
      const int64_t s_off = ((int64_t)(uint32_t)h.x) | ((int64_t)h.y << 32);
      double* out = sd.S + s_off;
      const int bsrc = h.z;
      const bool first = (h.w & 1) && sd.add_b, single = (h.w >> 1) & 1;
      const bool diag = (h.w >> 24) & 1;
      const int toI = (first && diag) ? __ldg(sd.item_toI + item) : 0;
      auto emit = [&](int r, int c, double a) {
        const int idx = r + c * 9;
        double v = -a;
        if (first) {
          if (bsrc >= 0) v += H[bsrc + idx];
          if (diag && r == c) v += dvec[toI + r];
        }
        if (single)
          out[idx] = v;
        else
          atomicAdd(out + idx, v);
      };
      emit(g, 2 * tq, acc0);
      emit(g, 2 * tq + 1, acc1);
      if (tq == 0) emit(8, g, row8);
      if (tq == 1) emit(g, 8, col8);
      if (lane == 2) emit(8, 8, corner);
    }
    if (!more) break;
    item = nitem;
    h = hn;
    cb ^= 1;
  }
}

// s_l += E y_I per E block (camera-major), then z_l = t_l - C^-1 s_l per landmark
__global__ void __launch_bounds__(kGThreads) schur_back_accum_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                                     SchurDev sd, const double* __restrict__ y) {
  __shared__ double stage[kGThreads / 32][32 * 27];
  if (ctrl->done) return;
  const int q = blockIdx.x * kGThreads + threadIdx.x;
  const bool valid = q < sd.n_entries;
  const int qc = valid ? q : sd.n_entries - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* __restrict__ H = sp.H[ctrl->init_idx];
  const int I = __ldg(sd.r_node + qc);
  const int l = __ldg(sd.r_lm + qc);
  const int dI = __ldg(sd.node_dim + I);
  const int eoff = __ldg(sd.r_eoff + qc);
  const double* yi = y + sd.node_toff[I];
  double s0 = 0, s1 = 0, s2 = 0;
  // a warp's 32 blocks are one contiguous run in camera-major order: read them coalesced through shared memory
  const int I0 = __shfl_sync(0xffffffffu, I, 0);
  const int eoff0 = __shfl_sync(0xffffffffu, eoff, 0);
  const bool staged = __all_sync(0xffffffffu, valid && I == I0 && dI == 9 && eoff == eoff0 + 27 * lane);
  if (staged) {
    double* st = stage[warp];
    const double* src = H + eoff0;
    {
      double t[27];
#pragma unroll
      for (int i = 0; i < 27; ++i) t[i] = SFX_LD_STREAM(src + i * 32 + lane);
#pragma unroll
      for (int i = 0; i < 27; ++i) st[i * 32 + lane] = t[i];
    }
    __syncwarp();
    const double* e = st + lane * 27;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      const double yc = yi[c];
      s0 += e[3 * c] * yc;
      s1 += e[3 * c + 1] * yc;
      s2 += e[3 * c + 2] * yc;
    }
  } else if (valid) {
    const double* e = H + eoff;
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < dI) {
        const double yc = yi[c];
        s0 += e[3 * c] * yc;
        s1 += e[3 * c + 1] * yc;
        s2 += e[3 * c + 2] * yc;
      }
  }
  if (!valid) return;
  atomicAdd(sd.sl + (size_t)l * 3, s0);
  atomicAdd(sd.sl + (size_t)l * 3 + 1, s1);
  atomicAdd(sd.sl + (size_t)l * 3 + 2, s2);
}
// ---- TMA-streamed variant of schur_back_accum_kernel ------------------------------------------------------------
// Persistent CTAs walk the slot array of H's camera columns (SchurDev::slot_*) tile by tile: one thread issues a
// cp.async.bulk (TMA, 1-D) of the next tile into a shared-memory ring and arms the stage's mbarrier with the byte
// count; all threads wait on the barrier's phase, take one slot each (27 doubles at stride 27: conflict-free), and a
// block barrier hands the stage back to the producer.  No registers or load instructions are spent on the stream.
__global__ void __launch_bounds__(kTmaSlots, kTmaCtas) schur_back_accum_tma_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                                            SchurDev sd, const double* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
  double* ring = reinterpret_cast<double*>(tma_smem);
  __shared__ __align__(8) uint64_t full[kTmaStages];
  if (ctrl->done) return;
  const double* __restrict__ H = sp.H[ctrl->init_idx] + sd.slot_base;
  const int n_tiles = (sd.n_slots + kTmaSlots - 1) / kTmaSlots;
  const int tid = threadIdx.x;
  auto tile_bytes = [&](int t) {
    const int ns = min(kTmaSlots, sd.n_slots - t * kTmaSlots);
    return (uint32_t)((ns * 27 * 8 + 15) & ~15);  // (an odd tail reads 8 bytes into the landmark blocks behind)
  };
  if (tid == 0) {
    for (int s = 0; s < kTmaStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < kTmaStages; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < n_tiles) {
        mbar_expect_tx(&full[s], tile_bytes(t));
        tma_load_1d(ring + (size_t)s * kTmaSlots * 27, H + (size_t)t * kTmaSlots * 27, tile_bytes(t), &full[s]);
      }
    }
  }
  int it = 0;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const int s = it % kTmaStages;
    const uint32_t parity = (it / kTmaStages) & 1;
    const int slot = t * kTmaSlots + tid;
    int l = -1, I = 0;
    if (slot < sd.n_slots) {
      l = __ldg(sd.slot_lm + slot);
      I = __ldg(sd.slot_node + slot);
    }
    mbar_wait(&full[s], parity);
    if (l >= 0) {
      const double* e = ring + (size_t)s * kTmaSlots * 27 + tid * 27;
      const double* yi = y + sd.node_toff[I];
      double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const double yc = __ldg(yi + c);
        s0 += e[3 * c] * yc;
        s1 += e[3 * c + 1] * yc;
        s2 += e[3 * c + 2] * yc;
      }
      atomicAdd(sd.sl + (size_t)l * 3, s0);
      atomicAdd(sd.sl + (size_t)l * 3 + 1, s1);
      atomicAdd(sd.sl + (size_t)l * 3 + 2, s2);
    }
    __syncthreads();  // every thread is done with the stage
    if (tid == 0) {
      const int tn = t + kTmaStages * gridDim.x;
      if (tn < n_tiles) {
        mbar_expect_tx(&full[s], tile_bytes(tn));
        tma_load_1d(ring + (size_t)s * kTmaSlots * 27, H + (size_t)tn * kTmaSlots * 27, tile_bytes(tn), &full[s]);
      }
    }
  }
}

__global__ void schur_back_final_kernel(const Ctrl* __restrict__ ctrl, SchurDev sd, const double* __restrict__ y,
                                        double* __restrict__ upd) {
  if (ctrl->done) return;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  for (int i = l; i < sd.reduced_dim; i += stride) upd[i] = -y[i];
  if (l >= sd.n_landmarks) return;
  const double* Ci = sd.cinv + (size_t)l * 9;
  const double* t = sd.tl + (size_t)l * 3;
  const double* s = sd.sl + (size_t)l * 3;
  const int to = sd.lm_toff[l];
#pragma unroll
  for (int a = 0; a < 3; ++a) upd[to + a] = -(t[a] - (Ci[a] * s[0] + Ci[a + 3] * s[1] + Ci[a + 6] * s[2]));
}

void launch_schur(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, const SchurDev& sd, const double* dvec) {
  schur_cinv_kernel<<<(sd.n_landmarks + 127) / 128, 128, 0, st>>>(ctrl, sp, sd, dvec); ++g_launches;
  {
    int zg = (int)((sd.s_values + 255) / 256);
    if (zg > 148 * 8) zg = 148 * 8;
    zero_kernel<<<zg, 256, 0, st>>>(ctrl, sd.S, sd.s_values); ++g_launches;
  }
  if (sd.fast3) {
    int n0 = sd.reduced_dim > sd.n_landmarks * 3 ? sd.reduced_dim : sd.n_landmarks * 3;
    int ig = (n0 + 255) / 256;
    if (ig > 148 * 8) ig = 148 * 8;
    if (ig * 256 < sd.reduced_dim) ig = (sd.reduced_dim + 255) / 256;
    schur_rhs_init_kernel<<<ig, 256, 0, st>>>(ctrl, sp, sd); ++g_launches;
    if (sd.wl != nullptr) {
      if (sd.items3 != nullptr && sd.s_diag_off != nullptr) {
        schur_diag_init_kernel<<<(sd.n_reduced_nodes * 81 + 255) / 256, 256, 0, st>>>(ctrl, sp, sd, dvec); ++g_launches;
        // (a cp.async.bulk-streamed variant of this kernel -- two-stage ring, whitening in place, bulk store of W -- was
        // measured at 0.69 ms with three and 0.63 ms with four CTAs per SM against 0.62 ms: the kernel is bound by its
        // per-tile shared-memory work and the read/write mix, not by load issue; see profiles/r02_results.md)
        schur_w_rhs_kernel<true><<<(sd.n_entries + kWThreads * kWChunks - 1) / (kWThreads * kWChunks), kWThreads, 0, st>>>(ctrl, sp, sd);
        ++g_launches;
      } else {
        schur_w_rhs_kernel<false><<<(sd.n_entries + kWThreads - 1) / kWThreads, kWThreads, 0, st>>>(ctrl, sp, sd); ++g_launches;
      }
      if (sd.items3 != nullptr) {
        int grid = sm_count() * SFX_S9_MINB;
        if (grid * kS9Warps > sd.n_items3) grid = (sd.n_items3 + kS9Warps - 1) / kS9Warps;
        if (grid > 0) { schur_s9_kernel<<<grid, kS9Warps * 32, 0, st>>>(ctrl, sp, sd, dvec); ++g_launches; }
        return;
      }
      schur_s2_kernel<<<(sd.n_items + kSchurWarps - 1) / kSchurWarps, kSchurWarps * 32, 0, st>>>(ctrl, sp, sd, dvec);
      ++g_launches;
      return;
    }
    schur_g_rhs_kernel<<<(sd.n_entries + kGThreads - 1) / kGThreads, kGThreads, 0, st>>>(ctrl, sp, sd); ++g_launches;
    schur_s_dmma_kernel<<<(sd.n_items + kSchurWarps - 1) / kSchurWarps, kSchurWarps * 32, 0, st>>>(ctrl, sp, sd, dvec);
    ++g_launches;
    return;
  }
  schur_s_kernel<<<(sd.n_items + kSchurWarps - 1) / kSchurWarps, kSchurWarps * 32, 0, st>>>(ctrl, sp, sd, dvec); ++g_launches;
  schur_rhs_kernel<<<(sd.n_reduced_nodes + 3) / 4, 128, 0, st>>>(ctrl, sp, sd); ++g_launches;
}

void launch_schur_back(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, const SchurDev& sd, const double* y,
                       double* upd) {
  int n = sd.n_landmarks > sd.reduced_dim ? sd.n_landmarks : sd.reduced_dim;
  if (sd.fast3) {
    if (sd.n_slots > 0) {
      static bool configured = false;
      const int smem = kTmaStages * kTmaTileBytes;
      if (!configured) {
        cudaFuncSetAttribute(schur_back_accum_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
      }
      const int n_tiles = (sd.n_slots + kTmaSlots - 1) / kTmaSlots;
      const int grid = n_tiles < sm_count() * kTmaCtas ? n_tiles : sm_count() * kTmaCtas;
      schur_back_accum_tma_kernel<<<grid, kTmaSlots, smem, st>>>(ctrl, sp, sd, y);
    } else {
      schur_back_accum_kernel<<<(sd.n_entries + kGThreads - 1) / kGThreads, kGThreads, 0, st>>>(ctrl, sp, sd, y);
    }
    ++g_launches;
    schur_back_final_kernel<<<(n + 127) / 128, 128, 0, st>>>(ctrl, sd, y, upd); ++g_launches;
    return;
  }
  schur_back_kernel<<<(n + 127) / 128, 128, 0, st>>>(ctrl, sp, sd, y, upd); ++g_launches;
}

// ------------------------------------------------------------------------------------------------
// K4: multifrontal Cholesky, one CTA per front
// ------------------------------------------------------------------------------------------------
constexpr int kFrontThreads = 256;
constexpr int kPanel = 8;

// Partial Cholesky of the leading w columns of the m x m (lower) matrix F with leading dimension ld;
// leaves the Schur complement in F[w:, w:].
__device__ void front_partial_cholesky(double* F, int ld, int m, int w, int* fail) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int kb = 0; kb < w; kb += kPanel) {
    const int nb = min(kPanel, w - kb);
    for (int k = kb; k < kb + nb; ++k) {
      __syncthreads();
      if (tid == 0) {
        double d = F[k + (size_t)k * ld];
        if (!(d > 0.0)) {
          *fail = 1;
          d = __longlong_as_double(0x7ff8000000000000LL);
        }
        F[k + (size_t)k * ld] = sqrt(d);
      }
      __syncthreads();
      const double piv = F[k + (size_t)k * ld];
      for (int i = k + 1 + tid; i < m; i += nt) F[i + (size_t)k * ld] /= piv;
      __syncthreads();
      // update the remaining panel columns
      const int pc = kb + nb - (k + 1);
      if (pc > 0)
        for (int i = k + 1 + tid; i < m; i += nt) {
          const double lik = F[i + (size_t)k * ld];
          for (int j = k + 1; j < kb + nb; ++j)
            if (i >= j) F[i + (size_t)j * ld] -= lik * F[j + (size_t)k * ld];
        }
    }
    __syncthreads();
    // trailing update: columns j >= kb+nb, rows i >= j
    for (int j = kb + nb + warp; j < m; j += nw) {
      double ljk[kPanel];
#pragma unroll
      for (int k = 0; k < kPanel; ++k) ljk[k] = k < nb ? F[j + (size_t)(kb + k) * ld] : 0.0;
      for (int i = j + lane; i < m; i += 32) {
        double v = F[i + (size_t)j * ld];
#pragma unroll
        for (int k = 0; k < kPanel; ++k)
          if (k < nb) v -= F[i + (size_t)(kb + k) * ld] * ljk[k];
        F[i + (size_t)j * ld] = v;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kFrontThreads) front_factor_kernel(Ctrl* ctrl, FrontDev fd,
                                                                     const double* __restrict__ sys_static, StatePtrs sp,
                                                                     int use_state_H, const double* __restrict__ dvec,
                                                                     int lvl_begin, int smem_m_max) {
  extern __shared__ double smem[];
  if (ctrl->done) return;
  const int s = fd.level_fronts[lvl_begin + blockIdx.x];
  const int w = fd.f_w[s], u = fd.f_u[s], m = w + u;
  const double* sys = use_state_H ? sp.H[ctrl->init_idx] : sys_static;
  double* Fg = fd.fronts + fd.f_off[s];
  const bool in_smem = m <= smem_m_max;
  double* F = in_smem ? smem : Fg;
  const int ld = m;
  const int tid = threadIdx.x, nt = blockDim.x;
  // zero (lower part suffices, zero all for simplicity)
  for (int64_t i = tid; i < (int64_t)m * m; i += nt) F[i] = 0.0;
  __syncthreads();
  // system matrix blocks
  for (int ci = fd.f_copy_ptr[s]; ci < fd.f_copy_ptr[s + 1]; ++ci) {
    const FrontCopy c = fd.copies[ci];
    const int ne = c.rows * c.cols;
    for (int e = tid; e < ne; e += nt) {
      const int r = e % c.rows, cc = e / c.rows;
      if (c.lower_only && r < cc) continue;
      const double v = sys[c.src + r + (int64_t)cc * c.src_ld];
      if (c.transposed)
        F[(c.dst_row + cc) + (size_t)(c.dst_col + r) * ld] += v;
      else
        F[(c.dst_row + r) + (size_t)(c.dst_col + cc) * ld] += v;
    }
  }
  __syncthreads();
  if (dvec != nullptr)
    for (int r = tid; r < w; r += nt) F[r + (size_t)r * ld] += dvec[fd.scalar_perm[fd.f_piv[s] + r]];
  __syncthreads();
  // extend-add of the children's update matrices (sequential over children: deterministic)
  for (int ci = fd.f_child_ptr[s]; ci < fd.f_child_ptr[s + 1]; ++ci) {
    const int c = fd.f_child[ci];
    const int wc = fd.f_w[c], uc = fd.f_u[c], mc = wc + uc;
    const double* U = fd.fronts + fd.f_off[c] + wc + (size_t)wc * mc;
    const int32_t* rel = fd.f_rel + fd.f_rows_ptr[c];
    for (int e = tid; e < uc * uc; e += nt) {
      const int i = e % uc, j = e / uc;
      if (i < j) continue;
      F[rel[i] + (size_t)rel[j] * ld] += U[i + (size_t)j * mc];
    }
    __syncthreads();
  }
  front_partial_cholesky(F, ld, m, w, &ctrl->chol_fail);
  if (in_smem) {
    for (int64_t e = tid; e < (int64_t)m * m; e += nt) {
      const int i = (int)(e % m), j = (int)(e / m);
      if (i >= j) Fg[e] = F[e];
    }
  }
}

// One WARP per front for levels whose fronts have at most 32 rows (pose-graph leaves and their parents: 6-12 pivots,
// a few dozen rows; tens of thousands of them per level): the front lives in an 8 KB slice of shared memory, lane i owns
// row i, every step is warp-synchronous (no block barrier), and a CTA of four warps works on four fronts.  Same
// operations on every entry in the same order as front_factor_kernel (updates of an entry arrive by increasing pivot).
constexpr int kWarpFrontM = kWarpFrontRows;
constexpr int kWarpFrontWarps = 4;
__global__ void __launch_bounds__(kWarpFrontWarps * 32) front_factor_warp_kernel(Ctrl* ctrl, FrontDev fd,
                                                                                  const double* __restrict__ sys_static,
                                                                                  StatePtrs sp, int use_state_H,
                                                                                  const double* __restrict__ dvec,
                                                                                  int lvl_begin, int lvl_count) {
  __shared__ double smem_f[kWarpFrontWarps][kWarpFrontM * kWarpFrontM];
  if (ctrl->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = blockIdx.x * kWarpFrontWarps + warp;
  if (fi >= lvl_count) return;
  const int s = fd.level_fronts[lvl_begin + fi];
  const int w = fd.f_w[s], u = fd.f_u[s], m = w + u;
  const double* sys = use_state_H ? sp.H[ctrl->init_idx] : sys_static;
  double* Fg = fd.fronts + fd.f_off[s];
  double* F = smem_f[warp];
  const int ld = m;
  for (int i = lane; i < m * m; i += 32) F[i] = 0.0;
  __syncwarp();
  for (int ci = fd.f_copy_ptr[s]; ci < fd.f_copy_ptr[s + 1]; ++ci) {
    const FrontCopy c = fd.copies[ci];
    const int ne = c.rows * c.cols;
    for (int e = lane; e < ne; e += 32) {
      const int r = e % c.rows, cc = e / c.rows;
      if (c.lower_only && r < cc) continue;
      const double v = sys[c.src + r + (int64_t)cc * c.src_ld];
      if (c.transposed)
        F[(c.dst_row + cc) + (c.dst_col + r) * ld] += v;
      else
        F[(c.dst_row + r) + (c.dst_col + cc) * ld] += v;
    }
    __syncwarp();  // (two blocks of a front never overlap, but keep the order of the block-barrier kernel)
  }
  if (dvec != nullptr && lane < w) F[lane + lane * ld] += dvec[fd.scalar_perm[fd.f_piv[s] + lane]];
  __syncwarp();
  for (int ci = fd.f_child_ptr[s]; ci < fd.f_child_ptr[s + 1]; ++ci) {
    const int c = fd.f_child[ci];
    const int wc = fd.f_w[c], uc = fd.f_u[c], mc = wc + uc;
    const double* U = fd.fronts + fd.f_off[c] + wc + (size_t)wc * mc;
    const int32_t* rel = fd.f_rel + fd.f_rows_ptr[c];
    for (int e = lane; e < uc * uc; e += 32) {
      const int i = e % uc, j = e / uc;
      if (i < j) continue;
      F[rel[i] + rel[j] * ld] += U[i + (size_t)j * mc];
    }
    __syncwarp();
  }
  // partial Cholesky, right-looking, lane = row
  for (int k = 0; k < w; ++k) {
    double d = F[k + k * ld];
    if (!(d > 0.0)) {
      ctrl->chol_fail = 1;
      d = __longlong_as_double(0x7ff8000000000000LL);
    }
    const double piv = sqrt(d);
    __syncwarp();
    double lik = 0.0;
    if (lane == k) F[k + k * ld] = piv;
    if (lane > k && lane < m) {
      lik = F[lane + k * ld] / piv;
      F[lane + k * ld] = lik;
    }
    __syncwarp();
    for (int j = k + 1; j < m; ++j) {
      const double ljk = F[j + k * ld];
      if (lane >= j && lane < m) F[lane + j * ld] -= lik * ljk;
    }
    __syncwarp();
  }
  for (int e = lane; e < m * m; e += 32) {
    const int i = e % m, j = e / m;
    if (i >= j) Fg[e] = F[e];
  }
}

// Threads per CTA of the one-CTA-per-front kernels: levels of tiny fronts (pose-graph leaves: 6 pivots, a few dozen
// rows) run more fronts per SM with 64 or 128 threads than with 256 mostly idle ones
static int small_front_threads(int max_m) { return max_m <= 32 ? 64 : max_m <= 64 ? 128 : kFrontThreads; }

void launch_front_factor(cudaStream_t st, Ctrl* ctrl, const FrontDev& fd, const double* sysvals_static, StatePtrs sp,
                         int use_state_H, const double* dvec, int lvl_begin, int lvl_count, int smem_m_max) {
  // smem_m_max: fronts with m <= smem_m_max are factored in shared memory (chosen per launch)
  static const bool no_warp_fronts = getenv("SFX_NO_WARP_FRONTS") != nullptr;  // A/B knob
  if (smem_m_max <= kWarpFrontM && !no_warp_fronts) {
    front_factor_warp_kernel<<<(lvl_count + kWarpFrontWarps - 1) / kWarpFrontWarps, kWarpFrontWarps * 32, 0, st>>>(
        ctrl, fd, sysvals_static, sp, use_state_H, dvec, lvl_begin, lvl_count);
    ++g_launches;
    return;
  }
  const size_t smem = (size_t)smem_m_max * smem_m_max * sizeof(double);
  front_factor_kernel<<<lvl_count, small_front_threads(smem_m_max), smem, st>>>(ctrl, fd, sysvals_static, sp, use_state_H,
                                                                                dvec, lvl_begin, smem_m_max); ++g_launches;
}

// ------------------------------------------------------------------------------------------------
// K5: supernodal triangular solves
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFrontThreads) front_solve_fwd_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd,
                                                                        const double* __restrict__ rhs_static,
                                                                        StatePtrs sp, int use_state_rhs, int lvl_begin) {
  extern __shared__ double f[];  // m doubles
  if (ctrl->done) return;
  const int s = fd.level_fronts[lvl_begin + blockIdx.x];
  const int w = fd.f_w[s], u = fd.f_u[s], m = w + u;
  const double* rhs = use_state_rhs ? sp.rhs[ctrl->init_idx] : rhs_static;
  const double* L = fd.fronts + fd.f_off[s];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int r = tid; r < m; r += nt) f[r] = r < w ? rhs[fd.scalar_perm[fd.f_piv[s] + r]] : 0.0;
  __syncthreads();
  for (int ci = fd.f_child_ptr[s]; ci < fd.f_child_ptr[s + 1]; ++ci) {
    const int c = fd.f_child[ci];
    const int uc = fd.f_u[c];
    const double* t = fd.twork + fd.f_toff[c];
    const int32_t* rel = fd.f_rel + fd.f_rows_ptr[c];
    for (int q = tid; q < uc; q += nt) f[rel[q]] += t[q];
    __syncthreads();
  }
  // L11 y = f1 (column oriented)
  for (int k = 0; k < w; ++k) {
    if (tid == 0) f[k] /= L[k + (size_t)k * m];
    __syncthreads();
    const double yk = f[k];
    for (int i = k + 1 + tid; i < w; i += nt) f[i] -= L[i + (size_t)k * m] * yk;
    __syncthreads();
  }
  // t = f2 - L21 y1
  for (int q = tid; q < u; q += nt) {
    double v = f[w + q];
    for (int k = 0; k < w; ++k) v -= L[(w + q) + (size_t)k * m] * f[k];
    fd.twork[fd.f_toff[s] + q] = v;
  }
  for (int r = tid; r < w; r += nt) fd.ywork[fd.f_piv[s] + r] = f[r];
}

__global__ void __launch_bounds__(kFrontThreads) front_solve_bwd_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd,
                                                                        int lvl_begin) {
  extern __shared__ double f[];  // m doubles: [x1 (w) | x2 (u)]
  if (ctrl->done) return;
  const int s = fd.level_fronts[lvl_begin + blockIdx.x];
  const int w = fd.f_w[s], u = fd.f_u[s], m = w + u;
  const double* L = fd.fronts + fd.f_off[s];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int32_t* rows = fd.f_rows + fd.f_rows_ptr[s];
  for (int r = tid; r < m; r += nt) f[r] = r < w ? fd.ywork[fd.f_piv[s] + r] : fd.ywork[rows[r - w]];
  __syncthreads();
  // g = y1 - L21^T x2 : one warp per column
  for (int r = warp; r < w; r += nw) {
    double v = 0;
    for (int q = lane; q < u; q += 32) v += L[(w + q) + (size_t)r * m] * f[w + q];
    v = warp_sum(v);
    if (lane == 0) f[r] -= v;
  }
  __syncthreads();
  // L11^T x1 = g (row oriented back substitution)
  for (int k = w - 1; k >= 0; --k) {
    if (tid == 0) f[k] /= L[k + (size_t)k * m];
    __syncthreads();
    const double xk = f[k];
    for (int i = tid; i < k; i += nt) f[i] -= L[k + (size_t)i * m] * xk;
    __syncthreads();
  }
  for (int r = tid; r < w; r += nt) fd.ywork[fd.f_piv[s] + r] = f[r];
}

void launch_front_solve_fwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const double* rhs_static,
                            StatePtrs sp, int use_state_rhs, int lvl_begin, int lvl_count, int smem_bytes) {
  front_solve_fwd_kernel<<<lvl_count, small_front_threads(smem_bytes / 8), smem_bytes, st>>>(
      ctrl, fd, rhs_static, sp, use_state_rhs, lvl_begin);
  ++g_launches;
}
void launch_front_solve_bwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, int lvl_begin, int lvl_count,
                            int smem_bytes) {
  front_solve_bwd_kernel<<<lvl_count, small_front_threads(smem_bytes / 8), smem_bytes, st>>>(ctrl, fd, lvl_begin);
  ++g_launches;
}
// opt in to large dynamic shared memory (process-wide maxima; launches pass their own size)
cudaError_t configure_front_kernels(int smem_m_max, int max_front) {
  static int cur_factor = 0, cur_solve = 48 * 1024;
  cudaError_t e = cudaSuccess;
  const int need_f = smem_m_max * smem_m_max * (int)sizeof(double);
  if (need_f > cur_factor) {
    e = cudaFuncSetAttribute(front_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, need_f);
    if (e != cudaSuccess) return e;
    cur_factor = need_f;
  }
  const int need_s = max_front * (int)sizeof(double);
  if (need_s > cur_solve) {
    e = cudaFuncSetAttribute(front_solve_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, need_s);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(front_solve_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, need_s);
    if (e != cudaSuccess) return e;
    cur_solve = need_s;
  }
  return e;
}

// out[scalar_perm[p]] = scale * ywork[p]
__global__ void unpermute_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, double* __restrict__ out, double scale) {
  if (ctrl->done) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < fd.n) out[fd.scalar_perm[p]] = scale * fd.ywork[p];
}
void launch_unpermute(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, double* out, double scale) {
  unpermute_kernel<<<(fd.n + 255) / 256, 256, 0, st>>>(ctrl, fd, out, scale); ++g_launches;
}

// ------------------------------------------------------------------------------------------------
// K6: retract
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void retract_rot3(const double* a, const double* v, double eps, double* out) {
  const double t0 = sqrt(eps * eps + v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double t1 = 0.5 * t0;
  const double s = sin(t1) / t0, c = cos(t1);
  const double bx = s * v[0], by = s * v[1], bz = s * v[2];
  double r0 = a[0] * c + a[1] * bz - a[2] * by + a[3] * bx;
  double r1 = -a[0] * bz + a[1] * c + a[2] * bx + a[3] * by;
  double r2 = a[2] * c + a[3] * bz + a[0] * by - a[1] * bx;
  double r3 = -a[2] * bz + a[3] * c - a[0] * bx - a[1] * by;
  const double n2 = r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
  if (n2 > 0) {
    const double n = sqrt(n2);
    r0 /= n;
    r1 /= n;
    r2 /= n;
    r3 /= n;
  }
  out[0] = r0;
  out[1] = r1;
  out[2] = r2;
  out[3] = r3;
}

__global__ void retract_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp, const int32_t* __restrict__ key_type,
                               const int32_t* __restrict__ key_voff, const int32_t* __restrict__ key_sdim,
                               const int32_t* __restrict__ key_tdim, const int32_t* __restrict__ key_itoff, int n_keys,
                               const double* __restrict__ upd) {
  if (ctrl->done) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_keys) return;
  const double* src = sp.values[ctrl->init_idx] + key_voff[k];
  double* dst = sp.values[ctrl->new_idx] + key_voff[k];
  const double* v = upd + key_itoff[k];
  const int type = key_type[k];
  const double eps = ctrl->epsilon;
  if (type == SFX_TYPE_VECTOR) {
    const int d = key_tdim[k];
    for (int i = 0; i < d; ++i) dst[i] = src[i] + v[i];
  } else if (type == SFX_TYPE_ROT3) {
    retract_rot3(src, v, eps, dst);
  } else {
    retract_rot3(src, v, eps, dst);
    dst[4] = src[4] + v[3];
    dst[5] = src[5] + v[4];
    dst[6] = src[6] + v[5];
  }
}
void launch_retract(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, const int32_t* key_type, const int32_t* key_voff,
                    const int32_t* key_sdim, const int32_t* key_tdim, const int32_t* key_itoff, int n_keys,
                    const double* upd) {
  retract_kernel<<<(n_keys + 127) / 128, 128, 0, st>>>(ctrl, sp, key_type, key_voff, key_sdim, key_tdim, key_itoff,
                                                       n_keys, upd); ++g_launches;
}

// ------------------------------------------------------------------------------------------------
// K7: reductions + LM bookkeeping
// ------------------------------------------------------------------------------------------------
constexpr int kRedBlocks = 296;
__global__ void __launch_bounds__(256) step_reduce_kernel(const Ctrl* __restrict__ ctrl, StatePtrs sp,
                                                          const double* __restrict__ upd, const double* __restrict__ dvec,
                                                          const double* __restrict__ last, int N,
                                                          double* __restrict__ partials, int i0) {
  if (ctrl->done) return;
  const double* rhs = sp.rhs[ctrl->init_idx];
  double a = 0, b = 0, c = 0, d = 0;
  for (int i = i0 + blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
    const double u = upd[i];
    a += u * (rhs[i] - dvec[i] * u);
    const double l = last[i];
    b += l * u;
    c += l * l;
    d += u * u;
  }
  a = block_sum<256>(a);
  b = block_sum<256>(b);
  c = block_sum<256>(c);
  d = block_sum<256>(d);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 4 + 0] = a;
    partials[blockIdx.x * 4 + 1] = b;
    partials[blockIdx.x * 4 + 2] = c;
    partials[blockIdx.x * 4 + 3] = d;
  }
}
__global__ void step_reduce_final_kernel(Ctrl* ctrl, const double* __restrict__ partials, int nb) {
  if (ctrl->done) return;
  const int q = threadIdx.x;  // 4 threads
  double v = 0;
  for (int i = 0; i < nb; ++i) v += partials[i * 4 + q];
  ctrl->red[1 + q] = v;
}
void launch_step_reduce(cudaStream_t st, Ctrl* ctrl, StatePtrs sp, const double* upd, const double* dvec,
                        const double* last_upd, int N, double* partials, int i0) {
  int nb = (N + 255) / 256;
  if (nb > kRedBlocks) nb = kRedBlocks;
  step_reduce_kernel<<<nb, 256, 0, st>>>(ctrl, sp, upd, dvec, last_upd, N, partials, i0); ++g_launches;
  step_reduce_final_kernel<<<1, 4, 0, st>>>(ctrl, partials, nb); ++g_launches;
}

// state_.Step(); iteration_++   (levenberg_marquardt_solver.tcc:145-149)
__global__ void lm_begin_kernel(Ctrl* c) {
  if (c->done) return;
  const int t = c->init_idx;
  c->init_idx = c->new_idx;
  c->new_idx = t;
  c->iteration++;
}

// after the first linearization of Init: SetBestToInit + stats[-1] (…tcc:151-186)
__global__ void lm_after_first_kernel(Ctrl* c) {
  if (c->done) return;
  // SetBestToInit after EvaluateFirst: on the first Iterate after every Reset / ResetState (…tcc:151-155)
  c->best_valid = 1;
  if (c->best_idx != c->init_idx) {
    if (c->best_idx != c->new_idx) c->free_idx = c->best_idx;
    c->best_idx = c->init_idx;
  }
  if (c->iteration != 0) return;  // FirstIterationStats only when the iteration counter was reset (…tcc:158-186)
  sfx_iteration& it = c->iters[0];
  it.iteration = -1;
  it.update_accepted = 0;
  it.current_lambda = c->lambda;
  it.new_error = c->err[c->init_idx];
  it.new_error_linear = 0;
  it.relative_reduction = 0;
  it.update_angle_change = 0;
  c->n_iters = 1;
  if (!isfinite(c->err[c->init_idx])) {
    c->done = 3;
    c->failure_reason = 2;
  }
}

// everything after Relinearize in Iterate() (…tcc:239-343)
__global__ void lm_end_kernel(Ctrl* c, const double* __restrict__ upd, double* __restrict__ last_upd, int N,
                              int* host_done) {
  if (c->done) return;
  if (threadIdx.x == 0) {
    const sfx_params& p = c->p;
    const double init_error = c->err[c->init_idx];
    const double new_error = c->err[c->new_idx];
    const double eps = c->epsilon;
    const double relative_reduction = (init_error - new_error) / (init_error + eps);
    const double new_error_linear = init_error + 0.5 * c->red[1];
    const double gain_ratio = (init_error - new_error) / (init_error - new_error_linear);
    int status = 0;
    if (relative_reduction > -p.early_exit_min_reduction / 10 && relative_reduction < p.early_exit_min_reduction)
      status = 1;
    else if (new_error < p.early_exit_min_absolute_error)
      status = 1;
    bool accept = relative_reduction > 0;
    double angle = 0;
    if (p.enable_bold_updates && c->have_last_update && !accept) {
      angle = c->red[2] / (sqrt(c->red[3]) * sqrt(c->red[4]));
      accept = ((1 - angle) * (1 - angle) * new_error) <= c->err[c->best_idx];
    }
    if (!accept && c->lambda >= p.lambda_upper_bound) {
      status = 3;
      c->failure_reason = 1;
    }
    if (c->chol_fail) {  // LLT met a non-positive pivot: the step is NaN (rejected below); counted, not hidden
      c->n_chol_fail++;
      c->chol_fail = 0;
    }
    if (!isfinite(c->red[4])) c->n_nonfinite_update++;
    sfx_iteration& it = c->iters[c->n_iters];
    it.iteration = c->iteration;
    it.current_lambda = c->lambda;
    it.new_error = new_error;
    it.new_error_linear = new_error_linear;
    it.relative_reduction = relative_reduction;
    if (!accept) {
      if (p.lambda_update_type == 1) {
        c->lambda *= p.lambda_up_factor;
      } else {
        c->lambda *= c->nu;
        c->nu *= 2;
      }
      const int t = c->init_idx;  // SwapNewAndInit
      c->init_idx = c->new_idx;
      c->new_idx = t;
    } else {
      if (p.lambda_update_type == 1) {
        c->lambda *= p.lambda_down_factor;
      } else {
        c->lambda *= fmax(1.0 / p.dynamic_lambda_update_gamma,
                          1.0 - (p.dynamic_lambda_update_beta - 1) *
                                    pow(2 * gain_ratio - 1, (double)p.dynamic_lambda_update_p));
        c->nu = 2;
      }
      c->have_last_update = 1;
      if (new_error <= c->err[c->best_idx]) {  // SetBestToNew
        c->best_valid = 1;
        if (c->best_idx != c->new_idx) {
          if (c->best_idx != c->init_idx) c->free_idx = c->best_idx;
          c->best_idx = c->new_idx;
        }
        c->best_index = c->n_iters;
      }
      if (c->best_idx == c->init_idx) {  // SetInitToNotBest
        c->init_idx = c->free_idx;
        c->free_idx = c->best_idx;
      }
    }
    c->lambda = fmin(fmax(c->lambda, p.lambda_lower_bound), p.lambda_upper_bound);
    it.update_angle_change = angle;
    it.update_accepted = accept ? 1 : 0;
    c->n_iters++;
    if (status) {
      if (status != 3) c->failure_reason = 0;
      c->done = status;
      *host_done = status;
    }
  }
}

// last_update_ = update_ when the step was accepted (the entry just written says so)
__global__ void save_last_update_kernel(const Ctrl* __restrict__ c, const double* __restrict__ upd,
                                        double* __restrict__ last_upd, int N) {
  if (c->n_iters < 1 || !c->iters[c->n_iters - 1].update_accepted) return;
  if (!c->p.enable_bold_updates) return;  // last_update_ is only read by the bold-update rule
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) last_upd[i] = upd[i];
}

void launch_lm_begin(cudaStream_t st, Ctrl* ctrl) { lm_begin_kernel<<<1, 1, 0, st>>>(ctrl); ++g_launches; }
void launch_lm_after_first_linearize(cudaStream_t st, Ctrl* ctrl) { lm_after_first_kernel<<<1, 1, 0, st>>>(ctrl); ++g_launches; }
void launch_lm_end(cudaStream_t st, Ctrl* ctrl, const double* upd, double* last_upd, int N, int* host_done) {
  lm_end_kernel<<<1, 32, 0, st>>>(ctrl, upd, last_upd, N, host_done); ++g_launches;
  int grid = (N + 255) / 256;
  if (grid > 296) grid = 296;
  save_last_update_kernel<<<grid, 256, 0, st>>>(ctrl, upd, last_upd, N); ++g_launches;
}

// ------------------------------------------------------------------------------------------------
// misc: copies / exports (parity hooks)
// ------------------------------------------------------------------------------------------------
__global__ void copy_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
// debug_stats: optimization_iteration_t::values / residual / update of the record about to be written
// (levenberg_marquardt_solver.tcc:166-171 for the record of iteration -1 = Init, :115-119 for New; the update in the
// reference's tangent order, zeros in the record of iteration -1)
__global__ void debug_snapshot_kernel(const Ctrl* __restrict__ c, StatePtrs sp, int first, int64_t n_values, int M,
                                      int cap, double* __restrict__ dv, double* __restrict__ dr,
                                      const double* __restrict__ upd, const int32_t* __restrict__ ref2int, int N,
                                      double* __restrict__ du) {
  if (c->done) return;
  if (first && c->iteration != 0) return;  // the record of iteration -1 exists only after a Reset
  const int blk = first ? c->init_idx : c->new_idx;
  const int rec = first ? 0 : c->n_iters;
  if (rec >= cap) return;
  const double* v = sp.values[blk];
  const double* r = sp.res[blk];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_values; i += stride) dv[(size_t)rec * n_values + i] = v[i];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) dr[(size_t)rec * M + i] = r[i];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
    du[(size_t)rec * N + i] = first ? 0.0 : upd[ref2int[i]];
}
void launch_debug_snapshot(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int first, int64_t n_values, int M, int cap,
                           double* dv, double* dr, const double* upd, const int32_t* ref2int, int N, double* du) {
  int64_t work = n_values > M ? n_values : M;
  if (N > work) work = N;
  int grid = (int)((work + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  debug_snapshot_kernel<<<grid, 256, 0, st>>>(ctrl, sp, first, n_values, M, cap, dv, dr, upd, ref2int, N, du); ++g_launches;
}

void launch_copy_values(cudaStream_t st, double* dst, const double* src, int64_t n) {
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  copy_kernel<<<grid, 256, 0, st>>>(dst, src, n); ++g_launches;
}
__global__ void export_csc_kernel(const double* __restrict__ Hv, const int32_t* __restrict__ src, int64_t nnz,
                                  double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) out[i] = Hv[src[i]];
}
void launch_export_csc(cudaStream_t st, const double* Hvals, const int32_t* csc_src, int64_t nnz, double* out) {
  int grid = (int)((nnz + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  export_csc_kernel<<<grid, 256, 0, st>>>(Hvals, csc_src, nnz, out); ++g_launches;
}
// out[ref] = in[ref2int[ref]]
__global__ void import_csc_kernel(const double* __restrict__ in, const int32_t* __restrict__ src, int64_t nnz,
                                  double* __restrict__ Hv) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) Hv[src[i]] = in[i];
}
void launch_import_csc(cudaStream_t st, const double* in, const int32_t* csc_src, int64_t nnz, double* Hvals) {
  int grid = (int)((nnz + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  import_csc_kernel<<<grid, 256, 0, st>>>(in, csc_src, nnz, Hvals); ++g_launches;
}
__global__ void set_unit_kernel(double* v, int j, int prev) {
  if (prev >= 0) v[prev] = 0.0;
  v[j] = 1.0;
}
void launch_set_unit(cudaStream_t st, double* v, int j, int prev) { set_unit_kernel<<<1, 1, 0, st>>>(v, j, prev); ++g_launches; }
__global__ void set_entry_kernel(double* v, int j, double value, int prev) {
  if (prev >= 0) v[prev] = 0.0;
  v[j] = value;
}
void launch_set_entry(cudaStream_t st, double* v, int j, double value, int prev) {
  set_entry_kernel<<<1, 1, 0, st>>>(v, j, value, prev); ++g_launches;
}

__global__ void permute_vec_kernel(const double* __restrict__ in, const int32_t* __restrict__ ref2int, int N,
                                   double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = in[ref2int[i]];
}
void launch_permute_vec(cudaStream_t st, const double* in, const int32_t* ref2int, int N, double* out) {
  permute_vec_kernel<<<(N + 255) / 256, 256, 0, st>>>(in, ref2int, N, out); ++g_launches;
}

}  // namespace sfx

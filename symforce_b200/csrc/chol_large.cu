// Large frontal matrices: tiled right-looking partial Cholesky executed as a task DAG by a
// persistent kernel (one launch per tree level), FP64 tensor-core (DMMA m8n8k4) tile products.
//
// A front with m = w + u rows is cut into 64-wide tiles, tile boundaries aligned to the pivot
// width w.  Tasks, in an order that keeps every dependency earlier in the list:
//   POTRF(k)      : factor diagonal tile (k,k), also forms L_kk^-1 (kept for TRSM and the solves)
//   TRSM(i,k)     : tile(i,k) <- tile(i,k) * L_kk^-T          (as a GEMM with L_kk^-1)
//   UPDATE(i,j,k) : tile(i,j) <- tile(i,j) - tile(i,k) tile(j,k)^T
// Each tile carries a version counter in global memory (number of updates applied, +1 once
// final); CTAs pull tasks from a queue and spin on the counters of their inputs
// (ld.acquire / st.release), so the diagonal critical path overlaps the trailing updates without
// kernel-launch boundaries.  Tile data is read with ld.global.cg (L2 is the coherence point).
#include <cstdio>

#include "kernels.cuh"

namespace sfx {

extern int64_t g_launches;

constexpr int kT = 64;        // tile size
constexpr int kLd = 68;       // smem leading dimension (== 4 mod 16: conflict-free DMMA fragment loads)
constexpr int kLargeThreads = 256;

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int tile_start(const LargeFront& lf, int t) {
  return t < lf.wt ? t * kT : lf.w + (t - lf.wt) * kT;
}
__device__ __forceinline__ int tile_size(const LargeFront& lf, int t) {
  return t < lf.wt ? min(kT, lf.w - t * kT) : min(kT, lf.m - (lf.w + (t - lf.wt) * kT));
}

// smem tile <- global (zero padded); identity padding on the diagonal when `ident`
__device__ __forceinline__ void load_tile(double* S, const double* __restrict__ G, int ldg, int nr, int nc, bool ident) {
  const int r = threadIdx.x & 63;
  for (int c = threadIdx.x >> 6; c < kT; c += kLargeThreads / 64) {
    double v = 0.0;
    if (r < nr && c < nc)
      v = __ldcg(G + r + (size_t)c * ldg);
    else if (ident && r == c)
      v = 1.0;
    S[r + c * kLd] = v;
  }
}

// acc += A * B^T over the 64-deep smem tiles; warp layout 4 (rows) x 2 (cols), warp tile 16 x 32
__device__ __forceinline__ void tile_gemm(const double* As, const double* Bs, double (&acc)[2][4][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = warp & 3, wc = warp >> 2;
  const int g = lane >> 2, tq = lane & 3;
  const double* ap = As + (wr * 16 + g) + tq * kLd;
  const double* bp = Bs + (wc * 32 + g) + tq * kLd;
#pragma unroll 4
  for (int kk = 0; kk < kT; kk += 4) {
    double a[2], b[4];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) a[rb] = ap[rb * 8 + kk * kLd];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) b[cb] = bp[cb * 8 + kk * kLd];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) mma884(acc[rb][cb][0], acc[rb][cb][1], a[rb], b[cb]);
  }
}

__global__ void __launch_bounds__(kLargeThreads) large_factor_kernel(Ctrl* ctrl, FrontDev fd, LargeDev ld, int t0,
                                                                     int t1, int level) {
  extern __shared__ double sm[];
  double* As = sm;
  double* Bs = sm + kT * kLd;
  __shared__ int s_task;
  if (ctrl->done) return;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wr = warp & 3, wc = warp >> 2;
  const int g = lane >> 2, tq = lane & 3;
  for (;;) {
    if (tid == 0) s_task = t0 + atomicAdd(&ld.queue[level], 1);
    __syncthreads();
    const int t = s_task;
    __syncthreads();
    if (t >= t1) break;
    const LargeTask task = ld.tasks[t];
    const LargeFront lf = ld.lf[task.lf];
    double* F = fd.fronts + lf.off;
    const int m = lf.m, nt = lf.nt;
    int* cnt = ld.counters + lf.cnt_off;
    const int k = task.k, i = task.i, j = task.j;
    if (task.type == 0) {
      // ---------------- POTRF(k) ----------------
      if (tid == 0)
        while (ld_acquire(cnt + k * nt + k) != k) __nanosleep(40);
      __syncthreads();
      const int s0 = tile_start(lf, k), nb = tile_size(lf, k);
      load_tile(As, F + s0 + (size_t)s0 * m, m, nb, nb, true);
      __syncthreads();
      // Cholesky of the tile fused with the forward substitution L X = I (X = L^-1 in Bs, layout
      // X[r + c*kLd]): per column one pivot, one scaling, one rank-1 update of both A and X.
      double* X = Bs;
      {
        const int r = tid & 63;
        for (int c = tid >> 6; c < kT; c += kLargeThreads / 64) X[r + c * kLd] = (r == c) ? 1.0 : 0.0;
      }
      __syncthreads();
      for (int c = 0; c < kT; ++c) {
        if (tid == 0) {
          double d = As[c + c * kLd];
          if (!(d > 0.0)) {
            ctrl->chol_fail = 1;
            d = __longlong_as_double(0x7ff8000000000000LL);
          }
          As[c + c * kLd] = sqrt(d);
        }
        __syncthreads();
        const double pinv = 1.0 / As[c + c * kLd];
        if (tid > c && tid < kT) As[tid + c * kLd] *= pinv;        // column c of L
        if (tid >= 64 && tid < 64 + c + 1) X[c + (tid - 64) * kLd] *= pinv;  // row c of X (cols 0..c)
        __syncthreads();
        {
          const int r = tid & 63;
          if (r > c) {
            const double lrc = As[r + c * kLd];
            for (int cc = c + 1 + (tid >> 6); cc <= r; cc += kLargeThreads / 64) As[r + cc * kLd] -= lrc * As[cc + c * kLd];
            for (int cc = (tid >> 6); cc <= c; cc += kLargeThreads / 64) X[r + cc * kLd] -= lrc * X[c + cc * kLd];
          }
        }
        __syncthreads();
      }
      double* linv = ld.linv + lf.linv_off + (size_t)k * kT * kT;
      {
        const int r = tid & 63;
        for (int c = tid >> 6; c < kT; c += kLargeThreads / 64) {
          linv[r + c * kT] = X[r + c * kLd];
          if (r < nb && c < nb && r >= c) F[(s0 + r) + (size_t)(s0 + c) * m] = As[r + c * kLd];
        }
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(cnt + k * nt + k, k + 1);
      }
    } else {
      // ---------------- TRSM(i,k) / UPDATE(i,j,k) ----------------
      const bool trsm = task.type == 1;
      if (tid == 0) {
        if (trsm) {
          while (ld_acquire(cnt + k * nt + k) < k + 1) __nanosleep(40);
          while (ld_acquire(cnt + i * nt + k) != k) __nanosleep(40);
        } else {
          while (ld_acquire(cnt + i * nt + k) < k + 1) __nanosleep(40);
          while (ld_acquire(cnt + j * nt + k) < k + 1) __nanosleep(40);
          while (ld_acquire(cnt + i * nt + j) != k) __nanosleep(40);
        }
      }
      __syncthreads();
      const int ri = tile_start(lf, i), ni = tile_size(lf, i);
      const int ck = tile_start(lf, k), nk = tile_size(lf, k);
      load_tile(As, F + ri + (size_t)ck * m, m, ni, nk, false);
      int cj, nj;
      if (trsm) {
        cj = ck;
        nj = nk;
        load_tile(Bs, ld.linv + lf.linv_off + (size_t)k * kT * kT, kT, kT, kT, false);
      } else {
        cj = tile_start(lf, j);
        nj = tile_size(lf, j);
        load_tile(Bs, F + cj + (size_t)ck * m, m, nj, nk, false);
      }
      __syncthreads();
      double acc[2][4][2];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) acc[rb][cb][0] = acc[rb][cb][1] = 0.0;
      double* C = F + ri + (size_t)cj * m;
      // all C loads first (independent, L2 latency overlapped), then subtract and store
      double cin[2][4][2];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = wr * 16 + rb * 8 + g;
            const int c = wc * 32 + cb * 8 + tq * 2 + e;
            cin[rb][cb][e] = (!trsm && r < ni && c < nj) ? __ldcg(C + r + (size_t)c * m) : 0.0;
          }
      tile_gemm(As, Bs, acc);
#pragma unroll
      for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = wr * 16 + rb * 8 + g;
            const int c = wc * 32 + cb * 8 + tq * 2 + e;
            if (r < ni && c < nj) C[r + (size_t)c * m] = trsm ? acc[rb][cb][e] : cin[rb][cb][e] - acc[rb][cb][e];
          }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(cnt + i * nt + (trsm ? k : j), k + 1);
      }
    }
  }
}

// ---- assembly of large fronts ---------------------------------------------------------------------
__global__ void large_zero_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld, int lf0) {
  if (ctrl->done) return;
  const LargeFront lf = ld.lf[lf0 + blockIdx.y];
  double* F = fd.fronts + lf.off;
  const int64_t n = (int64_t)lf.m * lf.m;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) F[e] = 0.0;
}

// one warp per job: system-matrix block copy, damping, or a column range of a child's update matrix
__global__ void __launch_bounds__(256) large_assemble_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld,
                                                              const double* __restrict__ sys_static, StatePtrs sp,
                                                              int use_state_H, const double* __restrict__ dvec, int j0,
                                                              int j1) {
  if (ctrl->done) return;
  const int lane = threadIdx.x & 31;
  const int job_id = j0 + blockIdx.x * 8 + (threadIdx.x >> 5);
  if (job_id >= j1) return;
  const LargeJob job = ld.jobs[job_id];
  const LargeFront lf = ld.lf[job.lf];
  double* F = fd.fronts + lf.off;
  const int m = lf.m;
  if (job.type == 0) {
    const double* sys = use_state_H ? sp.H[ctrl->init_idx] : sys_static;
    const FrontCopy c = fd.copies[job.idx];
    const int ne = c.rows * c.cols;
    for (int e = lane; e < ne; e += 32) {
      const int r = e % c.rows, cc = e / c.rows;
      if (c.lower_only && r < cc) continue;
      const double v = sys[c.src + r + (int64_t)cc * c.src_ld];
      double* dst = c.transposed ? F + (c.dst_row + cc) + (size_t)(c.dst_col + r) * m
                                 : F + (c.dst_row + r) + (size_t)(c.dst_col + cc) * m;
      atomicAdd(dst, v);
    }
  } else if (job.type == 2) {
    if (dvec != nullptr)
      for (int r = job.c0 + lane; r < job.c1; r += 32)
        atomicAdd(F + r + (size_t)r * m, dvec[fd.scalar_perm[fd.f_piv[lf.front] + r]]);
  } else {
    const int c = job.idx;
    const int wc = fd.f_w[c], uc = fd.f_u[c], mc = wc + uc;
    const double* U = fd.fronts + fd.f_off[c] + wc + (size_t)wc * mc;
    const int32_t* rel = fd.f_rel + fd.f_rows_ptr[c];
    for (int jc = job.c0; jc < job.c1; ++jc) {
      const int dj = rel[jc];
      for (int ic = jc + lane; ic < uc; ic += 32) atomicAdd(F + rel[ic] + (size_t)dj * m, U[ic + (size_t)jc * mc]);
    }
  }
}

void launch_large_level(cudaStream_t st, Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld, const LargeLevel& lv,
                        int level, const double* sys_static, StatePtrs sp, int use_state_H, const double* dvec) {
  if (lv.n_lf == 0) return;
  dim3 zg(96, lv.n_lf);
  large_zero_kernel<<<zg, 256, 0, st>>>(ctrl, fd, ld, lv.lf0); ++g_launches;
  const int nj = lv.j1 - lv.j0;
  if (nj > 0) {
    large_assemble_kernel<<<(nj + 7) / 8, 256, 0, st>>>(ctrl, fd, ld, sys_static, sp, use_state_H, dvec, lv.j0, lv.j1);
    ++g_launches;
  }
  const int ntask = lv.t1 - lv.t0;
  int grid = ntask < 148 * 3 ? ntask : 148 * 3;
  const size_t smem = 2 * kT * kLd * sizeof(double);
  large_factor_kernel<<<grid, kLargeThreads, smem, st>>>(ctrl, fd, ld, lv.t0, lv.t1, level); ++g_launches;
}

cudaError_t configure_large_kernels() {
  return cudaFuncSetAttribute(large_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(2 * kT * kLd * sizeof(double)));
}

// ---- triangular solves on large fronts (one CTA per front, panels of 64 with L_kk^-1) ------------
__global__ void __launch_bounds__(256) large_solve_fwd_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld,
                                                               const double* __restrict__ rhs_static, StatePtrs sp,
                                                               int use_state_rhs, int lf0) {
  extern __shared__ double f[];  // m + 64
  if (ctrl->done) return;
  const LargeFront lf = ld.lf[lf0 + blockIdx.x];
  const int s = lf.front;
  const int w = lf.w, m = lf.m, u = m - w;
  double* ytmp = f + m;
  const double* rhs = use_state_rhs ? sp.rhs[ctrl->init_idx] : rhs_static;
  const double* L = fd.fronts + lf.off;
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
  for (int kt = 0; kt < lf.wt; ++kt) {
    const int c0 = kt * kT, nb = min(kT, w - c0);
    const double* linv = ld.linv + lf.linv_off + (size_t)kt * kT * kT;
    if (tid < nb) {
      double v = 0.0;
      for (int q = 0; q <= tid; ++q) v += linv[tid + q * kT] * f[c0 + q];
      ytmp[tid] = v;
    }
    __syncthreads();
    if (tid < nb) f[c0 + tid] = ytmp[tid];
    for (int r = c0 + nb + tid; r < m; r += nt) {
      double v = f[r];
      for (int q = 0; q < nb; ++q) v -= L[r + (size_t)(c0 + q) * m] * ytmp[q];
      f[r] = v;
    }
    __syncthreads();
  }
  for (int q = tid; q < u; q += nt) fd.twork[fd.f_toff[s] + q] = f[w + q];
  for (int r = tid; r < w; r += nt) fd.ywork[fd.f_piv[s] + r] = f[r];
}

__global__ void __launch_bounds__(256) large_solve_bwd_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld,
                                                               int lf0) {
  extern __shared__ double f[];  // m + 64
  if (ctrl->done) return;
  const LargeFront lf = ld.lf[lf0 + blockIdx.x];
  const int s = lf.front;
  const int w = lf.w, m = lf.m;
  double* gt = f + m;
  const double* L = fd.fronts + lf.off;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int32_t* rows = fd.f_rows + fd.f_rows_ptr[s];
  for (int r = tid; r < m; r += nt) f[r] = r < w ? fd.ywork[fd.f_piv[s] + r] : fd.ywork[rows[r - w]];
  __syncthreads();
  for (int kt = lf.wt - 1; kt >= 0; --kt) {
    const int c0 = kt * kT, nb = min(kT, w - c0);
    // g_c = y_c - sum_{r >= c0+nb} L[r, c0+c] x_r   (one warp per column)
    for (int c = warp; c < nb; c += nw) {
      double v = 0.0;
      for (int r = c0 + nb + lane; r < m; r += 32) v += L[r + (size_t)(c0 + c) * m] * f[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) gt[c] = f[c0 + c] - v;
    }
    __syncthreads();
    // x = L_kk^-T g
    const double* linv = ld.linv + lf.linv_off + (size_t)kt * kT * kT;
    if (tid < nb) {
      double v = 0.0;
      for (int r = tid; r < nb; ++r) v += linv[r + tid * kT] * gt[r];
      f[c0 + tid] = v;
    }
    __syncthreads();
  }
  for (int r = tid; r < w; r += nt) fd.ywork[fd.f_piv[s] + r] = f[r];
}

void launch_large_solve_fwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld,
                            const LargeLevel& lv, const double* rhs_static, StatePtrs sp, int use_state_rhs) {
  if (lv.n_lf == 0) return;
  const size_t smem = (size_t)(lv.max_m + 64) * sizeof(double);
  large_solve_fwd_kernel<<<lv.n_lf, 256, smem, st>>>(ctrl, fd, ld, rhs_static, sp, use_state_rhs, lv.lf0); ++g_launches;
}
void launch_large_solve_bwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld,
                            const LargeLevel& lv) {
  if (lv.n_lf == 0) return;
  const size_t smem = (size_t)(lv.max_m + 64) * sizeof(double);
  large_solve_bwd_kernel<<<lv.n_lf, 256, smem, st>>>(ctrl, fd, ld, lv.lf0); ++g_launches;
}

}  // namespace sfx

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
#include <cstdlib>

#include "kernels.cuh"

namespace sfx {

extern int64_t g_launches;

constexpr int kT = 64;        // tile size
constexpr int kLd = 68;       // smem leading dimension (== 4 mod 16: conflict-free DMMA fragment loads)
constexpr int kLargeThreads = 256;
constexpr size_t kFactorSmem = 2 * 64 * 68 * sizeof(double);  // dynamic shared memory of large_factor_kernel: As | Bs

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
  const int r = threadIdx.x & 63, c0 = threadIdx.x >> 6;
  constexpr int kPer = kT / (kLargeThreads / 64);  // 16 columns per thread: all loads in flight before the first store
  double v[kPer];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int c = c0 + q * (kLargeThreads / 64);
    v[q] = (r < nr && c < nc) ? __ldcg(G + r + (size_t)c * ldg) : ((ident && r == c) ? 1.0 : 0.0);
  }
#pragma unroll
  for (int q = 0; q < kPer; ++q) S[r + (c0 + q * (kLargeThreads / 64)) * kLd] = v[q];
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

// 64x64 Cholesky of the tile in shared memory (lower part valid, ld = kLd), panels of 8 columns.
// A panel is factored by ONE warp with shuffles only (lane l holds rows l and l+32 of the panel in
// registers; per column: broadcast the pivot, rsqrt on every lane, scale, rank-1 update inside the
// panel with the pivot row's entries broadcast from their owner lane) -- no block barrier inside the
// 8-column chain.  The other warps join for the rank-8 trailing update; two barriers per panel.
constexpr int kPW = 8;
__device__ __forceinline__ void potrf_64_panels(double* As, int* fail) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
  for (int p = 0; p < kT / kPW; ++p) {
    const int c0 = p * kPW;
    if (warp == (p & 7)) {
      double a0[kPW], a1[kPW];
#pragma unroll
      for (int j = 0; j < kPW; ++j) {
        a0[j] = As[lane + (c0 + j) * kLd];
        a1[j] = As[lane + 32 + (c0 + j) * kLd];
      }
      const bool hi = c0 >= 32;  // pivot rows of this panel live in slot 1 (rows 32..63)
#pragma unroll
      for (int j = 0; j < kPW; ++j) {
        const int c = c0 + j;
        double d = __shfl_sync(0xffffffffu, hi ? a1[j] : a0[j], c & 31);
        if (!(d > 0.0)) {
          *fail = 1;
          d = __longlong_as_double(0x7ff8000000000000LL);
        }
        const double pinv = rsqrt(d);
        if (lane >= c) a0[j] *= pinv;       // rows lane (< 32)
        if (lane + 32 >= c) a1[j] *= pinv;  // rows lane + 32
#pragma unroll
        for (int jj = j + 1; jj < kPW; ++jj) {
          const int rr = c0 + jj;  // pivot-panel row whose entry in column c multiplies column jj
          const double l = __shfl_sync(0xffffffffu, hi ? a1[j] : a0[j], rr & 31);
          if (lane > c) a0[jj] -= a0[j] * l;
          if (lane + 32 > c) a1[jj] -= a1[j] * l;
        }
      }
#pragma unroll
      for (int j = 0; j < kPW; ++j) {
        As[lane + (c0 + j) * kLd] = lane >= c0 + j ? a0[j] : 0.0;
        As[lane + 32 + (c0 + j) * kLd] = lane + 32 >= c0 + j ? a1[j] : 0.0;
      }
    }
    __syncthreads();
    // trailing update: A[r][cc] -= sum_j L[r][c0+j] L[cc][c0+j] for cc >= c0 + 8, r >= cc
    {
      const int r = tid & 63;
      double lr[kPW];
#pragma unroll
      for (int j = 0; j < kPW; ++j) lr[j] = As[r + (c0 + j) * kLd];
      for (int cc = c0 + kPW + (tid >> 6); cc < kT; cc += kLargeThreads / 64)
        if (r >= cc) {
          double v = As[r + cc * kLd];
#pragma unroll
          for (int j = 0; j < kPW; ++j) v -= lr[j] * As[cc + (c0 + j) * kLd];
          As[r + cc * kLd] = v;
        }
    }
    __syncthreads();
  }
}

// ---- X = L^-1 for the 64x64 lower-triangular tile in shared memory, blocked by 16 ------------------
// diagonal 16x16 blocks by forward substitution (one warp each, one column per lane, registers);
// off-diagonal blocks level by level with FP64 tensor-core products:
//   X_ij = -X_ii * sum_{k=j}^{i-1} L_ik X_kj
// acc += A(16x16) * B(16x16), both column-major in shared memory (one warp)
__device__ __forceinline__ void warp_gemm16(const double* A, const double* B, double (&acc)[2][2][2]) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int kk = 0; kk < 16; kk += 4) {
    const double a0 = A[g + (kk + tq) * kLd], a1 = A[8 + g + (kk + tq) * kLd];
    const double b0 = B[(kk + tq) + g * kLd], b1 = B[(kk + tq) + (8 + g) * kLd];
    mma884(acc[0][0][0], acc[0][0][1], a0, b0);
    mma884(acc[0][1][0], acc[0][1][1], a0, b1);
    mma884(acc[1][0][0], acc[1][0][1], a1, b0);
    mma884(acc[1][1][0], acc[1][1][1], a1, b1);
  }
}
__device__ __forceinline__ void warp_store16(double* C, const double (&acc)[2][2][2], double scale) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int rb = 0; rb < 2; ++rb)
#pragma unroll
    for (int cb = 0; cb < 2; ++cb)
#pragma unroll
      for (int e = 0; e < 2; ++e) C[(rb * 8 + g) + (cb * 8 + tq * 2 + e) * kLd] = scale * acc[rb][cb][e];
}
// Ls: L (lower, zero above the diagonal), Xs: output (full tile written, zero above the diagonal)
__device__ void tri_inverse_64(const double* Ls, double* Xs) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // zero X
  for (int e = tid; e < kT * kT; e += kLargeThreads) Xs[(e & 63) + (e >> 6) * kLd] = 0.0;
  __syncthreads();
  // diagonal blocks: warp w < 4, lane j < 16 solves L_ww x = e_j
  if (warp < 4 && lane < 16) {
    const double* Lb = Ls + 16 * warp + 16 * warp * kLd;
    double x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double v = (r == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if (q < r) v -= Lb[r + q * kLd] * x[q];
      x[r] = v / Lb[r + r * kLd];
    }
    double* Xb = Xs + 16 * warp + (16 * warp + lane) * kLd;
#pragma unroll
    for (int r = 0; r < 16; ++r) Xb[r] = (r >= lane) ? x[r] : 0.0;
  }
  __syncthreads();
  // off-diagonal blocks, distance d = i - j; scratch for T = sum L_ik X_kj is the (unused) mirror block (j, i)
#pragma unroll 1
  for (int d = 1; d < 4; ++d) {
    if (warp < 4 - d) {
      const int j = warp, i = warp + d;
      double acc[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
      for (int k = j; k < i; ++k) warp_gemm16(Ls + 16 * i + 16 * k * kLd, Xs + 16 * k + 16 * j * kLd, acc);
      double* T = Xs + 16 * j + 16 * i * kLd;
      warp_store16(T, acc, 1.0);
      __syncwarp();
      double acc2[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
      warp_gemm16(Xs + 16 * i + 16 * i * kLd, T, acc2);
      __syncwarp();
      warp_store16(Xs + 16 * i + 16 * j * kLd, acc2, -1.0);
    }
    __syncthreads();
  }
  // clear the scratch (upper blocks)
  for (int e = tid; e < kT * kT; e += kLargeThreads) {
    const int r = e & 63, c = e >> 6;
    if ((r >> 4) < (c >> 4)) Xs[r + c * kLd] = 0.0;
  }
  __syncthreads();
}

// tile(i, jt) <- (trsm ? 0 : tile(i, jt)) -/+ A * B^T with A, B already in shared memory
__device__ __forceinline__ void gemm_store(double* F, int m, const LargeFront& lf, int i, int jt, const double* As,
                                           const double* Bs, bool trsm, double* keep /* optional smem copy */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = warp & 3, wc = warp >> 2;
  const int g = lane >> 2, tq = lane & 3;
  const int ri = tile_start(lf, i), ni = tile_size(lf, i);
  const int cj = tile_start(lf, jt), nj = tile_size(lf, jt);
  double acc[2][4][2];
#pragma unroll
  for (int rb = 0; rb < 2; ++rb)
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) acc[rb][cb][0] = acc[rb][cb][1] = 0.0;
  double* C = F + ri + (size_t)cj * m;
  tile_gemm(As, Bs, acc);
  if (!trsm) {
#pragma unroll
    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
      for (int cb = 0; cb < 4; ++cb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = wr * 16 + rb * 8 + g;
          const int c = wc * 32 + cb * 8 + tq * 2 + e;
          const double cin = (r < ni && c < nj) ? __ldcg(C + r + (size_t)c * m) : 0.0;
          acc[rb][cb][e] = cin - acc[rb][cb][e];
        }
  }
  if (keep != nullptr) __syncthreads();  // everyone finished reading As/Bs before `keep` (may alias) is written
#pragma unroll
  for (int rb = 0; rb < 2; ++rb)
#pragma unroll
    for (int cb = 0; cb < 4; ++cb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = wr * 16 + rb * 8 + g;
        const int c = wc * 32 + cb * 8 + tq * 2 + e;
        const bool in = r < ni && c < nj;
        if (in) C[r + (size_t)c * m] = acc[rb][cb][e];
        if (keep != nullptr) keep[r + c * kLd] = in ? acc[rb][cb][e] : 0.0;
      }
}

__device__ __forceinline__ void publish(int* flag, int v) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    st_release(flag, v);
  }
}
__device__ __forceinline__ void wait_eq(const int* flag, int v) {
  if (threadIdx.x == 0)
    while (ld_acquire(flag) != v) __nanosleep(32);
  __syncthreads();
}
__device__ __forceinline__ void wait_ge(const int* flag, int v) {
  if (threadIdx.x == 0)
    while (ld_acquire(flag) < v) __nanosleep(32);
  __syncthreads();
}

// ---- range updates: tile(i,j) -= sum_{kk in [k,k1)} L(i,kk) L(j,kk)^T with the accumulator in registers.
// The operand tiles are consumed as 64 x 32 half tiles, double buffered in shared memory; the next
// half tile pair is loaded into registers (ld.global.cg: tiles are produced by other CTAs of the same
// launch, and tile boundaries are not sector aligned, so L1 must not be involved) while the tensor
// cores work on the current one.
constexpr int kHalf = 32;
constexpr int kHalfDoubles = kHalf * kLd;
constexpr int kHalfPerThread = kT * kHalf / kLargeThreads;  // 8
__device__ __forceinline__ void load_half(double (&v)[kHalfPerThread], const double* F, int m, const LargeFront& lf,
                                          int t, int kk, int c0) {
  const int nr = tile_size(lf, t), nc = tile_size(lf, kk);
  const double* G = F + tile_start(lf, t) + (size_t)(kk * kT + c0) * m;
  const int r = threadIdx.x & 63;
#pragma unroll
  for (int q = 0; q < kHalfPerThread; ++q) {
    const int c = (threadIdx.x >> 6) + q * (kLargeThreads / 64);
    v[q] = (r < nr && c0 + c < nc) ? __ldcg(G + r + (size_t)c * m) : 0.0;
  }
}
__device__ __forceinline__ void store_half(double* S, const double (&v)[kHalfPerThread]) {
  const int r = threadIdx.x & 63;
#pragma unroll
  for (int q = 0; q < kHalfPerThread; ++q) S[r + ((threadIdx.x >> 6) + q * (kLargeThreads / 64)) * kLd] = v[q];
}
// acc += A * B^T over a 32-deep half tile pair
__device__ __forceinline__ void half_gemm(const double* As, const double* Bs, double (&acc)[2][4][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = warp & 3, wc = warp >> 2;
  const int g = lane >> 2, tq = lane & 3;
  const double* ap = As + (wr * 16 + g) + tq * kLd;
  const double* bp = Bs + (wc * 32 + g) + tq * kLd;
#pragma unroll
  for (int kk = 0; kk < kHalf; kk += 4) {
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

// ---- v2 critical path: panel Cholesky / panel substitution without shuffles or explicit inverses -----
// C(8x8 tile (rb, cb) of Cs) -= X[rb rows][c0..c0+8) * Y[cb rows][c0..c0+8)^T   (one warp, two DMMA k-steps)
__device__ __forceinline__ void rank8_tile(double* Cs, const double* Xs, const double* Ys, int rb, int cb, int c0) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  double* cp = Cs + (rb * 8 + g) + (cb * 8 + 2 * tq) * kLd;
  double d0 = cp[0], d1 = cp[kLd];
  const double* xp = Xs + (rb * 8 + g) + (c0 + tq) * kLd;
  const double* yp = Ys + (cb * 8 + g) + (c0 + tq) * kLd;
  mma884(d0, d1, -xp[0], yp[0]);
  mma884(d0, d1, -xp[4 * kLd], yp[4 * kLd]);
  cp[0] = d0;
  cp[kLd] = d1;
}

// 64x64 Cholesky of the tile in shared memory (lower part valid, ld = kLd) in panels of 8 columns.
// Panel step: every row thread (tid < 64, row >= c0) reads the 8x8 diagonal block (broadcast), factors
// it redundantly in registers and solves its own row against it -- no cross-lane traffic inside the
// 8-column chain; eight threads of a third warp do the same factorization and produce the columns of
// the inverse of the 8x8 block (s_binv[p], column-major 8x8, used by the TRSMs as a DMMA operand).
// The rank-8 trailing update (DMMA) is split: the next panel's tile column right away, all other
// tile columns by warps 3..7 while warps 0..2 are in the next panel's column chain.
__device__ void potrf_64_v2(double* As, double* s_binv, int* fail) {
  const int tid = threadIdx.x, warp = tid >> 5;
#pragma unroll 1
  for (int p = 0; p < 8; ++p) {
    const int c0 = 8 * p;
    const bool row_thread = tid < kT && tid >= c0;
    const bool inv_thread = tid >= kT && tid < kT + 8;
    double out[8];  // row thread: its row of L in this panel; inverse thread: its column of the block inverse
    if (row_thread || inv_thread) {
      double D[8][8], ri[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = j; i < 8; ++i) D[i][j] = As[(c0 + i) + (c0 + j) * kLd];
      double a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = row_thread ? As[tid + (c0 + j) * kLd] : 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double d = D[j][j];
        if (!(d > 0.0)) {
          *fail = 1;
          d = __longlong_as_double(0x7ff8000000000000LL);
        }
        ri[j] = rsqrt(d);
#pragma unroll
        for (int i = j + 1; i < 8; ++i) D[i][j] *= ri[j];
#pragma unroll
        for (int jj = j + 1; jj < 8; ++jj)
#pragma unroll
          for (int i = jj; i < 8; ++i) D[i][jj] -= D[i][j] * D[jj][j];
      }
      if (row_thread) {
        const int jr = tid - c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double v = a[j];
#pragma unroll
          for (int c = 0; c < j; ++c) v -= out[c] * D[j][c];
          out[j] = v * ri[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = jr >= j ? out[j] : 0.0;
      } else {
        // column e of the inverse of the 8x8 block: y = L^-1 e_e by forward substitution
        const int e = tid - kT;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          double v = (i == e) ? 1.0 : 0.0;
#pragma unroll
          for (int c = 0; c < i; ++c) v -= D[i][c] * out[c];
          out[i] = (i >= e) ? v * ri[i] : 0.0;
        }
      }
    }
    else if (warp >= 3 && p >= 1 && p <= 6) {
      // meanwhile warps 3..7 finish the trailing update of the PREVIOUS panel (tile columns >= p + 1;
      // tile column p, the one this panel factors, was updated right after the previous panel)
      int t = 0;
      for (int cb = p + 1; cb < 8; ++cb)
        for (int rb = cb; rb < 8; ++rb, ++t)
          if (t % 5 == warp - 3) rank8_tile(As, As, As, rb, cb, c0 - 8);
    }
    // Every thread has read the 8 x 8 diagonal block (rows c0 .. c0 + 7 of the panel) before its owner rows are
    // overwritten with L.  Without this barrier a warp that runs late (rows 32..63, the inverse threads) factors a
    // half-written block: a shared-memory race that stayed invisible while all warps took the same time to get here.
    __syncthreads();
    if (row_thread) {
#pragma unroll
      for (int j = 0; j < 8; ++j) As[tid + (c0 + j) * kLd] = out[j];
    } else if (inv_thread) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s_binv[p * 64 + i + 8 * (tid - kT)] = out[i];
    }
    __syncthreads();
    // rank-8 update of the next panel's tile column only (one tile per warp); the rest overlaps the next panel
    if (p < 7 && p + 1 + warp < 8) rank8_tile(As, As, As, p + 1 + warp, p + 1, c0);
    __syncthreads();
  }
}

// X <- X * L^-T for a 64x64 tile, one warp per 8-row strip held in registers as DMMA accumulator
// fragments xf[cb] = X[8w + g][8cb + 2tq + {0,1}]; rows are independent, so there is no barrier.  Per panel p:
//   X_p <- X_p * inv(L_pp)^T        (2 DMMA; inv(L_pp) from s_binv, X_p re-laid out as an A operand by shuffles)
//   X_cb -= X_p * L[cb, p]^T, cb > p (2 DMMA each, independent)
__device__ __forceinline__ void frag_to_a(const double (&c)[2], double& a0, double& a1) {
  // accumulator layout (row g: cols 2tq, 2tq+1) -> A operand layout (row g: col tq, col 4 + tq)
  const int lane = threadIdx.x & 31;
  const int g4 = lane & ~3, tq = lane & 3;
  const double lo0 = __shfl_sync(0xffffffffu, c[0], g4 + (tq >> 1));
  const double hi0 = __shfl_sync(0xffffffffu, c[1], g4 + (tq >> 1));
  const double lo1 = __shfl_sync(0xffffffffu, c[0], g4 + 2 + (tq >> 1));
  const double hi1 = __shfl_sync(0xffffffffu, c[1], g4 + 2 + (tq >> 1));
  a0 = (tq & 1) ? hi0 : lo0;
  a1 = (tq & 1) ? hi1 : lo1;
}
__device__ __forceinline__ void trsm_frag_64(double (&xf)[8][2], const double* Ls, const double* s_binv) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int c0 = 8 * p;
    double a0, a1;
    frag_to_a(xf[p], a0, a1);
    double d0 = 0.0, d1 = 0.0;
    mma884(d0, d1, a0, s_binv[p * 64 + g + 8 * tq]);
    mma884(d0, d1, a1, s_binv[p * 64 + g + 8 * (4 + tq)]);
    xf[p][0] = d0;
    xf[p][1] = d1;
    if (p < 7) {
      frag_to_a(xf[p], a0, a1);
      a0 = -a0;
      a1 = -a1;
#pragma unroll
      for (int cb = p + 1; cb < 8; ++cb) {
        const double* lp = Ls + (cb * 8 + g) + (c0 + tq) * kLd;
        mma884(xf[cb][0], xf[cb][1], a0, lp[0]);
        mma884(xf[cb][0], xf[cb][1], a1, lp[4 * kLd]);
      }
    }
  }
}
// fragments of warp w's strip <-> the 64 x 64 tile at G (global, ld ldg; rows < nr, cols < nc valid)
__device__ __forceinline__ void load_frag(double (&xf)[8][2], const double* __restrict__ G, int ldg, int nr, int nc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = warp * 8 + (lane >> 2), tq = lane & 3;
#pragma unroll
  for (int cb = 0; cb < 8; ++cb)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = cb * 8 + 2 * tq + e;
      xf[cb][e] = (r < nr && c < nc) ? __ldcg(G + r + (size_t)c * ldg) : 0.0;
    }
}
__device__ __forceinline__ void store_frag(const double (&xf)[8][2], double* G, int ldg, int nr, int nc, double* keep) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = warp * 8 + (lane >> 2), tq = lane & 3;
#pragma unroll
  for (int cb = 0; cb < 8; ++cb)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = cb * 8 + 2 * tq + e;
      if (r < nr && c < nc) G[r + (size_t)c * ldg] = xf[cb][e];
      if (keep != nullptr) keep[r + c * kLd] = xf[cb][e];
    }
}

// smem <- the final L_kk of pivot tile k from the front (strictly upper part zero, identity padding)
__device__ __forceinline__ void load_L(double* S, const double* __restrict__ F, int m, int s0, int nb) {
  const int r = threadIdx.x & 63, c0 = threadIdx.x >> 6;
  constexpr int kPer = kT / (kLargeThreads / 64);
  double v[kPer];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int c = c0 + q * (kLargeThreads / 64);
    if (r < nb && c < nb)
      v[q] = (r >= c) ? __ldcg(F + (s0 + r) + (size_t)(s0 + c) * m) : 0.0;
    else
      v[q] = (r == c) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int q = 0; q < kPer; ++q) S[r + (c0 + q * (kLargeThreads / 64)) * kLd] = v[q];
}

__device__ unsigned long long* g_trace = nullptr;
__device__ unsigned long long g_diag_stamps[8 * 512];  // debug: fine-grained DIAG phases  // debug: [task][4] = claim, deps ready, done (ns), smid
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
void set_factor_trace(unsigned long long* buf) { cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)); }
void get_diag_stamps(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_diag_stamps, sizeof(unsigned long long) * 8 * 512); }

__global__ void __launch_bounds__(kLargeThreads, 2) large_factor_kernel(Ctrl* ctrl, FrontDev fd, LargeDev ld, int t0,
                                                                     int t1, int level, int fwd) {
  extern __shared__ __align__(16) double sm[];
  double* As = sm;
  double* Bs = sm + kT * kLd;
  __shared__ int s_task;
  __shared__ double s_binv[512];  // inverses of the eight 8x8 diagonal blocks of the current L_kk
  if (ctrl->done) return;
  const int tid = threadIdx.x;
  // the trace pointer is read once: a __device__ variable is a global load, and the acquire polls of this kernel keep
  // invalidating L1, so `if (g_trace ...)` at every stamp cost three or four L2 round trips per task
  unsigned long long* const trace = g_trace;
  int cur_lf = -1;
  LargeFront lf_cache{};
  for (;;) {
    if (tid == 0) s_task = t0 + atomicAdd(&ld.queue[level], 1);
    __syncthreads();
    const int t = s_task;
    __syncthreads();
    if (t >= t1) break;
    const LargeTask task = ld.tasks[t];
    if (task.lf != cur_lf) {  // consecutive tasks of a CTA often belong to the same front
      lf_cache = ld.lf[task.lf];
      cur_lf = task.lf;
    }
    const LargeFront& lf = lf_cache;
    double* F = fd.fronts + lf.off;
    const int m = lf.m, nt = lf.nt;
    int* cnt = ld.counters + lf.cnt_off;
    const int k = task.k, i = task.i, j = task.j;
    if (trace && tid == 0) trace[(size_t)t * 4 + 0] = gtime();
    if (task.type == 3) {
      // ---------------- DIAG(k): POTRF(k), then TRSM(k+1,k) and UPDATE(k+1,k+1,k) on the critical path.
      // Sticky fronts (fused schedule): this CTA keeps the chain -- after UPDATE(k+1,k+1,k) the tile stays in shared
      // memory and the loop goes on with POTRF(k+1): no store + fence + publish + acquire + reload between two steps.
      __shared__ int s_pre;
      const bool sticky = lf.sticky != 0;
      bool handed = false;  // As already holds tile (kd,kd) with every update applied
      for (int kd = k;; ++kd) {
        double* binv_g = ld.binv + lf.linv_off / 8 + (size_t)kd * 512;
        // fused schedule: the children's update matrices are added by EXTEND-ADD tasks of this launch; everything
        // else in the front depends on POTRF(0), so this is the only wait on the assembly
        if (kd == 0 && lf.n_ea > 0) wait_ge(ld.counters + lf.asm_off, lf.n_ea);
        // tile (kd+1, kd) is usually ready before the diagonal tile: fetch it now (into Bs) if so
        if (tid == 0) s_pre = (kd + 1 < nt) && ld_acquire(cnt + (kd + 1) * nt + kd) == kd;
        if (!handed)
          wait_eq(cnt + kd * nt + kd, kd);
        else
          __syncthreads();
        if (trace && tid == 0 && kd == k) trace[(size_t)t * 4 + 1] = gtime();
        const int s0 = tile_start(lf, kd), nb = tile_size(lf, kd);
        const bool pre = s_pre != 0;
        if (!handed) load_tile(As, F + s0 + (size_t)s0 * m, m, nb, nb, true);
        if (pre) load_tile(Bs, F + tile_start(lf, kd + 1) + (size_t)s0 * m, m, tile_size(lf, kd + 1), nb, false);
        __syncthreads();
        if (trace && tid == 0 && kd < 512) g_diag_stamps[kd * 8 + 0] = gtime();
        potrf_64_v2(As, s_binv, &ctrl->chol_fail);
        __syncthreads();
        // first non-positive pivot tile of the factorization: (large front + 1) << 16 | pivot tile, for diagnostics
        if (tid == 0 && ctrl->chol_fail && ctrl->fail_where == 0) {
          bool nanp = false;
          for (int q = 0; q < nb; ++q) nanp = nanp || !(As[q + q * kLd] > 0.0);
          if (nanp) atomicCAS(&ctrl->fail_where, 0, ((task.lf + 1) << 16) | kd);
        }
        if (trace && tid == 0 && kd < 512) g_diag_stamps[kd * 8 + 1] = gtime();
        {
          const int r = tid & 63;
          for (int c = tid >> 6; c < kT; c += kLargeThreads / 64)
            if (r < nb && c < nb && r >= c) F[(s0 + r) + (size_t)(s0 + c) * m] = As[r + c * kLd];
          binv_g[tid] = s_binv[tid];
          binv_g[tid + 256] = s_binv[tid + 256];
        }
        publish(cnt + kd * nt + kd, kd + 1);
        if (trace && tid == 0 && kd < 512) g_diag_stamps[kd * 8 + 2] = gtime();
        if (trace && tid == 0 && kd == k) trace[(size_t)t * 4 + 3] = gtime();  // POTRF published
        handed = false;
        if (kd + 1 < nt) {
          // TRSM(kd+1, kd) against L_kk still in shared memory; result to the front and to Bs for the SYRK
          const int ri = tile_start(lf, kd + 1), ni = tile_size(lf, kd + 1);
          double xf[8][2];
          if (pre) {
            const int lane = tid & 31, warp = tid >> 5;
            const int r = warp * 8 + (lane >> 2), tq = lane & 3;
#pragma unroll
            for (int cb = 0; cb < 8; ++cb)
#pragma unroll
              for (int e = 0; e < 2; ++e) xf[cb][e] = Bs[r + (cb * 8 + 2 * tq + e) * kLd];
          } else {
            wait_eq(cnt + (kd + 1) * nt + kd, kd);
            load_frag(xf, F + ri + (size_t)s0 * m, m, ni, nb);
          }
          trsm_frag_64(xf, As, s_binv);
          store_frag(xf, F + ri + (size_t)s0 * m, m, ni, nb, Bs);
          publish(cnt + (kd + 1) * nt + kd, kd + 1);
          if (trace && tid == 0 && kd < 512) g_diag_stamps[kd * 8 + 3] = gtime();
          // UPDATE(kd+1, kd+1, kd)
          wait_eq(cnt + (kd + 1) * nt + (kd + 1), kd);
          if (sticky && kd + 1 < lf.wt) {
            // the updated tile goes straight into As for the next POTRF (identity padding restored); the copy in the
            // front is written too but nobody waits for it: its next reader is this CTA
            gemm_store(F, m, lf, kd + 1, kd + 1, Bs, Bs, false, As);
            __syncthreads();
            if (tid >= ni && tid < kT) As[tid + tid * kLd] = 1.0;
            handed = true;
          } else {
            gemm_store(F, m, lf, kd + 1, kd + 1, Bs, Bs, false, nullptr);
            publish(cnt + (kd + 1) * nt + (kd + 1), kd + 1);
          }
        }
        if (!handed) break;
      }
      if (trace && tid == 0) trace[(size_t)t * 4 + 2] = gtime();
    } else if (task.type == 5) {
      // ---------------- INV(k): L_kk^-1 for the triangular solves (off the critical path) ----------------
      wait_ge(cnt + k * nt + k, k + 1);
      if (trace && tid == 0) trace[(size_t)t * 4 + 1] = gtime();
      const int s0 = tile_start(lf, k), nb = tile_size(lf, k);
      load_L(As, F, m, s0, nb);
      __syncthreads();
      tri_inverse_64(As, Bs);
      double* linv = ld.linv + lf.linv_off + (size_t)k * kT * kT;
      {
        const int r = tid & 63;
        for (int c = tid >> 6; c < kT; c += kLargeThreads / 64) linv[r + c * kT] = Bs[r + c * kLd];
      }
      if (fwd)
        publish(ld.counters + lf.vc_off + nt + lf.wt + k, 1);  // L_kk^-1 ready for the fused forward substitution
      else
        __syncthreads();
      if (trace && tid == 0) trace[(size_t)t * 4 + 2] = gtime();
    } else if (task.type == 6) {
      // ---------------- EXTEND-ADD: update tile (i,j) of this (final) front -> its parent front ----------------
      const int wt = lf.wt;
      wait_ge(cnt + i * nt + j, wt);
      if (trace && tid == 0) trace[(size_t)t * 4 + 1] = gtime();
      const LargeFront pf = ld.lf[lf.parent_lf];
      double* Fp = fd.fronts + pf.off;
      const int mp = pf.m;
      const int32_t* rel = fd.f_rel + fd.f_rows_ptr[lf.front] - lf.w;  // indexed by the row of this front
      const int ri = tile_start(lf, i), ni = tile_size(lf, i);
      const int cj = tile_start(lf, j), nj = tile_size(lf, j);
      const int r = tid & 63, c0 = tid >> 6;
      if (r < ni) {
        const int pr = __ldg(rel + ri + r);
        constexpr int kPer = kT / (kLargeThreads / 64);
        double v[kPer];
        int pc[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q) {  // all loads in flight before the first RED
          const int c = c0 + q * (kLargeThreads / 64);
          const bool in = c < nj && ri + r >= cj + c;  // lower part (rel is increasing: it lands in the parent's lower part)
          v[q] = in ? __ldcg(F + (ri + r) + (size_t)(cj + c) * m) : 0.0;
          pc[q] = in ? __ldg(rel + cj + c) : -1;
        }
#pragma unroll
        for (int q = 0; q < kPer; ++q)
          if (pc[q] >= 0) atomicAdd(Fp + pr + (size_t)pc[q] * mp, v[q]);
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(ld.counters + pf.asm_off, 1);
      }
      if (trace && tid == 0) trace[(size_t)t * 4 + 2] = gtime();
    } else if (task.type >= 8) {
      // ---------------- fused forward substitution L y = b (fd.ywork gets y, pivot rows in elimination order) -----
      // 8: y_k = L_kk^-1 b_k     9: b_i -= L(i,k) y_k     10: update rows of b -> the parent's b
      if (!fwd) continue;  // a factorization without a right-hand side (covariances)
      int* vc = ld.counters + lf.vc_off;  // [nt] updates applied to row tile | [wt] y published | [wt] L_kk^-1 ready | [1]
      double* b = ld.fwd_b + lf.fb_off;
      const int wt = lf.wt;
      double* red = sm;          // 256 partial sums
      double* xs = sm + 256;     // 64 operand entries
      if (task.type == 8) {
        if (tid == 0) {
          if (k == 0)
            while (ld_acquire(vc + nt + 2 * wt) < lf.n_vch) __nanosleep(32);
          while (ld_acquire(vc + k) < k) __nanosleep(32);
          while (ld_acquire(vc + nt + wt + k) < 1) __nanosleep(32);
        }
        __syncthreads();
        if (trace && tid == 0) trace[(size_t)t * 4 + 1] = gtime();
        const int s0 = tile_start(lf, k), nb = tile_size(lf, k);
        if (tid < kT) xs[tid] = tid < nb ? __ldcg(b + s0 + tid) : 0.0;
        const double* linv = ld.linv + lf.linv_off + (size_t)k * kT * kT;
        const int r = tid & 63, gq = tid >> 6;
        double lv[kT / 4];
#pragma unroll
        for (int q = 0; q < kT / 4; ++q) lv[q] = __ldcg(linv + r + (size_t)(gq + 4 * q) * kT);
        __syncthreads();
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < kT / 4; ++q) v += lv[q] * xs[gq + 4 * q];
        red[gq * 64 + r] = v;
        __syncthreads();
        if (tid < nb) fd.ywork[fd.f_piv[lf.front] + s0 + tid] = red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid];
        publish(vc + nt + k, 1);
      } else if (task.type == 9) {
        if (tid == 0) {
          while (ld_acquire(vc + nt + k) < 1) __nanosleep(32);
          while (ld_acquire(cnt + i * nt + k) < k + 1) __nanosleep(32);
        }
        __syncthreads();
        if (trace && tid == 0) trace[(size_t)t * 4 + 1] = gtime();
        const int s0 = tile_start(lf, k), nk = tile_size(lf, k);
        const int ri = tile_start(lf, i), ni = tile_size(lf, i);
        if (tid < kT) xs[tid] = tid < nk ? __ldcg(fd.ywork + fd.f_piv[lf.front] + s0 + tid) : 0.0;
        const double* Lt = F + ri + (size_t)s0 * m;
        const int r = tid & 63, gq = tid >> 6;
        double lv[kT / 4];
#pragma unroll
        for (int q = 0; q < kT / 4; ++q) {
          const int c = gq + 4 * q;
          lv[q] = (r < ni && c < nk) ? __ldcg(Lt + r + (size_t)c * m) : 0.0;
        }
        __syncthreads();
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < kT / 4; ++q) v += lv[q] * xs[gq + 4 * q];
        red[gq * 64 + r] = v;
        __syncthreads();
        if (tid < ni) atomicAdd(b + ri + tid, -(red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid]));
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          atomicAdd(vc + i, 1);
        }
      } else {
        // every update row tile has received its wt contributions
        if (tid < 32) {
          for (int q = wt + tid; q < nt; q += 32)
            while (ld_acquire(vc + q) < wt) __nanosleep(32);
        }
        __syncthreads();
        if (trace && tid == 0) trace[(size_t)t * 4 + 1] = gtime();
        const LargeFront pf = ld.lf[lf.parent_lf];
        double* pb = ld.fwd_b + pf.fb_off;
        const int32_t* rel = fd.f_rel + fd.f_rows_ptr[lf.front];
        for (int q = tid; q < m - lf.w; q += kLargeThreads) atomicAdd(pb + __ldg(rel + q), __ldcg(b + lf.w + q));
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          atomicAdd(ld.counters + pf.vc_off + pf.nt + 2 * pf.wt, 1);
        }
      }
      if (trace && tid == 0) trace[(size_t)t * 4 + 2] = gtime();
    } else if (task.type == 1) {
      // ---------------- TRSM(i,k): strip-per-warp substitution in registers ----------------
      if (tid == 0) {
        while (ld_acquire(cnt + k * nt + k) < k + 1) __nanosleep(32);
        while (ld_acquire(cnt + i * nt + k) != k) __nanosleep(32);
      }
      __syncthreads();
          if (trace && tid == 0) trace[(size_t)t * 4 + 1] = gtime();
      const int s0 = tile_start(lf, k), nb = tile_size(lf, k);
      const int ri = tile_start(lf, i), ni = tile_size(lf, i);
      double xf[8][2];
      load_frag(xf, F + ri + (size_t)s0 * m, m, ni, nb);
      load_L(As, F, m, s0, nb);
      {
        const double* binv_g = ld.binv + lf.linv_off / 8 + (size_t)k * 512;
        s_binv[tid] = __ldcg(binv_g + tid);
        s_binv[tid + 256] = __ldcg(binv_g + tid + 256);
      }
      __syncthreads();
      trsm_frag_64(xf, As, s_binv);
      store_frag(xf, F + ri + (size_t)s0 * m, m, ni, nb, nullptr);
      publish(cnt + i * nt + k, k + 1);
      if (trace && tid == 0) trace[(size_t)t * 4 + 2] = gtime();
    } else if (task.type == 4) {
      // ---------------- UPDATE(i, j, [k, k1)) ----------------
      const int k1 = task.k1;
      const int nh = 2 * (k1 - k);
      double acc[2][4][2];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) acc[rb][cb][0] = acc[rb][cb][1] = 0.0;
      // one warp polls the version counters of every operand tile at once: when all of them are final already (the
      // usual case) the loop below runs without the per-step poll + block barrier
      __shared__ int s_all_ready;
      if (tid < 32) {
        bool ok = true;
        if (tid < k1 - k) ok = ld_acquire(cnt + i * nt + k + tid) >= k + tid + 1 && ld_acquire(cnt + j * nt + k + tid) >= k + tid + 1;
        ok = __all_sync(0xffffffffu, ok);
        if (tid == 0) s_all_ready = ok && (k1 - k) <= 32;
      }
      __syncthreads();
      const bool all_ready = s_all_ready != 0;
      // register-staged double buffer (ld.global.cg -> registers -> st.shared): measured faster than a cp.async ring
      // (8- or 16-byte copies, three stages) whose 104 KB of shared memory leaves the SM almost no L1
      double ra[kHalfPerThread], rb_[kHalfPerThread];
      auto fetch = [&](int h) {
        const int kk = k + (h >> 1);
        if ((h & 1) == 0 && !all_ready) {
          if (tid == 0) {
            while (ld_acquire(cnt + i * nt + kk) < kk + 1) __nanosleep(32);
            while (ld_acquire(cnt + j * nt + kk) < kk + 1) __nanosleep(32);
          }
          __syncthreads();
        }
        load_half(ra, F, m, lf, i, kk, (h & 1) * kHalf);
        load_half(rb_, F, m, lf, j, kk, (h & 1) * kHalf);
      };
      fetch(0);
      if (trace && tid == 0) trace[(size_t)t * 4 + 1] = gtime();
      store_half(sm, ra);
      store_half(sm + kHalfDoubles, rb_);
      __syncthreads();
      for (int h = 0; h < nh; ++h) {
        double* cur = sm + (h & 1) * 2 * kHalfDoubles;
        double* nxt = sm + ((h + 1) & 1) * 2 * kHalfDoubles;
        if (h + 1 < nh) fetch(h + 1);
        half_gemm(cur, cur + kHalfDoubles, acc);
        if (h + 1 < nh) {
          store_half(nxt, ra);
          store_half(nxt + kHalfDoubles, rb_);
        }
        __syncthreads();
      }
      // C tile: all earlier updates applied
      wait_eq(cnt + i * nt + j, k);
      {
        const int lane = tid & 31, warp = tid >> 5;
        const int wr = warp & 3, wc = warp >> 2;
        const int g = lane >> 2, tq = lane & 3;
        const int ri = tile_start(lf, i), ni = tile_size(lf, i);
        const int cj = tile_start(lf, j), nj = tile_size(lf, j);
        double* C = F + ri + (size_t)cj * m;
        // all 16 reads of the target in flight before the first store (they may alias for the compiler)
#pragma unroll
        for (int rb = 0; rb < 2; ++rb)
#pragma unroll
          for (int cb = 0; cb < 4; ++cb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int r = wr * 16 + rb * 8 + g;
              const int c = wc * 32 + cb * 8 + tq * 2 + e;
              const double cin = (r < ni && c < nj) ? __ldcg(C + r + (size_t)c * m) : 0.0;
              acc[rb][cb][e] = cin - acc[rb][cb][e];
            }
#pragma unroll
        for (int rb = 0; rb < 2; ++rb)
#pragma unroll
          for (int cb = 0; cb < 4; ++cb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int r = wr * 16 + rb * 8 + g;
              const int c = wc * 32 + cb * 8 + tq * 2 + e;
              if (r < ni && c < nj) C[r + (size_t)c * m] = acc[rb][cb][e];
            }
      }
      publish(cnt + i * nt + j, k1);
      if (trace && tid == 0) trace[(size_t)t * 4 + 2] = gtime();
    } else {
      // ---------------- UPDATE(i,j,k) ----------------
      const bool trsm = false;
      if (tid == 0) {
        if (trsm) {
          while (ld_acquire(cnt + k * nt + k) < k + 1) __nanosleep(32);
          while (ld_acquire(cnt + i * nt + k) != k) __nanosleep(32);
        } else {
          while (ld_acquire(cnt + i * nt + k) < k + 1) __nanosleep(32);
          while (ld_acquire(cnt + j * nt + k) < k + 1) __nanosleep(32);
          while (ld_acquire(cnt + i * nt + j) != k) __nanosleep(32);
        }
      }
      __syncthreads();
          if (trace && tid == 0) trace[(size_t)t * 4 + 1] = gtime();
      const int ck = tile_start(lf, k), nk = tile_size(lf, k);
      load_tile(As, F + tile_start(lf, i) + (size_t)ck * m, m, tile_size(lf, i), nk, false);
      if (trsm)
        load_tile(Bs, ld.linv + lf.linv_off + (size_t)k * kT * kT, kT, kT, kT, false);
      else
        load_tile(Bs, F + tile_start(lf, j) + (size_t)ck * m, m, tile_size(lf, j), nk, false);
      __syncthreads();
      gemm_store(F, m, lf, i, trsm ? k : j, As, Bs, trsm, nullptr);
      publish(cnt + i * nt + (trsm ? k : j), k + 1);
      if (trace && tid == 0) trace[(size_t)t * 4 + 2] = gtime();
    }
  }
}

// ---- assembly of large fronts ---------------------------------------------------------------------
// zeroes the lower-triangular tiles of every large front (column c from the top of its diagonal tile:
// rows >= c - 63 covers it wherever the tile boundary lies); nothing reads above the diagonal tiles
__global__ void __launch_bounds__(256) large_zero_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld, int n_jobs) {
  if (ctrl->done) return;
  // one job = a run of columns of one front with ~8k entries below the top of their diagonal tiles
  for (int jb = blockIdx.x; jb < n_jobs; jb += gridDim.x) {
    const int4 job = ld.zero_jobs[jb];  // {large front, first column, end column, -}
    const LargeFront lf = ld.lf[job.x];
    const int m = lf.m;
    double* F = fd.fronts + lf.off;
    for (int c = job.y; c < job.z; ++c) {
      double* col = F + (size_t)c * m;
      for (int r = max(0, c - 63) + threadIdx.x; r < m; r += 256) col[r] = 0.0;
    }
  }
}
void launch_large_zero(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld, int n_jobs) {
  if (n_jobs == 0) return;
  const int grid = n_jobs < 148 * 8 ? n_jobs : 148 * 8;
  large_zero_kernel<<<grid, 256, 0, st>>>(ctrl, fd, ld, n_jobs); ++g_launches;
}

// one warp per job: system-matrix block copy, damping, or a column range of a child's update matrix
__global__ void __launch_bounds__(256) large_assemble_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld,
                                                              const double* __restrict__ sys_static, StatePtrs sp,
                                                              int use_state_H, const double* __restrict__ dvec, int j0,
                                                              int j1, int plain) {
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
      if (plain)
        *dst = v;
      else
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
#pragma unroll 4
      for (int ic = jc + lane; ic < uc; ic += 32) atomicAdd(F + rel[ic] + (size_t)dj * m, U[ic + (size_t)jc * mc]);
    }
  }
}

void launch_large_preassemble(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld,
                              const double* sys_static, StatePtrs sp, int use_state_H, const double* dvec, int pre_j0,
                              int pre_j1, int damp_j0, int damp_j1) {
  if (pre_j1 > pre_j0) {
    large_assemble_kernel<<<(pre_j1 - pre_j0 + 7) / 8, 256, 0, st>>>(ctrl, fd, ld, sys_static, sp, use_state_H, dvec,
                                                                    pre_j0, pre_j1, 1);
    ++g_launches;
  }
  if (dvec != nullptr && damp_j1 > damp_j0) {
    large_assemble_kernel<<<(damp_j1 - damp_j0 + 7) / 8, 256, 0, st>>>(ctrl, fd, ld, sys_static, sp, use_state_H, dvec,
                                                                      damp_j0, damp_j1, 0);
    ++g_launches;
  }
}
void launch_large_level(cudaStream_t st, Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld, const LargeLevel& lv,
                        int level, const double* sys_static, StatePtrs sp, int use_state_H, const double* dvec,
                        int grid_cap) {
  if (lv.n_lf == 0) return;
  const int nj = lv.j1 - lv.j0;
  if (nj > 0) {
    large_assemble_kernel<<<(nj + 7) / 8, 256, 0, st>>>(ctrl, fd, ld, sys_static, sp, use_state_H, dvec, lv.j0, lv.j1, 0);
    ++g_launches;
  }
  const int ntask = lv.t1 - lv.t0;
  const int cap = large_factor_resident_ctas();
  int grid = ntask < cap ? ntask : cap;
  // single-front levels are bound by the diagonal chain, not by throughput: one CTA per SM is as fast as two
  // (measured) and leaves room for the forward-substitution CTAs that overlap them
  if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
  large_factor_kernel<<<grid, kLargeThreads, kFactorSmem, st>>>(ctrl, fd, ld, lv.t0, lv.t1, level, 0); ++g_launches;
}


// ---- triangular solves on large fronts -------------------------------------------------------------
// P cooperating CTAs per front; tile-row i of the front is owned by CTA i % P.  The dependency
// chain over the pivot tiles runs through global flags (forward) / per-tile contribution counters
// (backward) with ld.acquire / st.release; every L tile is read exactly once, spread over P SMs.
// All CTAs of a launch are co-resident (grid <= #SMs), which the spin-waits rely on.
__device__ __forceinline__ void tile_gemv(const double* __restrict__ Lt, int ldm, int nr, int nc,
                                          const double* __restrict__ x, double* red, double* out_sub) {
  // out_sub[r] -= sum_c Lt[r + c*ldm] * x[c]; 256 threads: r = tid & 63, 4 column groups
  const int r = threadIdx.x & 63, gq = threadIdx.x >> 6;
  double v = 0.0;
  if (r < nr)
    for (int c = gq; c < nc; c += 4) v += __ldcg(Lt + r + (size_t)c * ldm) * x[c];
  red[gq * 64 + r] = v;
  __syncthreads();
  if (threadIdx.x < 64 && r < nr) out_sub[r] -= red[r] + red[64 + r] + red[128 + r] + red[192 + r];
  __syncthreads();
}
__device__ __forceinline__ void tile_gemv_t(const double* __restrict__ Lt, int ldm, int nr, int nc,
                                            const double* __restrict__ x, double* __restrict__ out) {
  // out[c] = sum_r Lt[r + c*ldm] * x[r]; one warp per column (8 warps)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < nc; c += 8) {
    double v = 0.0;
    for (int r = lane; r < nr; r += 32) v += __ldcg(Lt + r + (size_t)c * ldm) * x[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[c] = v;
  }
}

__global__ void __launch_bounds__(256) large_solve_fwd_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld,
                                                               const double* __restrict__ rhs_static, StatePtrs sp,
                                                               int use_state_rhs, int lf0, int P) {
  extern __shared__ double smf[];  // own: ceil(nt/P)*64 | y: 64 | red: 256
  if (ctrl->done) return;
  const LargeFront lf = ld.lf[lf0 + blockIdx.x / P];
  const int p = blockIdx.x % P;
  const int s = lf.front, w = lf.w, m = lf.m, nt = lf.nt, wt = lf.wt;
  const int n_own = (nt - p + P - 1) / P;
  double* own = smf;
  double* yk = own + (size_t)((nt + P - 1) / P) * kT;
  double* red = yk + kT;
  const double* rhs = use_state_rhs ? sp.rhs[ctrl->init_idx] : rhs_static;
  const double* L = fd.fronts + lf.off;
  int* flag = ld.sflags + lf.flag_off;
  const int tid = threadIdx.x;
  // ---- assemble the owned part of the front right-hand side
  for (int q = tid; q < n_own * kT; q += 256) {
    const int i = p + (q / kT) * P, r = q % kT;
    const int row = tile_start(lf, i) + r;
    double v = 0.0;
    if (r < tile_size(lf, i) && row < w) v = rhs[fd.scalar_perm[fd.f_piv[s] + row]];
    own[q] = v;
  }
  __syncthreads();
  for (int ci = fd.f_child_ptr[s]; ci < fd.f_child_ptr[s + 1]; ++ci) {
    const int c = fd.f_child[ci];
    const int uc = fd.f_u[c];
    const double* t = fd.twork + fd.f_toff[c];
    const int32_t* rel = fd.f_rel + fd.f_rows_ptr[c];
    for (int q = tid; q < uc; q += 256) {
      const int row = rel[q];
      const int i = row < w ? row / kT : wt + (row - w) / kT;
      if (i % P == p) own[(i / P) * kT + (row - tile_start(lf, i))] += t[q];  // rows of one child are distinct
    }
    __syncthreads();
  }
  // ---- forward substitution over the pivot tiles
  for (int k = 0; k < wt; ++k) {
    const int nb = tile_size(lf, k);
    double* ypub = fd.ywork + fd.f_piv[s] + k * kT;
    if (k % P == p) {
      const double* linv = ld.linv + lf.linv_off + (size_t)k * kT * kT;
      const double* fk = own + (k / P) * kT;
      if (tid < nb) {
        double v = 0.0;
        for (int q = 0; q <= tid; ++q) v += linv[tid + q * kT] * fk[q];
        yk[tid] = v;
        ypub[tid] = v;
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(flag + k, 1);
      }
    } else {
      if (tid == 0)
        while (ld_acquire(flag + k) == 0) __nanosleep(20);
      __syncthreads();
      if (tid < nb) yk[tid] = __ldcg(ypub + tid);
      __syncthreads();
    }
    const int ck = k * kT;
    for (int i = k + 1 + ((p - (k + 1)) % P + P) % P; i < nt; i += P)
      tile_gemv(L + tile_start(lf, i) + (size_t)ck * m, m, tile_size(lf, i), nb, yk, red, own + (i / P) * kT);
  }
  // ---- update part -> twork (pivot part was published tile by tile)
  for (int q = tid; q < n_own * kT; q += 256) {
    const int i = p + (q / kT) * P, r = q % kT;
    if (i >= wt && r < tile_size(lf, i)) fd.twork[fd.f_toff[s] + (tile_start(lf, i) - w) + r] = own[q];
  }
}

__global__ void __launch_bounds__(256) large_solve_bwd_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld,
                                                               int lf0, int P) {
  extern __shared__ double smf[];  // x: 64 | g: 64 | tmp: 64
  if (ctrl->done) return;
  const LargeFront lf = ld.lf[lf0 + blockIdx.x / P];
  const int p = blockIdx.x % P;
  const int s = lf.front, w = lf.w, m = lf.m, nt = lf.nt, wt = lf.wt;
  double* xi = smf;
  double* g = xi + kT;
  const double* L = fd.fronts + lf.off;
  int* cntb = ld.sflags + lf.flag_off + wt;  // contributions received per pivot tile
  double* contrib = ld.contrib + lf.contrib_off;
  const int tid = threadIdx.x;
  const int32_t* rows = fd.f_rows + fd.f_rows_ptr[s];
  // owned tile rows, from the bottom up
  int i = nt - 1 - (((nt - 1 - p) % P) + P) % P;
  for (; i >= 0; i -= P) {
    const int ni = tile_size(lf, i), ri = tile_start(lf, i);
    if (i >= wt) {
      // update rows: x known from the ancestors
      if (tid < ni) xi[tid] = fd.ywork[rows[ri - w + tid]];
      __syncthreads();
    } else {
      // pivot tile i: wait for all contributions of rows below, then x_i = L_ii^-T (y_i - sum)
      if (tid == 0)
        while (ld_acquire(cntb + i) < nt - 1 - i) __nanosleep(20);
      __syncthreads();
      if (tid < ni) {
        double v = fd.ywork[fd.f_piv[s] + ri + tid];
        for (int r = nt - 1; r > i; --r) v -= __ldcg(contrib + ((size_t)i * nt + r) * kT + tid);
        g[tid] = v;
      }
      __syncthreads();
      const double* linv = ld.linv + lf.linv_off + (size_t)i * kT * kT;
      if (tid < ni) {
        double v = 0.0;
        for (int r = tid; r < ni; ++r) v += linv[r + tid * kT] * g[r];
        xi[tid] = v;
        fd.ywork[fd.f_piv[s] + ri + tid] = v;
      }
      __syncthreads();
    }
    // contributions of row i to the pivot tiles k < min(i, wt), nearest first
    for (int k = min(i, wt) - 1; k >= 0; --k) {
      double* out = contrib + ((size_t)k * nt + i) * kT;
      tile_gemv_t(L + ri + (size_t)(k * kT) * m, m, ni, tile_size(lf, k), xi, out);
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(cntb + k, 1);
      }
    }
  }
}


// ---- v2 solves: every L / L_kk^-1 tile a CTA will touch is known up front (the sequence depends only
// on the front geometry), so tiles stream into a 4-stage shared-memory ring with cp.async while the
// CTA waits on the dependency flags; the chain step is then flag latency + one smem mat-vec.
constexpr int kSStages = 4;
constexpr int kSLd = 65;  // conflict-free for both T*x (lanes over rows) and T^T*x (lanes over columns)
__device__ __forceinline__ void cp_async8(double* smem, const double* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void issue_tile(double* T, const double* __restrict__ G, int ldg, int nr, int nc) {
  const int r = threadIdx.x & 63;
  if (r < nr)
    for (int c = threadIdx.x >> 6; c < nc; c += kLargeThreads / 64) cp_async8(T + r + c * kSLd, G + r + (size_t)c * ldg);
}
// "LL" slots for values that cross CTAs inside one solve launch: a double travels as {lo, epoch, hi, epoch}
// in one 16-byte store; each 8-byte half validates itself, so the reader needs no separate flag
// round trip and the writer no fence.  epoch != 0 changes with every solve (slots start zeroed).
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned epoch) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(epoch),
               "r"((unsigned)(b >> 32)), "r"(epoch)
               : "memory");
}
__device__ __forceinline__ double ll_load(const uint4* p, unsigned epoch) {
  unsigned lo, f1, hi, f2;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(p) : "memory");
    if (f1 == epoch && f2 == epoch) break;
    __nanosleep(20);
  }
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
struct SolveItem {
  int kind, i, k;  // kind 0: L_kk^-1 tile of pivot tile k (i == k); 1: L tile (i, k)
};
// forward order of CTA p: for k = 0..wt-1: [L_kk^-1 if k % P == p], then the owned tile rows i > k
struct FwdIter {
  int k, i, phase;
  __device__ void init() { k = 0; i = 0; phase = 0; }
  __device__ bool next(const LargeFront& lf, int p, int P, SolveItem& it) {
    for (;;) {
      if (k >= lf.wt) return false;
      if (phase == 0) {
        phase = 1;
        i = k + 1 + (((p - (k + 1)) % P) + P) % P;
        if (k % P == p) {
          it = SolveItem{0, k, k};
          return true;
        }
      }
      if (i < lf.nt) {
        it = SolveItem{1, i, k};
        i += P;
        return true;
      }
      ++k;
      phase = 0;
    }
  }
};
// backward order of CTA p: owned tile rows from the bottom up: [L_ii^-1 if i < wt], then tiles (i, k), k descending
struct BwdIter {
  int i, k, phase;
  __device__ void init(const LargeFront& lf, int p, int P) {
    i = lf.nt - 1 - (((lf.nt - 1 - p) % P) + P) % P;
    k = 0;
    phase = 0;
  }
  __device__ bool next(const LargeFront& lf, int P, SolveItem& it) {
    for (;;) {
      if (i < 0) return false;
      if (phase == 0) {
        phase = 1;
        k = min(i, lf.wt) - 1;
        if (i < lf.wt) {
          it = SolveItem{0, i, i};
          return true;
        }
      }
      if (k >= 0) {
        it = SolveItem{1, i, k};
        --k;
        return true;
      }
      i -= P;
      phase = 0;
    }
  }
};
__device__ __forceinline__ void issue_item(double* T, const SolveItem& it, const LargeFront& lf, const double* L,
                                           const double* linv) {
  if (it.kind == 0)
    issue_tile(T, linv + (size_t)it.k * kT * kT, kT, kT, kT);
  else
    issue_tile(T, L + tile_start(lf, it.i) + (size_t)(it.k * kT) * lf.m, lf.m, tile_size(lf, it.i), tile_size(lf, it.k));
}
// red[gq*64 + r] = sum_{c = gq, gq+4, ..} T[r][c] * x[c]       (r < nr, c < nc)
__device__ __forceinline__ void smem_gemv_part(const double* T, int nr, int nc, const double* x, double* red) {
  const int r = threadIdx.x & 63, gq = threadIdx.x >> 6;
  double v = 0.0;
  if (r < nr) {
#pragma unroll 4
    for (int c = gq; c < nc; c += 4) v += T[r + c * kSLd] * x[c];
  }
  red[gq * 64 + r] = v;
}
// red[gq*64 + c] = sum_{r = gq, gq+4, ..} T[r][c] * x[r]
__device__ __forceinline__ void smem_gemv_t_part(const double* T, int nr, int nc, const double* x, double* red) {
  const int c = threadIdx.x & 63, gq = threadIdx.x >> 6;
  double v = 0.0;
  if (c < nc) {
#pragma unroll 4
    for (int r = gq; r < nr; r += 4) v += T[r + c * kSLd] * x[r];
  }
  red[gq * 64 + c] = v;
}

__global__ void __launch_bounds__(kLargeThreads, 1)
    large_solve_fwd2_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld, const double* __restrict__ rhs_static,
                            StatePtrs sp, int use_state_rhs, int lf0, int P, unsigned epoch) {
  extern __shared__ double smf[];  // stages | own: ceil(nt/P)*64 | y: 64 | red: 256
  if (ctrl->done) return;
  const LargeFront lf = ld.lf[lf0 + blockIdx.x / P];
  const int p = blockIdx.x % P;
  const int s = lf.front, w = lf.w, m = lf.m, nt = lf.nt, wt = lf.wt;
  const int n_own = (nt - p + P - 1) / P;
  double* stages = smf;
  double* own = stages + kSStages * kT * kSLd;
  double* yk = own + (size_t)((nt + P - 1) / P) * kT;
  double* red = yk + kT;
  const double* rhs = use_state_rhs ? sp.rhs[ctrl->init_idx] : rhs_static;
  const double* L = fd.fronts + lf.off;
  const double* linv = ld.linv + lf.linv_off;
  const int tid = threadIdx.x;
  // ---- start streaming the tiles of this CTA's sequence
  FwdIter prod, cons;
  prod.init();
  cons.init();
  SolveItem it;
  for (int st = 0; st < kSStages - 1; ++st) {
    if (prod.next(lf, p, P, it)) issue_item(stages + st * kT * kSLd, it, lf, L, linv);
    cp_async_commit();
  }
  // ---- assemble the owned part of the front right-hand side
  for (int q = tid; q < n_own * kT; q += kLargeThreads) {
    const int i = p + (q / kT) * P, r = q % kT;
    const int row = tile_start(lf, i) + r;
    double v = 0.0;
    if (r < tile_size(lf, i) && row < w) v = rhs[fd.scalar_perm[fd.f_piv[s] + row]];
    own[q] = v;
  }
  __syncthreads();
  for (int ci = fd.f_child_ptr[s]; ci < fd.f_child_ptr[s + 1]; ++ci) {
    const int c = fd.f_child[ci];
    const int uc = fd.f_u[c];
    const double* t = fd.twork + fd.f_toff[c];
    const int32_t* rel = fd.f_rel + fd.f_rows_ptr[c];
    for (int q = tid; q < uc; q += kLargeThreads) {
      const int row = rel[q];
      const int i = row < w ? row / kT : wt + (row - w) / kT;
      if (i % P == p) own[(i / P) * kT + (row - tile_start(lf, i))] += t[q];  // rows of one child are distinct
    }
    __syncthreads();
  }
  // ---- forward substitution
  int cur_k = -1, n = 0;
  while (cons.next(lf, p, P, it)) {
    {
      SolveItem nx;
      if (prod.next(lf, p, P, nx)) issue_item(stages + ((n + kSStages - 1) % kSStages) * kT * kSLd, nx, lf, L, linv);
      cp_async_commit();
    }
    const double* T = stages + (n % kSStages) * kT * kSLd;
    const int k = it.k;
    if (it.kind == 1 && k != cur_k) {
      // y_k of another CTA: every thread polls its own LL slot (tiles keep streaming in meanwhile)
      if (tid < kT) yk[tid] = tid < tile_size(lf, k) ? ll_load(ld.ll_y + fd.f_piv[s] + k * kT + tid, epoch) : 0.0;
      cur_k = k;
    }
    cp_async_wait<kSStages - 1>();
    __syncthreads();
    if (it.kind == 0) {
      // y_k = L_kk^-1 f_k, published for the other CTAs
      const int nb = tile_size(lf, k);
      smem_gemv_part(T, kT, kT, own + (k / P) * kT, red);
      __syncthreads();
      if (tid < kT) {
        const double v = red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid];
        yk[tid] = tid < nb ? v : 0.0;
        if (tid < nb) {
          fd.ywork[fd.f_piv[s] + k * kT + tid] = v;
          ll_store(ld.ll_y + fd.f_piv[s] + k * kT + tid, v, epoch);
        }
      }
      __syncthreads();
      cur_k = k;
    } else {
      const int i = it.i;
      const int ni = tile_size(lf, i);
      smem_gemv_part(T, ni, tile_size(lf, k), yk, red);
      __syncthreads();
      if (tid < ni) own[(i / P) * kT + tid] -= red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid];
      __syncthreads();
    }
    ++n;
  }
  cp_async_wait<0>();
  // ---- update part -> twork (pivot part was published tile by tile)
  for (int q = tid; q < n_own * kT; q += kLargeThreads) {
    const int i = p + (q / kT) * P, r = q % kT;
    if (i >= wt && r < tile_size(lf, i)) fd.twork[fd.f_toff[s] + (tile_start(lf, i) - w) + r] = own[q];
  }
}

__global__ void __launch_bounds__(kLargeThreads, 1)
    large_solve_bwd2_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld, int lf0, int P, unsigned epoch) {
  extern __shared__ double smf[];  // stages | x: 64 | g: 64 | red: 256
  if (ctrl->done) return;
  const LargeFront lf = ld.lf[lf0 + blockIdx.x / P];
  const int p = blockIdx.x % P;
  const int s = lf.front, w = lf.w, nt = lf.nt, wt = lf.wt;
  double* stages = smf;
  double* xi = stages + kSStages * kT * kSLd;
  double* g = xi + kT;
  double* red = g + kT;
  const double* L = fd.fronts + lf.off;
  const double* linv = ld.linv + lf.linv_off;
  uint4* contrib = ld.ll_contrib + lf.contrib_off;  // LL slots [pivot tile][row tile][64]
  const int tid = threadIdx.x;
  const int32_t* rows = fd.f_rows + fd.f_rows_ptr[s];
  BwdIter prod, cons;
  prod.init(lf, p, P);
  cons.init(lf, p, P);
  SolveItem it;
  for (int st = 0; st < kSStages - 1; ++st) {
    if (prod.next(lf, P, it)) issue_item(stages + st * kT * kSLd, it, lf, L, linv);
    cp_async_commit();
  }
  int cur_i = -1, n = 0;
  while (cons.next(lf, P, it)) {
    {
      SolveItem nx;
      if (prod.next(lf, P, nx)) issue_item(stages + ((n + kSStages - 1) % kSStages) * kT * kSLd, nx, lf, L, linv);
      cp_async_commit();
    }
    const double* T = stages + (n % kSStages) * kT * kSLd;
    const int i = it.i;
    const int ni = tile_size(lf, i), ri = tile_start(lf, i);
    if (it.kind == 0) {
      // pivot tile i: all contributions of the rows below (polled in their LL slots), then x_i = L_ii^-T (y_i - sum)
      {
        const int t = tid & 63, gq = tid >> 6;
        double v = 0.0;
        if (t < ni) {
          // four slots per round trip: the loads are issued together (most contributions arrived long ago); a slot
          // whose epoch is not there yet is polled afterwards
          for (int r0 = i + 1 + gq; r0 < nt; r0 += 16) {
            unsigned lo[4], f1[4], hi[4], f2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + 4 * u;
              if (r < nt) {
                const uint4* q = contrib + ((size_t)i * nt + r) * kT + t;
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(lo[u]), "=r"(f1[u]), "=r"(hi[u]), "=r"(f2[u])
                             : "l"(q)
                             : "memory");
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + 4 * u;
              if (r < nt) {
                if (f1[u] == epoch && f2[u] == epoch)
                  v += __longlong_as_double((long long)(((unsigned long long)hi[u] << 32) | lo[u]));
                else
                  v += ll_load(contrib + ((size_t)i * nt + r) * kT + t, epoch);
              }
            }
          }
        }
        red[gq * 64 + t] = v;
      }
      __syncthreads();
      if (tid < kT)
        g[tid] = tid < ni ? fd.ywork[fd.f_piv[s] + ri + tid] - (red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid]) : 0.0;
      cp_async_wait<kSStages - 1>();
      __syncthreads();
      smem_gemv_t_part(T, kT, kT, g, red);
      __syncthreads();
      if (tid < kT) {
        const double v = red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid];
        xi[tid] = tid < ni ? v : 0.0;
        if (tid < ni) fd.ywork[fd.f_piv[s] + ri + tid] = v;
      }
      __syncthreads();
      cur_i = i;
    } else {
      if (i != cur_i) {
        // update rows: x known from the ancestors
        if (tid < kT) xi[tid] = tid < ni ? fd.ywork[rows[ri - w + tid]] : 0.0;
        cur_i = i;
      }
      cp_async_wait<kSStages - 1>();
      __syncthreads();
      const int k = it.k;
      const int nk = tile_size(lf, k);
      smem_gemv_t_part(T, ni, nk, xi, red);
      __syncthreads();
      if (tid < nk)
        ll_store(contrib + ((size_t)k * nt + i) * kT + tid, red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid], epoch);
      __syncthreads();
    }
    ++n;
  }
  cp_async_wait<0>();
}

static bool solve_v1() { return getenv("SFX_SOLVE_V1") != nullptr; }  // read per launch: tests toggle it

void launch_large_solve_fwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld,
                            const LargeLevel& lv, const double* rhs_static, StatePtrs sp, int use_state_rhs,
                            unsigned epoch) {
  if (lv.n_lf == 0) return;
  const int P = lv.solve_p;
  const size_t smem = (size_t)(((lv.max_nt + P - 1) / P) * kT + kT + 256) * sizeof(double);
  if (solve_v1()) {
    large_solve_fwd_kernel<<<lv.n_lf * P, 256, smem, st>>>(ctrl, fd, ld, rhs_static, sp, use_state_rhs, lv.lf0, P);
  } else {
    const size_t smem2 = smem + (size_t)kSStages * kT * kSLd * sizeof(double);
    large_solve_fwd2_kernel<<<lv.n_lf * P, kLargeThreads, smem2, st>>>(ctrl, fd, ld, rhs_static, sp, use_state_rhs,
                                                                       lv.lf0, P, epoch);
  }
  ++g_launches;
}
void launch_large_solve_bwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld,
                            const LargeLevel& lv, unsigned epoch) {
  if (lv.n_lf == 0) return;
  const int P = lv.solve_p;
  if (solve_v1()) {
    large_solve_bwd_kernel<<<lv.n_lf * P, 256, 3 * kT * sizeof(double), st>>>(ctrl, fd, ld, lv.lf0, P);
  } else {
    const size_t smem2 = (size_t)(kSStages * kT * kSLd + 2 * kT + 256) * sizeof(double);
    large_solve_bwd2_kernel<<<lv.n_lf * P, kLargeThreads, smem2, st>>>(ctrl, fd, ld, lv.lf0, P, epoch);
  }
  ++g_launches;
}

// Right-hand sides of the fused fronts for the forward substitution inside the factor kernel: pivot rows from the
// system right-hand side, update rows zero, plus the update vectors of children factored by earlier launches.
__global__ void __launch_bounds__(256) large_fwd_init_kernel(const Ctrl* __restrict__ ctrl, FrontDev fd, LargeDev ld, int lf0,
                                                              int first_fused_level, const int32_t* __restrict__ f_level,
                                                              const double* __restrict__ rhs_static, StatePtrs sp,
                                                              int use_state_rhs) {
  if (ctrl->done) return;
  const LargeFront lf = ld.lf[lf0 + blockIdx.x];
  const int s = lf.front, w = lf.w, m = lf.m;
  const double* rhs = use_state_rhs ? sp.rhs[ctrl->init_idx] : rhs_static;
  double* b = ld.fwd_b + lf.fb_off;
  for (int r = threadIdx.x; r < m; r += 256) b[r] = r < w ? rhs[fd.scalar_perm[fd.f_piv[s] + r]] : 0.0;
  __syncthreads();
  for (int ci = fd.f_child_ptr[s]; ci < fd.f_child_ptr[s + 1]; ++ci) {
    const int c = fd.f_child[ci];
    if (f_level[c] >= first_fused_level) continue;  // added by a task of the factor kernel
    const int uc = fd.f_u[c];
    const double* t = fd.twork + fd.f_toff[c];
    const int32_t* rel = fd.f_rel + fd.f_rows_ptr[c];
    for (int q = threadIdx.x; q < uc; q += 256) b[rel[q]] += t[q];  // rows of one child are distinct
    __syncthreads();
  }
}
void launch_large_fwd_init(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld, int lf0, int n_lf,
                           int first_fused_level, const int32_t* f_level, const double* rhs_static, StatePtrs sp,
                           int use_state_rhs) {
  if (n_lf <= 0) return;
  large_fwd_init_kernel<<<n_lf, 256, 0, st>>>(ctrl, fd, ld, lf0, first_fused_level, f_level, rhs_static, sp, use_state_rhs);
  ++g_launches;
}

// All levels from the first fused one up in ONE launch: the assembly jobs that do not depend on fronts of this
// launch (children factored by earlier launches), then the tile tasks of every fused front in the order of the
// host-side list schedule (sfx_api.cu: build_fused_schedule), extend-adds included.
void launch_large_fused(cudaStream_t st, Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld, int t0, int t1, int j0,
                        int j1, int queue_slot, const double* sys_static, StatePtrs sp, int use_state_H,
                        const double* dvec, int fwd) {
  if (j1 > j0) {
    large_assemble_kernel<<<(j1 - j0 + 7) / 8, 256, 0, st>>>(ctrl, fd, ld, sys_static, sp, use_state_H, dvec, j0, j1, 0);
    ++g_launches;
  }
  const int ntask = t1 - t0;
  if (ntask <= 0) return;
  const int cap = large_factor_resident_ctas();
  const int grid = ntask < cap ? ntask : cap;
  large_factor_kernel<<<grid, kLargeThreads, kFactorSmem, st>>>(ctrl, fd, ld, t0, t1, queue_slot, fwd); ++g_launches;
}

// The spin-waits of the tile-DAG kernel need every CTA of a launch resident: grids are capped at what the device
// holds (occupancy x SM count), never at a hard-coded SM count.
static int g_resident_ctas = 0;
int large_factor_resident_ctas() {
  if (g_resident_ctas == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // the occupancy query answers 0 for this much dynamic shared memory until the kernel has been allowed to use it
    cudaFuncSetAttribute(large_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFactorSmem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, large_factor_kernel, kLargeThreads, kFactorSmem);
    g_resident_ctas = sms * (per_sm > 0 ? per_sm : 1);
  }
  return g_resident_ctas;
}

cudaError_t configure_large_kernels() {
  cudaError_t e = cudaFuncSetAttribute(large_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFactorSmem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(large_solve_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(large_solve_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

}  // namespace sfx

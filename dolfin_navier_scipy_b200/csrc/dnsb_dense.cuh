// dnsb_dense.cuh -- dense coarse-level solve  Y = alpha*(D X [+ mass term])
//
// D is the (n x n, row-major, fp64) inverse of the coarsest multigrid operator
// (for the cylinder meshes: the lumped pressure Schur complement itself),
// X is n x nb (members fastest).  For nb > 8 this is a skinny fp64 GEMM
// (2 n^2 nb flop, 8 n^2 bytes of D): CUDA-core DFMA bound for nb >= 32
// (B200: 64 DFMA/clk/SM), D-bandwidth bound below.  No tensor cores: fp64.
//
// Tiling: CTA = 128 threads computes a 64-row x TN-member tile over one K
// split; thread micro-tile 8 rows x TN/16 members (32 accumulators for TN=64),
// operands staged through shared memory (double buffered, register prefetch):
// per k, 4 LDS.128 (8 D values, warp-broadcast) + TN/16 LDS.64 feed 8*TN/16
// DFMA.  Split-K partial sums are combined by k_dense_epilogue in a fixed
// order (deterministic).
#pragma once
#include <cuda_runtime.h>

#define DGK_TM 64
#define DGK_TK 16

template <int TN>
__global__ void __launch_bounds__(128)
k_dense_gemm_splitk(const double *__restrict__ D, const double *__restrict__ X,
                    double *__restrict__ part, int n, int nb, int kchunk) {
  __shared__ __align__(16) double sD[2][DGK_TK][DGK_TM + 2];
  __shared__ __align__(16) double sX[2][DGK_TK][TN];
  constexpr int MC = TN / 16;              // member columns per thread
  constexpr int XQ = (DGK_TK * TN) / 128;  // X elements staged per thread
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const int row0 = blockIdx.x * DGK_TM;
  const int m0 = blockIdx.z * TN;
  const int kbeg = blockIdx.y * kchunk;
  const int kend = min(n, kbeg + kchunk);
  double acc[8][MC];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < MC; ++c) acc[r][c] = 0.0;
  double rdx[4], rdy[4], rx[XQ];

  auto gload = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + 128 * q;
      const int r = e >> 3, kk2 = e & 7;
      const int gi = row0 + r, gk = k0 + 2 * kk2;
      const double *p = D + (size_t)gi * n + gk;
      rdx[q] = (gi < n && gk < kend) ? p[0] : 0.0;
      rdy[q] = (gi < n && gk + 1 < kend) ? p[1] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < XQ; ++q) {
      const int e = tid + 128 * q;
      const int kk = e / TN, c = e % TN;
      const int gk = k0 + kk, gm = m0 + c;
      rx[q] = (gk < kend && gm < nb) ? X[(size_t)gk * nb + gm] : 0.0;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + 128 * q;
      const int r = e >> 3, kk2 = e & 7;
      sD[buf][2 * kk2][r] = rdx[q];
      sD[buf][2 * kk2 + 1][r] = rdy[q];
    }
#pragma unroll
    for (int q = 0; q < XQ; ++q) {
      const int e = tid + 128 * q;
      sX[buf][e / TN][e % TN] = rx[q];
    }
  };

  if (kbeg < kend) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += DGK_TK) {
    const bool more = k0 + DGK_TK < kend;
    if (more) gload(k0 + DGK_TK);
#pragma unroll
    for (int kk = 0; kk < DGK_TK; ++kk) {
      const double2 *pd = reinterpret_cast<const double2 *>(&sD[buf][kk][ty * 8]);
      const double2 d01 = pd[0], d23 = pd[1], d45 = pd[2], d67 = pd[3];
      const double d[8] = {d01.x, d01.y, d23.x, d23.y, d45.x, d45.y, d67.x, d67.y};
      double xv[MC];
#pragma unroll
      for (int c = 0; c < MC; ++c) xv[c] = sX[buf][kk][tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < MC; ++c) acc[r][c] += d[r] * xv[c];
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  double *out = part + (size_t)blockIdx.y * n * nb;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int gi = row0 + ty * 8 + r;
    if (gi >= n) continue;
#pragma unroll
    for (int c = 0; c < MC; ++c) {
      const int gm = m0 + tx + 16 * c;
      if (gm < nb) out[(size_t)gi * nb + gm] = acc[r][c];
    }
  }
}

// Y[i,m] = alpha*(sum_s part[s][i][m] + add_scale[m]*add_dinv[i]*X[i,m])
__global__ void k_dense_epilogue(const double *__restrict__ part, int nsplit,
                                 const double *__restrict__ X,
                                 double *__restrict__ Y, int n, int nb,
                                 double alpha,
                                 const double *__restrict__ add_dinv,
                                 const double *__restrict__ add_scale) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nn = (size_t)n * nb;
  if (t >= nn) return;
  double v = 0.0;
  for (int s = 0; s < nsplit; ++s) v += part[(size_t)s * nn + t];
  if (add_dinv) v += add_scale[t % nb] * add_dinv[t / nb] * X[t];
  Y[t] = alpha * v;
}

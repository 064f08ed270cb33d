// dnsb_dense.cuh -- dense coarse-level solve  Y = alpha*(D X [+ mass term])
//
// D is the (n x n, row-major, fp64) inverse of the coarsest multigrid operator
// (for the cylinder meshes: the lumped pressure Schur complement itself),
// X is n x nb (members fastest).  For nb > 8 this is a skinny fp64 GEMM
// (2 n^2 nb flop, 8 n^2 bytes of D): CUDA-core DFMA bound for nb >= 32
// (B200: 64 DFMA/clk/SM = 37 TFLOP/s), D-bandwidth bound below.  fp64 has no
// tcgen05 path, so this is a DFMA kernel.
//
// Work decomposition (stream-K): the (row tile, k step) units -- 64 rows x 16
// columns of D each -- are numbered row tile major and cut into P equal
// contiguous ranges, one per CTA, P = a fixed number of CTAs per SM; a CTA
// whose range crosses row tiles writes one partial tile per segment.  Every SM
// gets the same number of units (a plain split-K grid of 300 CTAs on 148 SMs
// left a third of the SM cycles idle).  k_dense_epilogue sums the partial
// tiles of a row in CTA order: deterministic.
//
// CTA = 128 threads, tile 64 rows x TN members, thread micro-tile 8 x TN/16
// (32 accumulators for TN=64); operands staged through shared memory (double
// buffered, register prefetch): per k, 4 LDS.128 (8 D values, warp-broadcast)
// + TN/16 LDS.64 feed 8*TN/16 DFMA.
#pragma once
#include <cuda_runtime.h>

#define DGK_TM 64
#define DGK_TK 16

struct DenseSplit {
  int ksteps;   // k steps per row tile = ceil(n / DGK_TK)
  int upc;      // units per CTA
  int nctas;    // P
  int maxseg;   // partial tiles per CTA
};

template <int TN>
__global__ void __launch_bounds__(128)
k_dense_gemm_streamk(const double *__restrict__ D, const double *__restrict__ X,
                     double *__restrict__ part, int n, int nb, DenseSplit sp) {
  __shared__ __align__(16) double sD[2][DGK_TK][DGK_TM + 2];
  __shared__ __align__(16) double sX[2][DGK_TK][TN];
  constexpr int MC = TN / 16;              // member columns per thread
  constexpr int XQ = (DGK_TK * TN) / 128;  // X elements staged per thread
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const int m0 = blockIdx.y * TN;
  const long utotal = (long)((n + DGK_TM - 1) / DGK_TM) * sp.ksteps;
  long u = (long)blockIdx.x * sp.upc;
  const long uend = min(utotal, u + sp.upc);
  int seg = 0;
  while (u < uend) {
    const int rt = (int)(u / sp.ksteps);
    const int ks0 = (int)(u - (long)rt * sp.ksteps);
    const int ks1 = (int)min((long)sp.ksteps, ks0 + (uend - u));
    const int row0 = rt * DGK_TM;
    const int kbeg = ks0 * DGK_TK, kend = min(n, ks1 * DGK_TK);
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

    __syncthreads();   // the previous segment is done with the buffers
    gload(kbeg);
    sstore(0);
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
    double *out = part + ((size_t)blockIdx.x * sp.maxseg + seg) * DGK_TM * nb;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int rl = ty * 8 + r;
      if (row0 + rl >= n) continue;
#pragma unroll
      for (int c = 0; c < MC; ++c) {
        const int gm = m0 + tx + 16 * c;
        if (gm < nb) out[(size_t)rl * nb + gm] = acc[r][c];
      }
    }
    u += ks1 - ks0;
    ++seg;
  }
}

// Y[i,m] = alpha*(sum over the CTAs c that hold a partial tile of row i
//                 + add_scale[m]*add_dinv[i]*X[i,m])
__global__ void k_dense_epilogue(const double *__restrict__ part, DenseSplit sp,
                                 const double *__restrict__ X,
                                 double *__restrict__ Y, int n, int nb,
                                 double alpha,
                                 const double *__restrict__ add_dinv,
                                 const double *__restrict__ add_scale) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * nb) return;
  const int i = (int)(t / nb), m = (int)(t % nb);
  const int rt = i / DGK_TM, rl = i % DGK_TM;
  const long u0 = (long)rt * sp.ksteps, u1 = u0 + sp.ksteps - 1;
  const int c0 = (int)(u0 / sp.upc), c1 = (int)(u1 / sp.upc);
  double v = 0.0;
  for (int c = c0; c <= c1; ++c) {
    const int first_rt = (int)(((long)c * sp.upc) / sp.ksteps);
    const int seg = rt - first_rt;
    v += part[(((size_t)c * sp.maxseg + seg) * DGK_TM + rl) * nb + m];
  }
  if (add_dinv) v += add_scale[m] * add_dinv[i] * X[t];
  Y[t] = alpha * v;
}

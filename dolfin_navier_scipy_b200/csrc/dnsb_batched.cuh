// dnsb_batched.cuh -- the batched (ensemble, nb > 1) row kernels of libdnsb200
//
// Layout: x[i*nb + m] (member fastest).  One thread owns one (row, member)
// pair; the 32 lanes of a warp therefore cover 32 consecutive members of one
// row (nb >= 32) or a few consecutive rows (nb < 32).  The CSR entries of the
// warp's rows are contiguous in memory: the warp loads them once, coalesced,
// into its own shared-memory slice (column offset pre-multiplied by nb, value
// pair (v1, v2) as one 16-byte word) and every lane then walks its row from
// shared memory, so that the only global loads in the inner loop are the
// (coalesced, independent, 4-way unrolled) gathers of x.  The sum over a row
// runs in CSR order with one accumulator: results do not depend on the chunking
// and are bit-reproducible.
//
// All kernels are HBM/L2-bandwidth bound (fp64, no tensor cores).
#pragma once
#include <cuda_runtime.h>

#define SPB_THREADS 256
#define SPB_WARPS (SPB_THREADS / 32)
#define SPB_CAP 96   // CSR entries staged per warp and chunk

struct SpbSmem2 {
  double2 v[SPB_WARPS][SPB_CAP];
  int off[SPB_WARPS][SPB_CAP];
};
struct SpbSmem1 {
  double v[SPB_WARPS][SPB_CAP];
  int off[SPB_WARPS][SPB_CAP];
};

// A CTA owns `gpc` consecutive groups of SPB_THREADS (row, member) pairs, i.e.
// a chunk of gpc*SPB_THREADS/nb consecutive rows, and walks them group by
// group: with the mesh-local (Hilbert) numbering of the unknowns the rows of a
// chunk share most of their columns, so the gathered x rows are served by the
// SM's L1 instead of L2 (measured reuse on cylinder_4: 9x for 64-row chunks).
// Inside the loop: (row, member) of this thread, row range of its warp.
#define SPB_FOR_GROUPS()                                                        \
  const long total_ = (long)A.nrows * nb;                                       \
  const int lane = threadIdx.x & 31;                                           \
  for (int g_ = 0; g_ < gpc; ++g_)

#define SPB_ROWMAP()                                                            \
  const long t_ = ((long)blockIdx.x * gpc + g_) * SPB_THREADS + threadIdx.x;    \
  const long tw0_ = t_ - lane;                                                  \
  if (tw0_ >= total_) break;                                                    \
  const bool valid = t_ < total_;                                               \
  const int row = valid ? (int)((unsigned)t_ / (unsigned)nb) : A.nrows - 1;     \
  const int m = valid ? (int)(t_ - (long)row * nb) : 0;                         \
  const int wrow0 = __shfl_sync(0xffffffffu, row, 0);                           \
  const int wrow1 = __shfl_sync(0xffffffffu, row, 31);

// sum_k (v1[k] + cm*v2[k]) * x[indices[k]*nb + m] over the row of this lane
template <bool HAS2>
__device__ __forceinline__ double
spb_rowdot(const CsrDev &A, double cm, const double *__restrict__ xm, int nb,
           int row, bool valid, int wrow0, int wrow1, int lane, double2 *sv2,
           double *sv1, int *soff) {
  const int e_lo = A.indptr[wrow0], e_hi = A.indptr[wrow1 + 1];
  const int k0 = valid ? A.indptr[row] : 0;
  const int k1 = valid ? A.indptr[row + 1] : 0;
  double acc = 0.0;
  for (int base = e_lo; base < e_hi; base += SPB_CAP) {
    const int cnt = min(SPB_CAP, e_hi - base);
    __syncwarp();
    for (int e = lane; e < cnt; e += 32) {
      soff[e] = A.indices[base + e] * nb;
      if (HAS2) sv2[e] = make_double2(A.v1[base + e], A.v2[base + e]);
      else sv1[e] = A.v1[base + e];
    }
    __syncwarp();
    int k = max(k0, base) - base;
    const int kend = min(k1, base + cnt) - base;
    for (; k + 4 <= kend; k += 4) {
      const double x0 = xm[soff[k]], x1 = xm[soff[k + 1]];
      const double x2 = xm[soff[k + 2]], x3 = xm[soff[k + 3]];
      if (HAS2) {
        const double2 a0 = sv2[k], a1 = sv2[k + 1], a2 = sv2[k + 2], a3 = sv2[k + 3];
        acc += (a0.x + cm * a0.y) * x0;
        acc += (a1.x + cm * a1.y) * x1;
        acc += (a2.x + cm * a2.y) * x2;
        acc += (a3.x + cm * a3.y) * x3;
      } else {
        acc += sv1[k] * x0;
        acc += sv1[k + 1] * x1;
        acc += sv1[k + 2] * x2;
        acc += sv1[k + 3] * x3;
      }
    }
    for (; k < kend; ++k) {
      const double xv = xm[soff[k]];
      if (HAS2) {
        const double2 a = sv2[k];
        acc += (a.x + cm * a.y) * xv;
      } else {
        acc += sv1[k] * xv;
      }
    }
  }
  return acc;
}

#define SPB_SMEM(HAS2)                                                          \
  __shared__ typename SpbSel<HAS2>::type sm_;                                   \
  double2 *sv2 = SpbSel<HAS2>::v2(sm_, threadIdx.x >> 5);                       \
  double *sv1 = SpbSel<HAS2>::v1(sm_, threadIdx.x >> 5);                        \
  int *soff = sm_.off[threadIdx.x >> 5];

template <bool HAS2> struct SpbSel;
template <> struct SpbSel<true> {
  typedef SpbSmem2 type;
  static __device__ __forceinline__ double2 *v2(SpbSmem2 &s, int w) { return s.v[w]; }
  static __device__ __forceinline__ double *v1(SpbSmem2 &, int) { return nullptr; }
};
template <> struct SpbSel<false> {
  typedef SpbSmem1 type;
  static __device__ __forceinline__ double2 *v2(SpbSmem1 &, int) { return nullptr; }
  static __device__ __forceinline__ double *v1(SpbSmem1 &s, int w) { return s.v[w]; }
};

// y = alpha*A*x + beta*z   (z may alias y; ignored when beta == 0)
// Algorithmic bytes: (12|20)*nnz + 4(nrows+1) + 8*nb*(ncols + nrows [+ nrows]).
template <bool HAS2>
__global__ void __launch_bounds__(SPB_THREADS)
k_spmm_b(CsrDev A, const double *__restrict__ coef, const double *__restrict__ x,
         const double *z, double *y, int nb, int gpc, double alpha, double beta) {
  SPB_SMEM(HAS2)
  SPB_FOR_GROUPS() {
    SPB_ROWMAP()
    const double cm = HAS2 ? coef[m] : 0.0;
    const double zin = (beta != 0.0 && valid) ? z[t_] : 0.0;
    const double acc = spb_rowdot<HAS2>(A, cm, x + m, nb, row, valid, wrow0, wrow1, lane, sv2, sv1, soff);
    if (valid) y[t_] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * zin;
  }
}

// Chebyshev start, fused with the gradient coupling of the block-triangular
// preconditioner (A = JT, may have nnz = 0 rows):
//   res = rv - JT*zp ;  d = dinv*res/theta      (z is NOT written: the first
//   step forms z = d0 + d1; with a single step the caller passes d = z)
__global__ void __launch_bounds__(SPB_THREADS)
k_cheb_init_b(CsrDev A, const double *__restrict__ zp, const double *__restrict__ rv,
              const double *__restrict__ dinv, double *__restrict__ res,
              double *__restrict__ d, int nb, int gpc, double inv_theta) {
  SPB_SMEM(false)
  SPB_FOR_GROUPS() {
    SPB_ROWMAP()
    const long te_ = valid ? t_ : 0;
    const double rv_ = rv[te_], di = dinv[te_];
    const double acc = spb_rowdot<false>(A, 0.0, zp + m, nb, row, valid, wrow0, wrow1, lane, sv2, sv1, soff);
    if (valid) {
      const double r = rv_ - acc;
      res[t_] = r;
      d[t_] = di * r * inv_theta;
    }
  }
}

// Chebyshev step:  r = res - F*d ;  dn = c1*d + c2*dinv*r ;  z (+)= dn
//   FIRST: z = d + dn (z not read);  LAST: res and dn are not written.
// Algorithmic bytes per (row, member): gather of d (8) + res (8) + dinv (8)
// + z read (8, not FIRST) + z write (8) + res/dn write (16, not LAST);
// matrix: 20 B/nnz (HAS2) shared by all members.
template <bool HAS2, bool FIRST, bool LAST>
__global__ void __launch_bounds__(SPB_THREADS)
k_cheb_step_b(CsrDev A, const double *__restrict__ coef, const double *__restrict__ d,
              const double *__restrict__ dinv, double *res, double *__restrict__ dn,
              double *z, int nb, int gpc, double c1, double c2) {
  SPB_SMEM(HAS2)
  SPB_FOR_GROUPS() {
    SPB_ROWMAP()
    const double cm = HAS2 ? coef[m] : 0.0;
    // operands of the update first: their latency overlaps the row product
    const long te_ = valid ? t_ : 0;
    const double r_old = res[te_], dold = d[te_], di = dinv[te_];
    const double z_old = FIRST ? 0.0 : z[te_];
    const double acc = spb_rowdot<HAS2>(A, cm, d + m, nb, row, valid, wrow0, wrow1, lane, sv2, sv1, soff);
    if (valid) {
      const double r = r_old - acc;
      const double dd = c1 * dold + c2 * di * r;
      if (!LAST) {
        res[t_] = r;
        dn[t_] = dd;
      }
      z[t_] = (FIRST ? dold : z_old) + dd;
    }
  }
}

// ---------------------------------------------------------------------------
// Gram-Schmidt of FGMRES, batched and deterministic.
// Block = RPB x nb threads (RPB a power of two), thread = (rr, m); the block
// owns a contiguous chunk of rows and the thread the rows r0+rr, r0+rr+RPB, ...
// of it (at most GS_RPT), which it keeps in registers.
// ---------------------------------------------------------------------------
#define GS_RPT 16   // rows per thread held in registers

// partial[(b*(nvec+1) + i)*nb + m] = sum_chunk V_i*w  (i < nvec),  i = nvec: w*w
// Algorithmic bytes: 8*n*nb*(nvec + 1).
__global__ void __launch_bounds__(256)
k_mdot_b(const double *__restrict__ V, size_t vstride, int nvec,
         const double *__restrict__ w, int n, int nb, int rpb, int rows_per_block,
         double *__restrict__ partial) {
  extern __shared__ double sred[];   // (nvec+1) x blockDim
  const int m = threadIdx.x % nb, rr = threadIdx.x / nb;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(n, r0 + rows_per_block);
  const int nthr = rpb * nb;
  double wr[GS_RPT];
  double ww = 0.0;
#pragma unroll
  for (int q = 0; q < GS_RPT; ++q) {
    const int r = r0 + rr + q * rpb;
    wr[q] = (r < r1) ? w[(size_t)r * nb + m] : 0.0;
    ww += wr[q] * wr[q];
  }
  for (int i = 0; i < nvec; ++i) {
    const double *vi = V + (size_t)i * vstride;
    double vv[GS_RPT];
#pragma unroll
    for (int q = 0; q < GS_RPT; ++q) {
      const int r = r0 + rr + q * rpb;
      vv[q] = (r < r1) ? vi[(size_t)r * nb + m] : 0.0;
    }
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < GS_RPT; ++q) acc += vv[q] * wr[q];
    sred[(size_t)i * nthr + threadIdx.x] = acc;
  }
  sred[(size_t)nvec * nthr + threadIdx.x] = ww;
  __syncthreads();
  // fixed-order sum over rr
  for (int o = threadIdx.x; o < (nvec + 1) * nb; o += nthr) {
    const int i = o / nb, mm = o % nb;
    double s = 0.0;
    for (int q = 0; q < rpb; ++q) s += sred[(size_t)i * nthr + q * nb + mm];
    partial[((size_t)blockIdx.x * (nvec + 1) + i) * nb + mm] = s;
  }
}

// vnext = w - sum_i h[i,m]*V_i  (not normalised);
// partial2[b*nb + m] = |vnext|^2 over the chunk of block b
// Algorithmic bytes: 8*n*nb*(nvec + 2).
__global__ void __launch_bounds__(256, 2)
k_gs_update_b(const double *__restrict__ V, size_t vstride, int nvec,
              const double *__restrict__ h, const double *__restrict__ w,
              double *__restrict__ vnext, int n, int nb, int rpb, int rows_per_block,
              double *__restrict__ partial2) {
  extern __shared__ double sred[];   // blockDim
  const int m = threadIdx.x % nb, rr = threadIdx.x / nb;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(n, r0 + rows_per_block);
  double wr[GS_RPT];
#pragma unroll
  for (int q = 0; q < GS_RPT; ++q) {
    const int r = r0 + rr + q * rpb;
    wr[q] = (r < r1) ? w[(size_t)r * nb + m] : 0.0;
  }
  for (int i = 0; i < nvec; ++i) {
    const double *vi = V + (size_t)i * vstride;
    const double hi = h[(size_t)i * nb + m];
    double vv[GS_RPT];
#pragma unroll
    for (int q = 0; q < GS_RPT; ++q) {
      const int r = r0 + rr + q * rpb;
      vv[q] = (r < r1) ? vi[(size_t)r * nb + m] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < GS_RPT; ++q) wr[q] -= hi * vv[q];
  }
  double nrm = 0.0;
#pragma unroll
  for (int q = 0; q < GS_RPT; ++q) {
    const int r = r0 + rr + q * rpb;
    if (r < r1) vnext[(size_t)r * nb + m] = wr[q];
    nrm += wr[q] * wr[q];
  }
  sred[threadIdx.x] = nrm;
  __syncthreads();
  if (rr == 0) {
    double s = 0.0;
    for (int q = 0; q < rpb; ++q) s += sred[q * nb + m];
    partial2[(size_t)blockIdx.x * nb + m] = s;
  }
}

// ---------------------------------------------------------------------------
// Two members per thread (nb even): the same kernels on double2 elements,
// x2[i*(nb/2) + mp] = (x[i*nb + 2mp], x[i*nb + 2mp + 1]).  The ncu source page
// of the one-member kernels (profiles/r1c) shows ~25 issued instructions per
// (entry, warp), two thirds of them address arithmetic, shared-memory reads of
// the staged entry and loop control -- all of it independent of the member.
// With a member pair per thread a warp covers 64 members per pass: that
// overhead, and the shared-memory wavefronts of the entries, are halved per
// member; the gather becomes one LDG.128.  Sums run in the same CSR order:
// results are bit-identical to the one-member kernels.
// ---------------------------------------------------------------------------
#define SPB2_ROWMAP()                                                           \
  const int nb2 = nb >> 1;                                                      \
  const long total_ = (long)A.nrows * nb2;                                      \
  const int lane = threadIdx.x & 31;                                            \
  const long t_ = (long)blockIdx.x * SPB_THREADS + threadIdx.x;                 \
  const long tw0_ = t_ - lane;                                                  \
  if (tw0_ >= total_) return;                                                   \
  const bool valid = t_ < total_;                                               \
  const int row = valid ? (int)((unsigned)t_ / (unsigned)nb2) : A.nrows - 1;    \
  const int mp = valid ? (int)(t_ - (long)row * nb2) : 0;                       \
  const int wrow0 = __shfl_sync(0xffffffffu, row, 0);                           \
  const int wrow1 = __shfl_sync(0xffffffffu, row, 31);

template <bool HAS2>
__device__ __forceinline__ double2
spb2_rowdot(const CsrDev &A, double2 cm, const double2 *__restrict__ xm, int nb2,
            int row, bool valid, int wrow0, int wrow1, int lane, double2 *sv2,
            double *sv1, int *soff) {
  const int e_lo = A.indptr[wrow0], e_hi = A.indptr[wrow1 + 1];
  const int k0 = valid ? A.indptr[row] : 0;
  const int k1 = valid ? A.indptr[row + 1] : 0;
  double ax = 0.0, ay = 0.0;
  for (int base = e_lo; base < e_hi; base += SPB_CAP) {
    const int cnt = min(SPB_CAP, e_hi - base);
    __syncwarp();
    for (int e = lane; e < cnt; e += 32) {
      soff[e] = A.indices[base + e] * nb2;
      if (HAS2) sv2[e] = make_double2(A.v1[base + e], A.v2[base + e]);
      else sv1[e] = A.v1[base + e];
    }
    __syncwarp();
    int k = max(k0, base) - base;
    const int kend = min(k1, base + cnt) - base;
    for (; k + 4 <= kend; k += 4) {
      const double2 x0 = xm[soff[k]], x1 = xm[soff[k + 1]];
      const double2 x2 = xm[soff[k + 2]], x3 = xm[soff[k + 3]];
      if (HAS2) {
        const double2 a0 = sv2[k], a1 = sv2[k + 1], a2 = sv2[k + 2], a3 = sv2[k + 3];
        ax += (a0.x + cm.x * a0.y) * x0.x;  ay += (a0.x + cm.y * a0.y) * x0.y;
        ax += (a1.x + cm.x * a1.y) * x1.x;  ay += (a1.x + cm.y * a1.y) * x1.y;
        ax += (a2.x + cm.x * a2.y) * x2.x;  ay += (a2.x + cm.y * a2.y) * x2.y;
        ax += (a3.x + cm.x * a3.y) * x3.x;  ay += (a3.x + cm.y * a3.y) * x3.y;
      } else {
        const double b0 = sv1[k], b1 = sv1[k + 1], b2 = sv1[k + 2], b3 = sv1[k + 3];
        ax += b0 * x0.x;  ay += b0 * x0.y;
        ax += b1 * x1.x;  ay += b1 * x1.y;
        ax += b2 * x2.x;  ay += b2 * x2.y;
        ax += b3 * x3.x;  ay += b3 * x3.y;
      }
    }
    for (; k < kend; ++k) {
      const double2 xv = xm[soff[k]];
      if (HAS2) {
        const double2 a = sv2[k];
        ax += (a.x + cm.x * a.y) * xv.x;
        ay += (a.x + cm.y * a.y) * xv.y;
      } else {
        const double b = sv1[k];
        ax += b * xv.x;
        ay += b * xv.y;
      }
    }
  }
  return make_double2(ax, ay);
}

template <bool HAS2>
__global__ void __launch_bounds__(SPB_THREADS)
k_spmm_b2(CsrDev A, const double *__restrict__ coef, const double2 *__restrict__ x,
          const double2 *z, double2 *y, int nb, double alpha, double beta) {
  SPB2_ROWMAP()
  SPB_SMEM(HAS2)
  const double2 cm = HAS2 ? reinterpret_cast<const double2 *>(coef)[mp] : make_double2(0.0, 0.0);
  const double2 zin = (beta != 0.0 && valid) ? z[t_] : make_double2(0.0, 0.0);
  const double2 acc = spb2_rowdot<HAS2>(A, cm, x + mp, nb2, row, valid, wrow0, wrow1, lane, sv2, sv1, soff);
  if (valid)
    y[t_] = (beta == 0.0) ? make_double2(alpha * acc.x, alpha * acc.y)
                          : make_double2(alpha * acc.x + beta * zin.x, alpha * acc.y + beta * zin.y);
}

__global__ void __launch_bounds__(SPB_THREADS)
k_cheb_init_b2(CsrDev A, const double2 *__restrict__ zp, const double2 *__restrict__ rv,
               const double2 *__restrict__ dinv, double2 *__restrict__ res,
               double2 *__restrict__ d, int nb, double inv_theta) {
  SPB2_ROWMAP()
  SPB_SMEM(false)
  const long te_ = valid ? t_ : 0;
  const double2 rv_ = rv[te_], di = dinv[te_];
  const double2 acc = spb2_rowdot<false>(A, make_double2(0.0, 0.0), zp + mp, nb2, row, valid, wrow0,
                                         wrow1, lane, sv2, sv1, soff);
  if (valid) {
    const double rx = rv_.x - acc.x, ry = rv_.y - acc.y;
    res[t_] = make_double2(rx, ry);
    d[t_] = make_double2(di.x * rx * inv_theta, di.y * ry * inv_theta);
  }
}

template <bool HAS2, bool FIRST, bool LAST>
__global__ void __launch_bounds__(SPB_THREADS)
k_cheb_step_b2(CsrDev A, const double *__restrict__ coef, const double2 *__restrict__ d,
               const double2 *__restrict__ dinv, double2 *res, double2 *__restrict__ dn,
               double2 *z, int nb, double c1, double c2) {
  SPB2_ROWMAP()
  SPB_SMEM(HAS2)
  const double2 cm = HAS2 ? reinterpret_cast<const double2 *>(coef)[mp] : make_double2(0.0, 0.0);
  const long te_ = valid ? t_ : 0;
  const double2 r_old = res[te_], dold = d[te_], di = dinv[te_];
  const double2 z_old = FIRST ? make_double2(0.0, 0.0) : z[te_];
  const double2 acc = spb2_rowdot<HAS2>(A, cm, d + mp, nb2, row, valid, wrow0, wrow1, lane, sv2, sv1, soff);
  if (valid) {
    const double rx = r_old.x - acc.x, ry = r_old.y - acc.y;
    const double ddx = c1 * dold.x + c2 * di.x * rx, ddy = c1 * dold.y + c2 * di.y * ry;
    if (!LAST) {
      res[t_] = make_double2(rx, ry);
      dn[t_] = make_double2(ddx, ddy);
    }
    z[t_] = make_double2((FIRST ? dold.x : z_old.x) + ddx, (FIRST ? dold.y : z_old.y) + ddy);
  }
}

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

#ifndef SPB_THREADS
#define SPB_THREADS 256
#endif
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
  dnsb_pdl_entry();
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
  dnsb_pdl_entry();
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
  dnsb_pdl_entry();
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

// The rows of w stay in registers; the basis vectors are streamed in two
// halves of GS_RPT/2 rows with the loads of the next half issued before the
// current one is consumed (software pipeline): 8-16 independent loads are in
// flight per thread at any time, without a bubble between two vectors.
#define GS_H (GS_RPT / 2)
// basis vectors: streamed once per kernel (evict-first in L2 with DNSB_L2_HINTS >= 2)
#if DNSB_L2_HINTS >= 2
#define GS_LDV(p) __ldcs(p)
#else
#define GS_LDV(p) (*(p))
#endif
#define GS_LOADV(dst, vec, half)                                                \
  _Pragma("unroll") for (int q = 0; q < GS_H; ++q) {                            \
    const int r = r0 + rr + ((half) * GS_H + q) * rpb;                          \
    dst[q] = (r < r1) ? GS_LDV((vec) + (size_t)r * nb + m) : 0.0;               \
  }
#define GS_LOAD(dst, vec, half)                                                 \
  _Pragma("unroll") for (int q = 0; q < GS_H; ++q) {                            \
    const int r = r0 + rr + ((half) * GS_H + q) * rpb;                          \
    dst[q] = (r < r1) ? (vec)[(size_t)r * nb + m] : 0.0;                        \
  }

// partial[(b*(nvec+1) + i)*nb + m] = sum_chunk V_i*w  (i < nvec),  i = nvec: w*w
// Algorithmic bytes: 8*n*nb*(nvec + 1).
__global__ void __launch_bounds__(256)
k_mdot_b(const double *__restrict__ V, size_t vstride, int nvec,
         const double *__restrict__ w, int n, int nb, int rpb, int rows_per_block,
         double *__restrict__ partial) {
  dnsb_pdl_entry();
  extern __shared__ double sred[];   // (nvec+1) x blockDim
  const int m = threadIdx.x % nb, rr = threadIdx.x / nb;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(n, r0 + rows_per_block);
  const int nthr = rpb * nb;
  double wa[GS_H], wb[GS_H], va[GS_H], vb[GS_H];
  GS_LOAD(wa, w, 0)
  GS_LOAD(wb, w, 1)
  if (nvec > 0) { GS_LOADV(va, V, 0) }
  double ww = 0.0;
#pragma unroll
  for (int q = 0; q < GS_H; ++q) ww += wa[q] * wa[q];
#pragma unroll
  for (int q = 0; q < GS_H; ++q) ww += wb[q] * wb[q];
  for (int i = 0; i < nvec; ++i) {
    const double *vi = V + (size_t)i * vstride;
    GS_LOADV(vb, vi, 1)
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < GS_H; ++q) acc += va[q] * wa[q];
    if (i + 1 < nvec) { GS_LOADV(va, vi + vstride, 0) }
#pragma unroll
    for (int q = 0; q < GS_H; ++q) acc += vb[q] * wb[q];
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
              double *__restrict__ partial2, const double *__restrict__ scale) {
  dnsb_pdl_entry();
  extern __shared__ double sred[];   // blockDim
  const int m = threadIdx.x % nb, rr = threadIdx.x / nb;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(n, r0 + rows_per_block);
  double wa[GS_H], wb[GS_H], va[GS_H], vb[GS_H];
  GS_LOAD(wa, w, 0)
  GS_LOAD(wb, w, 1)
  if (nvec > 0) { GS_LOADV(va, V, 0) }
  double hi = nvec > 0 ? h[m] : 0.0;
  for (int i = 0; i < nvec; ++i) {
    const double *vi = V + (size_t)i * vstride;
    GS_LOADV(vb, vi, 1)
    const double hn = (i + 1 < nvec) ? h[(size_t)(i + 1) * nb + m] : 0.0;
#pragma unroll
    for (int q = 0; q < GS_H; ++q) wa[q] -= hi * va[q];
    if (i + 1 < nvec) { GS_LOADV(va, vi + vstride, 0) }
#pragma unroll
    for (int q = 0; q < GS_H; ++q) wb[q] -= hi * vb[q];
    hi = hn;
  }
  double nrm = 0.0;
  // scale != null: the norm is known already (Pythagoras, k_gmres_givens): write the normalised vector
  const double sc = scale ? scale[m] : 1.0;
#pragma unroll
  for (int q = 0; q < GS_H; ++q) {
    const int ra = r0 + rr + q * rpb, rb = r0 + rr + (GS_H + q) * rpb;
    if (ra < r1) vnext[(size_t)ra * nb + m] = scale ? wa[q] * sc : wa[q];
    if (rb < r1) vnext[(size_t)rb * nb + m] = scale ? wb[q] * sc : wb[q];
  }
  if (scale) return;
#pragma unroll
  for (int q = 0; q < GS_H; ++q) nrm += wa[q] * wa[q];
#pragma unroll
  for (int q = 0; q < GS_H; ++q) nrm += wb[q] * wb[q];
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
          const double2 *z, double2 *y, int nb, int row_begin, double alpha, double beta) {
  dnsb_pdl_entry();
  // rows [row_begin, nrows): the tail after the paired rows, or everything
  const int nb2 = nb >> 1;
  const long total_ = (long)(A.nrows - row_begin) * nb2;
  const int lane = threadIdx.x & 31;
  const long tl_ = (long)blockIdx.x * SPB_THREADS + threadIdx.x;
  if (tl_ - lane >= total_) return;
  const bool valid = tl_ < total_;
  const int row = valid ? row_begin + (int)((unsigned)tl_ / (unsigned)nb2) : A.nrows - 1;
  const int mp = valid ? (int)(tl_ - (long)(row - row_begin) * nb2) : 0;
  const int wrow0 = __shfl_sync(0xffffffffu, row, 0);
  const int wrow1 = __shfl_sync(0xffffffffu, row, 31);
  const long t_ = (long)row * nb2 + mp;
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
  dnsb_pdl_entry();
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
  dnsb_pdl_entry();
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

// ---------------------------------------------------------------------------
// Row-pair kernels.  The ncu capture of the member-pair kernels
// (profiles/r1f_ncu_full.md) shows them bound by the L1 data path, not by
// DRAM (DRAM 21-25 %, L1 60 %, 50-65 % long-scoreboard stalls): every gathered
// x value travels L1 -> register for ONE multiply-add.  The two velocity
// components of a P2 node (rows 2k, 2k+1 of F, K and JT in the device
// numbering) have identical column lists, so one gather can feed both rows:
// a thread owns (row pair, member pair), i.e. a 2x2 register tile, and the L1
// traffic of the gather is halved.  `npairs` leading row pairs of the matrix
// are handled (csr->npair_rows / 2, verified on the host when the matrix is
// created); the sums run in CSR order per row: bit-identical results.
// ---------------------------------------------------------------------------
// Explicit fused operations shared by the row-pair kernels and the TMA-staged tile kernel
// (dnsb_tile.cuh): both evaluate EXACTLY the same sequence of roundings, whatever the compiler
// would choose to contract, so that the two paths are bit-identical.
//   acc += (a.x + c*a.y) * x
__device__ __forceinline__ double spp_acc(double acc, double c, double2 a, double x) {
  return __fma_rn(__fma_rn(c, a.y, a.x), x, acc);
}
//   dd = c1*dold + (c2*dinv)*r
__device__ __forceinline__ double spp_dir(double c1, double dold, double c2, double dinv, double r) {
  return __fma_rn(__dmul_rn(c2, dinv), r, __dmul_rn(c1, dold));
}

#ifndef SPP_CAP
#define SPP_CAP 192
#endif
// SPP_CAP: CSR entries (both rows of the warp's pairs) staged per warp

struct SppSmem2 {
  double2 v[SPB_WARPS][SPP_CAP];
  int off[SPB_WARPS][SPP_CAP];
};
struct SppSmem1 {
  double v[SPB_WARPS][SPP_CAP];
  int off[SPB_WARPS][SPP_CAP];
};
template <bool HAS2> struct SppSel;
template <> struct SppSel<true> {
  typedef SppSmem2 type;
  static __device__ __forceinline__ double2 *v2(SppSmem2 &s, int w) { return s.v[w]; }
  static __device__ __forceinline__ double *v1(SppSmem2 &, int) { return nullptr; }
};
template <> struct SppSel<false> {
  typedef SppSmem1 type;
  static __device__ __forceinline__ double2 *v2(SppSmem1 &, int) { return nullptr; }
  static __device__ __forceinline__ double *v1(SppSmem1 &s, int w) { return s.v[w]; }
};

#define SPP_MAP(HAS2)                                                           \
  const int nb2 = nb >> 1;                                                      \
  const long total_ = (long)npairs * nb2;                                       \
  const int lane = threadIdx.x & 31;                                            \
  const long t_ = (long)blockIdx.x * SPB_THREADS + threadIdx.x;                 \
  const long tw0_ = t_ - lane;                                                  \
  if (tw0_ >= total_) return;                                                   \
  const bool valid = t_ < total_;                                               \
  const int rp = valid ? (int)((unsigned)t_ / (unsigned)nb2) : npairs - 1;      \
  const int mp = valid ? (int)(t_ - (long)rp * nb2) : 0;                        \
  const int wp0 = __shfl_sync(0xffffffffu, rp, 0);                              \
  const int wp1 = __shfl_sync(0xffffffffu, rp, 31);                             \
  __shared__ typename SppSel<HAS2>::type sm_;                                   \
  double2 *sv2 = SppSel<HAS2>::v2(sm_, threadIdx.x >> 5);                       \
  double *sv1 = SppSel<HAS2>::v1(sm_, threadIdx.x >> 5);                        \
  int *soff = sm_.off[threadIdx.x >> 5];                                        \
  const size_t ta_ = (size_t)(2 * rp) * nb2 + mp, tb_ = ta_ + nb2;

// (accA, accB) of rows (2rp, 2rp+1) for the member pair of this lane
template <bool HAS2>
__device__ __forceinline__ void
spp_rowdots(const CsrDev &A, double2 cm, const double2 *__restrict__ xm, int nb2, int rp,
            bool valid, int wp0, int wp1, int lane, double2 *sv2, double *sv1, int *soff,
            double2 &accA, double2 &accB) {
  const int e_lo = A.indptr[2 * wp0], e_hi = A.indptr[2 * wp1 + 2];
  const int k0 = valid ? A.indptr[2 * rp] : 0;
  const int k1 = valid ? A.indptr[2 * rp + 1] : 0;
  const int L = k1 - k0;
  double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
  if (e_hi - e_lo <= SPP_CAP) {
    for (int e = lane; e < e_hi - e_lo; e += 32) {
      soff[e] = A.indices[e_lo + e] * nb2;
      if (HAS2) sv2[e] = make_double2(A.v1[e_lo + e], A.v2[e_lo + e]);
      else sv1[e] = A.v1[e_lo + e];
    }
    __syncwarp();
    int k = k0 - e_lo;
    const int kend = k1 - e_lo;
    for (; k + 2 <= kend; k += 2) {
      const double2 x0 = xm[soff[k]], x1 = xm[soff[k + 1]];
      if (HAS2) {
        const double2 a0 = sv2[k], a1 = sv2[k + 1], b0 = sv2[k + L], b1 = sv2[k + 1 + L];
        ax = spp_acc(ax, cm.x, a0, x0.x);  ay = spp_acc(ay, cm.y, a0, x0.y);
        bx = spp_acc(bx, cm.x, b0, x0.x);  by = spp_acc(by, cm.y, b0, x0.y);
        ax = spp_acc(ax, cm.x, a1, x1.x);  ay = spp_acc(ay, cm.y, a1, x1.y);
        bx = spp_acc(bx, cm.x, b1, x1.x);  by = spp_acc(by, cm.y, b1, x1.y);
      } else {
        const double a0 = sv1[k], a1 = sv1[k + 1], b0 = sv1[k + L], b1 = sv1[k + 1 + L];
        ax += a0 * x0.x;  ay += a0 * x0.y;  bx += b0 * x0.x;  by += b0 * x0.y;
        ax += a1 * x1.x;  ay += a1 * x1.y;  bx += b1 * x1.x;  by += b1 * x1.y;
      }
    }
    for (; k < kend; ++k) {
      const double2 xv = xm[soff[k]];
      if (HAS2) {
        const double2 a = sv2[k], b = sv2[k + L];
        ax = spp_acc(ax, cm.x, a, xv.x);  ay = spp_acc(ay, cm.y, a, xv.y);
        bx = spp_acc(bx, cm.x, b, xv.x);  by = spp_acc(by, cm.y, b, xv.y);
      } else {
        const double a = sv1[k], b = sv1[k + L];
        ax += a * xv.x;  ay += a * xv.y;  bx += b * xv.x;  by += b * xv.y;
      }
    }
  } else {
    // rows too long for the staging area: straight from global memory
    for (int k = k0; k < k1; ++k) {
      const double2 xv = xm[(size_t)A.indices[k] * nb2];
      const double a1 = A.v1[k], b1 = A.v1[k + L];
      const double a2 = HAS2 ? A.v2[k] : 0.0, b2 = HAS2 ? A.v2[k + L] : 0.0;
      ax += (a1 + cm.x * a2) * xv.x;  ay += (a1 + cm.y * a2) * xv.y;
      bx += (b1 + cm.x * b2) * xv.x;  by += (b1 + cm.y * b2) * xv.y;
    }
  }
  accA = make_double2(ax, ay);
  accB = make_double2(bx, by);
}

template <bool HAS2>
__global__ void __launch_bounds__(SPB_THREADS)
k_spmm_p2(CsrDev A, const double *__restrict__ coef, const double2 *__restrict__ x,
          const double2 *z, double2 *y, int nb, int npairs, double alpha, double beta) {
  dnsb_pdl_entry();
  SPP_MAP(HAS2)
  const double2 cm = HAS2 ? reinterpret_cast<const double2 *>(coef)[mp] : make_double2(0.0, 0.0);
  const double2 zero = make_double2(0.0, 0.0);
  const double2 za = (beta != 0.0 && valid) ? z[ta_] : zero;
  const double2 zb = (beta != 0.0 && valid) ? z[tb_] : zero;
  double2 a, b;
  spp_rowdots<HAS2>(A, cm, x + mp, nb2, rp, valid, wp0, wp1, lane, sv2, sv1, soff, a, b);
  if (valid) {
    if (beta == 0.0) {
      y[ta_] = make_double2(alpha * a.x, alpha * a.y);
      y[tb_] = make_double2(alpha * b.x, alpha * b.y);
    } else {
      y[ta_] = make_double2(alpha * a.x + beta * za.x, alpha * a.y + beta * za.y);
      y[tb_] = make_double2(alpha * b.x + beta * zb.x, alpha * b.y + beta * zb.y);
    }
  }
}

__global__ void __launch_bounds__(SPB_THREADS)
k_cheb_init_p2(CsrDev A, const double2 *__restrict__ zp, const double2 *__restrict__ rv,
               const double2 *__restrict__ dinv, double2 *__restrict__ res,
               double2 *__restrict__ d, int nb, int npairs, double inv_theta) {
  dnsb_pdl_entry();
  SPP_MAP(false)
  const size_t sa_ = valid ? ta_ : 0, sb_ = valid ? tb_ : 0;
  const double2 ra = rv[sa_], rb = rv[sb_], da = dinv[sa_], db = dinv[sb_];
  double2 a, b;
  spp_rowdots<false>(A, make_double2(0.0, 0.0), zp + mp, nb2, rp, valid, wp0, wp1, lane, sv2, sv1,
                     soff, a, b);
  if (valid) {
    const double rax = ra.x - a.x, ray = ra.y - a.y, rbx = rb.x - b.x, rby = rb.y - b.y;
    res[ta_] = make_double2(rax, ray);
    res[tb_] = make_double2(rbx, rby);
    d[ta_] = make_double2(da.x * rax * inv_theta, da.y * ray * inv_theta);
    d[tb_] = make_double2(db.x * rbx * inv_theta, db.y * rby * inv_theta);
  }
}

// 64 registers -> 4 CTAs per SM: measured 34.6 us per full step (3 CTAs at 80
// registers: 37.9 us, 5 CTAs at 48 registers with spills: 41.0 us); the
// operands of the update other than `res` are loaded after the row product
// to stay within 64 registers
#ifndef SPP_MINB
#define SPP_MINB 4
#endif
template <bool HAS2, bool FIRST, bool LAST>
__global__ void __launch_bounds__(SPB_THREADS, SPP_MINB)
k_cheb_step_p2(CsrDev A, const double *__restrict__ coef, const double2 *__restrict__ d,
               const double2 *__restrict__ dinv, double2 *res, double2 *__restrict__ dn,
               double2 *z, int nb, int npairs, double c1, double c2) {
  dnsb_pdl_entry();
  SPP_MAP(HAS2)
  const double2 cm = HAS2 ? reinterpret_cast<const double2 *>(coef)[mp] : make_double2(0.0, 0.0);
  const size_t sa_ = valid ? ta_ : 0, sb_ = valid ? tb_ : 0;
  const double2 zero = make_double2(0.0, 0.0);
  const double2 ra = res[sa_], rb = res[sb_];
  double2 a, b;
  spp_rowdots<HAS2>(A, cm, d + mp, nb2, rp, valid, wp0, wp1, lane, sv2, sv1, soff, a, b);
  const double2 oa = d[sa_], ob = d[sb_];
  const double2 da = dinv[sa_], db = dinv[sb_];
  const double2 za = FIRST ? zero : z[sa_], zb = FIRST ? zero : z[sb_];
  if (valid) {
    const double rax = ra.x - a.x, ray = ra.y - a.y, rbx = rb.x - b.x, rby = rb.y - b.y;
    const double dax = spp_dir(c1, oa.x, c2, da.x, rax), day = spp_dir(c1, oa.y, c2, da.y, ray);
    const double dbx = spp_dir(c1, ob.x, c2, db.x, rbx), dby = spp_dir(c1, ob.y, c2, db.y, rby);
    if (!LAST) {
      res[ta_] = make_double2(rax, ray);
      res[tb_] = make_double2(rbx, rby);
      dn[ta_] = make_double2(dax, day);
      dn[tb_] = make_double2(dbx, dby);
    }
    z[ta_] = make_double2((FIRST ? oa.x : za.x) + dax, (FIRST ? oa.y : za.y) + day);
    z[tb_] = make_double2((FIRST ? ob.x : zb.x) + dbx, (FIRST ? ob.y : zb.y) + dby);
  }
}

// y = alpha*A*x + beta*z for a matrix whose first 2*npairs rows pair up and
// whose remaining rows do not (the saddle-point matrix K = [F JT; J 0]: the
// velocity rows pair, the divergence rows J are long single rows).  One
// launch: the first `tail_blocks` CTAs take the (long, latency-bound) tail rows
// so that they overlap with the bulk of the paired rows.
template <bool HAS2>
__global__ void __launch_bounds__(SPB_THREADS)
k_spmm_k2(CsrDev A, const double *__restrict__ coef, const double2 *__restrict__ x,
          const double2 *z, double2 *y, int nb, int npairs, int tail_blocks,
          double alpha, double beta) {
  dnsb_pdl_entry();
  const int nb2 = nb >> 1;
  const int lane = threadIdx.x & 31;
  const double2 zero = make_double2(0.0, 0.0);
  if ((int)blockIdx.x < tail_blocks) {
    const int row_begin = 2 * npairs;
    const long total_ = (long)(A.nrows - row_begin) * nb2;
    const long tl_ = (long)blockIdx.x * SPB_THREADS + threadIdx.x;
    if (tl_ - lane >= total_) return;
    const bool valid = tl_ < total_;
    const int row = valid ? row_begin + (int)((unsigned)tl_ / (unsigned)nb2) : A.nrows - 1;
    const int mp = valid ? (int)(tl_ - (long)(row - row_begin) * nb2) : 0;
    const int wrow0 = __shfl_sync(0xffffffffu, row, 0);
    const int wrow1 = __shfl_sync(0xffffffffu, row, 31);
    const long t_ = (long)row * nb2 + mp;
    SPB_SMEM(HAS2)
    const double2 cm = HAS2 ? reinterpret_cast<const double2 *>(coef)[mp] : zero;
    const double2 zin = (beta != 0.0 && valid) ? z[t_] : zero;
    const double2 acc = spb2_rowdot<HAS2>(A, cm, x + mp, nb2, row, valid, wrow0, wrow1, lane, sv2, sv1, soff);
    if (valid)
      y[t_] = (beta == 0.0) ? make_double2(alpha * acc.x, alpha * acc.y)
                            : make_double2(alpha * acc.x + beta * zin.x, alpha * acc.y + beta * zin.y);
    return;
  }
  const long total_ = (long)npairs * nb2;
  const long t_ = (long)(blockIdx.x - tail_blocks) * SPB_THREADS + threadIdx.x;
  if (t_ - lane >= total_) return;
  const bool valid = t_ < total_;
  const int rp = valid ? (int)((unsigned)t_ / (unsigned)nb2) : npairs - 1;
  const int mp = valid ? (int)(t_ - (long)rp * nb2) : 0;
  const int wp0 = __shfl_sync(0xffffffffu, rp, 0);
  const int wp1 = __shfl_sync(0xffffffffu, rp, 31);
  __shared__ typename SppSel<HAS2>::type smp_;
  double2 *pv2 = SppSel<HAS2>::v2(smp_, threadIdx.x >> 5);
  double *pv1 = SppSel<HAS2>::v1(smp_, threadIdx.x >> 5);
  int *poff = smp_.off[threadIdx.x >> 5];
  const size_t ta_ = (size_t)(2 * rp) * nb2 + mp, tb_ = ta_ + nb2;
  const double2 cm = HAS2 ? reinterpret_cast<const double2 *>(coef)[mp] : zero;
  const double2 za = (beta != 0.0 && valid) ? z[ta_] : zero;
  const double2 zb = (beta != 0.0 && valid) ? z[tb_] : zero;
  double2 a, b;
  spp_rowdots<HAS2>(A, cm, x + mp, nb2, rp, valid, wp0, wp1, lane, pv2, pv1, poff, a, b);
  if (valid) {
    if (beta == 0.0) {
      y[ta_] = make_double2(alpha * a.x, alpha * a.y);
      y[tb_] = make_double2(alpha * b.x, alpha * b.y);
    } else {
      y[ta_] = make_double2(alpha * a.x + beta * za.x, alpha * a.y + beta * za.y);
      y[tb_] = make_double2(alpha * b.x + beta * zb.x, alpha * b.y + beta * zb.y);
    }
  }
}

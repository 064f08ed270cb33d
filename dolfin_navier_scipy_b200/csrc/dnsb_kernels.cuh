// dnsb_kernels.cuh -- device kernels of libdnsb200 (sm_100a, fp64, int32 idx)
//
// All vectors are batched member-fastest: x[i*nb + m].  Every kernel is
// memory/latency bound: no tensor-core work on this path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

// ---------------------------------------------------------------------------
// P2 tabulation on the 7-point degree-5 rule (filled by dnsb_ctx_create)
// ---------------------------------------------------------------------------
__constant__ double c_phi[7][6];       // phi_a(q)
__constant__ double c_dphi[7][6][3];   // d phi_a / d lambda_i (q)
__constant__ double c_qw[7];           // weights, sum = 1 (times area)
__constant__ double c_lam[7][3];       // barycentric coordinates of the points (= P1 basis)

struct CsrDev {
  int nrows, ncols, nnz;
  const int *__restrict__ indptr;
  const int *__restrict__ indices;
  const double *__restrict__ v1;
  const double *__restrict__ v2;   // may be null
};

// ---------------------------------------------------------------------------
// K1a: convection vector  c_(n,a) += int (grad(u1) u2)_a phi_n   per cell.
// thread <-> (cell, member); cells [c0,c1) share one colour => plain RMW.
// Algorithmic bytes per (cell, member): 24 (ids, amortised over members)
// + 40 (geometry, amortised) + 96 gather + 192 RMW = 352 B  (SURVEY 8d).
// ---------------------------------------------------------------------------
template <bool SAME>
__global__ void __launch_bounds__(128)
k_convvec(int c0, int c1, int ncell, const int *__restrict__ cn,
          const double *__restrict__ geom, const double *__restrict__ u1,
          const double *__restrict__ u2, double *__restrict__ out, int nb) {
  dnsb_pdl_entry();
  long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int cell = c0 + (int)(tid / nb);
  int m = (int)(tid % nb);
  if (cell >= c1) return;
  int n[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) n[k] = cn[k * ncell + cell];
  const double g1x = geom[0 * ncell + cell], g1y = geom[1 * ncell + cell];
  const double g2x = geom[2 * ncell + cell], g2y = geom[3 * ncell + cell];
  const double detj = geom[4 * ncell + cell];
  const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
  double U1[6][2], U2[6][2];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    U1[k][0] = u1[(size_t)(2 * n[k]) * nb + m];
    U1[k][1] = u1[(size_t)(2 * n[k] + 1) * nb + m];
    if (SAME) {
      U2[k][0] = U1[k][0];
      U2[k][1] = U1[k][1];
    } else {
      U2[k][0] = u2[(size_t)(2 * n[k]) * nb + m];
      U2[k][1] = u2[(size_t)(2 * n[k] + 1) * nb + m];
    }
  }
  double acc[6][2];
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k][0] = acc[k][1] = 0.0;
  const double wdet = 0.5 * fabs(detj);
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    double ux = 0, uy = 0, dxx = 0, dxy = 0, dyx = 0, dyy = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const double gx = c_dphi[q][a][0] * g0x + c_dphi[q][a][1] * g1x +
                        c_dphi[q][a][2] * g2x;
      const double gy = c_dphi[q][a][0] * g0y + c_dphi[q][a][1] * g1y +
                        c_dphi[q][a][2] * g2y;
      ux += U2[a][0] * c_phi[q][a];
      uy += U2[a][1] * c_phi[q][a];
      dxx += U1[a][0] * gx;   // d_x u_x
      dxy += U1[a][0] * gy;   // d_y u_x
      dyx += U1[a][1] * gx;   // d_x u_y
      dyy += U1[a][1] * gy;   // d_y u_y
    }
    const double w = c_qw[q] * wdet;
    const double ax = w * (dxx * ux + dxy * uy);
    const double ay = w * (dyx * ux + dyy * uy);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      acc[a][0] += ax * c_phi[q][a];
      acc[a][1] += ay * c_phi[q][a];
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    out[(size_t)(2 * n[k]) * nb + m] += acc[k][0];
    out[(size_t)(2 * n[k] + 1) * nb + m] += acc[k][1];
  }
}

// ---------------------------------------------------------------------------
// K1a, gather formulation (two launches, no colours, no atomics):
//   k_conv_elem:   thread <-> (cell, member): element vector (6 nodes x 2
//                  components) -> E[((cell*6 + a)*2 + c)*nb + m]
//   k_conv_gather: thread <-> (output dof, member): sum of the element
//                  contributions of the incident cells in the fixed order of
//                  the node->(cell, local node) table  => deterministic.
// Same algorithmic bytes as the coloured scatter (the read-modify-write of the
// 12 dofs becomes a write + a read of E), but every cell is in flight at once
// instead of one colour (a tenth of the mesh) per launch.  The gather can
// restrict itself to the inner dofs and apply the sign of the explicit
// convection term: nfc[i] = -c[inv[i]] (stokes_navier_utils.py:1136-1140).
// ---------------------------------------------------------------------------
template <bool SAME>
__global__ void __launch_bounds__(128)
k_conv_elem(int ncell, const int *__restrict__ cn, const double *__restrict__ geom,
            const double *__restrict__ u1, const double *__restrict__ u2,
            double *__restrict__ E, int nb) {
  dnsb_pdl_entry();
  long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int cell = (int)(tid / nb);
  int m = (int)(tid % nb);
  if (cell >= ncell) return;
  int n[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) n[k] = cn[k * ncell + cell];
  const double g1x = geom[0 * ncell + cell], g1y = geom[1 * ncell + cell];
  const double g2x = geom[2 * ncell + cell], g2y = geom[3 * ncell + cell];
  const double detj = geom[4 * ncell + cell];
  const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
  double U1[6][2], U2[6][2];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    U1[k][0] = u1[(size_t)(2 * n[k]) * nb + m];
    U1[k][1] = u1[(size_t)(2 * n[k] + 1) * nb + m];
    if (SAME) {
      U2[k][0] = U1[k][0];
      U2[k][1] = U1[k][1];
    } else {
      U2[k][0] = u2[(size_t)(2 * n[k]) * nb + m];
      U2[k][1] = u2[(size_t)(2 * n[k] + 1) * nb + m];
    }
  }
  double acc[6][2];
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k][0] = acc[k][1] = 0.0;
  const double wdet = 0.5 * fabs(detj);
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    double ux = 0, uy = 0, dxx = 0, dxy = 0, dyx = 0, dyy = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const double gx = c_dphi[q][a][0] * g0x + c_dphi[q][a][1] * g1x +
                        c_dphi[q][a][2] * g2x;
      const double gy = c_dphi[q][a][0] * g0y + c_dphi[q][a][1] * g1y +
                        c_dphi[q][a][2] * g2y;
      ux += U2[a][0] * c_phi[q][a];
      uy += U2[a][1] * c_phi[q][a];
      dxx += U1[a][0] * gx;
      dxy += U1[a][0] * gy;
      dyx += U1[a][1] * gx;
      dyy += U1[a][1] * gy;
    }
    const double w = c_qw[q] * wdet;
    const double ax = w * (dxx * ux + dxy * uy);
    const double ay = w * (dyx * ux + dyy * uy);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      acc[a][0] += ax * c_phi[q][a];
      acc[a][1] += ay * c_phi[q][a];
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    E[((size_t)(cell * 6 + k) * 2 + 0) * nb + m] = acc[k][0];
    E[((size_t)(cell * 6 + k) * 2 + 1) * nb + m] = acc[k][1];
  }
}

// out[o, m] = sign * sum_{(cell,a) incident to node(dof)} E[((cell*6+a)*2 + comp)*nb + m]
//   dofs == null: o = dof = all 2*nnodes dofs;  else dof = dofs[o] (inner dofs)
__global__ void k_conv_gather(int nout, const int *__restrict__ dofs,
                              const int *__restrict__ n2c_ptr, const int *__restrict__ n2c_idx,
                              const double *__restrict__ E, double *__restrict__ out, int nb,
                              double sign) {
  dnsb_pdl_entry();
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nout * nb) return;
  const int o = (int)(t / nb), m = (int)(t % nb);
  const int dof = dofs ? dofs[o] : o;
  const int node = dof >> 1, comp = dof & 1;
  double s = 0.0;
  for (int k = n2c_ptr[node]; k < n2c_ptr[node + 1]; ++k)
    s += E[((size_t)n2c_idx[k] * 2 + comp) * nb + m];
  out[t] = sign * s;
}

// ---------------------------------------------------------------------------
// K1b: convection matrices into the fixed CSR pattern.  thread <-> (cell,
// local row n): 6 N1 entries (block diagonal, written for both components),
// 24 N2 entries, 2 f3 entries.  Cells of one colour own disjoint slots.
//   N1[(n,a),(m,a)] = int (u0.grad phi_m) phi_n
//   N2[(n,a),(m,b)] = int d_b u0_a phi_m phi_n
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_convmats(int c0, int c1, int ncell, const int *__restrict__ cn,
           const double *__restrict__ geom, const int *__restrict__ slots,
           const double *__restrict__ u0, double *__restrict__ n1,
           double *__restrict__ n2, double *__restrict__ f3) {
  dnsb_pdl_entry();
  long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int cell = c0 + (int)(tid / 6);
  int n = (int)(tid % 6);
  if (cell >= c1) return;
  int nd[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) nd[k] = cn[k * ncell + cell];
  const double g1x = geom[0 * ncell + cell], g1y = geom[1 * ncell + cell];
  const double g2x = geom[2 * ncell + cell], g2y = geom[3 * ncell + cell];
  const double detj = geom[4 * ncell + cell];
  const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
  double U[6][2];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    U[k][0] = u0[2 * nd[k]];
    U[k][1] = u0[2 * nd[k] + 1];
  }
  double a1[6], a2[2][6][2], fx = 0, fy = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    a1[k] = 0;
    a2[0][k][0] = a2[0][k][1] = a2[1][k][0] = a2[1][k][1] = 0;
  }
  const double wdet = 0.5 * fabs(detj);
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    double gx[6], gy[6];
    double ux = 0, uy = 0, dxx = 0, dxy = 0, dyx = 0, dyy = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      gx[a] = c_dphi[q][a][0] * g0x + c_dphi[q][a][1] * g1x +
              c_dphi[q][a][2] * g2x;
      gy[a] = c_dphi[q][a][0] * g0y + c_dphi[q][a][1] * g1y +
              c_dphi[q][a][2] * g2y;
      ux += U[a][0] * c_phi[q][a];
      uy += U[a][1] * c_phi[q][a];
      dxx += U[a][0] * gx[a];
      dxy += U[a][0] * gy[a];
      dyx += U[a][1] * gx[a];
      dyy += U[a][1] * gy[a];
    }
    // phi_n(q) for this thread's row (n is not a compile-time constant)
    const double pn = c_phi[q][n];
    const double w = c_qw[q] * wdet * pn;
    fx += w * (dxx * ux + dxy * uy);
    fy += w * (dyx * ux + dyy * uy);
#pragma unroll
    for (int mm = 0; mm < 6; ++mm) {
      a1[mm] += w * (ux * gx[mm] + uy * gy[mm]);
      const double wp = w * c_phi[q][mm];
      a2[0][mm][0] += wp * dxx;   // (a=x, b=x): d_x u_x
      a2[0][mm][1] += wp * dxy;   // (a=x, b=y): d_y u_x
      a2[1][mm][0] += wp * dyx;   // (a=y, b=x): d_x u_y
      a2[1][mm][1] += wp * dyy;   // (a=y, b=y): d_y u_y
    }
  }
  // slots[(r*12 + c)*ncell + cell], r = 2n+a, c = 2m+b
#pragma unroll
  for (int a = 0; a < 2; ++a) {
#pragma unroll
    for (int mm = 0; mm < 6; ++mm) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int s = slots[(size_t)((2 * n + a) * 12 + 2 * mm + b) * ncell + cell];
        if (n2) n2[s] += a2[a][mm][b];
        if (n1 && a == b) n1[s] += a1[mm];
      }
    }
  }
  if (f3) {
    f3[2 * nd[n]] += fx;
    f3[2 * nd[n] + 1] += fy;
  }
}

// ---------------------------------------------------------------------------
// K1b, gather formulation (two launches, no colours, no atomics):
//   k_convmats_elem:   thread <-> (local row n, cell), cells fastest (coalesced
//                      input): element matrices stored cell by cell,
//                      6 N1 entries -> EN1[cell*36 + n*6 + m],
//                      24 N2 entries -> EN2[cell*144 + (2n+a)*12 + 2m+b]
//                      (the slots of one CSR row read whole 96-byte rows of it)
//   k_convmats_gather: thread <-> CSR slot of the fixed pattern: sums the
//                      element contributions listed for the slot (ascending
//                      cell order => deterministic); N1 gets the (a == b) ones.
// f3 = N(u0)u0 is the convection vector: the K1a kernels compute it.
// ---------------------------------------------------------------------------
// CTA = CME_CELLS consecutive cells x 6 local rows; the element matrices are
// staged in shared memory (row stride padded to CME_LD doubles: conflict-free
// 16-byte stores) and written out as one contiguous block per CTA.
#define CME_CELLS 32
#define CME_LD 182   // 144 + 36 entries per cell, + 2: stride = 4 words mod 32 banks
__global__ void __launch_bounds__(6 * CME_CELLS, 2)
k_convmats_elem(int ncell, const int *__restrict__ cn, const double *__restrict__ geom,
                const double *__restrict__ u0, double *__restrict__ EN1,
                double *__restrict__ EN2) {
  dnsb_pdl_entry();
  __shared__ __align__(16) double sm[CME_CELLS * CME_LD];
  const int cl = threadIdx.x % CME_CELLS, n = threadIdx.x / CME_CELLS;
  const int cell0 = blockIdx.x * CME_CELLS;
  const int nc = min(CME_CELLS, ncell - cell0);
  const int cell = cell0 + min(cl, nc - 1);   // surplus threads repeat the last cell (no divergence)
  int nd[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) nd[k] = cn[k * ncell + cell];
  const double g1x = geom[0 * ncell + cell], g1y = geom[1 * ncell + cell];
  const double g2x = geom[2 * ncell + cell], g2y = geom[3 * ncell + cell];
  const double detj = geom[4 * ncell + cell];
  const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
  double U[6][2];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    U[k][0] = u0[2 * nd[k]];
    U[k][1] = u0[2 * nd[k] + 1];
  }
  double a1[6], a2[2][6][2];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    a1[k] = 0;
    a2[0][k][0] = a2[0][k][1] = a2[1][k][0] = a2[1][k][1] = 0;
  }
  const double wdet = 0.5 * fabs(detj);
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    double gx[6], gy[6];
    double ux = 0, uy = 0, dxx = 0, dxy = 0, dyx = 0, dyy = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      gx[a] = c_dphi[q][a][0] * g0x + c_dphi[q][a][1] * g1x + c_dphi[q][a][2] * g2x;
      gy[a] = c_dphi[q][a][0] * g0y + c_dphi[q][a][1] * g1y + c_dphi[q][a][2] * g2y;
      ux += U[a][0] * c_phi[q][a];
      uy += U[a][1] * c_phi[q][a];
      dxx += U[a][0] * gx[a];
      dxy += U[a][0] * gy[a];
      dyx += U[a][1] * gx[a];
      dyy += U[a][1] * gy[a];
    }
    const double w = c_qw[q] * wdet * c_phi[q][n];
#pragma unroll
    for (int mm = 0; mm < 6; ++mm) {
      a1[mm] += w * (ux * gx[mm] + uy * gy[mm]);
      const double wp = w * c_phi[q][mm];
      a2[0][mm][0] += wp * dxx;
      a2[0][mm][1] += wp * dxy;
      a2[1][mm][0] += wp * dyx;
      a2[1][mm][1] += wp * dyy;
    }
  }
  double *row = sm + cl * CME_LD;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int mm = 0; mm < 6; ++mm)
      *reinterpret_cast<double2 *>(row + (2 * n + a) * 12 + 2 * mm) = make_double2(a2[a][mm][0], a2[a][mm][1]);
#pragma unroll
  for (int mm = 0; mm < 6; mm += 2)
    *reinterpret_cast<double2 *>(row + 144 + n * 6 + mm) = make_double2(a1[mm], a1[mm + 1]);
  __syncthreads();
  double *o2 = EN2 + (size_t)cell0 * 144, *o1 = EN1 + (size_t)cell0 * 36;
  for (int i = threadIdx.x; i < nc * 144; i += 6 * CME_CELLS) o2[i] = sm[(i / 144) * CME_LD + i % 144];
  for (int i = threadIdx.x; i < nc * 36; i += 6 * CME_CELLS) o1[i] = sm[(i / 36) * CME_LD + 144 + i % 36];
}

// ---------------------------------------------------------------------------
// Constant operators of the Taylor-Hood discretisation, per cell (SURVEY 8f-1;
// reference dolfin_to_sparrays.py:236-275, get_stokessysmats): same CTA shape
// and staging as k_convmats_elem.  Thread = (cell, local row n):
//   EN1[cell*36 + n*6 + m]            = int phi_n phi_m                         (mass, per component)
//   EN2[cell*144 + (2n+a)*12 + 2m+b]  = nu (d_ab grad phi_n.grad phi_m + d_a phi_m d_b phi_n)   (2 eps(u):grad(v))
//                                       or 2 nu d_ab grad phi_n.grad phi_m      (symgrad == 0)
//   EJ[cell*36 + k*12 + 2m+b]         = int lambda_k d_b phi_m      (k = n < 3: divergence rows)
//   EMP[cell*9 + k*3 + l]             = int lambda_k lambda_l       (pressure mass)
// The P2 vector blocks are summed into the fixed pattern by k_convmats_gather
// (EN1 -> the a == b entries, like N1), EJ / EMP by k_slot_gather.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(6 * CME_CELLS, 2)
k_stokes_elem(int ncell, const double *__restrict__ geom, double nu, int symgrad,
              double *__restrict__ EN1, double *__restrict__ EN2, double *__restrict__ EJ,
              double *__restrict__ EMP) {
  dnsb_pdl_entry();
  __shared__ __align__(16) double sm[CME_CELLS * CME_LD];
  const int cl = threadIdx.x % CME_CELLS, n = threadIdx.x / CME_CELLS;
  const int cell0 = blockIdx.x * CME_CELLS;
  const int nc = min(CME_CELLS, ncell - cell0);
  const int cell = cell0 + min(cl, nc - 1);
  const double g1x = geom[0 * ncell + cell], g1y = geom[1 * ncell + cell];
  const double g2x = geom[2 * ncell + cell], g2y = geom[3 * ncell + cell];
  const double detj = geom[4 * ncell + cell];
  const double g0x = -(g1x + g2x), g0y = -(g1y + g2y);
  const double wdet = 0.5 * fabs(detj);
  double mrow[6], a2[2][6][2], jrow[6][2], mp[3];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    mrow[k] = 0;
    jrow[k][0] = jrow[k][1] = 0;
    a2[0][k][0] = a2[0][k][1] = a2[1][k][0] = a2[1][k][1] = 0;
  }
  mp[0] = mp[1] = mp[2] = 0;
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    double gx[6], gy[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      gx[a] = c_dphi[q][a][0] * g0x + c_dphi[q][a][1] * g1x + c_dphi[q][a][2] * g2x;
      gy[a] = c_dphi[q][a][0] * g0y + c_dphi[q][a][1] * g1y + c_dphi[q][a][2] * g2y;
    }
    const double w = c_qw[q] * wdet;
    const double wphi = w * c_phi[q][n];
    const double wl = n < 3 ? w * c_lam[q][n] : 0.0;
#pragma unroll
    for (int mm = 0; mm < 6; ++mm) {
      mrow[mm] += wphi * c_phi[q][mm];
      const double kk = w * (gx[n] * gx[mm] + gy[n] * gy[mm]);
      if (symgrad) {
        // test (n,a), trial (m,b): d_ab grad.grad + d_a phi_m d_b phi_n
        a2[0][mm][0] += kk + w * gx[mm] * gx[n];
        a2[0][mm][1] += w * gx[mm] * gy[n];
        a2[1][mm][0] += w * gy[mm] * gx[n];
        a2[1][mm][1] += kk + w * gy[mm] * gy[n];
      } else {
        a2[0][mm][0] += 2.0 * kk;
        a2[1][mm][1] += 2.0 * kk;
      }
      jrow[mm][0] += wl * gx[mm];
      jrow[mm][1] += wl * gy[mm];
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) mp[l] += wl * c_lam[q][l];
  }
  double *row = sm + cl * CME_LD;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int mm = 0; mm < 6; ++mm)
      *reinterpret_cast<double2 *>(row + (2 * n + a) * 12 + 2 * mm) =
          make_double2(nu * a2[a][mm][0], nu * a2[a][mm][1]);
#pragma unroll
  for (int mm = 0; mm < 6; mm += 2)
    *reinterpret_cast<double2 *>(row + 144 + n * 6 + mm) = make_double2(mrow[mm], mrow[mm + 1]);
  if (n < 3 && cl < nc) {
    if (EJ) {
#pragma unroll
      for (int mm = 0; mm < 6; ++mm)
        *reinterpret_cast<double2 *>(EJ + (size_t)cell * 36 + n * 12 + 2 * mm) =
            make_double2(jrow[mm][0], jrow[mm][1]);
    }
    if (EMP) {
#pragma unroll
      for (int l = 0; l < 3; ++l) EMP[(size_t)cell * 9 + n * 3 + l] = mp[l];
    }
  }
  __syncthreads();
  double *o2 = EN2 + (size_t)cell0 * 144, *o1 = EN1 + (size_t)cell0 * 36;
  for (int i = threadIdx.x; i < nc * 144; i += 6 * CME_CELLS) o2[i] = sm[(i / 144) * CME_LD + i % 144];
  for (int i = threadIdx.x; i < nc * 36; i += 6 * CME_CELLS) o1[i] = sm[(i / 36) * CME_LD + 144 + i % 36];
}

// out[s] = sum of the element entries listed for slot s (ascending cell order: deterministic)
__global__ void k_slot_gather(int nslots, const int *__restrict__ sptr, const int *__restrict__ ssrc,
                              const double *__restrict__ E, double *__restrict__ out) {
  dnsb_pdl_entry();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  double v = 0.0;
  for (int k = sptr[s]; k < sptr[s + 1]; ++k) v += E[ssrc[k]];
  out[s] = v;
}

// contributions of slot s: src = cell*144 + e, e = r*12 + c (r = 2n+a, c = 2m+b)
__global__ void k_convmats_gather(int nnz, int ncell, const int *__restrict__ sptr,
                                  const int *__restrict__ ssrc, const double *__restrict__ EN1,
                                  const double *__restrict__ EN2, double *__restrict__ n1,
                                  double *__restrict__ n2) {
  dnsb_pdl_entry();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nnz) return;
  double s1 = 0.0, s2 = 0.0;
  for (int k = sptr[s]; k < sptr[s + 1]; ++k) {
    const int src = ssrc[k];
    const int cell = src / 144, e = src - cell * 144;
    const int r = e / 12, c = e - r * 12;
    if (n2) s2 += EN2[src];
    if (n1 && ((r ^ c) & 1) == 0) s1 += EN1[(size_t)cell * 36 + (r >> 1) * 6 + (c >> 1)];
  }
  if (n1) n1[s] = s1;
  if (n2) n2[s] = s2;
}

// ---------------------------------------------------------------------------
// K2: CSR SpMV / SpMM.  LPR lanes cooperate on one (row, member) pair; LPR=1
// is thread per (row, member) -- the batched case, where the nb lanes of a row
// read one matrix entry (broadcast) and nb contiguous x values (coalesced).
// Algorithmic bytes: 12*nnz (+8*nnz with v2) + 4(nrows+1) + 8*nb*(ncols+nrows).
// ---------------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ double csr_rowdot(const CsrDev &A, double coefm,
                                             const double *__restrict__ x,
                                             int nb, int row, int m, int lane,
                                             bool valid) {
  double acc = 0.0;
  if (valid) {
    const int k0 = A.indptr[row], k1 = A.indptr[row + 1];
    if (A.v2) {
      for (int k = k0 + lane; k < k1; k += LPR)
        acc += (A.v1[k] + coefm * A.v2[k]) * x[(size_t)A.indices[k] * nb + m];
    } else {
      for (int k = k0 + lane; k < k1; k += LPR)
        acc += A.v1[k] * x[(size_t)A.indices[k] * nb + m];
    }
  }
  if (LPR > 1) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
      acc += __shfl_down_sync(0xffffffffu, acc, o, LPR);
  }
  return acc;
}

#define DNSB_ROWMAP(LPR)                                                      \
  const long tid_ = (long)blockIdx.x * blockDim.x + threadIdx.x;              \
  const long grp_ = tid_ / LPR;                                               \
  const int lane = (int)(tid_ % LPR);                                         \
  const int row = (int)(grp_ / nb);                                           \
  const int m = (int)(grp_ % nb);                                             \
  const bool valid = row < A.nrows;

// y = alpha*A*x + beta*z   (z may alias y; z ignored when beta == 0)
template <int LPR>
__global__ void
k_spmm(CsrDev A, const double *__restrict__ coef, const double *__restrict__ x,
       const double *z, double *y, int nb, double alpha, double beta) {
  dnsb_pdl_entry();
  DNSB_ROWMAP(LPR)
  const double cm = (valid && coef) ? coef[m] : 0.0;
  const double acc = csr_rowdot<LPR>(A, cm, x, nb, row, m, lane, valid);
  if (valid && lane == 0) {
    const size_t i = (size_t)row * nb + m;
    y[i] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * z[i];
  }
}

// Chebyshev start, fused with the gradient coupling:
//   res = rv - JT*zp ;  d = dinv*res/theta
// (z is not written: the first step forms z = d0 + d1; for a single step the
// caller passes d = z)
template <int LPR>
__global__ void
k_cheb_init(CsrDev A /*JT*/, const double *__restrict__ zp,
            const double *__restrict__ rv, const double *__restrict__ dinv,
            double *__restrict__ res, double *__restrict__ d, int nb,
            double inv_theta) {
  dnsb_pdl_entry();
  DNSB_ROWMAP(LPR)
  const double acc = csr_rowdot<LPR>(A, 0.0, zp, nb, row, m, lane, valid);
  if (valid && lane == 0) {
    const size_t i = (size_t)row * nb + m;
    const double r = rv[i] - acc;
    res[i] = r;
    d[i] = dinv[i] * r * inv_theta;
  }
}

// plain variant (no coupling): res = r ; d = dinv*r/theta   (res may alias r)
__global__ void k_cheb_init_plain(const double *r, const double *__restrict__ dinv,
                                  double *res, double *__restrict__ d, size_t n,
                                  double inv_theta) {
  dnsb_pdl_entry();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double rr = r[i];
  res[i] = rr;
  d[i] = dinv[i] * rr * inv_theta;
}

// Chebyshev step:  r = res - F*d ;  dn = c1*d + c2*dinv*r ;  z (+)= dn
//   FIRST: z = d + dn (z not read);  LAST: res and dn are not written
template <int LPR, bool FIRST, bool LAST>
__global__ void
k_cheb_step(CsrDev A /*F*/, const double *__restrict__ coef,
            const double *__restrict__ d, const double *__restrict__ dinv,
            double *res, double *__restrict__ dn, double *z, int nb, double c1,
            double c2) {
  dnsb_pdl_entry();
  DNSB_ROWMAP(LPR)
  const double cm = (valid && coef) ? coef[m] : 0.0;
  const double acc = csr_rowdot<LPR>(A, cm, d, nb, row, m, lane, valid);
  if (valid && lane == 0) {
    const size_t i = (size_t)row * nb + m;
    const double r = res[i] - acc;
    const double dold = d[i];
    const double dd = c1 * dold + c2 * dinv[i] * r;
    if (!LAST) {
      res[i] = r;
      dn[i] = dd;
    }
    z[i] = (FIRST ? dold : z[i]) + dd;
  }
}

// dinv[i,m] = 1/(v1[dp_i] + coef[m]*v2[dp_i])
__global__ void k_diag_inv(CsrDev A, const int *__restrict__ diagpos,
                           const double *__restrict__ coef,
                           double *__restrict__ dinv, int nb) {
  dnsb_pdl_entry();
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)A.nrows * nb) return;
  const int row = (int)(t / nb), m = (int)(t % nb);
  const int k = diagpos[row];
  double dv = A.v1[k];
  if (A.v2 && coef) dv += coef[m] * A.v2[k];
  dinv[t] = 1.0 / dv;
}

// ---------------------------------------------------------------------------
// dense  Y = alpha * D * X   (D: n x n row-major, X: n x nb)
// small nb: one warp per row, lanes split the columns (GEMV-like, D-bandwidth
// bound: 8*n*n bytes).
// ---------------------------------------------------------------------------
template <int NBMAX>
__global__ void
k_dense_gemv(const double *__restrict__ D, const double *__restrict__ X,
             double *__restrict__ Y, int n, int nb, double alpha,
             const double *__restrict__ add_dinv,
             const double *__restrict__ add_scale) {
  dnsb_pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  double acc[NBMAX];
#pragma unroll
  for (int m = 0; m < NBMAX; ++m) acc[m] = 0.0;
  const double *drow = D + (size_t)warp * n;
  for (int j = lane; j < n; j += 32) {
    const double dv = drow[j];
#pragma unroll
    for (int m = 0; m < NBMAX; ++m)
      if (m < nb) acc[m] += dv * X[(size_t)j * nb + m];
  }
#pragma unroll
  for (int m = 0; m < NBMAX; ++m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      acc[m] += __shfl_down_sync(0xffffffffu, acc[m], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m < NBMAX; ++m)
      if (m < nb) {
        double v = acc[m];
        if (add_dinv)
          v += add_scale[m] * add_dinv[warp] * X[(size_t)warp * nb + m];
        Y[(size_t)warp * nb + m] = alpha * v;
      }
  }
}

// ---------------------------------------------------------------------------
// batched reductions (deterministic two-stage: block partials, then a serial
// sum over blocks in fixed order).  Block = RPB x nb threads (RPB power of 2).
// ---------------------------------------------------------------------------
// partial[b*nb+m] = sum over chunk of x*y  (norms / single dots)
__global__ void k_dot1(const double *__restrict__ x, const double *__restrict__ y,
                       int n, int nb, int rpb, double *__restrict__ partial) {
  dnsb_pdl_entry();
  extern __shared__ double sred[];
  const int m = threadIdx.x % nb;
  const int rr = threadIdx.x / nb;
  const int rows_per_block = (n + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(n, r0 + rows_per_block);
  double acc = 0.0;
  for (int r = r0 + rr; r < r1; r += rpb) {
    const size_t idx = (size_t)r * nb + m;
    acc += x[idx] * y[idx];
  }
  sred[threadIdx.x] = acc;
  __syncthreads();
  for (int s = rpb >> 1; s > 0; s >>= 1) {
    if (rr < s) sred[threadIdx.x] += sred[threadIdx.x + s * nb];
    __syncthreads();
  }
  if (rr == 0) partial[(size_t)blockIdx.x * nb + m] = sred[m];
}

// ---------------------------------------------------------------------------
// FGMRES scalar work, one thread per member.  Hessenberg columns are stored
// rotated (R factor): R[(i*(mr) + j)*nb + m] for i <= j.
// ---------------------------------------------------------------------------
struct GmresState {
  double *R;        // (mr+1)*mr*nb
  double *cs, *sn;  // mr*nb
  double *g;        // (mr+1)*nb
  double *h;        // (mr+1)*nb  -- CGS coefficients of the current column
  double *invh;     // nb         -- 1/h_{j+1,j} (0 for finished members)
  double *bnorm;    // nb
  double *resid;    // nb         -- current |residual|
  int *done;        // nb
  int *its;         // nb  iterations in the current cycle
  int *ittot;       // nb  total iterations of the solve
  int *flags;       // [0]: number of members not yet done
  double *relmax;   // nb  -- running max over solves of the final |r|/|b| (NaN sticks)
  int mr;
};

// relmax[m] = max(relmax[m], rel) with NaN kept (a NaN residual must surface)
__device__ __forceinline__ void gmres_track(GmresState &S, int m, double res) {
  const double bn = S.bnorm[m];
  const double rel = bn > 0.0 ? res / bn : (res > 0.0 ? res : 0.0);
  const double old = S.relmax[m];
  if (old == old && !(rel <= old)) S.relmax[m] = rel;   // a NaN, once recorded, stays
}

// members that hit maxit without reaching tol: their last residual counts
__global__ void k_gmres_track_unconverged(GmresState S, int nb) {
  dnsb_pdl_entry();
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < nb && !S.done[m]) gmres_track(S, m, S.resid[m]);
}

// start of a cycle: beta = |r| from partials; V0 scale = 1/beta
__global__ void k_gmres_begin(GmresState S, const double *__restrict__ partial,
                              int nblocks, int nb, double tol, int first_cycle) {
  dnsb_pdl_entry();
  int m = blockIdx.x * blockDim.x + threadIdx.x;   // single block launch
  if (m < nb) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * nb + m];
    const double beta = sqrt(s);
    if (first_cycle) S.ittot[m] = 0;
    S.its[m] = 0;
    S.resid[m] = beta;
    S.g[m] = beta;
    const bool fin = !(beta > tol * S.bnorm[m]);   // also catches NaN -> done
    S.done[m] = fin ? 1 : 0;
    S.invh[m] = fin ? 0.0 : 1.0 / beta;
    if (fin) gmres_track(S, m, beta);
  }
  __syncthreads();
  if (m == 0) {
    int cnt = 0, mx = 0;
    for (int k = 0; k < nb; ++k) {
      cnt += S.done[k] ? 0 : 1;
      mx = max(mx, S.ittot[k]);
    }
    S.flags[0] = cnt;
    S.flags[1] = mx;
  }
}

// bnorm[m] = |b_m| from partials
__global__ void k_set_bnorm(GmresState S, const double *__restrict__ partial,
                            int nblocks, int nb) {
  dnsb_pdl_entry();
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nb) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * nb + m];
  S.bnorm[m] = sqrt(s);
}

// after CGS of column j: h[0..j] known, |w|^2 in partial2.  Apply the old
// rotations, build the new one, update g, test convergence.
// Launched as ONE block of 1024 threads: the block first sums the per-CTA
// partial norms (32 members x 32 slices at a time, fixed order), then thread m
// does the scalar work of member m.
__global__ void __launch_bounds__(1024)
k_gmres_givens(GmresState S, const double *__restrict__ partial2,
               int nblocks, int nb, int j, double tol, int pyth, int staged) {
  dnsb_pdl_entry();
  __shared__ double sp[32][33];
  __shared__ double snorm[1024];
  if (pyth) {
    // |w - V h|^2 = |w|^2 - |h|^2 (V orthonormal): h[0..j] and <w, w> = h[j+1] come from ONE pass
    // over the basis (k_mdot / k_gs_tma<false>), so the update kernel can write the normalised vector
    // directly and needs no reduction of its own.  The cancellation only bites when w lies in the
    // span already (|w_orth| < 1e-6 |w|), i.e. when the member has converged; the Arnoldi relation
    // K Z = V H stays exact whatever number is used as h_{j+1,j}, because V_{j+1} is scaled with it.
    // The difference is floored at 1e-8 |w|^2: above the floor its relative error is < 1e-7; below
    // it the vector is under-normalised, which makes the residual estimate pessimistic, never
    // optimistic (no false convergence).
    const int c = threadIdx.x;
    if (c < nb) {
      const double ww = S.h[(size_t)(j + 1) * nb + c];
      double s = ww;
      for (int i = 0; i <= j; ++i) { const double hi = S.h[(size_t)i * nb + c]; s -= hi * hi; }
      snorm[c] = fmax(s, 1e-8 * ww);
    }
    __syncthreads();
  } else {
    const int cx = threadIdx.x & 31, by = threadIdx.x >> 5;
    for (int m0 = 0; m0 < nb; m0 += 32) {
      const int c = m0 + cx;
      double s0 = 0.0, s1 = 0.0;
      if (c < nb) {
        int b = by;
        for (; b + 32 < nblocks; b += 64) {
          const double p0 = partial2[(size_t)b * nb + c];
          const double p1 = partial2[(size_t)(b + 32) * nb + c];
          s0 += p0; s1 += p1;
        }
        for (; b < nblocks; b += 32) s0 += partial2[(size_t)b * nb + c];
      }
      sp[by][cx] = s0 + s1;
      __syncthreads();
      if (by == 0 && c < nb) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 32; ++q) t += sp[q][cx];
        snorm[c] = t;
      }
      __syncthreads();
    }
  }
  int m = threadIdx.x;
  // the rotations and the new column of every member, staged once (coalesced over members, all
  // loads of the block in flight together) instead of 3 dependent global loads per (member, i)
  // (staged == 0: the arrays do not fit the shared-memory window -- read them in place)
  extern __shared__ double sgiv[];   // cs | sn | h : (j + 2) * nb each (h: j + 1 entries used)
  const double *scs = S.cs, *ssn = S.sn, *sh = S.h;
  if (staged) {
    double *wcs = sgiv, *wsn = sgiv + (size_t)(j + 2) * nb, *wh = wsn + (size_t)(j + 2) * nb;
    for (int t = threadIdx.x; t < j * nb; t += blockDim.x) { wcs[t] = S.cs[t]; wsn[t] = S.sn[t]; }
    for (int t = threadIdx.x; t < (j + 1) * nb; t += blockDim.x) wh[t] = S.h[t];
    scs = wcs; ssn = wsn; sh = wh;
  }
  __syncthreads();
  int active = 0, myit = 0;
  if (m < nb) {
    if (S.done[m]) {
      S.invh[m] = 0.0;
      myit = S.ittot[m];
    } else {
      const double s = snorm[m];
      const double hn = sqrt(s);
      const int mr = S.mr;
      double *__restrict__ pR = S.R;
      double hprev = sh[m];   // h[0]
#pragma unroll 4
      for (int i = 0; i < j; ++i) {
        const double c = scs[(size_t)i * nb + m], sn = ssn[(size_t)i * nb + m];
        const double hnext = sh[(size_t)(i + 1) * nb + m];
        const double t = c * hprev + sn * hnext;
        hprev = -sn * hprev + c * hnext;
        pR[((size_t)i * mr + j) * nb + m] = t;
      }
      const double dd = hypot(hprev, hn);
      double c = 1.0, sn = 0.0;
      if (dd > 0.0) { c = hprev / dd; sn = hn / dd; }
      S.cs[(size_t)j * nb + m] = c;
      S.sn[(size_t)j * nb + m] = sn;
      S.R[((size_t)j * mr + j) * nb + m] = dd;
      const double gj = S.g[(size_t)j * nb + m];
      S.g[(size_t)(j + 1) * nb + m] = -sn * gj;
      S.g[(size_t)j * nb + m] = c * gj;
      const double res = fabs(sn * gj);
      S.resid[m] = res;
      S.its[m] = j + 1;
      myit = S.ittot[m] + 1;
      S.ittot[m] = myit;
      const bool fin = !(res > tol * S.bnorm[m]) || !(hn > 0.0);
      S.done[m] = fin ? 1 : 0;
      S.invh[m] = fin ? 0.0 : 1.0 / hn;
      if (fin) gmres_track(S, m, res);
      active = fin ? 0 : 1;
    }
  }
  // members still iterating and the largest iteration count: block-wide, no serial loop over members
  __shared__ int s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  const int cnt = __syncthreads_count(active);
  if (m < nb) atomicMax(&s_max, myit);
  __syncthreads();
  if (threadIdx.x == 0) {
    S.flags[0] = cnt;
    S.flags[1] = s_max;   // iterations actually needed so far (max over members)
  }
}

// out[c] = sum_b partial[b*count + c]; block = 32 outputs x 32 slices of b,
// coalesced in c, 4 independent loads in flight per thread, fixed summation
// order (deterministic)
__global__ void __launch_bounds__(1024)
k_reduce_partials2(const double *__restrict__ partial, int nblocks, int count,
                   double *__restrict__ out) {
  dnsb_pdl_entry();
  __shared__ double sp[32][33];
  const int cx = threadIdx.x & 31, by = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (c < count) {
    int b = by;
    // 8 independent loads per round: one memory round trip covers 256 blocks
    for (; b + 224 < nblocks; b += 256) {
      const double p0 = partial[(size_t)b * count + c];
      const double p1 = partial[(size_t)(b + 32) * count + c];
      const double p2 = partial[(size_t)(b + 64) * count + c];
      const double p3 = partial[(size_t)(b + 96) * count + c];
      const double p4 = partial[(size_t)(b + 128) * count + c];
      const double p5 = partial[(size_t)(b + 160) * count + c];
      const double p6 = partial[(size_t)(b + 192) * count + c];
      const double p7 = partial[(size_t)(b + 224) * count + c];
      s0 += p0; s1 += p1; s2 += p2; s3 += p3;
      s0 += p4; s1 += p5; s2 += p6; s3 += p7;
    }
    for (; b + 96 < nblocks; b += 128) {
      const double p0 = partial[(size_t)b * count + c];
      const double p1 = partial[(size_t)(b + 32) * count + c];
      const double p2 = partial[(size_t)(b + 64) * count + c];
      const double p3 = partial[(size_t)(b + 96) * count + c];
      s0 += p0; s1 += p1; s2 += p2; s3 += p3;
    }
    for (; b < nblocks; b += 32) s0 += partial[(size_t)b * count + c];
  }
  sp[by][cx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (by == 0 && c < count) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 32; ++q) t += sp[q][cx];
    out[c] = t;
  }
}

// y = R^-1 g (per member, using its[m] columns); stored in S.h[i*nb+m]
__global__ void k_gmres_solve_y(GmresState S, int nb, int jmax) {
  dnsb_pdl_entry();
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nb) return;
  const int k = S.its[m];
  const int mr = S.mr;
  for (int i = jmax - 1; i >= 0; --i) {
    if (i >= k) {
      S.h[(size_t)i * nb + m] = 0.0;
      continue;
    }
    double s = S.g[(size_t)i * nb + m];
    for (int l = i + 1; l < k; ++l)
      s -= S.R[((size_t)i * mr + l) * nb + m] * S.h[(size_t)l * nb + m];
    S.h[(size_t)i * nb + m] = s / S.R[((size_t)i * mr + i) * nb + m];
  }
}

// x += sum_i y[i*nb+m] * Z_i
__global__ void k_gmres_update_x(const double *__restrict__ Z, size_t zstride,
                                 int nvec, const double *__restrict__ y,
                                 double *__restrict__ x, size_t n, int nb) {
  dnsb_pdl_entry();
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * nb) return;
  const int m = (int)(idx % nb);
  double s = x[idx];
  for (int i = 0; i < nvec; ++i)
    s += y[(size_t)i * nb + m] * Z[(size_t)i * zstride + idx];
  x[idx] = s;
}

// out = x * scale[m]
__global__ void k_scale_member(const double *x, const double *__restrict__ scale,
                               double *out, size_t n, int nb) {
  dnsb_pdl_entry();
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * nb) return;
  out[idx] = x[idx] * scale[idx % nb];
}

// ---------------------------------------------------------------------------
// elementwise helpers of the time stepper
// ---------------------------------------------------------------------------
// z = a*x + b*y  (y may be null)
__global__ void k_axpby(double a, const double *x, double b, const double *y,
                        double *z, size_t n) {
  dnsb_pdl_entry();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  z[i] = y ? a * x[i] + b * y[i] : a * x[i];
}

// vfull[inv[i], m] = v[i, m]
__global__ void k_scatter_inner(const double *__restrict__ v,
                                const int *__restrict__ inv,
                                double *__restrict__ vfull, int nv, int nb) {
  dnsb_pdl_entry();
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nv * nb) return;
  const int i = (int)(t / nb), m = (int)(t % nb);
  vfull[(size_t)inv[i] * nb + m] = v[t];
}

// vfull[bcinds[k], m] = bcvals[k]   (later entries win: serial per index list
// is resolved on the host, which passes unique indices)
__global__ void k_set_bcs(const int *__restrict__ bcinds,
                          const double *__restrict__ bcvals,
                          double *__restrict__ vfull, int nbc, int nb) {
  dnsb_pdl_entry();
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nbc * nb) return;
  const int k = (int)(t / nb), m = (int)(t % nb);
  vfull[(size_t)bcinds[k] * nb + m] = bcvals[k];
}

// nfc[i,m] = -cfull[inv[i], m]
__global__ void k_gather_neg(const double *__restrict__ cfull,
                             const int *__restrict__ inv,
                             double *__restrict__ nfc, int nv, int nb) {
  dnsb_pdl_entry();
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nv * nb) return;
  const int i = (int)(t / nb), m = (int)(t % nb);
  nfc[t] = -cfull[(size_t)inv[i] * nb + m];
}

// rhs[i,m] += ca*na[i,m] + cb*nb_[i,m] + cf*fv[i] + sum_k (wa*ua[k,m] + wb*ub[k,m])*B[i,k]
//   na/nb_: convection terms; fv: constant forcing (shared by members);
//   ua/ub: input samples at two time levels (nk x nb), B: nv x nk (row-major)
__global__ void k_rhs_combine(double *__restrict__ rhs,
                              const double *__restrict__ na, double ca,
                              const double *__restrict__ nbv, double cb,
                              const double *__restrict__ fv, double cf,
                              const double *__restrict__ B, int nk,
                              const double *__restrict__ ua, double wa,
                              const double *__restrict__ ub, double wb,
                              int nv, int nb) {
  dnsb_pdl_entry();
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nv * nb) return;
  const int i = (int)(t / nb), m = (int)(t % nb);
  double r = rhs[t];
  if (na) r += ca * na[t];
  if (nbv) r += cb * nbv[t];
  if (fv) r += cf * fv[i];
  for (int k = 0; k < nk; ++k) {
    double u = 0.0;
    if (ua) u += wa * ua[(size_t)k * nb + m];
    if (ub) u += wb * ub[(size_t)k * nb + m];
    r += u * B[(size_t)i * nk + k];
  }
  rhs[t] = r;
}

// b[(nv+j), m] = fp[j]  (pressure part of the saddle rhs, shared by members)
__global__ void k_fill_rhsp(double *__restrict__ b, const double *__restrict__ fp,
                            int nv, int np, int nb) {
  dnsb_pdl_entry();
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)np * nb) return;
  const int j = (int)(t / nb);
  b[(size_t)nv * nb + t] = fp[j];
}

// p[j,m] = scale * x[(nv+j), m]
__global__ void k_extract_p(const double *__restrict__ x, double *__restrict__ p,
                            int nv, int np, int nb, double scale) {
  dnsb_pdl_entry();
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)np * nb) return;
  p[t] = scale * x[(size_t)nv * nb + t];
}


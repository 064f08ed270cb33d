/* Oracle (TEST INFRASTRUCTURE): plain-C restatement of the cell loop behind
 * `dolfin.assemble(inner(grad(u)*u, v)*dx)` in
 * dolfin_navier_scipy/dolfin_to_sparrays.py:462-470 -- the reference runs this
 * loop in FFC-generated C++, so the CPU baseline of bench.py uses this compiled
 * version rather than the numpy oracle.
 * Single thread, 7-point degree-5 rule, P2 vector element.
 *
 *   cells  : ncell*6 scalar P2 node ids (3 vertices, 3 edge midpoints)
 *   coords : nvert*2 vertex coordinates,  cverts: ncell*3 vertex ids
 *   u      : 2*nnodes (interleaved), out: 2*nnodes (zeroed here)
 */
#include <math.h>
#include <string.h>

static void tabulate(double phi[7][6], double dphi[7][6][3], double w[7]) {
  const double s15 = sqrt(15.0);
  const double a1 = (6.0 - s15) / 21.0, a2 = (6.0 + s15) / 21.0;
  const double w1 = (155.0 - s15) / 1200.0, w2 = (155.0 + s15) / 1200.0;
  const double qp[7][3] = {{1. / 3, 1. / 3, 1. / 3},
                           {1 - 2 * a1, a1, a1}, {a1, 1 - 2 * a1, a1}, {a1, a1, 1 - 2 * a1},
                           {1 - 2 * a2, a2, a2}, {a2, 1 - 2 * a2, a2}, {a2, a2, 1 - 2 * a2}};
  const double ww[7] = {9. / 40, w1, w1, w1, w2, w2, w2};
  for (int q = 0; q < 7; ++q) {
    const double l0 = qp[q][0], l1 = qp[q][1], l2 = qp[q][2];
    w[q] = ww[q];
    phi[q][0] = l0 * (2 * l0 - 1); phi[q][1] = l1 * (2 * l1 - 1);
    phi[q][2] = l2 * (2 * l2 - 1); phi[q][3] = 4 * l1 * l2;
    phi[q][4] = 4 * l0 * l2; phi[q][5] = 4 * l0 * l1;
    memset(dphi[q], 0, sizeof dphi[q]);
    dphi[q][0][0] = 4 * l0 - 1; dphi[q][1][1] = 4 * l1 - 1; dphi[q][2][2] = 4 * l2 - 1;
    dphi[q][3][1] = 4 * l2; dphi[q][3][2] = 4 * l1;
    dphi[q][4][0] = 4 * l2; dphi[q][4][2] = 4 * l0;
    dphi[q][5][0] = 4 * l1; dphi[q][5][1] = 4 * l0;
  }
}

void oracle_convvec(int ncell, int nnodes, const int *cells, const int *cverts,
                    const double *coords, const double *u, double *out) {
  double phi[7][6], dphi[7][6][3], w[7];
  tabulate(phi, dphi, w);
  memset(out, 0, sizeof(double) * 2 * (size_t)nnodes);
  for (int c = 0; c < ncell; ++c) {
    const int *nd = cells + 6 * c;
    const double *x0 = coords + 2 * cverts[3 * c], *x1 = coords + 2 * cverts[3 * c + 1],
                 *x2 = coords + 2 * cverts[3 * c + 2];
    const double detj = (x1[0] - x0[0]) * (x2[1] - x0[1]) - (x2[0] - x0[0]) * (x1[1] - x0[1]);
    const double g[3][2] = {{(x1[1] - x2[1]) / detj, (x2[0] - x1[0]) / detj},
                            {(x2[1] - x0[1]) / detj, (x0[0] - x2[0]) / detj},
                            {(x0[1] - x1[1]) / detj, (x1[0] - x0[0]) / detj}};
    double U[6][2], acc[6][2];
    for (int a = 0; a < 6; ++a) {
      U[a][0] = u[2 * nd[a]]; U[a][1] = u[2 * nd[a] + 1];
      acc[a][0] = acc[a][1] = 0.0;
    }
    const double wdet = 0.5 * fabs(detj);
    for (int q = 0; q < 7; ++q) {
      double ux = 0, uy = 0, dxx = 0, dxy = 0, dyx = 0, dyy = 0;
      for (int a = 0; a < 6; ++a) {
        const double gx = dphi[q][a][0] * g[0][0] + dphi[q][a][1] * g[1][0] + dphi[q][a][2] * g[2][0];
        const double gy = dphi[q][a][0] * g[0][1] + dphi[q][a][1] * g[1][1] + dphi[q][a][2] * g[2][1];
        ux += U[a][0] * phi[q][a]; uy += U[a][1] * phi[q][a];
        dxx += U[a][0] * gx; dxy += U[a][0] * gy;
        dyx += U[a][1] * gx; dyy += U[a][1] * gy;
      }
      const double ax = w[q] * wdet * (dxx * ux + dxy * uy);
      const double ay = w[q] * wdet * (dyx * ux + dyy * uy);
      for (int a = 0; a < 6; ++a) { acc[a][0] += ax * phi[q][a]; acc[a][1] += ay * phi[q][a]; }
    }
    for (int a = 0; a < 6; ++a) { out[2 * nd[a]] += acc[a][0]; out[2 * nd[a] + 1] += acc[a][1]; }
  }
}

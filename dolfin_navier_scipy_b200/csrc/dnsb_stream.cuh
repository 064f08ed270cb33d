// dnsb_stream.cuh -- CSR SpMV/SpMM with the matrix stream staged through shared
// memory by the TMA engine (1-D bulk copies, cp.async.bulk + mbarrier), for
// operators that do not fit L2 (refined meshes).
//
// The row kernels of dnsb_kernels.cuh / dnsb_batched.cuh walk
// indptr -> (column, value) -> x with three dependent global loads per row; on
// short P2 rows (about 23 entries) that leaves too few bytes in flight for
// HBM3e (25-29 % of the copy peak at 0.5 M rows).  Here one producer thread per
// CTA streams the (indptr, column, value) slices of row tiles into a ring of
// shared-memory stages.  The consumer threads take the staged entries one per
// thread, whatever row they belong to (all gathers of a tile are independent
// and in flight together), leave the products in the stage and then sum them
// row by row (a few lanes per row).  Only the gather of x goes to L1/L2.
// Persistent CTAs, row tiles dealt round-robin, summation order fixed =>
// deterministic.
//
// replaces: scipy.sparse CSR matvec at reference time_int_utils.py:123-127,
// stokes_navier_utils.py:139-143 (same arithmetic as k_spmm).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define SPT_ROWS 128          // most rows per tile (runtime `rt`, multiple of 4: 16-byte aligned indptr slices)
#define SPT_CONSUMERS 256     // consumer threads per CTA (2-8 lanes per tile row) + 1 producer warp
#define SPT_THREADS (SPT_CONSUMERS + 32)
#define SPT_MAX_STAGES 4

struct SptPlan {
  int ntiles = 0;
  int rt = SPT_ROWS; // rows per tile
  int cap = 0;       // entries per stage (max tile span, multiple of 16)
  int stages = 0;    // 0: operator not eligible
  int ctas_per_sm = 1;
  size_t stage_bytes = 0, smem = 0;
};

__device__ __forceinline__ uint32_t spt_smem(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void spt_mbar_init(uint64_t *b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spt_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void spt_mbar_expect(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(spt_smem(b)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void spt_mbar_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(spt_smem(b)) : "memory");
}
__device__ __forceinline__ void spt_mbar_wait(uint64_t *b, uint32_t parity) {
  const uint32_t a = spt_smem(b);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void spt_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   spt_smem(dst)),
               "l"(src), "r"(bytes), "r"(spt_smem(bar))
               : "memory");
}

// what is done with the row product acc_i = (A x)_i  (one right-hand side):
//   SPT_AXPBY      y = alpha*acc + beta*z                       (k_spmm)
//   SPT_CHEB_INIT  r = a - acc ; o1 = r ; o2 = b*r*alpha        (k_cheb_init: a = rv, b = dinv,
//                                                                 o1 = res, o2 = d, alpha = 1/theta)
//   SPT_CHEB_STEP  r = o1 - acc ; dd = alpha*x_i + beta*b*r ;   (k_cheb_step: x = d, b = dinv, o1 = res,
//                  [o1 = r ; o2 = dd] ; o3 = (FIRST ? x_i : o3) + dd        o2 = dn, o3 = z, alpha = c1, beta = c2)
enum { SPT_AXPBY = 0, SPT_CHEB_INIT = 1, SPT_CHEB_STEP = 2, SPT_CHEB_STEP_FIRST = 3,
       SPT_CHEB_STEP_LAST = 4, SPT_CHEB_STEP_ONLY = 5 };
struct SptEpi {
  const double *a, *b;
  double *o1, *o2, *o3;
  double alpha, beta;
};

template <bool H2, int EPI>
__global__ void __launch_bounds__(SPT_THREADS)
k_spmv_tma(CsrDev A, const double *__restrict__ coef, const double *__restrict__ x, SptEpi ep,
           int ntiles, int cap, int stages, int rt) {
  dnsb_pdl_entry();
  constexpr int NWARP = SPT_CONSUMERS / 32;
  constexpr int UN = 6;
  extern __shared__ __align__(128) unsigned char spt_raw[];
  __shared__ __align__(8) uint64_t full[SPT_MAX_STAGES], empty[SPT_MAX_STAGES];

  const size_t off_v2 = (size_t)cap * 8;
  const size_t off_ci = off_v2 + (H2 ? (size_t)cap * 8 : 0);
  const size_t off_ip = off_ci + (size_t)cap * 4;
  const size_t stage_bytes = (off_ip + (SPT_ROWS + 4) * 4 + 127) & ~(size_t)127;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      spt_mbar_init(&full[s], 1);
      spt_mbar_init(&empty[s], NWARP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == NWARP) {
    // ---- producer: one thread streams the tiles of this CTA ----
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % stages;
        const int r0 = t * rt, r1 = min(A.nrows, r0 + rt);
        const int k0 = A.indptr[r0] & ~3, k1 = (A.indptr[r1] + 3) & ~3;
        const uint32_t ne = (uint32_t)(k1 - k0);
        const uint32_t nip = (uint32_t)((r1 - r0 + 1 + 3) & ~3);
        unsigned char *st = spt_raw + (size_t)s * stage_bytes;
        if (it >= stages) spt_mbar_wait(&empty[s], ((it / stages) - 1) & 1);
        spt_mbar_expect(&full[s], ne * (H2 ? 20u : 12u) + nip * 4u);
        spt_bulk_g2s(st + off_ip, A.indptr + r0, nip * 4u, &full[s]);
        if (ne) {
          spt_bulk_g2s(st + off_ci, A.indices + k0, ne * 4u, &full[s]);
          spt_bulk_g2s(st, A.v1 + k0, ne * 8u, &full[s]);
          if (H2) spt_bulk_g2s(st + off_v2, A.v2 + k0, ne * 8u, &full[s]);
        }
      }
    }
    return;
  }

  // ---- consumers ----
  const int tid = threadIdx.x;
  const int lsh = rt <= 32 ? 3 : (rt <= 64 ? 2 : 1);   // lanes per row = 1 << lsh
  const double cm = (H2 && coef) ? coef[0] : 0.0;
  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int s = it % stages;
    spt_mbar_wait(&full[s], (it / stages) & 1);
    unsigned char *st = spt_raw + (size_t)s * stage_bytes;
    double *sv1 = reinterpret_cast<double *>(st);
    const double *sv2 = reinterpret_cast<const double *>(st + off_v2);
    const int *sci = reinterpret_cast<const int *>(st + off_ci);
    const int *sip = reinterpret_cast<const int *>(st + off_ip);
    const int r0 = t * rt, nr = min(A.nrows - r0, rt);
    const int k0 = sip[0] & ~3;
    const int ne = ((sip[nr] + 3) & ~3) - k0;
    // phase 1: one entry per thread, products left in place of the values
    for (int k = tid; k < ne; k += UN * SPT_CONSUMERS) {
      double v[UN], xv[UN];
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        const int kk = min(k + j * SPT_CONSUMERS, ne - 1);
        v[j] = sv1[kk];
        if (H2) v[j] += cm * sv2[kk];
        xv[j] = x[sci[kk]];
      }
#pragma unroll
      for (int j = 0; j < UN; ++j)
        if (k + j * SPT_CONSUMERS < ne) sv1[k + j * SPT_CONSUMERS] = v[j] * xv[j];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SPT_CONSUMERS) : "memory");
    // phase 2: SPT_CONSUMERS / rt lanes per row (a power of two, 2..8)
    {
      const int lr = tid >> lsh, h = tid & ((1 << lsh) - 1);
      double acc = 0.0;
      if (lr < nr) {
        const int b = sip[lr] - k0, e = sip[lr + 1] - k0;
        for (int k = b + h; k < e; k += 1 << lsh) acc += sv1[k];
      }
      for (int o = 1; o < (1 << lsh); o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lr < nr && h == 0) {
        const size_t i = (size_t)(r0 + lr);
        if (EPI == SPT_AXPBY) {
          ep.o1[i] = (ep.beta == 0.0) ? ep.alpha * acc : ep.alpha * acc + ep.beta * ep.a[i];
        } else if (EPI == SPT_CHEB_INIT) {
          const double r = ep.a[i] - acc;
          ep.o1[i] = r;
          ep.o2[i] = ep.b[i] * r * ep.alpha;
        } else {
          constexpr bool FIRST = EPI == SPT_CHEB_STEP_FIRST || EPI == SPT_CHEB_STEP_ONLY;
          constexpr bool LAST = EPI == SPT_CHEB_STEP_LAST || EPI == SPT_CHEB_STEP_ONLY;
          const double r = ep.o1[i] - acc;
          const double dold = x[i];
          const double dd = ep.alpha * dold + ep.beta * ep.b[i] * r;
          if (!LAST) {
            ep.o1[i] = r;
            ep.o2[i] = dd;
          }
          ep.o3[i] = (FIRST ? dold : ep.o3[i]) + dd;
        }
      }
    }
    // the stage is handed back to the async proxy (next bulk copy overwrites the products)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) spt_mbar_arrive(&empty[s]);
  }
}

// ---------------------------------------------------------------------------
// Gram-Schmidt kernels of FGMRES with the basis vectors streamed by the TMA
// engine: same block/thread mapping, partial-sum layout and summation order as
// k_mdot_b / k_gs_update_b (dnsb_batched.cuh) -- bit-identical results -- but
// the chunk [r0, r1) x nb of every basis vector (one contiguous block of
// memory) arrives in a shared-memory ring through one bulk copy issued by a
// producer thread, so that the bytes in flight per SM no longer depend on the
// registers of the consumer threads (the register-pipelined kernels reach
// about half of the HBM peak at 24 % occupancy).
//   UPDATE = false:  partial[(b*(nvec+1) + i)*nb + m] = sum_chunk V_i*w, i = nvec: w*w
//   UPDATE = true:   vnext = w - sum_i h[i,m] V_i ; partial[b*nb + m] = |vnext|^2 over the chunk
// Block = rpb*nb consumer threads (thread = (rr, m)) + one producer warp;
// requires nb even (16-byte bulk copies) and rows_per_block <= rpb*GST_RPT.
// ---------------------------------------------------------------------------
#define GST_RPT 16
#define GST_H (GST_RPT / 2)
#define GST_MAX_STAGES 4

template <bool UPDATE>
__global__ void __launch_bounds__(288)
k_gs_tma(const double *__restrict__ V, size_t vstride, int nvec, const double *__restrict__ h,
         const double *__restrict__ w, double *__restrict__ vnext, int n, int nb, int rpb,
         int rows_per_block, double *__restrict__ partial, int stages, const double *__restrict__ scale) {
  dnsb_pdl_entry();
  extern __shared__ __align__(128) unsigned char gst_raw[];
  __shared__ __align__(8) uint64_t full[GST_MAX_STAGES], empty[GST_MAX_STAGES];
  const int nthr = rpb * nb;                 // consumer threads
  const int ncw = nthr >> 5;                 // consumer warps (nthr is a multiple of 32)
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(n, r0 + rows_per_block);
  const size_t tile_doubles = (size_t)rows_per_block * nb;
  double *ring = reinterpret_cast<double *>(gst_raw);
  double *sred = ring + (size_t)stages * tile_doubles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      spt_mbar_init(&full[s], 1);
      spt_mbar_init(&empty[s], ncw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if ((int)threadIdx.x >= nthr) {
    // ---- producer: one bulk copy per basis vector ----
    if ((int)threadIdx.x == nthr) {
      const uint32_t bytes = (uint32_t)((size_t)(r1 - r0) * nb * sizeof(double));
      for (int i = 0; i < nvec; ++i) {
        const int s = i % stages;
        if (i >= stages) spt_mbar_wait(&empty[s], ((i / stages) - 1) & 1);
        spt_mbar_expect(&full[s], bytes);
        spt_bulk_g2s(ring + (size_t)s * tile_doubles, V + (size_t)i * vstride + (size_t)r0 * nb, bytes,
                     &full[s]);
      }
    }
    return;
  }

  // ---- consumers ----
  const int m = threadIdx.x % nb, rr = threadIdx.x / nb, lane = threadIdx.x & 31;
  double wa[GST_H], wb[GST_H];
#pragma unroll
  for (int q = 0; q < GST_H; ++q) {
    const int ra = r0 + rr + q * rpb, rb = r0 + rr + (GST_H + q) * rpb;
    wa[q] = (ra < r1) ? w[(size_t)ra * nb + m] : 0.0;
    wb[q] = (rb < r1) ? w[(size_t)rb * nb + m] : 0.0;
  }
  double hi = (UPDATE && nvec > 0) ? h[m] : 0.0;
  for (int i = 0; i < nvec; ++i) {
    const int s = i % stages;
    const double hn = (UPDATE && i + 1 < nvec) ? h[(size_t)(i + 1) * nb + m] : 0.0;
    spt_mbar_wait(&full[s], (i / stages) & 1);
    const double *tile = ring + (size_t)s * tile_doubles + m;
    double va[GST_H], vb[GST_H];
#pragma unroll
    for (int q = 0; q < GST_H; ++q) {
      const int la = rr + q * rpb, lb = rr + (GST_H + q) * rpb;
      va[q] = (r0 + la < r1) ? tile[(size_t)la * nb] : 0.0;
      vb[q] = (r0 + lb < r1) ? tile[(size_t)lb * nb] : 0.0;
    }
    if (UPDATE) {
#pragma unroll
      for (int q = 0; q < GST_H; ++q) wa[q] -= hi * va[q];
#pragma unroll
      for (int q = 0; q < GST_H; ++q) wb[q] -= hi * vb[q];
      hi = hn;
    } else {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < GST_H; ++q) acc += va[q] * wa[q];
#pragma unroll
      for (int q = 0; q < GST_H; ++q) acc += vb[q] * wb[q];
      sred[(size_t)i * nthr + threadIdx.x] = acc;
    }
    // the values have been consumed (the shared-memory reads are complete): hand the stage back
    __syncwarp();
    if (lane == 0) spt_mbar_arrive(&empty[s]);
  }
  double ww = 0.0;
  if (UPDATE) {
    // scale != null: the norm is known already (Pythagoras, k_gmres_givens): the normalised
    // vector is written directly and no partial norms are needed
    const double sc = scale ? scale[m] : 1.0;
#pragma unroll
    for (int q = 0; q < GST_H; ++q) {
      const int ra = r0 + rr + q * rpb, rb = r0 + rr + (GST_H + q) * rpb;
      if (ra < r1) vnext[(size_t)ra * nb + m] = scale ? wa[q] * sc : wa[q];
      if (rb < r1) vnext[(size_t)rb * nb + m] = scale ? wb[q] * sc : wb[q];
    }
    if (scale) return;
  }
#pragma unroll
  for (int q = 0; q < GST_H; ++q) ww += wa[q] * wa[q];
#pragma unroll
  for (int q = 0; q < GST_H; ++q) ww += wb[q] * wb[q];
  sred[(size_t)(UPDATE ? 0 : nvec) * nthr + threadIdx.x] = ww;
  asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");   // consumers only
  if (UPDATE) {
    if (rr == 0) {
      double s = 0.0;
      for (int q = 0; q < rpb; ++q) s += sred[q * nb + m];
      partial[(size_t)blockIdx.x * nb + m] = s;
    }
  } else {
    for (int o = threadIdx.x; o < (nvec + 1) * nb; o += nthr) {
      const int i = o / nb, mm = o % nb;
      double s = 0.0;
      for (int q = 0; q < rpb; ++q) s += sred[(size_t)i * nthr + q * nb + mm];
      partial[((size_t)blockIdx.x * (nvec + 1) + i) * nb + mm] = s;
    }
  }
}

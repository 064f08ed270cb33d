// dnsb_tile.cuh -- batched Chebyshev step with EVERYTHING staged by the TMA
// engine (ensemble hot path, nb = 64 members).
//
// The row-pair kernel k_cheb_step_p2 (dnsb_batched.cuh) is bound by the L1 data
// pipe, not by DRAM (ncu: l1tex__data_pipe_lsu_wavefronts 59 % of peak, DRAM
// 29 %): per stored entry a warp gathers a 512-byte row of x with one LDG.128,
// which the LSU works off at ~2 cycles per 128-byte line, and the matrix
// entries, the update operands (res, d, dinv, z) and the outputs all go through
// the same pipe behind a chain of dependent global loads (indptr -> entries ->
// gather -> operands).  Here nothing but the output stores goes through the
// LSU's global path:
//
//   * persistent CTAs, one per SM; tiles of TILE_RP consecutive row pairs (P2
//     nodes in Hilbert order);
//   * a producer warp streams, per tile, into a 3-stage shared-memory ring with
//     1-D bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP): the packed matrix
//     entries of the tile (pair-interleaved values, tile-local gather offsets)
//     and the UNIQUE rows of x the tile
//     references (precomputed on the host as runs of consecutive columns: 8 bulk
//     copies per tile on cylinder_4 instead of 63 single rows -- the TMA engine
//     needs ~70 ns per copy) -- every x row is fetched from L2 once per tile
//     instead of once per row pair;
//   * 8 consumer warps, one row pair each: the update operands (res, d, dinv, z
//     of the warp's own two rows: coalesced, address known up front) are loaded
//     into registers BEFORE the warp waits for the stage, the gathers are LDS.128
//     from the x tile (1 cycle per 128 bytes, no dependent global load left),
//     2x2 register tile per lane as before.  Same summation order as k_cheb_step_p2:
//     bit-identical results (test_tiled_chebyshev_is_bit_identical).
//
// Algorithmic bytes per launch: 18*nnz (packed entries: 4 values + 1 offset per
// column of a row pair) + 8*n*nb*(7 - FIRST - 2*LAST).
//
// replaces: the velocity-block part of the sparse LU solve of
// time_int_utils.py:89-91,132 (as the smoother of the preconditioner).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "dnsb_batched.cuh"   // spp_acc / spp_dir: the shared, explicitly fused arithmetic

#define TILE_RP 8            // row pairs per tile = consumer warps
#define TILE_MAX_STAGES 8    // ring depth is a plan parameter (TilePlan::stages / stages_f), at most this
#define TILE_THREADS (TILE_RP * 32 + 32)
#define TILE_NB 64           // members (a warp = 32 member pairs)
#define TILE_ROWB (TILE_NB * 8)   // bytes of one vector row

struct TilePlan {
  int ntiles = 0, npairs = 0;
  int umax = 0;       // most unique x rows of a tile
  int cap = 0;        // most pair-entries of a tile (aligned span, multiple of 4)
  size_t stage_bytes = 0, smem = 0;       // fp64 tiles (k_cheb_step_tile, k_spmm_tile)
  size_t stage_bytes_f = 0, smem_f = 0;   // fp32 tiles (k_cheb_step_tilef)
  int stages = 0, stages_f = 0;
  bool ok = false;
};

struct TileDev {
  const int *indptr;        // CSR row pointers of the matrix (pair p: entries [indptr[2p]/2, indptr[2p+2]/2) packed)
  const int *uptr;          // ntiles + 1: cumulative number of unique x rows
  const int *rptr;          // ntiles + 1: runs of consecutive unique columns per tile
  const int *runs;          // per run: first column, number of columns, first slot in the x tile
  const int *pidx;          // per pair-entry: byte offset of the x row inside the tile
  const double *pval;       // per pair-entry: (a.v1, a.v2, b.v1, b.v2)
  const int4 *tdesc;        // per tile: first staged pair-entry (aligned), number staged, unique x rows, runs
  const int *truns;         // per tile `rmax` runs x (first column, columns, first slot), fixed stride
  const int4 *pdesc;        // per row pair: (offset into the staged entries, entries per row, 0, 0)
  int rmax;
  int ntiles, npairs, umax, cap;
  int stages;               // ring depth
};

__device__ __forceinline__ uint32_t tl_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tl_mbar_init(uint64_t *b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tl_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void tl_mbar_expect(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tl_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tl_mbar_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tl_smem(b)) : "memory");
}
__device__ __forceinline__ void tl_mbar_wait(uint64_t *b, uint32_t parity) {
  const uint32_t a = tl_smem(b);
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
__device__ __forceinline__ void tl_bulk(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tl_smem(dst)),
               "l"(src), "r"(bytes), "r"(tl_smem(bar))
               : "memory");
}

// Chebyshev step:  r = res - F*d ;  dn = c1*d + c2*dinv*r ;  z (+)= dn
//   FIRST: z = d + dn (z not read);  LAST: res and dn are not written
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(TILE_THREADS, 1)
k_cheb_step_tile(TileDev T, const double *__restrict__ coef, const double *__restrict__ d,
                 const double *__restrict__ dinv, double *res, double *__restrict__ dn, double *z,
                 double c1, double c2) {
  extern __shared__ __align__(128) unsigned char tl_raw[];
  __shared__ __align__(8) uint64_t full[TILE_MAX_STAGES], empty[TILE_MAX_STAGES];
  const int nst = T.stages;
  // stage: x tile (umax rows) | packed values (cap x 32 B) | gather offsets (cap x 4 B)
  const size_t off_val = (size_t)T.umax * TILE_ROWB;
  const size_t off_idx = off_val + (size_t)T.cap * 32;
  const size_t stage_bytes = (off_idx + (size_t)T.cap * 4 + 127) & ~(size_t)127;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      tl_mbar_init(&full[s], 1);
      tl_mbar_init(&empty[s], TILE_RP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // barrier set-up and the first (static) tile descriptor are taken before the programmatic-launch wait
  if (warp != TILE_RP) dnsb_pdl_entry();
  if (warp == TILE_RP) {
    // ---- producer warp: the descriptor and the runs of the NEXT tile are loaded (one independent
    // round trip each, fixed addresses) while the copies of the current one are issued -- the loop
    // used to pay three dependent global round trips per tile (indptr -> run list -> addresses), which
    // bounded the whole kernel at ~2 us per tile whatever the number of bytes
    int it = 0;
    int t = blockIdx.x;
    int4 dsc = make_int4(0, 0, 0, 0);
    int rc = 0, rl = 0, rs = 0;
    if (t < T.ntiles) {
      dsc = T.tdesc[t];
      if (lane < T.rmax) {
        const int *rr = T.truns + ((size_t)t * T.rmax + lane) * 3;
        rc = rr[0]; rl = rr[1]; rs = rr[2];
      }
    }
    dnsb_pdl_entry();
    for (; t < T.ntiles; t += gridDim.x, ++it) {
      const int s = it % nst;
      const int tn = t + gridDim.x;
      int4 ndsc = dsc;
      int nrc = 0, nrl = 0, nrs = 0;
      if (tn < T.ntiles) {
        ndsc = T.tdesc[tn];
        if (lane < T.rmax) {
          const int *rr = T.truns + ((size_t)tn * T.rmax + lane) * 3;
          nrc = rr[0]; nrl = rr[1]; nrs = rr[2];
        }
      }
      if (it >= nst) {
        if (lane == 0) tl_mbar_wait(&empty[s], ((it / nst) - 1) & 1);
        __syncwarp();
      }
      unsigned char *st = tl_raw + (size_t)s * stage_bytes;
      if (lane == 0) {
        tl_mbar_expect(&full[s], (uint32_t)dsc.z * TILE_ROWB + (uint32_t)dsc.y * 36);
        tl_bulk(st + off_val, T.pval + (size_t)dsc.x * 4, (uint32_t)dsc.y * 32, &full[s]);
        tl_bulk(st + off_idx, T.pidx + dsc.x, (uint32_t)dsc.y * 4, &full[s]);
      }
      __syncwarp();
      if (lane < dsc.w)
        tl_bulk(st + (size_t)rs * TILE_ROWB, d + (size_t)rc * TILE_NB, (uint32_t)rl * TILE_ROWB, &full[s]);
      dsc = ndsc; rc = nrc; rl = nrl; rs = nrs;
    }
    return;
  }

  // ---- consumer warps: warp w owns row pair p0 + w of every tile ----
  const double2 cm = reinterpret_cast<const double2 *>(coef)[lane];
  const double2 zero = make_double2(0.0, 0.0);
  // update operands of the warp's own two rows (coalesced, addresses known up front): loaded
  // ONE TILE AHEAD into registers, so that their DRAM latency overlaps the previous tile
  struct Ops { double2 ra, rb, oa, ob, da, db, za, zb; int4 pd; };
  auto load_ops = [&](int tile) {
    Ops o;
    const int q = min(tile * TILE_RP + warp, T.npairs - 1);
    o.pd = T.pdesc[q];
    const size_t ia = (size_t)(2 * q) * (TILE_NB / 2) + lane, ib = ia + TILE_NB / 2;
    o.ra = reinterpret_cast<const double2 *>(res)[ia];  o.rb = reinterpret_cast<const double2 *>(res)[ib];
    o.oa = reinterpret_cast<const double2 *>(d)[ia];    o.ob = reinterpret_cast<const double2 *>(d)[ib];
    o.da = reinterpret_cast<const double2 *>(dinv)[ia]; o.db = reinterpret_cast<const double2 *>(dinv)[ib];
    o.za = FIRST ? zero : reinterpret_cast<const double2 *>(z)[ia];
    o.zb = FIRST ? zero : reinterpret_cast<const double2 *>(z)[ib];
    return o;
  };
  Ops cur = load_ops(min((int)blockIdx.x, T.ntiles - 1));
  int it = 0;
  for (int t = blockIdx.x; t < T.ntiles; t += gridDim.x, ++it) {
    const int s = it % nst;
    const int p0 = t * TILE_RP;
    const int rp = p0 + warp;
    const bool have = rp < T.npairs;
    const int kb = cur.pd.x, L = have ? cur.pd.y : 0;   // came with the operands, one tile ahead
    const size_t ta = (size_t)(2 * (have ? rp : 0)) * (TILE_NB / 2) + lane, tb = ta + TILE_NB / 2;
    const int tn = t + gridDim.x;
    Ops nxt = cur;
    if (tn < T.ntiles) nxt = load_ops(tn);
    tl_mbar_wait(&full[s], (it / nst) & 1);
    const unsigned char *st = tl_raw + (size_t)s * stage_bytes;
    if (have) {
      const double2 *sval = reinterpret_cast<const double2 *>(st + off_val) + (size_t)kb * 2;
      const int *sidx = reinterpret_cast<const int *>(st + off_idx) + kb;
      const unsigned char *xt = st + (size_t)lane * 16;
      double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
      int k = 0;
      for (; k + 2 <= L; k += 2) {
        const int o0 = sidx[k], o1 = sidx[k + 1];
        const double2 x0 = *reinterpret_cast<const double2 *>(xt + o0);
        const double2 x1 = *reinterpret_cast<const double2 *>(xt + o1);
        const double2 a0 = sval[2 * k], b0 = sval[2 * k + 1], a1 = sval[2 * k + 2], b1 = sval[2 * k + 3];
        ax = spp_acc(ax, cm.x, a0, x0.x);  ay = spp_acc(ay, cm.y, a0, x0.y);
        bx = spp_acc(bx, cm.x, b0, x0.x);  by = spp_acc(by, cm.y, b0, x0.y);
        ax = spp_acc(ax, cm.x, a1, x1.x);  ay = spp_acc(ay, cm.y, a1, x1.y);
        bx = spp_acc(bx, cm.x, b1, x1.x);  by = spp_acc(by, cm.y, b1, x1.y);
      }
      for (; k < L; ++k) {
        const double2 xv = *reinterpret_cast<const double2 *>(xt + sidx[k]);
        const double2 a = sval[2 * k], b = sval[2 * k + 1];
        ax = spp_acc(ax, cm.x, a, xv.x);  ay = spp_acc(ay, cm.y, a, xv.y);
        bx = spp_acc(bx, cm.x, b, xv.x);  by = spp_acc(by, cm.y, b, xv.y);
      }
      // the stage has been read: hand it back before the update and the stores
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(&empty[s]);
      const double rax = cur.ra.x - ax, ray = cur.ra.y - ay, rbx = cur.rb.x - bx, rby = cur.rb.y - by;
      const double dax = spp_dir(c1, cur.oa.x, c2, cur.da.x, rax), day = spp_dir(c1, cur.oa.y, c2, cur.da.y, ray);
      const double dbx = spp_dir(c1, cur.ob.x, c2, cur.db.x, rbx), dby = spp_dir(c1, cur.ob.y, c2, cur.db.y, rby);
      if (!LAST) {
        reinterpret_cast<double2 *>(res)[ta] = make_double2(rax, ray);
        reinterpret_cast<double2 *>(res)[tb] = make_double2(rbx, rby);
        reinterpret_cast<double2 *>(dn)[ta] = make_double2(dax, day);
        reinterpret_cast<double2 *>(dn)[tb] = make_double2(dbx, dby);
      }
      reinterpret_cast<double2 *>(z)[ta] =
          make_double2((FIRST ? cur.oa.x : cur.za.x) + dax, (FIRST ? cur.oa.y : cur.za.y) + day);
      reinterpret_cast<double2 *>(z)[tb] =
          make_double2((FIRST ? cur.ob.x : cur.zb.x) + dbx, (FIRST ? cur.ob.y : cur.zb.y) + dby);
    } else {
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(&empty[s]);
    }
    cur = nxt;
  }
}

// ---------------------------------------------------------------------------
// fp32 variant of the same kernel for the SMOOTHER INSIDE THE PRECONDITIONER.
// The Jacobi-Chebyshev iteration on the velocity block is a polynomial
// approximation of F^-1 that is accurate to a few per cent; it is applied inside
// a flexible GMRES whose bases, residuals and stopping test are fp64 (like the
// tensor-core Schur block, dnsb_tc.cuh).  Carrying its work vectors (res, d, z,
// dinv) and the matrix values in fp32 halves every byte the kernel moves -- DRAM
// operands, the L2 refetch of the x tile, the shared-memory gathers -- and the
// model of the solver (tools/solver_model.py) shows unchanged iteration counts.
// The last step writes z in fp64 (the preconditioned vector Z_j of FGMRES).
//   rows of x: 256 bytes (64 floats); packed entry: 4 floats + byte offset.
// ---------------------------------------------------------------------------
#define TILE_ROWBF (TILE_NB * 4)

struct TileDevF {
  const int *indptr, *uptr, *rptr, *runs;
  const int *pidx;          // byte offset of the x row inside the fp32 tile
  const float *pval;        // (a.v1, a.v2, b.v1, b.v2) as floats
  const int4 *tdesc;
  const int *truns;
  const int4 *pdesc;
  int rmax;
  int ntiles, npairs, umax, cap;
  int stages;               // ring depth
};

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(TILE_THREADS, 2)   // two CTAs per SM fit (registers, fp32 ring): see g_tilef_ctas
k_cheb_step_tilef(TileDevF T, const double *__restrict__ coef, const float *__restrict__ d,
                  const float *__restrict__ dinv, float *res, float *__restrict__ dn, float *zf,
                  double *__restrict__ zout, float c1, float c2) {
  extern __shared__ __align__(128) unsigned char tl_raw[];
  __shared__ __align__(8) uint64_t full[TILE_MAX_STAGES], empty[TILE_MAX_STAGES];
  const int nst = T.stages;
  const size_t off_val = (size_t)T.umax * TILE_ROWBF;
  const size_t off_idx = off_val + (size_t)T.cap * 16;
  const size_t stage_bytes = (off_idx + (size_t)T.cap * 4 + 127) & ~(size_t)127;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      tl_mbar_init(&full[s], 1);
      tl_mbar_init(&empty[s], TILE_RP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // barrier set-up and the first (static) tile descriptor are taken before the programmatic-launch wait
  if (warp != TILE_RP) dnsb_pdl_entry();
  if (warp == TILE_RP) {
    // ---- producer warp: the descriptor and the runs of the NEXT tile are loaded (one independent
    // round trip each, fixed addresses) while the copies of the current one are issued -- the loop
    // used to pay three dependent global round trips per tile (indptr -> run list -> addresses), which
    // bounded the whole kernel at ~2 us per tile whatever the number of bytes
    int it = 0;
    int t = blockIdx.x;
    int4 dsc = make_int4(0, 0, 0, 0);
    int rc = 0, rl = 0, rs = 0;
    if (t < T.ntiles) {
      dsc = T.tdesc[t];
      if (lane < T.rmax) {
        const int *rr = T.truns + ((size_t)t * T.rmax + lane) * 3;
        rc = rr[0]; rl = rr[1]; rs = rr[2];
      }
    }
    dnsb_pdl_entry();
    for (; t < T.ntiles; t += gridDim.x, ++it) {
      const int s = it % nst;
      const int tn = t + gridDim.x;
      int4 ndsc = dsc;
      int nrc = 0, nrl = 0, nrs = 0;
      if (tn < T.ntiles) {
        ndsc = T.tdesc[tn];
        if (lane < T.rmax) {
          const int *rr = T.truns + ((size_t)tn * T.rmax + lane) * 3;
          nrc = rr[0]; nrl = rr[1]; nrs = rr[2];
        }
      }
      if (it >= nst) {
        if (lane == 0) tl_mbar_wait(&empty[s], ((it / nst) - 1) & 1);
        __syncwarp();
      }
      unsigned char *st = tl_raw + (size_t)s * stage_bytes;
      if (lane == 0) {
        tl_mbar_expect(&full[s], (uint32_t)dsc.z * TILE_ROWBF + (uint32_t)dsc.y * 20);
        tl_bulk(st + off_val, T.pval + (size_t)dsc.x * 4, (uint32_t)dsc.y * 16, &full[s]);
        tl_bulk(st + off_idx, T.pidx + dsc.x, (uint32_t)dsc.y * 4, &full[s]);
      }
      __syncwarp();
      if (lane < dsc.w)
        tl_bulk(st + (size_t)rs * TILE_ROWBF, d + (size_t)rc * TILE_NB, (uint32_t)rl * TILE_ROWBF, &full[s]);
      dsc = ndsc; rc = nrc; rl = nrl; rs = nrs;
    }
    return;
  }

  const double2 cmd = reinterpret_cast<const double2 *>(coef)[lane];
  const float2 cm = make_float2((float)cmd.x, (float)cmd.y);
  const float2 zero = make_float2(0.f, 0.f);
  struct Ops { float2 ra, rb, oa, ob, da, db, za, zb; int4 pd; };
  auto load_ops = [&](int tile) {
    Ops o;
    const int q = min(tile * TILE_RP + warp, T.npairs - 1);
    o.pd = T.pdesc[q];
    const size_t ia = (size_t)(2 * q) * (TILE_NB / 2) + lane, ib = ia + TILE_NB / 2;
    o.ra = reinterpret_cast<const float2 *>(res)[ia];  o.rb = reinterpret_cast<const float2 *>(res)[ib];
    o.oa = reinterpret_cast<const float2 *>(d)[ia];    o.ob = reinterpret_cast<const float2 *>(d)[ib];
    o.da = reinterpret_cast<const float2 *>(dinv)[ia]; o.db = reinterpret_cast<const float2 *>(dinv)[ib];
    o.za = FIRST ? zero : reinterpret_cast<const float2 *>(zf)[ia];
    o.zb = FIRST ? zero : reinterpret_cast<const float2 *>(zf)[ib];
    return o;
  };
  Ops cur = load_ops(min((int)blockIdx.x, T.ntiles - 1));
  int it = 0;
  for (int t = blockIdx.x; t < T.ntiles; t += gridDim.x, ++it) {
    const int s = it % nst;
    const int p0 = t * TILE_RP;
    const int rp = p0 + warp;
    const bool have = rp < T.npairs;
    const int kb = cur.pd.x, L = have ? cur.pd.y : 0;   // came with the operands, one tile ahead
    const size_t ta = (size_t)(2 * (have ? rp : 0)) * (TILE_NB / 2) + lane, tb = ta + TILE_NB / 2;
    const int tn = t + gridDim.x;
    Ops nxt = cur;
    if (tn < T.ntiles) nxt = load_ops(tn);
    tl_mbar_wait(&full[s], (it / nst) & 1);
    const unsigned char *st = tl_raw + (size_t)s * stage_bytes;
    if (have) {
      const float4 *sval = reinterpret_cast<const float4 *>(st + off_val) + kb;
      const int *sidx = reinterpret_cast<const int *>(st + off_idx) + kb;
      const unsigned char *xt = st + (size_t)lane * 8;
      float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;
      int k = 0;
      for (; k + 2 <= L; k += 2) {
        const int o0 = sidx[k], o1 = sidx[k + 1];
        const float2 x0 = *reinterpret_cast<const float2 *>(xt + o0);
        const float2 x1 = *reinterpret_cast<const float2 *>(xt + o1);
        const float4 e0 = sval[k], e1 = sval[k + 1];
        ax = fmaf(fmaf(cm.x, e0.y, e0.x), x0.x, ax);  ay = fmaf(fmaf(cm.y, e0.y, e0.x), x0.y, ay);
        bx = fmaf(fmaf(cm.x, e0.w, e0.z), x0.x, bx);  by = fmaf(fmaf(cm.y, e0.w, e0.z), x0.y, by);
        ax = fmaf(fmaf(cm.x, e1.y, e1.x), x1.x, ax);  ay = fmaf(fmaf(cm.y, e1.y, e1.x), x1.y, ay);
        bx = fmaf(fmaf(cm.x, e1.w, e1.z), x1.x, bx);  by = fmaf(fmaf(cm.y, e1.w, e1.z), x1.y, by);
      }
      for (; k < L; ++k) {
        const float2 xv = *reinterpret_cast<const float2 *>(xt + sidx[k]);
        const float4 e = sval[k];
        ax = fmaf(fmaf(cm.x, e.y, e.x), xv.x, ax);  ay = fmaf(fmaf(cm.y, e.y, e.x), xv.y, ay);
        bx = fmaf(fmaf(cm.x, e.w, e.z), xv.x, bx);  by = fmaf(fmaf(cm.y, e.w, e.z), xv.y, by);
      }
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(&empty[s]);
      const float rax = cur.ra.x - ax, ray = cur.ra.y - ay, rbx = cur.rb.x - bx, rby = cur.rb.y - by;
      const float dax = fmaf(c2 * cur.da.x, rax, c1 * cur.oa.x), day = fmaf(c2 * cur.da.y, ray, c1 * cur.oa.y);
      const float dbx = fmaf(c2 * cur.db.x, rbx, c1 * cur.ob.x), dby = fmaf(c2 * cur.db.y, rby, c1 * cur.ob.y);
      if (!LAST) {
        reinterpret_cast<float2 *>(res)[ta] = make_float2(rax, ray);
        reinterpret_cast<float2 *>(res)[tb] = make_float2(rbx, rby);
        reinterpret_cast<float2 *>(dn)[ta] = make_float2(dax, day);
        reinterpret_cast<float2 *>(dn)[tb] = make_float2(dbx, dby);
      }
      const float zax = (FIRST ? cur.oa.x : cur.za.x) + dax, zay = (FIRST ? cur.oa.y : cur.za.y) + day;
      const float zbx = (FIRST ? cur.ob.x : cur.zb.x) + dbx, zby = (FIRST ? cur.ob.y : cur.zb.y) + dby;
      if (LAST) {
        reinterpret_cast<double2 *>(zout)[ta] = make_double2((double)zax, (double)zay);
        reinterpret_cast<double2 *>(zout)[tb] = make_double2((double)zbx, (double)zby);
      } else {
        reinterpret_cast<float2 *>(zf)[ta] = make_float2(zax, zay);
        reinterpret_cast<float2 *>(zf)[tb] = make_float2(zbx, zby);
      }
    } else {
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(&empty[s]);
    }
    cur = nxt;
  }
}

// Chebyshev start in fp32 storage (fp64 inputs): res = rv - JT*zp ; d = dinv*res/theta.
// Same row-pair mapping as k_cheb_init_p2 (thread = 2 rows x 2 members).
__global__ void __launch_bounds__(SPB_THREADS)
k_cheb_init_p2f(CsrDev A, const double2 *__restrict__ zp, const double2 *__restrict__ rv,
                const float2 *__restrict__ dinv, float2 *__restrict__ res, float2 *__restrict__ d,
                int nb, int npairs, float inv_theta) {
  dnsb_pdl_entry();
  SPP_MAP(false)
  const size_t sa_ = valid ? ta_ : 0, sb_ = valid ? tb_ : 0;
  const double2 ra = rv[sa_], rb = rv[sb_];
  const float2 da = dinv[sa_], db = dinv[sb_];
  double2 a, b;
  spp_rowdots<false>(A, make_double2(0.0, 0.0), zp + mp, nb2, rp, valid, wp0, wp1, lane, sv2, sv1,
                     soff, a, b);
  if (valid) {
    const float rax = (float)(ra.x - a.x), ray = (float)(ra.y - a.y);
    const float rbx = (float)(rb.x - b.x), rby = (float)(rb.y - b.y);
    res[ta_] = make_float2(rax, ray);
    res[tb_] = make_float2(rbx, rby);
    d[ta_] = make_float2(da.x * rax * inv_theta, da.y * ray * inv_theta);
    d[tb_] = make_float2(db.x * rbx * inv_theta, db.y * rby * inv_theta);
  }
}

// ---------------------------------------------------------------------------
// y = alpha*A*x + beta*z on the PAIRED rows of A with the same staging (fp64,
// same summation order as k_spmm_p2 / k_spmm_k2: bit-identical).  Used for the
// saddle-point matrix K = [F JT; J 0]: the unique rows of x of a tile include the
// pressure rows the gradient block refers to; the unpaired divergence rows are
// served by k_spmm_b2 as before.
// ---------------------------------------------------------------------------
// Warps beyond the producer (launch with TILE_THREADS + 32*ntw threads) serve the UNPAIRED tail rows of
// the matrix (the divergence rows J of K) beside the tile pipeline: a warp per row, lane = member pair, the
// row's entries loaded by the lanes and handed round with shuffles, 8 gathers from global memory in
// flight, sums in CSR order (bit-identical to k_spmm_b2).  They use no shared memory, so they fit next
// to the ring; a separate tail kernel could neither share the SM with the persistent tile CTA (its
// staging buffers) nor overlap on a second stream.
#define TILE_TAIL_WARPS_MAX 8
template <bool HASZ>
__global__ void __launch_bounds__(TILE_THREADS + 32 * TILE_TAIL_WARPS_MAX, 1)
k_spmm_tile(TileDev T, const double *__restrict__ coef, const double *__restrict__ x,
            const double *z, double *y, double alpha, double beta, CsrDev A, int row_begin) {
  extern __shared__ __align__(128) unsigned char tl_raw[];
  __shared__ __align__(8) uint64_t full[TILE_MAX_STAGES], empty[TILE_MAX_STAGES];
  const int nst = T.stages;
  const size_t off_val = (size_t)T.umax * TILE_ROWB;
  const size_t off_idx = off_val + (size_t)T.cap * 32;
  const size_t stage_bytes = (off_idx + (size_t)T.cap * 4 + 127) & ~(size_t)127;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      tl_mbar_init(&full[s], 1);
      tl_mbar_init(&empty[s], TILE_RP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double *d = x;   // the producer below streams rows of `d`

  // barrier set-up and the first (static) tile descriptor are taken before the programmatic-launch wait
  if (warp != TILE_RP) dnsb_pdl_entry();
  if (warp == TILE_RP) {
    int it = 0;
    int t = blockIdx.x;
    int4 dsc = make_int4(0, 0, 0, 0);
    int rc = 0, rl = 0, rs = 0;
    if (t < T.ntiles) {
      dsc = T.tdesc[t];
      if (lane < T.rmax) {
        const int *rr = T.truns + ((size_t)t * T.rmax + lane) * 3;
        rc = rr[0]; rl = rr[1]; rs = rr[2];
      }
    }
    dnsb_pdl_entry();
    for (; t < T.ntiles; t += gridDim.x, ++it) {
      const int s = it % nst;
      const int tn = t + gridDim.x;
      int4 ndsc = dsc;
      int nrc = 0, nrl = 0, nrs = 0;
      if (tn < T.ntiles) {
        ndsc = T.tdesc[tn];
        if (lane < T.rmax) {
          const int *rr = T.truns + ((size_t)tn * T.rmax + lane) * 3;
          nrc = rr[0]; nrl = rr[1]; nrs = rr[2];
        }
      }
      if (it >= nst) {
        if (lane == 0) tl_mbar_wait(&empty[s], ((it / nst) - 1) & 1);
        __syncwarp();
      }
      unsigned char *st = tl_raw + (size_t)s * stage_bytes;
      if (lane == 0) {
        tl_mbar_expect(&full[s], (uint32_t)dsc.z * TILE_ROWB + (uint32_t)dsc.y * 36);
        tl_bulk(st + off_val, T.pval + (size_t)dsc.x * 4, (uint32_t)dsc.y * 32, &full[s]);
        tl_bulk(st + off_idx, T.pidx + dsc.x, (uint32_t)dsc.y * 4, &full[s]);
      }
      __syncwarp();
      if (lane < dsc.w)
        tl_bulk(st + (size_t)rs * TILE_ROWB, d + (size_t)rc * TILE_NB, (uint32_t)rl * TILE_ROWB, &full[s]);
      dsc = ndsc; rc = nrc; rl = nrl; rs = nrs;
    }
    return;
  }

  const double2 cm = reinterpret_cast<const double2 *>(coef)[lane];
  const double2 zero = make_double2(0.0, 0.0);
  if (warp > TILE_RP) {
    const int tw = warp - TILE_RP - 1, ntw = (int)(blockDim.x >> 5) - TILE_RP - 1;
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    for (int row = row_begin + blockIdx.x * ntw + tw; row < A.nrows; row += gridDim.x * ntw) {
      const int k0 = A.indptr[row], k1 = A.indptr[row + 1];
      double ax = 0.0, ay = 0.0;
      for (int base = k0; base < k1; base += 32) {
        const int cnt = min(32, k1 - base);
        int col = 0;
        double v1 = 0.0, v2 = 0.0;
        if (lane < cnt) {
          col = A.indices[base + lane];
          v1 = A.v1[base + lane];
          v2 = A.v2[base + lane];
        }
        for (int k = 0; k < cnt; k += 8) {
          double2 xs[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = __shfl_sync(0xffffffffu, col, (k + u) & 31);
            xs[u] = (k + u < cnt) ? x2[(size_t)c * (TILE_NB / 2) + lane] : zero;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const double a1 = __shfl_sync(0xffffffffu, v1, (k + u) & 31);
            const double a2 = __shfl_sync(0xffffffffu, v2, (k + u) & 31);
            if (k + u < cnt) {
              ax = __fma_rn(__fma_rn(cm.x, a2, a1), xs[u].x, ax);
              ay = __fma_rn(__fma_rn(cm.y, a2, a1), xs[u].y, ay);
            }
          }
        }
      }
      const size_t t = (size_t)row * (TILE_NB / 2) + lane;
      if (HASZ) {
        const double2 zin = reinterpret_cast<const double2 *>(z)[t];
        reinterpret_cast<double2 *>(y)[t] = make_double2(alpha * ax + beta * zin.x, alpha * ay + beta * zin.y);
      } else {
        reinterpret_cast<double2 *>(y)[t] = make_double2(alpha * ax, alpha * ay);
      }
    }
    return;
  }
  struct Ops { double2 za, zb; int4 pd; };
  auto load_ops = [&](int tile) {
    Ops o;
    const int q = min(tile * TILE_RP + warp, T.npairs - 1);
    o.pd = T.pdesc[q];
    const size_t ia = (size_t)(2 * q) * (TILE_NB / 2) + lane, ib = ia + TILE_NB / 2;
    o.za = HASZ ? reinterpret_cast<const double2 *>(z)[ia] : zero;
    o.zb = HASZ ? reinterpret_cast<const double2 *>(z)[ib] : zero;
    return o;
  };
  Ops cur = load_ops(min((int)blockIdx.x, T.ntiles - 1));
  int it = 0;
  for (int t = blockIdx.x; t < T.ntiles; t += gridDim.x, ++it) {
    const int s = it % nst;
    const int rp = t * TILE_RP + warp;
    const bool have = rp < T.npairs;
    const int kb = cur.pd.x, L = have ? cur.pd.y : 0;
    const size_t ta = (size_t)(2 * (have ? rp : 0)) * (TILE_NB / 2) + lane, tb = ta + TILE_NB / 2;
    const int tn = t + gridDim.x;
    Ops nxt = cur;
    if (tn < T.ntiles) nxt = load_ops(tn);
    tl_mbar_wait(&full[s], (it / nst) & 1);
    const unsigned char *st = tl_raw + (size_t)s * stage_bytes;
    if (have) {
      const double2 *sval = reinterpret_cast<const double2 *>(st + off_val) + (size_t)kb * 2;
      const int *sidx = reinterpret_cast<const int *>(st + off_idx) + kb;
      const unsigned char *xt = st + (size_t)lane * 16;
      double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
      int k = 0;
      for (; k + 2 <= L; k += 2) {
        const int o0 = sidx[k], o1 = sidx[k + 1];
        const double2 x0 = *reinterpret_cast<const double2 *>(xt + o0);
        const double2 x1 = *reinterpret_cast<const double2 *>(xt + o1);
        const double2 a0 = sval[2 * k], b0 = sval[2 * k + 1], a1 = sval[2 * k + 2], b1 = sval[2 * k + 3];
        ax = spp_acc(ax, cm.x, a0, x0.x);  ay = spp_acc(ay, cm.y, a0, x0.y);
        bx = spp_acc(bx, cm.x, b0, x0.x);  by = spp_acc(by, cm.y, b0, x0.y);
        ax = spp_acc(ax, cm.x, a1, x1.x);  ay = spp_acc(ay, cm.y, a1, x1.y);
        bx = spp_acc(bx, cm.x, b1, x1.x);  by = spp_acc(by, cm.y, b1, x1.y);
      }
      for (; k < L; ++k) {
        const double2 xv = *reinterpret_cast<const double2 *>(xt + sidx[k]);
        const double2 a = sval[2 * k], b = sval[2 * k + 1];
        ax = spp_acc(ax, cm.x, a, xv.x);  ay = spp_acc(ay, cm.y, a, xv.y);
        bx = spp_acc(bx, cm.x, b, xv.x);  by = spp_acc(by, cm.y, b, xv.y);
      }
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(&empty[s]);
      if (HASZ) {
        reinterpret_cast<double2 *>(y)[ta] = make_double2(alpha * ax + beta * cur.za.x, alpha * ay + beta * cur.za.y);
        reinterpret_cast<double2 *>(y)[tb] = make_double2(alpha * bx + beta * cur.zb.x, alpha * by + beta * cur.zb.y);
      } else {
        reinterpret_cast<double2 *>(y)[ta] = make_double2(alpha * ax, alpha * ay);
        reinterpret_cast<double2 *>(y)[tb] = make_double2(alpha * bx, alpha * by);
      }
    } else {
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(&empty[s]);
    }
    cur = nxt;
  }
}

// ---------------------------------------------------------------------------
// Chebyshev start of the fp32 smoother with the same staging:  res = rv - JT*zp ;  d = dinv*res/theta
// (k_cheb_init_p2f: 15 us for 26 MB, bound by the dependent loads of its gather).  `T` is the tile plan
// of the gradient block JT (one value array: the second coefficient of a packed entry is zero), `x` =
// zp; sums in CSR order.  The stages are small (a tile touches ~20 pressure rows): two CTAs per SM.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS, 2)
k_cheb_init_tilef(TileDev T, const double *__restrict__ x, const double2 *__restrict__ rv,
                  const float2 *__restrict__ dinv, float2 *__restrict__ res, float2 *__restrict__ dout,
                  float inv_theta) {
  extern __shared__ __align__(128) unsigned char tl_raw[];
  __shared__ __align__(8) uint64_t full[TILE_MAX_STAGES], empty[TILE_MAX_STAGES];
  const int nst = T.stages;
  const size_t off_val = (size_t)T.umax * TILE_ROWB;
  const size_t off_idx = off_val + (size_t)T.cap * 32;
  const size_t stage_bytes = (off_idx + (size_t)T.cap * 4 + 127) & ~(size_t)127;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      tl_mbar_init(&full[s], 1);
      tl_mbar_init(&empty[s], TILE_RP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double *d = x;   // the producer below streams rows of `d`

  // barrier set-up and the first (static) tile descriptor are taken before the programmatic-launch wait
  if (warp != TILE_RP) dnsb_pdl_entry();
  if (warp == TILE_RP) {
    int it = 0;
    int t = blockIdx.x;
    int4 dsc = make_int4(0, 0, 0, 0);
    int rc = 0, rl = 0, rs = 0;
    if (t < T.ntiles) {
      dsc = T.tdesc[t];
      if (lane < T.rmax) {
        const int *rr = T.truns + ((size_t)t * T.rmax + lane) * 3;
        rc = rr[0]; rl = rr[1]; rs = rr[2];
      }
    }
    dnsb_pdl_entry();
    for (; t < T.ntiles; t += gridDim.x, ++it) {
      const int s = it % nst;
      const int tn = t + gridDim.x;
      int4 ndsc = dsc;
      int nrc = 0, nrl = 0, nrs = 0;
      if (tn < T.ntiles) {
        ndsc = T.tdesc[tn];
        if (lane < T.rmax) {
          const int *rr = T.truns + ((size_t)tn * T.rmax + lane) * 3;
          nrc = rr[0]; nrl = rr[1]; nrs = rr[2];
        }
      }
      if (it >= nst) {
        if (lane == 0) tl_mbar_wait(&empty[s], ((it / nst) - 1) & 1);
        __syncwarp();
      }
      unsigned char *st = tl_raw + (size_t)s * stage_bytes;
      if (lane == 0) {
        tl_mbar_expect(&full[s], (uint32_t)dsc.z * TILE_ROWB + (uint32_t)dsc.y * 36);
        tl_bulk(st + off_val, T.pval + (size_t)dsc.x * 4, (uint32_t)dsc.y * 32, &full[s]);
        tl_bulk(st + off_idx, T.pidx + dsc.x, (uint32_t)dsc.y * 4, &full[s]);
      }
      __syncwarp();
      if (lane < dsc.w)
        tl_bulk(st + (size_t)rs * TILE_ROWB, d + (size_t)rc * TILE_NB, (uint32_t)rl * TILE_ROWB, &full[s]);
      dsc = ndsc; rc = nrc; rl = nrl; rs = nrs;
    }
    return;
  }

  const double2 zero = make_double2(0.0, 0.0);
  struct Ops { double2 ra, rb; float2 da, db; int4 pd; };
  auto load_ops = [&](int tile) {
    Ops o;
    const int q = min(tile * TILE_RP + warp, T.npairs - 1);
    o.pd = T.pdesc[q];
    const size_t ia = (size_t)(2 * q) * (TILE_NB / 2) + lane, ib = ia + TILE_NB / 2;
    o.ra = rv[ia]; o.rb = rv[ib];
    o.da = dinv[ia]; o.db = dinv[ib];
    return o;
  };
  Ops cur = load_ops(min((int)blockIdx.x, T.ntiles - 1));
  int it = 0;
  for (int t = blockIdx.x; t < T.ntiles; t += gridDim.x, ++it) {
    const int s = it % nst;
    const int rp = t * TILE_RP + warp;
    const bool have = rp < T.npairs;
    const int kb = cur.pd.x, L = have ? cur.pd.y : 0;
    const size_t ta = (size_t)(2 * (have ? rp : 0)) * (TILE_NB / 2) + lane, tb = ta + TILE_NB / 2;
    const int tn = t + gridDim.x;
    Ops nxt = cur;
    if (tn < T.ntiles) nxt = load_ops(tn);
    tl_mbar_wait(&full[s], (it / nst) & 1);
    const unsigned char *st = tl_raw + (size_t)s * stage_bytes;
    if (have) {
      const double2 *sval = reinterpret_cast<const double2 *>(st + off_val) + (size_t)kb * 2;
      const int *sidx = reinterpret_cast<const int *>(st + off_idx) + kb;
      const unsigned char *xt = st + (size_t)lane * 16;
      double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
      int k = 0;
      for (; k + 2 <= L; k += 2) {
        const int o0 = sidx[k], o1 = sidx[k + 1];
        const double2 x0 = *reinterpret_cast<const double2 *>(xt + o0);
        const double2 x1 = *reinterpret_cast<const double2 *>(xt + o1);
        const double2 a0 = sval[2 * k], b0 = sval[2 * k + 1], a1 = sval[2 * k + 2], b1 = sval[2 * k + 3];
        ax = __fma_rn(a0.x, x0.x, ax);  ay = __fma_rn(a0.x, x0.y, ay);
        bx = __fma_rn(b0.x, x0.x, bx);  by = __fma_rn(b0.x, x0.y, by);
        ax = __fma_rn(a1.x, x1.x, ax);  ay = __fma_rn(a1.x, x1.y, ay);
        bx = __fma_rn(b1.x, x1.x, bx);  by = __fma_rn(b1.x, x1.y, by);
      }
      for (; k < L; ++k) {
        const double2 xv = *reinterpret_cast<const double2 *>(xt + sidx[k]);
        const double2 a = sval[2 * k], b = sval[2 * k + 1];
        ax = __fma_rn(a.x, xv.x, ax);  ay = __fma_rn(a.x, xv.y, ay);
        bx = __fma_rn(b.x, xv.x, bx);  by = __fma_rn(b.x, xv.y, by);
      }
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(&empty[s]);
      const float rax = (float)(cur.ra.x - ax), ray = (float)(cur.ra.y - ay);
      const float rbx = (float)(cur.rb.x - bx), rby = (float)(cur.rb.y - by);
      res[ta] = make_float2(rax, ray);
      res[tb] = make_float2(rbx, rby);
      dout[ta] = make_float2(cur.da.x * rax * inv_theta, cur.da.y * ray * inv_theta);
      dout[tb] = make_float2(cur.db.x * rbx * inv_theta, cur.db.y * rby * inv_theta);
    } else {
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(&empty[s]);
    }
    cur = nxt;
  }
}

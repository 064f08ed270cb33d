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
#define TILE_STAGES 3
#define TILE_THREADS (TILE_RP * 32 + 32)
#define TILE_NB 64           // members (a warp = 32 member pairs)
#define TILE_ROWB (TILE_NB * 8)   // bytes of one vector row

struct TilePlan {
  int ntiles = 0, npairs = 0;
  int umax = 0;       // most unique x rows of a tile
  int cap = 0;        // most pair-entries of a tile (aligned span, multiple of 4)
  size_t stage_bytes = 0, smem = 0;
  bool ok = false;
};

struct TileDev {
  const int *indptr;        // CSR row pointers of the matrix (pair p: entries [indptr[2p]/2, indptr[2p+2]/2) packed)
  const int *uptr;          // ntiles + 1: cumulative number of unique x rows
  const int *rptr;          // ntiles + 1: runs of consecutive unique columns per tile
  const int *runs;          // per run: first column, number of columns, first slot in the x tile
  const int *pidx;          // per pair-entry: byte offset of the x row inside the tile
  const double *pval;       // per pair-entry: (a.v1, a.v2, b.v1, b.v2)
  int ntiles, npairs, umax, cap;
  int dbg;   // timing experiments only (DNSB_TILE_DBG): 1 = no x copies, 2 = no gather loop, 4 = no entry copies
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
  __shared__ __align__(8) uint64_t full[TILE_STAGES], empty[TILE_STAGES];
  // stage: x tile (umax rows) | packed values (cap x 32 B) | gather offsets (cap x 4 B)
  const size_t off_val = (size_t)T.umax * TILE_ROWB;
  const size_t off_idx = off_val + (size_t)T.cap * 32;
  const size_t stage_bytes = (off_idx + (size_t)T.cap * 4 + 127) & ~(size_t)127;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TILE_STAGES; ++s) {
      tl_mbar_init(&full[s], 1);
      tl_mbar_init(&empty[s], TILE_RP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == TILE_RP) {
    // ---- producer warp ----
    int it = 0;
    for (int t = blockIdx.x; t < T.ntiles; t += gridDim.x, ++it) {
      const int s = it % TILE_STAGES;
      if (it >= TILE_STAGES) {
        if (lane == 0) tl_mbar_wait(&empty[s], ((it / TILE_STAGES) - 1) & 1);
        __syncwarp();
      }
      const int p0 = t * TILE_RP, p1 = min(T.npairs, p0 + TILE_RP);
      const int pe0 = T.indptr[2 * p0] >> 1, pe1 = T.indptr[2 * p1] >> 1;
      const int a0 = pe0 & ~3, a1 = (pe1 + 3) & ~3;
      const int nu = T.uptr[t + 1] - T.uptr[t];
      unsigned char *st = tl_raw + (size_t)s * stage_bytes;
      if (lane == 0) {
        tl_mbar_expect(&full[s], ((T.dbg & 1) ? 0u : (uint32_t)nu * TILE_ROWB) +
                                     ((T.dbg & 4) ? 0u : (uint32_t)(a1 - a0) * 36));
        if (!(T.dbg & 4)) {
          tl_bulk(st + off_val, T.pval + (size_t)a0 * 4, (uint32_t)(a1 - a0) * 32, &full[s]);
          tl_bulk(st + off_idx, T.pidx + a0, (uint32_t)(a1 - a0) * 4, &full[s]);
        }
      }
      __syncwarp();
      if (T.dbg & 1) continue;
      // the unique columns of a tile (Hilbert-ordered mesh nodes) form a few runs of consecutive
      // rows of x: one bulk copy per run (the TMA engine is slow on many small copies)
      const int r0 = T.rptr[t], nr = T.rptr[t + 1] - r0;
      for (int r = lane; r < nr; r += 32) {
        const int col = T.runs[3 * (r0 + r)], len = T.runs[3 * (r0 + r) + 1], sl = T.runs[3 * (r0 + r) + 2];
        tl_bulk(st + (size_t)sl * TILE_ROWB, d + (size_t)col * TILE_NB, (uint32_t)len * TILE_ROWB, &full[s]);
      }
    }
    return;
  }

  // ---- consumer warps: warp w owns row pair p0 + w of every tile ----
  const double2 cm = reinterpret_cast<const double2 *>(coef)[lane];
  const double2 zero = make_double2(0.0, 0.0);
  // update operands of the warp's own two rows (coalesced, addresses known up front): loaded
  // ONE TILE AHEAD into registers, so that their DRAM latency overlaps the previous tile
  struct Ops { double2 ra, rb, oa, ob, da, db, za, zb; };
  auto load_ops = [&](int tile) {
    Ops o;
    const int q = min(tile * TILE_RP + warp, T.npairs - 1);
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
    const int s = it % TILE_STAGES;
    const int p0 = t * TILE_RP;
    const int rp = p0 + warp;
    const bool have = rp < T.npairs;
    const int tpe0 = T.indptr[2 * p0] >> 1;
    const int k0 = have ? (T.indptr[2 * rp] >> 1) : 0;
    const int L = have ? (T.indptr[2 * rp + 1] - T.indptr[2 * rp]) : 0;
    const size_t ta = (size_t)(2 * (have ? rp : 0)) * (TILE_NB / 2) + lane, tb = ta + TILE_NB / 2;
    const int tn = t + gridDim.x;
    Ops nxt = cur;
    if (tn < T.ntiles) nxt = load_ops(tn);
    tl_mbar_wait(&full[s], (it / TILE_STAGES) & 1);
    const unsigned char *st = tl_raw + (size_t)s * stage_bytes;
    if (have) {
      const int kb = k0 - (tpe0 & ~3);
      const double2 *sval = reinterpret_cast<const double2 *>(st + off_val) + (size_t)kb * 2;
      const int *sidx = reinterpret_cast<const int *>(st + off_idx) + kb;
      const unsigned char *xt = st + (size_t)lane * 16;
      double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
      int k = 0;
      if (T.dbg & 2) k = L;
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

// dnsb_tc.cuh -- the dense Schur block of the preconditioner on the 5th-gen
// tensor cores:  Y = alpha * (D X [+ mass term]),  D = fp32 copy of the
// (np x np) inverse of the pressure Schur approximation, X = np x nb block of
// FGMRES residuals (fp64 in, fp64 out).
//
// Why reduced precision is legitimate HERE (and nowhere else on the path): the
// block is one factor of the PRECONDITIONER of a *flexible* GMRES.  The Krylov
// bases, the residuals, the stopping test and the solution update stay fp64, so
// the solve reaches the same fp64 tolerance; the Schur approximation
// J Z_2 J^T itself differs from the true Schur complement by ~10 %, five
// orders of magnitude more than the TF32 rounding (2^-11) of its application.
// tests: test_schur_tensor_core_block_* (iteration counts and per-step parity).
//
// Kernel (sm_100a only): split-K GEMM, CTA tile 128 rows x NP members
// (NP = nb rounded up to 16), K in blocks of 32 floats (one 128-byte swizzle
// atom).  D is stored PRE-TILED and PRE-SWIZZLED: tile (mtile, kblock) = 128 rows
// x 128 bytes in exactly the SWIZZLE_128B shared-memory image the UMMA
// descriptor expects (16-byte chunk c of row r at chunk position c ^ (r & 7)),
// 16 KB contiguous in global memory -- so a stage of A is ONE 1-D bulk copy
// (cp.async.bulk, SASS UBLKCP) streaming a contiguous block at full DRAM
// efficiency (the row-major copy fed through a 2-D tensor map -- 128 separate
// 128-byte row segments per stage, 15 KB apart -- reached 2.8 TB/s only).  The
// small X operand (K-major fp32 copy, L2 resident) comes through a 2-D tensor
// map (cp.async.bulk.tensor.2d, SWIZZLE_128B, SASS UTMALDG).  Warp 0 = TMA
// producer (ring of mbarrier stages), warp 1 = MMA issuer (tcgen05.mma.cta_group::1.kind::tf32,
// one elected thread, accumulator 128 lanes x NP columns in TMEM), warps 2-5 =
// epilogue (tcgen05.ld 32x32b -> registers -> fp32 partial tile in global
// memory).  k_tc_epilogue sums the split-K partials in split order
// (deterministic), applies alpha / the mass term and converts to fp64.
//
// Bytes per launch: 4*np^2 (D streamed once) + small; HBM bound: np = 3836:
// 59 MB -> 9 us at the measured copy peak (the fp64 DMMA kernel: 70 us at 73 %
// of the fp64 pipe).
//
// replaces: the pressure-Schur part of the sparse LU solve of
// time_int_utils.py:89-91,132 (as one block of the preconditioner).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#define TC_BM 128            // rows of D per CTA (UMMA M)
#define TC_BK 32             // floats of K per stage (128 bytes: one swizzle atom)
#define TC_UK 8              // UMMA K for kind::tf32 (32 bytes)
#define TC_THREADS 192
#define TC_MAX_STAGES 8

struct TcPlan {
  CUtensorMap mapB;
  int np_ = 0;        // NP: members padded to a multiple of 16
  int mtiles = 0, splits = 0, kblocks = 0, kb_per_split = 0, stages = 0;
  int ldx = 0;        // leading dimension (floats) of the transposed X copy
  size_t smem = 0;
  bool ok = false;
};

__device__ __forceinline__ uint32_t tc_smem(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tc_mbar_init(uint64_t *b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem(b)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *b, uint32_t parity) {
  const uint32_t a = tc_smem(b);
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
__device__ __forceinline__ void tc_tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(tc_smem(dst)), "l"(map), "r"(c0), "r"(c1), "r"(tc_smem(bar))
      : "memory");
}
// K-major operand tile, 128-byte swizzle, rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start >> 4 | LBO 1 | SBO 1024 >> 4 | version 1 | SWIZZLE_128B)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem(bar))
               : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// part[(split*mtiles*128 + row)*NP + col] (fp32)
__device__ __forceinline__ void tc_bulk_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc_smem(dst)),
               "l"(src), "r"(bytes), "r"(tc_smem(bar))
               : "memory");
}

// Dt: packed tiles [(mtile*kblocks + kb)][128 rows][32 floats], swizzled (k_tc_pack_d)
__global__ void __launch_bounds__(TC_THREADS, 1)
k_schur_tc(const float *__restrict__ Dt, const __grid_constant__ CUtensorMap mapB,
           float *__restrict__ part, int NP, int kblocks, int kb_per_split, int stages, int tmem_cols,
           int l2keep) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  __shared__ __align__(8) uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES], accum_full;
  __shared__ uint32_t tmem_base_s;
  // the dynamic window is only 16-byte aligned by contract: align the ring to 1024 (SWIZZLE_128B)
  unsigned char *ring = (unsigned char *)(((uintptr_t)tc_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TC_BM * TC_BK * 4, b_bytes = (uint32_t)NP * TC_BK * 4;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mtile = blockIdx.x, split = blockIdx.y;
  const int kb0 = split * kb_per_split, kb1 = min(kblocks, kb0 + kb_per_split);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      tc_mbar_init(&full[s], 1);
      tc_mbar_init(&empty[s], 1);
    }
    tc_mbar_init(&accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    // TMEM: tmem_cols (power of two >= 32) columns x 128 lanes of fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&tmem_base_s)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  dnsb_pdl_entry();   // barriers and the TMEM allocation above do not depend on the previous kernel

  if (warp == 0) {
    // ---- TMA producer ----
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % stages;
        if (i >= stages) tc_mbar_wait(&empty[s], ((i / stages) - 1) & 1);
        unsigned char *st = ring + (size_t)s * stage_bytes;
        tc_mbar_expect(&full[s], a_bytes + b_bytes);
        tc_bulk_1d(st, Dt + ((size_t)mtile * kblocks + kb0 + i) * (TC_BM * TC_BK), a_bytes, &full[s]);
        tc_tma_2d(st + a_bytes, &mapB, (kb0 + i) * TC_BK, 0, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, M = 128, N = NP ----
    // instruction descriptor (cute::UMMA::InstrDescriptor): c = F32 (1 << 4), a = b = TF32 (2 << 7,
    // 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) |
                           ((uint32_t)(TC_BM >> 4) << 24);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % stages;
      tc_mbar_wait(&full[s], (i / stages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t sa = tc_smem(ring + (size_t)s * stage_bytes), sb = sa + a_bytes;
        const uint64_t da = tc_desc(sa), db = tc_desc(sb);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UK; ++k)
          // advancing along K inside the swizzle atom: + k * 32 bytes (>> 4) on the start address
          tc_mma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (i | k) != 0);
        tc_commit(&empty[s]);               // stage free once these MMAs have read it
        if (i == nkb - 1) tc_commit(&accum_full);
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue: warp w reads the TMEM lanes of its quarter (w % 4) ----
    const int q = warp & 3;
    const int row = mtile * TC_BM + q * 32 + lane;
    float *out = part + ((size_t)split * gridDim.x * TC_BM + row) * NP;
    if (nkb > 0) {
      tc_mbar_wait(&accum_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = 0; c < NP; c += 16) {
        uint32_t r[16];
        tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float4 *o4 = reinterpret_cast<float4 *>(out + c);
#pragma unroll
        for (int v = 0; v < 4; ++v)
          o4[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                              __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
      }
    } else {
      for (int c = 0; c < NP; ++c) out[c] = 0.0f;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols)
                 : "memory");
}

// fp32 value rounded to the nearest TF32 (10-bit mantissa); the tensor core would truncate instead
__device__ __forceinline__ float tc_round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// packed, swizzled, TF32-rounded tiles of the n x n fp64 matrix `src`:
// tile (mt, kb) at dst + (mt*kblocks + kb)*4096 floats; element (r, k) of the tile (row mt*128 + r,
// column kb*32 + k) at float offset r*32 + (((k >> 2) ^ (r & 7)) << 2) + (k & 3); zero outside n
__global__ void k_tc_pack_d(const double *__restrict__ src, float *__restrict__ dst, int n, int mtiles,
                            int kblocks) {
  dnsb_pdl_entry();
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)mtiles * kblocks * TC_BM * TC_BK;
  if (t >= total) return;
  const int k = (int)(t & 31), r = (int)((t >> 5) & 127);
  const size_t tile = t >> 12;
  const int kb = (int)(tile % kblocks), mt = (int)(tile / kblocks);
  const int row = mt * TC_BM + r, col = kb * TC_BK + k;
  const float v = (row < n && col < n) ? tc_round_tf32((float)src[(size_t)row * n + col]) : 0.0f;
  dst[tile * (TC_BM * TC_BK) + (size_t)r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3))] = v;
}

// Xt[m*ldx + k] = tf32(x[k*nb + m])  (K-major copy of the right-hand sides; rows m >= nb stay zero)
__global__ void k_tc_pack_x(const double *__restrict__ x, float *__restrict__ xt, int n, int nb, int ldx) {
  dnsb_pdl_entry();
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, m = m0 + tx;
    tile[r][tx] = (k < n && m < nb) ? tc_round_tf32((float)x[(size_t)k * nb + m]) : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, k = k0 + tx;
    if (m < nb && k < ldx) xt[(size_t)m * ldx + k] = tile[tx][r];
  }
}

// y[i,m] = alpha * (sum_s part[s][i][m] + add_scale[m]*add_dinv[i]*x[i,m]), splits summed in order
__global__ void k_tc_epilogue(const float *__restrict__ part, int splits, size_t split_stride, int NP,
                              const double *__restrict__ x, double *__restrict__ y, int n, int nb,
                              double alpha, const double *__restrict__ add_dinv,
                              const double *__restrict__ add_scale) {
  dnsb_pdl_entry();
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * nb) return;
  const int i = (int)(t / nb), m = (int)(t % nb);
  double v = 0.0;
  for (int s = 0; s < splits; ++s) v += (double)part[(size_t)s * split_stride + (size_t)i * NP + m];
  if (add_dinv) v += add_scale[m] * add_dinv[i] * x[t];
  y[t] = alpha * v;
}

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
  dnsb_pdl_entry();
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
  dnsb_pdl_entry();
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

// ---------------------------------------------------------------------------
// Tensor-core variant of the stream-K dense solve: fp64 DMMA
// (mma.sync.m8n8k4.f64 -- the only fp64 matrix instruction; tcgen05 has no
// fp64 kind) fed by a 3-stage cp.async pipeline.  Same decomposition and
// partial-tile layout as k_dense_gemm_streamk (same epilogue); requires n and
// nb even (16-byte cp.async chunks).
//   CTA = 4 warps, tile 64 rows x 64 members; warp w owns rows 16w..16w+15:
//   2 m-tiles x 8 n-tiles of 8x8 accumulators (32 doubles per lane).
//   Shared tiles: D row-major [64][16] (row stride 20 doubles), X [16][64]
//   (row stride 68 doubles): the fragment loads are bank-conflict free.
// ---------------------------------------------------------------------------
#define DMM_STAGES 3
#define DMM_DS 20   // row stride of the D tile (doubles)
#define DMM_XS 68   // row stride of the X tile (doubles)
#define DMM_STAGE_DOUBLES (DGK_TM * DMM_DS + DGK_TK * DMM_XS)
#define DMM_SMEM_BYTES (DMM_STAGES * DMM_STAGE_DOUBLES * 8)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
// the same with an L2 evict-first policy: the dense inverse (118 MB at np = 3836) is read once per
// application and must not sweep the vectors and sparse operators of the other step kernels out of L2
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async16_stream(void *smem, const void *gmem, int src_bytes,
                                                  unsigned long long pol) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;\n" ::"r"(s), "l"(gmem),
               "r"(src_bytes), "l"(pol));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128)
k_dense_dmma_streamk(const double *__restrict__ D, const double *__restrict__ X,
                     double *__restrict__ part, int n, int nb, DenseSplit sp) {
  dnsb_pdl_entry();
  extern __shared__ __align__(16) double dsm[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int m0 = blockIdx.y * 64;
  const long utotal = (long)((n + DGK_TM - 1) / DGK_TM) * sp.ksteps;
  long u = (long)blockIdx.x * sp.upc;
  const long uend = min(utotal, u + sp.upc);
  int seg = 0;
  while (u < uend) {
    const int rt = (int)(u / sp.ksteps);
    const int ks0 = (int)(u - (long)rt * sp.ksteps);
    const int ks1 = (int)min((long)sp.ksteps, ks0 + (uend - u));
    const int row0 = rt * DGK_TM;
    const int nk = ks1 - ks0;
    double acc[2][8][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

#if DNSB_L2_HINTS
    const unsigned long long l2pol = l2_evict_first_policy();
#endif
    auto issue = [&](int kstep, int stage) {
      double *sD = dsm + (size_t)stage * DMM_STAGE_DOUBLES;
      double *sX = sD + DGK_TM * DMM_DS;
      const int k0 = (ks0 + kstep) * DGK_TK;
      // D tile: 64 rows x 8 chunks of 2 doubles
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = tid + 128 * q;
        const int r = c >> 3, ch = c & 7;
        const int gi = row0 + r, gk = k0 + 2 * ch;
        const bool ok = gi < n && gk < n;
#if DNSB_L2_HINTS
        cp_async16_stream(sD + r * DMM_DS + 2 * ch, ok ? (const void *)(D + (size_t)gi * n + gk) : (const void *)D,
                          ok ? 16 : 0, l2pol);
#else
        cp_async16(sD + r * DMM_DS + 2 * ch, ok ? (const void *)(D + (size_t)gi * n + gk) : (const void *)D,
                   ok ? 16 : 0);
#endif
      }
      // X tile: 16 rows x 32 chunks of 2 members
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = tid + 128 * q;
        const int r = c >> 5, ch = c & 31;
        const int gk = k0 + r, gm = m0 + 2 * ch;
        const bool ok = gk < n && gm < nb;
        cp_async16(sX + r * DMM_XS + 2 * ch, ok ? (const void *)(X + (size_t)gk * nb + gm) : (const void *)X,
                   ok ? 16 : 0);
      }
    };

    __syncthreads();   // the previous segment is done with the buffers
#pragma unroll
    for (int s = 0; s < DMM_STAGES - 1; ++s) {
      if (s < nk) issue(s, s);
      cp_async_commit();
    }
    for (int ks = 0; ks < nk; ++ks) {
      cp_async_wait<DMM_STAGES - 2>();
      __syncthreads();
      // prefetch the tile DMM_STAGES-1 steps ahead into the stage freed last
      const int nxt = ks + DMM_STAGES - 1;
      if (nxt < nk) issue(nxt, nxt % DMM_STAGES);
      cp_async_commit();
      const double *sD = dsm + (size_t)(ks % DMM_STAGES) * DMM_STAGE_DOUBLES;
      const double *sX = sD + DGK_TM * DMM_DS;
#pragma unroll
      for (int kk = 0; kk < DGK_TK; kk += 4) {
        const double a0 = sD[(16 * w + (lane >> 2)) * DMM_DS + kk + (lane & 3)];
        const double a1 = sD[(16 * w + 8 + (lane >> 2)) * DMM_DS + kk + (lane & 3)];
        double b[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) b[t] = sX[(kk + (lane & 3)) * DMM_XS + 8 * t + (lane >> 2)];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          dmma884(acc[0][t][0], acc[0][t][1], a0, b[t]);
          dmma884(acc[1][t][0], acc[1][t][1], a1, b[t]);
        }
      }
    }
    cp_async_wait<0>();
    double *out = part + ((size_t)blockIdx.x * sp.maxseg + seg) * DGK_TM * nb;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int rl = 16 * w + 8 * a + (lane >> 2);
      if (row0 + rl >= n) continue;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int gm = m0 + 8 * t + 2 * (lane & 3);
        if (gm + 1 < nb) {
          *reinterpret_cast<double2 *>(out + (size_t)rl * nb + gm) = make_double2(acc[a][t][0], acc[a][t][1]);
        } else if (gm < nb) {
          out[(size_t)rl * nb + gm] = acc[a][t][0];
        }
      }
    }
    u += nk;
    ++seg;
  }
}

// ---------------------------------------------------------------------------
// Optional reduced-precision variant of the dense Schur solve (DNSB_SCHUR_TF32=1,
// off by default): the dense inverse is kept as an fp32 copy and applied with
// the TF32 tensor-core instruction mma.sync.m16n8k8 in the 3xTF32 split
// (a = a_hi + a_lo, b = b_hi + b_lo; a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32
// accumulation): fp32-accurate products.  It is a PRECONDITIONER block inside a
// flexible GMRES whose residuals, bases and stopping test stay fp64, so the
// computed solution meets the same fp64 residual tolerance; only the iteration
// count can change.  Half the bytes of D (4 n^2) and no fp64-pipe bound.
// Same stream-K decomposition, partial-tile layout (fp64) and epilogue as the
// fp64 kernels.  Requires nb % 4 == 0; D32 has leading dimension ld (ld % 4 == 0).
//   CTA = 4 warps, tile 64 rows x 64 members, warp w owns rows 16w..16w+15:
//   8 n-tiles of m16n8 accumulators (32 floats per lane).
//   Shared tiles: D [64][16] floats (row stride 20), X [16][64] floats (row
//   stride 72): conflict-free fragment loads.
// ---------------------------------------------------------------------------
#define TFM_STAGES 6
#define TFM_DS 20
#define TFM_XS 72
#define TFM_STAGE_FLOATS (DGK_TM * TFM_DS + DGK_TK * TFM_XS)
#define TFM_SMEM_BYTES (TFM_STAGES * TFM_STAGE_FLOATS * 4)

__device__ __forceinline__ void tf32_split(float v, unsigned &hi, unsigned &lo) {
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(hi) : "f"(v));
  const float r = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32_1688(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void k_f64_to_f32(const double *__restrict__ src, float *__restrict__ dst, size_t rows,
                             size_t cols, size_t ld) {
  dnsb_pdl_entry();
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * ld) return;
  const size_t r = t / ld, c = t - r * ld;
  dst[t] = c < cols ? (float)src[r * cols + c] : 0.0f;
}

//   PASSES = 3: 3xTF32 as above; 2: a_hi (b_hi + b_lo) -- D rounded to TF32 once, X fp32-accurate;
//   1: a_hi b_hi (plain TF32)
template <int PASSES>
__global__ void __launch_bounds__(128)
k_dense_tf32_streamk(const float *__restrict__ D, int ld, const float *__restrict__ X,
                     double *__restrict__ part, int n, int nb, DenseSplit sp) {
  dnsb_pdl_entry();
  extern __shared__ __align__(16) float fsm[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int m0 = blockIdx.y * 64;
  const long utotal = (long)((n + DGK_TM - 1) / DGK_TM) * sp.ksteps;
  long u = (long)blockIdx.x * sp.upc;
  const long uend = min(utotal, u + sp.upc);
  int seg = 0;
#if DNSB_L2_HINTS
  const unsigned long long l2pol = l2_evict_first_policy();
#endif
  while (u < uend) {
    const int rt = (int)(u / sp.ksteps);
    const int ks0 = (int)(u - (long)rt * sp.ksteps);
    const int ks1 = (int)min((long)sp.ksteps, ks0 + (uend - u));
    const int row0 = rt * DGK_TM;
    const int nk = ks1 - ks0;
    float acc[8][4];
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.0f;

    auto issue = [&](int kstep, int stage) {
      float *sD = fsm + (size_t)stage * TFM_STAGE_FLOATS;
      float *sX = sD + DGK_TM * TFM_DS;
      const int k0 = (ks0 + kstep) * DGK_TK;
      // D tile: 64 rows x 4 chunks of 4 floats
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = tid + 128 * q;
        const int r = c >> 2, ch = c & 3;
        const int gi = row0 + r, gk = k0 + 4 * ch;
        const bool ok = gi < n && gk < ld;
        const void *src = ok ? (const void *)(D + (size_t)gi * ld + gk) : (const void *)D;
#if DNSB_L2_HINTS
        cp_async16_stream(sD + r * TFM_DS + 4 * ch, src, ok ? 16 : 0, l2pol);
#else
        cp_async16(sD + r * TFM_DS + 4 * ch, src, ok ? 16 : 0);
#endif
      }
      // X tile: 16 rows x 16 chunks of 4 members
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = tid + 128 * q;
        const int r = c >> 4, ch = c & 15;
        const int gk = k0 + r, gm = m0 + 4 * ch;
        const bool ok = gk < n && gm < nb;
        cp_async16(sX + r * TFM_XS + 4 * ch, ok ? (const void *)(X + (size_t)gk * nb + gm) : (const void *)X,
                   ok ? 16 : 0);
      }
    };

    __syncthreads();   // the previous segment is done with the buffers
#pragma unroll
    for (int s = 0; s < TFM_STAGES - 1; ++s) {
      if (s < nk) issue(s, s);
      cp_async_commit();
    }
    for (int ks = 0; ks < nk; ++ks) {
      cp_async_wait<TFM_STAGES - 2>();
      __syncthreads();
      const int nxt = ks + TFM_STAGES - 1;
      if (nxt < nk) issue(nxt, nxt % TFM_STAGES);
      cp_async_commit();
      const float *sD = fsm + (size_t)(ks % TFM_STAGES) * TFM_STAGE_FLOATS;
      const float *sX = sD + DGK_TM * TFM_DS;
#pragma unroll
      for (int kk = 0; kk < DGK_TK; kk += 8) {
        unsigned ah[4], al[4];
        tf32_split(sD[(16 * w + g) * TFM_DS + kk + t4], ah[0], al[0]);
        tf32_split(sD[(16 * w + 8 + g) * TFM_DS + kk + t4], ah[1], al[1]);
        tf32_split(sD[(16 * w + g) * TFM_DS + kk + 4 + t4], ah[2], al[2]);
        tf32_split(sD[(16 * w + 8 + g) * TFM_DS + kk + 4 + t4], ah[3], al[3]);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          unsigned bh0, bl0, bh1, bl1;
          tf32_split(sX[(kk + t4) * TFM_XS + 8 * t + g], bh0, bl0);
          tf32_split(sX[(kk + 4 + t4) * TFM_XS + 8 * t + g], bh1, bl1);
          if (PASSES >= 3) mma_tf32_1688(acc[t], al, bh0, bh1);
          if (PASSES >= 2) mma_tf32_1688(acc[t], ah, bl0, bl1);
          mma_tf32_1688(acc[t], ah, bh0, bh1);
        }
      }
    }
    cp_async_wait<0>();
    double *out = part + ((size_t)blockIdx.x * sp.maxseg + seg) * DGK_TM * nb;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int rl = 16 * w + 8 * a + g;
      if (row0 + rl >= n) continue;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int gm = m0 + 8 * t + 2 * t4;
        if (gm + 1 < nb) {
          *reinterpret_cast<double2 *>(out + (size_t)rl * nb + gm) =
              make_double2((double)acc[t][2 * a], (double)acc[t][2 * a + 1]);
        } else if (gm < nb) {
          out[(size_t)rl * nb + gm] = (double)acc[t][2 * a];
        }
      }
    }
    u += nk;
    ++seg;
  }
}

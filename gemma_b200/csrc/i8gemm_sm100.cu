// i8gemm_sm100.cu -- eigen-projection (U^T X)^T = G * U on the sm_100a tensor cores.
//
// Replaces fast_dgemm("T","N",1.0,U,Xlarge_sub,0.0,UtXlarge_sub) of LMM::Analyze
// (src/lmm.cpp:1521 / :1847): 2 n^2 flop per SNP, the dominant cost at n = 50 000.
//
// Genotypes are small integers (PLINK 2-bit: 0/1/2), so the product is made EXACT on the
// int8 tensor pipe with an error-free split of U (Ozaki scheme):
//   U[j][i] = s_i * sum_t d_t[j][i] * 256^(T-1-t) + eps,   d_t in int8,
//   s_i = max_j |U[j][i]| / (127.4 * 256^(T-1)),   |eps| <= s_i / 2
// (balanced base-256 digits of the integer rint(U / s_i); the top digit uses the whole int8 range, |d_0| <= 127, so the T planes
// keep log2(127.4) + 8 (T-1) bits of every entry relative to its column maximum).  Each digit
// plane times the int8 genotype tile accumulates exactly in int32 (|sum| <= 128*2*n < 2^31
// for n < 8e6); the T planes are recombined in FP64 in the epilogue.
// Mean-imputed missing genotypes are handled exactly as  U^T x = U^T z + mean * U^T q  with
// z = genotype with 0 at the holes (int8 GEMM) and q = hole indicator (sparse FP64 fix-up).
//
// Kernel (one CTA per SM, persistent, warp-specialised, no cluster):
//   warp 0  : TMA producer  -- cp.async.bulk.tensor 2D, 128B-swizzled K-major tiles,
//             A = 128 SNPs x 128 B of individuals, B = (T planes x NE eigenvectors) x 128 B
//   warp 1  : MMA issuer    -- tcgen05.mma.cta_group::1.kind::i8, M=128, N=T*NE (<=256), K=32,
//             accumulators in TMEM (2 x 256 columns, double buffered against the epilogue)
//   warp 2  : TMEM allocator
//   warps 4-7: epilogue     -- tcgen05.ld 32x32b, FP64 recombination of the T planes, scale,
//             store of U^T x rows (SNP-major, the layout the per-SNP kernel streams)
// All T planes of an eigenvector group live in the same N tile, so the genotype tile in
// shared memory is reused T times per load and no read-modify-write of C is needed.
#include "common.cuh"
#include <cuda.h>

namespace gb {

constexpr int I8_BM = 128;        // SNPs per tile (UMMA M, one TMEM lane per SNP)
constexpr int I8_BK = 128;        // K bytes per pipeline stage == one 128B swizzle row
constexpr int I8_UK = 32;         // UMMA K for 8-bit operands
constexpr int I8_STAGES = 4;
constexpr int I8_ACC_COLS = 256;  // TMEM columns per accumulator stage
constexpr int I8_THREADS = 256;
constexpr int I8_PANEL = 12;      // raster panel width (n-groups) for L2 reuse

struct I8Geom {
  int T, NE, N;            // planes, eigenvectors per tile, UMMA N = T*NE
  int n, n_padk;           // individuals, padded K extent in bytes
  int n_groups;            // ceil(n / NE)
};

static I8Geom make_geom(size_t n, int T) {
  I8Geom g;
  g.T = T;
  int ne = (256 / T) & ~7;                 // multiple of 8
  if ((T * ne) % 16 != 0) ne &= ~15;       // UMMA N must be a multiple of 16 for M = 128
  g.NE = ne; g.N = T * ne;
  g.n = (int)n; g.n_padk = (int)((n + I8_BK - 1) / I8_BK * I8_BK);
  g.n_groups = (int)((n + ne - 1) / ne);
  return g;
}

// ------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *tmap, uint64_t *bar, void *dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, int32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory matrix descriptor (tcgen05 "smem descriptor", version 1):
// rows of 128 B, 8-row groups 1024 B apart (SBO), 16-byte units, layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr, uint32_t lbo_units) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address  [0,14)
  d |= (uint64_t)(lbo_units & 0x3FFF) << 16;            // leading byte offset [16,30)
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;          // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell) [46,48)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B [61,64)
  return d;
}

// instruction descriptor for kind::i8: D = S32, A = a_fmt (0 u8 / 1 s8), B = s8, both K-major
__host__ __device__ constexpr uint32_t make_i8_idesc(int M, int N, int a_signed, int b_signed) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | (0u << 15) | (0u << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct I8KernelParams {
  int T, NE, N;
  int n, l;                 // eigenvectors (= individuals), SNPs
  int num_k_blocks;         // n_padk / 128
  int m_tiles, n_groups;
  int lbo_units;            // descriptor LBO field (kept runtime for bring-up)
  const double *scale;      // per eigenvector: s_i = colmax_i / (127.4 * 256^(T-1))
  double *C;                // l x n, ld = ldc   (mode 0)  |  K accumulator n x n, ld = ldc (mode 1)
  size_t ldc;
  int mode;                 // 0 = projection (T planes recombined, scaled, stored), 1 = kinship (K[i][j] += Z Z^T, lower triangle)
  const int2 *tiles;        // mode 1: explicit (m_blk, n_blk) list (lower-triangle tiles only)
  int num_tiles;
  // i8_gemm_pair2_kernel only
  const double *row_mean;   // mode 2: C[s][:] += row_mean[s] * (A . planes) * scale   (A = hole indicator rows)
  const int *tile_holes;    // mode 2: holes per 256-row tile; tiles without a hole are skipped
  const int *hole_switch;   // mode 2 (pair kernel): [0] = 1 when the batch's holes go through this GEMM pass, 0 = the sparse gather kernel has them
  int panel;                // raster panel width in units of NB eigenvector groups
  int stages;               // i8_gemm_pair_kernel: TMA pipeline stages
  unsigned int *wave_ctr;   // i8_gemm_pair_kernel: wave synchronisation counter (zeroed before the launch) or null
  int l2_hint;              // i8_gemm_pair_kernel: L2 eviction hints on the TMA loads (genotype panels evict_first, plane panels evict_last)
};

__device__ __forceinline__ void tile_coords_raster(int tile, int m_tiles, int n_groups, int panel_w, int &m_blk, int &n_grp) {
  // panels of I8_PANEL eigenvector groups; inside a panel SNP tiles vary slowest so that the ~148
  // concurrently running CTAs cover a ~12 x 12 block of (SNP tile, group) pairs: each A/B K-panel
  // streamed from HBM is shared by ~12 CTAs through L2.
  const int panel_tiles = panel_w * m_tiles;
  const int panel = tile / panel_tiles;
  const int first = panel * panel_w;
  const int width = (n_groups - first < panel_w) ? (n_groups - first) : panel_w;
  const int r = tile - panel * panel_tiles;
  m_blk = r / width;
  n_grp = first + r % width;
}

__device__ __forceinline__ void tile_coords(const I8KernelParams &p, int tile, int &m_blk, int &n_grp) {
  if (p.tiles) { const int2 t = p.tiles[tile]; m_blk = t.x; n_grp = t.y; }
  else tile_coords_raster(tile, p.m_tiles, p.n_groups, p.panel > 0 ? p.panel : I8_PANEL, m_blk, n_grp);
}

__global__ void __launch_bounds__(I8_THREADS, 1)
i8_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const I8KernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: A stages | B stages | barriers | tmem slot
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int a_bytes = I8_BM * I8_BK;
  const int b_bytes = p.N * I8_BK;
  uint8_t *smem_a = smem;
  uint8_t *smem_b = smem + I8_STAGES * a_bytes;
  uint64_t *bars = (uint64_t *)(smem_b + I8_STAGES * b_bytes);
  uint64_t *full = bars, *empty = bars + I8_STAGES;
  uint64_t *tfull = bars + 2 * I8_STAGES, *tempty = bars + 2 * I8_STAGES + 2;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * I8_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles ? p.num_tiles : p.m_tiles * p.n_groups;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_grp; tile_coords(p, tile, m_blk, n_grp);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], (uint32_t)(a_bytes + b_bytes));
          tma_load_2d(&tmap_a, &full[stage], smem_a + stage * a_bytes, kb * I8_BK, m_blk * I8_BM);
          tma_load_2d(&tmap_b, &full[stage], smem_b + stage * b_bytes, kb * I8_BK, n_grp * p.N);
          if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_i8_idesc(I8_BM, p.N, 0, 1);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * I8_ACC_COLS);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smem_a + stage * a_bytes), p.lbo_units);
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smem_b + stage * b_bytes), p.lbo_units);
#pragma unroll
          for (int k = 0; k < I8_BK / I8_UK; ++k) {
            // advance both start addresses by k*32 bytes inside the 128B swizzle row (16-byte units)
            tc_mma_i8(tmem_d, adesc + (uint64_t)(k * (I8_UK >> 4)), bdesc + (uint64_t)(k * (I8_UK >> 4)), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty[stage]);                 // frees the smem slot when these MMAs retire
          if (kb == p.num_k_blocks - 1) tc_commit(&tfull[acc]);
          if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;                         // == warp % 4 : TMEM lane quarter
    int acc = 0; uint32_t acc_phase = 0;
    const double w256 = 256.0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_grp; tile_coords(p, tile, m_blk, n_grp);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int s = m_blk * I8_BM + ew * 32 + lane;                       // SNP row of this thread
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * I8_ACC_COLS);
      const int i0 = n_grp * p.NE;
      if (p.mode == 1) {
        // kinship: exact integer tile Z Z^T added into the FP64 accumulator (lower triangle only)
        const int j0 = n_grp * p.N;
        double *krow = p.C + (size_t)s * p.ldc;
        for (int e0 = 0; e0 < p.N; e0 += 8) {
          int32_t d[8];
          tc_ld8(taddr + (uint32_t)e0, d);
          tc_ld_wait();
          if (s < p.l) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int j = j0 + e0 + q;
              if (j <= s) krow[j] += (double)d[q];
            }
          }
        }
      } else
      for (int e0 = 0; e0 < p.NE; e0 += 8) {
        double v[8];
        int32_t d[8];
        tc_ld8(taddr + (uint32_t)e0, d);
        tc_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (double)d[q];
        for (int t = 1; t < p.T; ++t) {
          tc_ld8(taddr + (uint32_t)(t * p.NE + e0), d);
          tc_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = fma(v[q], w256, (double)d[q]);
        }
        if (s < p.l) {
          double *crow = p.C + (size_t)s * p.ldc;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int i = i0 + e0 + q;
            if (i < p.n) crow[i] = v[q] * __ldg(p.scale + i);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster form one 256 x N tile.  Each CTA stages
// its own 128 SNP rows of A and HALF of the B rows; the pair's MMA reads both halves of B, so the
// shared-memory read traffic per MAC drops from (1/128 + 1/N) to (1/128 + 1/2N) bytes -- the single-CTA
// kernel above sits at ~97 B/clk of the 128 B/clk shared-memory port (tensor pipe 74-76% active).
//   - full[] barriers live in the leader (cluster rank 0): both CTAs' TMA loads complete_tx on it;
//   - the leader's single MMA thread issues tcgen05.mma.cta_group::2 and multicasts its commits to the
//     empty[] / tfull[] barriers of BOTH CTAs;
//   - every epilogue warp of both CTAs arrives (remotely for rank 1) on the leader's tempty[] barrier.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *tmap, uint32_t bar_cluster_addr, void *dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
// same with an L2 eviction-priority hint (createpolicy): the genotype panels of a wave are read once per wave (evict_first), the
// plane panels of a raster panel are re-read by the following waves (evict_last)
__device__ __forceinline__ void tma_load_2d_pair_hint(const CUtensorMap *tmap, uint32_t bar_cluster_addr, void *dst, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 2)     // <= 128 registers: leaves room for a co-resident per-SNP CTA
i8_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const I8KernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int halfN = p.N / 2;
  const int NS = p.stages;                                 // pipeline depth (shared memory permitting: 6 x 32 KB for the projection)
  const int a_bytes = I8_BM * I8_BK;
  const int b_bytes = halfN * I8_BK;                       // this CTA's half of the B rows
  uint8_t *smem_a = smem;
  uint8_t *smem_b = smem + NS * a_bytes;
  uint64_t *bars = (uint64_t *)(smem_b + NS * b_bytes);
  uint64_t *full = bars, *empty = bars + NS;
  uint64_t *tfull = bars + 2 * NS, *tempty = bars + 2 * NS + 2;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * NS + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  int num_tiles = p.tiles ? p.num_tiles : p.m_tiles * p.n_groups;          // m_tiles counts 256-row tiles here
  if (p.mode == 2 && p.hole_switch && __ldg(p.hole_switch) == 0) num_tiles = 0;   // the gather kernel handles this batch's holes (uniform over the grid)
  auto skip = [&](int m_blk) -> bool { return p.mode == 2 && __ldg(p.tile_holes + m_blk) == 0; };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }   // 4 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // peer barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      unsigned int wave = 0;
      const uint64_t pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++wave) {
        int m_blk, n_grp; tile_coords(p, tile, m_blk, n_grp);
        if (skip(m_blk)) continue;
        if (p.wave_ctr && wave > 0) {
          // Wave synchronisation: the ~74 tile pairs that run concurrently share their genotype / plane K-panels through L2
          // only while they walk K in phase (a panel is n bytes long: two tiles half a tile apart are 1 GB of other traffic
          // apart, and L2 holds 126 MB).  Left alone the pairs drift out of phase within a few waves and DRAM traffic climbs to
          // 20x the algorithmic bytes; starting every wave together keeps the reads at (a + b) panels per a x b tile block.
          // Every producer counts in after queueing its last K-block and waits (bounded) for the other CTAs before the next tile.
          const unsigned int want = gridDim.x * wave;
          unsigned int seen, spins = 0;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.wave_ctr) : "memory");
            if (seen >= want) break;
            __nanosleep(200);
          } while (++spins < 20000u);                      // ~4 ms: never a deadlock if part of the grid is not resident
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);             // local: released by the leader's multicast commit
          const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
          if (leader) mbar_expect_tx(&full[stage], (uint32_t)(2 * (a_bytes + b_bytes)));
          if (p.l2_hint) {
            tma_load_2d_pair_hint(&tmap_a, full_leader, smem_a + stage * a_bytes, kb * I8_BK, m_blk * 256 + (int)rank * I8_BM, pol_a);
            tma_load_2d_pair_hint(&tmap_b, full_leader, smem_b + stage * b_bytes, kb * I8_BK, n_grp * p.N + (int)rank * halfN, pol_b);
          } else {
            tma_load_2d_pair(&tmap_a, full_leader, smem_a + stage * a_bytes, kb * I8_BK, m_blk * 256 + (int)rank * I8_BM);
            tma_load_2d_pair(&tmap_b, full_leader, smem_b + stage * b_bytes, kb * I8_BK, n_grp * p.N + (int)rank * halfN);
          }
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        if (p.wave_ctr) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.wave_ctr) : "memory");
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_i8_idesc(256, p.N, 0, 1);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        if (p.mode == 2) { int m_blk, n_grp; tile_coords(p, tile, m_blk, n_grp); if (skip(m_blk)) continue; }
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * I8_ACC_COLS);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smem_a + stage * a_bytes), p.lbo_units);
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smem_b + stage * b_bytes), p.lbo_units);
#pragma unroll
          for (int k = 0; k < I8_BK / I8_UK; ++k)
            tc_mma_i8_pair(tmem_d, adesc + (uint64_t)(k * (I8_UK >> 4)), bdesc + (uint64_t)(k * (I8_UK >> 4)), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit_pair(&empty[stage]);
          if (kb == p.num_k_blocks - 1) tc_commit_pair(&tfull[acc]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, each its own 128 rows) =====================
    const int ew = warp - 4;
    int acc = 0; uint32_t acc_phase = 0;
    const double w256 = 256.0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int m_blk, n_grp; tile_coords(p, tile, m_blk, n_grp);
      if (skip(m_blk)) continue;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int s = m_blk * 256 + (int)rank * I8_BM + ew * 32 + lane;
      const double rm = (p.mode == 2 && s < p.l) ? __ldg(p.row_mean + s) : 0.0;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * I8_ACC_COLS);
      const int i0 = n_grp * p.NE;
      if (p.mode == 1) {
        const int j0 = n_grp * p.N;
        double *krow = p.C + (size_t)s * p.ldc;
        for (int e0 = 0; e0 < p.N; e0 += 8) {
          int32_t d[8];
          tc_ld8(taddr + (uint32_t)e0, d);
          tc_ld_wait();
          if (s < p.l) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int j = j0 + e0 + q;
              if (j <= s) krow[j] += (double)d[q];
            }
          }
        }
      } else {
        for (int e0 = 0; e0 < p.NE; e0 += 8) {
          double v[8];
          int32_t d[8];
          tc_ld8(taddr + (uint32_t)e0, d);
          tc_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = (double)d[q];
          for (int t = 1; t < p.T; ++t) {
            tc_ld8(taddr + (uint32_t)(t * p.NE + e0), d);
            tc_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fma(v[q], w256, (double)d[q]);
          }
          if (s < p.l) {
            double *crow = p.C + (size_t)s * p.ldc;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int i = i0 + e0 + q;
              if (i < p.n) {
                const double val = v[q] * __ldg(p.scale + i);
                if (p.mode == 2) crow[i] = fma(rm, val, crow[i]); else crow[i] = val;     // mode 2: + mean * U^T q (hole indicator rows)
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // nobody tears TMEM / smem down while the peer still uses it
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// CTA-pair kernel, TWO eigenvector groups per tile: the pair's 256 x 128 B genotype tile of a K-block is staged once and
// multiplied with the plane rows of NB = 2 neighbouring groups (accumulators at TMEM columns 0 and 256: all 512 columns, single
// buffered).  The L2 -> shared-memory bytes per MAC drop from (256 + 240) to (256 + 480) / 2 rows per 256 x 240 x 128 MACs (-26 %):
// at 15 TB/s of operand traffic the projection sat on the L2 fabric (~6.3 kB/clk chip-wide), not on the tensor pipe.
// The price is an epilogue that no longer overlaps the next tile's MMAs (~3 % of a tile).
// mode 0: C = (A . planes) * scale          (A = genotypes with 0 at the holes)
// mode 2: C += row_mean * (A . planes) * scale   (A = hole-indicator rows: the mean imputation of src/lmm.cpp:1819-1827 as a second
//         GEMM pass on the same kernel; tiles whose 256 SNPs have no hole are skipped)
template <int NB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1)
i8_gemm_pair2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const I8KernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int halfN = p.N / 2;
  const int a_bytes = I8_BM * I8_BK;
  const int b_bytes = halfN * I8_BK;                       // this CTA's half of ONE group's plane rows
  const int stage_bytes = a_bytes + NB * b_bytes;
  uint64_t *bars = (uint64_t *)(smem + I8_STAGES * stage_bytes);
  uint64_t *full = bars, *empty = bars + I8_STAGES;
  uint64_t *tfull = bars + 2 * I8_STAGES, *tempty = bars + 2 * I8_STAGES + 1;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * I8_STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int n_units = (p.n_groups + NB - 1) / NB;
  const int num_tiles = p.m_tiles * n_units;               // m_tiles counts 256-row tiles

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 8);              // 4 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto coords = [&](int tile, int &m_blk, int &unit) {
    const int panel_tiles = p.panel * p.m_tiles;
    const int pn = tile / panel_tiles;
    const int first = pn * p.panel;
    const int width = (n_units - first < p.panel) ? (n_units - first) : p.panel;
    const int r = tile - pn * panel_tiles;
    m_blk = r / width;
    unit = first + r % width;
  };
  auto skip = [&](int m_blk) -> bool { return p.mode == 2 && p.tile_holes && __ldg(p.tile_holes + m_blk) == 0; };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, unit; coords(tile, m_blk, unit);
        if (skip(m_blk)) continue;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
          if (leader) mbar_expect_tx(&full[stage], (uint32_t)(2 * stage_bytes));
          uint8_t *st = smem + stage * stage_bytes;
          tma_load_2d_pair(&tmap_a, full_leader, st, kb * I8_BK, m_blk * 256 + (int)rank * I8_BM);
#pragma unroll
          for (int b = 0; b < NB; ++b)        // rows past the last group are out of bounds of the tensor map: zero-filled
            tma_load_2d_pair(&tmap_b, full_leader, st + a_bytes + b * b_bytes, kb * I8_BK, (unit * NB + b) * p.N + (int)rank * halfN);
          if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_i8_idesc(256, p.N, 0, 1);
      int stage = 0; uint32_t phase = 0, acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, unit; coords(tile, m_blk, unit);
        if (skip(m_blk)) continue;
        mbar_wait(tempty, acc_phase ^ 1);                   // the epilogue of the previous tile has drained TMEM
        tc_fence_after();
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_sw128_kmajor_desc(sa, p.lbo_units);
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const uint64_t bdesc = make_sw128_kmajor_desc(sa + a_bytes + b * b_bytes, p.lbo_units);
#pragma unroll
            for (int k = 0; k < I8_BK / I8_UK; ++k)
              tc_mma_i8_pair(tmem_base + (uint32_t)(b * I8_ACC_COLS), adesc + (uint64_t)(k * (I8_UK >> 4)), bdesc + (uint64_t)(k * (I8_UK >> 4)),
                             idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit_pair(&empty[stage]);
          if (kb == p.num_k_blocks - 1) tc_commit_pair(tfull);
          if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
        }
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, each its own 128 rows) =====================
    const int ew = warp - 4;
    uint32_t acc_phase = 0;
    const double w256 = 256.0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int m_blk, unit; coords(tile, m_blk, unit);
      if (skip(m_blk)) continue;
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
      const int s = m_blk * 256 + (int)rank * I8_BM + ew * 32 + lane;
      const double rm = (p.mode == 2 && s < p.l) ? __ldg(p.row_mean + s) : 0.0;
      double *crow = p.C + (size_t)s * p.ldc;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int i0 = (unit * NB + b) * p.NE;
        if (i0 >= p.n) break;
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(b * I8_ACC_COLS);
        for (int e0 = 0; e0 < p.NE; e0 += 8) {
          double v[8];
          int32_t d[8];
          tc_ld8(taddr + (uint32_t)e0, d);
          tc_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = (double)d[q];
          for (int t = 1; t < p.T; ++t) {
            tc_ld8(taddr + (uint32_t)(t * p.NE + e0), d);
            tc_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fma(v[q], w256, (double)d[q]);
          }
          if (s < p.l) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int i = i0 + e0 + q;
              if (i < p.n) {
                const double val = v[q] * __ldg(p.scale + i);
                if (p.mode == 2) crow[i] = fma(rm, val, crow[i]); else crow[i] = val;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(tempty), 0));
      acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// U -> int8 digit planes
// column maxima of a row-major matrix with n rows (individuals), ncols columns, leading dimension ld
__global__ void col_absmax_kernel(const double *__restrict__ U, int n, double *__restrict__ colmax, int ncols = -1, size_t ld = 0) {
  if (ncols < 0) { ncols = n; ld = (size_t)n; }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncols) return;
  double m = 0.0;
  for (int j = 0; j < n; ++j) m = fmax(m, fabs(U[(size_t)j * ld + i]));
  colmax[i] = m;
}

// scale[i] = s_i = colmax_i / (127.4 * 256^(T-1)), mult[i] = 1 / s_i  (so that Q = rint(u * mult), |Q| <= 127.4 * 256^(T-1)).
// An all-zero column gets mult = scale = 0.
__global__ void col_scale_kernel(const double *__restrict__ colmax, int n, int T, double *__restrict__ scale,
                                 double *__restrict__ mult) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = colmax[i];
  double top = 127.4;
  for (int t = 1; t < T; ++t) top *= 256.0;
  if (m > 0.0 && isfinite(m)) { mult[i] = top / m; scale[i] = m / top; }
  else { mult[i] = 0.0; scale[i] = 0.0; }
}

// max over the columns of colmax (one block)
__global__ void __launch_bounds__(256) vec_max_kernel(const double *__restrict__ v, int n, double *__restrict__ out) {
  __shared__ double sh[8];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) m = fmax(m, v[i]);
  for (int o = 16; o >= 1; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) { for (int w = 1; w < 8; ++w) m = fmax(m, sh[w]); *out = m; }
}

// one 32x32 tile of U per block: read U[j][i] coalesced in i, write planes coalesced in j
__global__ void __launch_bounds__(256) slice_kernel(const double *__restrict__ U, int n, const double *__restrict__ mult,
                                                    int T, int NE, int n_padk, int8_t *__restrict__ planes, int ncols = -1, size_t ld = 0) {
  if (ncols < 0) { ncols = n; ld = (size_t)n; }               // square U by default; rectangular (n individuals x ncols) for the exact-sum vectors
  __shared__ long long q[32][33];
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, i = i0 + tx;
    long long v = 0;
    if (j < n && i < ncols) v = llrint(U[(size_t)j * ld + i] * mult[i]);
    q[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + tx;      // thread writes column (eigenvector) i, individual j
    if (i >= ncols || j >= n) continue;
    long long Q = q[tx][r];
    const int g = i / NE, e = i - g * NE;
    const size_t row0 = (size_t)g * (size_t)(T * NE) + (size_t)e;
    for (int t = T - 1; t >= 1; --t) {
      const long long d = ((Q + 128) & 255) - 128;          // balanced digit in [-128,127]
      Q = (Q - d) >> 8;
      planes[(row0 + (size_t)t * NE) * (size_t)n_padk + j] = (int8_t)d;
    }
    planes[row0 * (size_t)n_padk + j] = (int8_t)Q;           // |Q| <= 127: 127.4 + the carry of the balanced lower digits (< 0.51)
  }
}

// PLINK 2-bit rows -> int8 genotype rows (0 at missing / padding) + per-SNP mean and hole count
__global__ void __launch_bounds__(256) bed_to_i8_kernel(const unsigned char *__restrict__ bed, size_t bytes_per_snp,
                                                        const int *__restrict__ idx, int n, int n_padk, int l,
                                                        int8_t *__restrict__ G, double *__restrict__ mean,
                                                        int *__restrict__ nmiss, int8_t *__restrict__ Q,
                                                        int *__restrict__ tile_holes,
                                                        const double *__restrict__ vnull, int nq, int ldv, double *__restrict__ xex) {
  __shared__ int sh_sum[8], sh_miss[8];
  __shared__ double sh_x[8][2 * 4];
  // exact x-sums (LmmConst::xex): x . v_q over the observed genotypes and sum of v_q over the holes (the mean joins at the end)
  double ex[4] = {0.0, 0.0, 0.0, 0.0}, eh[4] = {0.0, 0.0, 0.0, 0.0};
  const int s = blockIdx.x;
  int8_t *g = G + (size_t)s * n_padk;
  int8_t *qr = Q ? Q + (size_t)s * n_padk : nullptr;     // hole-indicator row (second GEMM pass of the mean imputation)
  if (s >= l) {                                   // zero padding rows of the last tile
    for (int p = threadIdx.x; p < n_padk; p += 256) { g[p] = 0; if (qr) qr[p] = 0; }
    return;
  }
  const unsigned char *row = bed + (size_t)s * bytes_per_snp;
  int sum = 0, miss = 0;
  for (int p = threadIdx.x; p < n_padk; p += 256) {
    int8_t v = 0, hq = 0;
    if (p < n) {
      const size_t j = idx ? (size_t)idx[p] : (size_t)p;
      const unsigned b = (unsigned)row[j >> 2] >> (2 * (j & 3));
      const unsigned lo = b & 1u, hi = (b >> 1) & 1u;
      if (lo == 0) v = hi == 0 ? 2 : 1;
      else if (hi == 0) { miss++; hq = 1; }
      sum += v;
      if (vnull && (v | hq)) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nq) { const double t = __ldg(vnull + (size_t)q * ldv + p); if (hq) eh[q] += t; else ex[q] = fma((double)v, t, ex[q]); }
      }
    }
    g[p] = v;
    if (qr) qr[p] = hq;
  }
  if (vnull) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double a = ex[q], b = eh[q];
      for (int m = 16; m >= 1; m >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, m); b += __shfl_xor_sync(0xffffffffu, b, m); }
      if ((threadIdx.x & 31) == 0) { sh_x[threadIdx.x >> 5][2 * q] = a; sh_x[threadIdx.x >> 5][2 * q + 1] = b; }
    }
  }
  for (int m = 16; m >= 1; m >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, m); miss += __shfl_xor_sync(0xffffffffu, miss, m); }
  if ((threadIdx.x & 31) == 0) { sh_sum[threadIdx.x >> 5] = sum; sh_miss[threadIdx.x >> 5] = miss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ts = 0, tm = 0;
    for (int w = 0; w < 8; ++w) { ts += sh_sum[w]; tm += sh_miss[w]; }
    nmiss[s] = tm;
    mean[s] = (double)ts / (double)(n - tm);          // x_mean of src/lmm.cpp:1819
    if (tile_holes && tm) atomicAdd(tile_holes + (s >> 8), tm);
    if (vnull) {
      const double mu = (double)ts / (double)(n - tm);
      for (int q = 0; q < nq && q < 4; ++q) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh_x[w][2 * q]; b += sh_x[w][2 * q + 1]; }
        xex[(size_t)s * nq + q] = fma(mu, b, a);
      }
    }
  }
}

// Holes either way: few holes -> gather the rows of U at the holes (miss_fix_kernel: holes x n x 8 B of reads, 400 KB per hole at
// n = 50 000); many holes -> a second pass of the projection GEMM over the hole-indicator rows (mode 2: C += mean * U^T q, same
// cost as the main pass, independent of the hole count).  The gather costs holes x 8 n B / 6.5 TB/s, the GEMM pass about
// 2 T n^2 l / 3e15 s: the switch sits where the gather would take longer.  Decided on the device (no read-back).
__global__ void hole_switch_kernel(const int *__restrict__ tile_holes, int n_tiles, double holes_max, int *__restrict__ sw) {
  __shared__ double part[8];
  double t = 0.0;
  for (int k = threadIdx.x; k < n_tiles; k += 256) t += (double)tile_holes[k];
  for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) { for (int w = 1; w < 8; ++w) t += part[w]; sw[0] = (t > holes_max) ? 1 : 0; }
}

// U^T x += mean * sum_{j in holes} U[j][:]   for SNPs with missing genotypes
__global__ void __launch_bounds__(256) miss_fix_kernel(const unsigned char *__restrict__ bed, size_t bytes_per_snp,
                                                       const int *__restrict__ idx, int n,
                                                       const double *__restrict__ U, const double *__restrict__ mean,
                                                       const int *__restrict__ nmiss, double *__restrict__ C, size_t ldc,
                                                       const int *__restrict__ hole_switch, int ncols = -1, size_t ldu = 0) {
  if (ncols < 0) { ncols = n; ldu = (size_t)n; }             // rows of U (n columns) by default; rows of the exact-sum vectors V (ncols) otherwise
  constexpr int CAP = 2048;
  __shared__ int list[CAP];
  __shared__ int wcount[8];
  __shared__ int count;
  const int s = blockIdx.x;
  if (nmiss[s] == 0) return;
  if (hole_switch && __ldg(hole_switch) != 0) return;       // dense holes: the tensor-core hole pass has them
  const unsigned char *row = bed + (size_t)s * bytes_per_snp;
  const double m = mean[s];
  double *c = C + (size_t)s * ldc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // The holes of the SNP are collected in ascending order (ballot compaction: the summation order below is fixed) up to CAP at a
  // time -- in practice all of them: the gather only serves batches with few holes -- and the rows of U at the holes are then summed
  // in ONE sweep over the eigenvectors, 8 independent row loads in flight per thread.
  int start = 0;
  while (start < n) {
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    int p0 = start;
    for (; p0 < n; p0 += 256) {
      if (count + 256 > CAP) break;                          // uniform: `count` only changes between the barriers below
      const int p = p0 + threadIdx.x;
      bool hole = false;
      if (p < n) {
        const size_t j = idx ? (size_t)idx[p] : (size_t)p;
        const unsigned b = (unsigned)row[j >> 2] >> (2 * (j & 3));
        hole = (b & 3u) == 1u;                               // low = 1, high = 0 -> missing
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hole);
      if (lane == 0) wcount[warp] = __popc(bal);
      __syncthreads();
      int off = count;
      for (int w = 0; w < warp; ++w) off += wcount[w];
      if (hole) list[off + __popc(bal & ((1u << lane) - 1u))] = p;
      __syncthreads();
      if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += wcount[w]; count += t; }
      __syncthreads();
    }
    const int cnt = count;
    if (cnt > 0) {
      for (int i = threadIdx.x; i < ncols; i += 256) {
        double acc = 0.0;
        int q = 0;
        for (; q + 8 <= cnt; q += 8) {
          double v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __ldg(U + (size_t)list[q + k] * ldu + i);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc += v[k];           // left to right: the order of the holes
        }
        for (; q < cnt; ++q) acc += __ldg(U + (size_t)list[q] * ldu + i);
        c[i] += m * acc;
      }
    }
    __syncthreads();
    start = p0;
  }
}

// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

static bool make_tmap(CUtensorMap *tm, const void *base, uint64_t rows, uint64_t row_bytes, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {row_bytes, rows};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)I8_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool i8_available(gb200_ctx *) { return get_encode() != nullptr; }

// Number of int8 digit planes of U.  The planes keep every entry to within s_i / 2, s_i = colmax_i / (127.4 * 256^(T-1)); the
// dropped tails act as independent rounding noise, so a projected value (u_i . x) -- of typical size rms(x) for a unit vector --
// carries a relative error of about  colmax_i sqrt(n) / (sqrt(12) * 127.4 * 256^(T-1)).
//  * i8_default_planes(n): the U-independent worst case (colmax_i = 1, an eigenvector concentrated on one individual):
//    smallest T with that bound <= 2^-30 -- T = 5 up to n = 3.1e6.  Used when the planes are requested before U is known.
//  * i8_choose_planes(colmax, n): the same bound evaluated with the MEASURED largest column maximum of this U, target 2^-29
//    (1.9e-9 on a projected value; beta inherits it as 1.9e-9 x se, i.e. relative 1.9e-9 / |z-score| -- inside the 1e-6 parity
//    bar for every SNP with |z| > 0.002 -- and se / p-values at the 1e-9 level).  Eigenvectors of a kinship matrix of unrelated
//    individuals are delocalised (colmax ~ sqrt(4 ln n / n)): T = 4; family / population structure that concentrates an eigenvector
//    on few individuals raises colmax and with it T.  Small cohorts (n < 8192), where the projection is a minor cost, never go
//    below 5 planes.
int i8_default_planes(size_t n) {
  const double need = sqrt((double)(n > 1 ? n : 2)) / (sqrt(12.0) * 127.4) * 1073741824.0;     // bound * 2^30 at T = 1
  int T = 1 + (int)ceil(log2(need) / 8.0);
  if (T < 4) T = 4;
  if (T > 8) T = 8;
  return T;
}
int i8_choose_planes(double colmax_max, size_t n, bool linear_sums_exact = false) {
  if (!(colmax_max > 0.0) || !isfinite(colmax_max)) return 4;
  // with the exact linear x-sums (LmmConst::xsum) the projected U^T x only feeds sums QUADRATIC in x, where independent rounding noise
  // eps averages to eps / sqrt(n): the target relaxes from 2^-29 to 2^-21 and delocalised eigenvectors get by with 3 planes
  const double target = linear_sums_exact ? 2097152.0 : 536870912.0;                                      // 2^21 | 2^29
  const double need = colmax_max * sqrt((double)(n > 1 ? n : 2)) / (sqrt(12.0) * 127.4) * target;        // bound / target at T = 1
  int T = 1 + (int)ceil(log2(need) / 8.0);
  const int Tmin = linear_sums_exact ? 3 : 4;
  if (T < Tmin) T = Tmin;
  if (n < 8192 && T < 5) T = 5;
  if (T > 8) T = 8;
  return T;
}

// largest |entry| of U's columns -> the plane count i8_prepare will use (cached until the next setup)
int i8_effective_planes(gb200_ctx *c, int *T_out) {
  if (c->n_slices > 0) { *T_out = (int)c->n_slices; return GB200_OK; }
  if (!c->lmm_ready) { *T_out = c->n ? i8_default_planes(c->n) : 0; return GB200_OK; }
  if (c->i8.auto_T == 0) {
    DevBuf tmp;
    GB_CUDA(c, tmp.reserve(((size_t)c->n + 1) * sizeof(double)));
    col_absmax_kernel<<<((int)c->n + 255) / 256, 256, 0, c->stream>>>(c->dU.as<double>(), (int)c->n, tmp.as<double>());
    vec_max_kernel<<<1, 256, 0, c->stream>>>(tmp.as<double>(), (int)c->n, tmp.as<double>() + c->n);
    double m = 0.0;
    GB_CUDA(c, cudaMemcpyAsync(&m, tmp.as<double>() + c->n, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    tmp.release();
    c->i8.colmax_max = m;
    c->i8.auto_T = i8_choose_planes(m, c->n, c->i8.xs_ready && c->x_exact == 2);
  }
  *T_out = c->i8.auto_T;
  if (c->i8.no_xsum_consumer && *T_out < 4) *T_out = 4;     // a consumer without the exact linear sums (dosage rows) has used this context
  return GB200_OK;
}

int i8_prepare(gb200_ctx *c) {
  if (!c->lmm_ready) return set_err(c, GB200_ERR_STATE, "i8_prepare before lmm_setup");
  int T = 0;
  { const int rc = i8_effective_planes(c, &T); if (rc) return rc; }
  if (T < 2 || T > 8) return set_err(c, GB200_ERR_ARG, "n_slices must be in 2..8");
  if (c->i8.ready && c->i8.n_slices == T && c->i8.n == c->n) return GB200_OK;
  const I8Geom g = make_geom(c->n, T);
  const size_t rows = (size_t)g.n_groups * (size_t)g.N;
  const size_t bytes = rows * (size_t)g.n_padk;
  GB_CUDA(c, c->i8.slices.reserve(bytes));
  GB_CUDA(c, c->i8.scale.reserve((size_t)g.n * 3 * sizeof(double)));
  GB_CUDA(c, cudaMemsetAsync(c->i8.slices.p, 0, bytes, c->stream));
  double *scale = c->i8.scale.as<double>();
  double *colmax = scale + g.n;
  double *mult = colmax + g.n;
  col_absmax_kernel<<<(g.n + 255) / 256, 256, 0, c->stream>>>(c->dU.as<double>(), g.n, colmax);
  col_scale_kernel<<<(g.n + 255) / 256, 256, 0, c->stream>>>(colmax, g.n, T, scale, mult);
  dim3 grid((g.n + 31) / 32, (g.n + 31) / 32);
  slice_kernel<<<grid, 256, 0, c->stream>>>(c->dU.as<double>(), g.n, mult, T, g.NE, g.n_padk, c->i8.slices.as<int8_t>());
  GB_CUDA(c, cudaGetLastError());
  if (!c->i8.tmap_a) c->i8.tmap_a = aligned_alloc(64, sizeof(CUtensorMap));
  if (!c->i8.tmap_b) c->i8.tmap_b = aligned_alloc(64, sizeof(CUtensorMap));
  if (!make_tmap((CUtensorMap *)c->i8.tmap_b, c->i8.slices.p, rows, (uint64_t)g.n_padk, (uint32_t)g.N))
    return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the U planes");
  c->i8.n = c->n; c->i8.n_pad = (size_t)g.n_padk; c->i8.n_slices = T; c->i8.ready = true; c->i8.tmap_b_half = false;
  return GB200_OK;
}

// U^T x entries of the eigenvectors listed in idx <- their exact values from the side GEMM (columns col0.. of xs)
__global__ void xs_patch_kernel(const double *__restrict__ xs, size_t ld, int col0, const int *__restrict__ idx, int npatch, int l,
                                double *__restrict__ UtXt, size_t ldu) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= l * npatch) return;
  const int s = t / npatch, k = t - s * npatch;
  UtXt[(size_t)s * ldu + (size_t)__ldg(idx + k)] = xs[(size_t)s * ld + col0 + k];
}

// Digit planes (6) and tensor map of the exact-sum vectors V (n individuals x ncol, row-major in c->i8.xs_V): the side GEMM G . V of
// every int8-projected batch gives the sums linear in x at all hoisted lambdas (LmmConst::xsum).
int i8_xsum_prepare(gb200_ctx *c, int ncol) {
  I8State &S = c->i8;
  S.xs_ready = false;
  const int T = 6, NE = 40, N = T * NE;                       // same tile shape as the 6-plane projection (N = 240)
  const int n = (int)c->n, n_padk = (int)((c->n + I8_BK - 1) / I8_BK * I8_BK);
  const int groups = (ncol + NE - 1) / NE;
  const size_t rows = (size_t)groups * N, bytes = rows * (size_t)n_padk;
  GB_CUDA(c, S.xs_planes.reserve(bytes));
  GB_CUDA(c, S.xs_scale.reserve((size_t)ncol * 3 * sizeof(double)));
  GB_CUDA(c, cudaMemsetAsync(S.xs_planes.p, 0, bytes, c->stream));
  double *scale = S.xs_scale.as<double>(), *colmax = scale + ncol, *mult = colmax + ncol;
  col_absmax_kernel<<<(ncol + 255) / 256, 256, 0, c->stream>>>(S.xs_V.as<double>(), n, colmax, ncol, (size_t)ncol);
  col_scale_kernel<<<(ncol + 255) / 256, 256, 0, c->stream>>>(colmax, ncol, T, scale, mult);
  dim3 grid((ncol + 31) / 32, (n + 31) / 32);
  slice_kernel<<<grid, 256, 0, c->stream>>>(S.xs_V.as<double>(), n, mult, T, NE, n_padk, S.xs_planes.as<int8_t>(), ncol, (size_t)ncol);
  GB_CUDA(c, cudaGetLastError());
  if (!S.tmap_v) S.tmap_v = aligned_alloc(64, sizeof(CUtensorMap));
  if (!make_tmap((CUtensorMap *)S.tmap_v, S.xs_planes.p, rows, (uint64_t)n_padk, (uint32_t)(N / 2)))
    return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the exact-sum planes");
  S.xs_ncol = ncol; S.xs_T = T; S.xs_NE = NE; S.xs_groups = groups; S.xs_ld = (size_t)((ncol + 7) / 8 * 8);
  S.xs_ready = true;
  return GB200_OK;
}

int i8_project_bed(gb200_ctx *c, const unsigned char *bed_dev, const int *idx_dev, size_t ni_total, size_t l,
                   size_t bytes_per_snp, double *UtXt_dev) {
  (void)ni_total;
  int rc = i8_prepare(c);
  if (rc) return rc;
  const I8Geom g = make_geom(c->n, c->i8.n_slices);
  const bool pair = c->cta_pair != 0 && (g.N % 32 == 0 || g.N == 240);     // half of N must stay a multiple of 8 rows
  const bool pair2 = pair && c->gemm_groups == 2;                          // two eigenvector groups per tile + hole pass on the tensor pipe
  const size_t m_rows = pair ? 256 : I8_BM;
  const size_t l_pad = (l + m_rows - 1) / m_rows * m_rows;
  GB_CUDA(c, c->i8.geno.reserve(l_pad * (size_t)g.n_padk));
  GB_CUDA(c, c->i8.miss_mean.reserve(l_pad * (sizeof(double) + sizeof(int)) + (l_pad / 256 + 2) * sizeof(int)));
  double *mean = c->i8.miss_mean.as<double>();
  int *nmiss = reinterpret_cast<int *>(mean + l_pad);
  int *tile_holes = nmiss + l_pad;
  const bool hole_gemm = pair && !pair2 && c->hole_gemm != 0;   // hybrid: gather when holes are few, tensor-core hole pass when they are many
  if (pair2 || hole_gemm) {
    GB_CUDA(c, c->i8.holeq.reserve(l_pad * (size_t)g.n_padk));
    GB_CUDA(c, cudaMemsetAsync(tile_holes, 0, (l_pad / 256 + 2) * sizeof(int), c->stream));
  }
  const int nq = (int)c->n_cvt + 1;
  const bool want_xsum = c->i8.xs_ready && c->x_exact == 2 && pair && !pair2 && !c->overlap;
  const bool want_xex = !want_xsum && c->vnull_ready && c->x_exact && nq <= 4 && !c->overlap;
  c->i8.xs_valid = false;
  c->i8.xex_valid = false;
  if (want_xex) GB_CUDA(c, c->i8.xex.reserve(l_pad * (size_t)nq * sizeof(double)));
  {
  ProfScope ps(c, "decode");
  bed_to_i8_kernel<<<(unsigned)l_pad, 256, 0, c->stream>>>(bed_dev, bytes_per_snp, idx_dev, g.n, g.n_padk, (int)l,
                                                           c->i8.geno.as<int8_t>(), mean, nmiss,
                                                           (pair2 || hole_gemm) ? c->i8.holeq.as<int8_t>() : nullptr,
                                                           (pair2 || hole_gemm) ? tile_holes : nullptr,
                                                           want_xex ? c->dVnull.as<double>() : nullptr, nq, (int)c->n_c,
                                                           want_xex ? c->i8.xex.as<double>() : nullptr);
  GB_CUDA(c, cudaGetLastError());
  }
  if (want_xex) { c->i8.xex_valid = true; c->i8.xex_for = UtXt_dev; c->i8.xex_l = l; }
  if (!make_tmap((CUtensorMap *)c->i8.tmap_a, c->i8.geno.p, l_pad, (uint64_t)g.n_padk, (uint32_t)I8_BM))
    return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the genotype tile");
  I8KernelParams p;
  p.T = g.T; p.NE = g.NE; p.N = g.N; p.n = g.n; p.l = (int)l;
  p.num_k_blocks = g.n_padk / I8_BK;
  p.m_tiles = (int)(l_pad / m_rows); p.n_groups = g.n_groups;
  p.lbo_units = 1;
  p.scale = c->i8.scale.as<double>();
  p.C = UtXt_dev; p.ldc = c->n_c; p.mode = 0; p.tiles = nullptr; p.num_tiles = 0;
  p.row_mean = nullptr; p.tile_holes = nullptr; p.panel = c->gemm_panel > 0 ? (int)c->gemm_panel : (pair2 ? 6 : (pair ? 9 : I8_PANEL)); p.stages = I8_STAGES; p.wave_ctr = nullptr; p.hole_switch = nullptr; p.l2_hint = 0;
  const size_t smem = 1024 + (size_t)I8_STAGES * (I8_BM * I8_BK + (size_t)g.N * I8_BK) + 256;
  const int tiles = p.m_tiles * p.n_groups;
  if (pair2) {
    // B tensor map with a half-N box; the SAME kernel runs twice: genotypes (store), then hole indicators (accumulate with the SNP mean)
    if (!make_tmap((CUtensorMap *)c->i8.tmap_b, c->i8.slices.p, (uint64_t)g.n_groups * (uint64_t)g.N, (uint64_t)g.n_padk, (uint32_t)(g.N / 2)))
      return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the U planes (pair)");
    c->i8.tmap_b_half = true;
    if (!c->i8.tmap_q) c->i8.tmap_q = aligned_alloc(64, sizeof(CUtensorMap));
    if (!make_tmap((CUtensorMap *)c->i8.tmap_q, c->i8.holeq.p, l_pad, (uint64_t)g.n_padk, (uint32_t)I8_BM))
      return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the hole-indicator tile");
    GB_CUDA(c, cudaFuncSetAttribute(i8_gemm_pair2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int units = (p.n_groups + 1) / 2;
    int pairs = c->num_sms / 2; if (pairs > p.m_tiles * units) pairs = p.m_tiles * units; if (pairs < 1) pairs = 1;
    const size_t smem2 = 1024 + (size_t)I8_STAGES * (I8_BM * I8_BK + (size_t)g.N * I8_BK) + 256;      // A + 2 x half of B per stage
    {
      ProfScope ps(c, "utx");
      i8_gemm_pair2_kernel<2><<<2 * pairs, I8_THREADS, smem2, c->stream>>>(*(CUtensorMap *)c->i8.tmap_a, *(CUtensorMap *)c->i8.tmap_b, p);
      GB_CUDA(c, cudaGetLastError());
    }
    p.mode = 2; p.row_mean = mean; p.tile_holes = tile_holes;
    ProfScope ps2(c, "fix");
    i8_gemm_pair2_kernel<2><<<2 * pairs, I8_THREADS, smem2, c->stream>>>(*(CUtensorMap *)c->i8.tmap_q, *(CUtensorMap *)c->i8.tmap_b, p);
    GB_CUDA(c, cudaGetLastError());
    return GB200_OK;
  }
  GB_CUDA(c, cudaFuncSetAttribute(i8_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (pair) {
    // B tensor map with a half-N box (each CTA of the pair loads its own half of the plane rows)
    if (!make_tmap((CUtensorMap *)c->i8.tmap_b, c->i8.slices.p, (uint64_t)g.n_groups * (uint64_t)g.N, (uint64_t)g.n_padk, (uint32_t)(g.N / 2)))
      return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the U planes (pair)");
    c->i8.tmap_b_half = true;
    GB_CUDA(c, cudaFuncSetAttribute(i8_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int pairs = c->num_sms / 2; if (pairs > tiles) pairs = tiles; if (pairs < 1) pairs = 1;
    const size_t stage_pair = (size_t)I8_BM * I8_BK + (size_t)(g.N / 2) * I8_BK;                            // A + half of B per stage
    int ns = c->gemm_stages > 0 ? (int)c->gemm_stages : 6;
    while (ns > 2 && 1024 + (size_t)ns * stage_pair + 256 > 227 * 1024) --ns;
    p.stages = ns; p.l2_hint = (int)c->gemm_l2hint;
    if (c->gemm_wave_sync && 2 * pairs == c->num_sms / 2 * 2 && tiles > pairs) {       // whole-chip persistent grid: every CTA is resident
      GB_CUDA(c, c->i8.wave_ctr.reserve(sizeof(unsigned int)));
      GB_CUDA(c, cudaMemsetAsync(c->i8.wave_ctr.p, 0, sizeof(unsigned int), c->stream));
      p.wave_ctr = c->i8.wave_ctr.as<unsigned int>();
    }
    const size_t smem_pair = 1024 + (size_t)ns * stage_pair + 256;
    {
      ProfScope ps(c, "utx");
      i8_gemm_pair_kernel<<<2 * pairs, I8_THREADS, smem_pair, c->stream>>>(*(CUtensorMap *)c->i8.tmap_a, *(CUtensorMap *)c->i8.tmap_b, p);
      GB_CUDA(c, cudaGetLastError());
    }
    if (want_xsum) {
      // side GEMM: the same genotype tiles against the 6 planes of the exact-sum vectors, then mean * (sum of V rows at the holes)
      I8State &S = c->i8;
      GB_CUDA(c, S.xs_out.reserve(l_pad * S.xs_ld * sizeof(double)));
      I8KernelParams q = p;
      q.T = S.xs_T; q.NE = S.xs_NE; q.N = S.xs_T * S.xs_NE; q.n = S.xs_ncol; q.n_groups = S.xs_groups;
      q.scale = S.xs_scale.as<double>(); q.C = S.xs_out.as<double>(); q.ldc = S.xs_ld; q.mode = 0; q.wave_ctr = nullptr; q.panel = 4;
      const size_t stage_v = (size_t)I8_BM * I8_BK + (size_t)(q.N / 2) * I8_BK;
      int nsv = 6; while (nsv > 2 && 1024 + (size_t)nsv * stage_v + 256 > 227 * 1024) --nsv;
      q.stages = nsv;
      const int tiles_v = q.m_tiles * q.n_groups;
      int pairs_v = c->num_sms / 2; if (pairs_v > tiles_v) pairs_v = tiles_v; if (pairs_v < 1) pairs_v = 1;
      {
        ProfScope ps(c, "xsum");
        i8_gemm_pair_kernel<<<2 * pairs_v, I8_THREADS, 1024 + (size_t)nsv * stage_v + 256, c->stream>>>(*(CUtensorMap *)S.tmap_a, *(CUtensorMap *)S.tmap_v, q);
        GB_CUDA(c, cudaGetLastError());
        miss_fix_kernel<<<(unsigned)l, 256, 0, c->stream>>>(bed_dev, bytes_per_snp, idx_dev, g.n, S.xs_V.as<double>(), mean, nmiss, S.xs_out.as<double>(),
                                                            S.xs_ld, nullptr, S.xs_ncol, (size_t)S.xs_ncol);
        GB_CUDA(c, cudaGetLastError());
      }
      S.xs_valid = true; S.xs_for = UtXt_dev; S.xs_l = l;
    }
    if (hole_gemm) {
      int *sw = tile_holes + (l_pad / 256 + 1);
      // gather: holes * 8 n bytes at ~4 TB/s;  hole pass: the main pass again, ~2 T n^2 l_pad / 2.8e15 s
      const double t_gemm = 2.0 * (double)g.T * (double)g.n * (double)g.n * (double)l_pad / 2.8e15;
      const double holes_max = t_gemm * 4.0e12 / (8.0 * (double)g.n);
      hole_switch_kernel<<<1, 256, 0, c->stream>>>(tile_holes, (int)(l_pad / 256), holes_max, sw);
      if (!c->i8.tmap_q) c->i8.tmap_q = aligned_alloc(64, sizeof(CUtensorMap));
      if (!make_tmap((CUtensorMap *)c->i8.tmap_q, c->i8.holeq.p, l_pad, (uint64_t)g.n_padk, (uint32_t)I8_BM))
        return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the hole-indicator tile");
      I8KernelParams ph = p;
      ph.mode = 2; ph.row_mean = mean; ph.tile_holes = tile_holes; ph.hole_switch = sw; ph.wave_ctr = nullptr;
      ProfScope ps(c, "fix");
      i8_gemm_pair_kernel<<<2 * pairs, I8_THREADS, smem_pair, c->stream>>>(*(CUtensorMap *)c->i8.tmap_q, *(CUtensorMap *)c->i8.tmap_b, ph);
      GB_CUDA(c, cudaGetLastError());
      miss_fix_kernel<<<(unsigned)l, 256, 0, c->stream>>>(bed_dev, bytes_per_snp, idx_dev, g.n, c->dU.as<double>(), mean, nmiss, UtXt_dev, c->n_c, sw);
      GB_CUDA(c, cudaGetLastError());
      if (want_xsum && c->i8.xs_npatch > 0) {
        xs_patch_kernel<<<(unsigned)((l * (size_t)c->i8.xs_npatch + 255) / 256), 256, 0, c->stream>>>(c->i8.xs_out.as<double>(), c->i8.xs_ld, c->i8.xs_patch0,
            c->i8.xs_patch_idx.as<int>(), c->i8.xs_npatch, (int)l, UtXt_dev, c->n_c);
        GB_CUDA(c, cudaGetLastError());
      }
      return GB200_OK;
    }
  } else {
    if (c->i8.tmap_b_half) {
      if (!make_tmap((CUtensorMap *)c->i8.tmap_b, c->i8.slices.p, (uint64_t)g.n_groups * (uint64_t)g.N, (uint64_t)g.n_padk, (uint32_t)g.N))
        return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the U planes");
      c->i8.tmap_b_half = false;
    }
    const int grid = tiles < c->num_sms ? tiles : c->num_sms;
    ProfScope ps(c, "utx");
    i8_gemm_kernel<<<grid, I8_THREADS, smem, c->stream>>>(*(CUtensorMap *)c->i8.tmap_a, *(CUtensorMap *)c->i8.tmap_b, p);
    GB_CUDA(c, cudaGetLastError());
  }
  ProfScope ps2(c, "fix");
  miss_fix_kernel<<<(unsigned)l, 256, 0, c->stream>>>(bed_dev, bytes_per_snp, idx_dev, g.n, c->dU.as<double>(), mean,
                                                      nmiss, UtXt_dev, c->n_c, nullptr);
  GB_CUDA(c, cudaGetLastError());
  if (want_xsum && c->i8.xs_npatch > 0) {
    xs_patch_kernel<<<(unsigned)((l * (size_t)c->i8.xs_npatch + 255) / 256), 256, 0, c->stream>>>(c->i8.xs_out.as<double>(), c->i8.xs_ld, c->i8.xs_patch0,
        c->i8.xs_patch_idx.as<int>(), c->i8.xs_npatch, (int)l, UtXt_dev, c->n_c);
    GB_CUDA(c, cudaGetLastError());
  }
  return GB200_OK;
}

// ==========================================================================================
// Dosage-valued (BIMBAM mean-genotype) batches on the same tensor-core projection.  The reference hands the parsed doubles to
// cblas_dgemm (src/lmm.cpp:1590-1618, :1521).  A mean-genotype file prints its values with a few decimals (the mouse example: 0/1/2
// with an occasional "1.03" or "0.4375"), so every non-missing entry of a SNP row is q / 10^d for an integer q and a small d.  The
// row is then EXACTLY representable by one to three unsigned base-256 digit rows (q < 256, < 65536, < 2^24), each of which goes
// through the int8 GEMM like a PLINK genotype row; the epilogue pass recombines  U^T x = (sum_k 256^k U^T q_k) / 10^d  in FP64 and
// the holes get their  mean * sum_{j missing} U[j][:]  like the PLINK path.  Rows that are plain 0/1/2 cost what a PLINK row costs.
// A batch that contains a value with more than 6 decimals (or a negative one) is not taken: the caller keeps the FP64 GEMM.
__global__ void __launch_bounds__(256) geno_classify_kernel(const double *__restrict__ G, int n, size_t ldg, double *__restrict__ mean,
                                                            int *__restrict__ nmiss, int *__restrict__ dsel, int *__restrict__ nplanes) {
  __shared__ double sh_sum[8], sh_max[8];
  __shared__ int sh_miss[8], sh_d[8];
  const int s = blockIdx.x;
  const double *g = G + (size_t)s * ldg;
  double sum = 0.0, xmax = 0.0;
  int miss = 0, dreq = 0;                       // dreq = 7: not representable with <= 6 decimals
  for (int p = threadIdx.x; p < n; p += 256) {
    const double x = g[p];
    if (isnan(x)) { miss++; continue; }
    sum += x;
    if (!(x >= 0.0) || !(x < 16.0)) { dreq = 7; continue; }
    xmax = fmax(xmax, x);
    double sc = 1.0;
    int d = 0;
    for (; d <= 6; ++d, sc *= 10.0) {
      const double v = x * sc;
      if (fabs(v - rint(v)) <= 1e-8) break;
    }
    dreq = d > dreq ? d : dreq;
  }
  for (int m = 16; m >= 1; m >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, m); xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, m));
    miss += __shfl_xor_sync(0xffffffffu, miss, m); const int o = __shfl_xor_sync(0xffffffffu, dreq, m); dreq = o > dreq ? o : dreq;
  }
  if ((threadIdx.x & 31) == 0) { sh_sum[threadIdx.x >> 5] = sum; sh_max[threadIdx.x >> 5] = xmax; sh_miss[threadIdx.x >> 5] = miss; sh_d[threadIdx.x >> 5] = dreq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tx = 0.0; int tm = 0, td = 0;
    for (int w = 0; w < 8; ++w) { ts += sh_sum[w]; tx = fmax(tx, sh_max[w]); tm += sh_miss[w]; td = sh_d[w] > td ? sh_d[w] : td; }
    nmiss[s] = tm;
    mean[s] = ts / (double)(n - tm);                 // x_mean of src/lmm.cpp:1599-1604 (NaN for an all-missing row, as there)
    int T = 0;
    if (td <= 6) {
      double sc = 1.0; for (int d = 0; d < td; ++d) sc *= 10.0;
      const double qmax = rint(tx * sc);
      T = qmax < 256.0 ? 1 : (qmax < 65536.0 ? 2 : (qmax < 16777216.0 ? 3 : 0));
    }
    dsel[s] = td; nplanes[s] = T;                    // T == 0: this batch stays on the FP64 path
  }
}

// exclusive scan of nplanes over one chunk of SNPs -> digit-row offsets; info = {rows, #rows not representable, all rows plain (d = 0, one plane)}
__global__ void __launch_bounds__(1024) geno_rowoff_kernel(const int *__restrict__ nplanes, const int *__restrict__ dsel, int l, int *__restrict__ rowoff,
                                                           int *__restrict__ info) {
  __shared__ int part[1024];
  const int per = (l + 1023) / 1024, b0 = threadIdx.x * per;
  int t = 0, bad = 0, fancy = 0;
  for (int k = 0; k < per && b0 + k < l; ++k) { const int T = nplanes[b0 + k]; t += T; bad += (T == 0); fancy += (T != 1 || dsel[b0 + k] != 0); }
  part[threadIdx.x] = t;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int off = part[threadIdx.x] - t;
  for (int k = 0; k < per && b0 + k < l; ++k) { rowoff[b0 + k] = off; off += nplanes[b0 + k]; }
  if (bad) atomicAdd(&info[1], bad);
  if (fancy) atomicAdd(&info[2], fancy);
  if (threadIdx.x == 1023) info[0] = part[1023];
}

__global__ void __launch_bounds__(256) geno_to_planes_kernel(const double *__restrict__ G, int n, size_t ldg, int n_padk, const int *__restrict__ dsel,
                                                             const int *__restrict__ nplanes, const int *__restrict__ rowoff, uint8_t *__restrict__ A) {
  const int s = blockIdx.x;
  const int T = nplanes[s], d = dsel[s];
  const double *g = G + (size_t)s * ldg;
  double sc = 1.0;
  for (int k = 0; k < d; ++k) sc *= 10.0;
  uint8_t *a0 = A + (size_t)rowoff[s] * n_padk;
  for (int p = threadIdx.x; p < n_padk; p += 256) {
    unsigned q = 0u;
    if (p < n) { const double x = g[p]; if (!isnan(x)) q = (unsigned)rint(x * sc); }      // holes enter as 0; their mean term is added afterwards
    for (int k = 0; k < T; ++k) a0[(size_t)k * n_padk + p] = (uint8_t)((q >> (8 * k)) & 255u);
  }
}

// U^T x = (sum_k 256^k C_k) / 10^d  +  mean * sum_{j missing} U[j][:]      (one CTA per SNP)
__global__ void __launch_bounds__(256) geno_combine_kernel(const double *__restrict__ G, int n, size_t ldg, const double *__restrict__ Ct, size_t ldc,
                                                           const int *__restrict__ dsel, const int *__restrict__ nplanes, const int *__restrict__ rowoff,
                                                           const double *__restrict__ mean, const int *__restrict__ nmiss,
                                                           const double *__restrict__ U, double *__restrict__ out, size_t ldo, int in_place) {
  constexpr int CAP = 1024;
  __shared__ int list[CAP];
  __shared__ int count;
  const int s = blockIdx.x;
  const int T = nplanes[s], d = dsel[s];
  double p10 = 1.0;
  for (int k = 0; k < d; ++k) p10 *= 10.0;
  const double *c0 = Ct + (size_t)rowoff[s] * ldc;
  double *o = out + (size_t)s * ldo;
  if (!in_place) {
    for (int i = threadIdx.x; i < n; i += 256) {
      double v = c0[(size_t)(T - 1) * ldc + i];
      for (int k = T - 2; k >= 0; --k) v = fma(v, 256.0, c0[(size_t)k * ldc + i]);
      o[i] = v / p10;
    }
  }
  if (nmiss[s] == 0) return;
  const double m = mean[s];
  const double *g = G + (size_t)s * ldg;
  for (int base = 0; base < n; base += CAP) {
    __syncthreads();
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    const int hi_p = min(n, base + CAP);
    for (int p = base + threadIdx.x; p < hi_p; p += 256)
      if (isnan(g[p])) list[atomicAdd(&count, 1)] = p;
    __syncthreads();
    const int cnt = count;
    if (cnt > 0) {
      if (threadIdx.x == 0)       // fixed summation order
        for (int a = 1; a < cnt; ++a) { int key = list[a], b2 = a - 1; while (b2 >= 0 && list[b2] > key) { list[b2 + 1] = list[b2]; --b2; } list[b2 + 1] = key; }
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += 256) {
        double acc = 0.0;
        for (int q = 0; q < cnt; ++q) acc += U[(size_t)list[q] * n + i];
        o[i] += m * acc;
      }
    }
  }
}

// G_dev: l x n SNP-major doubles (ld ldg), NaN = missing.  *taken = false (and nothing written) when a value is not a short decimal.
int i8_project_geno(gb200_ctx *c, const double *G_dev, size_t l, size_t ldg, double *UtXt_dev, bool *taken) {
  *taken = false;
  c->i8.no_xsum_consumer = true;      // these rows get no exact linear sums from the side GEMM: at least 4 planes from now on
  int rc = i8_prepare(c);
  if (rc) return rc;
  const I8Geom g = make_geom(c->n, c->i8.n_slices);
  if (!(c->cta_pair != 0 && (g.N % 32 == 0 || g.N == 240))) return GB200_OK;          // only the CTA-pair kernel carries this path
  const size_t CH = 2048;                                   // SNPs per GEMM launch (bounds the digit-row scratch: <= 3 x 2048 rows)
  GB_CUDA(c, c->i8.miss_mean.reserve(l * (sizeof(double) + 4 * sizeof(int)) + 64));
  double *mean = c->i8.miss_mean.as<double>();
  int *nmiss = reinterpret_cast<int *>(mean + l), *dsel = nmiss + l, *nplanes = dsel + l, *rowoff = nplanes + l, *info = rowoff + l;
  {
    ProfScope ps(c, "decode");
    geno_classify_kernel<<<(unsigned)l, 256, 0, c->stream>>>(G_dev, g.n, ldg, mean, nmiss, dsel, nplanes);
    GB_CUDA(c, cudaGetLastError());
  }
  // representable at all?  (one small read-back per batch)
  {
    GB_CUDA(c, cudaMemsetAsync(info, 0, 4 * sizeof(int), c->stream));
    geno_rowoff_kernel<<<1, 1024, 0, c->stream>>>(nplanes, dsel, (int)l, rowoff, info);
    int h[4];
    GB_CUDA(c, cudaMemcpyAsync(h, info, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (h[1] != 0) return GB200_OK;
  }
  GB_CUDA(c, c->i8.geno.reserve((3 * CH + 256) * (size_t)g.n_padk));
  GB_CUDA(c, cudaFuncSetAttribute(i8_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (!make_tmap((CUtensorMap *)c->i8.tmap_b, c->i8.slices.p, (uint64_t)g.n_groups * (uint64_t)g.N, (uint64_t)g.n_padk, (uint32_t)(g.N / 2)))
    return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the U planes (pair)");
  c->i8.tmap_b_half = true;
  for (size_t s0 = 0; s0 < l; s0 += CH) {
    const size_t lc = l - s0 < CH ? l - s0 : CH;
    int h[4];
    GB_CUDA(c, cudaMemsetAsync(info, 0, 4 * sizeof(int), c->stream));
    geno_rowoff_kernel<<<1, 1024, 0, c->stream>>>(nplanes + s0, dsel + s0, (int)lc, rowoff + s0, info);
    GB_CUDA(c, cudaMemcpyAsync(h, info, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(c, cudaStreamSynchronize(c->stream));
    const size_t R = (size_t)h[0], R_pad = (R + 255) / 256 * 256;
    const bool plain = (h[2] == 0);                         // every row is 0..255 integers: the GEMM writes U^T x itself
    {
      ProfScope ps(c, "decode");
      GB_CUDA(c, cudaMemsetAsync(c->i8.geno.as<uint8_t>() + R * (size_t)g.n_padk, 0, (R_pad - R) * (size_t)g.n_padk, c->stream));
      geno_to_planes_kernel<<<(unsigned)lc, 256, 0, c->stream>>>(G_dev + s0 * ldg, g.n, ldg, g.n_padk, dsel + s0, nplanes + s0, rowoff + s0,
                                                                 c->i8.geno.as<uint8_t>());
      GB_CUDA(c, cudaGetLastError());
    }
    double *Ct = UtXt_dev + s0 * c->n_c;
    if (!plain) { GB_CUDA(c, c->dTmp.reserve(R_pad * c->n_c * sizeof(double))); Ct = c->dTmp.as<double>(); }
    if (!make_tmap((CUtensorMap *)c->i8.tmap_a, c->i8.geno.p, R_pad, (uint64_t)g.n_padk, (uint32_t)I8_BM))
      return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the dosage digit rows");
    I8KernelParams p;
    p.T = g.T; p.NE = g.NE; p.N = g.N; p.n = g.n; p.l = (int)R;
    p.num_k_blocks = g.n_padk / I8_BK;
    p.m_tiles = (int)(R_pad / 256); p.n_groups = g.n_groups;
    p.lbo_units = 1; p.scale = c->i8.scale.as<double>();
    p.C = Ct; p.ldc = c->n_c; p.mode = 0; p.tiles = nullptr; p.num_tiles = 0;
    p.row_mean = nullptr; p.tile_holes = nullptr; p.panel = c->gemm_panel > 0 ? (int)c->gemm_panel : 9;
    const size_t stage_pair = (size_t)I8_BM * I8_BK + (size_t)(g.N / 2) * I8_BK;
    int ns = c->gemm_stages > 0 ? (int)c->gemm_stages : 6;
    while (ns > 2 && 1024 + (size_t)ns * stage_pair + 256 > 227 * 1024) --ns;
    p.stages = ns; p.wave_ctr = nullptr; p.hole_switch = nullptr; p.l2_hint = 0;
    const int tiles = p.m_tiles * p.n_groups;
    int pairs = c->num_sms / 2; if (pairs > tiles) pairs = tiles; if (pairs < 1) pairs = 1;
    if (c->gemm_wave_sync && 2 * pairs == c->num_sms / 2 * 2 && tiles > pairs) {
      GB_CUDA(c, c->i8.wave_ctr.reserve(sizeof(unsigned int)));
      GB_CUDA(c, cudaMemsetAsync(c->i8.wave_ctr.p, 0, sizeof(unsigned int), c->stream));
      p.wave_ctr = c->i8.wave_ctr.as<unsigned int>();
    }
    {
      ProfScope ps(c, "utx");
      i8_gemm_pair_kernel<<<2 * pairs, I8_THREADS, 1024 + (size_t)ns * stage_pair + 256, c->stream>>>(*(CUtensorMap *)c->i8.tmap_a, *(CUtensorMap *)c->i8.tmap_b, p);
      GB_CUDA(c, cudaGetLastError());
    }
    ProfScope ps2(c, "fix");
    geno_combine_kernel<<<(unsigned)lc, 256, 0, c->stream>>>(G_dev + s0 * ldg, g.n, ldg, Ct, c->n_c, dsel + s0, nplanes + s0, rowoff + s0, mean + s0,
                                                             nmiss + s0, c->dU.as<double>(), UtXt_dev + s0 * c->n_c, c->n_c, plain ? 1 : 0);
    GB_CUDA(c, cudaGetLastError());
  }
  *taken = true;
  return GB200_OK;
}

// ==========================================================================================
// Kinship on the tensor cores: K_raw = Z Z^T is exact in int32 for 0/1/2 genotypes, and for
// SNPs without missing calls the centred product of BimbamKin/PlinkKin (gemma_io.cpp:1511-1554)
// is  sum_s (z_s - m_s 1)(z_s - m_s 1)^T = Z Z^T - a 1^T - 1 a^T + (sum_s m_s^2) 1 1^T,
// a = sum_s m_s z_s, so one int8 GEMM plus rank-one FP64 terms replaces the FP64 dgemm.
// Batches that contain a missing genotype (or -gk 2) take the FP64 path.

// .bed (SNP-major 2-bit) -> individual-major int8 tile, plus per-SNP integer sum / missing count
__global__ void __launch_bounds__(256) bed_transpose_i8_kernel(const unsigned char *__restrict__ bed, size_t bps, int n,
                                                               int l, int8_t *__restrict__ Zt, size_t pitch, size_t col0,
                                                               int *__restrict__ sum, int *__restrict__ nmiss,
                                                               unsigned long long *__restrict__ qbits, size_t qpitch) {
  __shared__ int8_t tile[128][132];
  const int s0 = blockIdx.x * 128, i0 = blockIdx.y * 128;
  const int r = threadIdx.x >> 1, half = threadIdx.x & 1;      // SNP row r, individuals half*64 .. +63
  const int s = s0 + r;
  int my_sum = 0, my_miss = 0;
  for (int bq = 0; bq < 16; ++bq) {
    const int ib = i0 + half * 64 + bq * 4;                    // first individual of this byte
    unsigned byte = 0xFFu;                                     // 11 11 11 11 -> genotype 0, harmless padding
    bool inb = false;
    if (s < l && ib < n) { byte = bed[(size_t)s * bps + (size_t)(ib >> 2)]; inb = true; }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int8_t v = 0;
      if (inb && ib + q < n) {
        const unsigned b = (byte >> (2 * q)) & 3u;
        if (b == 0u) v = 2; else if (b == 2u) v = 1; else if (b == 1u) { my_miss++; v = -1; }   // -1 marks "missing" inside the tile only
        my_sum += (v > 0) ? v : 0;
      }
      tile[half * 64 + bq * 4 + q][r] = v;
    }
  }
  if (s < l) { if (my_sum) atomicAdd(&sum[s], my_sum); if (my_miss) atomicAdd(&nmiss[s], my_miss); }
  __syncthreads();
  // write: thread -> individual row (r), 64 SNP columns (half)
  const int ii = i0 + r;
  if (ii < n) {
    int8_t *dst = Zt + (size_t)ii * pitch + col0 + (size_t)s0 + (size_t)half * 64;
    unsigned long long mask = 0ull;                      // bit t: individual ii is missing at SNP s0 + half*64 + t
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int4 w;
      int8_t *wb = reinterpret_cast<int8_t *>(&w);
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        int8_t v = tile[r][half * 64 + q * 16 + t];
        if (v < 0) { mask |= 1ull << (q * 16 + t); v = 0; }
        wb[t] = v;
      }
      *reinterpret_cast<int4 *>(dst + q * 16) = w;
    }
    if (qbits) qbits[(size_t)ii * qpitch + ((col0 + (size_t)s0) >> 6) + (size_t)half] = mask;
  }
}

// mean_s, missing flag, beta += sum mean_s^2
__global__ void kin_snp_stats_kernel(const int *__restrict__ sum, const int *__restrict__ nmiss, int n, int l,
                                     double *__restrict__ mean, double *__restrict__ beta_flag /* [0]=beta [1]=flag */) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  double m2 = 0.0, fl = 0.0;
  if (s < l) {
    const int nm = nmiss[s];
    const double m = (double)sum[s] / (double)(n - nm);
    mean[s] = m; m2 = m * m;
    fl = (double)nm;                                    // total number of missing genotypes in the chunk
  }
  m2 = warp_allsum(m2); fl = warp_allsum(fl);
  if ((threadIdx.x & 31) == 0) { if (m2 != 0.0) atomicAdd(&beta_flag[2], m2); if (fl != 0.0) atomicAdd(&beta_flag[1], fl); }
}

// a[i] += sum_s mean_s * Z[s][i] over the staged columns [col0, col0+l)   (one warp per individual)
__global__ void __launch_bounds__(256) kin_a_kernel(const int8_t *__restrict__ Zt, size_t pitch, size_t col0, int l, int n,
                                                    const double *__restrict__ mean, double *__restrict__ a_pending) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const int8_t *row = Zt + (size_t)i * pitch + col0;
  double acc = 0.0;
  for (int s = lane; s < l; s += 32) acc = fma((double)row[s], mean[s], acc);
  acc = warp_allsum(acc);
  if (lane == 0) a_pending[i] += acc;
}

// Missing genotypes on the int8 path.  The reference imputes a missing entry to the SNP mean, i.e. its centred value is
// 0 (src/gemma_io.cpp:1688-1706): with z = 0 at missing entries and q the missing indicator,
//   sum_s (z_si - m_s (1-q_si)) (z_sj - m_s (1-q_sj))
//     = [Z Z^T]_ij - a_i - a_j + beta  +  X_ij + X_ji + G3_ij - b_i - b_j,
//   X_ij = sum_s m_s z_si q_sj,  G3_ij = sum_s m_s^2 q_si q_sj,  b_i = sum_s m_s^2 q_si.
// X and G3 are sparse in q: for each individual j this kernel walks the SNPs where j is missing (bit rows written by the
// transposer) and adds  m_s z_s[i] + (m_s^2 / 2) q_s[i]  into row j of Y (so that Y + Y^T = X + X^T + G3), reading the
// 2-bit SNP rows straight from the .bed chunk.  Work = (#missing entries) x n instead of a dense FP64 GEMM.
constexpr int KFIX_CAP = 512;      // SNPs staged per pass: index + 4-entry value table (2 m, m^2/2, m, 0) indexed by the .bed code
__global__ void __launch_bounds__(256) kin_miss_fix_kernel(const unsigned char *__restrict__ bed, size_t bps, int n, int l,
                                                           const unsigned long long *__restrict__ qbits, size_t qpitch, size_t word0,
                                                           const double *__restrict__ mean, double *__restrict__ Y,
                                                           double *__restrict__ b) {
  __shared__ int list_s[KFIX_CAP];
  __shared__ __align__(16) double lut[KFIX_CAP][4];
  __shared__ int warp_cnt[8];
  __shared__ double warp_b[8];
  const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i_base = (blockIdx.y * 256 + tid) * 16;
  const unsigned long long *qrow = qbits + (size_t)j * qpitch + word0;
  const int nwords = (l + 63) >> 6;
  const size_t byte0 = (size_t)(i_base >> 2);
  const bool active = i_base < n;
  const bool whole = active && byte0 + 4 <= bps;
  double acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.0;
  double bsum = 0.0;
  bool any = false;
  for (int w0 = 0; w0 < nwords; w0 += 256) {
    // every thread owns one 64-SNP word of the bit row; positions in the staged list follow the SNP order (deterministic sums)
    const int w = w0 + tid;
    unsigned long long bits = (w < nwords) ? __ldg(qrow + w) : 0ull;
    const int cnt = __popcll(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_cnt[wid] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int c = warp_cnt[q]; if (q < wid) base += c; total += c; }
    const int first = base + incl - cnt;                  // list position of this thread's first set bit
    __syncthreads();
    if (total == 0) continue;
    any = true;
    for (int seg0 = 0; seg0 < total; seg0 += KFIX_CAP) {
      // stage the SNPs [seg0, seg0 + KFIX_CAP) of this pass
      {
        unsigned long long bb = bits; int pos = first;
        while (bb) {
          const int k = __ffsll((long long)bb) - 1;
          bb &= bb - 1;
          if (pos >= seg0 && pos < seg0 + KFIX_CAP) {
            const int sidx = w * 64 + k;
            const double m = __ldg(mean + sidx);
            list_s[pos - seg0] = sidx;
            lut[pos - seg0][0] = m + m; lut[pos - seg0][1] = 0.5 * m * m; lut[pos - seg0][2] = m; lut[pos - seg0][3] = 0.0;
            if (blockIdx.y == 0) bsum += m * m;
          }
          ++pos;
        }
      }
      __syncthreads();
      const int cntseg = (total - seg0 < KFIX_CAP) ? (total - seg0) : KFIX_CAP;
      if (active) {
        int e = 0;
        for (; e + 4 <= cntseg; e += 4) {
          unsigned g[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const unsigned char *row = bed + (size_t)list_s[e + u] * bps + byte0;
            if (whole) g[u] = (unsigned)row[0] | ((unsigned)row[1] << 8) | ((unsigned)row[2] << 16) | ((unsigned)row[3] << 24);
            else {
              g[u] = 0xFFFFFFFFu;
              for (int q = 0; q < 4; ++q) if (byte0 + (size_t)q < bps) g[u] = (g[u] & ~(0xFFu << (8 * q))) | ((unsigned)row[q] << (8 * q));
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double *L = lut[e + u];
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[t] += L[(g[u] >> (2 * t)) & 3u];
          }
        }
        for (; e < cntseg; ++e) {
          const unsigned char *row = bed + (size_t)list_s[e] * bps + byte0;
          unsigned g = 0xFFFFFFFFu;
          for (int q = 0; q < 4; ++q) if (byte0 + (size_t)q < bps) g = (g & ~(0xFFu << (8 * q))) | ((unsigned)row[q] << (8 * q));
          const double *L = lut[e];
#pragma unroll
          for (int t = 0; t < 16; ++t) acc[t] += L[(g >> (2 * t)) & 3u];
        }
      }
      __syncthreads();
    }
  }
  if (!any) return;                                        // uniform over the CTA
  if (active) {
    double *yrow = Y + (size_t)j * (size_t)n;
#pragma unroll
    for (int t = 0; t < 16; ++t)
      if (i_base + t < n) yrow[i_base + t] += acc[t];
  }
  if (blockIdx.y == 0) {                                   // b[j] += sum of m_s^2 over the SNPs where j is missing (fixed reduction order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    if (lane == 0) warp_b[wid] = bsum;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int q = 0; q < 8; ++q) t += warp_b[q]; b[j] += t; }
  }
}

__global__ void kin_commit_kernel(double *a, const double *a_pending, double *beta_flag, int n) {
  // fold the pending (checked, missing-free) chunk into the running totals
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += a_pending[i];
  if (i == 0) { beta_flag[0] += beta_flag[2]; }
}

__global__ void kin_finish_kernel(double *K, size_t n, size_t ld, const double *__restrict__ a, const double *__restrict__ beta_flag,
                                  const double *__restrict__ Y, const double *__restrict__ b, double inv_ns) {
  // rows on grid.x (limit 2^31 - 1), column blocks on grid.y: n / 256 stays far below the 65 535 limit of grid.y
  const size_t j = (size_t)blockIdx.y * blockDim.x + threadIdx.x, i = blockIdx.x;
  if (j > i || i >= n) return;
  double v = K[i * ld + j] - a[i] - a[j] + beta_flag[0];
  if (Y) v += (Y[i * n + j] + Y[j * n + i]) - b[i] - b[j];      // missing-genotype terms (kin_miss_fix_kernel)
  K[i * ld + j] = v * inv_ns;
}

static bool make_tmap_rows(CUtensorMap *tm, const void *base, uint64_t rows, uint64_t inner_bytes, uint64_t pitch_bytes,
                           uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {inner_bytes, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)I8_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool kin_i8_eligible(gb200_ctx *c) {
  return c->kin_path != 1 && c->kin_mode == 1 && c->kin_n >= 1024 && get_encode() != nullptr;
}

static int kin_i8_setup(gb200_ctx *c) {
  I8State &S = c->i8;
  const size_t n = c->kin_n;
  if (S.kin_n == n && S.kin_cap) return GB200_OK;
  size_t cap = ((size_t)1 << 31) / n;                 // <= 2 GiB of staged int8 genotypes
  if (cap > 131072) cap = 131072;
  cap = cap / 128 * 128;
  if (cap < 1024) cap = 1024;
  GB_CUDA(c, S.kin_zt.reserve(n * cap));
  GB_CUDA(c, S.kin_stats.reserve(cap * (2 * sizeof(int) + sizeof(double))));
  GB_CUDA(c, S.kin_a.reserve((2 * n + 8) * sizeof(double)));
  GB_CUDA(c, S.kin_qbits.reserve(n * (cap / 64) * sizeof(unsigned long long)));
  // lower-triangle tiles: 128-row x 256-column tiles that intersect j <= i
  std::vector<int2> tiles;
  const int m_tiles = (int)((n + 127) / 128);
  for (int m = 0; m < m_tiles; ++m)
    for (int nb = 0; nb * 256 <= m * 128 + 127; ++nb) tiles.push_back(make_int2(m, nb));
  S.kin_num_tiles = (int)tiles.size();
  const int m2_tiles = (int)((n + 255) / 256);            // CTA-pair tiles: 256 x 256, nb <= m
  for (int m = 0; m < m2_tiles; ++m)
    for (int nb = 0; nb <= m; ++nb) tiles.push_back(make_int2(m, nb));
  S.kin_num_tiles_pair = (int)tiles.size() - S.kin_num_tiles;
  GB_CUDA(c, S.kin_tiles.reserve(tiles.size() * sizeof(int2)));
  GB_CUDA(c, cudaMemcpyAsync(S.kin_tiles.p, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (!S.tmap_ka) S.tmap_ka = aligned_alloc(64, sizeof(CUtensorMap));
  if (!S.tmap_kb) S.tmap_kb = aligned_alloc(64, sizeof(CUtensorMap));
  S.kin_cap = cap; S.kin_n = n; S.kin_fill = 0;
  return GB200_OK;
}

int kin_i8_begin(gb200_ctx *c) {
  I8State &S = c->i8;
  S.kin_used = false; S.kin_fill = 0; S.kin_y_used = false;
  if (!kin_i8_eligible(c)) return GB200_OK;
  int rc = kin_i8_setup(c);
  if (rc) return rc;
  GB_CUDA(c, cudaMemsetAsync(S.kin_a.p, 0, (2 * c->kin_n + 8) * sizeof(double), c->stream));
  return GB200_OK;
}

int kin_i8_flush(gb200_ctx *c) {
  I8State &S = c->i8;
  if (S.kin_fill == 0) return GB200_OK;
  const size_t n = c->kin_n, kbytes = (S.kin_fill + 127) / 128 * 128;
  if (!make_tmap_rows((CUtensorMap *)S.tmap_ka, S.kin_zt.p, n, kbytes, S.kin_cap, 128) ||
      !make_tmap_rows((CUtensorMap *)S.tmap_kb, S.kin_zt.p, n, kbytes, S.kin_cap, 256))
    return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the kinship genotype matrix");
  I8KernelParams p;
  p.T = 1; p.NE = 256; p.N = 256; p.n = (int)n; p.l = (int)n;
  p.num_k_blocks = (int)(kbytes / I8_BK);
  p.m_tiles = (int)((n + 127) / 128); p.n_groups = (int)((n + 255) / 256);
  p.lbo_units = 1; p.scale = nullptr;
  p.C = c->dK.as<double>(); p.ldc = n; p.mode = 1;
  p.tiles = S.kin_tiles.as<int2>(); p.num_tiles = S.kin_num_tiles;
  p.row_mean = nullptr; p.tile_holes = nullptr; p.panel = I8_PANEL; p.stages = I8_STAGES; p.wave_ctr = nullptr; p.hole_switch = nullptr; p.l2_hint = 0;
  const size_t smem = 1024 + (size_t)I8_STAGES * (I8_BM * I8_BK + (size_t)256 * I8_BK) + 256;
  GB_CUDA(c, cudaFuncSetAttribute(i8_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (c->kin_cta_pair != 0) {
    // pair tiles: 256 x 256, lower-triangle tiles nb <= m; B box = 128 rows (half of N)
    if (!make_tmap_rows((CUtensorMap *)S.tmap_kb, S.kin_zt.p, n, kbytes, S.kin_cap, 128))
      return set_err(c, GB200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the kinship genotype matrix (pair)");
    p.tiles = S.kin_tiles.as<int2>() + S.kin_num_tiles; p.num_tiles = S.kin_num_tiles_pair;
    GB_CUDA(c, cudaFuncSetAttribute(i8_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int pairs = c->num_sms / 2; if (pairs > p.num_tiles) pairs = p.num_tiles; if (pairs < 1) pairs = 1;
    ProfScope ps(c, "kin");
    i8_gemm_pair_kernel<<<2 * pairs, I8_THREADS, smem, c->stream>>>(*(CUtensorMap *)S.tmap_ka, *(CUtensorMap *)S.tmap_kb, p);
    GB_CUDA(c, cudaGetLastError());
  } else {
    const int grid = p.num_tiles < c->num_sms ? p.num_tiles : c->num_sms;
    ProfScope ps(c, "kin");
    i8_gemm_kernel<<<grid, I8_THREADS, smem, c->stream>>>(*(CUtensorMap *)S.tmap_ka, *(CUtensorMap *)S.tmap_kb, p);
    GB_CUDA(c, cudaGetLastError());
  }
  S.kin_fill = 0;
  return GB200_OK;
}

// Try to stage `l` SNPs (device .bed rows) for the int8 kinship GEMM.  *taken = false when the chunk holds a
// missing genotype (the caller then runs the FP64 path on it); nothing is staged in that case.
int kin_i8_add_chunk(gb200_ctx *c, const unsigned char *bed_dev, size_t l, size_t bytes_per_snp, bool *taken) {
  I8State &S = c->i8;
  *taken = false;
  const size_t n = c->kin_n;
  if (l > S.kin_cap) return set_err(c, GB200_ERR_ARG, "kinship chunk larger than the staging capacity");
  if (S.kin_fill + l > S.kin_cap) { int rc = kin_i8_flush(c); if (rc) return rc; }
  int *sum = S.kin_stats.as<int>();
  int *nmiss = sum + S.kin_cap;
  double *mean = reinterpret_cast<double *>(nmiss + S.kin_cap);
  double *a = S.kin_a.as<double>(), *a_pending = a + n, *beta_flag = a + 2 * n;   // [0] beta, [1] flag, [2] pending beta
  const size_t lpad = (l + 127) / 128 * 128;
  {
    ProfScope ps(c, "decode", 4);
    GB_CUDA(c, cudaMemsetAsync(sum, 0, 2 * S.kin_cap * sizeof(int), c->stream));
    GB_CUDA(c, cudaMemsetAsync(a_pending, 0, n * sizeof(double), c->stream));
    GB_CUDA(c, cudaMemsetAsync(beta_flag + 1, 0, 2 * sizeof(double), c->stream));
    dim3 grid((unsigned)(lpad / 128), (unsigned)((n + 127) / 128));
    bed_transpose_i8_kernel<<<grid, 256, 0, c->stream>>>(bed_dev, bytes_per_snp, (int)n, (int)l, S.kin_zt.as<int8_t>(), S.kin_cap,
                                                         S.kin_fill, sum, nmiss, S.kin_qbits.as<unsigned long long>(),
                                                         S.kin_cap / 64);
    kin_snp_stats_kernel<<<(unsigned)((l + 255) / 256), 256, 0, c->stream>>>(sum, nmiss, (int)n, (int)l, mean, beta_flag);
    kin_a_kernel<<<(unsigned)((n + 7) / 8), 256, 0, c->stream>>>(S.kin_zt.as<int8_t>(), S.kin_cap, S.kin_fill, (int)l, (int)n, mean, a_pending);
    GB_CUDA(c, cudaGetLastError());
  }
  double flag = 0.0;
  GB_CUDA(c, cudaMemcpyAsync(&flag, beta_flag + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  GB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (flag > c->kin_miss_max * (double)n * (double)l) {
    // too many missing genotypes for the sparse correction to pay: un-stage the chunk (zero its columns), FP64 path
    GB_CUDA(c, cudaMemset2DAsync(S.kin_zt.as<int8_t>() + S.kin_fill, S.kin_cap, 0, lpad, n, c->stream));
    return GB200_OK;
  }
  if (flag != 0.0) {
    // sparse missing-genotype terms of this chunk (the integer GEMM sees z = 0 at those entries)
    if (!S.kin_y_used) {
      GB_CUDA(c, S.kin_y.reserve((n * n + n) * sizeof(double)));
      GB_CUDA(c, cudaMemsetAsync(S.kin_y.p, 0, (n * n + n) * sizeof(double), c->stream));
      S.kin_y_used = true;
    }
    ProfScope ps(c, "fix", 1);
    dim3 grid((unsigned)n, (unsigned)((n + 4095) / 4096));
    kin_miss_fix_kernel<<<grid, 256, 0, c->stream>>>(bed_dev, bytes_per_snp, (int)n, (int)l, S.kin_qbits.as<unsigned long long>(),
                                                     S.kin_cap / 64, S.kin_fill / 64, mean, S.kin_y.as<double>(),
                                                     S.kin_y.as<double>() + n * n);
    GB_CUDA(c, cudaGetLastError());
  }
  kin_commit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a, a_pending, beta_flag, (int)n);
  GB_CUDA(c, cudaGetLastError());
  S.kin_fill += lpad;                                   // keep the fill pointer 128-aligned (padding columns are zero)
  S.kin_used = true;
  c->kin_ns += l;
  *taken = true;
  return GB200_OK;
}

// flush + apply the rank-one centring terms and the 1/ns scaling (replaces launch_scale at finish)
int kin_i8_finish(gb200_ctx *c, double inv_ns) {
  I8State &S = c->i8;
  int rc = kin_i8_flush(c);
  if (rc) return rc;
  const size_t n = c->kin_n;
  double *a = S.kin_a.as<double>();
  dim3 grid((unsigned)n, (unsigned)((n + 255) / 256));
  const double *Y = S.kin_y_used ? S.kin_y.as<double>() : nullptr;
  kin_finish_kernel<<<grid, 256, 0, c->stream>>>(c->dK.as<double>(), n, n, a, a + 2 * n, Y, Y ? Y + n * n : nullptr, inv_ns);
  GB_CUDA(c, cudaGetLastError());
  return GB200_OK;
}

}  // namespace gb

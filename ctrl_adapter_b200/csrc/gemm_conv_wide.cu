// The multi-tap tcgen05 GEMM of gemm_conv_kernel.cuh with 256 x 320 CTA-pair tiles for the N = 320 / 640 / 960 / 1920
// convolutions and linears (default dispatch since round 2: kernel checks green on hardware, profiles/r2_experiments.md;
// CA_GEMM_BN320=0 falls back to the 160-wide tiles).
//
// Why: the operand feed of the GEMM is bound by L2 -> SM bandwidth (~6.3 KB / clk chip-wide, profiles/README.md).  With
// BN = 160 tiles a k-block moves 16 KB (A) + 10 KB (B) per CTA for 128 x 160 x 64 MACs and the tensor pipe tops out
// at ~52 % (measured 0.87-1.1 PFLOP/s on the N = 320 / 640 convs); a 320-wide tile reuses the same A block for twice
// the columns: 16 + 20 KB for 128 x 320 x 64, i.e. 31 % less traffic per MAC (-> ~75 %).
//
// A 128-lane x 320-column fp32 accumulator cannot be double buffered in the 512 TMEM columns.  Instead the tile is
// accumulated as TWO 160-column halves in THREE rotating TMEM regions (3 x 160 = 480 columns): tile j uses regions
// (2j) % 3 and (2j+1) % 3, each half is drained by its own group of four epilogue warps, and the MMA issuer of the next
// tile only needs the half that was drained first -- so half of every epilogue overlaps the next main loop.  Region
// uses are strictly round robin (use u = 2j + g -> region u % 3, previous uses of that region = u / 3), which is all
// the phase bookkeeping there is.
//
// Everything else (TMA 5-D implicit-GEMM boxes, CTA pairs with cta_group::2, M = 256, the mbarrier stage ring, the lean
// bias / bias+residual epilogues with packed bf16 math, swizzled staging, coalesced stores) is the validated kernel's.
// Each CTA stages its 160 weight rows per k-block as two 80-row chunks (rows [r*80, r*80+80) of either half), so that
// MMA #h (N = 160: 80 rows from each CTA) produces the contiguous output columns [h*160, h*160+160).
#include "gemm_conv_kernel.cuh"

namespace ca {

struct WideCfg {
  static constexpr int kBN = 320, kHalf = 160, kRegions = 3;
  static constexpr uint32_t kABytes = kBM * kBK * 2;              // 16 KB
  static constexpr uint32_t kBChunkRows = 80;                     // rows of one half staged by this CTA
  static constexpr uint32_t kBChunkBytes = kBChunkRows * kBK * 2;  // 10 KB (a multiple of the 1 KB swizzle atom)
  static constexpr uint32_t kBBytes = 2 * kBChunkBytes;           // 20 KB
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;      // 36 KB
  static constexpr uint32_t kBudget = 232448 - 1024 - 256 - kStagingBytes;
  static constexpr int kStages = kBudget / kStageBytes;           // 4
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};
static_assert(WideCfg::kStages >= 3, "stage ring too short");

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_conv_wide_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                      const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
  static_assert(EPI == EPI_PLAIN || EPI == EPI_RES, "wide tiles: bias / bias+residual epilogues only");
  using Cfg = WideCfg;
  constexpr int NCTA = 2;
  CA_PDL_TRIGGER();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* smem_stage = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + kStagingBytes);
  uint64_t* full_bar = bars;                         // [kStages]  (leader's instance is the live one)
  uint64_t* empty_bar = bars + Cfg::kStages;         // [kStages]
  uint64_t* acc_full = bars + 2 * Cfg::kStages;      // [3] one per TMEM region
  uint64_t* acc_empty = acc_full + Cfg::kRegions;    // [3] (leader's instance)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + Cfg::kRegions);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  const int group = blockIdx.x / NCTA;
  const int ngroups = gridDim.x / NCTA;

  const int tiles_m = p.ntile[0] * p.ntile[1] * p.ntile[2] * p.ntile[3];
  const int tiles_mg = (tiles_m + NCTA - 1) / NCTA;
  const int total_tiles = tiles_mg * p.n_tiles_n;
  const int chunks0 = (p.src_c[0] + kBK - 1) / kBK;
  const int chunks1 = (p.nsrc > 1) ? (p.src_c[1] + kBK - 1) / kBK : 0;
  const int kiters = p.ntaps * (chunks0 + chunks1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a0);
    if (p.nsrc > 1) tma_prefetch_desc(&tmap_a1);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < Cfg::kRegions; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], NCTA * 4);  // a region is drained by ONE group of four epilogue warps in each CTA
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  CA_PDL_WAIT();

  auto tile_coords = [&](int tile, int& tn, int (&org)[4]) {
    tn = tile % p.n_tiles_n;
    int tm = (tile / p.n_tiles_n) * NCTA + rank;
    org[0] = (tm % p.ntile[0]) * p.box[0]; tm /= p.ntile[0];
    org[1] = (tm % p.ntile[1]) * p.box[1]; tm /= p.ntile[1];
    org[2] = (tm % p.ntile[2]) * p.box[2]; tm /= p.ntile[2];
    org[3] = tm * p.box[3];
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = group; tile < total_tiles; tile += ngroups) {
        int tn, org[4];
        tile_coords(tile, tn, org);
        const int n_lo = tn * Cfg::kBN + rank * static_cast<int>(Cfg::kBChunkRows);  // this CTA's rows of half 0
        const int n_hi = n_lo + Cfg::kHalf;                                          // ... and of half 1
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int c1 = org[0] + p.tap_off[tap][0], c2 = org[1] + p.tap_off[tap][1];
          const int c3 = org[2] + p.tap_off[tap][2], c4 = org[3] + p.tap_off[tap][3];
          const int kbase = tap * p.k_per_tap;
          for (int ch = 0; ch < chunks0 + chunks1; ++ch) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            const bool second = ch >= chunks0;
            const int cc = (second ? (ch - chunks0) : ch) * kBK;
            const int ca = cc + (second ? p.src_c0_off[1] : p.src_c0_off[0]) + p.tap_c_off[tap];
            const int kw = kbase + (second ? p.src_c[0] : 0) + cc;
            const CUtensorMap* ta = second ? &tmap_a1 : &tmap_a0;
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            tma_load_5d_pair(smem_a + stage * Cfg::kABytes, ta, &full_bar[stage], ca, c1, c2, c3, c4);
            uint8_t* bdst = smem_b + stage * Cfg::kBBytes;
            tma_load_2d_pair(bdst, &tmap_w, &full_bar[stage], kw, n_lo);
            tma_load_2d_pair(bdst + Cfg::kBChunkBytes, &tmap_w, &full_bar[stage], kw, n_hi);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM * NCTA, Cfg::kHalf, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t u = 0;  // region-use counter: tile j uses u = 2j and 2j + 1
      for (int tile = group; tile < total_tiles; tile += ngroups, u += 2) {
        const uint32_t r0 = u % 3, r1 = (u + 1) % 3;
        mbar_wait(&acc_empty[r0], ((u / 3) & 1) ^ 1);        // drained first by the previous user's epilogue group
        mbar_wait(&acc_empty[r1], (((u + 1) / 3) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + r0 * Cfg::kHalf, d1 = tmem_base + r1 * Cfg::kHalf;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db0 = umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            const uint64_t db1 = umma_smem_desc_sw128(b_addr + Cfg::kBChunkBytes + k * 32, 16, 1024);
            const uint32_t accumulate = (it | k) != 0 ? 1u : 0u;
            umma_bf16_ss_pair(d0, da, db0, idesc, accumulate);
            umma_bf16_ss_pair(d1, da, db1, idesc, accumulate);
          }
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&acc_full[r0]);
        umma_commit_pair(&acc_full[r1]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: warp group g drains the g-th 160-column half of every tile =====================
    constexpr bool kRes = EPI == EPI_RES;
    const int ew = warp - 4;
    const int q = warp & 3;
    const int g = ew >> 2;
    const int r = q * 32 + lane;
    uint8_t* stg_out = smem_stage + ew * kWarpStageBytes;
    uint8_t* stg_res = smem_stage + (kEpiWarps + ew) * kWarpStageBytes;
    const int seg = lane & 7;
    const int rsub = lane >> 3;
    const bool has_bias = p.bias != nullptr;
    uint32_t u = static_cast<uint32_t>(g);  // this group's region use: u = 2j + g
    for (int tile = group; tile < total_tiles; tile += ngroups, u += 2) {
      int tn, org[4];
      tile_coords(tile, tn, org);
      int rr = r;
      const int i1 = rr % p.box[0]; rr /= p.box[0];
      const int i2 = rr % p.box[1]; rr /= p.box[1];
      const int i3 = rr % p.box[2]; rr /= p.box[2];
      const int i4 = rr;
      const int o[4] = {org[0] + i1, org[1] + i2, org[2] + i3, org[3] + i4};
      const bool row_ok = o[0] < p.odim[0] && o[1] < p.odim[1] && o[2] < p.odim[2] && o[3] < p.odim[3];
      const long long out_off = o[0] * p.ostride[0] + o[1] * p.ostride[1] + o[2] * p.ostride[2] + o[3] * p.ostride[3];
      const long long res_off = o[0] * p.rstride[0] + o[1] * p.rstride[1] + o[2] * p.rstride[2] + o[3] * p.rstride[3];
      const unsigned ok_mask = __ballot_sync(0xffffffffu, row_ok);
      const int col0 = tn * Cfg::kBN + g * Cfg::kHalf;  // first output column of this group's half
      const uint32_t region = u % 3;

      auto unit_valid = [&](int u0) { return min(64, min(Cfg::kHalf - u0, p.n_out - (col0 + u0))); };  // multiple of 8
      auto prefetch = [&](int u0) {
        const int ucol = col0 + u0;
        const int uvalid = unit_valid(u0);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rowi = it * 4 + rsub;
          const long long roff = __shfl_sync(0xffffffffu, res_off, rowi);
          uint8_t* dst = stg_res + rowi * kStageRowBytes + ((seg ^ (rowi & 7)) << 4);
          if (((ok_mask >> rowi) & 1u) && seg * 8 < uvalid) cp_async_16(dst, p.residual + roff + ucol + seg * 8);
          else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
      };
      if (kRes && col0 < p.n_out) prefetch(0);

      mbar_wait(&acc_full[region], (u / 3) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + region * Cfg::kHalf + (static_cast<uint32_t>(q * 32) << 16);

      for (int u0 = 0; u0 < Cfg::kHalf; u0 += 64) {
        const int ucol = col0 + u0;
        if (ucol >= p.n_out) break;
        const int uvalid = unit_valid(u0);
        if (kRes) {
          cp_async_wait_all();
          __syncwarp();
        }
#pragma unroll
        for (int hsel = 0; hsel < 2; ++hsel) {
          if (hsel * 32 >= uvalid) break;
          const int c0 = u0 + hsel * 32;
          const int nvalid = uvalid - hsel * 32;
          uint32_t va[32];
          tmem_ld_32x32(t_row + c0, va);  // the last unit is 32 wide: hsel 1 is skipped there (uvalid <= 32)
          const float4* bias_a = reinterpret_cast<const float4*>(p.bias + col0 + c0);
          tmem_ld_wait();
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (j8 * 8 >= nvalid) break;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(va[j8 * 8 + e]);
            if (has_bias) {
              const float4 b0 = __ldg(bias_a + j8 * 2), b1 = __ldg(bias_a + j8 * 2 + 1);
              x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
              x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
            }
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack_bf16x2(x[2 * e], x[2 * e + 1]);
            if constexpr (kRes) {
              const uint4 rr4 = *reinterpret_cast<const uint4*>(stg_res + lane * kStageRowBytes +
                                                                (((hsel * 4 + j8) ^ (lane & 7)) << 4));
              w[0] = add_bf16x2(w[0], rr4.x); w[1] = add_bf16x2(w[1], rr4.y);
              w[2] = add_bf16x2(w[2], rr4.z); w[3] = add_bf16x2(w[3], rr4.w);
            }
            *reinterpret_cast<uint4*>(stg_out + lane * kStageRowBytes + (((hsel * 4 + j8) ^ (lane & 7)) << 4)) =
                make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        __syncwarp();
        if (kRes && u0 + 64 < Cfg::kHalf && col0 + u0 + 64 < p.n_out) prefetch(u0 + 64);
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rowi = it * 4 + rsub;
          const long long ooff = __shfl_sync(0xffffffffu, out_off, rowi);
          if (((ok_mask >> rowi) & 1u) && seg * 8 < uvalid) {
            const uint4 uu = *reinterpret_cast<const uint4*>(stg_out + rowi * kStageRowBytes + ((seg ^ (rowi & 7)) << 4));
            *reinterpret_cast<uint4*>(outp + ooff + ucol + seg * 8) = uu;
          }
        }
        __syncwarp();
      }
      // every TMEM read of this region is complete -> hand it back to the (leader's) MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&acc_empty[region]);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

template <int EPI>
static cudaError_t launch_wide(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const GemmParams& p,
                               int grid, cudaStream_t stream) {
  using Cfg = WideCfg;
  static bool attr_set = false;
  auto kern = gemm_conv_wide_kernel<EPI>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Cfg::kSmemBytes));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#ifndef CA_NO_PDL
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = 2;
#endif
  return cudaLaunchKernelEx(&cfg, kern, a0, a1, w, p);
}

cudaError_t launch_gemm_wide(int epi, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w,
                             const GemmParams& p, int grid, cudaStream_t stream) {
  switch (epi) {
    case EPI_PLAIN: return launch_wide<EPI_PLAIN>(a0, a1, w, p, grid, stream);
    case EPI_RES: return launch_wide<EPI_RES>(a0, a1, w, p, grid, stream);
    default: return cudaErrorInvalidValue;  // the caller only selects 320-wide tiles for these two epilogues
  }
}

}  // namespace ca

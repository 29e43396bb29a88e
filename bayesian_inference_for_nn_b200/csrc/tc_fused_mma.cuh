// tc_fused_mma.cuh — the fused layer-1 GEMM (int8 slices, tcgen05) + layer-2 kernel whose epilogue runs the two SMALL
// layer-2 products on the tensor cores as well (mma.sync m16n8k16, bf16 hi/lo split, fp32 accumulation).
//
// Why: ncu on tc_g1_layer2_fused<.., I8> (profiles/r2_ncu_full_c3_i8_kernels.txt) shows the kernel bound by its epilogue:
// 55 warp instructions per (row, hidden unit), 12 of them the FFMA2 of  z2 += a1 W2  and  dZ1 = dZ2 W2^T  (two FMA-pipe
// cycles each), fed by 1 536 LDS.128 of W2 fragments per tile (4 wavefronts each) and followed by a shuffle
// reduce-scatter; the tcgen05 pipe is 34 % busy.  Both products are GEMMs whose register layouts are already there:
//   * tcgen05.ld.16x256b hands lane (g, t) the columns {2t, 2t+1} of rows g, g+8 per 8-column block — that IS the
//     accumulator layout of mma.m16n8, which is the A-fragment layout of mma.m16n8k16: two adjacent 8-unit blocks of a1
//     are one A fragment (k = 16 hidden units), no shuffles;
//   * the product over the quad's units happens inside the MMA, so the partial logits come out already reduced over
//     the warp's 128 units, spread over the quad by class (lane t: classes 2t, 2t+1, 8+2t, 9+2t of its four rows);
//   * dZ1 = dZ2 W2^T is m16 (rows) x n8 (units) x k16 (classes, zero padded): its output fragment is again the
//     (rows g, g+8; units 2t, 2t+1) ownership of the TMEM layout, so mask, slicing and the transposed stores are unchanged.
// W2 lives in shared memory ONCE per chain as bf16 hi / lo [256 units][16 classes] (32-byte rows): ldmatrix.trans gives the
// B fragments of the logits product (k = units), plain ldmatrix those of the delta product (k = classes): 32 ldmatrix.x4
// per warp and tile instead of 192 LDS.128.  fp32-grade products as everywhere: hi*hi + lo*hi + hi*lo.
// Everything else (TMA producer, tcgen05 issuer, int8 reconstruction, loss, stores) is tc_g1_layer2_fused<CP, false, I8>.
#pragma once
#include "tc_i8.cuh"

namespace pyb {

// SIXTEEN epilogue warps (4 TMEM lane quadrants x 4 quarters of the hidden units): with both accumulators in TMEM the
// accumulator cannot be double-buffered, so phase A of the epilogue is serial with the next tile's MMAs (cycle = T_A +
// T_mma); eight warps left every unit half idle (issue 38 %, tensor 53 %, L1 65 %, L2 58 %: latency-bound), sixteen halve T_A.
constexpr int TFM_THREADS = 640;                 // 16 epilogue warps + one control warpgroup (TMA, tcgen05 issuer, TMEM allocator, idle)
struct TfmCfg {
  static constexpr int STAGES = 4;
  static constexpr int SMEM = STAGES * TP_STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers*/ + 2048 /*bias x2*/ + 128 /*b2 x2*/ +
                              4096 /*cw, zq x2*/ + 16384 /*W2 staging fp32 [256][16]*/ + 2 * 8192 /*W2 bf16 hi, lo [256][16]*/ +
                              32768 /*partial logits [4][16][128]*/ + 8192 /*dZ2 fragments [4 quadrants][32 rows][64 B]*/;
};

// 16 lanes x 2 column blocks of 8: rows {g, g+8} of the 16 TMEM lanes starting at the address's lane, raw 32-bit words
__device__ __forceinline__ void tc_ld_16x256b_x2_raw(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}

// ZI8: dZ1^T leaves as two int8 slices (softmax-CE, a-priori scale), else as bf16 hi/lo for the bf16x3 dW1 GEMM
template <int CP, bool ZI8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TFM_THREADS, 1)
tc_fused_i8_mma(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const TcGemmParams p, const Layer2Params l2) {
  constexpr int STAGES = TfmCfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint64_t* bars = (uint64_t*)(smem + STAGES * TP_STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]   (leader CTA)
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]   (one per CTA)
  uint64_t* tmem_full = bars + 2 * STAGES;       // (one per CTA)
  uint64_t* tmem_empty = bars + 2 * STAGES + 1;  // (leader CTA)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * STAGES + 2);
  float* bias_s = (float*)(smem + STAGES * TP_STAGE_BYTES + 1024);          // [2][256]
  float* b2_s = bias_s + 512;                                               // [2][16]
  float* cw_s = b2_s + 32;                                                  // [2][256]
  float* zq_s = cw_s + 512;                                                 // [2][256]
  float* w2st_s = zq_s + 512;                                               // [256][16] fp32: cp.async target of the next item
  uint8_t* w2h_s = reinterpret_cast<uint8_t*>(w2st_s + 256 * 16);           // [256][16] bf16 hi
  uint8_t* w2l_s = w2h_s + 8192;                                            // [256][16] bf16 lo
  float* zx_s = reinterpret_cast<float*>(w2l_s + 8192);                     // [4 quarters][16][128]
  uint4* dz_s = reinterpret_cast<uint4*>(zx_s + 4 * 16 * 128);              // [4 quadrants][32 rows][4 slots]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int nk = (p.K + 63) / 64;
  const int H = p.H;
  const int half_rows = H >> 1;
  const uint32_t cta_bytes = 2 * TC_A_TILE_BYTES + 2 * (uint32_t)half_rows * 64;

  if (warp == 16 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
  }
  if (warp == 17 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 32);                   // 16 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 18) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= 16) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
   if (warp == 16) {
    // ===== TMA producer (both CTAs): own 128 rows of the X slices, own half of the chain's W1^T slices =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        int b, mp, split;
        tc_decode(p, item, b, mp, split);
        const int arow = p.a_row0 + (mp * 2 + (int)rank) * 128;
        const int brow = b * H + (int)rank * half_rows;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait_sleep(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stage_base + stage * TP_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * cta_bytes);
          const int k0 = kc * 64;
          tma_load_2d_pair(st, &tmA_hi, &full_bar[stage], k0, arow);
          tma_load_2d_pair(st + TC_A_TILE_BYTES, &tmA_lo, &full_bar[stage], k0, arow);
          tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES, &tmB_hi, &full_bar[stage], k0, brow);
          tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES + 8192, &tmB_lo, &full_bar[stage], k0, brow);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
   } else if (warp == 17) {
    // ===== tcgen05 issuer (leader CTA only): hi*hi -> columns [0, 256), hi*lo + lo*hi -> [256, 512) =====
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(H >> 3) << 17) | ((256u >> 4) << 24);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      const int k_tail = p.K - (nk - 1) * 64;
      long long t_wait = 0, t_loop = 0, t_fill = 0;
      for (int item = cluster_id; item < p.total_items; item += n_clusters, ++it) {
        const long long c0 = clock64();
        mbar_wait_sleep(tmem_empty, (uint32_t)(it & 1) ^ 1);
        tc_fence_after();
        const long long c1 = clock64();
        t_wait += c1 - c0;
        for (int kc = 0; kc < nk; ++kc) {
          const long long f0 = clock64();
          mbar_wait_sleep(&full_bar[stage], phase);
          tc_fence_after();
          t_fill += clock64() - f0;
          const uint32_t st = smem_u32(stage_base + stage * TP_STAGE_BYTES);
          const int nks = (kc == nk - 1 && k_tail <= 32) ? 1 : 2;
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t koff = ks * 32;
            const uint64_t ah = make_smem_desc_sw64(st + koff);
            const uint64_t al = make_smem_desc_sw64(st + TC_A_TILE_BYTES + koff);
            const uint64_t bh = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + koff);
            const uint64_t bl = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + 8192 + koff);
            const uint32_t accum = (kc != 0) || (ks != 0);
            tc_mma_i8_pair(tmem_base, ah, bh, idesc, accum);
            tc_mma_i8_pair(tmem_base + 256, ah, bl, idesc, accum);
            tc_mma_i8_pair(tmem_base + 256, al, bh, idesc, 1);
          }
          tc_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(tmem_full);
        t_loop += clock64() - c1;
      }
      if (l2.dbg) {
        l2.dbg[blockIdx.x * 8 + 0] = (unsigned long long)t_wait;     // issuer: waiting for the epilogue to drain TMEM
        l2.dbg[blockIdx.x * 8 + 1] = (unsigned long long)t_loop;     // issuer: k loops (issue + waiting for stages)
        l2.dbg[blockIdx.x * 8 + 2] = (unsigned long long)t_fill;     // issuer: of which waiting for TMA fills
        l2.dbg[blockIdx.x * 8 + 3] = (unsigned long long)it;
      }
    }
   }
  } else {
    // the pool is what the CTA got at launch (640 x 96): 16 x 104 + 4 x 40 warp-registers fit, 16 x 112 + 4 x 40 do not
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===== epilogue (both CTAs): 16 warps = 4 TMEM lane quadrants x 4 quarters of the hidden units =====
    const int half = warp >> 2;                                    // which quarter of the hidden units (0..3)
    const int quad = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    const int eall = threadIdx.x;                                 // 0..511
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int Hh = H >> 2;                                        // hidden units per quarter: 32 or 64
    const int C = l2.C;
    const float invN = l2.scale / (float)l2.N;
    const int r_own = 2 * (t & 1) + (t >> 1);                     // the row (of this thread's 4) whose loss / dZ2 it computes
    const int row_own = quad * 32 + g + 8 * r_own;
    const int pos0 = quad * 32 + 4 * g;                           // storage position of this thread's 4 rows (fused_row_pos)
    const int hbase = half * Hh + 2 * t;
    uint16_t* z2_hi = reinterpret_cast<uint16_t*>(l2.z2_hi);
    uint16_t* z2_lo = reinterpret_cast<uint16_t*>(l2.z2_lo);
    const uint32_t w2h_a = smem_u32(w2h_s), w2l_a = smem_u32(w2l_s);
    // ldmatrix row addresses of this lane (bytes inside a [units][16 classes] bf16 array with 32-byte rows)
    const uint32_t offA = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * 32 + (lane >> 4) * 16);   // .trans: k = units
    const uint32_t offB = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * 32 + ((lane >> 3) & 1) * 16);   // plain:  k = classes
    for (int i = eall; i < 256 * 16; i += 512) w2st_s[i] = 0.f;   // padded classes (c >= C) stay zero
    if (eall < 32) b2_s[eall] = 0.f;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    auto fetch_consts = [&](int item, int buf) {
      int b, mp, split;
      tc_decode(p, item, b, mp, split);
      const float* th = l2.theta + (int64_t)b * l2.P;
      if (eall < H) {
        const int h = eall;
        const float* src = th + l2.w2_off + (int64_t)h * C;
        const uint32_t dst = smem_u32(w2st_s + h * 16);
#pragma unroll
        for (int c = 0; c < CP; ++c)
          if (c < C) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)(c * 4)), "l"(src + c) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(bias_s + buf * 256 + h)),
                     "l"(p.bias + (int64_t)b * p.bias_stride + h) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(cw_s + buf * 256 + h)),
                     "l"(l2.cw + (int64_t)b * H + h) : "memory");
        if (ZI8)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(zq_s + buf * 256 + h)),
                       "l"(l2.zq + (int64_t)b * H + h) : "memory");
      }
      if (eall < C)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(b2_s + buf * 16 + eall)),
                     "l"(th + l2.b2_off + eall) : "memory");
    };
    if (cluster_id < p.total_items) fetch_consts(cluster_id, 0);
    int it = 0, b_prev = -1, cbuf = 1;
    const bool ce = l2.loss_kind == PYB_LOSS_SPARSE_CE;
    long long e_top = 0, e_wait = 0, e_A = 0, e_rest = 0, e_x = 0, e_sm = 0;
    const int dfl = l2.dbg_flags;
    for (int item = cluster_id; item < p.total_items; item += n_clusters, ++it) {
      const long long s0 = clock64();
      int b, mp, split;
      tc_decode(p, item, b, mp, split);
      const int mt = mp * 2 + (int)rank;
      const bool new_chain = b != b_prev;                          // uniform over the CTA: constants and W2 fragments follow the chain
      b_prev = b;
      cbuf ^= new_chain ? 1 : 0;
      const float* bsb = bias_s + cbuf * 256 + hbase;
      const float* cwb = cw_s + cbuf * 256 + hbase;
      const float* zqb = zq_s + cbuf * 256 + hbase;
      const float* b2b = b2_s + cbuf * 16;
      float sxr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int rg = mt * 128 + quad * 32 + g + 8 * r;
        sxr[r] = rg < p.M_valid ? __ldg(l2.sx + rg) : 0.f;
      }
      const int row_g = mt * 128 + row_own;
      const int yi = (ce && row_g < p.M_valid) ? __ldg(l2.y_i + row_g) : 0;     // long before the softmax needs it
      if (new_chain) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("bar.sync 1, 512;" ::: "memory");             // staging complete; every reader of the last chain's fragments is done
      if (eall < H) {
        // this chain's W2 row -> bf16 hi / lo (32-byte rows): what ldmatrix reads in both phases
        const float4* src = reinterpret_cast<const float4*>(w2st_s + eall * 16);
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 v = src[q4];
          split_pair(v.x, v.y, hw[2 * q4], lw[2 * q4]);
          split_pair(v.z, v.w, hw[2 * q4 + 1], lw[2 * q4 + 1]);
        }
        uint4* dh = reinterpret_cast<uint4*>(w2h_s + eall * 32);
        uint4* dl = reinterpret_cast<uint4*>(w2l_s + eall * 32);
        dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
        dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");             // fragments visible; the staging rows may be refilled
      }
      const long long s1 = clock64();
      mbar_wait_sleep(tmem_full, (uint32_t)(it & 1));
      tc_fence_after();
      const long long s2 = clock64();
      const int64_t blk_e = ((((int64_t)b * p.out_tiles + mt) * H) + hbase) * 128 + pos0;     // element index in the block layout
      uint2* pa_hi = reinterpret_cast<uint2*>(p.out_hi) + (blk_e >> 2);
      uint2* pa_lo = reinterpret_cast<uint2*>(p.out_lo) + (blk_e >> 2);
      // ---- phase A: a1 = relu(z1) -> A1^T hi/lo, mask bits, partial logits on mma.sync
      float acc[2][2][4];                                          // [rows {g,g+8} | {g+16,g+24}][classes 0-7 | 8-15][fragment]
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[a][n][e] = 0.f;
      uint32_t mask[2] = {0u, 0u};
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        if (ch * 32 < Hh) {
          uint32_t m = 0u;
#pragma unroll
          for (int kbp = 0; kbp < 2; ++kbp) {                      // one k16 tile of hidden units = two 8-column blocks
            const uint32_t col = (uint32_t)(half * Hh + ch * 32 + kbp * 16);
            uint32_t rh[2][8], rc[2][8];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              tc_ld_16x256b_x2_raw(tmem_base + lane_addr + ((uint32_t)(16 * hf) << 16) + col, rh[hf]);
              tc_ld_16x256b_x2_raw(tmem_base + lane_addr + ((uint32_t)(16 * hf) << 16) + 256u + col, rc[hf]);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t ahw[4][2], alw[4][2];                          // [row][block of the pair]: bf16 pairs (units 2t, 2t+1)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int kb = 2 * kbp + kk;
              const float2 bb = *reinterpret_cast<const float2*>(bsb + ch * 32 + 8 * kb);
              const float2 cc = *reinterpret_cast<const float2*>(cwb + ch * 32 + 8 * kb);
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const int j = kk * 4 + (r & 1) * 2;
                const float s0 = fmaf((float)(int)rc[r >> 1][j], 1.0f / 254.0f, (float)(int)rh[r >> 1][j]);
                const float s1 = fmaf((float)(int)rc[r >> 1][j + 1], 1.0f / 254.0f, (float)(int)rh[r >> 1][j + 1]);
                const float a0 = fmaxf(fmaf(s0, sxr[r] * cc.x, bb.x), 0.f);
                const float a1 = fmaxf(fmaf(s1, sxr[r] * cc.y, bb.y), 0.f);
                m |= (a0 > 0.f ? 1u : 0u) << (kb * 8 + r * 2);
                m |= (a1 > 0.f ? 1u : 0u) << (kb * 8 + r * 2 + 1);
                split_pair(a0, a1, ahw[r][kk], alw[r][kk]);
              }
              // A1^T: unit i of this block, this thread's 4 rows adjacent -> one 8-byte store per array
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const uint32_t sel = i ? 0x7632u : 0x5410u;
                const int w_off = (ch * 32 + 8 * kb + i) * 32;
                if (dfl & 4) continue;
                __stcs(pa_hi + w_off, make_uint2(__byte_perm(ahw[0][kk], ahw[1][kk], sel), __byte_perm(ahw[2][kk], ahw[3][kk], sel)));
                __stcs(pa_lo + w_off, make_uint2(__byte_perm(alw[0][kk], alw[1][kk], sel), __byte_perm(alw[2][kk], alw[3][kk], sel)));
              }
            }
            // logits: [16 rows x 16 units] x [16 units x 16 classes] per row pair group, hi*hi + lo*hi + hi*lo
            uint32_t bh[4], bl[4];
            const uint32_t u0 = (uint32_t)(half * Hh + ch * 32 + kbp * 16) * 32u;
            ldmatrix_x4_trans(w2h_a + u0 + offA, bh);
            ldmatrix_x4_trans(w2l_a + u0 + offA, bl);
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              const uint32_t Ah[4] = {ahw[2 * a][0], ahw[2 * a + 1][0], ahw[2 * a][1], ahw[2 * a + 1][1]};
              const uint32_t Al[4] = {alw[2 * a][0], alw[2 * a + 1][0], alw[2 * a][1], alw[2 * a + 1][1]};
#pragma unroll
              for (int n = 0; n < 2; ++n) {
                if (dfl & 16) { acc[a][n][0] += __uint_as_float(Ah[n] ^ bh[n]); continue; }
                mma_bf16_16816(acc[a][n], Ah, bh[2 * n], bh[2 * n + 1]);
                mma_bf16_16816(acc[a][n], Al, bh[2 * n], bh[2 * n + 1]);
                mma_bf16_16816(acc[a][n], Ah, bl[2 * n], bl[2 * n + 1]);
              }
            }
          }
          mask[ch] = m;
        }
      }
      // the accumulators are no longer needed: hand them back to the tcgen05 issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader_relaxed(tmem_empty);
      const long long s3 = clock64();
      if (item + n_clusters < p.total_items) {
        int bn, mpn, sn;
        tc_decode(p, item + n_clusters, bn, mpn, sn);
        if (bn != b) fetch_consts(item + n_clusters, cbuf ^ 1);
      }
      // ---- the partial logits of the two unit halves meet in shared memory (lane t holds classes 2t, 2t+1, 8+2t, 9+2t)
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int row = quad * 32 + g + 8 * (2 * a + (e >> 1));
            const int cls = 8 * n + 2 * t + (e & 1);
            zx_s[(half * 16 + cls) * 128 + row] = acc[a][n][e];
          }
      // the four warps of a TMEM lane quadrant (one per quarter of the units) share these rows and nobody else does
      asm volatile("bar.sync %0, 128;" ::"r"(4 + quad) : "memory");
      const long long s4 = clock64();
      float zf[CP], dz[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c)
        zf[c] = ((b2b[c] + zx_s[c * 128 + row_own]) + zx_s[(16 + c) * 128 + row_own]) +
                (zx_s[(32 + c) * 128 + row_own] + zx_s[(48 + c) * 128 + row_own]);
      asm volatile("bar.sync %0, 128;" ::"r"(8 + quad) : "memory");   // all four have read: the next item's partial logits may land
      float loss_r = 0.f;
      if (dfl & 32) {
#pragma unroll
        for (int c = 0; c < CP; ++c) dz[c] = zf[c];
      } else if (ce) {
        // l2_loss_dz's sparse cross-entropy branch with the label already in a register and one exp per class
#pragma unroll
        for (int c = 0; c < CP; ++c) dz[c] = 0.f;
        if (row_g < p.M_valid) {
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < CP; ++c) if (c < C) mx = fmaxf(mx, zf[c]);
          float se = 0.f, zy = 0.f;
#pragma unroll
          for (int c = 0; c < CP; ++c)
            if (c < C) { dz[c] = expf(zf[c] - mx); se += dz[c]; zy = c == yi ? zf[c] : zy; }
          loss_r = logf(se) - (zy - mx);
          const float inv = 1.0f / se;
#pragma unroll
          for (int c = 0; c < CP; ++c)
            if (c < C) dz[c] = (dz[c] * inv - (c == yi ? 1.f : 0.f)) * invN;
        }
      } else
      l2_loss_dz<CP>(l2, row_g, row_g < p.M_valid, zf, dz, loss_r, invN);
      {
        // every quarter holds the same dZ2 rows: quarter q stores / sums the classes c = q (mod 4), quarter 3 the loss
        const int64_t blk2 = (((int64_t)b * p.out_tiles + mt) * L2_CMAX) * 128 + pos0 + r_own;
        const int64_t grp = (int64_t)b * l2.n_groups + mt * 4 + quad;
        float mine = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
          if ((c & 3) == half) {
            __nv_bfloat16 hb, lb;
            split_bf16(dz[c], hb, lb);
            z2_hi[blk2 + c * 128] = __bfloat16_as_ushort(hb);
            z2_lo[blk2 + c * 128] = __bfloat16_as_ushort(lb);
            const float sm = warp_sum(dz[c]);
            if (lane == c) mine = sm;
          }
        }
        if (lane < CP && (lane & 3) == half) l2.b2_partial[grp * L2_CMAX + lane] = mine;
        if (half == 3) {
          const float ls = warp_sum(loss_r);
          if (lane == 0) l2.loss_partial[grp] = (double)ls;
        }
      }
      // ---- dZ2 of the row this lane finished -> bf16 hi / lo A fragments for the whole quad, through shared memory:
      //      slot s of a row = {hi(2s, 2s+1), hi(2s+8, 2s+9), lo(2s, 2s+1), lo(2s+8, 2s+9)} (both halves write the same values)
      {
        uint32_t dh[8], dl[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float x0 = (2 * q < CP) ? dz[(2 * q < CP) ? 2 * q : 0] : 0.f;
          const float x1 = (2 * q + 1 < CP) ? dz[(2 * q + 1 < CP) ? 2 * q + 1 : 0] : 0.f;
          split_pair(x0, x1, dh[q], dl[q]);
        }
        uint4* drow = dz_s + (quad * 32 + g + 8 * r_own) * 4;
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) drow[s4] = make_uint4(dh[s4], dh[4 + s4], dl[s4], dl[4 + s4]);
      }
      __syncwarp();
      uint32_t Dh[2][4], Dl[2][4];                                  // A fragments of dZ2: [row pair group][a0..a3]
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint4 f = dz_s[(quad * 32 + g + 8 * r) * 4 + t];
        Dh[r >> 1][(r & 1)] = f.x; Dh[r >> 1][2 + (r & 1)] = f.y;
        Dl[r >> 1][(r & 1)] = f.z; Dl[r >> 1][2 + (r & 1)] = f.w;
      }
      __syncwarp();                                                 // the rows are re-written by the next item
      const long long s5 = clock64();
      uint2* pz_hi = reinterpret_cast<uint2*>(l2.zt_hi) + (blk_e >> 2);
      uint2* pz_lo = reinterpret_cast<uint2*>(l2.zt_lo) + (blk_e >> 2);
      uint32_t* pzi_hi = reinterpret_cast<uint32_t*>(l2.zi_hi + (ZI8 ? blk_e : 0));
      uint32_t* pzi_lo = reinterpret_cast<uint32_t*>(l2.zi_lo + (ZI8 ? blk_e : 0));
      // ---- phase B: dZ1 = (dZ2 W2^T) * relu'(z1) on mma.sync: m16 rows x n8 units x k16 classes
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        if (ch * 32 < Hh) {
          const uint32_t m = mask[ch];
#pragma unroll
          for (int kbp = 0; kbp < 2; ++kbp) {
            uint32_t bh[4], bl[4];
            const uint32_t v0 = (uint32_t)(half * Hh + ch * 32 + kbp * 16) * 32u;
            ldmatrix_x4(w2h_a + v0 + offB, bh);
            ldmatrix_x4(w2l_a + v0 + offB, bl);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int kb = 2 * kbp + kk;
              float d2[2][4];
#pragma unroll
              for (int a = 0; a < 2; ++a) {
#pragma unroll
                for (int e = 0; e < 4; ++e) d2[a][e] = 0.f;
                if (dfl & 8) { d2[a][0] = __uint_as_float(Dh[a][0] ^ bh[kk]); d2[a][1] = __uint_as_float(Dl[a][1] ^ bl[kk]); continue; }
                mma_bf16_16816(d2[a], Dh[a], bh[2 * kk], bh[2 * kk + 1]);
                mma_bf16_16816(d2[a], Dl[a], bh[2 * kk], bh[2 * kk + 1]);
                mma_bf16_16816(d2[a], Dh[a], bl[2 * kk], bl[2 * kk + 1]);
              }
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                float d[4] = {d2[0][i], d2[0][2 + i], d2[1][i], d2[1][2 + i]};     // rows g, g+8, g+16, g+24 of unit 2t + i
#pragma unroll
                for (int r = 0; r < 4; ++r) d[r] = ((m >> (kb * 8 + r * 2 + i)) & 1u) ? d[r] : 0.f;
                const int w_off = (ch * 32 + 8 * kb + i) * 32;
                if ((dfl & 2) && d[0] != 12345.f) continue;
                if (ZI8) {
                  uint32_t hw, lw;
                  slice4_i8(d, zqb[ch * 32 + 8 * kb + i], hw, lw);
                  __stcs(pzi_hi + w_off, hw);
                  __stcs(pzi_lo + w_off, lw);
                } else {
                  uint32_t hw0, lw0, hw1, lw1;
                  split_pair(d[0], d[1], hw0, lw0);
                  split_pair(d[2], d[3], hw1, lw1);
                  __stcs(pz_hi + w_off, make_uint2(hw0, hw1));
                  __stcs(pz_lo + w_off, make_uint2(lw0, lw1));
                }
              }
            }
          }
        }
      }
      e_top += s1 - s0; e_wait += s2 - s1; e_A += s3 - s2; e_rest += clock64() - s5; e_x += s4 - s3; e_sm += s5 - s4;
    }
    if (l2.dbg && warp == 0 && lane == 0 && (blockIdx.x & 1)) {
      l2.dbg[blockIdx.x * 8 + 0] = (unsigned long long)e_x;          // (odd CTAs' slots 0 / 1 are free) logits exchange incl. the barrier
      l2.dbg[blockIdx.x * 8 + 1] = (unsigned long long)e_sm;         // softmax / loss / dZ2 fragments
    }
    if (l2.dbg && warp == 0 && lane == 0) {
      l2.dbg[blockIdx.x * 8 + 4] = (unsigned long long)e_top;        // epilogue warp 0: constants / barriers at the top of an item
      l2.dbg[blockIdx.x * 8 + 5] = (unsigned long long)e_wait;       // waiting for the accumulators
      l2.dbg[blockIdx.x * 8 + 6] = (unsigned long long)e_A;          // phase A (TMEM held)
      l2.dbg[blockIdx.x * 8 + 7] = (unsigned long long)e_rest;       // phase B
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 18) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace pyb

// tc_fused.cuh — layer-1 GEMM with layer 2, the loss and both back-propagated deltas in its epilogue.
#pragma once
#include "tc_layer2.cuh"

namespace pyb {

// ------------------------------------------------------------------------------------------
// G1 + layer 2 in ONE kernel (relu hidden layer, H = 128 or 256): the CTA-pair GEMM above with an epilogue
// that never lets the hidden activations leave the SM before layer 2 has consumed them.
// The epilogue reads the TMEM accumulator with tcgen05.ld.16x256b: lane (g = lane/4, t = lane%4) of a warp
// receives, for each 8-column block, columns {2t, 2t+1} of rows g and g+8 (two loads: + rows g+16, g+24).
// One thread therefore owns FOUR data rows x a quarter of the hidden units of its warp's column half, and every
// W2 row it fetches from shared memory feeds 4 rows (the 32x32b layout, thread == row, re-reads W2 for every
// row: measured shared-memory-bandwidth bound, 12.6 ms against 7.7 ms of MMA work).  Per 128-row tile of a chain:
//   phase A  a1 = relu(z1 + b1) -> A1^T hi/lo (kept for the dW2 GEMM), relu mask bits in registers,
//            partial logits z2 += a1 * W2 (packed fp32x2 FMAs); TMEM accumulator released to the MMA warp
//   reduce   partial logits: quad reduce-scatter by shuffles (lane t ends up with ONE complete row), the two
//            column halves meet in shared memory; softmax-CE / MSE and dZ2 once per row; quad all-gather of dZ2
//   phase B  dZ1 = (dZ2 W2^T) * mask -> dZ1^T hi/lo for the dW1 GEMM
// Transposed stores: inside a 128-row block the rows are kept in the order fused_row_pos() — the 4 rows one thread
// owns are adjacent, so a hidden unit's 4 values leave as ONE 8-byte store and the 8 lanes sharing t write 64
// contiguous bytes; the dW2 GEMM contracts two arrays written this way, the dW1 GEMM uses an [X^T;1] copy in the
// same row order.
// ------------------------------------------------------------------------------------------
constexpr int TF_THREADS = 384;                 // 8 epilogue warps + 4 control warps
template <int CP> struct TfCfg {
  static constexpr int STAGES = 5;
  static constexpr int W2_BYTES = 2 * 256 * CP * 4;            // [2][256*CP] fp32 (fragment-interleaved), double-buffered
  static constexpr int ZX_BYTES = 2 * CP * 128 * 4;            // [2 halves][CP][128 rows] partial logits
  static constexpr int SMEM = STAGES * TP_STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers*/ + 2048 /*bias x2*/ +
                              128 /*b2 x2*/ + W2_BYTES + ZX_BYTES;
  static constexpr int SMEM_I8 = SMEM + 4096;                  // + [2][256] W1 column scales, [2][256] dZ1 quantisation factors
};
// same 16x256b.x4 load, raw 32-bit words (the int32 accumulators of the int8-slice scheme)
__device__ __forceinline__ void tc_ld_16x256b_x4_raw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// two int8 slices of 4 values d[r] * zq (|d * zq| <= 127 by the caller's bound): hi = rint(q), lo = rint((q - hi) * 254),
// one byte per row, row 0 in the lowest byte.  Round-to-nearest through the float32 magic number 1.5 * 2^23: the low
// mantissa byte of q + M is rint(q) in two's complement — four full-rate FP instructions per value instead of two
// FRND + two F2I (quarter rate) and the clamps; q - hi comes out of one fused multiply-add, so it is exact.
__device__ __forceinline__ void slice4_i8(const float* d, float zq, uint32_t& hw, uint32_t& lw) {
  constexpr float M = 12582912.0f;
  uint32_t hb[4], lb[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float t = fmaf(d[r], zq, M);                // M + rint(q)
    const float h = t - M;                            // rint(q)
    const float e = fmaf(d[r], zq, -h);               // q - rint(q), exact
    const float u = fmaf(e, 254.0f, M);               // M + rint((q - hi) * 254)
    hb[r] = __float_as_uint(t);
    lb[r] = __float_as_uint(u);
  }
  // byte 0 of each word -> one packed word
  hw = __byte_perm(__byte_perm(hb[0], hb[1], 0x0040), __byte_perm(hb[2], hb[3], 0x0040), 0x5410);
  lw = __byte_perm(__byte_perm(lb[0], lb[1], 0x0040), __byte_perm(lb[2], lb[3], 0x0040), 0x5410);
}

// 16 accumulator columns... 32 columns x rows {g, g+8} of the 16 TMEM lanes starting at the address's lane
__device__ __forceinline__ void tc_ld_16x256b_x4(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// W2[h][c] inside the per-chain shared-memory copy: the 4 lanes of a quad (h = 8*kb + 2*t + i) read one
// contiguous 64-byte segment per (kb, i, c/4) -> conflict-free for every class padding CP
template <int CP>
__device__ __forceinline__ int w2_slot(int h, int c) {
  return ((((h >> 3) * 2 + (h & 1)) * (CP / 4) + (c >> 2)) * 4 + ((h >> 1) & 3)) * 4 + (c & 3);
}
// bf16 hi/lo words of a row pair (x0 = even row, x1 = odd row)
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hw, uint32_t& lw) {
  const __nv_bfloat162 hp = __floats2bfloat162_rn(x0, x1);
  hw = *reinterpret_cast<const uint32_t*>(&hp);
  const __nv_bfloat162 lp = __floats2bfloat162_rn(x0 - __uint_as_float(hw << 16), x1 - __uint_as_float(hw & 0xffff0000u));
  lw = *reinterpret_cast<const uint32_t*>(&lp);
}

// FWD = false: training step (everything above).  FWD = true: forward only (posterior predictive): phase A, the logit
// reduction and the output activation; out[b][row][c] is the only thing written — no A1^T, no deltas, no loss.
// I8 != 0: the int8-slice scheme (tc_i8.cuh).  The four tensor maps address int8 slice tensors (X rows centred by the
// feature means and scaled per row, W1^T scaled per hidden unit), a stage covers 64 K-elements, kind::i8 MMAs accumulate
// hi*hi into TMEM columns [0, 256) and hi*lo + lo*hi (weight 1/254) into [256, 512) as EXACT int32 — half the MMA slots
// and half the operand bytes of bf16x3, no truncating accumulator.  Both accumulators fill TMEM, so the accumulator is
// single-buffered: phase A of the epilogue is exposed, the rest of it still overlaps the next tile's main loop.
// z1 = s_x[row] * (s_w[unit] / 127^2) * (hh + cross / 254) + b1'[unit], b1' = b1 + mu^T W1 (the data were centred).
// I8 == 2: dZ1^T leaves as two int8 slices against the a-priori scale s_z[unit] >= max |dZ1[:, unit]| (softmax-CE only).
template <int CP, bool FWD, int I8 = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TF_THREADS, 1)
tc_g1_layer2_fused(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   const TcGemmParams p, const Layer2Params l2) {
  constexpr int STAGES = TfCfg<CP>::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint8_t* stage_base = smem;
  uint64_t* bars = (uint64_t*)(smem + STAGES * TP_STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]   (leader CTA)
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]   (one per CTA)
  uint64_t* tmem_full = bars + 2 * STAGES;       // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]        (leader CTA)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * STAGES + 4);
  float* bias_s = (float*)(smem + STAGES * TP_STAGE_BYTES + 1024);          // [2][256]
  float* b2_s = bias_s + 512;                                               // [2][16]
  float* W2_s = b2_s + 32;                                                  // [2][256*CP]
  float* zx_s = W2_s + 2 * 256 * CP;                                        // [2][CP][128]
  float* cw_s = zx_s + 2 * CP * 128;                                        // [2][256]   (I8)
  float* zq_s = cw_s + 512;                                                 // [2][256]   (I8 == 2)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  constexpr int KSTAGE = I8 ? 64 : TC_BK;         // K elements per stage (a stage row is 64 bytes either way)
  const int nk = (p.K + KSTAGE - 1) / KSTAGE;
  const int H = p.H;
  const int half_rows = H >> 1;
  const uint32_t cta_bytes = 2 * TC_A_TILE_BYTES + 2 * (uint32_t)half_rows * TC_BK * 2;

  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 16); mbar_init(&tmem_empty[1], 16);   // 8 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 10) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // warp roles: 0-7 epilogue, 8 TMA producer, 9 MMA issuer, 10 TMEM allocator.  The SMSP arbiter favours the
  // highest warp id, so the single-thread issuers sit ABOVE the epilogue warps and never queue behind them.
  if (warp >= 8) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");                  // the control warpgroup gives registers ...
   if (warp == 8) {
    // ===== TMA producer (both CTAs): own 128 rows of X, own half of W1^T[b] =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        int b, mp, split;
        tc_decode(p, item, b, mp, split);
        const int arow = p.a_row0 + (mp * 2 + (int)rank) * 128;
        const int brow = b * H + (int)rank * half_rows;
        for (int kc = 0; kc < nk; ++kc) {
          (I8 ? mbar_wait_sleep(&empty_bar[stage], phase ^ 1) : mbar_wait(&empty_bar[stage], phase ^ 1));
          uint8_t* st = stage_base + stage * TP_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * cta_bytes);
          const int k0 = kc * KSTAGE;
          tma_load_2d_pair(st, &tmA_hi, &full_bar[stage], k0, arow);
          tma_load_2d_pair(st + TC_A_TILE_BYTES, &tmA_lo, &full_bar[stage], k0, arow);
          tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES, &tmB_hi, &full_bar[stage], k0, brow);
          tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES + 8192, &tmB_lo, &full_bar[stage], k0, brow);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0 && lane == 0) {
      // bf16: D = F32, A = B = BF16; int8 slices: D = S32, A = B = signed 8-bit; K-major, N = H, M = 256 across the pair
      const uint32_t idesc = (I8 ? (2u << 4) : (1u << 4)) | (1u << 7) | (1u << 10) | ((uint32_t)(H >> 3) << 17) | ((256u >> 4) << 24);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      const int k_tail = p.K - (nk - 1) * KSTAGE;                // valid K elements of the last stage
      for (int item = cluster_id; item < p.total_items; item += n_clusters, ++it) {
        const int acc = I8 ? 0 : (it & 1);
        const uint32_t acc_phase = I8 ? (uint32_t)(it & 1) : (uint32_t)((it >> 1) & 1);
        (I8 ? mbar_wait_sleep(&tmem_empty[acc], acc_phase ^ 1) : mbar_wait(&tmem_empty[acc], acc_phase ^ 1));
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * 256;
        for (int kc = 0; kc < nk; ++kc) {
          (I8 ? mbar_wait_sleep(&full_bar[stage], phase) : mbar_wait(&full_bar[stage], phase));
          tc_fence_after();
          const uint32_t st = smem_u32(stage_base + stage * TP_STAGE_BYTES);
          const int nks = (kc == nk - 1 && k_tail <= KSTAGE / 2) ? 1 : 2;
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t koff = ks * 32;
            const uint64_t ah = make_smem_desc_sw64(st + koff);
            const uint64_t al = make_smem_desc_sw64(st + TC_A_TILE_BYTES + koff);
            const uint64_t bh = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + koff);
            const uint64_t bl = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + 8192 + koff);
            const uint32_t accum = (kc != 0) || (ks != 0);
            if (I8) {
              tc_mma_i8_pair(d, ah, bh, idesc, accum);            // hi * hi           -> columns [0, 256)
              tc_mma_i8_pair(d + 256, ah, bl, idesc, accum);      // hi * lo + lo * hi -> columns [256, 512)
              tc_mma_i8_pair(d + 256, al, bh, idesc, 1);
            } else {
              tc_mma_bf16_pair(d, ah, bh, idesc, accum);
              tc_mma_bf16_pair(d, al, bh, idesc, 1);
              tc_mma_bf16_pair(d, ah, bl, idesc, 1);
            }
          }
          tc_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(&tmem_full[acc]);
      }
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");                // ... to the two epilogue warpgroups
    // ===== fused epilogue (both CTAs): 8 warps = 4 TMEM lane quadrants x 2 halves of the hidden units =====
    const int half = warp >> 2;
    const int quad = warp & 3;                                    // TMEM lane quadrant (hardware: warp % 4)
    const int g = lane >> 2, t = lane & 3;
    const int eall = threadIdx.x;                                 // 0..255
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int Hh = H >> 1;                                        // hidden units per half: 64 or 128
    const int C = l2.C;
    const float invN = l2.scale / (float)l2.N;
    const int r_own = 2 * (t & 1) + (t >> 1);                     // the row (of this thread's 4) whose logits it completes
    const int row_own = quad * 32 + g + 8 * r_own;                // ... inside the 128-row tile
    const int pos0 = quad * 32 + 4 * g;                           // storage position of this thread's 4 rows (fused_row_pos)
    const int hbase = half * Hh + 2 * t;                          // this thread's first hidden unit
    uint16_t* z2_hi = reinterpret_cast<uint16_t*>(l2.z2_hi);
    uint16_t* z2_lo = reinterpret_cast<uint16_t*>(l2.z2_lo);
    // per-chain constants (W2, b1, b2) travel global -> shared memory with cp.async, issued one item ahead into the
    // buffer of the accumulator that item will use; slots of padded classes (c >= C) are zeroed once and never written
    for (int i = eall; i < 2 * 256 * CP; i += 256) W2_s[i] = 0.f;
    if (eall < 32) b2_s[eall] = 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    auto fetch_consts = [&](int item, int buf) {
      int b, mp, split;
      tc_decode(p, item, b, mp, split);
      const float* th = l2.theta + (int64_t)b * l2.P;
      if (eall < H) {
        const int h = eall;
        const float* src = th + l2.w2_off + (int64_t)h * C;
        const uint32_t dst = smem_u32(W2_s + buf * 256 * CP + w2_slot<CP>(h, 0));
#pragma unroll
        for (int c = 0; c < CP; ++c)
          if (c < C)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)(((c >> 2) * 16 + (c & 3)) * 4)),
                         "l"(src + c) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(bias_s + buf * 256 + h)),
                     "l"(p.bias + (int64_t)b * p.bias_stride + h) : "memory");
      }
      if (eall < C)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(b2_s + buf * 16 + eall)),
                     "l"(th + l2.b2_off + eall) : "memory");
      if (I8 && eall < H) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(cw_s + buf * 256 + eall)),
                     "l"(l2.cw + (int64_t)b * H + eall) : "memory");
        if (I8 == 2 && !FWD)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(zq_s + buf * 256 + eall)),
                       "l"(l2.zq + (int64_t)b * H + eall) : "memory");
      }
    };
    if (cluster_id < p.total_items) fetch_consts(cluster_id, 0);
    int it = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters, ++it) {
      int b, mp, split;
      tc_decode(p, item, b, mp, split);
      const int mt = mp * 2 + (int)rank;
      const int cbuf = it & 1;                                    // buffer of the per-chain constants
      const int acc = I8 ? 0 : cbuf;                              // TMEM accumulator (single-buffered with int8 slices)
      const uint32_t acc_phase = I8 ? (uint32_t)(it & 1) : (uint32_t)((it >> 1) & 1);
      float* bs = bias_s + cbuf * 256;
      float* b2b = b2_s + cbuf * 16;
      float* W2b = W2_s + cbuf * 256 * CP;
      float sxr[4] = {0.f, 0.f, 0.f, 0.f};                        // row scales of this thread's 4 data rows (int8 slices)
      if (I8) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int rg = mt * 128 + quad * 32 + g + 8 * r;
          sxr[r] = rg < p.M_valid ? __ldg(l2.sx + rg) : 0.f;
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");           // this thread's share of the constants has landed
      asm volatile("bar.sync 1, 256;" ::: "memory");             // constants visible; zx_s readers of the last item done
      (I8 ? mbar_wait_sleep(&tmem_full[acc], acc_phase) : mbar_wait(&tmem_full[acc], acc_phase));
      tc_fence_after();
      // (hidden unit hbase, this thread's 4 rows) inside this (chain, tile) block, in 8-byte units (4 bf16)
      const int64_t blk_w = (((((int64_t)b * p.out_tiles + mt) * H) + hbase) * 128 + pos0) >> 2;
      uint2* pa_hi = reinterpret_cast<uint2*>(p.out_hi) + blk_w;
      uint2* pa_lo = reinterpret_cast<uint2*>(p.out_lo) + blk_w;
      uint2* pz_hi = reinterpret_cast<uint2*>(l2.zt_hi) + blk_w;
      uint2* pz_lo = reinterpret_cast<uint2*>(l2.zt_lo) + blk_w;
      const float4* w4b = reinterpret_cast<const float4*>(W2b) + (((half * Hh) >> 3) * 2 * (CP / 4)) * 4 + t;
      const float* bsb = bs + hbase;
      const float* cwb = cw_s + cbuf * 256 + hbase;
      const float* zqb = zq_s + cbuf * 256 + hbase;
      uint32_t* pzi_hi = nullptr; uint32_t* pzi_lo = nullptr;      // dZ1^T int8 slices: 4 rows = one 32-bit word
      if (I8 == 2 && !FWD) {
        const int64_t blk_b = ((((int64_t)b * p.out_tiles + mt) * H) + hbase) * 128 + pos0;   // byte == element
        pzi_hi = reinterpret_cast<uint32_t*>(l2.zi_hi + blk_b);
        pzi_lo = reinterpret_cast<uint32_t*>(l2.zi_lo + blk_b);
      }
      // ---- phase A  (rows >= M_valid need no masking: their X rows are TMA zero fill, so a1 = relu(b1) stays
      //      finite, and their dZ2 is zero, which zeroes dZ1 and every gradient contribution)
      float2 z[4][CP / 2];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < CP / 2; ++c) z[r][c] = make_float2(0.f, 0.f);
      uint32_t mask[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch * 32 < Hh) {
          const uint32_t col = (uint32_t)(acc * 256 + half * Hh + ch * 32);
          float v[32];                                             // [rows g, g+8 | rows g+16, g+24][4 col blocks][2 rows][2 cols]
          if (I8) {
            // exact int32 sums: hh (|.| <= 127^2 K < 2^24 for K <= 1040: exact in fp32) + cross / 254
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t rh[16], rc[16];
              tc_ld_16x256b_x4_raw(tmem_base + lane_addr + ((uint32_t)(16 * hf) << 16) + col, rh);
              tc_ld_16x256b_x4_raw(tmem_base + lane_addr + ((uint32_t)(16 * hf) << 16) + 256u + col, rc);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 16; ++j)
                v[16 * hf + j] = fmaf((float)(int)rc[j], 1.0f / 254.0f, (float)(int)rh[j]);
            }
          } else {
            tc_ld_16x256b_x4(tmem_base + lane_addr + col, v);
            tc_ld_16x256b_x4(tmem_base + lane_addr + (16u << 16) + col, v + 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
          uint32_t m = 0u;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const float2 bb = *reinterpret_cast<const float2*>(bsb + ch * 32 + 8 * kb);
            float2 cc = make_float2(0.f, 0.f);
            if (I8) cc = *reinterpret_cast<const float2*>(cwb + ch * 32 + 8 * kb);
            float a[4][2];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float* vv = v + (r >> 1) * 16 + kb * 4 + (r & 1) * 2;
              if (I8) {
                a[r][0] = fmaxf(fmaf(vv[0], sxr[r] * cc.x, bb.x), 0.f);
                a[r][1] = fmaxf(fmaf(vv[1], sxr[r] * cc.y, bb.y), 0.f);
              } else {
                a[r][0] = fmaxf(vv[0] + bb.x, 0.f);
                a[r][1] = fmaxf(vv[1] + bb.y, 0.f);
              }
              if (!FWD) {
                m |= (a[r][0] > 0.f ? 1u : 0u) << (kb * 8 + r * 2);
                m |= (a[r][1] > 0.f ? 1u : 0u) << (kb * 8 + r * 2 + 1);
              }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              // logits: one W2 row from shared memory feeds this thread's 4 rows
              const float4* w4 = w4b + (((ch * 4 + kb) * 2 + i) * (CP / 4)) * 4;
#pragma unroll
              for (int c4 = 0; c4 < CP / 4; ++c4) {
                const float4 w = w4[c4 * 4];
                const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const float2 aa = make_float2(a[r][i], a[r][i]);
                  z[r][2 * c4] = __ffma2_rn(aa, w01, z[r][2 * c4]);
                  z[r][2 * c4 + 1] = __ffma2_rn(aa, w23, z[r][2 * c4 + 1]);
                }
              }
              if (!FWD) {
                // A1^T: this thread's 4 rows of hidden unit hbase + ... are adjacent in the block's row order
                uint32_t hw0, lw0, hw1, lw1;
                split_pair(a[0][i], a[1][i], hw0, lw0);
                split_pair(a[2][i], a[3][i], hw1, lw1);
                const int w_off = (ch * 32 + 8 * kb + i) * 32;
                __stcs(pa_hi + w_off, make_uint2(hw0, hw1));     // streaming: 18 GB per launch must not evict X / W1^T from L2
                __stcs(pa_lo + w_off, make_uint2(lw0, lw1));
              }
            }
          }
          mask[ch] = m;
        }
      }
      // the accumulator is no longer needed: hand it back to the MMA warp before the rest of the epilogue
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader_relaxed(&tmem_empty[acc]);
      // next item's constants fly while this item's reductions and phase B run; buffer [cbuf ^ 1] was last read in
      // the previous item, which every epilogue thread has left (they all passed this item's bar.sync 1)
      if (item + n_clusters < p.total_items) fetch_consts(item + n_clusters, cbuf ^ 1);
      // ---- partial logits: quad reduce-scatter (fixed order), lane t keeps row r_own
      float zo[CP];
      {
        const bool b0 = t & 1, b1 = t & 2;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
          const float z0 = (c & 1) ? z[0][c >> 1].y : z[0][c >> 1].x, z1 = (c & 1) ? z[1][c >> 1].y : z[1][c >> 1].x;
          const float z2v = (c & 1) ? z[2][c >> 1].y : z[2][c >> 1].x, z3 = (c & 1) ? z[3][c >> 1].y : z[3][c >> 1].x;
          // stage 1 (xor 1): lanes with t&1 == 0 keep rows {0,1}, the others rows {2,3}
          float k0 = b0 ? z2v : z0, k1 = b0 ? z3 : z1;
          const float g0 = __shfl_xor_sync(0xffffffffu, b0 ? z0 : z2v, 1);
          const float g1 = __shfl_xor_sync(0xffffffffu, b0 ? z1 : z3, 1);
          k0 = b0 ? g0 + k0 : k0 + g0;                            // always (t even) + (t odd)
          k1 = b0 ? g1 + k1 : k1 + g1;
          // stage 2 (xor 2): lanes with t&2 == 0 keep the first of their two rows
          const float kk = b1 ? k1 : k0;
          const float gg = __shfl_xor_sync(0xffffffffu, b1 ? k0 : k1, 2);
          zo[c] = b1 ? gg + kk : kk + gg;                         // always (t < 2) + (t >= 2)
        }
      }
      // ---- the two column halves meet in shared memory
#pragma unroll
      for (int c = 0; c < CP; ++c) zx_s[(half * CP + c) * 128 + row_own] = zo[c];
      asm volatile("bar.sync 2, 256;" ::: "memory");
      float zf[CP], dz[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) zf[c] = (b2b[c] + zx_s[c * 128 + row_own]) + zx_s[(CP + c) * 128 + row_own];
      const int row_g = mt * 128 + row_own;
      if (FWD) {
        // predictive: softmax / output activation of the finished row, one thread per row
        if (half == 0 && row_g < p.M_valid) {
          float* o = l2.fwd_out + ((int64_t)b * l2.N + row_g) * C;
          if (l2.out_act == PYB_ACT_SOFTMAX) {
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < CP; ++c) if (c < C) mx = fmaxf(mx, zf[c]);
            float se = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) if (c < C) se += expf(zf[c] - mx);
            const float inv = 1.0f / se;
#pragma unroll
            for (int c = 0; c < CP; ++c) if (c < C) o[c] = expf(zf[c] - mx) * inv;
          } else {
#pragma unroll
            for (int c = 0; c < CP; ++c) if (c < C) o[c] = act_apply(zf[c], l2.out_act);
          }
        }
        continue;
      }
      float loss_r = 0.f;
      l2_loss_dz<CP>(l2, row_g, row_g < p.M_valid, zf, dz, loss_r, invN);
      if (half == 0) {
        // dZ2^T for the dW2 GEMM, per-warp partial sums of the loss and of db2 (one warp == 32 rows)
        const int64_t blk2 = (((int64_t)b * p.out_tiles + mt) * L2_CMAX) * 128 + pos0 + r_own;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
          __nv_bfloat16 hb, lb;
          split_bf16(dz[c], hb, lb);
          z2_hi[blk2 + c * 128] = __bfloat16_as_ushort(hb);
          z2_lo[blk2 + c * 128] = __bfloat16_as_ushort(lb);
        }
        const int64_t grp = (int64_t)b * l2.n_groups + mt * 4 + quad;
        float mine = 0.f;                                         // lane c keeps the sum of class c
#pragma unroll
        for (int c = 0; c < CP; ++c) {
          const float sm = warp_sum(dz[c]);
          if (lane == c) mine = sm;
        }
        if (lane < CP) l2.b2_partial[grp * L2_CMAX + lane] = mine;
        const float ls = warp_sum(loss_r);                        // 32 rows in fp32; the per-chain total is summed in fp64
        if (lane == 0) l2.loss_partial[grp] = (double)ls;
      }
      // ---- quad all-gather of dZ2: row r lives in lane t = (r >> 1) | ((r & 1) << 1)
      float2 dzp[4][CP / 2];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int src = (lane & ~3) | ((r >> 1) | ((r & 1) << 1));
#pragma unroll
        for (int c = 0; c < CP / 2; ++c) {
          dzp[r][c].x = __shfl_sync(0xffffffffu, dz[2 * c], src);
          dzp[r][c].y = __shfl_sync(0xffffffffu, dz[2 * c + 1], src);
        }
      }
      // ---- phase B: dZ1 = (dZ2 W2^T) * relu'(z1)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch * 32 < Hh) {
          const uint32_t m = mask[ch];
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float4* w4 = w4b + (((ch * 4 + kb) * 2 + i) * (CP / 4)) * 4;
              float2 s[4];
#pragma unroll
              for (int r = 0; r < 4; ++r) s[r] = make_float2(0.f, 0.f);
#pragma unroll
              for (int c4 = 0; c4 < CP / 4; ++c4) {
                const float4 w = w4[c4 * 4];
                const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  s[r] = __ffma2_rn(dzp[r][2 * c4], w01, s[r]);
                  s[r] = __ffma2_rn(dzp[r][2 * c4 + 1], w23, s[r]);
                }
              }
              float d[4];
#pragma unroll
              for (int r = 0; r < 4; ++r) d[r] = ((m >> (kb * 8 + r * 2 + i)) & 1u) ? s[r].x + s[r].y : 0.f;
              const int w_off = (ch * 32 + 8 * kb + i) * 32;
              if (I8 == 2) {
                uint32_t hw, lw;
                slice4_i8(d, zqb[ch * 32 + 8 * kb + i], hw, lw);
                __stcs(pzi_hi + w_off, hw);                      // (unit, 4 adjacent rows): 32 words per 128-byte block row
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
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 10) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace pyb

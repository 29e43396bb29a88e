// tc_gemm.cuh — the persistent, warp-specialised bf16x3 split GEMM kernels (TMA -> mbarrier ring -> tcgen05.mma ->
// TMEM -> epilogue): the 1-CTA kernel (two accumulators sharing each B stage), the CTA-pair kernel (cta_group::2,
// double-buffered accumulator) and the dual-accumulator hidden-major gradient GEMM.
#pragma once
#include "tc_ptx.cuh"

namespace pyb {

// ------------------------------------------------------------------------------------------
// GEMM kernel
// ------------------------------------------------------------------------------------------
constexpr int TC_BK = 32;                       // K elements per stage
constexpr int TC_STAGES = 3;
constexpr int TC_A_TILE_BYTES = 128 * TC_BK * 2;    // 8 KB
constexpr int TC_B_TILE_BYTES = 256 * TC_BK * 2;    // 16 KB (H <= 256)
constexpr int TC_STAGE_BYTES = 4 * TC_A_TILE_BYTES + 2 * TC_B_TILE_BYTES;   // 64 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers + bias*/ + 1024;
constexpr int TC_THREADS = 640;

enum { EPI_BIAS_ACT_T_SPLIT = 0, EPI_STORE = 1 };

struct TcGemmParams {
  int K, n_mtiles, n_pairs, n_batch, H;
  int a_batch_rows;            // A row offset per chain (0: A shared by all chains); unused when A is blocked
  int a_blocked, b_blocked, k_tiles;   // operand addressing (see tma_load_operand); k_tiles = K/128
  int a_box_rows;              // rows of one A TMA box (128, or H when a per-chain A has fewer rows)
  int order, sub_batch, total_items;
  int epi;
  // EPI_BIAS_ACT_T_SPLIT: a = act(D[row][col] + bias[b][col]) -> bf16 hi/lo at [b*H + col][row]  (row < M_valid)
  const float* bias; int64_t bias_stride; int act;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; int out_tiles;   // blocked [b][tile][col][128] output
  // EPI_STORE: out[b*out_stride + row*out_ld + col] = D[row][col]   (row < M_valid, col < N_valid)
  float* out; int64_t out_stride; int out_ld;
  int M_valid, N_valid;
  int a_row0;                  // first A row of this launch (row-sharded callers)
  // split-K: the tcgen05 fp32 accumulator TRUNCATES on every accumulate (measured -4e-8 relative per MMA,
  // -4.5e-4 after 60000-long reductions), so long K loops are cut into k_splits independent accumulations of
  // chunks_per_split 32-element chunks whose partial results are summed afterwards in a fixed order
  int k_splits, chunks_per_split; int64_t split_stride;
  // pair kernel, 'hidden-major' gradient GEMM: chain b selects the (blocked) A operand, bt in [0, n_btiles) selects
  // the B row tile [bt*H, bt*H+H) (+b_row0) and the output is stored transposed: out[(col0 + col)*out_ld + row]
  int n_btiles, b_row0, transpose_out;
  int vec_store;               // EPI_STORE: every output row segment is 16-byte aligned -> float4 stores
  int n_cols_total;            // >0: chain b owns columns [b*H, min((b+1)*H, n_cols_total)) of one wide output
  int chain_pairs, n_chains;   // dual hidden-major kernel with 128 hidden units: the two CTAs of a pair take two CHAINS
                               // (2b, 2b+1 < n_chains) instead of the two halves of one chain's 256 hidden units
  int sym_skip;                // pair kernel, A == B (Gram matrix): tiles strictly below the diagonal (256-col tile b <
                               // 256-row tile mp) are skipped; the caller mirrors the upper triangle afterwards
};

__device__ __forceinline__ void tc_decode(const TcGemmParams& p, int item, int& b, int& mp, int& split) {
  if (p.k_splits > 1) {          // [split][chain][pair]: the pairs sharing one B k-range run side by side, and the
    mp = item % p.n_pairs;       // k-range of the shared A operand stays L2-resident while all chains pass over it
    int r = item / p.n_pairs;
    b = r % p.n_batch;
    split = r / p.n_batch;
    return;
  }
  split = 0;
  if (p.order == 0) {            // chain-major: the pairs of one chain run side by side (G2)
    b = item / p.n_pairs;
    mp = item - b * p.n_pairs;
  } else {                       // sub-batched: [sub-batch][pair][chain in sub-batch] (G1)
    int per_sb = p.n_pairs * p.sub_batch;
    int sb = item / per_sb;
    int rem = item - sb * per_sb;
    int first = sb * p.sub_batch;
    int size = min(p.sub_batch, p.n_batch - first);
    mp = rem / size;
    b = first + (rem - mp * size);
  }
}

template <int ACT>
__device__ __forceinline__ float act_apply_t(float z) {
  if (ACT == PYB_ACT_RELU) return fmaxf(z, 0.0f);
  if (ACT == PYB_ACT_TANH) return tanhf(z);
  if (ACT == PYB_ACT_SIGMOID) return 1.0f / (1.0f + expf(-z));
  return z;
}

template <int EPI, int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_bf16x3(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint8_t* stage_base = smem;
  uint64_t* bars = (uint64_t*)(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [TC_STAGES]
  uint64_t* empty_bar = bars + TC_STAGES;       // [TC_STAGES]
  uint64_t* tmem_full = bars + 2 * TC_STAGES;
  uint64_t* tmem_empty = bars + 2 * TC_STAGES + 1;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * TC_STAGES + 2);
  float* bias_s = (float*)(smem + TC_STAGES * TC_STAGE_BYTES + 1024);   // [256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (p.K + TC_BK - 1) / TC_BK;
  const uint32_t b_bytes = (uint32_t)p.H * TC_BK * 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 512);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int b, mp, split;
        tc_decode(p, item, b, mp, split);
        const int kc_begin = split * p.chunks_per_split, kc_end = min(nk, kc_begin + p.chunks_per_split);
        const int mt0 = mp * 2;
        const int n_mt = (mt0 + 1 < p.n_mtiles) ? 2 : 1;
        const uint32_t bytes = (uint32_t)n_mt * 2 * (uint32_t)p.a_box_rows * TC_BK * 2 + 2 * b_bytes;
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stage_base + stage * TC_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], bytes);
          const int k0 = kc * TC_BK;
          for (int mt = 0; mt < n_mt; ++mt) {
            const int arow = (p.a_blocked ? 0 : p.a_row0 + b * p.a_batch_rows) + (mt0 + mt) * 128;
            tma_load_operand(st + mt * TC_A_TILE_BYTES, &tmA_hi, &full_bar[stage], p.a_blocked, k0, arow, b, p.k_tiles);
            tma_load_operand(st + (2 + mt) * TC_A_TILE_BYTES, &tmA_lo, &full_bar[stage], p.a_blocked, k0, arow, b, p.k_tiles);
          }
          const int brow = p.b_blocked ? 0 : b * p.H;
          tma_load_operand(st + 4 * TC_A_TILE_BYTES, &tmB_hi, &full_bar[stage], p.b_blocked, k0, brow, b, p.k_tiles);
          tma_load_operand(st + 4 * TC_A_TILE_BYTES + TC_B_TILE_BYTES, &tmB_lo, &full_bar[stage], p.b_blocked, k0, brow, b,
                           p.k_tiles);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=BF16, both K-major, N = H, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.H >> 3) << 17) | ((128u >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      const int k_tail = p.K - (nk - 1) * TC_BK;                 // valid K elements of the last chunk
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int b, mp, split;
        tc_decode(p, item, b, mp, split);
        const int kc_begin = split * p.chunks_per_split, kc_end = min(nk, kc_begin + p.chunks_per_split);
        const int mt0 = mp * 2;
        const int n_mt = (mt0 + 1 < p.n_mtiles) ? 2 : 1;
        mbar_wait(tmem_empty, acc_phase ^ 1);                    // epilogue has drained the accumulators
        tc_fence_after();
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(stage_base + stage * TC_STAGE_BYTES);
          const int nks = (kc == nk - 1 && k_tail <= 16) ? 1 : 2;
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t koff = ks * 32;                       // 16 bf16 = 32 bytes along K inside the atom
            const uint64_t bh = make_smem_desc_sw64(st + 4 * TC_A_TILE_BYTES + koff);
            const uint64_t bl = make_smem_desc_sw64(st + 4 * TC_A_TILE_BYTES + TC_B_TILE_BYTES + koff);
            for (int mt = 0; mt < n_mt; ++mt) {
              const uint64_t ah = make_smem_desc_sw64(st + mt * TC_A_TILE_BYTES + koff);
              const uint64_t al = make_smem_desc_sw64(st + (2 + mt) * TC_A_TILE_BYTES + koff);
              const uint32_t d = tmem_base + (uint32_t)mt * 256;
              tc_mma_bf16(d, ah, bh, idesc, (kc != kc_begin) || (ks != 0));
              tc_mma_bf16(d, al, bh, idesc, 1);
              tc_mma_bf16(d, ah, bl, idesc, 1);
            }
          }
          tc_commit(&empty_bar[stage]);                          // smem slot reusable once these MMAs retire
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tmem_full);                                    // accumulators complete
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 16 warps = (accumulator, column half) x 4 lane quadrants; TMEM lane == tile row =====
    const int grp = ((warp - 4) >> 2) & 1;                       // which accumulator / m-tile of the pair
    const int half = (warp - 4) >> 3;                            // which half of the accumulator's columns
    const int et = (threadIdx.x - 128) & 127;                    // row inside the 128-row tile
    const int eall = threadIdx.x - 128;                          // 0..511
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int b, mp, split;
      tc_decode(p, item, b, mp, split);
      const int mt0 = mp * 2;
      const int n_mt = (mt0 + 1 < p.n_mtiles) ? 2 : 1;
      if (EPI == EPI_BIAS_ACT_T_SPLIT) {
        asm volatile("bar.sync 1, 512;" ::: "memory");           // previous item's readers are done
        for (int c = eall; c < p.H; c += 512) bias_s[c] = p.bias ? p.bias[(int64_t)b * p.bias_stride + c] : 0.f;
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
      mbar_wait(tmem_full, acc_phase);
      tc_fence_after();
      if (grp < n_mt) {
        const int row = (mt0 + grp) * 128 + et;
        const bool valid = row < p.M_valid;
        const int c_split = ((p.H + 63) >> 6) << 5;              // first half: [0, c_split), second: [c_split, H)
        for (int c0 = half ? c_split : 0; c0 < (half ? p.H : min(c_split, p.H)); c0 += 32) {
          float v[32];
          tc_ld32(tmem_base + lane_base + (uint32_t)(grp * 256 + c0), v);
          if (EPI == EPI_BIAS_ACT_T_SPLIT) {
            // even lanes own even columns, odd lanes odd columns; the partner lane's value arrives by
            // shuffle so that (row, row+1) leave as one 32-bit bf16x2 word: half the store instructions,
            // 64 B contiguous per half-warp.  All lanes take part in the shuffles (rows >= M_valid too).
            const int odd = lane & 1;
            const bool v_even = (row & ~1) < p.M_valid, v_odd = (row | 1) < p.M_valid;   // rows >= M_valid store zeros
            const int64_t w0 = ((((int64_t)b * p.out_tiles + (mt0 + grp)) * p.H + c0 + odd) * 128 + (et & ~1)) >> 1;
            uint32_t* ohi = reinterpret_cast<uint32_t*>(p.out_hi) + w0;
            uint32_t* olo = reinterpret_cast<uint32_t*>(p.out_lo) + w0;
            const bool full = (c0 + 32 <= p.H);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float a_e = act_apply_t<ACT>(v[j] + bias_s[c0 + j]);
              const float a_o = act_apply_t<ACT>(v[j + 1] + bias_s[c0 + j + 1]);
              const float recv = __shfl_xor_sync(0xffffffffu, odd ? a_e : a_o, 1);
              const float x0 = v_even ? (odd ? recv : a_e) : 0.f;       // row & ~1
              const float x1 = v_odd ? (odd ? a_o : recv) : 0.f;        // (row & ~1) + 1
              if (full || c0 + j + odd < p.H) {
                const __nv_bfloat162 hp = __floats2bfloat162_rn(x0, x1);           // one cvt.rn.bf16x2.f32
                const uint32_t hw = *reinterpret_cast<const uint32_t*>(&hp);
                const float r0 = x0 - __uint_as_float(hw << 16);
                const float r1 = x1 - __uint_as_float(hw & 0xffff0000u);
                const __nv_bfloat162 lp = __floats2bfloat162_rn(r0, r1);
                ohi[j * 64] = hw;                                                  // column c0+j+odd is j*64 words on
                olo[j * 64] = *reinterpret_cast<const uint32_t*>(&lp);
              }
            }
          } else {
            if (valid) {
              float* o = p.out + (int64_t)split * p.split_stride + (int64_t)b * p.out_stride + (int64_t)row * p.out_ld + c0;
              const int nvalid = p.n_cols_total > 0 ? min(p.N_valid, p.n_cols_total - b * p.H) : p.N_valid;
              if (p.vec_store && c0 + 32 <= nvalid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (c0 + j < nvalid) o[j] = v[j];
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);
      acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256 x H tile per SM pair.  Each CTA owns 128 rows of A and HALF
// of the B stage (the tensor core reads the peer's half through the pair), so the L2->SM traffic per
// flop is the same as the two-accumulator kernel above, but an accumulator is only 128 x H per SM:
// TMEM holds TWO of them and the epilogue of item i overlaps the MMAs of item i+1.
// Roles per CTA: warp 8 TMA producer (own A rows, own half of B; completion is signalled on the
// leader CTA's barrier), warp 9 MMA issuer (leader CTA only), warp 10 TMEM allocator, warps 0-7 epilogue.
// ------------------------------------------------------------------------------------------
constexpr int TP_STAGES = 6;
constexpr int TP_STAGE_BYTES = 2 * TC_A_TILE_BYTES + 2 * 8192;     // A hi/lo (128 rows) + half of B hi/lo (<=128 rows)
constexpr int TP_SMEM_BYTES = TP_STAGES * TP_STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers*/ + 2048 /*bias x2*/;
constexpr int TP_THREADS = 384;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA loads: executed by both CTAs, the transaction bytes land on the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {       // arrive on the same barrier in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on the leader CTA's copy of `bar`
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

// same, without release semantics: the caller has already ordered its TMEM reads with tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync and publishes no memory through this barrier (a cluster-scope release would
// wait for every global store the warp has in flight)
__device__ __forceinline__ void mbar_arrive_leader_relaxed(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_decode_pair(const TcGemmParams& p, int item, int& b, int& mp, int& split, int& bt) {
  if (p.n_btiles > 0) {          // [split][chain][B tile]: the B tiles sharing one A k-range run side by side, and the
    bt = item % p.n_btiles;      // k-range of the shared B operand (a few tens of MB) stays L2-resident for all chains
    int r = item / p.n_btiles;
    b = r % p.n_batch;
    split = r / p.n_batch;
    mp = 0;
    return;
  }
  bt = -1;
  tc_decode(p, item, b, mp, split);
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int EPI, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TP_THREADS, 1)
tc_gemm_pair_bf16x3(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint8_t* stage_base = smem;
  uint64_t* bars = (uint64_t*)(smem + TP_STAGES * TP_STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [TP_STAGES]   (used in the leader CTA)
  uint64_t* empty_bar = bars + TP_STAGES;        // [TP_STAGES]   (one per CTA: own smem slot is free)
  uint64_t* tmem_full = bars + 2 * TP_STAGES;    // [2]           (one per CTA)
  uint64_t* tmem_empty = bars + 2 * TP_STAGES + 2;   // [2]       (leader CTA: both epilogues have drained)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * TP_STAGES + 4);
  float* bias_s = (float*)(smem + TP_STAGES * TP_STAGE_BYTES + 1024);   // [2][256]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int nk = (p.K + TC_BK - 1) / TC_BK;
  const int half_rows = p.H >> 1;                // rows of B this CTA stages
  const uint32_t cta_bytes = 2 * TC_A_TILE_BYTES + 2 * (uint32_t)half_rows * TC_BK * 2;

  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < TP_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
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
  cluster_sync_all();                            // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // warp roles: 0-7 epilogue, 8 TMA producer, 9 MMA issuer, 10 TMEM allocator (the SMSP arbiter favours the
  // highest warp id: the single-thread issuers must not queue behind the epilogue warps)
  if (warp == 8) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        int b, mp, split, bt;
        tc_decode_pair(p, item, b, mp, split, bt);
        if (p.sym_skip && b < mp) continue;
        const int kc_begin = split * p.chunks_per_split, kc_end = min(nk, kc_begin + p.chunks_per_split);
        const int arow = (p.a_blocked ? 0 : p.a_row0) + (mp * 2 + (int)rank) * 128;
        const int brow = (bt >= 0 ? p.b_row0 + bt * p.H : b * p.H) + (int)rank * half_rows;
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stage_base + stage * TP_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * cta_bytes);     // bytes of BOTH CTAs
          const int k0 = kc * TC_BK;
          if (p.a_blocked) {
            tma_load_3d_pair(st, &tmA_hi, &full_bar[stage], k0 & 127, arow, b * p.k_tiles + (k0 >> 7));
            tma_load_3d_pair(st + TC_A_TILE_BYTES, &tmA_lo, &full_bar[stage], k0 & 127, arow, b * p.k_tiles + (k0 >> 7));
          } else {
            tma_load_2d_pair(st, &tmA_hi, &full_bar[stage], k0, arow);
            tma_load_2d_pair(st + TC_A_TILE_BYTES, &tmA_lo, &full_bar[stage], k0, arow);
          }
          tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES, &tmB_hi, &full_bar[stage], k0, brow);
          tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES + 8192, &tmB_lo, &full_bar[stage], k0, brow);
          if (++stage == TP_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0 && lane == 0) {
      // D=F32, A=B=BF16, K-major, N = H, M = 256 across the pair
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.H >> 3) << 17) | ((256u >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const int k_tail = p.K - (nk - 1) * TC_BK;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        int b, mp, split, bt;
        tc_decode_pair(p, item, b, mp, split, bt);
        if (p.sym_skip && b < mp) continue;
        const int kc_begin = split * p.chunks_per_split, kc_end = min(nk, kc_begin + p.chunks_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);               // both epilogues drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * 256;
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(stage_base + stage * TP_STAGE_BYTES);
          const int nks = (kc == nk - 1 && k_tail <= 16) ? 1 : 2;
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t koff = ks * 32;
            const uint64_t ah = make_smem_desc_sw64(st + koff);
            const uint64_t al = make_smem_desc_sw64(st + TC_A_TILE_BYTES + koff);
            const uint64_t bh = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + koff);
            const uint64_t bl = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + 8192 + koff);
            tc_mma_bf16_pair(d, ah, bh, idesc, (kc != kc_begin) || (ks != 0));
            tc_mma_bf16_pair(d, al, bh, idesc, 1);
            tc_mma_bf16_pair(d, ah, bl, idesc, 1);
          }
          tc_commit_pair(&empty_bar[stage]);                      // frees the slot in BOTH CTAs
          if (++stage == TP_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(&tmem_full[acc]);                          // accumulator ready in BOTH CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 8) {
    // ===== epilogue (both CTAs): 8 warps = 4 lane quadrants x 2 column halves of this CTA's 128 x H tile =====
    const int half = warp >> 2;
    const int et = threadIdx.x & 127;
    const int eall = threadIdx.x;                                 // 0..255
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      int b, mp, split, bt;
      tc_decode_pair(p, item, b, mp, split, bt);
      if (p.sym_skip && b < mp) continue;
      const int mt = mp * 2 + (int)rank;
      float* bs = bias_s + acc * 256;
      if (EPI == EPI_BIAS_ACT_T_SPLIT) {
        // bias_s[acc] was last read two items ago; the tmem_empty/tmem_full hand-shake orders those reads
        for (int c = eall; c < p.H; c += 256) bs[c] = p.bias ? p.bias[(int64_t)b * p.bias_stride + c] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      {
        const int row = mt * 128 + et;
        const bool valid = row < p.M_valid;
        const int c_split = ((p.H + 63) >> 6) << 5;
        for (int c0 = half ? c_split : 0; c0 < (half ? p.H : min(c_split, p.H)); c0 += 32) {
          float v[32];
          tc_ld32(tmem_base + lane_base + (uint32_t)(acc * 256 + c0), v);
          if (EPI == EPI_BIAS_ACT_T_SPLIT) {
            const int odd = lane & 1;
            const bool v_even = (row & ~1) < p.M_valid, v_odd = (row | 1) < p.M_valid;
            const int64_t w0 = ((((int64_t)b * p.out_tiles + mt) * p.H + c0 + odd) * 128 + (et & ~1)) >> 1;
            uint32_t* ohi = reinterpret_cast<uint32_t*>(p.out_hi) + w0;
            uint32_t* olo = reinterpret_cast<uint32_t*>(p.out_lo) + w0;
            const bool full = (c0 + 32 <= p.H);
            const bool in_range = mt < p.out_tiles;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float a_e = act_apply_t<ACT>(v[j] + bs[c0 + j]);
              const float a_o = act_apply_t<ACT>(v[j + 1] + bs[c0 + j + 1]);
              const float recv = __shfl_xor_sync(0xffffffffu, odd ? a_e : a_o, 1);
              const float x0 = v_even ? (odd ? recv : a_e) : 0.f;
              const float x1 = v_odd ? (odd ? a_o : recv) : 0.f;
              if (in_range && (full || c0 + j + odd < p.H)) {
                const __nv_bfloat162 hp = __floats2bfloat162_rn(x0, x1);
                const uint32_t hw = *reinterpret_cast<const uint32_t*>(&hp);
                const float r0 = x0 - __uint_as_float(hw << 16);
                const float r1 = x1 - __uint_as_float(hw & 0xffff0000u);
                const __nv_bfloat162 lp = __floats2bfloat162_rn(r0, r1);
                ohi[j * 64] = hw;
                olo[j * 64] = *reinterpret_cast<const uint32_t*>(&lp);
              }
            }
          } else {
            if (p.transpose_out) {
              // out[(col0 + col) * out_ld + row]: consecutive lanes hold consecutive rows -> 128 B per warp store
              if (valid) {
                const int col0 = p.b_row0 + bt * p.H + c0;
                float* o = p.out + (int64_t)split * p.split_stride + (int64_t)b * p.out_stride + (int64_t)col0 * p.out_ld + row;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.n_cols_total && c0 + j < p.H) o[(int64_t)j * p.out_ld] = v[j];
              }
            } else if (valid) {
              float* o = p.out + (int64_t)split * p.split_stride + (int64_t)b * p.out_stride + (int64_t)row * p.out_ld + c0;
              const int nvalid = p.n_cols_total > 0 ? min(p.N_valid, p.n_cols_total - b * p.H) : p.N_valid;
              if (p.vec_store && c0 + 32 <= nvalid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (c0 + j < nvalid) o[j] = v[j];
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peer may still be reading this CTA's shared memory
  if (warp == 10) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// Hidden-major gradient GEMM with TWO accumulators per CTA pair: D[h, f] = sum_r A[h, r] B[f, r] for two
// adjacent feature tiles (bt, bt+1) of [X^T;1] at once.  The pair kernel above re-reads the chain's dZ1^T
// (the A operand, 61 MB per chain) once per feature tile — 4 times at D+1 = 785 — and ncu shows it bound by
// L2->SM bandwidth (64 GB per launch at ~5900 B/clk, tensor pipe 85.7 %).  Here every A stage feeds both
// tiles (A traffic halves, 64 -> 47 GB); the two 256-column accumulators fill TMEM, so the epilogue is not
// overlapped, which costs ~2 % at one epilogue per 8192-row split-K segment.
// Roles: warps 0-7 epilogue (lane quadrant x accumulator), 8 TMA producer, 9 MMA issuer, 10 TMEM allocator.
// ------------------------------------------------------------------------------------------
constexpr int TD_STAGES = 4;
constexpr int TD_STAGE_BYTES = 2 * TC_A_TILE_BYTES + 4 * 8192;     // A hi/lo (128 rows) + half of B hi/lo for two tiles
constexpr int TD_SMEM_BYTES = TD_STAGES * TD_STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers*/;

__device__ __forceinline__ void td_decode(const TcGemmParams& p, int item, int& b, int& split, int& bt0, int& n_t) {
  const int n_btp = (p.n_btiles + 1) >> 1;       // [split][chain][tile pair]: the shared operand's k-range stays in L2
  const int btp = item % n_btp;
  const int r = item / n_btp;
  b = r % p.n_batch;
  split = r / p.n_batch;
  bt0 = 2 * btp;
  n_t = (bt0 + 1 < p.n_btiles) ? 2 : 1;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TP_THREADS, 1)
tc_gemm_pair_dual_bf16x3(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                         const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                         const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint64_t* bars = (uint64_t*)(smem + TD_STAGES * TD_STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [TD_STAGES]   (leader CTA)
  uint64_t* empty_bar = bars + TD_STAGES;        // [TD_STAGES]   (one per CTA)
  uint64_t* tmem_full = bars + 2 * TD_STAGES;    // (one per CTA)
  uint64_t* tmem_empty = bars + 2 * TD_STAGES + 1;   // (leader CTA: both epilogues have drained)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * TD_STAGES + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int nk = (p.K + TC_BK - 1) / TC_BK;
  const int half_rows = p.H >> 1;                // rows of each B tile this CTA stages
  const uint32_t b_bytes = (uint32_t)half_rows * TC_BK * 2;

  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < TD_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 16);                   // 8 epilogue warps x 2 CTAs
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

  if (warp == 8) {
    // ===== TMA producer (both CTAs): own 128 hidden units of dZ1^T[b], own half of each feature tile =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        int b, split, bt0, n_t;
        td_decode(p, item, b, split, bt0, n_t);
        const int kc_begin = split * p.chunks_per_split, kc_end = min(nk, kc_begin + p.chunks_per_split);
        const int arow = p.chain_pairs ? 0 : (int)rank * 128;
        const int ab = p.chain_pairs ? 2 * b + (int)rank : b;     // (an odd batch's last partner reads another chain's
        const uint32_t cta_bytes = 2 * TC_A_TILE_BYTES + 2 * (uint32_t)n_t * b_bytes;   // block or zeros: never stored)
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stage_base + stage * TD_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * cta_bytes);
          const int k0 = kc * TC_BK;
          tma_load_3d_pair(st, &tmA_hi, &full_bar[stage], k0 & 127, arow, ab * p.k_tiles + (k0 >> 7));
          tma_load_3d_pair(st + TC_A_TILE_BYTES, &tmA_lo, &full_bar[stage], k0 & 127, arow, ab * p.k_tiles + (k0 >> 7));
          for (int j = 0; j < n_t; ++j) {
            const int brow = p.b_row0 + (bt0 + j) * p.H + (int)rank * half_rows;
            tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES + (2 * j) * 8192, &tmB_hi, &full_bar[stage], k0, brow);
            tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES + (2 * j + 1) * 8192, &tmB_lo, &full_bar[stage], k0, brow);
          }
          if (++stage == TD_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.H >> 3) << 17) | ((256u >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      const int k_tail = p.K - (nk - 1) * TC_BK;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        int b, split, bt0, n_t;
        td_decode(p, item, b, split, bt0, n_t);
        const int kc_begin = split * p.chunks_per_split, kc_end = min(nk, kc_begin + p.chunks_per_split);
        mbar_wait(tmem_empty, acc_phase ^ 1);                     // both epilogues drained the accumulators
        tc_fence_after();
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(stage_base + stage * TD_STAGE_BYTES);
          const int nks = (kc == nk - 1 && k_tail <= 16) ? 1 : 2;
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t koff = ks * 32;
            const uint64_t ah = make_smem_desc_sw64(st + koff);
            const uint64_t al = make_smem_desc_sw64(st + TC_A_TILE_BYTES + koff);
            const uint32_t accum = (kc != kc_begin) || (ks != 0);
            for (int j = 0; j < n_t; ++j) {
              const uint64_t bh = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + (2 * j) * 8192 + koff);
              const uint64_t bl = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + (2 * j + 1) * 8192 + koff);
              const uint32_t d = tmem_base + (uint32_t)j * 256;
              tc_mma_bf16_pair(d, ah, bh, idesc, accum);
              tc_mma_bf16_pair(d, al, bh, idesc, 1);
              tc_mma_bf16_pair(d, ah, bl, idesc, 1);
            }
          }
          tc_commit_pair(&empty_bar[stage]);
          if (++stage == TD_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(tmem_full);
        acc_phase ^= 1;
      }
    }
  } else if (warp < 8) {
    // ===== epilogue (both CTAs): warp = (accumulator, lane quadrant); transposed partial-sum stores =====
    const int j = warp >> 2;                                      // which accumulator / feature tile of the pair
    const int et = threadIdx.x & 127;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc_phase = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      int b, split, bt0, n_t;
      td_decode(p, item, b, split, bt0, n_t);
      mbar_wait(tmem_full, acc_phase);
      tc_fence_after();
      if (j < n_t) {
        const int row = p.chain_pairs ? et : (int)rank * 128 + et;          // hidden unit
        const int oc = p.chain_pairs ? 2 * b + (int)rank : b;               // chain whose gradient this CTA holds
        const int colbase = p.b_row0 + (bt0 + j) * p.H;
        float* ob = p.out + (int64_t)split * p.split_stride + (int64_t)oc * p.out_stride + row;
        for (int c0 = 0; c0 < p.H; c0 += 32) {
          float v[32];
          tc_ld32(tmem_base + lane_base + (uint32_t)(j * 256 + c0), v);
          if (row < p.M_valid && (!p.chain_pairs || oc < p.n_chains)) {
            float* o = ob + (int64_t)(colbase + c0) * p.out_ld;
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (colbase + c0 + q < p.n_cols_total && c0 + q < p.H) o[(int64_t)q * p.out_ld] = v[q];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tmem_empty);
      acc_phase ^= 1;
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

// tc_path.cu — tensor-core path for the MNIST-width case (BASELINE config C3/C4/C5 shapes):
// two-layer Dense MLP whose first layer is a real dense contraction (D x H with H <= 256).
//
// What runs where, per chain batch (Bc chains), per log-posterior evaluation:
//   k_split_transpose  theta -> W1^T split into bf16 hi/lo, K-major [Bc*H, D]       (SIMT, tiny)
//   tc_gemm<G1>   Z1 = X W1 (+b1, act) -> A1^T split bf16 hi/lo [Bc*H, Npad]         (tcgen05 + TMA)
//   k_layer2      per row: z2 = a1 W2 + b2, softmax-CE / MSE, dZ2,
//                 dZ1 = (dZ2 W2^T) act'(a1) -> dZ1^T, dZ2^T split bf16 hi/lo; loss, db2  (SIMT, HBM-bound)
//   tc_gemm<G3>   dW2 = A1^T dZ2 -> grad[:, w2_off : w2_off + H*C]                   (tcgen05 + TMA, HBM-bound)
//   tc_gemm<G2>   [dW1; db1] = [X^T; 1] dZ1 -> grad[:, 0 : D*H + H]                  (tcgen05 + TMA)
// fp32-grade products on bf16 tensor cores: every operand is split x = hi + lo (bf16 each) and the
// MMA issues hi*hi + lo*hi + hi*lo into one fp32 TMEM accumulator (relative error ~2^-16 per
// product, far inside the 1e-4 parity budget; single-pass BF16/TF32 would not be).
//
// The GEMM kernel is ONE persistent, warp-specialised kernel used for both contractions:
//   D[M, H] = A[M, K] * B[H, K]^T,  A shared by all chains, B per chain.
//   G1: A = X [N, D],  B = W1^T[b] [H, D],     M = rows,     K = D   (784)
//   G2: A = [X^T;1] [D+1, N], B = dZ1^T[b] [H, N], M = D+1,  K = rows (60000)
// CTA tile = 2 x (128 x H) accumulators in TMEM (2*256 = 512 columns) sharing each B stage, K
// streamed in 32-element (64 B, SWIZZLE_64B) chunks through a 3-stage TMA/mbarrier ring:
// 64 KB/stage for 2*3*128*256*32*2 flop => ~42 B/clk/SM of L2->SM traffic at full MMA rate.
// Roles: warp 0 TMA producer, warp 1 MMA issuer (one elected thread), warp 2 TMEM allocator,
// warps 4-19 epilogue: (accumulator, column half) x 4 lane quadrants (TMEM lane == output row).
#include "common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace pyb {

// ------------------------------------------------------------------------------------------
// PTX helpers (sm_100a)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// operand tile load: plain 2-D [rows, K] tensor, or row-tile-blocked 3-D [chain*k_tiles][rows][128] tensor
// (K = data rows, blocked in tiles of 128 so that one (chain, tile) block is contiguous in HBM)
__device__ __forceinline__ void tma_load_operand(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int blocked,
                                                 int k0, int row0, int b, int k_tiles) {
  if (blocked) tma_load_3d(smem_dst, map, bar, k0 & 127, row0, b * k_tiles + (k0 >> 7));
  else tma_load_2d(smem_dst, map, bar, k0, row0);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major operand tile with 64-byte rows, SWIZZLE_64B: 8-row atoms of 512 B (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // start address, bits [0,14)
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)4 << 61;                            // layout type SWIZZLE_64B
  return d;
}

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
        const int arow = (int)rank * 128;
        const uint32_t cta_bytes = 2 * TC_A_TILE_BYTES + 2 * (uint32_t)n_t * b_bytes;
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stage_base + stage * TD_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * cta_bytes);
          const int k0 = kc * TC_BK;
          tma_load_3d_pair(st, &tmA_hi, &full_bar[stage], k0 & 127, arow, b * p.k_tiles + (k0 >> 7));
          tma_load_3d_pair(st + TC_A_TILE_BYTES, &tmA_lo, &full_bar[stage], k0 & 127, arow, b * p.k_tiles + (k0 >> 7));
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
        const int row = (int)rank * 128 + et;                     // hidden unit
        const int colbase = p.b_row0 + (bt0 + j) * p.H;
        float* ob = p.out + (int64_t)split * p.split_stride + (int64_t)b * p.out_stride + row;
        for (int c0 = 0; c0 < p.H; c0 += 32) {
          float v[32];
          tc_ld32(tmem_base + lane_base + (uint32_t)(j * 256 + c0), v);
          if (row < p.M_valid) {
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

// ------------------------------------------------------------------------------------------
// SIMT helpers around the GEMMs
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// src [R, C] fp32 (row stride lds) -> hi/lo [R, C] bf16 (row stride ldd), same orientation
__global__ void k_split_rows(const float* src, int64_t R, int C, int64_t lds, __nv_bfloat16* hi, __nv_bfloat16* lo,
                             int64_t ldd) {
  const int64_t total = R * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / C;
    int c = (int)(i - r * C);
    __nv_bfloat16 h, l;
    split_bf16(src[r * lds + c], h, l);
    hi[r * ldd + c] = h;
    lo[r * ldd + c] = l;
  }
}

// src [R, C] fp32 (batch stride sb) -> transposed hi/lo [C(+ones row), Rpad] bf16 per batch element.
// grid (ceil(C/32), ceil(R/32), batch), block (32, 8)
// Row order inside a 128-row block as the fused G1+layer-2 epilogue stores it: the 4 rows {g, g+8, g+16, g+24} of a
// 32-row TMEM quadrant that one thread owns become 4 CONSECUTIVE elements (one 8-byte store).  Any operand that is
// contracted against those arrays over the data rows ([X^T;1] in the dW1 GEMM) must use the same order.
__host__ __device__ __forceinline__ int fused_row_pos(int r) { return (r & ~31) | ((r & 7) << 2) | ((r >> 3) & 3); }

__global__ void k_split_transpose(const float* src, int64_t sb, int R, int C, int64_t lds, __nv_bfloat16* hi,
                                  __nv_bfloat16* lo, int64_t db, int64_t ldd, int perm = 0) {
  __shared__ float t[32][33];
  const float* s = src + (int64_t)blockIdx.z * sb;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < R && c < C) ? s[(int64_t)r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) {
      __nv_bfloat16 h, l;
      split_bf16(t[threadIdx.x][i], h, l);
      int64_t o = (int64_t)blockIdx.z * db + (int64_t)c * ldd + (perm ? fused_row_pos(r) : r);
      hi[o] = h;
      lo[o] = l;
    }
  }
}

__global__ void k_fill_ones_perm(__nv_bfloat16* p, int n) {      // p[fused_row_pos(r)] = 1 for r < n
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[fused_row_pos(i)] = __float2bfloat16_rn(1.0f);
}
__global__ void k_fill_bf16(__nv_bfloat16* p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = __float2bfloat16_rn(v);
}

// Layer 2 forward/backward.  One thread owns TWO adjacent data rows (packed bf16x2 loads/stores are
// then 128 B per warp); a block walks 256-row tiles of one chain.  a1 = hi + lo is read ONCE: for
// relu the derivative mask is kept as 2 x 256 bits in registers; other activations re-read a1.
constexpr int L2_CMAX = 16;
constexpr int L2_ROWS = 256;    // rows per tile (128 threads x 2)
struct Layer2Params {
  const __nv_bfloat16* a_hi; const __nv_bfloat16* a_lo;     // A1^T [Bc*H][ld]
  __nv_bfloat16* zt_hi; __nv_bfloat16* zt_lo;               // dZ1^T [Bc*H][ld]
  __nv_bfloat16* z2_hi; __nv_bfloat16* z2_lo;               // dZ2^T [Bc*16][ld]
  int k_tiles;                 // 128-row tiles per chain; arrays are blocked [chain][tile][unit][128]
  const float* theta; int64_t P; int64_t w2_off, b2_off;
  int H, C, N, act1, out_act, loss_kind;
  const int32_t* y_i; const float* y_f;
  float scale;                 // n_train (or 1): dZ = scale * d(mean loss)/dz
  double* loss_partial;        // [Bc][n_groups]
  float* b2_partial;           // [Bc][n_groups][16]
  int n_groups, n_tiles;
};
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
template <int CP, typename LT>
__device__ __forceinline__ void l2_loss_dz(const Layer2Params& p, int r, bool valid, const float* z, float* dz,
                                           LT& loss_acc, float invN) {
  const int C = p.C;
#pragma unroll
  for (int c = 0; c < CP; ++c) dz[c] = 0.f;
  if (!valid) return;
  if (p.loss_kind == PYB_LOSS_SPARSE_CE) {
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CP; ++c) if (c < C) mx = fmaxf(mx, z[c]);
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) if (c < C) se += expf(z[c] - mx);
    const int yi = p.y_i[r];
    float zy = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) if (c == yi) zy = z[c];
    loss_acc += (LT)(logf(se) - (zy - mx));
    const float inv = 1.0f / se;
#pragma unroll
    for (int c = 0; c < CP; ++c)
      if (c < C) dz[c] = (expf(z[c] - mx) * inv - (c == yi ? 1.f : 0.f)) * invN;
  } else {
    float acc = 0.f;
    const float sc = 2.0f * invN / (float)C;
#pragma unroll
    for (int c = 0; c < CP; ++c)
      if (c < C) {
        float a = act_apply(z[c], p.out_act);
        float df = a - p.y_f[(int64_t)r * C + c];
        acc += df * df;
        dz[c] = sc * df * act_grad_from_output(a, p.out_act);
      }
    loss_acc += (LT)(acc / (float)C);
  }
}
// CP = class count padded to a multiple of 4 (register tile of the per-row logits)
template <int CP>
__global__ void __launch_bounds__(128, 4) k_layer2(Layer2Params p) {
  __shared__ __align__(16) float W2s[256 * CP];
  __shared__ float b2s[CP];
  __shared__ uint32_t mask_s[2][8][128];
  __shared__ float redf[4][CP];
  __shared__ double scratch[32];
  const int t = threadIdx.x, b = blockIdx.y;
  const int H = p.H, C = p.C;
  const float* th = p.theta + (int64_t)b * p.P;
  for (int i = t; i < H * CP; i += 128) {
    int h = i / CP, c = i % CP;
    W2s[i] = (c < C) ? th[p.w2_off + (int64_t)h * C + c] : 0.f;
  }
  if (t < CP) b2s[t] = (t < C) ? th[p.b2_off + t] : 0.f;
  __syncthreads();
  // blocked layout, in packed bf16x2 words: ((chain*k_tiles + tile128)*rows_per_block + unit)*64 + pair
  const uint32_t* ah = reinterpret_cast<const uint32_t*>(p.a_hi);
  const uint32_t* al = reinterpret_cast<const uint32_t*>(p.a_lo);
  uint32_t* zh = reinterpret_cast<uint32_t*>(p.zt_hi);
  uint32_t* zl = reinterpret_cast<uint32_t*>(p.zt_lo);
  uint32_t* z2h = reinterpret_cast<uint32_t*>(p.z2_hi);
  uint32_t* z2l = reinterpret_cast<uint32_t*>(p.z2_lo);
  const int64_t ld2 = 64;                  // pitch between hidden units inside a block, in packed pairs
  float accb[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) accb[c] = 0.f;
  double loss_acc = 0.0;
  const float invN = p.scale / (float)p.N;
  const bool relu = p.act1 == PYB_ACT_RELU;
  const int nhb = (H + 31) >> 5;
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int r0 = tile * L2_ROWS + 2 * t;
    const int64_t tile128 = (int64_t)b * p.k_tiles + tile * 2 + (t >> 6);   // this thread's 128-row block
    const int64_t col = tile128 * H * 64 + (t & 63);                        // word offset of (unit 0, this row pair)
    const int64_t col2 = tile128 * L2_CMAX * 64 + (t & 63);                 // same for the dZ2^T blocks
    const bool v0 = r0 < p.N, v1 = r0 + 1 < p.N;
    // ---- phase A: z2 = a1 W2 + b2 for both rows; relu mask bits go to shared memory
    float z0[CP], z1[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) { z0[c] = b2s[c]; z1[c] = b2s[c]; }
    for (int hb = 0; hb < nhb; ++hb) {
      uint32_t m0 = 0u, m1 = 0u;
      const uint32_t* ahp = ah + (int64_t)(hb * 32) * ld2 + col;
      const uint32_t* alp = al + (int64_t)(hb * 32) * ld2 + col;
      const int jn = min(32, H - hb * 32);
#pragma unroll 4
      for (int j = 0; j < jn; ++j) {
        const float2 hi = unpack_bf16x2(ahp[(int64_t)j * ld2]);
        const float2 lo = unpack_bf16x2(alp[(int64_t)j * ld2]);
        const float a0 = hi.x + lo.x, a1v = hi.y + lo.y;
        m0 |= (a0 > 0.f ? 1u : 0u) << j;
        m1 |= (a1v > 0.f ? 1u : 0u) << j;
        const float4* w = reinterpret_cast<const float4*>(&W2s[(hb * 32 + j) * CP]);
#pragma unroll
        for (int q4 = 0; q4 < CP / 4; ++q4) {
          const float4 wv = w[q4];
          z0[q4 * 4 + 0] = fmaf(a0, wv.x, z0[q4 * 4 + 0]); z1[q4 * 4 + 0] = fmaf(a1v, wv.x, z1[q4 * 4 + 0]);
          z0[q4 * 4 + 1] = fmaf(a0, wv.y, z0[q4 * 4 + 1]); z1[q4 * 4 + 1] = fmaf(a1v, wv.y, z1[q4 * 4 + 1]);
          z0[q4 * 4 + 2] = fmaf(a0, wv.z, z0[q4 * 4 + 2]); z1[q4 * 4 + 2] = fmaf(a1v, wv.z, z1[q4 * 4 + 2]);
          z0[q4 * 4 + 3] = fmaf(a0, wv.w, z0[q4 * 4 + 3]); z1[q4 * 4 + 3] = fmaf(a1v, wv.w, z1[q4 * 4 + 3]);
        }
      }
      mask_s[0][hb][t] = m0;
      mask_s[1][hb][t] = m1;
    }
    // ---- loss and dZ2 (scaled); dZ2^T stored split for the dW2 GEMM
    float dz0[CP], dz1[CP];
    l2_loss_dz<CP>(p, r0, v0, z0, dz0, loss_acc, invN);
    l2_loss_dz<CP>(p, r0 + 1, v1, z1, dz1, loss_acc, invN);
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      accb[c] += dz0[c] + dz1[c];
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(dz0[c], h0, l0);
      split_bf16(dz1[c], h1, l1);
      z2h[(int64_t)c * ld2 + col2] = pack_bf16x2(h0, h1);
      z2l[(int64_t)c * ld2 + col2] = pack_bf16x2(l0, l1);
    }
    // ---- phase A2: dZ1 = (dZ2 W2^T) * act'(a1) -> split bf16, transposed packed store
    for (int hb = 0; hb < nhb; ++hb) {
      const uint32_t m0 = mask_s[0][hb][t], m1 = mask_s[1][hb][t];   // own writes: no barrier needed
      uint32_t* zhp = zh + (int64_t)(hb * 32) * ld2 + col;
      uint32_t* zlp = zl + (int64_t)(hb * 32) * ld2 + col;
      const int jn = min(32, H - hb * 32);
#pragma unroll 4
      for (int j = 0; j < jn; ++j) {
        const int h = hb * 32 + j;
        const float4* w = reinterpret_cast<const float4*>(&W2s[h * CP]);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int q4 = 0; q4 < CP / 4; ++q4) {
          const float4 wv = w[q4];
          d0 = fmaf(dz0[q4 * 4 + 0], wv.x, d0); d1 = fmaf(dz1[q4 * 4 + 0], wv.x, d1);
          d0 = fmaf(dz0[q4 * 4 + 1], wv.y, d0); d1 = fmaf(dz1[q4 * 4 + 1], wv.y, d1);
          d0 = fmaf(dz0[q4 * 4 + 2], wv.z, d0); d1 = fmaf(dz1[q4 * 4 + 2], wv.z, d1);
          d0 = fmaf(dz0[q4 * 4 + 3], wv.w, d0); d1 = fmaf(dz1[q4 * 4 + 3], wv.w, d1);
        }
        if (relu) {
          d0 = ((m0 >> j) & 1u) ? d0 : 0.f;
          d1 = ((m1 >> j) & 1u) ? d1 : 0.f;
        } else {
          const float2 hi = unpack_bf16x2(ah[(int64_t)h * ld2 + col]);
          const float2 lo = unpack_bf16x2(al[(int64_t)h * ld2 + col]);
          d0 *= act_grad_from_output(hi.x + lo.x, p.act1);
          d1 *= act_grad_from_output(hi.y + lo.y, p.act1);
        }
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(d0, h0, l0);
        split_bf16(d1, h1, l1);
        zhp[(int64_t)j * ld2] = pack_bf16x2(h0, h1);
        zlp[(int64_t)j * ld2] = pack_bf16x2(l0, l1);
      }
    }
  }
  // ---- per-block partials: db2 (sum of dZ2 over the block's rows) and the loss
  const int lane = t & 31, w = t >> 5;
#pragma unroll
  for (int c = 0; c < CP; ++c) {
    float sum = warp_sum(accb[c]);
    if (lane == 0) redf[w][c] = sum;
  }
  __syncthreads();
  if (t < L2_CMAX)
    p.b2_partial[((int64_t)b * p.n_groups + blockIdx.x) * L2_CMAX + t] =
        (t < CP) ? redf[0][t] + redf[1][t] + redf[2][t] + redf[3][t] : 0.f;
  double tot = block_sum<double>(loss_acc, scratch);
  if (t == 0) p.loss_partial[(int64_t)b * p.n_groups + blockIdx.x] = tot;
}


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
};
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

template <int CP>
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

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int nk = (p.K + TC_BK - 1) / TC_BK;
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
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stage_base + stage * TP_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * cta_bytes);
          const int k0 = kc * TC_BK;
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
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(H >> 3) << 17) | ((256u >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const int k_tail = p.K - (nk - 1) * TC_BK;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * 256;
        for (int kc = 0; kc < nk; ++kc) {
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
            tc_mma_bf16_pair(d, ah, bh, idesc, (kc != 0) || (ks != 0));
            tc_mma_bf16_pair(d, al, bh, idesc, 1);
            tc_mma_bf16_pair(d, ah, bl, idesc, 1);
          }
          tc_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
    };
    if (cluster_id < p.total_items) fetch_consts(cluster_id, 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      int b, mp, split;
      tc_decode(p, item, b, mp, split);
      const int mt = mp * 2 + (int)rank;
      float* bs = bias_s + acc * 256;
      float* b2b = b2_s + acc * 16;
      float* W2b = W2_s + acc * 256 * CP;
      asm volatile("cp.async.wait_all;" ::: "memory");           // this thread's share of the constants has landed
      asm volatile("bar.sync 1, 256;" ::: "memory");             // constants visible; zx_s readers of the last item done
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      // (hidden unit hbase, this thread's 4 rows) inside this (chain, tile) block, in 8-byte units (4 bf16)
      const int64_t blk_w = (((((int64_t)b * p.out_tiles + mt) * H) + hbase) * 128 + pos0) >> 2;
      uint2* pa_hi = reinterpret_cast<uint2*>(p.out_hi) + blk_w;
      uint2* pa_lo = reinterpret_cast<uint2*>(p.out_lo) + blk_w;
      uint2* pz_hi = reinterpret_cast<uint2*>(l2.zt_hi) + blk_w;
      uint2* pz_lo = reinterpret_cast<uint2*>(l2.zt_lo) + blk_w;
      const float4* w4b = reinterpret_cast<const float4*>(W2b) + (((half * Hh) >> 3) * 2 * (CP / 4)) * 4 + t;
      const float* bsb = bs + hbase;
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
          tc_ld_16x256b_x4(tmem_base + lane_addr + col, v);
          tc_ld_16x256b_x4(tmem_base + lane_addr + (16u << 16) + col, v + 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          uint32_t m = 0u;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const float2 bb = *reinterpret_cast<const float2*>(bsb + ch * 32 + 8 * kb);
            float a[4][2];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float* vv = v + (r >> 1) * 16 + kb * 4 + (r & 1) * 2;
              a[r][0] = fmaxf(vv[0] + bb.x, 0.f);
              a[r][1] = fmaxf(vv[1] + bb.y, 0.f);
              m |= (a[r][0] > 0.f ? 1u : 0u) << (kb * 8 + r * 2);
              m |= (a[r][1] > 0.f ? 1u : 0u) << (kb * 8 + r * 2 + 1);
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
              // A1^T: this thread's 4 rows of hidden unit hbase + ... are adjacent in the block's row order
              uint32_t hw0, lw0, hw1, lw1;
              split_pair(a[0][i], a[1][i], hw0, lw0);
              split_pair(a[2][i], a[3][i], hw1, lw1);
              const int w_off = (ch * 32 + 8 * kb + i) * 32;
              __stcs(pa_hi + w_off, make_uint2(hw0, hw1));       // streaming: 18 GB per launch must not evict X / W1^T from L2
              __stcs(pa_lo + w_off, make_uint2(lw0, lw1));
            }
          }
          mask[ch] = m;
        }
      }
      // the accumulator is no longer needed: hand it back to the MMA warp before the rest of the epilogue
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader_relaxed(&tmem_empty[acc]);
      // next item's constants fly while this item's reductions and phase B run; buffer [acc ^ 1] was last read in
      // the previous item, which every epilogue thread has left (they all passed this item's bar.sync 1)
      if (item + n_clusters < p.total_items) fetch_consts(item + n_clusters, acc ^ 1);
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
      float loss_r = 0.f;
      const int row_g = mt * 128 + row_own;
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
              uint32_t hw0, lw0, hw1, lw1;
              split_pair(d[0], d[1], hw0, lw0);
              split_pair(d[2], d[3], hw1, lw1);
              const int w_off = (ch * 32 + 8 * kb + i) * 32;
              __stcs(pz_hi + w_off, make_uint2(hw0, hw1));
              __stcs(pz_lo + w_off, make_uint2(lw0, lw1));
            }
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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

// forward-only layer 2 (posterior predictive): out[b][row][c] = softmax(a1 W2 + b2) or act(.)
struct Layer2FwdParams {
  const __nv_bfloat16* a_hi; const __nv_bfloat16* a_lo; int k_tiles;
  const float* theta; int64_t P; int64_t w2_off, b2_off;
  int H, C, N, out_act, n_tiles;
  float* out;
};
__global__ void __launch_bounds__(128) k_layer2_fwd(Layer2FwdParams p) {
  __shared__ __align__(16) float W2s[256 * L2_CMAX];
  __shared__ float b2s[L2_CMAX];
  const int t = threadIdx.x, b = blockIdx.y, H = p.H, C = p.C;
  const float* th = p.theta + (int64_t)b * p.P;
  for (int i = t; i < H * L2_CMAX; i += 128) {
    int h = i / L2_CMAX, c = i % L2_CMAX;
    W2s[i] = (c < C) ? th[p.w2_off + (int64_t)h * C + c] : 0.f;
  }
  if (t < L2_CMAX) b2s[t] = (t < C) ? th[p.b2_off + t] : 0.f;
  __syncthreads();
  const uint32_t* ah = reinterpret_cast<const uint32_t*>(p.a_hi);
  const uint32_t* al = reinterpret_cast<const uint32_t*>(p.a_lo);
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int r0 = tile * L2_ROWS + 2 * t;
    const int64_t tile128 = (int64_t)b * p.k_tiles + tile * 2 + (t >> 6);
    const int64_t col = tile128 * H * 64 + (t & 63);
    float z0[L2_CMAX], z1[L2_CMAX];
#pragma unroll
    for (int c = 0; c < L2_CMAX; ++c) { z0[c] = b2s[c]; z1[c] = b2s[c]; }
#pragma unroll 4
    for (int h = 0; h < H; ++h) {
      const float2 hi = unpack_bf16x2(ah[(int64_t)h * 64 + col]);
      const float2 lo = unpack_bf16x2(al[(int64_t)h * 64 + col]);
      const float a0 = hi.x + lo.x, a1v = hi.y + lo.y;
      const float4* w = reinterpret_cast<const float4*>(&W2s[h * L2_CMAX]);
#pragma unroll
      for (int q4 = 0; q4 < L2_CMAX / 4; ++q4) {
        const float4 wv = w[q4];
        z0[q4 * 4 + 0] = fmaf(a0, wv.x, z0[q4 * 4 + 0]); z1[q4 * 4 + 0] = fmaf(a1v, wv.x, z1[q4 * 4 + 0]);
        z0[q4 * 4 + 1] = fmaf(a0, wv.y, z0[q4 * 4 + 1]); z1[q4 * 4 + 1] = fmaf(a1v, wv.y, z1[q4 * 4 + 1]);
        z0[q4 * 4 + 2] = fmaf(a0, wv.z, z0[q4 * 4 + 2]); z1[q4 * 4 + 2] = fmaf(a1v, wv.z, z1[q4 * 4 + 2]);
        z0[q4 * 4 + 3] = fmaf(a0, wv.w, z0[q4 * 4 + 3]); z1[q4 * 4 + 3] = fmaf(a1v, wv.w, z1[q4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = r0 + rr;
      if (r >= p.N) continue;
      float* z = rr ? z1 : z0;
      float* o = p.out + ((int64_t)b * p.N + r) * C;
      if (p.out_act == PYB_ACT_SOFTMAX) {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < L2_CMAX; ++c) if (c < C) mx = fmaxf(mx, z[c]);
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < L2_CMAX; ++c) if (c < C) se += expf(z[c] - mx);
        const float inv = 1.0f / se;
#pragma unroll
        for (int c = 0; c < L2_CMAX; ++c) if (c < C) o[c] = expf(z[c] - mx) * inv;
      } else {
#pragma unroll
        for (int c = 0; c < L2_CMAX; ++c) if (c < C) o[c] = act_apply(z[c], p.out_act);
      }
    }
  }
}

// grad[b][b2_off + c] = sum_g b2_partial[b][g][c]; loss[b] = sum_g loss_partial / N.  Fixed summation order
// (16 interleaved strands per class, combined in order): deterministic for any number of groups.  blockDim = 256.
__global__ void k_layer2_reduce(const float* b2_partial, const double* loss_partial, int n_groups, int C, float* grad,
                                int64_t P, int64_t b2_off, float* loss_out, int N) {
  __shared__ double scratch[32];
  __shared__ float strands[16][L2_CMAX + 1];
  const int b = blockIdx.x, t = threadIdx.x;
  {
    const int c = t & 15, j = t >> 4;                 // class, strand
    float s = 0.f;
    if (c < C)
      for (int g = j; g < n_groups; g += 16) s += b2_partial[((int64_t)b * n_groups + g) * L2_CMAX + c];
    strands[j][c] = s;
  }
  __syncthreads();
  if (t < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += strands[j][t];
    grad[(int64_t)b * P + b2_off + t] = s;
  }
  double a = 0.0;
  for (int g = t; g < n_groups; g += blockDim.x) a += loss_partial[(int64_t)b * n_groups + g];
  double tot = block_sum<double>(a, scratch);
  if (t == 0 && loss_out) loss_out[b] = (float)(tot / (double)N);
}

// out[b*out_stride + i] = sum_s part[s*split_stride + b*part_stride + i]   (fixed order => deterministic)
__global__ void k_reduce_ksplits(const float* part, int splits, int64_t split_stride, int64_t part_stride, int64_t count,
                                 float* out, int64_t out_stride) {
  const int b = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(int64_t)k * split_stride + (int64_t)b * part_stride + i];
    out[(int64_t)b * out_stride + i] = s;
  }
}
// one accumulator sees at most TC_SPLIT_CHUNKS chunks (8192 K elements): <= 1536 truncating accumulations
constexpr int TC_SPLIT_CHUNKS = 256;

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    PYB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    PYB_REQUIRE(f != nullptr && q == cudaDriverEntryPointSuccess, PYB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
    fn = (PFN_encodeTiled)f;
  }
  return fn;
}
// 2-D bf16 tensor [rows, k] with row pitch ld_elems, box [TC_BK, box_rows], SWIZZLE_64B, zero OOB fill
// row-tile-blocked bf16 tensor [blocks][rows_per_block][128]: box [TC_BK, box_rows, 1], SWIZZLE_64B
static CUtensorMap make_map_blocked(const void* base, int64_t blocks, int rows_per_block, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[3] = {128, (cuuint64_t)rows_per_block, (cuuint64_t)blocks};
  cuuint64_t strides[2] = {256, (cuuint64_t)rows_per_block * 256};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PYB_REQUIRE(r == CUDA_SUCCESS, PYB_ERR_CUDA, "cuTensorMapEncodeTiled (blocked) failed");
  return m;
}
static CUtensorMap make_map(const void* base, int64_t k, int64_t rows, int64_t ld_elems, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  PYB_REQUIRE((ld_elems * 2) % 16 == 0, PYB_ERR_INVALID, "tensor map pitch must be a multiple of 16 bytes");
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PYB_REQUIRE(r == CUDA_SUCCESS, PYB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return m;
}

struct TcData {   // split bf16 operands derived from one [N, D] fp32 matrix resident in HBM
  bool ready = false;
  int64_t N = 0, Npad = 0;
  int D = 0;
  DevBuf<__nv_bfloat16> x_hi, x_lo, xt_hi, xt_lo;       // [N][D], [D+1][Npad]
  DevBuf<__nv_bfloat16> xtq_hi, xtq_lo;                 // [D+1][Npad] with the fused epilogue's row order (fused_row_pos)
  CUtensorMap mX_hi, mX_lo, mXT_hi, mXT_lo, mXTp_hi, mXTp_lo;   // mXTp: half-tile boxes for the hidden-major pair GEMM
  CUtensorMap mXTq_hi, mXTq_lo, mXTqp_hi, mXTqp_lo;     // the same two views of xtq
};
struct TcState {
  TcData train, aux;                                    // resident training set / minibatch or test inputs
  int D = 0, H = 0, C = 0;
  int64_t cap_chains = 0, cap_blocks = 0;               // capacity: chains, and 128-row blocks over all chains
  DevBuf<__nv_bfloat16> w_hi, w_lo;                     // W1^T  [chains*H][D]
  DevBuf<__nv_bfloat16> z_hi, z_lo;                     // dZ1^T blocked [block][H][128]
  DevBuf<__nv_bfloat16> a_hi, a_lo;                     // A1^T  blocked [block][H][128]
  DevBuf<__nv_bfloat16> z2_hi, z2_lo;                   // dZ2^T blocked [block][16][128]
  DevBuf<float> b2_partial;
  DevBuf<float> kpart, gpart;                           // split-K partial sums (gradient GEMMs / exported GEMM)
  DevBuf<double> loss_partial;
  CUtensorMap mW_hi, mW_lo, mWp_hi, mWp_lo, mZ_hi, mZ_lo, mZa_hi, mZa_lo, mA_hi, mA_lo, mZ2_hi, mZ2_lo;
};
static TcState* tc_state(pyb_handle* h) {
  if (!h->tc) h->tc = new TcState();
  return (TcState*)h->tc;
}
void tc_release(pyb_handle* h) {
  if (h->tc) delete (TcState*)h->tc;
  h->tc = nullptr;
}
void tc_invalidate_dataset(pyb_handle* h) {
  if (h->tc) ((TcState*)h->tc)->train.ready = false;
}

// shape conditions of the tensor path for a batch of n_rows data rows
bool tc_supported_rows(pyb_handle* h, int64_t n_rows) {
  const Model& m = h->model;
  if (m.n_layers != 2) return false;
  const LayerDesc& L1 = m.layer[0];
  const LayerDesc& L2 = m.layer[1];
  if (!L1.use_bias || !L2.use_bias) return false;
  if (L1.fan_out % 16 != 0 || L1.fan_out < 16 || L1.fan_out > 256) return false;
  if (L1.fan_in % 8 != 0 || L1.fan_in < 64) return false;
  if (L2.fan_out > L2_CMAX) return false;
  if (L1.act == PYB_ACT_SOFTMAX) return false;
  return n_rows >= 128;
}
bool tc_supported(pyb_handle* h, int64_t S) {
  (void)S;
  return h->have_data && tc_supported_rows(h, h->N);
}

template <int EPI, int ACT>
static void launch_gemm_inst(pyb_handle* h, int grid, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                             const CUtensorMap& b_hi, const CUtensorMap& b_lo, const TcGemmParams& p) {
  // per device and per instantiation; cheap enough to repeat (a process may drive several GPUs)
  PYB_CUDA(cudaFuncSetAttribute(tc_gemm_bf16x3<EPI, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  tc_gemm_bf16x3<EPI, ACT><<<grid, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(a_hi, a_lo, b_hi, b_lo, p);
}
template <int EPI, int ACT>
static void launch_pair_inst(pyb_handle* h, int grid, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                             const CUtensorMap& b_hi, const CUtensorMap& b_lo, const TcGemmParams& p) {
  PYB_CUDA(cudaFuncSetAttribute(tc_gemm_pair_bf16x3<EPI, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES));
  tc_gemm_pair_bf16x3<EPI, ACT><<<grid, TP_THREADS, TP_SMEM_BYTES, h->stream>>>(a_hi, a_lo, b_hi, b_lo, p);
}
// CTA-pair kernel: plain 2-D operands, H a multiple of 32, an even number of M tiles' worth of work per item
static bool pair_ok(const pyb_handle* h, const TcGemmParams& p) {
  if (!h->opt_tc_pair || p.b_blocked || p.H < 32 || p.n_mtiles < 2) return false;
  if (p.n_btiles > 0) return p.a_blocked && p.n_mtiles == 2 && (p.H % 16) == 0;   // hidden-major gradient GEMM
  if ((p.H % 32) != 0) return false;
  return !p.a_blocked && p.a_batch_rows == 0;
}
// bp_hi/bp_lo: the SAME B tensors described with half-height TMA boxes (H/2 rows) for the CTA-pair kernel, or null
static void launch_gemm_tc(pyb_handle* h, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                           const CUtensorMap& b_hi, const CUtensorMap& b_lo, TcGemmParams p, double flops,
                           const CUtensorMap* bp_hi = nullptr, const CUtensorMap* bp_lo = nullptr) {
  if (p.k_splits <= 1) { p.k_splits = 1; p.chunks_per_split = (p.K + TC_BK - 1) / TC_BK; p.split_stride = 0; }
  p.vec_store = p.epi == EPI_STORE && ((uintptr_t)p.out % 16 == 0) && (p.out_stride % 4 == 0) && (p.out_ld % 4 == 0) &&
                (p.split_stride % 4 == 0);
  int grid = std::min(p.total_items, h->sm_count);
  prof_begin(h);
  if (bp_hi && bp_lo && pair_ok(h, p) && p.n_btiles > 1 && p.epi == EPI_STORE && p.transpose_out && h->opt_tc_dual &&
      p.H <= 256 && (p.H / 2) * TC_BK * 2 <= 8192) {
    // hidden-major gradient GEMM: two feature tiles per item share every A stage
    p.total_items = p.n_batch * p.k_splits * ((p.n_btiles + 1) / 2);
    grid = std::min(2 * p.total_items, (h->sm_count / 2) * 2);
    PYB_CUDA(cudaFuncSetAttribute(tc_gemm_pair_dual_bf16x3, cudaFuncAttributeMaxDynamicSharedMemorySize, TD_SMEM_BYTES));
    tc_gemm_pair_dual_bf16x3<<<grid, TP_THREADS, TD_SMEM_BYTES, h->stream>>>(a_hi, a_lo, *bp_hi, *bp_lo, p);
    prof_end(h, flops);
    count_launch(h);
    return;
  }
  if (bp_hi && bp_lo && pair_ok(h, p)) {
    grid = std::min(2 * p.total_items, (h->sm_count / 2) * 2);
    if (p.epi == EPI_STORE) launch_pair_inst<EPI_STORE, 0>(h, grid, a_hi, a_lo, *bp_hi, *bp_lo, p);
    else if (p.act == PYB_ACT_RELU) launch_pair_inst<EPI_BIAS_ACT_T_SPLIT, PYB_ACT_RELU>(h, grid, a_hi, a_lo, *bp_hi, *bp_lo, p);
    else if (p.act == PYB_ACT_TANH) launch_pair_inst<EPI_BIAS_ACT_T_SPLIT, PYB_ACT_TANH>(h, grid, a_hi, a_lo, *bp_hi, *bp_lo, p);
    else if (p.act == PYB_ACT_SIGMOID) launch_pair_inst<EPI_BIAS_ACT_T_SPLIT, PYB_ACT_SIGMOID>(h, grid, a_hi, a_lo, *bp_hi, *bp_lo, p);
    else launch_pair_inst<EPI_BIAS_ACT_T_SPLIT, PYB_ACT_LINEAR>(h, grid, a_hi, a_lo, *bp_hi, *bp_lo, p);
    prof_end(h, flops);
    count_launch(h);
    return;
  }
  if (p.epi == EPI_STORE) launch_gemm_inst<EPI_STORE, 0>(h, grid, a_hi, a_lo, b_hi, b_lo, p);
  else if (p.act == PYB_ACT_RELU) launch_gemm_inst<EPI_BIAS_ACT_T_SPLIT, PYB_ACT_RELU>(h, grid, a_hi, a_lo, b_hi, b_lo, p);
  else if (p.act == PYB_ACT_TANH) launch_gemm_inst<EPI_BIAS_ACT_T_SPLIT, PYB_ACT_TANH>(h, grid, a_hi, a_lo, b_hi, b_lo, p);
  else if (p.act == PYB_ACT_SIGMOID) launch_gemm_inst<EPI_BIAS_ACT_T_SPLIT, PYB_ACT_SIGMOID>(h, grid, a_hi, a_lo, b_hi, b_lo, p);
  else launch_gemm_inst<EPI_BIAS_ACT_T_SPLIT, PYB_ACT_LINEAR>(h, grid, a_hi, a_lo, b_hi, b_lo, p);
  prof_end(h, flops);
  count_launch(h);
}


template <int CP>
static void launch_fused_inst(pyb_handle* h, int grid, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                              const CUtensorMap& b_hi, const CUtensorMap& b_lo, const TcGemmParams& p, const Layer2Params& l2) {
  PYB_CUDA(cudaFuncSetAttribute(tc_g1_layer2_fused<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TfCfg<CP>::SMEM));
  tc_g1_layer2_fused<CP><<<grid, TF_THREADS, TfCfg<CP>::SMEM, h->stream>>>(a_hi, a_lo, b_hi, b_lo, p, l2);
}
// fused G1 + layer 2 applies to: relu hidden layer of 128 or 256 units, CTA pairs enabled
static bool fused_ok(const pyb_handle* h, int H, int act1) {
  return h->opt_tc_pair && h->opt_tc_fuse && act1 == PYB_ACT_RELU && (H == 128 || H == 256);
}

static void tc_prepare_data(pyb_handle* h, TcData& d, const float* X, int64_t N, bool need_xt) {
  const Model& m = h->model;
  const int D = m.layer[0].fan_in;
  d.N = N; d.D = D;
  d.Npad = ((N + L2_ROWS - 1) / L2_ROWS) * L2_ROWS;
  const int64_t Npad = d.Npad;
  d.x_hi.alloc(N * D); d.x_lo.alloc(N * D);
  k_split_rows<<<(unsigned)std::min<int64_t>((N * D + 255) / 256, 65535), 256, 0, h->stream>>>(X, N, D, D, d.x_hi.p, d.x_lo.p, D);
  count_launch(h);
  d.mX_hi = make_map(d.x_hi.p, D, N, D, 128);
  d.mX_lo = make_map(d.x_lo.p, D, N, D, 128);
  if (need_xt) {
    d.xt_hi.alloc((int64_t)(D + 1) * Npad); d.xt_lo.alloc((int64_t)(D + 1) * Npad);
    PYB_CUDA(cudaMemsetAsync(d.xt_hi.p, 0, (size_t)(D + 1) * Npad * 2, h->stream));
    PYB_CUDA(cudaMemsetAsync(d.xt_lo.p, 0, (size_t)(D + 1) * Npad * 2, h->stream));
    dim3 g2((D + 31) / 32, (unsigned)((N + 31) / 32), 1), blk(32, 8);
    k_split_transpose<<<g2, blk, 0, h->stream>>>(X, 0, (int)N, D, D, d.xt_hi.p, d.xt_lo.p, 0, Npad);
    k_fill_bf16<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(d.xt_hi.p + (int64_t)D * Npad, N, 1.0f);  // ones row -> db1
    count_launch(h, 2);
    d.mXT_hi = make_map(d.xt_hi.p, Npad, D + 1, Npad, 128);
    d.mXT_lo = make_map(d.xt_lo.p, Npad, D + 1, Npad, 128);
    const int n_t = (D + 1 + 255) / 256, Ht = (((D + 1 + n_t - 1) / n_t) + 15) / 16 * 16;
    d.mXTp_hi = make_map(d.xt_hi.p, Npad, D + 1, Npad, Ht / 2);
    d.mXTp_lo = make_map(d.xt_lo.p, Npad, D + 1, Npad, Ht / 2);
    if (fused_ok(h, m.layer[0].fan_out, m.layer[0].act)) {
      d.xtq_hi.alloc((int64_t)(D + 1) * Npad); d.xtq_lo.alloc((int64_t)(D + 1) * Npad);
      PYB_CUDA(cudaMemsetAsync(d.xtq_hi.p, 0, (size_t)(D + 1) * Npad * 2, h->stream));
      PYB_CUDA(cudaMemsetAsync(d.xtq_lo.p, 0, (size_t)(D + 1) * Npad * 2, h->stream));
      k_split_transpose<<<g2, blk, 0, h->stream>>>(X, 0, (int)N, D, D, d.xtq_hi.p, d.xtq_lo.p, 0, Npad, 1);
      k_fill_ones_perm<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(d.xtq_hi.p + (int64_t)D * Npad, (int)N);
      count_launch(h, 2);
      d.mXTq_hi = make_map(d.xtq_hi.p, Npad, D + 1, Npad, 128);
      d.mXTq_lo = make_map(d.xtq_lo.p, Npad, D + 1, Npad, 128);
      d.mXTqp_hi = make_map(d.xtq_hi.p, Npad, D + 1, Npad, Ht / 2);
      d.mXTqp_lo = make_map(d.xtq_lo.p, Npad, D + 1, Npad, Ht / 2);
    }
  }
  d.ready = true;
}

// chain batch for S chains over Npad rows; grows the shared buffers (and their tensor maps) on demand
static int64_t tc_prepare_bufs(pyb_handle* h, TcState* st, int64_t S, int64_t Npad, bool backward) {
  const Model& m = h->model;
  st->D = m.layer[0].fan_in; st->H = m.layer[0].fan_out; st->C = m.layer[1].fan_out;
  const int H = st->H, D = st->D;
  const int64_t k_tiles = Npad / 128;
  // intermediates: A1^T hi/lo (+ dZ1^T hi/lo) = 4 (+4) bytes per (chain, hidden, row)
  int64_t per_chain = (int64_t)H * Npad * (backward ? 8 : 4) + (int64_t)H * D * 4;
  int64_t budget = (int64_t)(std::max(h->opt_workspace_mb, 20000.0) * 1024.0 * 1024.0);
  int64_t bc = std::max<int64_t>(1, budget / per_chain);
  if (h->opt_chain_batch > 0) bc = std::min<int64_t>(bc, h->opt_chain_batch);
  bc = std::min<int64_t>(bc, S);
  if (bc >= h->sm_count) bc = (bc / h->sm_count) * h->sm_count;      // whole waves of per-chain GEMM work
  bc = std::min<int64_t>(bc, 16384);
  const int64_t need_blocks = bc * k_tiles;
  if (bc > st->cap_chains || need_blocks > st->cap_blocks) {
    PYB_CUDA(cudaStreamSynchronize(h->stream));
    st->cap_chains = std::max(st->cap_chains, bc);
    st->cap_blocks = std::max(st->cap_blocks, need_blocks);
    const int64_t cb = st->cap_blocks, cc = st->cap_chains;
    st->w_hi.alloc(cc * H * D); st->w_lo.alloc(cc * H * D);
    st->a_hi.alloc(cb * H * 128); st->a_lo.alloc(cb * H * 128);
    st->z_hi.alloc(cb * H * 128); st->z_lo.alloc(cb * H * 128);
    st->z2_hi.alloc(cb * L2_CMAX * 128); st->z2_lo.alloc(cb * L2_CMAX * 128);
    // every block a kernel reads is fully rewritten by the producer kernels of the same evaluation
    // (G1 zero-fills rows >= N, k_layer2 writes whole 256-row tiles), so no clearing is needed
    st->mW_hi = make_map(st->w_hi.p, D, cc * H, D, H);
    st->mW_lo = make_map(st->w_lo.p, D, cc * H, D, H);
    st->mWp_hi = make_map(st->w_hi.p, D, cc * H, D, std::max(H / 2, 8));
    st->mWp_lo = make_map(st->w_lo.p, D, cc * H, D, std::max(H / 2, 8));
    st->mZ_hi = make_map_blocked(st->z_hi.p, cb, H, H);
    st->mZ_lo = make_map_blocked(st->z_lo.p, cb, H, H);
    st->mZa_hi = make_map_blocked(st->z_hi.p, cb, H, std::min(H, 128));   // dZ1^T as an A operand (128-row boxes)
    st->mZa_lo = make_map_blocked(st->z_lo.p, cb, H, std::min(H, 128));
    st->mA_hi = make_map_blocked(st->a_hi.p, cb, H, std::min(H, 128));
    st->mA_lo = make_map_blocked(st->a_lo.p, cb, H, std::min(H, 128));
    st->mZ2_hi = make_map_blocked(st->z2_hi.p, cb, L2_CMAX, L2_CMAX);
    st->mZ2_lo = make_map_blocked(st->z2_lo.p, cb, L2_CMAX, L2_CMAX);
  }
  return bc;
}

static void tc_pack_and_g1(pyb_handle* h, TcState* st, TcData& d, const float* th, int nb,
                           const Layer2Params* fused = nullptr) {
  const Model& m = h->model;
  const LayerDesc& L1 = m.layer[0];
  const int D = st->D, H = st->H;
  const int64_t N = d.N, Npad = d.Npad, P = m.P;
  // W1 [D,H] per chain -> W1^T hi/lo [H, D]
  dim3 g((H + 31) / 32, (D + 31) / 32, nb), blk(32, 8);
  k_split_transpose<<<g, blk, 0, h->stream>>>(th + L1.w_off, P, D, H, H, st->w_hi.p, st->w_lo.p, (int64_t)H * D, D);
  count_launch(h);
  // G1: A1^T = act(X W1 + b1)^T, split bf16 (all Npad/128 row tiles: rows >= N are written as zeros)
  TcGemmParams p = {};
  p.K = D; p.n_mtiles = (int)(Npad / 128); p.n_pairs = (p.n_mtiles + 1) / 2; p.n_batch = nb; p.H = H;
  p.a_batch_rows = 0; p.a_box_rows = 128; p.order = 1; p.sub_batch = 32; p.total_items = p.n_pairs * nb;
  p.epi = EPI_BIAS_ACT_T_SPLIT;
  p.bias = th + L1.b_off; p.bias_stride = P; p.act = L1.act;
  p.out_hi = st->a_hi.p; p.out_lo = st->a_lo.p; p.out_tiles = (int)(Npad / 128);
  p.M_valid = (int)N; p.N_valid = H;
  if (fused) {
    // G1 with layer 2 (+ loss, dZ2, dZ1) in its epilogue
    const int grid = std::min(2 * p.total_items, (h->sm_count / 2) * 2);
    const int C = fused->C;
    prof_begin(h);
    if (C <= 4) launch_fused_inst<4>(h, grid, d.mX_hi, d.mX_lo, st->mWp_hi, st->mWp_lo, p, *fused);
    else if (C <= 8) launch_fused_inst<8>(h, grid, d.mX_hi, d.mX_lo, st->mWp_hi, st->mWp_lo, p, *fused);
    else if (C <= 12) launch_fused_inst<12>(h, grid, d.mX_hi, d.mX_lo, st->mWp_hi, st->mWp_lo, p, *fused);
    else launch_fused_inst<16>(h, grid, d.mX_hi, d.mX_lo, st->mWp_hi, st->mWp_lo, p, *fused);
    prof_end(h, 2.0 * N * (D * (double)H + 3.0 * H * C) * nb);
    count_launch(h);
    return;
  }
  launch_gemm_tc(h, d.mX_hi, d.mX_lo, st->mW_hi, st->mW_lo, p, 2.0 * N * D * (double)H * nb, &st->mWp_hi, &st->mWp_lo);
}

static void tc_eval_on(pyb_handle* h, TcState* st, TcData& d, const int32_t* y_i, const float* y_f, const float* theta,
                       int64_t S, float scale, float* loss_out, float* grad_out) {
  const Model& m = h->model;
  const int64_t Bc = tc_prepare_bufs(h, st, S, d.Npad, true);
  const LayerDesc& L1 = m.layer[0];
  const LayerDesc& L2 = m.layer[1];
  const int D = st->D, H = st->H, C = st->C;
  const int64_t N = d.N, Npad = d.Npad, P = m.P;
  const int n_tiles = (int)(Npad / L2_ROWS);
  const bool fused = fused_ok(h, H, L1.act) && C <= L2_CMAX;
  // fused: one partial per (128-row tile, lane quadrant); unfused: one per k_layer2 block
  const int n_groups = fused ? (int)(Npad / 128) * 4
                             : std::min(n_tiles, std::max(1, (int)((8 * (int64_t)h->sm_count + Bc - 1) / Bc)));
  st->b2_partial.alloc((size_t)Bc * n_groups * L2_CMAX);
  st->loss_partial.alloc((size_t)Bc * n_groups);
  for (int64_t b0 = 0; b0 < S; b0 += Bc) {
    const int nb = (int)std::min<int64_t>(Bc, S - b0);
    const float* th = theta + b0 * P;
    float* gr = grad_out + b0 * P;
    // layer 2 + loss + dZ1^T, dZ2^T: inside G1's epilogue when the shape allows, else a separate pass over A1^T
    {
      Layer2Params p = {};
      p.a_hi = st->a_hi.p; p.a_lo = st->a_lo.p; p.zt_hi = st->z_hi.p; p.zt_lo = st->z_lo.p;
      p.z2_hi = st->z2_hi.p; p.z2_lo = st->z2_lo.p; p.k_tiles = (int)(Npad / 128);
      p.theta = th; p.P = P; p.w2_off = L2.w_off; p.b2_off = L2.b_off;
      p.H = H; p.C = C; p.N = (int)N; p.act1 = L1.act; p.out_act = L2.act; p.loss_kind = h->loss_kind;
      p.y_i = y_i; p.y_f = y_f; p.scale = scale;
      p.loss_partial = st->loss_partial.p; p.b2_partial = st->b2_partial.p; p.n_groups = n_groups;
      p.n_tiles = n_tiles;
      if (fused) {
        tc_pack_and_g1(h, st, d, th, nb, &p);
      } else {
        tc_pack_and_g1(h, st, d, th, nb);
        dim3 g(n_groups, nb);
        if (C <= 4) k_layer2<4><<<g, 128, 0, h->stream>>>(p);
        else if (C <= 8) k_layer2<8><<<g, 128, 0, h->stream>>>(p);
        else if (C <= 12) k_layer2<12><<<g, 128, 0, h->stream>>>(p);
        else k_layer2<16><<<g, 128, 0, h->stream>>>(p);
      }
      k_layer2_reduce<<<nb, 256, 0, h->stream>>>(st->b2_partial.p, st->loss_partial.p, n_groups, C, gr, P, L2.b_off,
                                                loss_out ? loss_out + b0 : nullptr, (int)N);
      count_launch(h, 2);
    }
    // the two reductions over the data rows are split-K (accumulator truncation, see TcGemmParams)
    const int nk_rows = (int)(Npad / TC_BK);
    const int splits = (nk_rows + TC_SPLIT_CHUNKS - 1) / TC_SPLIT_CHUNKS;
    const int64_t cnt1 = (int64_t)(D + 1) * H, cnt2 = (int64_t)H * C;
    if (splits > 1) st->kpart.alloc((size_t)splits * nb * std::max(cnt1, cnt2));
    // G3: dW2[h][c] = sum_r a1[r][h] dZ2[r][c]   (A = A1^T per chain, B = dZ2^T per chain, N = 16)
    {
      TcGemmParams p = {};
      p.K = (int)Npad; p.n_mtiles = (H + 127) / 128; p.n_pairs = (p.n_mtiles + 1) / 2; p.n_batch = nb; p.H = L2_CMAX;
      p.a_blocked = 1; p.b_blocked = 1; p.k_tiles = (int)(Npad / 128); p.a_box_rows = std::min(H, 128);
      p.a_batch_rows = H; p.order = 0; p.sub_batch = nb;
      p.epi = EPI_STORE; p.out_ld = C; p.M_valid = H; p.N_valid = C;
      if (splits > 1) {
        p.k_splits = splits; p.chunks_per_split = TC_SPLIT_CHUNKS; p.split_stride = nb * cnt2;
        p.out = st->kpart.p; p.out_stride = cnt2;
      } else {
        p.out = gr + L2.w_off; p.out_stride = P;
      }
      p.total_items = p.n_pairs * nb * std::max(splits, 1);
      launch_gemm_tc(h, st->mA_hi, st->mA_lo, st->mZ2_hi, st->mZ2_lo, p, 2.0 * N * (double)H * C * nb);
      if (splits > 1) {
        dim3 rg((unsigned)std::min<int64_t>((cnt2 + 255) / 256, 64), nb);
        k_reduce_ksplits<<<rg, 256, 0, h->stream>>>(st->kpart.p, splits, nb * cnt2, cnt2, cnt2, gr + L2.w_off, P);
        count_launch(h);
      }
    }
    // G2: [dW1; db1] = [X^T; 1] dZ1
    if (H == 256 && h->opt_tc_pair) {
      // hidden-major on the CTA-pair kernel: D[h, f] = sum_r dZ1^T[h, r] [X^T;1][f, r].  M = 256 hidden units is
      // exactly one CTA pair (no M padding), the D+1 feature rows are the N dimension in tiles of 256 plus one
      // narrow remainder tile, the accumulators are double-buffered so the partial-sum stores overlap the MMAs
      // (N = D+1 = 785 -> 4 tiles of 208 columns: 6 % padding instead of the 12.5 % of 7 x 128 M tiles)
      const int n_t = (D + 1 + 255) / 256;
      const int Ht = (((D + 1 + n_t - 1) / n_t) + 15) / 16 * 16;
      TcGemmParams p = {};
      p.K = (int)Npad; p.n_mtiles = 2; p.n_pairs = 1; p.n_batch = nb; p.H = Ht;
      p.a_blocked = 1; p.b_blocked = 0; p.k_tiles = (int)(Npad / 128); p.a_box_rows = 128;
      p.n_btiles = n_t; p.b_row0 = 0; p.transpose_out = 1; p.n_cols_total = D + 1;
      p.epi = EPI_STORE; p.out_ld = H; p.M_valid = H; p.N_valid = Ht;
      if (splits > 1) {
        p.k_splits = splits; p.chunks_per_split = TC_SPLIT_CHUNKS; p.split_stride = nb * cnt1;
        p.out = st->kpart.p; p.out_stride = cnt1;
      } else {
        p.out = gr; p.out_stride = P;
      }
      p.total_items = nb * std::max(splits, 1) * n_t;
      launch_gemm_tc(h, st->mZa_hi, st->mZa_lo, fused ? d.mXTq_hi : d.mXT_hi, fused ? d.mXTq_lo : d.mXT_lo, p,
                     2.0 * N * (double)(D + 1) * H * nb, fused ? &d.mXTqp_hi : &d.mXTp_hi, fused ? &d.mXTqp_lo : &d.mXTp_lo);
    } else {
      TcGemmParams p = {};
      p.K = (int)Npad; p.n_mtiles = (D + 1 + 127) / 128; p.n_pairs = (p.n_mtiles + 1) / 2; p.n_batch = nb; p.H = H;
      p.a_blocked = 0; p.b_blocked = 1; p.k_tiles = (int)(Npad / 128); p.a_box_rows = 128;
      p.a_batch_rows = 0; p.order = 0; p.sub_batch = nb;
      p.epi = EPI_STORE; p.out_ld = H; p.M_valid = D + 1; p.N_valid = H;
      if (splits > 1) {
        p.k_splits = splits; p.chunks_per_split = TC_SPLIT_CHUNKS; p.split_stride = nb * cnt1;
        p.out = st->kpart.p; p.out_stride = cnt1;
      } else {
        p.out = gr; p.out_stride = P;
      }
      p.total_items = p.n_pairs * nb * std::max(splits, 1);
      launch_gemm_tc(h, fused ? d.mXTq_hi : d.mXT_hi, fused ? d.mXTq_lo : d.mXT_lo, st->mZ_hi, st->mZ_lo, p,
                     2.0 * N * (double)(D + 1) * H * nb);
    }
    if (splits > 1) {
      dim3 rg((unsigned)std::min<int64_t>((cnt1 + 255) / 256, 256), nb);
      k_reduce_ksplits<<<rg, 256, 0, h->stream>>>(st->kpart.p, splits, nb * cnt1, cnt1, cnt1, gr, P);
      count_launch(h);
    }
  }
  PYB_CUDA(cudaGetLastError());
}

void tc_eval(pyb_handle* h, const float* theta, int64_t S, float scale, float* loss_out, float* grad_out) {
  TcState* st = tc_state(h);
  if (!st->train.ready) tc_prepare_data(h, st->train, h->X.p, h->N, true);
  tc_eval_on(h, st, st->train, h->y_i.p, h->y_f.p, theta, S, scale, loss_out, grad_out);
}

// loss + gradient on an arbitrary device-resident batch (SVGD minibatches): operands are re-derived per call
void tc_eval_batch(pyb_handle* h, const float* Xb, const int32_t* yb_i, const float* yb_f, int64_t Nb, const float* theta,
                   int64_t S, float scale, float* loss_out, float* grad_out) {
  TcState* st = tc_state(h);
  tc_prepare_data(h, st->aux, Xb, Nb, true);
  tc_eval_on(h, st, st->aux, yb_i, yb_f, theta, S, scale, loss_out, grad_out);
}

// forward only: out [S, N, C] (softmax / output activation applied) for device-resident inputs x [N, D]
void tc_forward(pyb_handle* h, const float* theta, int64_t S, const float* x, int64_t N, float* out) {
  TcState* st = tc_state(h);
  const Model& m = h->model;
  tc_prepare_data(h, st->aux, x, N, false);
  TcData& d = st->aux;
  const int64_t Bc = tc_prepare_bufs(h, st, S, d.Npad, false);
  const LayerDesc& L2 = m.layer[1];
  const int64_t P = m.P;
  for (int64_t b0 = 0; b0 < S; b0 += Bc) {
    const int nb = (int)std::min<int64_t>(Bc, S - b0);
    const float* th = theta + b0 * P;
    tc_pack_and_g1(h, st, d, th, nb);
    Layer2FwdParams p = {};
    p.a_hi = st->a_hi.p; p.a_lo = st->a_lo.p; p.k_tiles = (int)(d.Npad / 128);
    p.theta = th; p.P = P; p.w2_off = L2.w_off; p.b2_off = L2.b_off;
    p.H = st->H; p.C = st->C; p.N = (int)N; p.out_act = L2.act; p.n_tiles = (int)(d.Npad / L2_ROWS);
    p.out = out + b0 * N * st->C;
    dim3 g(std::min(p.n_tiles, 64), nb);
    k_layer2_fwd<<<g, 128, 0, h->stream>>>(p);
    count_launch(h);
  }
  PYB_CUDA(cudaGetLastError());
}

// ---- building blocks exported to svgd.cu (Gram matrix and Stein contraction on the tensor cores) ----
void tc_split_rows(pyb_handle* h, const float* src, int64_t R, int C, int64_t lds, void* hi, void* lo, int64_t ldd) {
  k_split_rows<<<(unsigned)std::min<int64_t>((R * C + 255) / 256, 65535), 256, 0, h->stream>>>(
      src, R, C, lds, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ldd);
  count_launch(h);
}
// src [R, C] fp32 -> hi/lo [C, R] bf16 with row pitch ldd
void tc_split_transpose(pyb_handle* h, const float* src, int R, int C, int64_t lds, void* hi, void* lo, int64_t ldd) {
  dim3 g((C + 31) / 32, (R + 31) / 32, 1), blk(32, 8);
  k_split_transpose<<<g, blk, 0, h->stream>>>(src, 0, R, C, lds, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, 0, ldd);
  count_launch(h);
}
// out[m][n] = sum_k A[a_row0+m][k] B[n][k] for m < M, n < Nn; split operands, K-major, pitches in elements
void tc_gemm_split(pyb_handle* h, const void* a_hi, const void* a_lo, int64_t lda, int64_t a_rows_total, int a_row0, int M,
                   const void* b_hi, const void* b_lo, int64_t ldb, int Nn, int64_t K, float* out, int64_t ldc) {
  PYB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, PYB_ERR_INVALID, "operand pitches must be multiples of 8 elements");
  CUtensorMap ma_h = make_map(a_hi, K, a_rows_total, lda, 128), ma_l = make_map(a_lo, K, a_rows_total, lda, 128);
  const int Hn = Nn >= 256 ? 256 : ((Nn + 15) / 16) * 16;
  CUtensorMap mb_h = make_map(b_hi, K, Nn, ldb, Hn), mb_l = make_map(b_lo, K, Nn, ldb, Hn);
  TcGemmParams p = {};
  p.K = (int)K; p.n_mtiles = (M + 127) / 128; p.n_pairs = (p.n_mtiles + 1) / 2; p.n_batch = (Nn + Hn - 1) / Hn; p.H = Hn;
  p.a_row0 = a_row0; p.a_batch_rows = 0; p.a_box_rows = 128; p.order = 1; p.sub_batch = 8;
  p.epi = EPI_STORE; p.out = out; p.out_stride = Hn; p.out_ld = (int)ldc; p.M_valid = M; p.N_valid = Hn; p.n_cols_total = Nn;
  const int nk = (int)((K + TC_BK - 1) / TC_BK);
  const int splits = (nk + TC_SPLIT_CHUNKS - 1) / TC_SPLIT_CHUNKS;
  TcState* st = tc_state(h);
  if (splits > 1) {      // long reductions (Gram over P = 1e5 parameters): split-K partials, fixed-order sum
    st->gpart.alloc((size_t)splits * M * ldc);
    p.k_splits = splits; p.chunks_per_split = TC_SPLIT_CHUNKS; p.split_stride = (int64_t)M * ldc; p.out = st->gpart.p;
  }
  p.total_items = p.n_pairs * p.n_batch * std::max(splits, 1);
  CUtensorMap mp_h = make_map(b_hi, K, Nn, ldb, std::max(Hn / 2, 8)), mp_l = make_map(b_lo, K, Nn, ldb, std::max(Hn / 2, 8));
  launch_gemm_tc(h, ma_h, ma_l, mb_h, mb_l, p, 2.0 * M * (double)Nn * (double)K, &mp_h, &mp_l);
  if (splits > 1) {
    const int64_t cnt = (int64_t)M * ldc;
    dim3 rg((unsigned)std::min<int64_t>((cnt + 255) / 256, 4096), 1);
    k_reduce_ksplits<<<rg, 256, 0, h->stream>>>(st->gpart.p, splits, cnt, 0, cnt, out, 0);
    count_launch(h);
  }
}

// debug / unit-test entry: D[M,Nn] = A[M,K] B[Nn,K]^T through the tcgen05 kernel (host pointers)
void tc_debug_gemm(pyb_handle* h, const float* A, const float* B, int M, int Nn, int K, float* Dout) {
  PYB_REQUIRE(Nn % 16 == 0 && Nn >= 16 && Nn <= 256 && K % 8 == 0, PYB_ERR_INVALID, "Nn%16, Nn<=256, K%8 required");
  DevBuf<float> dA, dB, dD;
  DevBuf<__nv_bfloat16> ah, al, bh, bl;
  dA.alloc((size_t)M * K); dB.alloc((size_t)Nn * K); dD.alloc((size_t)M * Nn);
  ah.alloc((size_t)M * K); al.alloc((size_t)M * K); bh.alloc((size_t)Nn * K); bl.alloc((size_t)Nn * K);
  PYB_CUDA(cudaMemcpyAsync(dA.p, A, (size_t)M * K * 4, cudaMemcpyHostToDevice, h->stream));
  PYB_CUDA(cudaMemcpyAsync(dB.p, B, (size_t)Nn * K * 4, cudaMemcpyHostToDevice, h->stream));
  PYB_CUDA(cudaMemsetAsync(dD.p, 0xff, (size_t)M * Nn * 4, h->stream));
  k_split_rows<<<1024, 256, 0, h->stream>>>(dA.p, M, K, K, ah.p, al.p, K);
  k_split_rows<<<1024, 256, 0, h->stream>>>(dB.p, Nn, K, K, bh.p, bl.p, K);
  CUtensorMap ma_h = make_map(ah.p, K, M, K, 128), ma_l = make_map(al.p, K, M, K, 128);
  CUtensorMap mb_h = make_map(bh.p, K, Nn, K, Nn), mb_l = make_map(bl.p, K, Nn, K, Nn);
  TcGemmParams p = {};
  p.K = K; p.n_mtiles = (M + 127) / 128; p.n_pairs = (p.n_mtiles + 1) / 2; p.n_batch = 1; p.H = Nn;
  p.order = 0; p.sub_batch = 1; p.total_items = p.n_pairs;
  p.epi = EPI_STORE; p.out = dD.p; p.out_stride = 0; p.out_ld = Nn; p.M_valid = M; p.N_valid = Nn; p.a_batch_rows = 0; p.a_box_rows = 128;
  CUtensorMap mp_h = make_map(bh.p, K, Nn, K, std::max(Nn / 2, 8)), mp_l = make_map(bl.p, K, Nn, K, std::max(Nn / 2, 8));
  launch_gemm_tc(h, ma_h, ma_l, mb_h, mb_l, p, 2.0 * M * Nn * (double)K, &mp_h, &mp_l);
  PYB_CUDA(cudaMemcpyAsync(Dout, dD.p, (size_t)M * Nn * 4, cudaMemcpyDeviceToHost, h->stream));
  PYB_CUDA(cudaStreamSynchronize(h->stream));
  PYB_CUDA(cudaGetLastError());
}

}  // namespace pyb

extern "C" int pyb_debug_tc_gemm(pyb_handle* h, const float* A, const float* B, int32_t M, int32_t Nn, int32_t K,
                                 float* D) {
  try {
    if (!h || !A || !B || !D) throw pyb::Error(PYB_ERR_INVALID, "NULL argument");
    PYB_CUDA(cudaSetDevice(h->device));
    pyb::tc_debug_gemm(h, A, B, M, Nn, K, D);
  } catch (const pyb::Error& e) {
    pyb::set_last_error(e.what());
    return e.code;
  }
  return PYB_OK;
}

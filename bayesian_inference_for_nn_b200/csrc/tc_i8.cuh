// tc_i8.cuh — the int8-slice operand scheme of the tensor path (tcgen05.mma kind::i8): operand producers and the
// hidden-major dW1 GEMM.  (The forward GEMM on slices is the I8 mode of tc_g1_layer2_fused, tc_fused.cuh.)
//
// x = s / 127 * (hi + lo / 254) with int8 hi, lo in [-127, 127] and ONE scale s per index of the non-contracted axis (it
// factors out of the dot product).  A product of two such operands is hh + (hl + lh) / 254 (the ll term, 2^-16 of the
// result, is dropped): 3 kind::i8 MMAs (K = 32, twice the kind::f16 rate) per 32 K-elements instead of the 6 kind::f16
// MMAs of bf16x3, 2 operand bytes per element instead of 4, and the int32 accumulators are exact — no truncating fp32
// accumulation, split-K only for parallelism and the int32 range.  The hl + lh sum carries the weight 1 / 254 and needs
// its own accumulator (TMEM columns [256, 512)).
//
// Where the precision goes (tools/study_split_precision.py, tools/study_i8_plan.py; 1e-4 parity budget): 16-bit FIXED
// point pays the max / rms of a weight column (3.3 for 784 Gaussians) that bf16x3's floating parts do not — the forward
// GEMM's W1 slices decide the gradient error (3.5e-5 - 6e-5 norm-wise), the data slices and the whole backward GEMM
// add 5e-6.  W1's error enters as X dW1, so the data are CENTRED by their feature means before slicing (the mean part
// mu^T W1 is exact: it is folded into the bias by the pack kernel): for U[0,1) data that halves the error.
#pragma once
#include "tc_fused.cuh"

namespace pyb {

// ------------------------------------------------------------------------------------------
// operand producers
// ------------------------------------------------------------------------------------------
// column statistics of X [N, D] in two deterministic levels: grid (ceil(D / 32), RC) blocks of (32, 32) threads each reduce
// one chunk of rows to a partial sum (double) and a partial absmax per feature, k_col_stats_final adds the RC partials
// in a fixed order: mean[d], absmax[d]
constexpr int I8_STAT_CHUNKS = 64;
__global__ void k_col_stats(const float* __restrict__ X, int64_t N, int D, double* __restrict__ psum, float* __restrict__ pmax) {
  __shared__ double ssum[32][33];
  __shared__ float smax[32][33];
  const int d = blockIdx.x * 32 + threadIdx.x;
  const int64_t per = (N + gridDim.y - 1) / gridDim.y;
  const int64_t r_begin = (int64_t)blockIdx.y * per, r_end = min(N, r_begin + per);
  double s = 0.0;
  float m = 0.f;
  if (d < D)
    for (int64_t r = r_begin + threadIdx.y; r < r_end; r += 32) {
      const float v = X[r * D + d];
      s += (double)v;
      m = fmaxf(m, fabsf(v));
    }
  ssum[threadIdx.y][threadIdx.x] = s;
  smax[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
    double tot = 0.0;
    float mm = 0.f;
    for (int j = 0; j < 32; ++j) { tot += ssum[j][threadIdx.x]; mm = fmaxf(mm, smax[j][threadIdx.x]); }
    psum[(int64_t)blockIdx.y * D + d] = tot;
    pmax[(int64_t)blockIdx.y * D + d] = mm;
  }
}
__global__ void k_col_stats_final(const double* __restrict__ psum, const float* __restrict__ pmax, int chunks, int64_t N, int D,
                                  float* __restrict__ mean, float* __restrict__ absmax) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  double tot = 0.0;
  float mm = 0.f;
  for (int j = 0; j < chunks; ++j) { tot += psum[(int64_t)j * D + d]; mm = fmaxf(mm, pmax[(int64_t)j * D + d]); }
  mean[d] = (float)(tot / (double)N);
  absmax[d] = mm;
}
__device__ __forceinline__ void slice_i8(float q, int& hi, int& lo) {   // q = x * 127 / s
  const float h = fminf(fmaxf(rintf(q), -127.f), 127.f);
  const float l = fminf(fmaxf(rintf((q - h) * 254.0f), -127.f), 127.f);
  hi = __float2int_rn(h);
  lo = __float2int_rn(l);
}
// forward operand: rows of X centred by the feature means, one scale per data row.  One warp per row; Dk = pitch (bytes)
__global__ void k_slice_x_rows_i8(const float* __restrict__ X, int64_t N, int D, const float* __restrict__ mean,
                                  int8_t* __restrict__ hi, int8_t* __restrict__ lo, float* __restrict__ sx, int Dk) {
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < N; r += (int64_t)gridDim.x * wpb) {
    const float* x = X + r * D;
    float m = 0.f;
    for (int k = lane; k < D; k += 32) m = fmaxf(m, fabsf(x[k] - mean[k]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float s = m > 0.f ? m : 1.f;
    if (lane == 0) sx[r] = s;
    const float inv = 127.0f / s;
    for (int k4 = lane * 4; k4 < Dk; k4 += 128) {
      uint32_t hw = 0, lw = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int h = 0, l = 0;
        if (k4 + j < D) slice_i8((x[k4 + j] - mean[k4 + j]) * inv, h, l);
        hw |= ((uint32_t)h & 0xffu) << (8 * j);
        lw |= ((uint32_t)l & 0xffu) << (8 * j);
      }
      *reinterpret_cast<uint32_t*>(hi + r * Dk + k4) = hw;
      *reinterpret_cast<uint32_t*>(lo + r * Dk + k4) = lw;
    }
  }
}
// backward operand: [X^T; 1] slices [D + 1][Npad] in the fused epilogue's row order, one scale per FEATURE (the data rows
// are the contracted axis here); sf[d] = absmax[d] (1 for an all-zero feature), sf[D] = 1 with the ones row = 127.
// grid (ceil(D/32), ceil(N/32)), block (32, 8); the arrays are zeroed beforehand
__global__ void k_slice_xt_i8(const float* __restrict__ X, int64_t N, int D, const float* __restrict__ absmax,
                              int8_t* __restrict__ hi, int8_t* __restrict__ lo, int64_t Npad) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = (int64_t)r0 + i;
    const int c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < N && c < D) ? X[r * D + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const int64_t r = (int64_t)r0 + threadIdx.x;
    if (c < D && r < N) {
      const float s = absmax[c] > 0.f ? absmax[c] : 1.f;
      int h, l;
      slice_i8(t[threadIdx.x][i] * (127.0f / s), h, l);
      const int64_t o = (int64_t)c * Npad + (r & ~(int64_t)127) + fused_row_pos((int)(r & 127));
      hi[o] = (int8_t)h;
      lo[o] = (int8_t)l;
    }
  }
}
__global__ void k_xt_i8_tail(int8_t* hi_ones, int64_t N, const float* absmax, float* sf, int D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) hi_ones[(i & ~(int64_t)127) + fused_row_pos((int)(i & 127))] = 127;   // 127 / 127 * 1 = the ones row (db1)
  if (i <= D) sf[i] = (i == D) ? (1.0f / 127.0f) : ((absmax[i] > 0.f ? absmax[i] : 1.f) / 127.0f);   // dequantisation factors
}

// per chain b and hidden unit h (grid (ceil(H/32), nb), block (32, 8)):
//   W1 [D, H] -> W1^T slices [nb*H][Dk] (K-major), scale s_w = max_d |W1[d, h]|, cw = s_w / 127^2,
//   b1c = b1 + sum_d mean[d] W1[d, h] (the centred-data correction, double accumulation in a fixed order),
//   zq = 127 / s_z and zd = s_z / 127 with s_z = 1.001 * (max_c W2[h, c] - min_c W2[h, c]) * scale / N >= |dZ1[:, h]|:
//   dZ1 = sum_c dZ2[c] W2[h, c] with sum_c dZ2[c] = 0 and sum_c |dZ2[c]| <= 2 scale / N for the softmax cross-entropy
__global__ void k_pack_w1_i8(const float* __restrict__ theta, int64_t P, int64_t w1_off, int64_t b1_off, int64_t w2_off,
                             int D, int H, int C, const float* __restrict__ mean, float invN, int8_t* __restrict__ hi,
                             int8_t* __restrict__ lo, int Dk, float* __restrict__ cw, float* __restrict__ b1c,
                             float* __restrict__ zq, float* __restrict__ zd) {
  __shared__ float smax[8][33];
  __shared__ double ssum[8][33];
  __shared__ float s_inv[32];
  __shared__ float t[32][33];
  const int b = blockIdx.y, h0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;
  const float* th = theta + (int64_t)b * P;
  const float* W1 = th + w1_off;
  const int h = h0 + tx;
  float m = 0.f;
  double s = 0.0;
  if (h < H)
    for (int d = ty; d < D; d += 8) {
      const float w = W1[(int64_t)d * H + h];
      m = fmaxf(m, fabsf(w));
      s += (double)mean[d] * (double)w;
    }
  smax[ty][tx] = m;
  ssum[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float mm = 0.f;
    double tot = 0.0;
    for (int j = 0; j < 8; ++j) { mm = fmaxf(mm, smax[j][tx]); tot += ssum[j][tx]; }
    const float sw = mm > 0.f ? mm : 1.f;
    s_inv[tx] = 127.0f / sw;
    if (h < H) {
      cw[(int64_t)b * H + h] = sw * (1.0f / (127.0f * 127.0f));
      b1c[(int64_t)b * H + h] = (float)((double)th[b1_off + h] + tot);
      const float* w2 = th + w2_off + (int64_t)h * C;
      float lo2 = w2[0], hi2 = w2[0];
      for (int c = 1; c < C; ++c) { lo2 = fminf(lo2, w2[c]); hi2 = fmaxf(hi2, w2[c]); }
      // + the softmax probabilities sum to 1 only to float32 rounding: sum_c dZ2[c] = O(1e-6) / N multiplies the row's offset
      float sz = (1.001f * (hi2 - lo2) + 8e-6f * fmaxf(fabsf(hi2), fabsf(lo2))) * invN;
      if (!(sz > 0.f) || !isfinite(sz)) sz = 1.f;               // W2 row constant (or not finite): dZ1 is 0 (or NaN anyway)
      zq[(int64_t)b * H + h] = 127.0f / sz;
      zd[(int64_t)b * H + h] = sz * (1.0f / 127.0f);
    }
  }
  __syncthreads();
  // transpose 32 x 32 tiles through shared memory: rows of the slice tensors are hidden units, 4 K-elements per store
  for (int d0 = 0; d0 < Dk; d0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      const int d = d0 + i;
      t[i][tx] = (d < D && h < H) ? W1[(int64_t)d * H + h] : 0.f;
    }
    __syncthreads();
    {
      const int u = ty * 4 + (tx >> 3);            // unit inside the tile: 0..31
      const int k4 = (tx & 7) * 4;                 // 4 consecutive K-elements
      if (h0 + u < H && d0 + k4 < Dk) {
        uint32_t hw = 0, lw = 0;
        const float inv = s_inv[u];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int hh, ll;
          slice_i8(t[k4 + j][u] * inv, hh, ll);
          hw |= ((uint32_t)hh & 0xffu) << (8 * j);
          lw |= ((uint32_t)ll & 0xffu) << (8 * j);
        }
        const int64_t o = ((int64_t)b * H + h0 + u) * Dk + d0 + k4;
        *reinterpret_cast<uint32_t*>(hi + o) = hw;
        *reinterpret_cast<uint32_t*>(lo + o) = lw;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Hidden-major dW1 GEMM on int8 slices: D[h, f] = sum_r dZ1^T[h, r] [X^T;1][f, r] for ONE feature tile per item.
// CTA pair, M = 256 hidden units = one chain, N = Ht feature rows, K = data rows in split-K segments (exact int32
// accumulation: the segments exist for parallelism, L2 locality of the shared operand and the int32 range, not for
// rounding).  hh accumulates in TMEM columns [0, Ht), hl + lh in [256, 256 + Ht); the epilogue (not overlapped: one per
// segment, ~4 % of an item) stores s_z[h] / 127 * s_f[f] / 127 * (hh + cross / 254) transposed as fp32 partial sums.
// A = dZ1^T slices, blocked [chain][tile][H][128] (3-D u8 map), B = [X^T;1] slices [D+1][Npad] (2-D u8 map, half-tile boxes).
// Roles: warps 0-7 epilogue, 8 TMA producer, 9 MMA issuer, 10 TMEM allocator.
// ------------------------------------------------------------------------------------------
constexpr int TI_STAGES = 6;
constexpr int TI_SMEM_BYTES = TI_STAGES * TP_STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers*/;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TP_THREADS, 1)
tc_gemm_pair_dw1_i8(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const TcGemmParams p, const float* __restrict__ zd, const float* __restrict__ sf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint64_t* bars = (uint64_t*)(smem + TI_STAGES * TP_STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [TI_STAGES]   (leader CTA)
  uint64_t* empty_bar = bars + TI_STAGES;        // [TI_STAGES]   (one per CTA)
  uint64_t* tmem_full = bars + 2 * TI_STAGES;    // (one per CTA)
  uint64_t* tmem_empty = bars + 2 * TI_STAGES + 1;   // (leader CTA: both epilogues have drained)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * TI_STAGES + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int nk = (p.K + 63) / 64;                // stages of 64 data rows
  const int half_rows = p.H >> 1;                // rows of the feature tile this CTA stages
  const uint32_t cta_bytes = 2 * TC_A_TILE_BYTES + 2 * (uint32_t)half_rows * 64;

  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < TI_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
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
    // ===== TMA producer (both CTAs): own 128 hidden units of dZ1^T[b], own half of the feature tile =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        int b, mp, split, bt;
        tc_decode_pair(p, item, b, mp, split, bt);
        const int kc_begin = split * p.chunks_per_split, kc_end = min(nk, kc_begin + p.chunks_per_split);
        const int arow = (int)rank * 128;
        const int brow = p.b_row0 + bt * p.H + (int)rank * half_rows;
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait_sleep(&empty_bar[stage], phase ^ 1);
          uint8_t* st = stage_base + stage * TP_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * cta_bytes);
          const int k0 = kc * 64;
          tma_load_3d_pair(st, &tmA_hi, &full_bar[stage], k0 & 127, arow, b * p.k_tiles + (k0 >> 7));
          tma_load_3d_pair(st + TC_A_TILE_BYTES, &tmA_lo, &full_bar[stage], k0 & 127, arow, b * p.k_tiles + (k0 >> 7));
          tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES, &tmB_hi, &full_bar[stage], k0, brow);
          tma_load_2d_pair(st + 2 * TC_A_TILE_BYTES + 8192, &tmB_lo, &full_bar[stage], k0, brow);
          if (++stage == TI_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0 && lane == 0) {
      // D = S32, A = B = signed 8-bit, K-major, N = Ht, M = 256 across the pair
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.H >> 3) << 17) | ((256u >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        int b, mp, split, bt;
        tc_decode_pair(p, item, b, mp, split, bt);
        const int kc_begin = split * p.chunks_per_split, kc_end = min(nk, kc_begin + p.chunks_per_split);
        mbar_wait_sleep(tmem_empty, acc_phase ^ 1);
        tc_fence_after();
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait_sleep(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(stage_base + stage * TP_STAGE_BYTES);
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t koff = ks * 32;
            const uint64_t ah = make_smem_desc_sw64(st + koff);
            const uint64_t al = make_smem_desc_sw64(st + TC_A_TILE_BYTES + koff);
            const uint64_t bh = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + koff);
            const uint64_t bl = make_smem_desc_sw64(st + 2 * TC_A_TILE_BYTES + 8192 + koff);
            const uint32_t accum = (kc != kc_begin) || (ks != 0);
            tc_mma_i8_pair(tmem_base, ah, bh, idesc, accum);
            tc_mma_i8_pair(tmem_base + 256, ah, bl, idesc, accum);
            tc_mma_i8_pair(tmem_base + 256, al, bh, idesc, 1);
          }
          tc_commit_pair(&empty_bar[stage]);
          if (++stage == TI_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(tmem_full);
        acc_phase ^= 1;
      }
    }
  } else if (warp < 8) {
    // ===== epilogue (both CTAs): warp = (column half, lane quadrant); transposed partial-sum stores =====
    const int chalf = warp >> 2;
    const int et = threadIdx.x & 127;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int n_chunks = (p.H + 15) >> 4;                          // 16-column chunks of the feature tile
    const int ch_begin = chalf ? (n_chunks + 1) >> 1 : 0, ch_end = chalf ? n_chunks : (n_chunks + 1) >> 1;
    uint32_t acc_phase = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      int b, mp, split, bt;
      tc_decode_pair(p, item, b, mp, split, bt);
      const int row = (int)rank * 128 + et;                         // hidden unit
      const float dq = row < p.M_valid ? zd[(int64_t)b * p.M_valid + row] : 0.f;
      mbar_wait_sleep(tmem_full, acc_phase);
      tc_fence_after();
      const int colbase = p.b_row0 + bt * p.H;
      float* ob = p.out + (int64_t)split * p.split_stride + (int64_t)b * p.out_stride + row;
      for (int ch = ch_begin; ch < ch_end; ++ch) {
        const int c0 = ch * 16;
        uint32_t rh[16], rc[16];
        tc_ld16_raw(tmem_base + lane_base + (uint32_t)c0, rh);
        tc_ld16_raw(tmem_base + lane_base + (uint32_t)(256 + c0), rc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < p.M_valid) {
          float* o = ob + (int64_t)(colbase + c0) * p.out_ld;
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (colbase + c0 + q < p.n_cols_total && c0 + q < p.H) {
              const float acc = fmaf((float)(int)rc[q], 1.0f / 254.0f, (float)(int)rh[q]);
              o[(int64_t)q * p.out_ld] = acc * (dq * __ldg(sf + colbase + c0 + q));
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

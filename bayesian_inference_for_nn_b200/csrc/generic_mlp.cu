// generic_mlp.cu — shape-agnostic fp32 SIMT path: batched Dense forward/backward for S parameter
// vectors at once.  Any depth/width the Keras JSON describes runs here; the fused small-width
// kernel (fused_small.cu) and the tcgen05 path (tc_path.cu) are specialisations that are checked
// against this one.
//
// Replaces, for S samples at once: model(X) (HMC.py:155, SVGD.py:106, BayesianModel.py:124),
// the Keras loss (HMC.py:157, SVGD.py:107 via Dataset.py:152-159) and tape.gradient
// (HMC.py:132-135, SVGD.py:110).
//
// HBM layout: activations of a chain batch are [Bc, N, width] fp32 row-major per layer; the
// parameters stay in the flat [S, P] particle buffer and every GEMM addresses W/b/dW/db in place
// through (offset, batch stride = P).
#include "common.cuh"

namespace pyb {

struct GemmArgs {
  int M, N, K;
  const float* A; int lda; int64_t strideA;
  const float* B; int ldb; int64_t strideB;
  float* C; int ldc; int64_t strideC;
  const float* bias; int64_t strideBias;   // fwd epilogue (nullable)
  int act;                                 // fwd epilogue activation
  const float* mask; int ldmask; int64_t strideMask; int mask_act;  // bwd_a epilogue: *= act'(mask)
  int splits, kchunk;                      // split-K (bwd_w): C is [splits][batch] partials when splits>1
  int64_t strideSplit;
};

constexpr int BM = 64, BN = 64, BK = 16;

// C[M,N] = op(A)[M,K] * op(B)[K,N]; TA: A stored [K,M] (lda = row length of the stored matrix),
// TB: B stored [N,K].
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_sgemm(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int batch = blockIdx.z / g.splits, split = blockIdx.z % g.splits;
  const float* A = g.A + batch * g.strideA;
  const float* B = g.B + batch * g.strideB;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int kbeg = split * g.kchunk;
  const int kend = min(g.K, kbeg + g.kchunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int e = tid + r * 256;
      int m, k;
      if (TA) { m = e & (BM - 1); k = e >> 6; } else { k = e & (BK - 1); m = e >> 4; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < kend) v = TA ? A[(int64_t)gk * g.lda + gm] : A[(int64_t)gm * g.lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int e = tid + r * 256;
      int n, k;
      if (TB) { k = e & (BK - 1); n = e >> 4; } else { n = e & (BN - 1); k = e >> 6; }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < kend) v = TB ? B[(int64_t)gn * g.ldb + gk] : B[(int64_t)gk * g.ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* C = g.C + batch * g.strideC + (g.splits > 1 ? split * g.strideSplit : 0);
  const float* bias = g.bias ? g.bias + batch * g.strideBias : nullptr;
  const float* mask = g.mask ? g.mask + batch * g.strideMask : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      v = act_apply(v, g.act);
      if (mask) v *= act_grad_from_output(mask[(int64_t)gm * g.ldmask + gn], g.mask_act);
      C[(int64_t)gm * g.ldc + gn] = v;
    }
  }
}

// sum split-K partials in a fixed order (deterministic): out[b][i] = sum_s part[s][b][i]
__global__ void k_reduce_splits(const float* part, int splits, int64_t strideSplit, int64_t strideB,
                                float* out, int64_t strideOut, int count) {
  int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[k * strideSplit + b * strideB + i];
    out[b * strideOut + i] = s;
  }
}

// db[b][c] = sum_rows dZ[b][row][c]   (block = 32 columns x 8 row-lanes)
__global__ void k_colsum(const float* dz, int64_t strideB, int rows, int cols, float* out, int64_t strideOut) {
  __shared__ float sm[8][33];
  int b = blockIdx.y;
  int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < cols)
    for (int r = threadIdx.y; r < rows; r += 8) s += dz[b * strideB + (int64_t)r * cols + c];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    out[b * strideOut + c] = t;
  }
}

// loss + dZ_L.  One thread per (chain, row).  SPARSE_CE: logits in `out`; MSE: out = act(z).
// dZ = scale * d(mean loss)/dz.  loss partials are per block (double), reduced in a fixed order.
__global__ void k_loss(const float* out, int64_t strideB, int rows, int C, int loss_kind, int out_act,
                       const int32_t* y_i, const float* y_f, float scale, float* dz, double* loss_partial,
                       int write_dz) {
  __shared__ double scratch[32];
  int b = blockIdx.y;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  double li = 0.0;
  if (r < rows) {
    const float* o = out + b * strideB + (int64_t)r * C;
    float* d = dz + b * strideB + (int64_t)r * C;
    if (loss_kind == PYB_LOSS_SPARSE_CE) {
      float mx = -INFINITY;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, o[c]);
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += expf(o[c] - mx);
      int yi = y_i[r];
      float lse = logf(se);
      li = (double)(lse - (o[yi] - mx));
      if (write_dz) {
        float inv = 1.0f / se, sc = scale / (float)rows;
        for (int c = 0; c < C; ++c) d[c] = (expf(o[c] - mx) * inv - (c == yi ? 1.f : 0.f)) * sc;
      }
    } else {
      float acc = 0.f;
      float sc = scale * 2.0f / ((float)rows * (float)C);
      for (int c = 0; c < C; ++c) {
        float a = o[c];
        float df = a - y_f[(int64_t)r * C + c];
        acc += df * df;
        if (write_dz) d[c] = sc * df * act_grad_from_output(a, out_act);
      }
      li = (double)(acc / (float)C);
    }
  }
  double tot = block_sum<double>(li, scratch);
  if (threadIdx.x == 0) loss_partial[(int64_t)b * gridDim.x + blockIdx.x] = tot;
}

__global__ void k_loss_finish(const double* loss_partial, int nblk, int rows, float* loss_out) {
  __shared__ double scratch[32];
  int b = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += loss_partial[(int64_t)b * nblk + i];
  double tot = block_sum<double>(s, scratch);
  if (threadIdx.x == 0) loss_out[b] = (float)(tot / (double)rows);
}

__global__ void k_softmax_rows(float* z, int64_t total_rows, int C) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= total_rows) return;
  float* o = z + r * C;
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, o[c]);
  float se = 0.f;
  for (int c = 0; c < C; ++c) se += expf(o[c] - mx);
  float inv = 1.0f / se;
  for (int c = 0; c < C; ++c) o[c] = expf(o[c] - mx) * inv;
}

template <bool TA, bool TB>
static void launch_gemm(pyb_handle* h, const GemmArgs& g, int batch) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, batch * g.splits);
  prof_begin(h);
  k_sgemm<TA, TB><<<grid, 256, 0, h->stream>>>(g);
  prof_end(h, 2.0 * g.M * g.N * (double)g.K * batch);
  count_launch(h);
}

int64_t generic_chain_batch(pyb_handle* h, int64_t S, int64_t N, bool backward) {
  const Model& m = h->model;
  int64_t widths = 0;
  for (int l = 0; l < m.n_layers; ++l) widths += m.layer[l].fan_out;
  int64_t per_chain = N * (widths + (backward ? 2 * (int64_t)m.max_width : 0)) * 4;
  int64_t budget = (int64_t)(h->opt_workspace_mb * 1024.0 * 1024.0);
  int64_t bc = budget / (per_chain > 0 ? per_chain : 1);
  if (bc < 1) bc = 1;
  if (h->opt_chain_batch > 0 && bc > h->opt_chain_batch) bc = h->opt_chain_batch;
  if (bc > S) bc = S;
  // gridDim.z limit
  if (bc > 16384) bc = 16384;
  return bc;
}

static void ensure_ws(pyb_handle* h, int64_t Bc, int64_t N, bool backward) {
  Workspace& ws = h->ws;
  const Model& m = h->model;
  while ((int)ws.act.size() < m.n_layers) ws.act.push_back(new DevBuf<float>());
  for (int l = 0; l < m.n_layers; ++l) ws.act[l]->alloc((size_t)Bc * N * m.layer[l].fan_out);
  if (backward) {
    ws.dz_a.alloc((size_t)Bc * N * m.max_width);
    ws.dz_b.alloc((size_t)Bc * N * m.max_width);
  }
  int nblk = (int)((N + 255) / 256);
  ws.loss_partial.alloc((size_t)Bc * nblk);
  ws.Bc = Bc;
  ws.N = N;
}

static void forward_batch(pyb_handle* h, const float* theta, int64_t nb, const float* X, int64_t N) {
  const Model& m = h->model;
  Workspace& ws = h->ws;
  const float* a_prev = X;
  int64_t strideA = 0;
  for (int l = 0; l < m.n_layers; ++l) {
    const LayerDesc& L = m.layer[l];
    GemmArgs g = {};
    g.M = (int)N; g.N = L.fan_out; g.K = L.fan_in;
    g.A = a_prev; g.lda = L.fan_in; g.strideA = strideA;
    g.B = theta + L.w_off; g.ldb = L.fan_out; g.strideB = m.P;
    g.C = ws.act[l]->p; g.ldc = L.fan_out; g.strideC = N * L.fan_out;
    g.bias = L.use_bias ? theta + L.b_off : nullptr; g.strideBias = m.P;
    g.act = (L.act == PYB_ACT_SOFTMAX) ? PYB_ACT_LINEAR : L.act;
    g.splits = 1; g.kchunk = L.fan_in;
    launch_gemm<false, false>(h, g, (int)nb);
    a_prev = ws.act[l]->p;
    strideA = N * L.fan_out;
  }
}

void generic_eval(pyb_handle* h, const float* theta, int64_t S, const float* X, const int32_t* y_i,
                  const float* y_f, int64_t N, float scale, float* loss_out, float* grad_out) {
  const Model& m = h->model;
  const bool backward = grad_out != nullptr;
  int64_t Bc = generic_chain_batch(h, S, N, backward);
  ensure_ws(h, Bc, N, backward);
  Workspace& ws = h->ws;
  const LayerDesc& Llast = m.layer[m.n_layers - 1];
  const int C = Llast.fan_out;
  const int nblk = (int)((N + 255) / 256);
  for (int64_t b0 = 0; b0 < S; b0 += Bc) {
    int64_t nb = (S - b0 < Bc) ? (S - b0) : Bc;
    const float* th = theta + b0 * m.P;
    forward_batch(h, th, nb, X, N);
    float* dz_cur = ws.dz_a.p;
    float* dz_nxt = ws.dz_b.p;
    {
      dim3 grid(nblk, (unsigned)nb);
      k_loss<<<grid, 256, 0, h->stream>>>(ws.act[m.n_layers - 1]->p, N * C, (int)N, C, h->loss_kind, Llast.act,
                                          y_i, y_f, scale, backward ? dz_cur : ws.act[m.n_layers - 1]->p,
                                          ws.loss_partial.p, backward ? 1 : 0);
      count_launch(h);
      if (loss_out) {
        k_loss_finish<<<(unsigned)nb, 256, 0, h->stream>>>(ws.loss_partial.p, nblk, (int)N, loss_out + b0);
        count_launch(h);
      }
    }
    if (!backward) continue;
    float* gr = grad_out + b0 * m.P;
    for (int l = m.n_layers - 1; l >= 0; --l) {
      const LayerDesc& L = m.layer[l];
      const float* a_in = (l == 0) ? X : ws.act[l - 1]->p;
      int64_t strideAin = (l == 0) ? 0 : N * L.fan_in;
      // dW = a_in^T dZ
      {
        GemmArgs g = {};
        g.M = L.fan_in; g.N = L.fan_out; g.K = (int)N;
        g.A = a_in; g.lda = L.fan_in; g.strideA = strideAin;
        g.B = dz_cur; g.ldb = L.fan_out; g.strideB = N * L.fan_out;
        g.act = PYB_ACT_LINEAR;
        int tiles = ((L.fan_in + BM - 1) / BM) * ((L.fan_out + BN - 1) / BN);
        int64_t want = (2 * (int64_t)h->sm_count + tiles * nb - 1) / (tiles * nb);
        int splits = (int)(want < 1 ? 1 : (want > 64 ? 64 : want));
        int kchunk = (int)((N + splits - 1) / splits);
        kchunk = ((kchunk + BK - 1) / BK) * BK;
        splits = (int)((N + kchunk - 1) / kchunk);
        g.splits = splits; g.kchunk = kchunk;
        int64_t wsz = (int64_t)L.fan_in * L.fan_out;
        if (splits > 1) {
          ws.partial.alloc((size_t)splits * nb * wsz);
          g.C = ws.partial.p; g.ldc = L.fan_out; g.strideC = wsz; g.strideSplit = nb * wsz;
          launch_gemm<true, false>(h, g, (int)nb);
          dim3 rg((unsigned)((wsz + 255) / 256 > 1024 ? 1024 : (wsz + 255) / 256), (unsigned)nb);
          k_reduce_splits<<<rg, 256, 0, h->stream>>>(ws.partial.p, splits, nb * wsz, wsz, gr + L.w_off, m.P, (int)wsz);
          count_launch(h);
        } else {
          g.C = gr + L.w_off; g.ldc = L.fan_out; g.strideC = m.P;
          launch_gemm<true, false>(h, g, (int)nb);
        }
      }
      if (L.use_bias) {
        dim3 grid((L.fan_out + 31) / 32, (unsigned)nb), blk(32, 8);
        k_colsum<<<grid, blk, 0, h->stream>>>(dz_cur, N * L.fan_out, (int)N, L.fan_out, gr + L.b_off, m.P);
        count_launch(h);
      }
      if (l > 0) {
        // dA_prev = dZ W^T, masked by act'(a_prev)
        GemmArgs g = {};
        g.M = (int)N; g.N = L.fan_in; g.K = L.fan_out;
        g.A = dz_cur; g.lda = L.fan_out; g.strideA = N * L.fan_out;
        g.B = th + L.w_off; g.ldb = L.fan_out; g.strideB = m.P;
        g.C = dz_nxt; g.ldc = L.fan_in; g.strideC = N * L.fan_in;
        g.act = PYB_ACT_LINEAR;
        g.mask = ws.act[l - 1]->p; g.ldmask = L.fan_in; g.strideMask = N * L.fan_in;
        g.mask_act = m.layer[l - 1].act;
        g.splits = 1; g.kchunk = L.fan_out;
        launch_gemm<false, true>(h, g, (int)nb);
        float* t = dz_cur; dz_cur = dz_nxt; dz_nxt = t;
      }
    }
  }
  PYB_CUDA(cudaGetLastError());
}

void generic_forward(pyb_handle* h, const float* theta, int64_t S, const float* X, int64_t N, float* out) {
  const Model& m = h->model;
  int64_t Bc = generic_chain_batch(h, S, N, false);
  ensure_ws(h, Bc, N, false);
  Workspace& ws = h->ws;
  const int C = m.out_dim;
  for (int64_t b0 = 0; b0 < S; b0 += Bc) {
    int64_t nb = (S - b0 < Bc) ? (S - b0) : Bc;
    forward_batch(h, theta + b0 * m.P, nb, X, N);
    float* o = ws.act[m.n_layers - 1]->p;
    if (m.layer[m.n_layers - 1].act == PYB_ACT_SOFTMAX) {
      int64_t rows = nb * N;
      k_softmax_rows<<<(unsigned)((rows + 255) / 256), 256, 0, h->stream>>>(o, rows, C);
      count_launch(h);
    }
    PYB_CUDA(cudaMemcpyAsync(out + b0 * N * C, o, (size_t)nb * N * C * sizeof(float), cudaMemcpyDeviceToDevice,
                             h->stream));
  }
  PYB_CUDA(cudaGetLastError());
}

}  // namespace pyb

// sgmc.cu — S independent stochastic-gradient chains on one minibatch per step: SGLD and SWAG as S-batched variants
// of the same minibatch gradient kernels the SVGD step uses.
//   SGLD.step  (Pyesian/optimizers/SGLD.py:46-95):  theta -= lr_n * (g + noise), noise ~ N(0, stddev = lr_n)  (:67-68,
//              i.e. lr_n^2 * z — the reference's scaling, kept), then for EVERY step n the running first and second
//              moments  mean = (mean*n + theta)/(n+1),  sq = (sq*n + theta^2)/(n+1)  (:79-86).
//   SWAG.step  (Pyesian/optimizers/SWAG.py:43-94):  theta -= lr * g (:62-64); when n % frequency == 0 the same moment
//              updates WEIGHTED BY n (:75-80) and a column theta - mean of the deviation matrix: appended while it has
//              fewer than k columns, otherwise written over the LAST column (:83-89 keeps columns 0..k-2).
// One fused element pass per step (theta, g, mean, sq read; theta, mean, sq and one deviation column written: HBM
// bound, 32 B per parameter); the gradient is the shared eval_on_batch (tensor path when the shape allows).
#include "common.cuh"
#include <algorithm>
#include <math.h>

namespace pyb {

// Keras Dense defaults (what tf.keras.models.model_from_json builds, SGLD.py:138): glorot_uniform kernel
// U(-l, l), l = sqrt(6 / (fan_in + fan_out)), zero bias.  One Philox block gives 4 uniforms.
__global__ void k_sg_glorot(float* theta, int64_t P, int64_t off, int64_t count, float limit, uint64_t seed,
                            int64_t chain_offset, uint32_t layer) {
  const int64_t s = blockIdx.y;
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (base >= count) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)(base >> 2), (uint32_t)(chain_offset + s), layer, STREAM_INIT, (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  for (int j = 0; j < 4; ++j)
    if (base + j < count) {
      const float u = ((float)(r[j] >> 8) + 0.5f) * 5.9604644775390625e-08f;   // (0, 1)
      theta[s * P + off + base + j] = (2.f * u - 1.f) * limit;
    }
}
__global__ void k_sg_broadcast(float* theta, const float* src, int64_t P, int rows) {
  const int64_t s = blockIdx.y;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < P; e += (int64_t)gridDim.x * blockDim.x)
    theta[s * P + e] = src[(rows == 1 ? 0 : s) * P + e];
}

// the fused update over the FLAT [S*P] arrays (they are contiguous, so a thread's four elements are one 16-byte access
// whatever P is; they may straddle two chains or two Philox blocks).  dev_col = nullptr: no deviation column this step;
// moments = 0: parameters only.  All loads are issued before the first store.
__global__ void __launch_bounds__(256) k_sg_update(float* __restrict__ theta, const float* __restrict__ g,
                                                   float* __restrict__ mean, float* __restrict__ sq,
                                                   float* __restrict__ dev_col, int64_t dev_stride,
                                                   const float* __restrict__ inj, int64_t P, int64_t total, float lr,
                                                   float noise_scale, int kind, int moments, float n, uint64_t seed,
                                                   int64_t chain_offset, uint32_t iter) {
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (base >= total) return;
  const bool full = base + 4 <= total;
  float t[4], gg[4], m[4], q[4], z[4] = {0.f, 0.f, 0.f, 0.f};
  auto ld4 = [](const float* p, float* o) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  };
  auto st4 = [](float* p, const float* o) { *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]); };
  if (full) {
    ld4(theta + base, t);
    ld4(g + base, gg);
    if (moments) { ld4(mean + base, m); ld4(sq + base, q); }
    if (kind == PYB_SG_SGLD && inj) ld4(inj + base, z);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = base + j < total;
      t[j] = ok ? theta[base + j] : 0.f;
      gg[j] = ok ? g[base + j] : 0.f;
      m[j] = (ok && moments) ? mean[base + j] : 0.f;
      q[j] = (ok && moments) ? sq[base + j] : 0.f;
      if (kind == PYB_SG_SGLD && inj && ok) z[j] = inj[base + j];
    }
  }
  int64_t s = base / P, e = base - s * P;       // chain and element of the first of the four
  if (kind == PYB_SG_SGLD && !inj) {
    float zz[4];
    int64_t have_s = -1, have_b = -1;
    int64_t ss = s, ee = e;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if ((ee >> 2) != have_b || ss != have_s) {
        philox_normal4((uint32_t)(ee >> 2), (uint32_t)(chain_offset + ss), iter, STREAM_SGLD, seed, zz);
        have_b = ee >> 2; have_s = ss;
      }
      const int l = (int)(ee & 3);
      z[j] = l == 0 ? zz[0] : (l == 1 ? zz[1] : (l == 2 ? zz[2] : zz[3]));
      if (++ee == P) { ee = 0; ++ss; }
    }
  }
  // every product and sum is rounded on its own (no FMA contraction), like the eager float32 ops of the reference
  float d[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    // var.assign_add(-lr * (grad + noise)), noise = stddev * z with stddev = lr (SGLD.py:67-68); SWAG: assign_sub(lr * g)
    t[j] = (kind == PYB_SG_SGLD) ? __fadd_rn(t[j], __fmul_rn(-lr, __fadd_rn(gg[j], __fmul_rn(noise_scale, z[j]))))
                                 : __fsub_rn(t[j], __fmul_rn(lr, gg[j]));
    if (moments) {
      m[j] = __fdiv_rn(__fadd_rn(__fmul_rn(m[j], n), t[j]), n + 1.0f);
      q[j] = __fdiv_rn(__fadd_rn(__fmul_rn(q[j], n), __fmul_rn(t[j], t[j])), n + 1.0f);
      d[j] = __fsub_rn(t[j], m[j]);
    }
  }
  if (full) {
    st4(theta + base, t);
    if (moments) { st4(mean + base, m); st4(sq + base, q); }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (base + j < total) {
        theta[base + j] = t[j];
        if (moments) { mean[base + j] = m[j]; sq[base + j] = q[j]; }
      }
  }
  if (moments && dev_col) {          // column c of chain s sits at s * k * P + c * P: not 16-byte aligned in general
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (base + j < total) dev_col[s * dev_stride + e] = d[j];
      if (++e == P) { e = 0; ++s; }
    }
  }
}
__global__ void k_sg_loss_mean(const float* loss, int64_t S, double* out) {
  __shared__ double scratch[32];
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < S; i += blockDim.x) a += (double)loss[i];
  a = block_sum(a, scratch);
  if (threadIdx.x == 0) out[0] = a / (double)S;
}

void sg_init(pyb_handle* h, int64_t S, int64_t chain_offset, int kind, int k_dev, int frequency, const float* theta0,
             int theta0_rows) {
  PYB_REQUIRE(h->have_data, PYB_ERR_STATE, "dataset must be set first");
  PYB_REQUIRE(S > 0 && S <= 65535, PYB_ERR_INVALID, "S must be in [1, 65535]");
  PYB_REQUIRE(kind == PYB_SG_SGLD || kind == PYB_SG_SWAG, PYB_ERR_INVALID, "bad kind");
  PYB_REQUIRE(kind != PYB_SG_SWAG || (k_dev >= 2 && frequency >= 1), PYB_ERR_INVALID, "SWAG needs k >= 2 and frequency >= 1");
  PYB_REQUIRE(!theta0 || theta0_rows == 1 || theta0_rows == S, PYB_ERR_INVALID, "theta0 must have 1 or S rows");
  SgState& sg = h->sg;
  const Model& m = h->model;
  const int64_t P = m.P;
  sg.S = S; sg.offset = chain_offset; sg.kind = kind; sg.n = 0; sg.cols = 0;
  sg.k = (kind == PYB_SG_SWAG) ? k_dev : 0;
  sg.freq = (kind == PYB_SG_SWAG) ? frequency : 1;
  sg.theta.alloc(S * P); sg.g.alloc(S * P); sg.mean.alloc(S * P); sg.sq.alloc(S * P); sg.loss.alloc(S);
  sg.mean_loss.alloc(1);
  if (sg.k) sg.dev.alloc((size_t)S * sg.k * P);
  PYB_CUDA(cudaMemsetAsync(sg.mean.p, 0, S * P * sizeof(float), h->stream));
  PYB_CUDA(cudaMemsetAsync(sg.sq.p, 0, S * P * sizeof(float), h->stream));
  if (sg.k) PYB_CUDA(cudaMemsetAsync(sg.dev.p, 0, (size_t)S * sg.k * P * sizeof(float), h->stream));
  if (theta0) {
    DevBuf<float> tmp;
    tmp.alloc((size_t)theta0_rows * P);
    PYB_CUDA(cudaMemcpyAsync(tmp.p, theta0, (size_t)theta0_rows * P * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    dim3 grid((unsigned)std::min<int64_t>((P + 255) / 256, 1024), (unsigned)S);
    k_sg_broadcast<<<grid, 256, 0, h->stream>>>(sg.theta.p, tmp.p, P, theta0_rows);
    count_launch(h);
    PYB_CUDA(cudaStreamSynchronize(h->stream));
  } else {
    PYB_CUDA(cudaMemsetAsync(sg.theta.p, 0, S * P * sizeof(float), h->stream));
    for (int l = 0; l < m.n_layers; ++l) {
      const LayerDesc& L = m.layer[l];
      const int64_t count = (int64_t)L.fan_in * L.fan_out;
      const float limit = sqrtf(6.0f / (float)(L.fan_in + L.fan_out));
      dim3 grid((unsigned)((count + 1023) / 1024), (unsigned)S);
      k_sg_glorot<<<grid, 256, 0, h->stream>>>(sg.theta.p, P, L.w_off, count, limit, h->seed, chain_offset, (uint32_t)l);
      count_launch(h);
    }
    PYB_CUDA(cudaStreamSynchronize(h->stream));
  }
  PYB_CUDA(cudaGetLastError());
  sg.inited = true;
}

void sg_step(pyb_handle* h, const int32_t* idx, int64_t B, double lr, const float* noise, float* loss_out,
             double* mean_loss_out) {
  SgState& sg = h->sg;
  PYB_REQUIRE(sg.inited, PYB_ERR_STATE, "pyb_sg_init must be called first");
  const Model& m = h->model;
  const int64_t P = m.P, S = sg.S;
  const float* Xb = h->X.p;
  const int32_t* yb_i = h->y_i.p;
  const float* yb_f = h->y_f.p;
  int64_t Nb = h->N;
  PYB_CUDA(cudaEventRecord(h->ev0, h->stream));
  if (idx) {
    PYB_REQUIRE(B > 0, PYB_ERR_INVALID, "B must be > 0 with batch_idx");
    sg.idx.alloc(B);
    sg.Xb.alloc(B * m.in_dim);
    if (h->loss_kind == PYB_LOSS_SPARSE_CE) sg.yb_i.alloc(B); else sg.yb_f.alloc(B * m.out_dim);
    PYB_CUDA(cudaMemcpyAsync(sg.idx.p, idx, B * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    gather_batch(h, sg.idx.p, B, sg.Xb.p, sg.yb_i.p, sg.yb_f.p);
    Xb = sg.Xb.p; yb_i = sg.yb_i.p; yb_f = sg.yb_f.p; Nb = B;
  }
  DevBuf<float> inj;
  if (noise && sg.kind == PYB_SG_SGLD) {
    inj.alloc(S * P);
    PYB_CUDA(cudaMemcpyAsync(inj.p, noise, S * P * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  }
  // mean minibatch loss and its gradient for every chain (SGLD.py:54-65, SWAG.py:52-61)
  eval_on_batch(h, sg.theta.p, S, Xb, yb_i, yb_f, Nb, 1.0f, sg.loss.p, sg.g.p);
  const int moments = (sg.n % sg.freq == 0) ? 1 : 0;
  float* dev_col = nullptr;
  if (moments && sg.k) {
    const int col = (sg.cols == sg.k) ? sg.k - 1 : sg.cols;      // full: the last column is replaced (SWAG.py:84-86)
    dev_col = sg.dev.p + (int64_t)col * P;
    if (sg.cols < sg.k) sg.cols += 1;
  }
  const int64_t total = S * P;
  k_sg_update<<<(unsigned)((total + 1023) / 1024), 256, 0, h->stream>>>(
      sg.theta.p, sg.g.p, sg.mean.p, sg.sq.p, dev_col, (int64_t)sg.k * P, inj.p, P, total, (float)lr, (float)lr, sg.kind,
      moments, (float)sg.n, h->seed, sg.offset, (uint32_t)sg.n);
  count_launch(h);
  k_sg_loss_mean<<<1, 256, 0, h->stream>>>(sg.loss.p, S, sg.mean_loss.p);
  count_launch(h);
  sg.n += 1;
  PYB_CUDA(cudaEventRecord(h->ev1, h->stream));
  double ml = 0.0;
  PYB_CUDA(cudaMemcpyAsync(&ml, sg.mean_loss.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (loss_out) PYB_CUDA(cudaMemcpyAsync(loss_out, sg.loss.p, S * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  PYB_CUDA(cudaStreamSynchronize(h->stream));
  PYB_CUDA(cudaGetLastError());
  if (mean_loss_out) *mean_loss_out = ml;
  float ms = 0.f;
  PYB_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_device_ms = ms;
}

void sg_get(pyb_handle* h, float* theta, float* mean, float* sq, float* dev, int32_t* n_cols, int64_t* n_steps) {
  SgState& sg = h->sg;
  PYB_REQUIRE(sg.inited, PYB_ERR_STATE, "pyb_sg_init must be called first");
  const size_t bytes = (size_t)sg.S * h->model.P * sizeof(float);
  if (theta) PYB_CUDA(cudaMemcpyAsync(theta, sg.theta.p, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (mean) PYB_CUDA(cudaMemcpyAsync(mean, sg.mean.p, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (sq) PYB_CUDA(cudaMemcpyAsync(sq, sg.sq.p, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (dev && sg.k) PYB_CUDA(cudaMemcpyAsync(dev, sg.dev.p, bytes * sg.k, cudaMemcpyDeviceToHost, h->stream));
  PYB_CUDA(cudaStreamSynchronize(h->stream));
  if (n_cols) *n_cols = sg.cols;
  if (n_steps) *n_steps = sg.n;
}

}  // namespace pyb

/* pyesian_b200.h — C ABI of libpyesian_b200.so
 *
 * B200-native (sm_100a) implementation of ONE hot path of leoelm/Bayesian_inference_for_NN
 * ("Pyesian"): the particle-batched log-posterior forward/backward of a Keras Dense MLP over the
 * whole dataset for S weight samples at once, fused with the HMC leapfrog / Hamiltonian /
 * Metropolis update and the SVGD Stein step, plus the BayesianModel posterior predictive.
 *
 * The reference has no FFI layer of its own (it is pure Python on TensorFlow eager); the entry
 * points below are what a ctypes binding inside the reference's optimizer classes would bind.
 * Each one cites the reference code (path:line under the reference checkout) it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types, no exceptions across the boundary.
 *   - every call returns 0 (PYB_OK) or a negative pyb_status; pyb_last_error() gives the message
 *     (thread-local).
 *   - the library owns all device memory behind the handle; caller-owned buffers are only read
 *     or written for the duration of the call.  `mem` says whether a caller buffer is host or
 *     device memory (device pointers come from DLPack capsules on the Python side).
 *   - device pointers: the library works on its own non-blocking stream, which is ordered against no other stream.  Every
 *     entry point that READS caller-owned device memory (pyb_set_dataset with PYB_MEM_DEVICE, W / x of pyb_predict and
 *     pyb_predict_uncertainty, src / dst of pyb_gather_rows) first waits for the whole device (cudaDeviceSynchronize), so a
 *     tensor produced on any stream just before the call is complete when it is read; results the library WRITES to
 *     caller-owned device memory are complete when the call returns (every call ends with a stream synchronisation).
 *   - sharded runs (pyb_svgd_set_comm / pyb_set_comm): every rank must hold the same number of particles; the first
 *     sharded step checks it (one all-reduce) and fails with PYB_ERR_INVALID instead of hanging in a collective.
 *   - one handle = one GPU; a handle is not thread-safe; there is NO CPU fallback: without a
 *     usable sm_100 device pyb_create fails with PYB_ERR_CUDA.
 *   - flat parameter order everywhere = model.layers order -> (kernel [in,out] C-order, bias)
 *     (HMC.py:178-183, SVGD.py:159-160,230-239, BayesianModel.py:73-77).
 */
#ifndef PYESIAN_B200_H
#define PYESIAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYB_ABI_VERSION 1

typedef enum {
  PYB_OK = 0,
  PYB_ERR_INVALID = -1,     /* bad argument */
  PYB_ERR_CUDA = -2,        /* CUDA runtime/driver failure, or no sm_100 device */
  PYB_ERR_STATE = -3,       /* call made in the wrong state (e.g. run before init) */
  PYB_ERR_UNSUPPORTED = -4, /* model/loss combination outside the hot path */
  PYB_ERR_OOM = -5          /* device memory */
} pyb_status;

typedef enum { PYB_ACT_LINEAR = 0, PYB_ACT_RELU = 1, PYB_ACT_SOFTMAX = 2, PYB_ACT_TANH = 3,
               PYB_ACT_SIGMOID = 4 } pyb_activation;
typedef enum { PYB_LOSS_SPARSE_CE = 0, PYB_LOSS_MSE = 1 } pyb_loss;
typedef enum { PYB_MEM_HOST = 0, PYB_MEM_DEVICE = 1 } pyb_mem;
typedef enum { PYB_PRIOR_SCALAR = 0, PYB_PRIOR_PER_VARIABLE = 1, PYB_PRIOR_PER_ELEMENT = 2 } pyb_prior_form;
typedef enum { PYB_HMC_REFERENCE = 0, PYB_HMC_CANONICAL = 1 } pyb_hmc_semantics;
typedef enum { PYB_SVGD_REFERENCE_LIVE = 0, PYB_SVGD_CANONICAL_MEDIAN = 1 } pyb_svgd_semantics;
/* which device path evaluates the MLP; AUTO picks by shape */
typedef enum { PYB_PATH_AUTO = 0, PYB_PATH_GENERIC = 1, PYB_PATH_FUSED_SMALL = 2, PYB_PATH_TENSOR = 3 } pyb_path;

typedef enum { PYB_SG_SGLD = 0, PYB_SG_SWAG = 1 } pyb_sg_kind;
typedef enum { PYB_UQ_CANONICAL = 0, PYB_UQ_REFERENCE = 1 } pyb_uq_semantics;

typedef struct pyb_handle pyb_handle;

/* Dense stack parsed from the Keras model JSON (model.to_json(); consumed at HMC.py:56,
 * SVGD.py:222, BayesianModel.py:18).  The JSON parse itself is host logic on the Python side. */
typedef struct {
  int32_t n_layers;
  int32_t in_dim;
  const int32_t* units;      /* [n_layers] */
  const int32_t* activation; /* [n_layers] pyb_activation */
  const int32_t* use_bias;   /* [n_layers] 0/1 */
} pyb_model_desc;

typedef struct {
  double mean_loss;      /* mean over chains of the loss step() would return (HMC.py:96,104) */
  double accept_rate;    /* accepted / total since the phase started (HMC.py:112,122) */
  int64_t n_accepted;    /* over all local chains and iterations of this call */
  int64_t n_total;
  int64_t n_nan;         /* proposals whose Hamiltonian difference was NaN (rejected) */
  int64_t grad_evals;    /* chain x position evaluations actually executed in this call (L per iteration and chain,
                            plus one when the carried evaluation at the start position is not available) */
  double device_ms;      /* CUDA-event time of the call's kernels on the handle's stream */
  int64_t kernel_launches;
} pyb_hmc_diag;

/* ---- library ---- */
int pyb_version(void);
const char* pyb_last_error(void);
int pyb_device_count(int32_t* n_out);

/* ---- handle / model (replaces tf.keras.models.model_from_json + variable creation,
 *      HMC.py:56, SVGD.py:222, BayesianModel.py:18) ---- */
int pyb_create(const pyb_model_desc* desc, int32_t device_id, uint64_t seed, pyb_handle** out);
int pyb_destroy(pyb_handle* h);
int pyb_param_count(const pyb_handle* h, int64_t* n_params_out);
/* knobs: "path" (pyb_path), "workspace_mb", "chain_batch", "profile" (per-launch CUDA-event timing of the
 * tensor-core kernels, read back with "prof_ms"/"prof_flops"/"prof_launches"), and three A/B switches of the tensor
 * path, all default 1: "tc_pair" (CTA-pair cta_group::2 kernels), "tc_fuse" (layer 2, loss and both deltas inside
 * the layer-1 GEMM's epilogue), "tc_dual" (two feature tiles per item in the dW1 GEMM).  Switching them off selects
 * the older kernels that compute the same quantities (used by the tests and by tools/kernel_cycles.sh);
 * "predict_sharded" (default 0, see pyb_set_comm); "hmc_carry" (default 1): the loss and loss gradient at a chain's
 * current position are carried from the previous iteration (its end point if accepted, its start if rejected) instead
 * of being re-evaluated as HMC.py:80,82 do - L instead of L+1 full-data evaluations per iteration, identical results. */
int pyb_set_option(pyb_handle* h, const char* key, double value);
/* read-outs: "path_used", "kernel_launches", "last_device_ms", "workspace_bytes", "tensor_path_ok" */
int pyb_get_info(const pyb_handle* h, const char* key, double* value_out);

/* ---- inputs ---- */
/* Full training batch kept resident in HBM (HMC.py:63-65: ONE full-dataset batch frozen for the
 * run; SVGD.py:220-221 draws minibatches from the same pool).  X [N, in_dim] float32 row-major;
 * y int32 [N] (SPARSE_CE) or float32 [N, out_dim] (MSE).  n_train multiplies the mean loss in the
 * potential (HMC.py:158); pass N (or <=0) for the reference behaviour. */
int pyb_set_dataset(pyb_handle* h, const float* X, int64_t N, const void* y, int32_t loss_kind,
                    int32_t mem, int64_t n_train);
/* GaussianPrior.get_model_priors (GaussianPrior.py:100-121, :28-47): sigma = rho used RAW.
 * SCALAR: 1 value each; PER_VARIABLE: one per trainable variable (kernel, bias, kernel, ...);
 * PER_ELEMENT: n_params values.  Host pointers. */
int pyb_set_prior_gaussian(pyb_handle* h, const float* mean, const float* sigma, int32_t form);

/* ---- HMC (HMC.py:45-72 compile, :74-104 step, :106-126 train, :128-171 helpers) ---- */
/* S local chains whose global ids are chain_offset .. chain_offset+S-1 (RNG counters use the
 * global id, so a sharded run reproduces the unsharded one).  q0 NULL => every chain starts at
 * the prior mean (HMC.py:69-72); else host float32 [S, P]. */
int pyb_hmc_init(pyb_handle* h, int64_t S, int64_t chain_offset, double epsilon, double m, int32_t L,
                 int32_t semantics, const float* q0);
/* test hook: momenta [S,P] and/or uniforms [S] (host) for the NEXT iteration only. */
int pyb_hmc_inject(pyb_handle* h, const float* p, const float* u);
/* n_iters iterations of HMC.step(sampling, burning) for all chains, no host sync inside. */
int pyb_hmc_run(pyb_handle* h, int32_t n_iters, int32_t burning, int32_t sampling, pyb_hmc_diag* diag_out);
/* parity hook: U = -log prior + n_train*mean_loss, the mean loss, and dU/dq for S given
 * positions (HMC._potential_energy HMC.py:149-159 and the gradient _step_p takes :128-136).
 * All host pointers; U/loss/grad may be NULL. */
int pyb_hmc_eval(pyb_handle* h, const float* q, int64_t S, float* U_out, float* loss_out, float* grad_out);
int pyb_hmc_get_state(pyb_handle* h, float* q_out, float* p_out);
/* per-chain values of the LAST iteration (any may be NULL). */
int pyb_hmc_last(pyb_handle* h, float* U0, float* K0, float* U1, float* K1, float* log_alpha,
                 int32_t* accepted, float* loss);
/* samples + frequencies of every chain (HMC.py:75-77,92-104; HMC.result :176-184), chain-major,
 * within a chain in acceptance order. */
int pyb_hmc_reset_samples(pyb_handle* h);
int pyb_hmc_sample_count(pyb_handle* h, int64_t* n_out);
int pyb_hmc_samples(pyb_handle* h, float* samples_out, int32_t* freq_out, int32_t* chain_out);

/* ---- SVGD (SVGD.py:219-228 compile, :143-157 init, :84-141 step, :54-68 + :183-202 live kernel,
 *      :165-181 median-heuristic kernel, legacy Adam created :226 applied :120) ---- */
/* particles0 NULL => one prior sample per element (Philox); else host float64 [S, P]. */
int pyb_svgd_init(pyb_handle* h, int64_t S, int64_t particle_offset, double lr, int32_t semantics,
                  const double* particles0);
/* One step on the minibatch given by row indices into the resident dataset (host int32 [B]);
 * batch_idx NULL => the full dataset.  loss_out = mean over particles of the minibatch loss. */
int pyb_svgd_step(pyb_handle* h, const int32_t* batch_idx, int64_t B, double* loss_out);
/* parity hook: phi for given particles X [S,P] (float64) and gradients G [S,P] (float32):
 * CANONICAL_MEDIAN: phi=(K G + dxkxy)/S with the median bandwidth, h_out = h.
 * REFERENCE_LIVE : row i = ((sum_k K_ik) G_i + 2 sum_k K_ik (x_i-x_k))/S with gamma=1 (Jacobi
 * evaluation of the formula the live sweep applies row by row). */
int pyb_svgd_phi(pyb_handle* h, const double* X, const float* G, int64_t S, int32_t semantics,
                 float* phi_out, double* h_out);
int pyb_svgd_get_particles(pyb_handle* h, double* particles_out);
/* The validation pass of SVGD.step (SVGD.py:126-129: every particle's loss over the whole validation split, summed
 * / M) without moving the particles: the held-out set (host X [N, in_dim] float32, y as in pyb_set_dataset) is uploaded
 * once, pyb_svgd_validation_loss evaluates the mean loss of every resident particle on it (per_particle_out [S] host,
 * may be NULL) and returns their mean (all-reduced over the ranks of a sharded run). */
int pyb_svgd_set_validation(pyb_handle* h, const float* X, const void* y, int64_t N);
int pyb_svgd_validation_loss(pyb_handle* h, double* mean_loss_out, float* per_particle_out);
/* NCCL plumbing for sharded particles: all ranks pass the same 128-byte ncclUniqueId.
 * Options of the sharded canonical step on the tensor path (all ranks must set them alike): "svgd_pshard" (default 1:
 * Stein phase sharded over the parameters; 0: row-sharded with all-gathers), "svgd_p2p" (default 1: the gradient rows and
 * the updated particle blocks are exchanged by stores of the library's own kernels into peer memory - CUDA IPC between
 * processes of one host, plain peer access inside one process - with one-float NCCL barriers; falls back to ncclSend /
 * ncclRecv when a mapping is refused; 0: always NCCL; read-out "svgd_p2p" tells which one runs), "svgd_gram_sync" and
 * "svgd_halves" (A/B switches of the pipeline, default 0), "svgd_chain_fused" (default 1: the reduction / median /
 * kernel-matrix chain in half the launches), "select_compact" (median radix select of a local set: 0 default, 1 with
 * candidate compaction, 2 with every digit picked inside the next pass), "live_cta" (default 1: the reference-live sweep of a particle
 * set that fits one CTA's shared memory runs in one CTA); read-outs "svgd_phase_ms_0..6" (with "profile" on): CUDA-event
 * split of the last sharded step. */
int pyb_svgd_set_comm(pyb_handle* h, int32_t rank, int32_t world, const void* nccl_unique_id_128);
/* The same communicator under its general name.  With option "predict_sharded" = 1, pyb_predict and
 * pyb_predict_uncertainty treat W as this rank's share of the weight samples (BayesianModel.predict's nb_samples
 * draws split over the ranks) and all-reduce their moment sums: every rank returns the mean / variance / uncertainty
 * matrices over ALL ranks' samples; all_out stays local.  All ranks must make the call. */
int pyb_set_comm(pyb_handle* h, int32_t rank, int32_t world, const void* nccl_unique_id_128);
int pyb_nccl_unique_id(void* out_128);

/* ---- S-batched stochastic-gradient chains: SGLD (SGLD.py:46-95 step, :115-121 schedule on the host, :133-143
 *      compile) and SWAG (SWAG.py:43-94 step, :97-113 compile) on the same minibatch gradient kernels ----
 * S local chains with global ids chain_offset.. (Philox counters use the global id).  theta0 NULL => Keras Dense
 * defaults per chain (glorot_uniform kernels, zero biases: what model_from_json builds, SGLD.py:138); else host
 * float32 [theta0_rows, P] with theta0_rows = 1 (every chain starts from the same weights: SWAG's starting_model,
 * SWAG.py:105-106) or S.  k_dev / frequency: SWAG's deviation-matrix width and moment update period (ignored by SGLD,
 * whose moments are updated every step and whose deviation matrix is never read, SGLD.py:146-161). */
int pyb_sg_init(pyb_handle* h, int64_t S, int64_t chain_offset, int32_t kind, int32_t k_dev, int32_t frequency,
                const float* theta0, int32_t theta0_rows);
/* One step of every chain on the minibatch batch_idx (host int32 [B] row indices into the resident dataset; NULL =>
 * the full dataset).  SGLD: theta -= lr * (g + lr * z) (the reference draws its noise with stddev = lr and multiplies
 * by lr again, SGLD.py:67-68); noise = host float32 [S, P] standard normals injected for this step (test hook) or
 * NULL => Philox.  SWAG: theta -= lr * g.  Then the running moments / deviation column as the reference updates
 * them.  loss_out [S] (host, may be NULL) = each chain's minibatch loss BEFORE the update; mean_loss_out = their mean. */
int pyb_sg_step(pyb_handle* h, const int32_t* batch_idx, int64_t B, double lr, const float* noise, float* loss_out,
                double* mean_loss_out);
/* State read-back (any pointer may be NULL): theta, mean, sq_mean host float32 [S, P]; dev [S, k_dev, P] (column c of
 * chain s contiguous; columns >= *n_cols are zero); n_cols = deviation columns filled; n_steps = steps taken. */
int pyb_sg_get(pyb_handle* h, float* theta, float* mean, float* sq_mean, float* dev, int32_t* n_cols, int64_t* n_steps);

/* ---- posterior predictive (BayesianModel.predict BayesianModel.py:106-129; Plotter.py:244) ----
 * W [n,P] weight samples, weight [n] or NULL (=1), x [Nt,in_dim]; mean/var [Nt,out_dim] with
 * NaN->0 per element and population variance; all_out [n,Nt,out_dim] or NULL.  weight and the outputs are host
 * pointers; W and x may be host OR device pointers (device memory is read in place: weight samples that already
 * live in HBM skip the n*P*4-byte upload). */
int pyb_predict(pyb_handle* h, const float* W, int64_t n, const float* weight, const float* x,
                int64_t Nt, float* mean_out, float* var_out, float* all_out);

/* Metrics.classification_uncertainty (Metrics.py:344-375): for the same n weight samples and inputs, with integer
 * labels y [Nt] (host), per data row the matrices  aleatoric = sum_k w_k (diag(p_k) - p_k p_k^T)  and epistemic, divided
 * by `divisor` (the reference divides by its n_samples ARGUMENT, :368-369).
 *   PYB_UQ_REFERENCE: what the reference's code computes (pinned by goldens produced by the reference method itself,
 *     tests/golden/reference_metrics.npz): its accumulators are never reset between rows, so row r holds the running sum
 *     over rows 0..r (:352-366), and its epistemic term is D D^T with D = reshape(p, (-1,1)) - one_hot(label) BROADCAST
 *     to D_ij = p_i - onehot_j (:362-363), i.e. C p p^T - p 1^T - 1 p^T + 1 1^T, independent of the label.
 *   PYB_UQ_CANONICAL: per-row matrices with epistemic = sum_k w_k (p_k - onehot(y)) (p_k - onehot(y))^T.
 * Outputs are host float32 [Nt, Ce, Ce] with Ce = out_dim, or 2 for a one-unit output widened to [1-p, p] (:357-359);
 * total = epistemic + aleatoric; mean_out [Nt, out_dim] may be NULL.  At most 32 classes. */
int pyb_predict_uncertainty(pyb_handle* h, const float* W, int64_t n, const float* weight, const float* x, int64_t Nt,
                            const int32_t* y, int32_t semantics, double divisor, float* total_out,
                            float* aleatoric_out, float* epistemic_out, float* mean_out);

/* ---- device-resident arrays owned by the caller (posterior samples kept in HBM between predict calls:
 *      BayesianModel.predict re-draws nb_samples weight vectors from the SAME Sampled on every call,
 *      BayesianModel.py:106-129 / Sampled.py:29-32) ----
 * pyb_buffer_create allocates `bytes` on the handle's device and, if host_or_null is given, uploads them;
 * pyb_gather_rows writes dst[k] = src[idx[k]] for rows of row_len floats (src, dst device pointers, idx on the host);
 * buffers outlive nothing: free them with pyb_buffer_destroy before pyb_destroy. */
int pyb_buffer_create(pyb_handle* h, const void* host_or_null, int64_t bytes, void** dev_out);
int pyb_buffer_destroy(pyb_handle* h, void* dev);
int pyb_gather_rows(pyb_handle* h, const float* src, const int64_t* idx, int64_t n, int64_t row_len, float* dst);

/* ---- diagnostic entry points (used by tests/, never by the host classes) ----
 * pyb_debug_tc_gemm: D[M, Nn] = A[M, K] B[Nn, K]^T through the tcgen05 bf16x3 GEMM kernel (host pointers; Nn % 16 == 0,
 *   Nn <= 256, K % 8 == 0): the unit test of the kernel every tensor-path GEMM is built from.
 * pyb_debug_relu_mask: relu'(z1) exactly as the LAST tensor-path evaluation on the resident dataset used it, for one
 *   chain of its chain batch: mask_out [N, H] uint8 on the host.  HMC._step_p (HMC.py:128-136) differentiates through
 *   Keras' relu, whose derivative is discontinuous at 0; two correct float32 implementations disagree about the units
 *   whose pre-activation lies within rounding of zero, so the full-size parity test evaluates the float64 oracle with
 *   the device's own mask (tests/test_gpu_fullsize.py). */
int pyb_debug_tc_gemm(pyb_handle* h, const float* A, const float* B, int32_t M, int32_t Nn, int32_t K, float* D);
int pyb_debug_relu_mask(pyb_handle* h, int64_t chain, uint8_t* mask_out);

#ifdef __cplusplus
}
#endif
#endif /* PYESIAN_B200_H */

"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Tolerances are BASELINE.json's: log-prob and gradients within 1e-4 relative (fp32), leapfrog
trajectories within 1e-3 after L steps, identical accept/reject except where |log alpha - log u| is
below 1e-5, posterior predictive mean/variance within Monte-Carlo error (exact for injected draws).
"""
import numpy as np
import pytest

from conftest import load_golden, rel_err, spec_from_golden

pytestmark = pytest.mark.gpu

from bayesian_inference_for_nn_b200 import _lib, keras_json  # noqa: E402
from bayesian_inference_for_nn_b200.engine import Engine  # noqa: E402

TOL_LOGP = 1e-4     # relative, log-prob / loss / energies
TOL_GRAD = 1e-4     # relative (norm-wise), gradients
TOL_TRAJ = 1e-3     # relative (norm-wise), q_L / p_L after L leapfrog steps
PATHS = {"generic": _lib.PATH_GENERIC, "auto": _lib.PATH_AUTO}


def make_engine(g, O, path="generic", seed=0):
    spec = spec_from_golden(g, O)
    js = keras_json.make_sequential_json(spec.in_dim, spec.units, spec.acts, spec.use_bias)
    eng = Engine(keras_json.parse_model_json(js), seed=seed)
    eng.set_option("path", PATHS[path])
    eng.set_dataset(g["X"], g["y"], int(g["loss_kind"]))
    eng.set_prior([0.0], [float(g["sigma"])], _lib.PRIOR_SCALAR)
    return eng, spec


GOLD = ["hmc_c1_mini", "hmc_c1_canonical", "hmc_c3_mini", "hmc_regression", "hmc_deep_mse"]


@pytest.mark.parametrize("path", ["generic", "auto"])
@pytest.mark.parametrize("name", GOLD)
def test_logprob_and_gradient_match_golden(oracle, name, path):
    g = load_golden(name)
    eng, _ = make_engine(g, oracle, path)
    U, loss, grad = eng.hmc_eval(g["q"])
    np.testing.assert_allclose(U, g["U"], rtol=TOL_LOGP)
    np.testing.assert_allclose(loss, g["loss"], rtol=TOL_LOGP)
    for s in range(grad.shape[0]):
        assert rel_err(grad[s], g["grad"][s]) < TOL_GRAD


@pytest.mark.parametrize("path", ["generic", "auto"])
@pytest.mark.parametrize("name", GOLD)
def test_hmc_iteration_with_injected_randomness(oracle, name, path):
    g = load_golden(name)
    eng, _ = make_engine(g, oracle, path)
    S = int(g["S"])
    eng.hmc_init(S, float(g["eps"]), float(g["m"]), int(g["L"]), int(g["semantics"]), q0=g["q"])
    eng.hmc_inject(p=g["p"], u=g["u"])
    d = eng.hmc_run(1, burning=False, sampling=True)
    last = eng.hmc_last()
    for k in ("U0", "K0", "U1", "K1"):
        np.testing.assert_allclose(last[k], g[k], rtol=TOL_LOGP, err_msg=k)
    # decisions identical except inside the 1e-5 band around log u
    la, lu = g["log_alpha"], np.log(np.maximum(g["u"].astype(np.float64), 1e-300))
    decisive = np.abs(la - lu) > 1e-5
    np.testing.assert_array_equal(last["accept"][decisive], g["accept"][decisive])
    assert np.all(np.abs(last["log_alpha"] - la) <= 2e-4 * np.maximum(np.abs(g["U0"]), 1.0))
    q, p = eng.hmc_state()
    same = last["accept"] == g["accept"]
    for s in np.where(same)[0]:
        assert rel_err(q[s], g["q_out"][s]) < TOL_TRAJ
        assert rel_err(p[s], g["pL"][s]) < TOL_TRAJ
    assert d["n_total"] == S and d["n_accepted"] == int(last["accept"].sum())
    assert d["grad_evals"] == S * (int(g["L"]) + 1) and d["kernel_launches"] > 0
    # bookkeeping: first sampling iteration seeds [q0] and appends accepted states (HMC.py:75-77,92-96)
    samples, freq, chain = eng.hmc_samples()
    exp_n = S + int(last["accept"].sum())
    assert samples.shape[0] == exp_n and int(freq.sum()) == 2 * S
    k = 0
    for s in range(S):
        assert chain[k] == s
        np.testing.assert_array_equal(samples[k], g["q"][s])
        if last["accept"][s]:
            assert freq[k] == 1 and freq[k + 1] == 1 and chain[k + 1] == s
            np.testing.assert_array_equal(samples[k + 1], q[s])
            k += 2
        else:
            assert freq[k] == 2
            k += 1


@pytest.mark.parametrize("path", ["generic", "auto"])
def test_trajectory_endpoints_before_accept(oracle, path):
    """q_L itself (not only the post-accept state): force acceptance by burning."""
    g = load_golden("hmc_c1_mini")
    eng, _ = make_engine(g, oracle, path)
    eng.hmc_init(int(g["S"]), float(g["eps"]), float(g["m"]), int(g["L"]), int(g["semantics"]), q0=g["q"])
    eng.hmc_inject(p=g["p"], u=g["u"])
    d = eng.hmc_run(1, burning=True, sampling=False)
    q, p = eng.hmc_state()
    assert d["accept_rate"] == 1.0
    assert rel_err(q, g["qL"]) < TOL_TRAJ and rel_err(p, g["pL"]) < TOL_TRAJ


@pytest.mark.parametrize("path", ["generic", "auto"])
def test_device_rng_matches_philox_restatement(oracle, path):
    O = oracle
    g = load_golden("hmc_c1_mini")
    seed, S, m = 0x1234ABCD5678, 3, 0.5
    eng, spec = make_engine(g, O, path, seed=seed)
    eng.hmc_init(S, 1e-3, m, 1, _lib.HMC_REFERENCE, chain_offset=5)
    eng.hmc_run(2, burning=True, sampling=False)       # iteration counters 0 and 1
    # momenta of the third iteration, read back after a zero-step-size iteration (p unchanged by kicks)
    eng2, _ = make_engine(g, O, path, seed=seed)
    eng2.hmc_init(S, 0.0, m, 1, _lib.HMC_REFERENCE, chain_offset=5)
    eng2.hmc_run(1, burning=True, sampling=False)
    _, p = eng2.hmc_state()
    want = O.philox_normals(seed, np.arange(5, 5 + S), 0, O.STREAM_MOMENTUM, spec.n_params) * np.float32(m)
    np.testing.assert_allclose(p, want, rtol=2e-5, atol=2e-6)
    # canonical semantics draws with std sqrt(m)
    eng3, _ = make_engine(g, O, path, seed=seed)
    eng3.hmc_init(S, 0.0, m, 1, _lib.HMC_CANONICAL, chain_offset=5)
    eng3.hmc_run(1, burning=True, sampling=False)
    _, p3 = eng3.hmc_state()
    np.testing.assert_allclose(p3, want / np.float32(m) * np.float32(np.sqrt(m)), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("path", ["generic", "auto"])
def test_multi_iteration_run_matches_oracle_with_device_rng(oracle, path):
    """Whole phase (several iterations, device Philox momenta and uniforms) against the oracle fed
    the restated streams: states, accept decisions and the Sampled bookkeeping."""
    O = oracle
    g = load_golden("hmc_c1_mini")
    seed, S, n_it = 77, 5, 4
    eps, m, L = 2e-3, 1.0, 4
    eng, spec = make_engine(g, O, path, seed=seed)
    eng.hmc_init(S, eps, m, L, _lib.HMC_REFERENCE)
    mu, sg = O.expand_prior(spec, 0.0, float(g["sigma"]))
    prob = O.Problem(spec, g["X"], g["y"], int(g["loss_kind"]), mu, sg)
    q = np.zeros((S, spec.n_params), np.float32)         # chains start at the prior mean (HMC.py:69-72)
    book = O.SampleBook(S)
    acc_total = 0
    ambiguous = False
    for it in range(n_it):
        p = O.philox_normals(seed, np.arange(S), it, O.STREAM_MOMENTUM, spec.n_params) * np.float32(m)
        u = O.philox_uniforms(seed, np.arange(S), it)
        r = O.hmc_iteration(prob, q, p, u, eps, m, L, False, O.HMC_REFERENCE, np.float64)
        lu = np.log(np.maximum(u.astype(np.float64), 1e-300))
        ambiguous |= bool(np.any(np.abs(r["log_alpha"] - lu) < 1e-3))
        book.record(q, r["q"].astype(np.float32), r["accept"])
        q = r["q"].astype(np.float32)
        acc_total += int(r["accept"].sum())
    d = eng.hmc_run(n_it, burning=False, sampling=True)
    if ambiguous:
        pytest.skip("a decision fell inside the tolerance band for this seed")
    assert d["n_accepted"] == acc_total and d["n_total"] == S * n_it
    qd, _ = eng.hmc_state()
    assert rel_err(qd, q) < TOL_TRAJ
    samples, freq, chain = eng.hmc_samples()
    k = 0
    for s in range(S):
        for j, f in enumerate(book.freq[s]):
            assert chain[k] == s and freq[k] == f
            assert rel_err(samples[k], book.samples[s][j]) < TOL_TRAJ or np.abs(samples[k] - book.samples[s][j]).max() < 1e-6
            k += 1
    assert k == samples.shape[0]


@pytest.mark.parametrize("path", ["generic", "auto"])
def test_negative_sigma_rejects_everything_after_burn_in(oracle, path):
    """SURVEY B-1 / HMC_classification.py:50: GaussianPrior(0,-1) => NaN Hamiltonian."""
    g = load_golden("hmc_c1_mini")
    eng, _ = make_engine(g, oracle, path)
    eng.set_prior([0.0], [-1.0], _lib.PRIOR_SCALAR)
    eng.hmc_init(4, 0.005, 0.5, 5, _lib.HMC_REFERENCE)
    d = eng.hmc_run(3, burning=True, sampling=False)
    assert d["accept_rate"] == 1.0 and d["n_nan"] == 12
    q_burn, _ = eng.hmc_state()
    assert np.abs(q_burn).max() > 0
    d = eng.hmc_run(5, burning=False, sampling=True)
    assert d["n_accepted"] == 0 and d["n_nan"] == 20
    q_after, _ = eng.hmc_state()
    np.testing.assert_array_equal(q_after, q_burn)
    samples, freq, _ = eng.hmc_samples()
    assert samples.shape[0] == 4 and freq.tolist() == [6, 6, 6, 6]


@pytest.mark.parametrize("path", ["generic", "auto"])
def test_sharded_chains_reproduce_the_unsharded_run(oracle, path):
    """Chains never interact and RNG counters use the global chain id: running chains [0,6) in one
    handle or as [0,3)+[3,6) in two handles gives the same states."""
    g = load_golden("hmc_c1_mini")
    full, _ = make_engine(g, oracle, path, seed=9)
    full.hmc_init(6, 2e-3, 1.0, 3, _lib.HMC_REFERENCE)
    full.hmc_run(3, burning=False, sampling=True)
    qf, _ = full.hmc_state()
    parts = []
    for off in (0, 3):
        e, _ = make_engine(g, oracle, path, seed=9)
        e.hmc_init(3, 2e-3, 1.0, 3, _lib.HMC_REFERENCE, chain_offset=off)
        e.hmc_run(3, burning=False, sampling=True)
        parts.append(e.hmc_state()[0])
    np.testing.assert_allclose(np.concatenate(parts), qf, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("S", [1, 5, 30, 200])
def test_cluster_spread_small_chains_match_the_one_cta_kernel(oracle, S):
    """Few chains: the small-width kernel spreads each chain over a cluster of 8/4/2 CTAs (row ranges + DSMEM sum of
    the partial gradients).  Same trajectory as one CTA per chain up to summation order, same accept decisions, and
    the one-chain case still matches the float64 oracle."""
    O = oracle
    rng = np.random.default_rng(S)
    N, D, H, Cc, L, eps = 1600, 2, 50, 2, 6, 5e-3
    X = rng.standard_normal((N, D)).astype(np.float32)
    y = (X[:, 0] * X[:, 1] > 0).astype(np.int32)
    spec_o = O.MLPSpec(D, [H, Cc], ["relu", "softmax"])
    q = (rng.standard_normal((S, spec_o.n_params)) * 0.3).astype(np.float32)
    p0 = rng.standard_normal((S, spec_o.n_params)).astype(np.float32)
    u = rng.random(S).astype(np.float32)
    outs = []
    for cluster in (1, 0):
        eng = Engine(keras_json.parse_model_json(keras_json.make_sequential_json(D, [H, Cc], ["relu", "softmax"])))
        eng.set_option("fs_cluster", cluster)
        eng.set_dataset(X, y, _lib.LOSS_SPARSE_CE)
        eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
        eng.hmc_init(S, eps, 1.0, L, _lib.HMC_REFERENCE, q0=q)
        eng.hmc_inject(p=p0, u=u)
        eng.hmc_run(1, burning=False, sampling=True)
        assert int(eng.info("path_used")) == _lib.PATH_FUSED_SMALL
        qs, ps = eng.hmc_state()
        outs.append((qs, ps, eng.hmc_last()))
        eng.close()
    (qa, pa, la), (qb, pb, lb) = outs
    np.testing.assert_allclose(qa, qb, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(pa, pb, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(la["U1"], lb["U1"], rtol=1e-5)
    np.testing.assert_array_equal(la["accept"], lb["accept"])
    if S == 1:
        mu, sg = O.expand_prior(spec_o, 0.0, 1.0)
        want = O.hmc_iteration(O.Problem(spec_o, X, y, O.LOSS_SPARSE_CE, mu, sg), q, p0, u, eps, 1.0, L, False,
                               O.HMC_REFERENCE, np.float64)
        assert rel_err(pa, want["pL"]) < 1e-3 and abs(la["U1"][0] - want["U1"][0]) < 1e-4 * abs(want["U1"][0])


def test_chain_batching_does_not_change_results(oracle):
    g = load_golden("hmc_c1_mini")
    a, _ = make_engine(g, oracle)
    b, _ = make_engine(g, oracle)
    b.set_option("chain_batch", 1)
    Ua, la, ga = a.hmc_eval(g["q"])
    Ub, lb, gb = b.hmc_eval(g["q"])
    np.testing.assert_array_equal(Ua, Ub)
    np.testing.assert_array_equal(ga, gb)


# ---- SVGD -------------------------------------------------------------------------------------
def svgd_engine(g, semantics):
    js = keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"])
    eng = Engine(keras_json.parse_model_json(js))
    eng.set_dataset(g["X"], g["y"], _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    eng.svgd_init(6, float(g["lr"]), semantics, particles0=g["particles0"])
    return eng


def test_svgd_live_steps_match_oracle():
    g = load_golden("svgd_mini")
    eng = svgd_engine(g, _lib.SVGD_REFERENCE_LIVE)
    losses = [eng.svgd_step(ix) for ix in g["idx"]]
    np.testing.assert_allclose(losses, g["live_losses"], rtol=TOL_LOGP)
    # Adam's first steps move every coordinate by ~lr regardless of |phi|: compare with an absolute
    # budget of 1e-3 * lr per step on top of the relative one
    assert np.abs(eng.svgd_particles() - g["live_particles"]).max() < 2e-3 * float(g["lr"]) * len(g["idx"]) + 1e-6


@pytest.mark.parametrize("variant", ["one_cta", "cooperative", "per_particle_launches"])
def test_svgd_live_sweep_variants_agree(variant):
    """the sequential live sweep as ONE CTA with the particles in shared memory (default for the reference's sizes), as one
    cooperative launch with grid barriers, and as 2 S launches: the same update up to the float64 reduction order of
    the kernel rows, and all of them inside the oracle's budget"""
    g = load_golden("svgd_mini")
    eng = svgd_engine(g, _lib.SVGD_REFERENCE_LIVE)
    eng.set_option("live_cta", 1 if variant == "one_cta" else 0)
    eng.set_option("live_fused", 0 if variant == "per_particle_launches" else 1)
    losses = [eng.svgd_step(ix) for ix in g["idx"]]
    np.testing.assert_allclose(losses, g["live_losses"], rtol=TOL_LOGP)
    assert np.abs(eng.svgd_particles() - g["live_particles"]).max() < 2e-3 * float(g["lr"]) * len(g["idx"]) + 1e-6
    assert int(eng.info("kernel_launches")) > 0


def test_svgd_canonical_steps_match_oracle():
    g = load_golden("svgd_mini")
    eng = svgd_engine(g, _lib.SVGD_CANONICAL_MEDIAN)
    losses = [eng.svgd_step(ix) for ix in g["idx"]]
    np.testing.assert_allclose(losses, g["can_losses"], rtol=TOL_LOGP)
    assert np.abs(eng.svgd_particles() - g["can_particles"]).max() < 2e-3 * float(g["lr"]) * len(g["idx"]) + 1e-6


def test_svgd_phi_hook_median_bandwidth_is_exact(oracle):
    g = load_golden("svgd_mini")
    eng = svgd_engine(g, _lib.SVGD_CANONICAL_MEDIAN)
    phi, h = eng.svgd_phi(g["particles0"], g["G"], _lib.SVGD_CANONICAL_MEDIAN)
    assert abs(h - float(g["h_hook"])) < 1e-9 * float(g["h_hook"])
    assert rel_err(phi, g["phi_hook"]) < 1e-5
    # odd and even particle counts exercise both median branches
    rng = np.random.default_rng(5)
    # (from 64 particles on the select compacts its candidates after two passes: 80, 131 and the duplicated rows below —
    # groups of EQUAL distances around the median — exercise that branch (select_compact = 1); 0 (default) is the plain
    # eight-pass select, 2 the same passes with every digit picked inside the next pass (the kernels the sharded step's
    # side chain uses))
    for S in (5, 8, 33, 80, 131):
        X = rng.standard_normal((S, 252)) * 0.3
        if S == 131:
            X[40:80] = X[0:40]                      # exact ties: 40 duplicated particles
        G = rng.standard_normal((S, 252)).astype(np.float32)
        want, h_ref, _ = oracle.svgd_phi_canonical(X.astype(np.float32).astype(np.float64), G)
        for compact in (2, 1, 0):
            eng.set_option("select_compact", compact)
            phi, h = eng.svgd_phi(X.astype(np.float32).astype(np.float64), G, _lib.SVGD_CANONICAL_MEDIAN)
            assert abs(h - h_ref) < 1e-9 * h_ref and rel_err(phi, want) < 1e-5, (S, compact, h, h_ref)
    eng.set_option("select_compact", 0)


def test_svgd_live_formula_hook(oracle):
    rng = np.random.default_rng(6)
    g = load_golden("svgd_mini")
    eng = svgd_engine(g, _lib.SVGD_REFERENCE_LIVE)
    X = (rng.standard_normal((7, 252)) * 0.05).astype(np.float32).astype(np.float64)
    G = rng.standard_normal((7, 252)).astype(np.float32)
    phi, _ = eng.svgd_phi(X, G, _lib.SVGD_REFERENCE_LIVE)
    diff = X[:, None, :] - X[None, :, :]
    K = np.exp(-(diff ** 2).sum(-1))
    want = (K.sum(1)[:, None] * G + 2 * (K[:, :, None] * diff).sum(1)) / 7
    assert rel_err(phi, want) < 1e-5


def test_svgd_device_init_draws_from_the_prior():
    js = keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"])
    eng = Engine(keras_json.parse_model_json(js), seed=3)
    g = load_golden("svgd_mini")
    eng.set_dataset(g["X"], g["y"], _lib.LOSS_SPARSE_CE)
    eng.set_prior([1.5], [0.25], _lib.PRIOR_SCALAR)
    eng.svgd_init(64, 1e-3, _lib.SVGD_REFERENCE_LIVE)
    p = eng.svgd_particles()
    assert abs(p.mean() - 1.5) < 0.01 and abs(p.std() - 0.25) < 0.01
    loss = eng.svgd_step(None)          # full batch
    assert np.isfinite(loss)


# ---- predictive -------------------------------------------------------------------------------
def test_predictive_matches_golden():
    g = load_golden("predict_mini")
    js = keras_json.make_sequential_json(2, [50, 2], ["relu", "softmax"])
    eng = Engine(keras_json.parse_model_json(js))
    mean, var, allo = eng.predict(g["W"], g["x"], weights=g["freq"].astype(np.float32), want_all=True)
    np.testing.assert_allclose(mean, g["mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(var, g["var"], rtol=1e-3, atol=1e-6)
    np.testing.assert_array_equal(mean.max(-1) < 0.7, g["mask"])
    mean1, var1, _ = eng.predict(g["W"], g["x"])
    np.testing.assert_allclose(mean1, g["mean_unweighted"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(allo.mean(0), g["mean_unweighted"], rtol=1e-4, atol=1e-6)
    # NaN -> 0 per element (BayesianModel.py:125)
    W = g["W"].copy()
    W[0, :] = np.nan
    mean_nan, _, all_nan = eng.predict(W, g["x"], want_all=True)
    assert np.all(all_nan[0] == 0) and np.all(np.isfinite(mean_nan))


def test_errors_cross_the_boundary_as_status_codes():
    js = keras_json.make_sequential_json(2, [4, 2], ["relu", "softmax"])
    eng = Engine(keras_json.parse_model_json(js))
    with pytest.raises(_lib.PyesianB200Error, match="pyb_set_dataset"):
        eng.hmc_init(2, 0.1, 1.0, 2)
    x = np.zeros((8, 2), np.float32)
    with pytest.raises(_lib.PyesianB200Error, match="label out of range"):
        eng.set_dataset(x, np.full(8, 5, np.int32), _lib.LOSS_SPARSE_CE)
    with pytest.raises(_lib.PyesianB200Error, match="softmax"):
        eng.set_dataset(x, np.zeros((8, 2), np.float32), _lib.LOSS_MSE)
    eng.set_dataset(x, np.zeros(8, np.int32), _lib.LOSS_SPARSE_CE)
    eng.set_prior([0.0], [1.0], _lib.PRIOR_SCALAR)
    with pytest.raises(_lib.PyesianB200Error, match="pyb_hmc_init"):
        eng.hmc_run(1)
    with pytest.raises(_lib.PyesianB200Error):
        eng.hmc_init(2, 0.1, 1.0, 0)


def test_auto_path_selection(oracle):
    """AUTO picks the fused one-launch kernel for small-width nets and the generic path otherwise."""
    g = load_golden("hmc_c1_mini")
    eng, _ = make_engine(g, oracle, "auto")
    eng.hmc_init(4, 1e-3, 1.0, 3)
    d = eng.hmc_run(2, burning=False, sampling=True)
    assert int(eng.info("path_used")) == _lib.PATH_FUSED_SMALL
    assert d["kernel_launches"] == 2 * 3          # trajectory kernel + slot scan + select/record per iteration
    g3 = load_golden("hmc_c3_mini")                # 96 rows: below the tensor path's minimum tile
    eng3, _ = make_engine(g3, oracle, "auto")
    eng3.hmc_eval(g3["q"])
    assert int(eng3.info("path_used")) == _lib.PATH_GENERIC


@pytest.mark.parametrize("name,path", [("hmc_c1_mini", "generic"), ("hmc_c3_mini", "generic"), ("hmc_c3_mini", "auto")])
def test_carried_evaluation_is_bit_identical_to_re_evaluation(oracle, name, path):
    """"hmc_carry": the loss / gradient at q0 come from the previous iteration (its end point if accepted, its start
    if rejected) instead of being re-evaluated as HMC.py:80,82 do.  Same bits, L instead of L+1 evaluations."""
    g = load_golden(name)
    S, L, n_it = 6, 3, 7
    spec = spec_from_golden(g, oracle)
    eps = 1e-2 if spec.in_dim < 100 else 8e-2        # large enough that some proposals are rejected (oracle-calibrated)
    runs = {}
    for carry in (1, 0):
        eng, _ = make_engine(g, oracle, path, seed=5)
        eng.set_option("hmc_carry", carry)
        eng.hmc_init(S, eps, 1.0, L, _lib.HMC_REFERENCE)
        eng.hmc_run(2, burning=True, sampling=False)
        d1 = eng.hmc_run(n_it, burning=False, sampling=True)
        # two calls, and a change of data in between: the carried evaluation must be dropped, not reused
        eng.set_dataset(g["X"][::-1].copy(), g["y"][::-1].copy(), int(g["loss_kind"]))
        d2 = eng.hmc_run(2, burning=False, sampling=True)
        runs[carry] = (eng.hmc_state(), eng.hmc_last(), eng.hmc_samples(), d1, d2)
        if int(eng.info("path_used")) == _lib.PATH_FUSED_SMALL:
            pytest.skip("the one-launch small-width kernel evaluates everything in shared memory")
        eng.close()
    (qa, pa), la, sa, d1a, d2a = runs[1]
    (qb, pb), lb, sb, d1b, d2b = runs[0]
    np.testing.assert_array_equal(qa, qb)
    np.testing.assert_array_equal(pa, pb)
    for k in ("U0", "U1", "K0", "K1", "log_alpha", "accept", "loss"):
        np.testing.assert_array_equal(la[k], lb[k], err_msg=k)
    for a, b in zip(sa, sb):
        np.testing.assert_array_equal(a, b)
    assert 0 < d1a["n_accepted"] < S * n_it, "the case must contain accepted AND rejected proposals"
    assert d1a["n_accepted"] == d1b["n_accepted"] and d2a["n_accepted"] == d2b["n_accepted"]
    assert d1b["grad_evals"] == S * n_it * (L + 1) and d2b["grad_evals"] == S * 2 * (L + 1)
    assert d1a["grad_evals"] == S * n_it * L                 # the burn-in call left its end-point evaluation behind
    assert d2a["grad_evals"] == S * (2 * L + 1)              # new dataset: one fresh evaluation at the start position

import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import pyesian_oracle as O
from test_gpu_tensor import engine, problem
from conftest import rel_err
from bayesian_inference_for_nn_b200 import _lib
D,H,Cc,N,S=784,256,10,512,3
spec, prob, q, out_act, _ = problem(O, D, H, Cc, N, S, seed=1, act="relu", loss="ce")
U64, loss64, g64 = O.potential(prob, q, np.float64)
eng = engine(D,H,Cc,"relu",out_act)
eng.set_option("tc_i8", 2)
eng.set_dataset(prob.X, prob.y, prob.loss_kind)
eng.set_prior([0.0],[1.0],_lib.PRIOR_SCALAR)
eng.set_option("path", _lib.PATH_TENSOR)
for em in (0,1):
    eng.set_option("tc_epi_mma", em)
    U, ls, g = eng.hmc_eval(q)
    print("epi_mma", em, "loss", ls, "loss64", loss64, "grad err", [float("%.2e" % rel_err(g[s], g64[s])) for s in range(S)])

"""Generate tests/golden/gmm_*.npz -- the independent pin for inst/clusterSVGs.stan (R/clusterSVGsBayes.R:139-160).

A literal transcription of /root/reference/inst/clusterSVGs.stan, statement by statement (the K x N double loop of the model
block included), evaluated with torch in fp64: values from the Stan lpdfs written the way Stan Math evaluates `~` statements
(propto drops terms whose arguments are all data / constants [U]), gradients from autograd.  The simplex is Stan's stick-breaking
transform with its log-Jacobian [U: stan::math::simplex_constrain], `<lower=0>` is exp with Jacobian log sigma.  Shares no code with
oracle/bvg_oracle.c (analytic gradient, reverse pass through the transform written by hand) or with csrc/k_gmm.cu.

Run:  python tests/golden/make_golden_gmm.py     (deterministic; rewrites the .npz files)
"""
import math
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_default_dtype(torch.float64)
LOG_SQRT_2PI = 0.5 * math.log(2 * math.pi)


def stan_log_prob(th, X, K, jacobian, propto):
    N, D = X.shape
    y, um, us = th[:K - 1], th[K - 1:K - 1 + K * D], th[K - 1 + K * D:]
    # simplex[K] theta
    theta, stick, lj = [], torch.tensor(1.0), torch.tensor(0.0)
    for k in range(K - 1):
        z = torch.sigmoid(y[k] - math.log(K - 1 - k))
        theta.append(stick * z)
        lj = lj + torch.log(stick) + torch.log(z) + torch.log1p(-z)
        stick = stick - theta[-1]
    theta.append(stick)
    theta = torch.stack(theta)
    mu = um.reshape(D, K).T                    # matrix[K, D]: column-major serialisation
    sigma = torch.exp(us.reshape(K, D))        # array[K] vector<lower=0>[D]
    lj = lj + us.sum()
    lp = torch.tensor(0.0)
    # to_vector(mu) ~ normal(0, 7);
    lp = lp + (-0.5 * (mu / 7.0) ** 2).sum() - (0 if propto else (LOG_SQRT_2PI + math.log(7.0)) * K * D)
    # for (i in 1:K) sigma[i] ~ normal(0, 3);
    for i in range(K):
        lp = lp + (-0.5 * (sigma[i] / 3.0) ** 2).sum() - (0 if propto else (LOG_SQRT_2PI + math.log(3.0)) * D)
    # theta ~ dirichlet(rep_vector(2, K));
    lp = lp + ((2.0 - 1.0) * torch.log(theta)).sum() + (0 if propto else math.lgamma(2.0 * K) - K * math.lgamma(2.0))
    # log_theta_const + the N x K loop with log_sum_exp
    ltc = [torch.log(theta[i]) - 0.5 * (D * math.log(2 * math.pi) + 2 * torch.log(sigma[i]).sum()) for i in range(K)]
    for i in range(N):
        comp = []
        for j in range(K):
            scaled = (X[i] - mu[j]) / sigma[j]
            comp.append(ltc[j] - 0.5 * (scaled * scaled).sum())
        lp = lp + torch.logsumexp(torch.stack(comp), 0)
    return lp + (lj if jacobian else 0.0)


def make(name, N, D, K, seed):
    rng = np.random.default_rng(seed)
    cent = rng.normal(0, 3, (K, D))
    X = cent[rng.integers(0, K, N)] + rng.normal(0, 0.8, (N, D))
    P = K - 1 + 2 * K * D
    thetas = np.stack([np.zeros(P), rng.normal(0, 0.3, P), rng.normal(0, 1.0, P)])
    out = dict(X=X, K=K, thetas=thetas)
    Xt = torch.tensor(X)
    for jac in (0, 1):
        for pro in (0, 1):
            lps, grads = [], []
            for th in thetas:
                t = torch.tensor(th, requires_grad=True)
                lp = stan_log_prob(t, Xt, K, bool(jac), bool(pro))
                lp.backward()
                lps.append(float(lp.detach()))
                grads.append(t.grad.numpy().copy())
            out[f"lp_j{jac}_p{pro}"] = np.array(lps)
            out[f"grad_j{jac}_p{pro}"] = np.stack(grads)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "P =", P, out["lp_j1_p0"])


if __name__ == "__main__":
    make("gmm_small", 40, 3, 3, 1)
    make("gmm_default_shape", 60, 30, 5, 2)     # n.PCs = 30, n.clust = 5 (R/clusterSVGsBayes.R:65-66)

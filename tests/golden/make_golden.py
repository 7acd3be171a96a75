"""Generate tests/golden/*.npz -- an INDEPENDENT pin for the CPU oracle and the CUDA engine.

The reference (R + CmdStan) cannot run in this image and its own tests
hold no numeric vectors, so the pin is a *literal transcription* of the Stan programs
(/root/reference/inst/approxGP.stan, inst/approxGP4.stan and the depth-adjusted inst/approxGP2.stan,
inst/approxGP3.stan), statement by statement -- including
the spot_id/gene_id gather loop that the oracle and the engine optimise away -- evaluated with
scipy.stats log-densities (values) and torch-autograd fp64 (gradients).  It shares no code with
oracle/bvg_oracle.c (analytic gradient) or with the CUDA kernels.

Also pinned here: the basis builder against scipy.linalg.qr(pivoting=True) (LAPACK dgeqp3, the
routine R's qr(LAPACK=TRUE) calls) and the Matern closed forms against scipy.special.kv, i.e. the
besselK expression in /root/reference/R/maternKernel.R:22-26.

Run:  python tests/golden/make_golden.py     (deterministic; rewrites the .npz files)
"""
import os

import numpy as np
import scipy.linalg
import scipy.special
import scipy.stats as st
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_default_dtype(torch.float64)
LOG_SQRT_2PI = 0.5 * np.log(2 * np.pi)


# ----------------------------------------------------------------------------------------------
# Stan lpdfs, written the way Stan Math evaluates `~` statements (propto drops terms whose
# arguments are all data/constants) [U].
# ----------------------------------------------------------------------------------------------
def std_normal(x, propto):
    return (-0.5 * x * x).sum() - (0 if propto else LOG_SQRT_2PI * x.numel())


def normal_const_scale(x, sigma, propto):  # normal(x | 0, sigma) with sigma a literal
    return (-0.5 * (x / sigma) ** 2).sum() - (0 if propto else (LOG_SQRT_2PI + np.log(sigma)) * x.numel())


def normal_lpdf(y, mean, sigma, propto):  # sigma is a parameter: -log sigma always kept
    n = y.numel()
    return (-0.5 * ((y - mean) / sigma) ** 2).sum() - n * torch.log(sigma) - (0 if propto else LOG_SQRT_2PI * n)


def lognormal_lpdf(y, mu, sigma, propto):
    n = y.numel()
    return (-torch.log(y) - 0.5 * ((torch.log(y) - mu) / sigma) ** 2).sum() - n * torch.log(sigma) - (
        0 if propto else LOG_SQRT_2PI * n)


def cauchy_trunc0(x, sigma, propto):  # cauchy(0, sigma) T[0, ]
    lp = -torch.log1p((x / sigma) ** 2)
    if not propto:
        lp = lp - np.log(np.pi) - np.log(sigma) + np.log(2.0)  # lpdf constants, minus log(1 - cdf(0))
    return lp


def neg_binomial_2_log(y, eta, phi):
    # lchoose(y + phi - 1, y) + y*eta - phi*log1p_exp(eta - log phi) - y*(log phi + log1p_exp(eta - log phi))
    L = torch.nn.functional.softplus(eta - torch.log(phi))
    return (torch.lgamma(y + phi) - torch.lgamma(y + 1) - torch.lgamma(phi) + y * eta - phi * L
            - y * (torch.log(phi) + L)).sum()


# ----------------------------------------------------------------------------------------------
# Literal model blocks.  theta: unconstrained vector in declaration order.
# ----------------------------------------------------------------------------------------------
def stan_gp(theta, phi, y, spot_id, gene_id, G, k, jacobian, propto):
    """inst/approxGP.stan"""
    i = 0
    lj = 0.0

    def take(n, lower0=False):
        nonlocal i, lj
        u = theta[i:i + n]
        i += n
        if lower0:
            lj = lj + u.sum()
            return torch.exp(u)
        return u

    mu_beta0 = take(1)          # :13
    sigma_beta0 = take(1, True)  # :14
    z_beta0 = take(G)           # :15
    mu_amplitude = take(1)      # :16
    sigma_amplitude = take(1, True)  # :17
    amplitude = take(G, True)   # :18
    mu_alpha = take(k)          # :19
    sigma_alpha = take(k, True)  # :20
    z_alpha_t = take(k * G).reshape(G, k).T  # :21 matrix[k, G], column-major
    sigma_y = take(1, True)     # :22
    assert i == theta.numel()
    beta0 = mu_beta0 + sigma_beta0 * z_beta0  # :26
    amplitude_sq = amplitude ** 2             # :27
    alpha_t = mu_alpha[:, None].expand(k, G) + sigma_alpha[:, None] * z_alpha_t  # :32
    phi_alpha = phi @ alpha_t                 # :33
    w = phi_alpha[spot_id, gene_id]           # :35-37
    lp = std_normal(z_alpha_t, propto)        # :38
    lp = lp + std_normal(z_beta0, propto)     # :39
    lp = lp + normal_const_scale(mu_beta0, 2.0, propto)   # :40
    lp = lp + std_normal(sigma_beta0, propto)             # :41
    lp = lp + normal_const_scale(mu_amplitude, 2.0, propto)  # :42
    lp = lp + std_normal(sigma_amplitude, propto)         # :43
    lp = lp + lognormal_lpdf(amplitude, mu_amplitude, sigma_amplitude, propto)  # :44
    lp = lp + normal_const_scale(mu_alpha, 2.0, propto)   # :45
    lp = lp + std_normal(sigma_alpha, propto)             # :46
    lp = lp + std_normal(sigma_y, propto)                 # :47
    lp = lp + normal_lpdf(y, beta0[gene_id] + amplitude_sq[gene_id] * w, sigma_y, propto)  # :48
    return lp + (lj if jacobian else 0.0)


def stan_gp4(theta, phi, y, spot_id, gene_id, G, k, jacobian, propto):
    """inst/approxGP4.stan"""
    i = 0
    lj = 0.0

    def take(n, lower0=False):
        nonlocal i, lj
        u = theta[i:i + n]
        i += n
        if lower0:
            lj = lj + u.sum()
            return torch.exp(u)
        return u

    mu_beta0 = take(1)
    sigma_beta0 = take(1, True)
    z_beta0 = take(G)
    phi_nb = take(1, True)       # :16
    amplitude = take(G, True)    # :17
    mu_amplitude = take(1)
    sigma_amplitude = take(1, True)
    mu_alpha = take(k)
    sigma_alpha = take(k, True)
    z_alpha_t = take(k * G).reshape(G, k).T
    assert i == theta.numel()
    beta0 = mu_beta0 + sigma_beta0 * z_beta0
    amplitude_sq = amplitude ** 2
    alpha_t = mu_alpha[:, None].expand(k, G) + sigma_alpha[:, None] * z_alpha_t
    phi_alpha = phi @ alpha_t
    lp = std_normal(z_alpha_t, propto)
    lp = lp + normal_const_scale(mu_beta0, 2.0, propto)
    lp = lp + std_normal(sigma_beta0, propto)
    lp = lp + std_normal(z_beta0, propto)
    lp = lp + normal_const_scale(mu_amplitude, 2.0, propto)
    lp = lp + std_normal(sigma_amplitude, propto)
    lp = lp + lognormal_lpdf(amplitude, mu_amplitude, sigma_amplitude, propto)
    lp = lp + cauchy_trunc0(phi_nb, 2.0, propto).sum()     # :41
    lp = lp + normal_const_scale(mu_alpha, 2.0, propto)
    lp = lp + std_normal(sigma_alpha, propto)
    w = phi_alpha[spot_id, gene_id]                         # :44-47
    lp = lp + neg_binomial_2_log(y, beta0[gene_id] + amplitude_sq[gene_id] * w, phi_nb)  # :48
    return lp + (lj if jacobian else 0.0)


def stan_gp2(theta, phi, y, spot_id, gene_id, G, k, jacobian, propto, gene_depths):
    """inst/approxGP2.stan"""
    i = 0
    lj = 0.0

    def take(n, lower0=False):
        nonlocal i, lj
        u = theta[i:i + n]
        i += n
        if lower0:
            lj = lj + u.sum()
            return torch.exp(u)
        return u

    beta0 = take(1)             # :14
    beta1 = take(1)             # :15
    sigma_y = take(1, True)     # :16
    amplitude = take(G, True)   # :17
    mu_amplitude = take(1)      # :18
    sigma_amplitude = take(1, True)  # :19
    mu_alpha = take(k)          # :20
    sigma_alpha = take(k, True)  # :21
    z_alpha_t = take(k * G).reshape(G, k).T  # :22
    assert i == theta.numel()
    amplitude_sq = amplitude ** 2            # :26
    alpha_t = mu_alpha[:, None].expand(k, G) + sigma_alpha[:, None] * z_alpha_t  # :31
    phi_alpha = phi @ alpha_t                # :32
    w = phi_alpha[spot_id, gene_id]          # :33-36
    lp = std_normal(z_alpha_t, propto)       # :37
    lp = lp + normal_const_scale(beta0, 2.0, propto)         # :38
    lp = lp + normal_const_scale(beta1, 2.0, propto)         # :39
    lp = lp + normal_const_scale(mu_alpha, 2.0, propto)      # :40
    lp = lp + std_normal(sigma_alpha, propto)                # :41
    lp = lp + normal_const_scale(mu_amplitude, 2.0, propto)  # :42
    lp = lp + std_normal(sigma_amplitude, propto)            # :43
    lp = lp + normal_const_scale(sigma_y, 2.0, propto)       # :44
    lp = lp + lognormal_lpdf(amplitude, mu_amplitude, sigma_amplitude, propto)  # :45
    lp = lp + normal_lpdf(y, beta0 + beta1 * gene_depths[gene_id] + amplitude_sq[gene_id] * w, sigma_y, propto)  # :46
    return lp + (lj if jacobian else 0.0)


def stan_gp3(theta, phi, y, spot_id, gene_id, G, k, jacobian, propto, gene_depths):
    """inst/approxGP3.stan"""
    i = 0
    lj = 0.0

    def take(n, lower0=False):
        nonlocal i, lj
        u = theta[i:i + n]
        i += n
        if lower0:
            lj = lj + u.sum()
            return torch.exp(u)
        return u

    beta0 = take(1)             # :14
    beta1 = take(1)             # :15
    phi_nb = take(1, True)      # :16
    amplitude = take(G, True)   # :17
    mu_amplitude = take(1)
    sigma_amplitude = take(1, True)
    mu_alpha = take(k)
    sigma_alpha = take(k, True)
    z_alpha_t = take(k * G).reshape(G, k).T
    assert i == theta.numel()
    amplitude_sq = amplitude ** 2
    lp = normal_const_scale(beta0, 2.0, propto)              # :30
    lp = lp + normal_const_scale(beta1, 2.0, propto)         # :31
    alpha_t = mu_alpha[:, None].expand(k, G) + sigma_alpha[:, None] * z_alpha_t
    phi_alpha = phi @ alpha_t
    lp = lp + std_normal(z_alpha_t, propto)
    lp = lp + normal_const_scale(mu_alpha, 2.0, propto)
    lp = lp + std_normal(sigma_alpha, propto)
    lp = lp + normal_const_scale(mu_amplitude, 2.0, propto)
    lp = lp + std_normal(sigma_amplitude, propto)
    lp = lp + lognormal_lpdf(amplitude, mu_amplitude, sigma_amplitude, propto)
    lp = lp + cauchy_trunc0(phi_nb, 2.0, propto).sum()       # :41
    w = phi_alpha[spot_id, gene_id]                           # :42-45
    lp = lp + neg_binomial_2_log(y, beta0 + beta1 * gene_depths[gene_id] + amplitude_sq[gene_id] * w, phi_nb)  # :46
    return lp + (lj if jacobian else 0.0)


def scipy_value_check_gp23(theta, phi, y, G, k, depths, nb):
    """values only (propto = FALSE, jacobian = FALSE) for the depth-adjusted programs"""
    th = theta
    b0, b1, tail = th[0], th[1], np.exp(th[2])
    amp = np.exp(th[3:3 + G]); mu_a, s_a = th[3 + G], np.exp(th[4 + G])
    mu_al = th[5 + G:5 + G + k]; s_al = np.exp(th[5 + G + k:5 + G + 2 * k])
    z = th[5 + G + 2 * k:].reshape(G, k).T
    alpha = mu_al[:, None] + s_al[:, None] * z
    eta = (b0 + b1 * depths)[None, :] + (amp ** 2)[None, :] * (phi @ alpha)
    lp = st.norm.logpdf(z).sum() + st.norm.logpdf(b0, 0, 2) + st.norm.logpdf(b1, 0, 2)
    lp += st.norm.logpdf(mu_a, 0, 2) + st.norm.logpdf(s_a) + st.lognorm.logpdf(amp, s=s_a, scale=np.exp(mu_a)).sum()
    lp += st.norm.logpdf(mu_al, 0, 2).sum() + st.norm.logpdf(s_al).sum()
    if nb:
        lp += st.cauchy.logpdf(tail, 0, 2) - st.cauchy.logsf(0, 0, 2)
        lp += st.nbinom.logpmf(y, tail, tail / (tail + np.exp(eta))).sum()
    else:
        lp += st.norm.logpdf(tail, 0, 2) + st.norm.logpdf(y, eta, tail).sum()
    return lp


def scipy_value_check_gp(theta, phi, y, G, k):
    """values only, with scipy.stats frozen distributions (propto = FALSE, jacobian = FALSE)"""
    th = theta
    mu_b, s_b = th[0], np.exp(th[1]); zb = th[2:2 + G]
    mu_a, s_a = th[2 + G], np.exp(th[3 + G]); amp = np.exp(th[4 + G:4 + 2 * G])
    mu_al = th[4 + 2 * G:4 + 2 * G + k]; s_al = np.exp(th[4 + 2 * G + k:4 + 2 * G + 2 * k])
    z = th[4 + 2 * G + 2 * k:4 + 2 * G + 2 * k + k * G].reshape(G, k).T
    s_y = np.exp(th[-1])
    alpha = mu_al[:, None] + s_al[:, None] * z
    mean = (mu_b + s_b * zb)[None, :] + (amp ** 2)[None, :] * (phi @ alpha)
    lp = st.norm.logpdf(z).sum() + st.norm.logpdf(zb).sum() + st.norm.logpdf(mu_b, 0, 2) + st.norm.logpdf(s_b)
    lp += st.norm.logpdf(mu_a, 0, 2) + st.norm.logpdf(s_a) + st.lognorm.logpdf(amp, s=s_a, scale=np.exp(mu_a)).sum()
    lp += st.norm.logpdf(mu_al, 0, 2).sum() + st.norm.logpdf(s_al).sum() + st.norm.logpdf(s_y)
    lp += st.norm.logpdf(y, mean, s_y).sum()
    return lp


def scipy_value_check_gp4(theta, phi, y, G, k):
    th = theta
    mu_b, s_b = th[0], np.exp(th[1]); zb = th[2:2 + G]
    ph = np.exp(th[2 + G]); amp = np.exp(th[3 + G:3 + 2 * G])
    mu_a, s_a = th[3 + 2 * G], np.exp(th[4 + 2 * G])
    mu_al = th[5 + 2 * G:5 + 2 * G + k]; s_al = np.exp(th[5 + 2 * G + k:5 + 2 * G + 2 * k])
    z = th[5 + 2 * G + 2 * k:].reshape(G, k).T
    alpha = mu_al[:, None] + s_al[:, None] * z
    eta = (mu_b + s_b * zb)[None, :] + (amp ** 2)[None, :] * (phi @ alpha)
    m = np.exp(eta)
    lp = st.norm.logpdf(z).sum() + st.norm.logpdf(zb).sum() + st.norm.logpdf(mu_b, 0, 2) + st.norm.logpdf(s_b)
    lp += st.norm.logpdf(mu_a, 0, 2) + st.norm.logpdf(s_a) + st.lognorm.logpdf(amp, s=s_a, scale=np.exp(mu_a)).sum()
    lp += st.norm.logpdf(mu_al, 0, 2).sum() + st.norm.logpdf(s_al).sum()
    lp += st.cauchy.logpdf(ph, 0, 2) - st.cauchy.logsf(0, 0, 2)
    lp += st.nbinom.logpmf(y, ph, ph / (ph + m)).sum()
    return lp


def make_case(model, M, G, k, seed, n_theta=3, depth=False):
    rng = np.random.default_rng(seed)
    coords = rng.normal(size=(M, 2))
    raw = np.exp(-0.5 * ((coords[:, None, :] - coords[rng.choice(M, k, replace=False)][None]) ** 2).sum(-1))
    phi, _ = np.linalg.qr(raw)
    amp = np.exp(rng.normal(-0.5, 0.7, G))
    alpha = rng.normal(size=(k, G))
    beta = rng.normal(0, 0.5, G)
    mean = beta[None] + (amp ** 2)[None] * (phi @ alpha) * np.sqrt(M)
    if model == "gaussian":
        y = mean + rng.normal(size=(M, G))
        y = (y - y.mean()) / y.std(ddof=1)  # global z-score, findSpatiallyVariableFeaturesBayes.R:185
        fn, sfn = stan_gp, scipy_value_check_gp
    else:
        y = rng.negative_binomial(2.0, 2.0 / (2.0 + np.exp(np.log(0.5) + np.clip(mean, -4, 4)))).astype(np.float64)
        fn, sfn = stan_gp4, scipy_value_check_gp4
    D = 5 + 2 * G + 2 * k + k * G
    depths = None
    if depth:
        # gene_depths = log(rowSums(counts)) (R/findSpatiallyVariableFeaturesBayes.R:303); synthetic stand-in
        depths = np.log(rng.integers(500, 50000, G).astype(np.float64))
        td = torch.tensor(depths)
        base = stan_gp2 if model == "gaussian" else stan_gp3
        fn = lambda th, a, b, c, d, G_, k_, j, p: base(th, a, b, c, d, G_, k_, j, p, td)  # noqa: E731
        sfn = lambda t, ph, yy, G_, k_: scipy_value_check_gp23(t, ph, yy, G_, k_, depths, model != "gaussian")  # noqa: E731
        D = 5 + G + 2 * k + k * G
    gene_id = np.repeat(np.arange(G), M)  # gene-major long vector (:176-183)
    spot_id = np.tile(np.arange(M), G)
    ylong = y.T.reshape(-1)  # y[(g-1)*M + s]
    tphi, ty = torch.tensor(phi), torch.tensor(ylong)
    tsi, tgi = torch.tensor(spot_id), torch.tensor(gene_id)
    thetas = rng.normal(0, 0.6, size=(n_theta, D))
    thetas[0] = 0.0  # init = 0 (the L-BFGS start, :348)
    out = dict(phi=phi, y=y, thetas=thetas, M=M, G=G, k=k)
    if depth:
        out["gene_depths"] = depths
    for jac in (0, 1):
        for pro in (0, 1):
            lps, grads = [], []
            for t in thetas:
                th = torch.tensor(t, requires_grad=True)
                lp = fn(th, tphi, ty, tsi, tgi, G, k, bool(jac), bool(pro)).sum()
                lp.backward()
                lps.append(float(lp))
                grads.append(th.grad.numpy().copy())
            out[f"lp_j{jac}_p{pro}"] = np.array(lps)
            out[f"grad_j{jac}_p{pro}"] = np.array(grads)
    # independent value check (scipy frozen distributions) for propto=0, jacobian=0
    sv = np.array([sfn(t, phi, y, G, k) for t in thetas])
    assert np.allclose(sv, out["lp_j0_p0"], rtol=1e-12, atol=1e-9), (sv, out["lp_j0_p0"])
    out["lp_scipy_j0_p0"] = sv
    return out


def make_basis_case(seed=7, M=211, k=12):
    rng = np.random.default_rng(seed)
    # jittered hex-like lattice, z-scored like coop::scaler (:160)
    xy = rng.uniform(0, 10, size=(M, 2))
    xy = (xy - xy.mean(0)) / xy.std(0, ddof=1)
    centers = xy[rng.choice(M, k, replace=False)] + 0.01 * rng.normal(size=(k, 2))
    d = np.sqrt(((centers[:, None] - centers[None]) ** 2).sum(-1))
    lscale = float(np.median(d[np.triu_indices(k, 1)]))  # :205-207
    d2 = ((xy[:, None, :] - centers[None]) ** 2).sum(-1)
    out = dict(coords=xy, centers=centers, lscale=lscale)
    for name, nu in (("exp_quad", 0.0), ("matern0.5", 0.5), ("matern1.5", 1.5), ("matern2.5", 2.5), ("periodic", 0.0)):
        if name == "exp_quad":
            raw = np.exp(-d2 / (2 * lscale ** 2))  # R/expQuadKernel.R:17
        elif name == "periodic":
            period = 3.0
            raw = np.exp(-2 * np.sin(np.pi * np.sqrt(d2) / period) ** 2 / lscale ** 2)  # R/periodicKernel.R:20
        else:
            s = np.sqrt(2 * nu) * np.sqrt(d2) / lscale  # R/maternKernel.R:22-26, besselK form
            with np.errstate(invalid="ignore", divide="ignore"):
                raw = (2 ** (1 - nu) / scipy.special.gamma(nu)) * s ** nu * scipy.special.kv(nu, s)
            raw[~np.isfinite(raw)] = 1.0
        Q, R, piv = scipy.linalg.qr(raw, mode="economic", pivoting=True)  # dgeqp3 + dorgqr
        out[f"raw_{name}"] = raw
        out[f"Q_{name}"] = Q
        out[f"piv_{name}"] = piv
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "gp_small.npz"), **make_case("gaussian", 37, 5, 4, 1))
    np.savez_compressed(os.path.join(HERE, "gp4_small.npz"), **make_case("nb", 41, 6, 3, 2))
    np.savez_compressed(os.path.join(HERE, "gp_ragged.npz"), **make_case("gaussian", 131, 9, 20, 3, n_theta=2))
    np.savez_compressed(os.path.join(HERE, "gp4_ragged.npz"), **make_case("nb", 67, 7, 20, 4, n_theta=2))
    np.savez_compressed(os.path.join(HERE, "gp2_small.npz"), **make_case("gaussian", 43, 6, 5, 5, depth=True))
    np.savez_compressed(os.path.join(HERE, "gp3_small.npz"), **make_case("nb", 39, 5, 4, 6, depth=True))
    np.savez_compressed(os.path.join(HERE, "gp2_ragged.npz"), **make_case("gaussian", 97, 11, 20, 7, n_theta=2, depth=True))
    np.savez_compressed(os.path.join(HERE, "gp3_ragged.npz"), **make_case("nb", 71, 9, 20, 8, n_theta=2, depth=True))
    np.savez_compressed(os.path.join(HERE, "basis.npz"), **make_basis_case())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

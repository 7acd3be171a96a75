"""Synthetic datasets of the shapes BASELINE.json names.  Deterministic."""
import numpy as np

from .api import length_scale_kmeans, scale_coords


def visium_grid(rows=78, cols=64):
    """full Visium hex lattice: 4,992 spots, 100 um pitch"""
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    x = (2 * c + (r % 2)) * 50.0
    y = r * 50.0 * np.sqrt(3.0)
    return np.stack([x.ravel(), y.ravel()], 1)


def masked_blob(coords, n_keep, seed=312):
    """fixed-seed connected blob of n_keep spots (stand-in for the 2,444 in-tissue spots of seu_brain)"""
    rng = np.random.default_rng(seed)
    c0 = coords[rng.integers(len(coords))]
    w = rng.uniform(0.7, 1.3, 2)
    d = (((coords - c0) * w) ** 2).sum(1)
    return coords[np.sort(np.argsort(d, kind="stable")[:n_keep])]


def random_points(n, seed):
    return np.random.default_rng(seed).uniform(0, 1, size=(n, 2)) * 10000.0


def det_kmeans(X, k, seed=312, iters=100):
    """deterministic k-means (k-means++ init, Lloyd) whose centroids are stored with the dataset"""
    rng = np.random.default_rng(seed)
    M = X.shape[0]
    C = [X[rng.integers(M)]]
    d2 = ((X - C[0]) ** 2).sum(1)
    for _ in range(1, k):
        C.append(X[rng.choice(M, p=d2 / d2.sum())])
        d2 = np.minimum(d2, ((X - C[-1]) ** 2).sum(1))
    C = np.array(C)
    for _ in range(iters):
        lab = np.argmin(((X[:, None, :] - C[None]) ** 2).sum(-1) if M * k < 5e7 else _chunked_d2(X, C), 1)
        newC = np.array([X[lab == j].mean(0) if np.any(lab == j) else C[j] for j in range(k)])
        if np.allclose(newC, C, atol=1e-10):
            break
        C = newC
    return C


def _chunked_d2(X, C):
    out = np.empty((X.shape[0], C.shape[0]))
    for i in range(0, X.shape[0], 65536):
        out[i:i + 65536] = ((X[i:i + 65536, None, :] - C[None]) ** 2).sum(-1)
    return out


def expression(phi, G, likelihood="gaussian", seed=312, dtype=np.float64, signal=None):
    """y = beta_g + amp_g^2 * signal * Phi alpha_g + eps (gaussian, then the global z-score of :185) or
    NB2(mean = exp(log 0.5 + beta_g + amp_g^2 * signal * Phi alpha_g), phi = 2).  30% of genes have amp = 0.
    Phi has orthonormal columns (entries ~ 1/sqrt(M)), so `signal` defaults to sqrt(M / k) to give the
    spatial term unit variance per unit amplitude^2; the true amplitudes are returned for rank checks."""
    rng = np.random.default_rng(seed)
    M, k = phi.shape
    if signal is None:
        signal = np.sqrt(M / k)
    amp = np.exp(rng.normal(-1.0, 1.0, G))
    amp[rng.random(G) < 0.3] = 0.0
    beta = rng.normal(0, 0.5, G)
    Y = np.empty((M, G), dtype=dtype, order="F")
    step = max(1, int(2e7 // M))
    for g0 in range(0, G, step):
        g1 = min(G, g0 + step)
        alpha = rng.normal(size=(k, g1 - g0))
        mean = beta[None, g0:g1] + (amp[g0:g1] ** 2)[None] * signal * (phi @ alpha)
        if likelihood == "gaussian":
            Y[:, g0:g1] = mean + rng.normal(size=mean.shape)
        else:
            mu = np.exp(np.log(0.5) + np.clip(mean, -6, 6))
            Y[:, g0:g1] = rng.negative_binomial(2.0, 2.0 / (2.0 + mu))
    if likelihood == "gaussian":
        m = Y.mean()
        s = Y.std(ddof=1)
        Y -= m
        Y /= s
    return Y, amp


def dataset(name, basis_fn, G=None, seed=312):
    """name: cfg1 | cfg2 | cfg3 | cfg4 | tiny.  basis_fn(Xscaled, centers, lscale, kernel, nu) -> Phi."""
    spec = {
        "tiny": dict(coords=lambda: visium_grid(20, 16), G=64, k=8, kernel="matern", nu=2.5, lik="gaussian"),
        "cfg1": dict(coords=lambda: masked_blob(visium_grid(), 2444), G=3000, k=20, kernel="matern", nu=2.5, lik="gaussian"),
        "cfg2": dict(coords=visium_grid, G=3000, k=20, kernel="matern", nu=2.5, lik="gaussian"),
        "cfg3": dict(coords=visium_grid, G=3000, k=20, kernel="exp_quad", nu=0.0, lik="nb"),
        "cfg4": dict(coords=lambda: random_points(50000, 1), G=10000, k=20, kernel="matern", nu=1.5, lik="gaussian"),
    }[name]
    X = scale_coords(spec["coords"]())
    C = det_kmeans(X, spec["k"], seed)
    ls = length_scale_kmeans(C)
    phi = basis_fn(X, C, ls, spec["kernel"], spec["nu"])
    Gn = spec["G"] if G is None else G
    Y, amp = expression(phi, Gn, spec["lik"], seed)
    if spec["lik"] == "nb":
        Y = np.asfortranarray(Y.astype(np.int32))
    return dict(name=name, X=X, centers=C, lscale=ls, phi=phi, Y=Y, amp=amp, likelihood=spec["lik"],
                kernel=spec["kernel"], nu=spec["nu"], k=spec["k"])

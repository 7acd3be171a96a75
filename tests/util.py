import numpy as np

from oracle import oracle as orc


def make_problem(M, G, k, model="gaussian", seed=0, kernel="matern", nu=2.5):
    """small synthetic problem with an oracle-built basis (CPU only)"""
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-1.7, 1.7, (M, 2))
    C = xy[rng.choice(M, k, replace=False)] + 1e-3 * rng.normal(size=(k, 2))
    d = np.sqrt(((C[:, None] - C[None]) ** 2).sum(-1))
    ls = float(np.median(d[np.triu_indices(k, 1)])) if k > 1 else 1.0
    phi = orc.basis(xy, C, ls, kernel, nu)
    amp = np.exp(rng.normal(-1, 1, G))
    amp[rng.random(G) < 0.3] = 0
    alpha = rng.normal(size=(k, G))
    beta = rng.normal(0, 0.5, G)
    mean = beta[None] + (amp ** 2)[None] * (phi @ alpha) * np.sqrt(M / k)
    if model == "gaussian":
        y = mean + rng.normal(size=(M, G))
        y = (y - y.mean()) / y.std(ddof=1)
    else:
        y = rng.negative_binomial(2.0, 2.0 / (2.0 + np.exp(np.log(0.5) + np.clip(mean, -5, 5)))).astype(np.float64)
    return dict(xy=xy, centers=C, lscale=ls, phi=phi, y=np.asfortranarray(y), amp=amp)


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))

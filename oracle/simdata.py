"""Synthetic inputs for the VGA hot path -- TEST/BENCH INFRASTRUCTURE (CPU side).

Restates the data-generating process of ``sim_data_log``
(/root/reference/R/simulate_poisson_matrix.R:9-61) and the README example
(/root/reference/README.md:33-48) with NumPy's RNG (R's RNG stream cannot be
reproduced without R), and builds the solver state the VGA step consumes inside
``ebpmf_log``: the flash factors with the two fixed offset
columns, ``EL = [log l0 + c, 1, L~]`` and ``EF = [1, log f0, F~]`` (F10,
R/ebpmf_log_flash_init.R:31-45), a start ``M``, ``sigma2`` and ``R2``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def sim_data_log(n, p, K=3, d=None, add_background=True, background_unif_range=(1.0, 2.0),
                 pi0=0.8, var1=1.0, var_e=0.0, seed=12345, log_offset=0.0):
    """One draw of sim_data_log's DGP (R/simulate_poisson_matrix.R:25-53).

    ``log_offset`` adds a global constant to log(lambda) so that the configs'
    5-10 % densities can be hit (the package defaults give ~85-90 % non-zeros).
    """
    rng = np.random.default_rng(seed)
    d = np.ones(K) if d is None else np.asarray(d, dtype=np.float64)
    FF = rng.normal(0.0, np.sqrt(var1), size=(p, K)) * rng.binomial(1, 1 - pi0, size=(p, K))   # :29-31
    L = rng.normal(0.0, 1.0, size=(n, K)) * np.sqrt(d)[None, :]                                   # :33-36
    if add_background:
        l0 = rng.uniform(background_unif_range[0], background_unif_range[1], size=n)              # :38
        f0 = rng.uniform(background_unif_range[0], background_unif_range[1], size=p)              # :39
    else:
        l0 = np.ones(n); f0 = np.ones(p)
    loglam = np.log(l0)[:, None] + np.log(f0)[None, :] + L @ FF.T + log_offset                   # :42,47
    if var_e > 0:
        loglam = loglam + rng.normal(0.0, np.sqrt(var_e), size=(n, p))
    Y = rng.poisson(np.exp(loglam)).astype(np.float64)                                            # :48
    return dict(Y=np.asfortranarray(Y), Factor=FF, Loading=L, L0=l0, F0=f0, log_offset=log_offset)


def offset_for_density(target, K=3, background_unif_range=(1.0, 2.0), pi0=0.8, var1=1.0, seed=7,
                       sample=2048):
    """Bisect the global log-offset c so that mean(1-exp(-lambda)) == target on a sample."""
    rng = np.random.default_rng(seed)
    FF = rng.normal(0.0, np.sqrt(var1), size=(sample, K)) * rng.binomial(1, 1 - pi0, size=(sample, K))
    L = rng.normal(size=(sample, K))
    l0 = rng.uniform(*background_unif_range, size=sample)
    f0 = rng.uniform(*background_unif_range, size=sample)
    base = np.log(l0)[:, None] + np.log(f0)[None, :] + L @ FF.T
    lo, hi = -30.0, 10.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        dens = np.mean(-np.expm1(-np.exp(base + mid)))
        if dens < target:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def readme_example(N=1000, p=100, K=3, seed=12345):
    """README.md:33-48 recipe shapes (C1): block Ftrue of 1/2/3, l0 = f0 = 0, dense Y."""
    rng = np.random.default_rng(seed)
    Ftrue = np.zeros((p, K))
    Ftrue[0:20, 0] = 1; Ftrue[20:40, 1] = 2; Ftrue[40:60, 2] = 3
    Ltrue = rng.normal(size=(N, K))
    Y = rng.poisson(np.exp(Ltrue @ Ftrue.T)).astype(np.float64)
    return dict(Y=np.asfortranarray(Y), Factor=Ftrue, Loading=Ltrue, L0=np.ones(N), F0=np.ones(p),
                log_offset=0.0)


def solver_state(sim, start="cold", noise=0.2, var_type="by_col", seed=999):
    """Inputs of one VGA step as ebpmf_log would hold them.

    EL = [log l0 + c, 1, L + N(0,noise^2)], EF = [1, log f0, F + N(0,noise^2)]
    (K_tot = K+2); M0 = Beta ("warm") or log(Y+0.5) ("cold");
    sigma2 ~ U(0.05,1) of the var_type's length; R2 ~ U(0, n|p|n*p).
    """
    rng = np.random.default_rng(seed)
    Y = sim["Y"]; n, p = Y.shape
    L = sim["Loading"]; F = sim["Factor"]
    EL = np.column_stack([np.log(sim["L0"]) + sim["log_offset"], np.ones(n),
                          L + noise * rng.normal(size=L.shape)])
    EF = np.column_stack([np.ones(p), np.log(sim["F0"]), F + noise * rng.normal(size=F.shape)])
    Beta = EL @ EF.T
    M0 = Beta.copy() if start == "warm" else np.log(Y + 0.5)
    ns = {"by_col": p, "by_row": n, "constant": 1}[var_type]
    sigma2 = rng.uniform(0.05, 1.0, size=ns)
    R2 = rng.uniform(0.0, {"by_col": n, "by_row": p, "constant": n * p}[var_type], size=ns)
    return dict(M0=np.asfortranarray(M0), EL=np.asfortranarray(EL), EF=np.asfortranarray(EF),
                sigma2=sigma2, R2=R2, var_type=var_type)


def to_dgCMatrix(Y):
    """Dense -> the three dgCMatrix slots (@p int32, @i int32 0-based sorted, @x double)."""
    Yc = sp.csc_matrix(Y)
    Yc.sort_indices()
    return Yc.indptr.astype(np.int32), Yc.indices.astype(np.int32), Yc.data.astype(np.float64)

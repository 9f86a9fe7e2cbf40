"""Host-side statistics around the hot path: LDSC regression weights
(kgwas/utils.py:397-434), evaluation metrics (utils.py:41-45), Storey-Tibshirani p-value
re-weighting (kgwas/eval_utils.py:509-596) and the calibration search (eval_utils.py:11-28).
Plain numpy / pandas, same formulas; nothing here touches the GPU."""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy import interpolate
from scipy.stats import pearsonr


def ldsc_regression_weights(ld, w_ld, N, M, hsq, intercept=None, ii=None):
    """w = 1 / (2 (1 + hsq*N*ld/M)^2) * 1 / w_ld   with ld, w_ld floored at 1 (utils.py:397-434)."""
    M = float(M)
    intercept = 1 if intercept is None else intercept
    hsq = min(max(hsq, 0.0), 1.0)
    ld = np.fmax(ld, 1.0)
    w_ld = np.fmax(w_ld, 1.0)
    c = hsq * N / M
    het_w = 1.0 / (2 * np.square(intercept + np.multiply(c, ld)))
    return np.multiply(het_w, 1.0 / w_ld)


def compute_metrics(results, binary=False, coverage=None, uncertainty_reg=1, loss_fct=None):
    pred, truth = np.asarray(results["pred"], dtype=np.float64), np.asarray(results["truth"], dtype=np.float64)
    return {"mse": float(np.mean((pred - truth) ** 2)), "pearsonr": pearsonr(pred, truth)[0]}


def storey_pi_estimator(pvalue: np.ndarray) -> float:
    """Storey & Tibshirani (2003) pi0 of one bin: cubic spline of #(p > lambda) / (m (1 - lambda)) at the last lambda."""
    total = float(len(pvalue))
    lam = np.arange(0.05, 0.95, 0.05)
    pi0 = np.array([(pvalue > l).sum() / (total * (1 - l)) for l in lam])
    ok = np.isfinite(pi0)
    lam, pi0 = lam[ok], pi0[ok]
    est = float(interpolate.CubicSpline(lam, pi0)(lam[-1]))
    return min(est, 1.0)


def storey_ribshirani_integrate(gwas_data: pd.DataFrame, column: str = "pred", num_bins: int = 100) -> np.ndarray:
    """Bin SNPs by ``column`` quantiles, estimate pi0 per bin, weight p-values by (1-pi0)/pi0 normalised to mean 1."""
    num_bins = float(num_bins)
    q = gwas_data[column].quantile(np.arange(0, 1 + 1 / (num_bins + 1), 1 / num_bins))
    q.iloc[0] = q.iloc[0] - 1          # widen the outer edges (labels 0.0 and 1.0 of the quantile index,
    q.iloc[-1] = q.iloc[-1] + 1        # eval_utils.py:548-549) so that every value falls inside a bin
    q = q.drop_duplicates()
    n_bins = len(q) - 1
    gwas_data["bin_number"] = pd.cut(gwas_data[column], q, labels=np.arange(n_bins))
    if gwas_data["P"].min() < 0 or gwas_data["P"].max() > 1:
        print("detected p-values < 0 or > 1, please double check. we clipped it to 0-1 for now...")
        gwas_data["P"] = gwas_data["P"].clip(lower=0, upper=1)
    pi0 = np.full(len(gwas_data), np.nan)
    bins = gwas_data["bin_number"].to_numpy()
    pvals = gwas_data["P"].to_numpy()
    for i in range(n_bins):
        idx = bins == i
        if idx.any():
            pi0[idx] = min(max(storey_pi_estimator(pvals[idx]), 1e-5), 1 - 1e-5)
    gwas_data["pi0"] = pi0
    weights = (1 - pi0) / pi0
    weights = weights / np.nanmean(weights)
    gwas_data["weights"] = weights
    pw = pvals / weights
    over = pw > 1
    pw[over] = pvals[over]              # keep the original p-value when the weighted one exceeds 1
    pw[np.isnan(pw)] = 1
    gwas_data["P_weighted"] = pw
    return pw


def find_closest_x(df_pred: pd.DataFrame, lower_bound=0, upper_bound=200, tolerance=0.01):
    """Bisection for the scale s.t. #{1e-3 < s*P_weighted < 1e-2} matches #{1e-3 < P < 1e-2} (eval_utils.py:11-28)."""
    upper, lower = 1e-2, 1e-3
    pw, p = df_pred.P_weighted.values, df_pred.P.values
    res2 = int(((p < upper) & (p > lower)).sum())
    mid = (lower_bound + upper_bound) / 2
    while lower_bound <= upper_bound:
        mid = (lower_bound + upper_bound) / 2
        res1 = int(((pw * mid < upper) & (pw * mid > lower)).sum())
        result = res1 / res2
        if abs(result - 1) < tolerance:
            return mid
        if result > 1:
            lower_bound = mid + tolerance
        else:
            upper_bound = mid - tolerance
    return mid

"""Golden vectors for step 2's efficient-resampling exact test, produced by the reference's OWN compiled code.

    make -C oracle          # builds oracle/_ref/libskat_exact_ref.so from /root/reference/src/SAIGE/src/Binary_*.cpp
    python tests/golden/make_er_golden.py

Writes tests/golden/er_golden.json.  Each case holds the inputs of one variant (genotype / fitted probability / residual of
its k carriers, mean fitted probability of everybody else, n, ncase) and what the reference returns for it:
`prob` from GetProb (Binary_global.cpp:113-119 -> HyperGeo::Run / Get_lprob) and `pvalue` = pval - pval_same / 2 from
SKAT_Exact (Binary_global.cpp:69-80 -> ComputeExact::Init / Run / GetPvalues).  The Armadillo glue around those two calls
(SKATExactBin_Work and SKATExactBin_ComputeProb_Group, ER_binary_func.cpp:23-85, 186-278) needs RcppArmadillo and is
restated below, call for call, to build their arguments.  The reference library runs in its _STAND_ALONE_ build, whose
lCombinations sums logarithms instead of calling R's lchoose (Binary_HyperGeo.cpp:172-190): same value to ~1e-12.
"""
import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "..", "oracle", "_ref", "libskat_exact_ref.so")


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def reference_er(lib, g1, p1, res1, p2mean, n, ncase, epsilon=1e-6):
    k = len(g1)
    # SKATExactBin_ComputeProb_Group (ER_binary_func.cpp:23-85)
    p1c = np.where(p1 >= 1, 0.999, p1)
    group, weight = [], []
    for i in range(10):
        a1, a2 = i / 10.0, (i + 1) / 10.0
        sel = (p1c >= a1) & ((p1c < a2) if i + 1 < 10 else (p1c <= a2))
        if sel.any():
            pm = p1c[sel].mean()
            weight.append(pm / (1 - pm))
            group.append(int(sel.sum()))
    p2odd = p2mean / (1 - p2mean)
    weight.append(p2odd)
    weight = np.array(weight) / p2odd
    group.append(n - k)
    group = np.array(group, dtype=np.int32)
    prob = np.zeros(k + 1)
    lib.ref_get_prob(k, len(group), int(ncase), _ip(group), _dp(weight), _dp(prob))
    prob_in = prob.copy()
    # SKATExactBin_ComputProb_New (:113-143) in the all-exact regime, SKATExactBin_Work (:186-278)
    total_k = np.array([int(round(_choose(k, i))) for i in range(k + 1)], dtype=np.int32)
    total = int(total_k.sum())
    is_exact = np.ones(k + 1, dtype=np.int32)
    Z0 = np.ascontiguousarray(g1 * (-p1))
    Z1 = np.ascontiguousarray(g1 * (1 - p1))
    odds = np.ascontiguousarray(p1 / (1 - p1))
    p1_adj = np.ascontiguousarray(p1 / p1.mean())
    resarray = np.array(sorted(np.nonzero(res1 > 0)[0]), dtype=np.int32)
    nres_k = np.array([len(resarray)], dtype=np.int32)
    if len(resarray) == 0:
        resarray = np.zeros(1, dtype=np.int32)
    pval, pval1, minp = np.zeros(1), np.zeros(1), np.zeros(1)
    lib.ref_skat_exact(_ip(resarray), 1, _ip(nres_k), _dp(Z0), _dp(Z1), k, 1, total, _ip(total_k), _dp(prob), _dp(odds),
                       _dp(p1_adj), _ip(is_exact), _dp(pval), _dp(pval1), _dp(minp), 1, ctypes.c_double(epsilon))
    return prob_in, float(pval[0] - pval1[0] / 2)


def _choose(n, r):
    from math import comb
    return comb(n, r)


def main():
    lib = ctypes.CDLL(LIB)
    rng = np.random.default_rng(20261017)
    cases = []
    shapes = [(1000, 100), (1000, 500), (5000, 37), (200000, 20000), (200000, 100000), (50, 10), (30, 25)]
    for t in range(160):
        n, ncase = shapes[t % len(shapes)]
        k = 1 + t % 10 if t < 120 else 1 + t % 4
        kind = t % 5
        if kind == 0:
            p1 = rng.uniform(0.001, 0.2, size=k)                 # the usual unbalanced case-control fit
        elif kind == 1:
            p1 = rng.uniform(0.02, 0.98, size=k)                 # every probability class
        elif kind == 2:
            p1 = np.full(k, rng.uniform(0.01, 0.5))              # one class, tied odds
        elif kind == 3:
            p1 = rng.choice([0.1, 0.2, 0.35, 0.9, 1.0 - 1e-9], size=k)    # class boundaries
        else:
            p1 = rng.uniform(0.0005, 0.02, size=k)
        g1 = rng.choice([1.0, 2.0], size=k, p=[0.85, 0.15])
        if g1.sum() > 10:
            g1[:] = 1.0
        y1 = (rng.uniform(size=k) < np.clip(3 * p1, 0.05, 0.9)).astype(np.float64)
        if y1.sum() > ncase:
            y1[:] = 0
        res1 = y1 - p1
        p2mean = float(rng.uniform(0.01, 0.6)) if kind != 0 else ncase / n
        prob, pv = reference_er(lib, g1, p1, res1, p2mean, n, ncase)
        cases.append(dict(n=n, ncase=ncase, g1=g1.tolist(), p1=p1.tolist(), res1=res1.tolist(), p2mean=p2mean,
                          prob=prob.tolist(), pvalue=pv))
    with open(os.path.join(HERE, "er_golden.json"), "w") as f:
        json.dump(dict(source="oracle/_ref/libskat_exact_ref.so = /root/reference/src/SAIGE/src/Binary_{ComputeExact,HyperGeo,"
                              "global}.cpp compiled with -D_STAND_ALONE_ (oracle/Makefile)", epsilon=1e-6, cases=cases), f, indent=0)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()

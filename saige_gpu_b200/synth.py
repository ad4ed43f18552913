"""Synthetic workload of SURVEY.md 8(d): per-marker A1 frequency f ~ U(0.05, 0.5), genotypes Binomial(2, f) i.i.d.
from a counter-based integer hash (so any shard can be generated independently, on device, bit-identically to the
oracle's generator), 22 contiguous chromosome blocks with human-like proportions, logistic phenotype with two
covariates.  Bench/test input only -- not part of the drop-in boundary."""
import numpy as np

# approximate share of the autosomal genome per chromosome 1..22
_CHR_FRAC = np.array([8.2, 8.0, 6.6, 6.3, 6.0, 5.7, 5.3, 4.8, 4.6, 4.5, 4.5, 4.4, 3.8, 3.6, 3.4, 3.0, 2.8, 2.7, 2.0, 2.1,
                      1.6, 1.7])


def thresholds(M, seed):
    """f and the integer thresholds: genotype = (u >= t0) + (u >= t1) for u uniform on [0, 2^32)."""
    rng = np.random.default_rng(seed)
    f = rng.uniform(0.05, 0.5, size=M)
    t0 = np.floor((1 - f) ** 2 * 4294967296.0).astype(np.uint64).clip(0, 4294967295).astype(np.uint32)
    t1 = np.floor((1 - f * f) * 4294967296.0).astype(np.uint64).clip(0, 4294967295).astype(np.uint32)
    return f, t0, t1


def chromosomes(M):
    """Chromosome label (1..22) of each of M markers, contiguous blocks."""
    edges = np.floor(np.cumsum(_CHR_FRAC) / _CHR_FRAC.sum() * M).astype(np.int64)
    edges[-1] = M
    chrs = np.zeros(M, dtype=np.int64)
    lo = 0
    for c, hi in enumerate(edges):
        chrs[lo:hi] = c + 1
        lo = hi
    return chrs


def phenotype(N, seed, prevalence=0.1, gterm=None):
    """x1 ~ N(0,1), x2 ~ Bernoulli(0.5), binary y with the given prevalence; `gterm` is the polygenic part of the
    liability (a combination of standardised genotypes), unstructured noise when not given."""
    rng = np.random.default_rng(seed + 1)
    x1 = rng.normal(size=N)
    x2 = rng.integers(0, 2, size=N).astype(np.float64)
    if gterm is None:
        gterm = rng.normal(scale=0.6, size=N)
    eta = np.log(prevalence / (1 - prevalence)) + 0.5 * x1 + 0.3 * x2 + gterm
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-eta))).astype(np.float64)
    yq = 0.5 * x1 + 0.3 * x2 + gterm + rng.normal(size=N)
    X = np.column_stack([np.ones(N), x1, x2])
    return y, yq, X


def raw_bed(n_samples, n_markers, seed, miss_rate=0.0, n_classes=32, pool_bytes=1 << 20):
    """Body of a PLINK .bed (SNP-major, ceil(n/4) bytes per marker, no magic bytes) for throughput measurements of the
    ingest and step-2 paths: markers fall into `n_classes` A1-frequency classes on [0.05, 0.5]; each class owns a pool
    of bytes whose four 2-bit codes are i.i.d. Hardy-Weinberg draws (00 = hom A1, 10 = het, 11 = hom A2, 01 = missing
    at `miss_rate`), and a marker row is a slice of its class pool at a random offset.  Fast (a memcpy per marker) and
    statistically plain; parity tests use the oracle's generator instead."""
    rng = np.random.default_rng(seed)
    B0 = (n_samples + 3) // 4
    pool_bytes = max(pool_bytes, 2 * B0)
    fs = np.linspace(0.05, 0.5, n_classes)
    codes = np.arange(256)
    pools = []
    for f in fs:
        p = np.array([f * f, 0.0, 2 * f * (1 - f), (1 - f) ** 2]) * (1.0 - miss_rate)      # code 0, 1, 2, 3
        p[1] = miss_rate
        pb = np.ones(256)
        for j in range(4):
            pb *= p[(codes >> (2 * j)) & 3]
        pools.append(rng.choice(256, size=pool_bytes, p=pb / pb.sum()).astype(np.uint8))
    bed = np.empty(B0 * n_markers, dtype=np.uint8)
    cls = rng.integers(0, n_classes, size=n_markers)
    off = rng.integers(0, pool_bytes - B0, size=n_markers)
    for m in range(n_markers):
        bed[m * B0:(m + 1) * B0] = pools[cls[m]][off[m]:off[m] + B0]
    return bed

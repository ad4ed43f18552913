"""Helpers shared by test modules (not collected: no test_ prefix)."""
import numpy as np


def write_plink(prefix, bed_body, n_fam, n_bim, chrs=None):
    """A PLINK fileset around a SNP-major .bed body (tests build synthetic cohorts with the oracle's generator)."""
    with open(prefix + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        f.write(np.ascontiguousarray(bed_body, dtype=np.uint8).tobytes())
    with open(prefix + ".bim", "w") as f:
        for m in range(n_bim):
            f.write("%d\tsnp%d\t0\t%d\tA\tC\n" % (1 if chrs is None else chrs[m], m, m + 1))
    with open(prefix + ".fam", "w") as f:
        for i in range(n_fam):
            f.write("f%d i%d 0 0 1 -9\n" % (i, i))

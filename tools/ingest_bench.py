"""setgeno throughput from a host-resident PLINK .bed body (and from a .bed file): GB/s of raw genotype bytes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import oracle as O
from saige_gpu_b200 import SaigeB200
N0, M0 = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50_000, 100_000)
t = time.time(); bed = O.synth_bed(N0, M0, seed=3, miss_rate=0.01); print("synth %.1fs, %.2f GB" % (time.time() - t, bed.nbytes / 1e9))
g = SaigeB200(); g.setminMAFforGRM(0.01); g.setmaxMissingRateforGRM(0.15)
sub = np.arange(1, N0 + 1); ind = np.ones(N0, np.uint8)
for rep in range(2):
    t = time.time(); g.setgeno_mem(bed, N0, M0, sub, ind); dt = time.time() - t
    print("setgeno_mem: %.3f s  -> %.2f GB/s of .bed   (M=%d of %d pass QC)" % (dt, bed.nbytes / dt / 1e9, g.M, M0))
path = "/tmp/ingest_test"
with open(path + ".bed", "wb") as f:
    f.write(bytes([0x6C, 0x1B, 0x01])); f.write(bed.tobytes())
with open(path + ".bim", "w") as f:
    f.write("".join("1\trs%d\t0\t%d\tA\tG\n" % (i, i) for i in range(M0)))
with open(path + ".fam", "w") as f:
    f.write("".join("f%d i%d 0 0 0 -9\n" % (i, i) for i in range(N0)))
t = time.time(); g.setgeno(path + ".bed", path + ".bim", path + ".fam", sub, ind); dt = time.time() - t
print("setgeno (file, page cache warm): %.3f s -> %.2f GB/s" % (dt, bed.nbytes / dt / 1e9))
# a phenotyped subset in shuffled order exercises the gather path of the repack kernel
rng = np.random.default_rng(0); keep = np.sort(rng.choice(N0, N0 * 9 // 10, replace=False)); sub2 = rng.permutation(keep) + 1
ind2 = np.zeros(N0, np.uint8); ind2[keep] = 1
t = time.time(); g.setgeno_mem(bed, N0, M0, sub2, ind2); dt = time.time() - t
print("setgeno_mem (90%% subset, shuffled): %.3f s -> %.2f GB/s" % (dt, bed.nbytes / dt / 1e9))

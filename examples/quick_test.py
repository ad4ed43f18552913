"""The reference's bundled quick test, files in / files out, on one B200 (needs a CUDA device; no CPU fallback).

    python __graft_entry__.py          # build libsaige_b200.so
    python examples/quick_test.py      # step 1 on the 1000-sample x 10k-marker set, step 2 on the 100-marker files

Same calls and argument names as the reference's extdata/step1_fitNULLGLMM.R and extdata/step2_SPAtests.R
(docs: /root/reference/docs/installation.md:110, extdata/cmd.sh); the inputs are the fixtures under tests/golden/."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from saige_gpu_b200 import SaigeB200, SPAGMMATtest, fitNULLGLMM  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
out = tempfile.mkdtemp(prefix="saige_b200_")
g = SaigeB200(device=0)
fit = fitNULLGLMM(g, plinkFile=os.path.join(G, "grm10k"), phenoFile=os.path.join(G, "pheno_1000samples.txt"), phenoCol="y_binary",
                  covarColList=["x1", "x2"], sampleIDColinphenoFile="IID", traitType="binary", outputPrefix=os.path.join(out, "example"),
                  LOCO=False, nrun=30, IsOverwriteVarianceRatioFile=True)
print("step 1: tau =", fit["modglmm"]["theta"], " variance ratio =", fit["varianceRatio"], " ->", fit["modelFile"], fit["varRatioFile"])
p = os.path.join(G, "step2_100markers")
for label, src in (("PLINK", dict(bedFile=p + ".bed", bimFile=p + ".bim", famFile=p + ".fam")),
                   ("VCF", dict(vcfFile=p + ".vcf.gz", vcfField="GT"))):
    n = SPAGMMATtest(g, GMMATmodelFile=fit["modelFile"], varianceRatioFile=fit["varRatioFile"], LOCO=False, min_MAC=20,
                     SAIGEOutputFile=os.path.join(out, "step2_%s.txt" % label), is_Firth_beta=True, pCutoffforFirth=0.05,
                     return_rows=False, **src)
    print("step 2 (%s): %d variants tested -> %s" % (label, n, os.path.join(out, "step2_%s.txt" % label)))
g.close()

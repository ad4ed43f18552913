"""Copies the small reference DATA fixtures this repo's tests need into tests/golden/ (the GPU box has no
/root/reference) and records golden values derived from them.  Run once in the build container:

    python tests/golden/make_golden.py

Fixtures (data, not source) come from /root/reference/src/SAIGE/extdata/input:
  plinkforGRM_1000samples_10kMarkers.{bed,bim,fam,frq}   -- pins decode / allele counts / MAF QC
  nfam_100_nindep_0_step1_includeMoreRareVariants_poly_22chr_random1000.{bed,bim,fam} -- 22-chromosome LOCO set
  pheno_1000samples.txt_withdosages_withBothTraitTypes.txt -- phenotypes/covariates of the bundled example
  genotype_100markers.{bed,bim,fam} + ../output/example_binary.{rda,varianceRatio.txt} + ../output/
  genotype_100markers_marker_plink.txt -- step-2 inputs and the reference's golden result table (32 variants);
  genotype_100markers_marker_vcf.txt (LOCO off) and genotype_100markers_marker_bgen.txt (alleles swapped) -- two more
"""
import os
import shutil
import sys

REF = "/root/reference/src/SAIGE/extdata/input"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = [
    ("plinkforGRM_1000samples_10kMarkers.bed", "grm10k.bed"),
    ("plinkforGRM_1000samples_10kMarkers.bim", "grm10k.bim"),
    ("plinkforGRM_1000samples_10kMarkers.fam", "grm10k.fam"),
    ("plinkforGRM_1000samples_10kMarkers.frq", "grm10k.frq"),
    ("nfam_100_nindep_0_step1_includeMoreRareVariants_poly_22chr_random1000.bed", "chr22_1000.bed"),
    ("nfam_100_nindep_0_step1_includeMoreRareVariants_poly_22chr_random1000.bim", "chr22_1000.bim"),
    ("nfam_100_nindep_0_step1_includeMoreRareVariants_poly_22chr_random1000.fam", "chr22_1000.fam"),
    ("pheno_1000samples.txt_withdosages_withBothTraitTypes.txt", "pheno_1000samples.txt"),
    # step 2: inputs + the reference's own golden result table
    ("genotype_100markers.bed", "step2_100markers.bed"),
    ("genotype_100markers.bim", "step2_100markers.bim"),
    ("genotype_100markers.fam", "step2_100markers.fam"),
    ("../output/example_binary.rda", "example_binary.rda"),
    ("../output/example_binary.varianceRatio.txt", "example_binary.varianceRatio.txt"),
    ("../output/genotype_100markers_marker_plink.txt", "step2_100markers_golden.txt"),
    # two more golden tables of the same 100 markers and model: the run without LOCO (produced from the VCF copy of the
    # genotypes) and the run with the alleles the other way round (produced from the BGEN copy: Allele2 is the major
    # allele there, AF_Allele2 up to 0.99, so every row goes through the reference's flip branch)
    ("../output/genotype_100markers_marker_vcf.txt", "step2_100markers_golden_noLOCO.txt"),
    ("../output/genotype_100markers_marker_bgen.txt", "step2_100markers_golden_flipped.txt"),
    # conditional analysis of the same markers on rs13 and rs79 (--condition=1:13:A:C,1:79:A:C, extdata/cmd.sh:128-139)
    ("../output/genotype_100markers_marker_vcf_cond.txt", "step2_100markers_golden_cond.txt"),
    # the VCF and BGEN copies themselves (the files those two tables were produced from) + two small files with missing calls
    ("genotype_100markers.vcf.gz", "step2_100markers.vcf.gz"),
    ("genotype_100markers.bgen", "step2_100markers.bgen"),
    ("genotype_10markers.missingness.vcf.gz", "missing_10markers.vcf.gz"),
    ("genotype_10markers.missingness.bgen", "missing_10markers.bgen"),
    ("dosage_10markers.vcf.gz", "dosage_10markers.vcf.gz"),
    # a genome-wide-significant variant (p = 3.5e-7 after SPA): one-marker VCF of 10,000 samples, its model and result
    ("nfam_1000_MAF0.2_nMarker1_nseed200.vcf", "positive_signal_1marker.vcf"),
    ("../output/example_binary_positive_signal.rda", "positive_signal.rda"),
    ("../output/example_binary_positive_signal.varianceRatio.txt", "positive_signal.varianceRatio.txt"),
    ("../output/example_binary_positive_signal.assoc.step2.txt", "positive_signal_golden.txt"),
]

if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    for src, dst in FILES:
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        os.chmod(os.path.join(HERE, dst), 0o644)
        print("copied", src, "->", dst)
    # plink2's allele counts of the 22-chromosome set (128,868 markers), restricted to the 1000 markers of chr22_1000.bim:
    # an independent tool's counts for a second genotype file (the .frq of the 10k set pins the first)
    want = [l.split()[1] for l in open(os.path.join(HERE, "chr22_1000.bim"))]
    rows = {}
    for l in open(os.path.join(REF, "nfam_100_nindep_0_step1_includeMoreRareVariants_poly_22chr.acount")):
        t = l.split()
        if t[0].startswith("#"):
            header = l
        else:
            rows[t[1]] = l
    with open(os.path.join(HERE, "chr22_1000.acount"), "w") as f:
        f.write(header)
        for mid in want:
            f.write(rows[mid])
    print("wrote chr22_1000.acount (%d markers)" % len(want))

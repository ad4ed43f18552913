#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle/): cuts the reference's solver-layer functions out of
/root/reference/src/SAIGE/src/SAIGE_fitGLMM_fast.cpp AT BUILD TIME and writes them, unmodified, to oracle/_ref/fg_extract.inc
(git-ignored; reference sources are never copied into the repository).  oracle/ref_fg/fg_ref_shim.cpp includes that file and
compiles it against oracle/ref_fg/mini_arma.h.  The whole file cannot be built here (R, Rcpp, RcppParallel, Armadillo, MPI,
cuBLAS), these functions can: they only need vector algebra, the genotype object's diagonal and the GRM product, which the shim
supplies.

usage: extract_ref.py <SAIGE_fitGLMM_fast.cpp> <out.inc> [solver|cpu]"""
import re
import sys

# in dependency order (the file has no prototypes for them); each is the first line of the definition
WANT = [
    r"arma::fvec getDiagOfSigma\(arma::fvec& wVec, arma::fvec& tauVec\)\{",
    r"arma::fvec getDiagOfSigma_LOCO\(arma::fvec& wVec, arma::fvec& tauVec\)\{",
    r"arma::fcolvec getCrossprod\(arma::fcolvec& bVec, arma::fvec& wVec, arma::fvec& tauVec\)\{",
    r"arma::fcolvec getCrossprod_LOCO\(arma::fcolvec& bVec, arma::fvec& wVec, arma::fvec& tauVec\)\{",
    r"arma::fvec getPCG1ofSigmaAndVector\(arma::fvec& wVec,  arma::fvec& tauVec, arma::fvec& bVec, int maxiterPCG, float tolPCG\)\{",
    r"arma::fvec getPCG1ofSigmaAndVector_LOCO\(arma::fvec& wVec,  arma::fvec& tauVec, arma::fvec& bVec, int maxiterPCG, float tolPCG\)\{",
    r"Rcpp::NumericVector nb\(int n\) \{",
    r"float calCV\(arma::fvec& xVec\)\{",
    r"float GetTrace\(arma::fmat Sigma_iX, arma::fmat& Xmat, arma::fvec& wVec, arma::fvec& tauVec, arma::fmat& cov1, int nrun, int maxiterPCG, float tolPCG, float traceCVcutoff\)\{",
    r"Rcpp::List getCoefficients\(arma::fvec& Yvec, arma::fmat& Xmat, arma::fvec& wVec,  arma::fvec& tauVec, int maxiterPCG, float tolPCG\)\{",
    r"Rcpp::List getCoefficients_LOCO\(arma::fvec& Yvec, arma::fmat& Xmat, arma::fvec& wVec,  arma::fvec& tauVec, int maxiterPCG, float tolPCG\)\{",
    r"Rcpp::List getAIScore\(arma::fvec& Yvec, arma::fmat& Xmat, arma::fvec& wVec,  arma::fvec& tauVec,",
    r"Rcpp::List fitglmmaiRPCG\(arma::fvec& Yvec, arma::fmat& Xmat, arma::fvec &wVec,  arma::fvec &tauVec,",
    r"arma::fmat getSigma_X\(arma::fvec& wVec, arma::fvec& tauVec,arma::fmat& Xmat, int maxiterPCG, float tolPCG\)\{",
    r"arma::fvec  getSigma_G\(arma::fvec& wVec, arma::fvec& tauVec,arma::fvec& Gvec, int maxiterPCG, float tolPCG\)\{",
    r"arma::fvec GetTrace_q\(arma::fmat Sigma_iX, arma::fmat& Xmat, arma::fvec& wVec, arma::fvec& tauVec, arma::fmat& cov1,  int nrun, int maxiterPCG, float tolPCG, float traceCVcutoff\)\{",
    r"Rcpp::List getAIScore_q\(arma::fvec& Yvec, arma::fmat& Xmat, arma::fvec& wVec,  arma::fvec& tauVec,",
    r"Rcpp::List fitglmmaiRPCG_q\(arma::fvec& Yvec, arma::fmat& Xmat, arma::fvec &wVec,  arma::fvec &tauVec,",
]


# the genotype store and the CPU product (built only into libfg_refcpu.so): the whole genoClass, the OpenMP marker loop and the
# small exports that configure / query it
WANT_CPU = [
    r"class genoClass\{",
    r"arma::fvec parallelCrossProdOpenMP\(int startIndex, int endIndex, arma::fcolvec &bVec, int &count\)",
    r"arma::fvec parallelCrossProd\(arma::fcolvec & bVec\) \{",
    r"arma::fvec parallelCrossProd_full\(arma::fcolvec & bVec, int & markerNum\) \{",
    r"arma::fvec parallelCrossProd_LOCO\(arma::fcolvec & bVec\) \{",
    r"void setgeno\(std::string bedfile, std::string bimfile, std::string famfile, std::vector<int> & subSampleInGeno, std::vector<bool> & indicatorGenoSamplesWithPheno, float memoryChunk, bool isDiagofKinSetAsOne\)",
    r"arma::ivec Get_OneSNP_Geno\(int SNPIdx\)",
    r"arma::ivec Get_OneSNP_Geno_forVarRatio\(int SNPIdx\)",
    r"void setStartEndIndex\(int startIndex, int endIndex, int chromIndex\)\{",
    r"void setStartEndIndexVec\( arma::ivec & startIndex_vec,  arma::ivec & endIndex_vec\)\{",
    r"void setminMAFforGRM\(float minMAFforGRM\)\{",
    r"void setmaxMissingRateforGRM\(float maxMissingforGRM\)\{",
    r"void set_Diagof_StdGeno_LOCO\(\)\{",
    r"void setminMAC_VarianceRatio\(float t_minMACVarRatio, float t_maxMACVarRatio, bool t_isVarianceRatioinGeno\)\{",
]


def function_end(text, start):
    """index just past the brace that closes the first '{' at or after `start`; comments, strings and char literals skipped"""
    i, depth, n = start, 0, len(text)
    while i < n:
        c = text[i]
        if text.startswith("//", i):
            i = text.index("\n", i)
        elif text.startswith("/*", i):
            i = text.index("*/", i) + 2
            continue
        elif c == '"' or c == "'":
            q = c
            i += 1
            while text[i] != q:
                i += 2 if text[i] == "\\" else 1
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise SystemExit("unbalanced braces after offset %d" % start)


def main(src, out, which="solver"):
    text = open(src, encoding="utf-8", errors="replace").read()
    parts = []
    for pat in (WANT_CPU if which == "cpu" else WANT):
        hits = [m for m in re.finditer("^" + pat, text, flags=re.M)]
        if len(hits) != 1:
            raise SystemExit("expected exactly one definition matching %r, found %d" % (pat, len(hits)))
        s = hits[0].start()
        e = function_end(text, s)
        if pat.startswith("class "):
            e = text.index(";", e) + 1                          # the class definition ends with "};"
            tail = "\n// the reference's global instance (FG.cpp:1188)\ngenoClass geno;\n"
        else:
            tail = ""
        line0 = text.count("\n", 0, s) + 1
        parts.append("// ---- %s:%d-%d ----\n%s\n%s" % (src, line0, line0 + text.count("\n", s, e), text[s:e], tail))
    with open(out, "w") as f:
        f.write("// GENERATED at build time by oracle/ref_fg/extract_ref.py from the reference tree; not part of the repository.\n")
        f.write("\n".join(parts))
    print("extracted %d reference functions -> %s" % (len(parts), out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "solver")

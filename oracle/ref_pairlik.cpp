// TEST INFRASTRUCTURE — known-answer generator for SURVEY §8 row a19.
// Links against the UNMODIFIED reference objects (oracle/_ref/obj) and calls the reference's own
// calcLikelihood / calcLikelihoodConsensus (src/assembler/nuclassembleUtil.cpp:95-200, :203-374),
// getNuclRevFragment (:67-76), initDeamProbabilities (:821-1007) and getSeqErrorProf (:49-65) on seeded
// random read pairs.  Output (stdout) -> tests/golden/pairlik_kat.txt via oracle/make_pairlik_kat.py.
//
// line format (tab separated):
//   L <isRev> <qStart> <qEnd> <dbStart> <dbEnd> <maxAln> <query> <target as stored> <sLenNorm %.17g> <sRatio %.17g>
//   C <isRev> <qLen> <qStart> <qEnd> <dbStart> <dbEnd> <maxAln> <consensus (3*qLen)> <target as stored> <sLenNorm> <sRatio>
// coordinates are in the oriented frame the reference works in (target reverse-complemented when isRev).
#include "NucleotideMatrix.h"
#include "LocalParameters.h"
#include "nuclassembleUtil.h"
#include <cstdio>
#include <string>
#include <vector>

static unsigned long long g_s = 88172645463325252ull;
static unsigned rnd() { g_s ^= g_s << 13; g_s ^= g_s >> 7; g_s ^= g_s << 17; return (unsigned) (g_s >> 11); }
static unsigned rndn(unsigned n) { return rnd() % n; }

static std::string revcomp_plain(const std::string &s) {
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); i++) {
        char c = s[s.size() - 1 - i];
        r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
    return r;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_pairlik 5p.prof 3p.prof\n"); return 1; }
    LocalParameters::getLocalInstance();
    std::string nm = "nucleotide.out", md;
    { FILE *f = fopen("/root/reference/lib/mmseqs/data/nucleotide.out", "rb"); char b[4096]; size_t n = fread(b, 1, 4096, f); md.assign(b, n); fclose(f); }
    char *ser = BaseMatrix::serialize(nm, md);
    NucleotideMatrix subMat(ser, 1.0, 0.0);
    std::vector<substitutionRates> s5, s3;
    std::vector<diNucleotideProb> D(11), Drev;   // sized by the caller, correction.cpp:184-189
    initDeamProbabilities(argv[1], argv[2], s5, s3, D, Drev);
    diNucleotideProb seqErr;
    getSeqErrorProf(seqErr, 0.001);   // ancientReadsResults.cpp:172
    const float randAlnPenal = 0.85f, excessPenal = 0.0625f;   // LocalParameters.h:295-296
    const char *ACGT = "ACGT";
    for (int c = 0; c < 400; c++) {
        const bool cons = c >= 200;
        const bool isRev = rndn(2);
        const bool left = rndn(2);      // target overhangs the query's left end (qStart == 0, dbEnd == dbLen-1)
        unsigned qLen = 30 + rndn(120), aln = 12 + rndn(qLen - 12), over = 1 + rndn(40);
        unsigned dbLen = aln + over;
        if (dbLen < 11) dbLen = 11;
        std::string q(qLen, 'A');
        for (auto &ch : q) ch = ACGT[rndn(4)];
        // oriented target: the overlapping query part with damage-like and random substitutions, plus an overhang
        std::string tOri(dbLen, 'A');
        for (auto &ch : tOri) ch = ACGT[rndn(4)];
        unsigned qS = left ? 0 : qLen - aln, dS = left ? dbLen - aln : 0;
        for (unsigned i = 0; i < aln; i++) {
            char b = q[qS + i];
            unsigned r = rndn(100);
            const unsigned mis = (c % 3 == 0) ? 0 : (c % 3 == 1) ? 2 : 6, dam = (i < 5 || i + 5 >= aln) ? 25 : 2;
            if (r < mis) b = ACGT[rndn(4)];
            else if (r < mis + dam && b == 'C') b = 'T';
            else if (r < mis + 2 * dam && b == 'G') b = 'A';
            tOri[dS + i] = b;
        }
        if (c % 7 == 3) tOri[rndn(dbLen)] = 'N';
        if (c % 11 == 5) q[rndn(qLen)] = 'N';
        std::string tStored = isRev ? revcomp_plain(tOri) : tOri;
        // the reference's own orientation step
        std::string tUse = tStored;
        if (isRev) { char *r = getNuclRevFragment(&tStored[0], tStored.size(), &subMat); tUse.assign(r, tStored.size()); delete[] r; }
        scorePerRes sr;
        sr.r.qStartPos = qS; sr.r.qEndPos = qS + aln - 1; sr.r.dbStartPos = dS; sr.r.dbEndPos = dS + aln - 1;
        sr.r.qLen = qLen; sr.r.dbLen = dbLen; sr.r.alnLength = aln; sr.r.isRevToAlignment = isRev;
        // consensus mode counts every column where consensus and target are both called (up to dbLen); maxAln below that
        // count makes the reference's unsigned `excess` wrap — kept in a few cases on purpose
        unsigned maxAln = (cons && c % 4 != 1 ? dbLen : aln) + (c % 4 == 0 ? 0 : rndn(12));
        std::vector<diNucleotideProb> &tab = isRev ? Drev : D;   // r_s_pair, ancientReadsResults.cpp:60
        if (!cons) {
            calcLikelihood(sr, &q[0], tUse.c_str(), tab, maxAln, randAlnPenal, seqErr, excessPenal);
            printf("L\t%d\t%u\t%u\t%u\t%u\t%u\t%s\t%s\t%.17g\t%.17g\n", (int) isRev, qS, qS + aln - 1, dS, dS + aln - 1, maxAln,
                   q.c_str(), tStored.c_str(), sr.sLenNorm, sr.sRatio);
        } else {
            // consensus string as consensusCaller lays it out: 3*qLen, the query in the middle third, 'N' where no call
            std::string consensus(3 * qLen, 'N');
            for (unsigned i = 0; i < qLen; i++) consensus[qLen + i] = q[i];
            int lo = left ? (int) qLen - (int) over : (int) (2 * qLen), hi = left ? (int) qLen : (int) (2 * qLen + over);
            for (int i = lo; i < hi && i < (int) (3 * qLen); i++)   // overhang: mostly the target's own bases, some uncalled
                if (i >= 0 && rndn(10) < 8) consensus[i] = rndn(20) ? tOri[left ? i - lo : aln + (i - lo)] : ACGT[rndn(4)];
            if (c % 5 == 1) consensus[qLen + rndn(qLen)] = 'N';
            if (over > qLen) { c--; continue; }   // offset > qLen underflows consensusStart in the reference
            calcLikelihoodConsensus(sr, consensus, qLen, tUse.c_str(), tab, maxAln, randAlnPenal, seqErr, excessPenal);
            printf("C\t%d\t%u\t%u\t%u\t%u\t%u\t%u\t%s\t%s\t%.17g\t%.17g\n", (int) isRev, qLen, qS, qS + aln - 1, dS, dS + aln - 1, maxAln,
                   consensus.c_str(), tStored.c_str(), sr.sLenNorm, sr.sRatio);
        }
    }
    return 0;
}

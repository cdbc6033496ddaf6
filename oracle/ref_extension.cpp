// TEST INFRASTRUCTURE — known-answer generator for SURVEY §8(f) row N1 (scoring front half of ancient_read_assemble).
// Links against the UNMODIFIED reference objects (oracle/_ref/obj).  Per query of a sequence DB + alignment DB it walks the
// per-alignment glue of doNuclAssembly1 (src/assembler/ancientReadsResults.cpp:200-370: orientation, identity recompute,
// candidate filter — restated here line by line because it lives inside the reference's OpenMP loop) and calls the reference's
// own consensusCaller (nuclassembleUtil.cpp:570-702), updateSeqIdConsensusReads (:377-470), getNuclRevFragment (:67-76) and
// r_s_pair / calcLikelihoodConsensus (ancientReadsResults.cpp:48-70, nuclassembleUtil.cpp:203-374).
// Output (stdout) -> tests/golden/extension_kat.txt.gz via oracle/make_extension_kat.py:
//   Q <query index> <unsafe> <maxAlnLeft> <maxAlnRight> <consensus, 3*qLen letters>
//   A <line of the query> <candidate> <isRev> <qStart> <qEnd> <dbStart> <dbEnd> <seqId %.9g> <rySeqId %.9g> <scored> <sLenNorm %.17g> <sRatio %.17g>
#include "DBReader.h"
#include "LocalParameters.h"
#include "Matcher.h"
#include "NucleotideMatrix.h"
#include "nuclassembleUtil.h"
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

scorePerRes r_s_pair(Matcher::result_t res, std::string &consensus, char *targetSeq, unsigned int querySeqLen, std::vector<diNucleotideProb> &subDeamDiNuc,
                     std::vector<diNucleotideProb> &subDeamDiNucRev, unsigned int &maxLeft, unsigned int &maxRight, float randAlnPenal,
                     diNucleotideProb &seqErrMatch, float excessPenal);

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: ref_extension seqDB alnDB 5p.prof 3p.prof unsafe(0|1)\n"); return 1; }
    LocalParameters &par = LocalParameters::getLocalInstance();
    par.ancientUnsafe = atoi(argv[5]) != 0;
    par.seqIdThr = 0.9f;   // --min-seq-id 0.9 of the assembler workflow (ASSEMBLE_RESULT_PAR, tests/golden/cfg1_stage_argv.log)
    std::string nm = "nucleotide.out", md;
    { FILE *f = fopen("/root/reference/lib/mmseqs/data/nucleotide.out", "rb"); char b[4096]; size_t n = fread(b, 1, 4096, f); md.assign(b, n); fclose(f); }
    char *ser = BaseMatrix::serialize(nm, md);
    NucleotideMatrix *subMat = new NucleotideMatrix(ser, 1.0, 0.0);
    std::string seqPath = argv[1], alnPath = argv[2];
    DBReader<unsigned int> *sequenceDbr = new DBReader<unsigned int>(seqPath.c_str(), (seqPath + ".index").c_str(), 1, DBReader<unsigned int>::USE_DATA | DBReader<unsigned int>::USE_INDEX);
    sequenceDbr->open(DBReader<unsigned int>::NOSORT);
    DBReader<unsigned int> *alnReader = new DBReader<unsigned int>(alnPath.c_str(), (alnPath + ".index").c_str(), 1, DBReader<unsigned int>::USE_DATA | DBReader<unsigned int>::USE_INDEX);
    alnReader->open(DBReader<unsigned int>::NOSORT);
    std::unordered_map<char, int> ryMap = {{'A', 0}, {'C', 1}, {'G', 0}, {'T', 1}};
    std::vector<diNucleotideProb> subDeamDiNuc(11), subDeamDiNucRev;
    std::vector<substitutionRates> sub5p, sub3p;
    initDeamProbabilities(argv[3], argv[4], sub5p, sub3p, subDeamDiNuc, subDeamDiNucRev);
    diNucleotideProb seqErrMatch;
    getSeqErrorProf(seqErrMatch, 0.001);
    const float randAlnPenal = par.randomAlignPenal;
    const unsigned int thread_idx = 0;
    std::vector<Matcher::result_t> alignments;
    for (size_t id = 0; id < sequenceDbr->getSize(); id++) {
        unsigned int queryKey = sequenceDbr->getDbKey(id);
        char *querySeq = sequenceDbr->getData(id, thread_idx);
        unsigned int querySeqLen = sequenceDbr->getSeqLen(id);
        char *alnData = alnReader->getDataByDBKey(queryKey, thread_idx);
        alignments.clear();
        Matcher::readAlignmentResults(alignments, alnData);
        if (alignments.size() < 2) continue;
        std::vector<int> cand(alignments.size(), 0), scored(alignments.size(), 0);
        std::vector<double> sLen(alignments.size(), 0.0), sRat(alignments.size(), 0.0);
        // ancientReadsResults.cpp:200-236 (orientation)
        for (size_t alnIdx = 0; alnIdx < alignments.size(); alnIdx++) {
            unsigned int dbStartPos = alignments[alnIdx].dbStartPos, dbEndPos = alignments[alnIdx].dbEndPos;
            unsigned int qStartPos = alignments[alnIdx].qStartPos, qEndPos = alignments[alnIdx].qEndPos;
            const bool rightStart = dbStartPos == 0 && qEndPos == (querySeqLen - 1);
            const bool leftStart = qStartPos == 0 && dbEndPos == (alignments[alnIdx].dbLen - 1);
            if (!rightStart && !leftStart) continue;
            if (alignments[alnIdx].qStartPos > alignments[alnIdx].qEndPos) {
                std::swap(alignments[alnIdx].qStartPos, alignments[alnIdx].qEndPos);
                unsigned int d0 = alignments[alnIdx].dbStartPos;
                alignments[alnIdx].dbStartPos = alignments[alnIdx].dbLen - alignments[alnIdx].dbEndPos - 1;
                alignments[alnIdx].dbEndPos = alignments[alnIdx].dbLen - d0 - 1;
                alignments[alnIdx].isRevToAlignment = true;
            } else alignments[alnIdx].isRevToAlignment = false;
        }
        // :247-293 (identity recompute)
        for (unsigned int idx = 0; idx < alignments.size(); idx++) {
            unsigned int dbStartPos = alignments[idx].dbStartPos, dbEndPos = alignments[idx].dbEndPos;
            unsigned int qStartPos = alignments[idx].qStartPos, qEndPos = alignments[idx].qEndPos;
            const bool rightStart = dbStartPos == 0 && qEndPos == (querySeqLen - 1);
            const bool leftStart = qStartPos == 0 && dbEndPos == (alignments[idx].dbLen - 1);
            if (!rightStart && !leftStart) continue;
            unsigned int aln2updateId = sequenceDbr->getId(alignments[idx].dbKey);
            unsigned int aln2updateLen = sequenceDbr->getSeqLen(aln2updateId);
            if (aln2updateId == queryKey) continue;
            char *aln2updateSequ = sequenceDbr->getData(aln2updateId, thread_idx);
            std::string aln2updateSeq;
            if (alignments[idx].isRevToAlignment) {
                char *tmp = getNuclRevFragment(aln2updateSequ, aln2updateLen, subMat);
                aln2updateSeq = std::string(tmp, aln2updateLen);
                delete[] tmp;
            } else aln2updateSeq = std::string(aln2updateSequ, aln2updateLen);
            int idCnt = 0, idRyCnt = 0;
            for (int i = alignments[idx].qStartPos; i <= alignments[idx].qEndPos; i++) {
                idCnt += (querySeq[i] == aln2updateSeq[alignments[idx].dbStartPos + (i - alignments[idx].qStartPos)]) ? 1 : 0;
                idRyCnt += (ryMap[querySeq[i]] == ryMap[aln2updateSeq[alignments[idx].dbStartPos + (i - alignments[idx].qStartPos)]]) ? 1 : 0;
            }
            alignments[idx].seqId = static_cast<float>(idCnt) / alignments[idx].alnLength;
            alignments[idx].rySeqId = static_cast<float>(idRyCnt) / alignments[idx].alnLength;
        }
        // :294-315 (candidates)
        std::vector<Matcher::result_t> notContig; std::vector<size_t> origin;
        for (size_t alnIdx = 0; alnIdx < alignments.size(); alnIdx++) {
            unsigned int dbStartPos = alignments[alnIdx].dbStartPos, dbEndPos = alignments[alnIdx].dbEndPos;
            unsigned int qStartPos = alignments[alnIdx].qStartPos, qEndPos = alignments[alnIdx].qEndPos;
            const bool rightStart = dbStartPos == 0 && qEndPos == (querySeqLen - 1);
            const bool leftStart = qStartPos == 0 && dbEndPos == (alignments[alnIdx].dbLen - 1);
            if (!rightStart && !leftStart) continue;
            bool noOffset = (alignments[alnIdx].dbLen - alignments[alnIdx].alnLength) == 0;
            unsigned int targetId = sequenceDbr->getId(alignments[alnIdx].dbKey);
            bool isContig = sequenceDbr->getExtData(targetId);
            if (isContig == false && alignments[alnIdx].alnLength >= 30 && alignments[alnIdx].seqId >= par.seqIdThr && !noOffset) {
                notContig.push_back(alignments[alnIdx]); origin.push_back(alnIdx); cand[alnIdx] = 1;
            }
        }
        std::string consensus;
        unsigned int maxAlnLeft = 0, maxAlnRight = 0;
        consensus.assign(3 * querySeqLen, 'N');
        consensusCaller(consensus, notContig, sequenceDbr, querySeq, querySeqLen, queryKey, thread_idx, par, subMat);
        updateSeqIdConsensusReads(notContig, sequenceDbr, consensus, querySeqLen, thread_idx, subMat, maxAlnLeft, maxAlnRight);
        // :330-365 (scores)
        for (size_t alnIdx = 0; alnIdx < notContig.size(); alnIdx++) {
            unsigned int targetId = sequenceDbr->getId(notContig[alnIdx].dbKey);
            char *targetSeq = sequenceDbr->getData(targetId, thread_idx);
            unsigned int targetSeqLen = sequenceDbr->getSeqLen(targetId);
            bool del = false;
            if (notContig[alnIdx].isRevToAlignment) { targetSeq = getNuclRevFragment(targetSeq, targetSeqLen, subMat); del = true; }
            bool notInside = notContig[alnIdx].dbLen != notContig[alnIdx].alnLength;
            const bool rightStart = notContig[alnIdx].dbStartPos == 0;
            const bool leftStart = notContig[alnIdx].qStartPos == 0;
            const bool isNotIdentity = (notContig[alnIdx].dbKey != queryKey);
            if ((rightStart || leftStart) && notInside && isNotIdentity && notContig[alnIdx].rySeqId >= par.rySeqIdThr && notContig[alnIdx].seqId >= par.seqIdThr) {
                scorePerRes toAdd = r_s_pair(notContig[alnIdx], consensus, targetSeq, querySeqLen, subDeamDiNuc, subDeamDiNucRev, maxAlnLeft, maxAlnRight, randAlnPenal,
                                             seqErrMatch, par.excessPenal);
                scored[origin[alnIdx]] = 1; sLen[origin[alnIdx]] = toAdd.sLenNorm; sRat[origin[alnIdx]] = toAdd.sRatio;
            }
            if (del) delete[] targetSeq;
            alignments[origin[alnIdx]] = notContig[alnIdx];   // the consensus-based identities
        }
        printf("Q\t%zu\t%d\t%u\t%u\t%s\n", id, (int) par.ancientUnsafe, maxAlnLeft, maxAlnRight, consensus.c_str());
        for (size_t a = 0; a < alignments.size(); a++)
            printf("A\t%zu\t%d\t%d\t%d\t%d\t%d\t%d\t%.9g\t%.9g\t%d\t%.17g\t%.17g\n", a, cand[a], (int) alignments[a].isRevToAlignment, alignments[a].qStartPos,
                   alignments[a].qEndPos, alignments[a].dbStartPos, alignments[a].dbEndPos, alignments[a].seqId, alignments[a].rySeqId, scored[a], sLen[a], sRat[a]);
    }
    fprintf(stderr, "thresholds: seqIdThr %g rySeqIdThr %g likelihoodThreshold %g minCovSafe %u randomAlignPenal %g excessPenal %g\n", par.seqIdThr, par.rySeqIdThr,
            par.likelihoodThreshold, par.minCovSafe, par.randomAlignPenal, par.excessPenal);
    return 0;
}

// TEST INFRASTRUCTURE — known-answer generator for SURVEY §8(f) row N3 (scoring front half of ancient_contig_merge).
// Compiled against the UNMODIFIED reference: links oracle/_ref/obj and includes src/assembler/ancientContigsResults.cpp where it
// lies (its object is left out of the link), because the Beta-binomial comparator and the queue type are file-local there.
// Per query of a sequence DB + alignment DB it walks the per-alignment glue of doNuclAssembly2
// (src/assembler/ancientContigsResults.cpp:187-262: orientation, identity recompute, merge-candidate filter, queue filter —
// restated here line by line because it lives inside the reference's OpenMP loop) and calls the reference's own consensusCaller
// (nuclassembleUtil.cpp:570-702), updateSeqIdConsensus (:704-790), ancientMatchCount (:1050-1180, with deamMatches :1011-1047) and
// pushes into the reference's own QueueByScoreNuclContigs (ancientContigsResults.cpp:24-72).
// Output (stdout) -> tests/golden/merge_kat_*.txt.gz via oracle/make_merge_kat.py:
//   Q <query index> <unsafe> <FNV-1a 64 of the 3*qLen consensus letters, hex>
//   A <line> <contig> <isRev> <qStart> <qEnd> <dbStart> <dbEnd> <seqId0 %.9g> <rySeqId0 %.9g> <seqId %.9g> <rySeqId %.9g> <alnLengthCons> <queued> <deamMatch %.9g>
//   P <lines of the query in the order the reference's priority queue pops them>
#include "/root/reference/src/assembler/ancientContigsResults.cpp"
#include <cstdio>

static unsigned long long fnv1a(const std::string &s) {
    unsigned long long h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: ref_merge seqDB alnDB 5p.prof 3p.prof unsafe(0|1) [mergeSeqIdThr]\n"); return 1; }
    LocalParameters &par = LocalParameters::getLocalInstance();
    par.ancientUnsafe = atoi(argv[5]) != 0;
    par.mergeSeqIdThr = argc > 6 ? (float) atof(argv[6]) : 0.99f;   // --min-merge-seq-id 0.99 (tests/golden/cfg1_stage_argv.log)
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
        const size_t nA = alignments.size();
        std::vector<int> isContig(nA, 0), queued(nA, 0);
        std::vector<float> sid0(nA), ry0(nA);
        std::vector<Matcher::result_t> contigs; std::vector<size_t> origin;
        // ancientContigsResults.cpp:187-236
        for (unsigned int idx = 0; idx < nA; idx++) {
            unsigned int aln2updateId = sequenceDbr->getId(alignments[idx].dbKey);
            unsigned int aln2updateLen = sequenceDbr->getSeqLen(aln2updateId);
            char *aln2updateSequ = sequenceDbr->getData(aln2updateId, thread_idx);
            std::string aln2updateSeq;
            if (alignments[idx].qStartPos > alignments[idx].qEndPos) {
                alignments[idx].isRevToAlignment = true;
                std::swap(alignments[idx].qStartPos, alignments[idx].qEndPos);
                unsigned int dbStartPos = alignments[idx].dbStartPos;
                alignments[idx].dbStartPos = alignments[idx].dbLen - alignments[idx].dbEndPos - 1;
                alignments[idx].dbEndPos = alignments[idx].dbLen - dbStartPos - 1;
                char *tmp = getNuclRevFragment(aln2updateSequ, aln2updateLen, subMat);
                aln2updateSeq = std::string(tmp, aln2updateLen);
                delete[] tmp;
            } else {
                aln2updateSeq = std::string(aln2updateSequ, aln2updateLen);
                alignments[idx].isRevToAlignment = false;
            }
            int idCnt = 0, idRyCnt = 0;
            for (int i = alignments[idx].qStartPos; i <= alignments[idx].qEndPos; i++) {
                idCnt += (querySeq[i] == aln2updateSeq[alignments[idx].dbStartPos + (i - alignments[idx].qStartPos)]) ? 1 : 0;
                idRyCnt += (ryMap[querySeq[i]] == ryMap[aln2updateSeq[alignments[idx].dbStartPos + (i - alignments[idx].qStartPos)]]) ? 1 : 0;
            }
            alignments[idx].seqId = static_cast<float>(idCnt) / alignments[idx].alnLength;
            alignments[idx].rySeqId = static_cast<float>(idRyCnt) / alignments[idx].alnLength;
            sid0[idx] = alignments[idx].seqId; ry0[idx] = alignments[idx].rySeqId;
            bool isNotIdentity = (queryKey != alignments[idx].dbKey);
            if (alignments[idx].seqId >= par.mergeSeqIdThr && alignments[idx].rySeqId >= par.rySeqIdThr && isNotIdentity) {
                contigs.push_back(alignments[idx]); origin.push_back(idx); isContig[idx] = 1;
            }
        }
        std::string consensus(3 * querySeqLen, 'N');
        consensusCaller(consensus, contigs, sequenceDbr, querySeq, querySeqLen, queryKey, thread_idx, par, subMat);
        updateSeqIdConsensus(contigs, sequenceDbr, consensus, querySeqLen, thread_idx, subMat);
        // :246-262
        unsigned int lefts = 0, rights = 0, others = 0;
        QueueByScoreNuclContigs alnQueue;
        for (size_t alnIdx = 0; alnIdx < contigs.size(); alnIdx++) {
            unsigned int minAlnLen = 500;
            minAlnLen = (contigs[alnIdx].alnLength < minAlnLen) ? std::min(minAlnLen, static_cast<unsigned int>(0.2 * contigs[alnIdx].dbLen)) : minAlnLen;
            if (contigs[alnIdx].seqId >= par.mergeSeqIdThr && contigs[alnIdx].rySeqId >= par.rySeqIdThr && contigs[alnIdx].alnLength >= minAlnLen) {
                std::vector<diNucleotideProb> subDeamDiNucRef = contigs[alnIdx].isRevToAlignment ? subDeamDiNucRev : subDeamDiNuc;
                float ancientMatches = ancientMatchCount(contigs[alnIdx], consensus, querySeqLen, subDeamDiNucRef, sequenceDbr, thread_idx, subMat, lefts, rights, others);
                contigs[alnIdx].deamMatch = ancientMatches;
                queued[origin[alnIdx]] = 1;
                Matcher::result_t tagged = contigs[alnIdx];
                tagged.score = (int) origin[alnIdx];   // the comparator never reads score: carries the line through the queue
                alnQueue.push(tagged);
            }
            alignments[origin[alnIdx]] = contigs[alnIdx];
        }
        printf("Q\t%zu\t%d\t%016llx\n", id, (int) par.ancientUnsafe, fnv1a(consensus));
        for (size_t a = 0; a < nA; a++)
            printf("A\t%zu\t%d\t%d\t%d\t%d\t%d\t%d\t%.9g\t%.9g\t%.9g\t%.9g\t%u\t%d\t%.9g\n", a, isContig[a], (int) alignments[a].isRevToAlignment, alignments[a].qStartPos,
                   alignments[a].qEndPos, alignments[a].dbStartPos, alignments[a].dbEndPos, sid0[a], ry0[a], alignments[a].seqId, alignments[a].rySeqId,
                   isContig[a] ? alignments[a].alnLengthCons : 0u, queued[a], queued[a] ? alignments[a].deamMatch : 0.f);
        printf("P");
        while (!alnQueue.empty()) { printf("\t%d", alnQueue.top().score); alnQueue.pop(); }
        printf("\n");
    }
    fprintf(stderr, "thresholds: mergeSeqIdThr %g rySeqIdThr %g minCovSafe %d\n", par.mergeSeqIdThr, par.rySeqIdThr, par.minCovSafe);
    return 0;
}

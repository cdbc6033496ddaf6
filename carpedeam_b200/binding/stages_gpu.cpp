// In-tree binding of the three hot-path modules of the reference to libcarpedeam_gpu.so (north_star: "host code stays C++ inside
// the existing carpedeam module ... and calls CUDA through a thin C-ABI layer so each stage is a drop-in for its CPU counterpart").
//
// This translation unit is compiled AGAINST THE REFERENCE'S OWN HEADERS (Parameters, DBReader, DBWriter, Debug, QueryMatcher, Matcher)
// and linked with the unmodified reference objects in place of three entry points:
//     int kmermatcher(int, const char **, const Command &)        lib/mmseqs/src/linclust/kmermatcher.cpp:786
//     int rescorediagonal(int, const char **, const Command &)    lib/mmseqs/src/alignment/rescorediagonal.cpp:381
//     int correction(int, const char **, const Command &)         src/assembler/correction.cpp:492
//     int cyclecheck(int, const char **, const Command &)         src/assembler/cyclecheck.cpp:31
// The reference's objects keep every other symbol (setLinearFilterDefault, setKmerLengthAndAlphabet, doRescorediagonal, doCorrection, ...);
// the build (oracle/build_ref.py --gpu-stages) renames the three original entry points to *_cpu with `objcopy --redefine-sym`, no reference
// source is touched.  Argument parsing, parameter printing, DB opening / closing, text formats and the error convention
// (Debug(ERROR) + EXIT) are the reference's; only the hot loops are replaced.  Calls whose parameters lie outside the path the GPU library
// implements (protein DBs, spaced / masked k-mers, rescore modes 1, 2, 4, compressed DBs, other matrices) go to the reference's own
// implementation (*_cpu); the nucleotide hot path has no CPU fallback: without a usable device the stage ends with an error.
#include <cstring>
#include <string>
#include <vector>

#include "DBReader.h"
#include "DBWriter.h"
#include "Debug.h"
#include "LocalParameters.h"
#include "MMseqsMPI.h"
#include "Matcher.h"
#include "Parameters.h"
#include "QueryMatcher.h"
#include "Util.h"
#include "kmermatcher.h"

#include "carpedeam_host.h"

int kmermatcher_cpu(int argc, const char **argv, const Command &command);       // the reference's own entry points, renamed at link time
int rescorediagonal_cpu(int argc, const char **argv, const Command &command);
int correction_cpu(int argc, const char **argv, const Command &command);
int cyclecheck_cpu(int argc, const char **argv, const Command &command);
void setCycleCheckDefaults(LocalParameters *p);   // src/assembler/cyclecheck.cpp:26 (stays in the reference's object)

namespace {

void check(cdg_ctx *ctx, int rc, const char *what) {
    if (rc != CDG_OK) {
        Debug(Debug::ERROR) << what << " failed: " << cdg_last_error(ctx) << "\n";
        EXIT(EXIT_FAILURE);
    }
}

cdg_ctx *gpu_context() {
    cdg_ctx *ctx = NULL;
    const char *dev = getenv("CDG_DEVICE");
    if (cdg_create(dev ? atoi(dev) : 0, &ctx) != CDG_OK) {
        Debug(Debug::ERROR) << "No usable CUDA device (the GPU stages have no CPU fallback)\n";
        EXIT(EXIT_FAILURE);
    }
    return ctx;
}

// a call outside the GPU path goes to the reference's own entry point, which parses argv again: forget what this one parsed
// (parseParameters refuses a parameter whose wasSet flag is already set, Parameters.cpp:1536-1540)
int delegate(int (*cpu)(int, const char **, const Command &), int argc, const char **argv, const Command &command) {
    for (size_t i = 0; i < command.params->size(); i++) (*command.params)[i]->wasSet = false;
    return cpu(argc, argv, command);
}
bool single_process() {
#ifdef HAVE_MPI
    return MMseqsMPI::numProc <= 1;
#else
    return true;
#endif
}

bool plain_nucleotide_db(const Parameters &par, DBReader<unsigned int> &dbr) {
    return Parameters::isEqualDbtype(dbr.getDbtype(), Parameters::DBTYPE_NUCLEOTIDES) && !dbr.isCompressed() && par.compressed == 0 &&
           par.scoringMatrixFile.nucleotides != NULL && strstr(par.scoringMatrixFile.nucleotides, "nucleotide.out") != NULL;
}

// the sequence DB as the flat arrays cdg_upload_seqs takes (entries `BASES\n\0`; a DB whose data lies in several files is gathered)
struct FlatSeqs {
    std::vector<uint8_t> data; std::vector<uint64_t> off; std::vector<uint32_t> len; std::vector<uint8_t> ext; std::vector<uint32_t> key;
};
void flatten(DBReader<unsigned int> &dbr, FlatSeqs &s) {
    const size_t n = dbr.getSize();
    s.off.resize(n); s.len.resize(n); s.ext.resize(n); s.key.resize(n);
    size_t total = 0;
    for (size_t id = 0; id < n; id++) { s.off[id] = total; total += dbr.getEntryLen(id); }
    s.data.resize(total);
#pragma omp parallel for schedule(static)
    for (size_t id = 0; id < n; id++) {
        memcpy(s.data.data() + s.off[id], dbr.getData(id, 0), dbr.getEntryLen(id));
        s.len[id] = (uint32_t) dbr.getSeqLen(id); s.ext[id] = dbr.getExtData(id) ? 1 : 0; s.key[id] = dbr.getDbKey(id);
    }
}
void upload(cdg_ctx *ctx, const FlatSeqs &s) {
    check(ctx, cdg_upload_seqs(ctx, s.data.data(), s.data.size(), s.off.data(), s.len.data(), s.ext.data(), s.len.size()), "cdg_upload_seqs");
}

}   // namespace

// ---- kmermatcher: replaces kmermatcherInner (kmermatcher.cpp:596-740) --------------------------------------------------------------
int kmermatcher(int argc, const char **argv, const Command &command) {
    MMseqsMPI::init(argc, argv);
    Parameters &par = Parameters::getInstance();
    setLinearFilterDefault(&par);
    par.parseParameters(argc, argv, command, true, 0, MMseqsParameter::COMMAND_CLUSTLINEAR);
    DBReader<unsigned int> seqDbr(par.db1.c_str(), par.db1Index.c_str(), par.threads, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    seqDbr.open(DBReader<unsigned int>::NOSORT);
    const int querySeqType = seqDbr.getDbtype();
    setKmerLengthAndAlphabet(par, seqDbr.getAminoAcidDBSize(), querySeqType);
    const bool onPath = plain_nucleotide_db(par, seqDbr) && par.spacedKmer == 0 && par.maskMode == 0 && par.maskLowerCaseMode == 0 && par.ignoreMultiKmer &&
                        par.adjustKmerLength == 0 && par.kmerSize >= 1 && par.kmerSize <= 31 && seqDbr.getMaxSeqLen() <= par.maxSeqLen && single_process();
    if (!onPath) {
        seqDbr.close();
        return delegate(kmermatcher_cpu, argc, argv, command);
    }
    par.printParameters(command.cmd, argc, argv, *command.params);
    Debug(Debug::INFO) << "Database size: " << seqDbr.getSize() << " type: " << seqDbr.getDbTypeName() << " (GPU k-mer matcher)\n";
    FlatSeqs s;
    flatten(seqDbr, s);
    cdg_ctx *ctx = gpu_context();
    upload(ctx, s);
    cdg_kmer_params p;
    p.k = par.kmerSize; p.kmers_per_seq = par.kmersPerSequence; p.kmers_per_seq_scale = par.kmersPerSequenceScale.nucleotides; p.hash_shift = (uint64_t) par.hashShift;
    p.include_only_extendable = par.includeOnlyExtendable ? 1 : 0; p.hash_lo = 0; p.hash_hi = 65535;
    p.split_passes = 0;                       // as many hash-range passes as the free device memory asks for (kmermatcher.cpp:615-627)
    p.cov_thr = par.covThr; p.cov_mode = par.covMode;
    cdg_hits h;
    check(ctx, cdg_kmermatch(ctx, &p, &h), "cdg_kmermatch");
    // writeKmerMatcherResult (:815-930) + the singleton fill (:710-730): one entry per sequence key
    const size_t n = seqDbr.getSize();
    std::vector<uint64_t> first(n + 1, 0);
    for (uint64_t i = 0; i < h.n_hits; i++) first[h.rep[i] + 1]++;
    for (size_t q = 0; q < n; q++) first[q + 1] += first[q];
    DBWriter dbw(par.db2.c_str(), par.db2Index.c_str(), 1, par.compressed, Parameters::DBTYPE_PREFILTER_REV_RES);
    dbw.open();
    std::string entry;
    char buffer[100];
    for (size_t q = 0; q < n; q++) {
        entry.clear();
        hit_t self; self.seqId = s.key[q]; self.prefScore = 0; self.diagonal = 0;
        entry.append(buffer, QueryMatcher::prefilterHitToBuffer(buffer, self));
        for (uint64_t i = first[q]; i < first[q + 1]; i++) {
            hit_t t; t.seqId = s.key[h.target[i]]; t.prefScore = h.score[i]; t.diagonal = (unsigned short) h.diag[i];
            entry.append(buffer, QueryMatcher::prefilterHitToBuffer(buffer, t));
        }
        // only the singleton fill carries the ext flag (:727); entries of representatives are written without (:860, :923)
        dbw.writeData(entry.c_str(), entry.size(), s.key[q], 0, first[q + 1] > first[q] ? false : (s.ext[q] != 0));
    }
    dbw.close(false, false);
    cdg_destroy(ctx);
    seqDbr.close();
    return EXIT_SUCCESS;
}

// ---- rescorediagonal: replaces the OpenMP loop of doRescorediagonal (rescorediagonal.cpp:125-360) -------------------------------------
int rescorediagonal(int argc, const char **argv, const Command &command) {
    MMseqsMPI::init(argc, argv);
    Parameters &par = Parameters::getInstance();
    par.parseParameters(argc, argv, command, true, 0, 0);
    if (par.wrappedScoring && par.rescoreMode != Parameters::RESCORE_MODE_HAMMING) {
        Debug(Debug::ERROR) << "ERROR: wrapped scoring is only allowed with RESCORE_MODE_HAMMING\n";
        return EXIT_FAILURE;
    }
    const bool hamming = par.rescoreMode == Parameters::RESCORE_MODE_HAMMING, ungapped = par.rescoreMode == Parameters::RESCORE_MODE_END_TO_END_ALIGNMENT;
    bool onPath = (hamming || ungapped) && par.db1.compare(par.db2) == 0 && !par.filterHits && !par.addBacktrace && par.sortResults == 0 && !par.includeIdentity &&
                  single_process();
    if (ungapped) onPath = onPath && par.covThr == 0.0f && par.alnLenThr == 0 && par.seqIdMode == Parameters::SEQ_ID_ALN_LEN;   // the assembler's setting
    DBReader<unsigned int> *seqDbr = NULL;
    if (onPath) {
        seqDbr = new DBReader<unsigned int>(par.db1.c_str(), par.db1Index.c_str(), par.threads, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
        seqDbr->open(DBReader<unsigned int>::NOSORT);
        onPath = plain_nucleotide_db(par, *seqDbr);
        if (!onPath) { seqDbr->close(); delete seqDbr; }
    }
    if (!onPath) return delegate(rescorediagonal_cpu, argc, argv, command);
    DBReader<unsigned int> resultReader(par.db3.c_str(), par.db3Index.c_str(), par.threads, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    resultReader.open(DBReader<unsigned int>::LINEAR_ACCCESS);
    const int dbtype = hamming ? resultReader.getDbtype() : Parameters::DBTYPE_ALIGNMENT_RES;
    FlatSeqs s;
    flatten(*seqDbr, s);
    cdg_ctx *ctx = gpu_context();
    upload(ctx, s);
    // prefilter lines of every query, target keys -> indices (QueryMatcher::parsePrefilterHits, QueryMatcher.h:81-104)
    const size_t n = seqDbr->getSize();
    std::vector<uint64_t> lineOff(n + 1, 0);
    std::vector<uint32_t> tgt; std::vector<int32_t> score; std::vector<uint16_t> diag;
    std::vector<char> hasEntry(n, 0);
    for (size_t q = 0; q < n; q++) {
        const size_t rid = resultReader.getId(s.key[q]);
        if (rid != UINT_MAX) {
            hasEntry[q] = 1;
            std::vector<hit_t> hits = QueryMatcher::parsePrefilterHits(resultReader.getData(rid, 0));
            for (size_t i = 0; i < hits.size(); i++) {
                const size_t tid = seqDbr->getId(hits[i].seqId);
                if (tid == UINT_MAX) { Debug(Debug::ERROR) << "Invalid database read for database data file=" << par.db2 << "\n"; EXIT(EXIT_FAILURE); }
                tgt.push_back((uint32_t) tid); score.push_back(hits[i].prefScore); diag.push_back(hits[i].diagonal);
            }
        }
        lineOff[q + 1] = tgt.size();
    }
    check(ctx, cdg_set_pref_lines(ctx, lineOff.data(), n, tgt.data(), score.data(), diag.data(), tgt.size()), "cdg_set_pref_lines");
    DBWriter resultWriter(par.db4.c_str(), par.db4Index.c_str(), 1, par.compressed, dbtype);
    resultWriter.open();
    std::string entry;
    char buffer[1024];
    if (hamming) {
        cdg_hamming_params hp; hp.cov_thr = par.covThr; hp.cov_mode = par.covMode; hp.min_seq_id = par.seqIdThr; hp.seq_id_mode = par.seqIdMode;
        hp.min_aln_len = par.alnLenThr; hp.wrapped_scoring = par.wrappedScoring ? 1 : 0;
        cdg_hamming_out r;
        check(ctx, cdg_rescore_hamming(ctx, &hp, &r), "cdg_rescore_hamming");
        for (size_t q = 0; q < n; q++) {
            if (!hasEntry[q]) continue;
            entry.clear();
            for (uint64_t l = r.line_off[q]; l < r.line_off[q + 1]; l++) {
                if (!r.keep[l]) continue;
                hit_t hit; hit.seqId = s.key[r.target[l]]; hit.prefScore = r.pref_score[l]; hit.diagonal = (unsigned short) r.diag[l];   // :323-329
                entry.append(buffer, QueryMatcher::prefilterHitToBuffer(buffer, hit));
            }
            resultWriter.writeData(entry.c_str(), entry.size(), s.key[q], 0);
        }
    } else {
        cdg_rescore_params rp; rp.min_seq_id = par.seqIdThr; rp.evalue_thr = par.evalThr; rp.db_residues = 0;   // same DB: getAminoAcidDBSize()
        cdg_alns r;
        check(ctx, cdg_rescore(ctx, &rp, &r), "cdg_rescore");
        for (size_t q = 0; q < n; q++) {
            if (!hasEntry[q]) continue;
            entry.clear();
            for (uint64_t l = r.line_off[q]; l < r.line_off[q + 1]; l++) {
                if (!r.keep[l]) continue;
                const uint32_t t = r.target[l];
                const int alnLen = std::max(abs(r.q_end[l] - r.q_start[l]), abs(r.db_end[l] - r.db_start[l])) + 1;
                Matcher::result_t res(s.key[t], r.bits[l], 0.0f, 0.0f, r.seq_id[l], r.evalue[l], alnLen, r.q_start[l], r.q_end[l], s.len[q], r.db_start[l], r.db_end[l],
                                      s.len[t], "");   // :278-279 (the coverage fields are not written)
                entry.append(buffer, Matcher::resultToBuffer(buffer, res, false, false));
            }
            resultWriter.writeData(entry.c_str(), entry.size(), s.key[q], 0);
        }
    }
    resultWriter.close();
    resultReader.close();
    cdg_destroy(ctx);
    seqDbr->close();
    delete seqDbr;
    return EXIT_SUCCESS;
}

// ---- ancient_correction: replaces the parallel region of doCorrection (correction.cpp:158-478) ----------------------------------------
int correction(int argc, const char **argv, const Command &command) {
    LocalParameters &par = LocalParameters::getLocalInstance();
    par.parseParameters(argc, argv, command, true, 0, 0);
    MMseqsMPI::init(argc, argv);
    par.allowDeletion = false;
    DBReader<unsigned int> sequenceDbr(par.db1.c_str(), par.db1Index.c_str(), par.threads, DBReader<unsigned int>::USE_DATA | DBReader<unsigned int>::USE_INDEX);
    sequenceDbr.open(DBReader<unsigned int>::NOSORT);
    if (!plain_nucleotide_db(par, sequenceDbr) || !single_process()) {
        sequenceDbr.close();
        return delegate(correction_cpu, argc, argv, command);
    }
    Debug(Debug::INFO) << "Do correction (GPU).\n";
    DBReader<unsigned int> alnReader(par.db2.c_str(), par.db2Index.c_str(), par.threads, DBReader<unsigned int>::USE_DATA | DBReader<unsigned int>::USE_INDEX);
    alnReader.open(DBReader<unsigned int>::NOSORT);
    FlatSeqs s;
    flatten(sequenceDbr, s);
    cdg_ctx *ctx = gpu_context();
    upload(ctx, s);
    cdg_damage dmg; char err[512];
    if (cdh_build_damage(par.ancientDamagePath.c_str(), 0.01, &dmg, err, sizeof err) != 0) {   // correction.cpp:153-156, :184-197
        Debug(Debug::ERROR) << err << "\n";
        EXIT(EXIT_FAILURE);
    }
    check(ctx, cdg_set_damage(ctx, &dmg), "cdg_set_damage");
    const size_t n = sequenceDbr.getSize();
    std::vector<uint64_t> off(n + 1, 0);
    std::vector<uint32_t> tgt; std::vector<float> sid; std::vector<int32_t> qs, qe, ds, de;
    std::vector<Matcher::result_t> alignments;
    for (size_t q = 0; q < n; q++) {
        alignments.clear();
        char *alnData = alnReader.getDataByDBKey(s.key[q], 0);
        if (alnData != NULL) Matcher::readAlignmentResults(alignments, alnData);   // :205-207
        for (size_t i = 0; i < alignments.size(); i++) {
            const size_t tid = sequenceDbr.getId(alignments[i].dbKey);
            if (tid == UINT_MAX) { Debug(Debug::ERROR) << "Invalid database read for database data file=" << par.db1 << "\n"; EXIT(EXIT_FAILURE); }
            tgt.push_back((uint32_t) tid); sid.push_back(alignments[i].seqId);
            qs.push_back(alignments[i].qStartPos); qe.push_back(alignments[i].qEndPos); ds.push_back(alignments[i].dbStartPos); de.push_back(alignments[i].dbEndPos);
        }
        off[q + 1] = tgt.size();
    }
    check(ctx, cdg_set_alignments(ctx, off.data(), n, tgt.data(), sid.data(), qs.data(), qe.data(), ds.data(), de.data(), tgt.size()), "cdg_set_alignments");
    cdg_corr_params cp; cp.min_seqid_corr_reads = par.seqIdThr; cp.min_ryseq_id_corr_reads = par.corrReadsRySeqId;   // :297, :359
    std::vector<uint8_t> out(s.data.size());
    uint64_t changed = 0;
    check(ctx, cdg_correct(ctx, &cp, out.data(), &changed), "cdg_correct");
    DBWriter resultWriter(par.db3.c_str(), par.db3Index.c_str(), 1, par.compressed, sequenceDbr.getDbtype());
    resultWriter.open();
    for (size_t q = 0; q < n; q++)   // :465-470: corrected bases + '\n', same key and ext flag
        resultWriter.writeData((const char *) out.data() + s.off[q], s.len[q] + 1, s.key[q], 0, s.ext[q] != 0);
    resultWriter.close(true);
    alnReader.close();
    cdg_destroy(ctx);
    sequenceDbr.close();
    Debug(Debug::INFO) << "\nDone. " << changed << " bases corrected.\n";
    return EXIT_SUCCESS;
}

// ---- cyclecheck: replaces the OpenMP region of cyclecheck (cyclecheck.cpp:72-262) ---------------------------------------------------
int cyclecheck(int argc, const char **argv, const Command &command) {
    LocalParameters &par = LocalParameters::getLocalInstance();
    setCycleCheckDefaults(&par);
    par.parseParameters(argc, argv, command, true, 0, 0);
    DBReader<unsigned int> seqDbr(par.db1.c_str(), par.db1Index.c_str(), par.threads, DBReader<unsigned int>::USE_DATA | DBReader<unsigned int>::USE_INDEX);
    seqDbr.open(DBReader<unsigned int>::NOSORT);
    if (!plain_nucleotide_db(par, seqDbr) || par.kmerSize != 22) {
        seqDbr.close();
        return delegate(cyclecheck_cpu, argc, argv, command);
    }
    DBWriter cycleResultWriter(par.db2.c_str(), par.db2Index.c_str(), 1, par.compressed, Parameters::DBTYPE_NUCLEOTIDES);
    cycleResultWriter.open();
    FlatSeqs s;
    flatten(seqDbr, s);
    cdg_ctx *ctx = gpu_context();
    upload(ctx, s);
    cdg_cycle_params cp; cp.k = par.kmerSize; cp.max_seq_len = (uint32_t) par.maxSeqLen;
    cdg_cycle_out co;
    check(ctx, cdg_cyclecheck(ctx, &cp, &co), "cdg_cyclecheck");
    for (size_t id = 0; id < seqDbr.getSize(); id++) {
        if (s.len[id] >= par.maxSeqLen)
            Debug(Debug::WARNING) << "Sequence " << seqDbr.getDbKey(id) << " too long. It will be skipped. Max length allowed is " << par.maxSeqLen << "\n";
        if (co.split_diagonal[id] == 0) continue;
        char *nuclSeq = seqDbr.getData(id, 0);                       // :238-252
        unsigned int len = seqDbr.getEntryLen(id) - 1;
        std::string seq;
        if (par.chopCycle) {
            seq = std::string(nuclSeq, co.split_diagonal[id]);
            seq.push_back('\n');
            nuclSeq = (char *) seq.c_str();
            len = seq.size();
        }
        cycleResultWriter.writeData(nuclSeq, len, seqDbr.getDbKey(id), 0);
    }
    cycleResultWriter.close(true);
    cdg_destroy(ctx);
    seqDbr.close();
    return EXIT_SUCCESS;
}

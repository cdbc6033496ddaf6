// carpedeam_gpu — drop-in stage binary: the three hot-path modules of the reference with the same positional
// database arguments and the flags the assembler workflow passes (data/nuclassemble.sh:105-126):
//
//   carpedeam_gpu kmermatcher        <seqDB> <prefDB>                 [flags]   (kmermatcher.cpp:786)
//   carpedeam_gpu rescorediagonal    <qDB> <tDB> <prefDB> <alnDB>     [flags]   (rescorediagonal.cpp:381)
//   carpedeam_gpu ancient_correction <seqDB> <alnDB> <outDB>          [flags]   (correction.cpp:492)
// and, either side of the loop:
//   carpedeam_gpu rescorediagonal    ... --rescore-mode 0 [--wrapped-scoring 1]  (linclust's Hamming pre-clustering, linclust.sh:28-32)
//   carpedeam_gpu cyclecheck         <seqDB> <cycleDB>                [flags]   (cyclecheck.cpp:31)
//   carpedeam_gpu createdb           <fasta/fastq[.gz]> ... <seqDB>   [flags]   (createdb.cpp; host only)
//   carpedeam_gpu iteration          <seqDB> <prefDB> <alnDB> <corrDB> [flags]  (the three hot-path stages fused: one process, one upload, three DBs)
//
// Host side only: DBReader/DBWriter-compatible file access and the text formats; all arithmetic happens in
// libcarpedeam_gpu.so through the C ABI.  Errors print a message and exit(EXIT_FAILURE), as the reference does.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/carpedeam_host.h"
#include "formats.h"
#include "mmdb.h"

using namespace cdh;

namespace {

struct Args {
    std::vector<std::string> pos;
    std::map<std::string, std::string> opt;
    std::string get(const std::string &k, const std::string &def) const { auto it = opt.find(k); return it == opt.end() ? def : it->second; }
    double num(const std::string &k, double def) const { auto it = opt.find(k); return it == opt.end() ? def : atof(it->second.c_str()); }
};
Args parse_args(int argc, char **argv, int first) {
    Args a;
    for (int i = first; i < argc; i++) {
        std::string s = argv[i];
        if (s.size() > 1 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9')) {   // every MMseqs2 flag takes one value
            if (i + 1 < argc) { a.opt[s] = argv[i + 1]; i++; } else a.opt[s] = "";
        } else a.pos.push_back(s);
    }
    return a;
}

void ck(cdg_ctx *ctx, int rc, const char *what) {
    if (rc != 0) die(std::string(what) + " failed: " + cdg_last_error(ctx));
}

// The implementation is pinned to the values the assembler workflow passes (tests/golden/cfg1_stage_argv.log): a flag given with
// any other value would silently change the reference's output, so it is an error here.  Flags that cannot affect the
// output (--threads, -v, --db-load-mode, --split-memory-limit, ...) are accepted as they are.
struct Pinned { const char *flag; const char *value; };
void require_pinned(const Args &a, const char *stage, std::initializer_list<Pinned> table) {
    for (const Pinned &p : table) {
        auto it = a.opt.find(p.flag);
        if (it == a.opt.end()) continue;
        const bool numeric = p.value[0] == '-' || (p.value[0] >= '0' && p.value[0] <= '9');
        const bool same = numeric ? atof(it->second.c_str()) == atof(p.value) : it->second.find(p.value) != std::string::npos;
        if (!same) die(std::string(stage) + ": " + p.flag + " " + it->second + " is not supported by carpedeam_gpu (implemented: " + p.flag + " " + p.value + ")");
    }
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int host_threads(const Args &a) {
    int t = (int) a.num("--threads", 0);
#ifdef _OPENMP
    if (t <= 0) t = omp_get_max_threads();
#else
    t = 1;
#endif
    return t < 1 ? 1 : t;
}

// Entries of queries [0, n) built by `threads` host threads, each over a contiguous block of queries with its own writer (the
// reference's per-thread DBWriter, rescorediagonal.cpp:125-146,340-352), merged in query order into one data file.
template <class F>
void write_entries_parallel(const std::string &path, int dbtype, size_t n, int threads, F entryOf) {
    if (threads > (int) (n / 1024 + 1)) threads = (int) (n / 1024 + 1);
    std::vector<DbWriter> parts;
    for (int t = 0; t < threads; t++) parts.emplace_back(path, dbtype);
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; t++) {
        const size_t lo = n * (size_t) t / (size_t) threads, hi = n * (size_t) (t + 1) / (size_t) threads;
        std::string entry;
        for (size_t q = lo; q < hi; q++) {
            entry.clear();
            uint32_t key = 0; uint8_t ext = 0;
            entryOf(q, entry, key, ext);
            parts[t].write(entry.data(), entry.size(), key, ext);
        }
    }
    parts[0].close_parts(threads > 1 ? &parts[1] : nullptr, (size_t) (threads - 1));
}

// per-query variable-length records parsed by `threads` host threads into one CSR: parseOf(q, push) calls push(record)
template <class R, class F>
void parse_entries_parallel(size_t n, int threads, std::vector<uint64_t> &off, std::vector<R> &records, F parseOf) {
    if (threads > (int) (n / 1024 + 1)) threads = (int) (n / 1024 + 1);
    std::vector<std::vector<R>> part(threads);
    off.assign(n + 1, 0);
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; t++) {
        const size_t lo = n * (size_t) t / (size_t) threads, hi = n * (size_t) (t + 1) / (size_t) threads;
        for (size_t q = lo; q < hi; q++) {
            const size_t before = part[t].size();
            parseOf(q, part[t]);
            off[q + 1] = part[t].size() - before;   // count, turned into an offset below
        }
    }
    for (size_t q = 0; q < n; q++) off[q + 1] += off[q];
    records.resize(off[n]);
    size_t pos = 0;
    for (int t = 0; t < threads; t++) { std::copy(part[t].begin(), part[t].end(), records.begin() + pos); pos += part[t].size(); }
}

cdg_ctx *make_ctx() {
    cdg_ctx *ctx = nullptr;
    const char *dev = getenv("CDG_DEVICE");
    if (cdg_create(dev ? atoi(dev) : 0, &ctx) != 0) die("No usable CUDA device (carpedeam_gpu has no CPU fallback)");
    return ctx;
}

void upload(cdg_ctx *ctx, const Db &db, std::vector<uint32_t> &lens) {
    if (db.dbtype != DBTYPE_NUCLEOTIDES && db.dbtype != 0) die("Only nucleotide sequence databases are supported");
    lens.resize(db.size());
    for (size_t i = 0; i < db.size(); i++) lens[i] = (uint32_t) (db.lengths[i] >= 2 ? db.lengths[i] - 2 : 0);   // DBReader::getSeqLen
    ck(ctx, cdg_upload_seqs(ctx, db.data.data(), db.data.size(), db.offsets.data(), lens.data(), db.ext.data(), db.size()), "cdg_upload_seqs");
}

int kmermatcher(int argc, char **argv) {
    Args a = parse_args(argc, argv, 2);
    if (a.pos.size() < 2) die("Usage: kmermatcher <i:sequenceDB> <o:prefDB> [options]");
    require_pinned(a, "kmermatcher", {{"--sub-mat", "nucleotide.out"}, {"--alph-size", "nucl:5"}, {"--spaced-kmer-mode", "0"}, {"--adjust-kmer-len", "0"},
                                      {"--mask", "0"}, {"--mask-lower-case", "0"}, {"--ignore-multi-kmer", "1"}, {"--compressed", "0"}});
    const double t0 = now_ms();
    Db seq; open_db(a.pos[0], seq);
    if (a.opt.count("--max-seq-len"))
        for (size_t i = 0; i < seq.size(); i++)
            if (seq.lengths[i] >= 2 && (double) (seq.lengths[i] - 2) > a.num("--max-seq-len", 65535)) die("kmermatcher: a sequence is longer than --max-seq-len (truncation is not implemented)");
    const double t1 = now_ms();
    cdg_ctx *ctx = make_ctx();
    std::vector<uint32_t> lens;
    upload(ctx, seq, lens);
    cdg_kmer_params p;
    p.k = (int) a.num("-k", 0);
    if (p.k == 0) die("kmermatcher: -k is required for nucleotide databases in this build");
    p.kmers_per_seq = (int) a.num("--kmer-per-seq", 21);          // Parameters.cpp default; the assembler passes 200
    p.kmers_per_seq_scale = (float) a.num("--kmer-per-seq-scale", 0.2);
    p.hash_shift = (uint64_t) a.num("--hash-shift", 67);
    p.include_only_extendable = (int) a.num("--include-only-extendable", 0);
    p.hash_lo = 0; p.hash_hi = 65535;
    p.split_passes = 0;   // --split-memory-limit 0: as many hash-range passes as the free device memory asks for (usually one)
    p.cov_thr = (float) a.num("-c", 0.0); p.cov_mode = (int) a.num("--cov-mode", 0);   // assignGroup's canBeCovered (kmermatcher.cpp:526-529)
    if (p.cov_mode < 0 || p.cov_mode > 5) die("kmermatcher: unknown --cov-mode");
    cdg_hits h;
    ck(ctx, cdg_kmermatch(ctx, &p, &h), "cdg_kmermatch");
    const double t2 = now_ms();
    // first hit of every representative (the hit list is sorted by representative)
    std::vector<uint64_t> first(seq.size() + 1, 0);
    for (uint64_t i = 0; i < h.n_hits; i++) first[h.rep[i] + 1]++;
    for (size_t q = 0; q < seq.size(); q++) first[q + 1] += first[q];
    write_entries_parallel(a.pos[1], DBTYPE_PREFILTER_REV_RES, seq.size(), host_threads(a), [&](size_t q, std::string &entry, uint32_t &key, uint8_t &ext) {
        char buf[64];
        entry.append(buf, format_pref_line(buf, seq.keys[q], 0, 0));   // `rep 0 0` / singleton `key 0 0` (kmermatcher.cpp:836, :710-730)
        for (uint64_t i = first[q]; i < first[q + 1]; i++) entry.append(buf, format_pref_line(buf, seq.keys[h.target[i]], h.score[i], h.diag[i]));
        // only the singleton fill carries the ext flag (kmermatcher.cpp:727); entries of representatives are written without (:860,:923)
        key = seq.keys[q]; ext = first[q + 1] > first[q] ? 0 : seq.ext[q];
    });
    fprintf(stderr, "carpedeam_gpu kmermatcher: read DB %.1f ms, upload + GPU + export %.1f ms, format + write DB %.1f ms\n", t1 - t0, t2 - t1, now_ms() - t2);
    cdg_destroy(ctx);
    return EXIT_SUCCESS;
}

// prefilter lines of every query of `seq` from a pref DB, target keys replaced by indices
void load_pref_lines(const Args &a, const Db &seq, const Db &pref, std::vector<uint64_t> &lineOff, std::vector<uint32_t> &tgt, std::vector<int32_t> &score,
                     std::vector<uint16_t> &diag) {
    std::vector<PrefLine> all;
    parse_entries_parallel(seq.size(), host_threads(a), lineOff, all, [&](size_t q, std::vector<PrefLine> &out) {
        long pi = pref.find(seq.keys[q]);
        const size_t before = out.size();
        if (pi >= 0) parse_pref_entry((const char *) pref.data.data() + pref.offsets[pi], out);
        for (size_t i = before; i < out.size(); i++) {
            long ti = seq.find(out[i].targetKey);
            if (ti < 0) die("Invalid database read for database data file=" + a.pos[1]);
            out[i].targetKey = (uint32_t) ti;
        }
    });
    tgt.resize(all.size()); score.resize(all.size()); diag.resize(all.size());
    for (size_t i = 0; i < all.size(); i++) { tgt[i] = all[i].targetKey; score[i] = all[i].score; diag[i] = all[i].diag; }
}

// rescorediagonal --rescore-mode 0 (HAMMING_PAR of linclust, lib/mmseqs/src/workflow/Linclust.cpp:103-112): prefilter-format output,
// one entry per query of the pref DB (rescorediagonal.cpp:323-352)
int rescore_hamming(const Args &a) {
    require_pinned(a, "rescorediagonal", {{"--sub-mat", "nucleotide.out"}, {"--filter-hits", "0"}, {"-a", "0"}, {"--add-self-matches", "0"}, {"--sort-results", "0"},
                                          {"--compressed", "0"}});
    Db seq; open_db(a.pos[0], seq);
    Db pref; open_db(a.pos[2], pref);
    cdg_ctx *ctx = make_ctx();
    std::vector<uint32_t> lens;
    upload(ctx, seq, lens);
    std::vector<uint64_t> lineOff; std::vector<uint32_t> tgt; std::vector<int32_t> score; std::vector<uint16_t> diag;
    load_pref_lines(a, seq, pref, lineOff, tgt, score, diag);
    ck(ctx, cdg_set_pref_lines(ctx, lineOff.data(), seq.size(), tgt.data(), score.data(), diag.data(), tgt.size()), "cdg_set_pref_lines");
    cdg_hamming_params p;
    p.cov_thr = (float) a.num("-c", 0.0); p.cov_mode = (int) a.num("--cov-mode", 0); p.min_seq_id = (float) a.num("--min-seq-id", 0.0);
    p.seq_id_mode = (int) a.num("--seq-id-mode", 0); p.min_aln_len = (int) a.num("--min-aln-len", 0); p.wrapped_scoring = (int) a.num("--wrapped-scoring", 0);
    cdg_hamming_out r;
    ck(ctx, cdg_rescore_hamming(ctx, &p, &r), "cdg_rescore_hamming");
    // entries only for the keys of the pref DB (the reference iterates the result DB, rescorediagonal.cpp:143-147)
    std::vector<size_t> withEntry;
    for (size_t q = 0; q < seq.size(); q++) if (pref.find(seq.keys[q]) >= 0) withEntry.push_back(q);
    write_entries_parallel(a.pos[3], pref.dbtype ? pref.dbtype : DBTYPE_PREFILTER_REV_RES, withEntry.size(), host_threads(a),
                           [&](size_t i, std::string &entry, uint32_t &key, uint8_t &ext) {
        const size_t q = withEntry[i];
        char buf[64];
        for (uint64_t l = r.line_off[q]; l < r.line_off[q + 1]; l++)
            if (r.keep[l]) entry.append(buf, format_pref_line(buf, seq.keys[r.target[l]], r.pref_score[l], (int16_t) r.diag[l]));
        key = seq.keys[q]; ext = 0;
    });
    fprintf(stderr, "carpedeam_gpu rescorediagonal --rescore-mode 0: %llu of %llu lines kept\n", (unsigned long long) r.n_kept, (unsigned long long) r.n_lines);
    cdg_destroy(ctx);
    return EXIT_SUCCESS;
}

// cyclecheck <i:sequenceDB> <o:cycleDB> (src/assembler/cyclecheck.cpp:31-269): only the reported sequences are written, whole or
// (--chop-cycle 1) cut at the split diagonal
int cyclecheck(int argc, char **argv) {
    Args a = parse_args(argc, argv, 2);
    if (a.pos.size() < 2) die("Usage: cyclecheck <i:sequenceDB> <o:cycleDB> [options]");
    require_pinned(a, "cyclecheck", {{"--compressed", "0"}});
    Db seq; open_db(a.pos[0], seq);
    cdg_ctx *ctx = make_ctx();
    std::vector<uint32_t> lens;
    upload(ctx, seq, lens);
    cdg_cycle_params p; p.k = 22; p.max_seq_len = (uint32_t) a.num("--max-seq-len", 65535);
    cdg_cycle_out r;
    ck(ctx, cdg_cyclecheck(ctx, &p, &r), "cdg_cyclecheck");
    const bool chop = (int) a.num("--chop-cycle", 0) != 0;
    std::vector<size_t> cyc;
    for (size_t q = 0; q < seq.size(); q++) {
        if (lens[q] >= p.max_seq_len) fprintf(stderr, "Sequence %u too long. It will be skipped. Max length allowed is %u\n", seq.keys[q], p.max_seq_len);
        if (r.split_diagonal[q]) cyc.push_back(q);
    }
    write_entries_parallel(a.pos[1], DBTYPE_NUCLEOTIDES, cyc.size(), host_threads(a), [&](size_t i, std::string &entry, uint32_t &key, uint8_t &ext) {
        const size_t q = cyc[i];
        entry.assign((const char *) seq.data.data() + seq.offsets[q], chop ? r.split_diagonal[q] : lens[q]);
        entry.push_back('\n');
        key = seq.keys[q]; ext = 0;
    });
    fprintf(stderr, "carpedeam_gpu cyclecheck: %llu of %zu sequences reported\n", (unsigned long long) r.n_cyclic, seq.size());
    cdg_destroy(ctx);
    return EXIT_SUCCESS;
}

// createdb <i:fastaFile1[.gz]> ... <o:sequenceDB> (lib/mmseqs/src/util/createdb.cpp): host only
int createdb(int argc, char **argv) {
    Args a = parse_args(argc, argv, 2);
    if (a.pos.size() < 2) die("Usage: createdb <i:fastaFile1[.gz]> ... <i:fastaFileN[.gz]> <o:sequenceDB> [options]");
    require_pinned(a, "createdb", {{"--compressed", "0"}, {"--createdb-mode", "1"}, {"--dbtype", "0"}, {"--id-offset", "0"}});
    std::vector<const char *> files;
    for (size_t i = 0; i + 1 < a.pos.size(); i++) files.push_back(a.pos[i].c_str());
    cdh_seqset s; char err[512];
    if (cdh_createdb(files.data(), (int) files.size(), (int) a.num("--shuffle", 1), (int) a.num("--write-lookup", 1), a.pos.back().c_str(), &s, err, sizeof err) != 0) die(err);
    fprintf(stderr, "carpedeam_gpu createdb: %llu sequences, %llu residues\n", (unsigned long long) s.n, (unsigned long long) s.total_residues);
    cdh_free_seqset(&s);
    return EXIT_SUCCESS;
}

int rescorediagonal(int argc, char **argv) {
    Args a = parse_args(argc, argv, 2);
    if (a.pos.size() < 4) die("Usage: rescorediagonal <i:queryDB> <i:targetDB> <i:prefDB> <o:alnDB> [options]");
    if (a.pos[0] != a.pos[1]) die("rescorediagonal: query and target database must be the same (assembler usage)");
    const int mode = (int) a.num("--rescore-mode", 0);
    if (mode == 0) return rescore_hamming(a);   // linclust's pre-clustering step
    if (mode != 3) die("rescorediagonal: --rescore-mode 3 (assembler) and 0 (linclust pre-clustering) are implemented");
    require_pinned(a, "rescorediagonal", {{"--sub-mat", "nucleotide.out"}, {"--wrapped-scoring", "0"}, {"--filter-hits", "0"}, {"-a", "0"}, {"-c", "0"},
                                          {"--min-aln-len", "0"}, {"--seq-id-mode", "0"}, {"--add-self-matches", "0"}, {"--sort-results", "0"},
                                          {"--compressed", "0"}});
    const double t0 = now_ms();
    Db seq; open_db(a.pos[0], seq);
    Db pref; open_db(a.pos[2], pref);
    const double t1 = now_ms();
    cdg_ctx *ctx = make_ctx();
    std::vector<uint32_t> lens;
    upload(ctx, seq, lens);
    std::vector<uint64_t> lineOff;
    std::vector<PrefLine> all;   // targetKey replaced by the target's index
    parse_entries_parallel(seq.size(), host_threads(a), lineOff, all, [&](size_t q, std::vector<PrefLine> &out) {
        long pi = pref.find(seq.keys[q]);
        const size_t before = out.size();
        if (pi >= 0) parse_pref_entry((const char *) pref.data.data() + pref.offsets[pi], out);
        for (size_t i = before; i < out.size(); i++) {
            long ti = seq.find(out[i].targetKey);
            if (ti < 0) die("Invalid database read for database data file=" + a.pos[1]);
            out[i].targetKey = (uint32_t) ti;
        }
    });
    std::vector<uint32_t> tgt(all.size()); std::vector<int32_t> score(all.size()); std::vector<uint16_t> diag(all.size());
    for (size_t i = 0; i < all.size(); i++) { tgt[i] = all[i].targetKey; score[i] = all[i].score; diag[i] = all[i].diag; }
    const double t2 = now_ms();
    ck(ctx, cdg_set_pref_lines(ctx, lineOff.data(), seq.size(), tgt.data(), score.data(), diag.data(), tgt.size()), "cdg_set_pref_lines");
    cdg_rescore_params p;
    p.min_seq_id = (float) a.num("--min-seq-id", 0.0);
    p.evalue_thr = a.num("-e", 0.001);
    p.db_residues = 0;   // same DB: getAminoAcidDBSize() = sum of sequence lengths
    cdg_alns r;
    ck(ctx, cdg_rescore(ctx, &p, &r), "cdg_rescore");
    const double t3 = now_ms();
    write_entries_parallel(a.pos[3], DBTYPE_ALIGNMENT_RES, seq.size(), host_threads(a), [&](size_t q, std::string &entry, uint32_t &key, uint8_t &ext) {
        char buf[256];
        for (uint64_t l = r.line_off[q]; l < r.line_off[q + 1]; l++) {
            if (!r.keep[l]) continue;
            uint32_t t = r.target[l];
            entry.append(buf, format_aln_line(buf, seq.keys[t], r.bits[l], r.seq_id[l], r.evalue[l], r.q_start[l], r.q_end[l], (int32_t) lens[q],
                                              r.db_start[l], r.db_end[l], (int32_t) lens[t]));
        }
        key = seq.keys[q]; ext = 0;
    });
    fprintf(stderr, "carpedeam_gpu rescorediagonal: read DBs %.1f ms, parse pref text %.1f ms, upload + GPU + export %.1f ms, format + write DB %.1f ms\n",
            t1 - t0, t2 - t1, t3 - t2, now_ms() - t3);
    cdg_destroy(ctx);
    return EXIT_SUCCESS;
}

int ancient_correction(int argc, char **argv) {
    Args a = parse_args(argc, argv, 2);
    if (a.pos.size() < 3) die("Usage: ancient_correction <i:sequenceDB> <i:alnDB> <o:correctedDB> [options]");
    require_pinned(a, "ancient_correction", {{"--compressed", "0"}});
    const double t0 = now_ms();
    Db seq; open_db(a.pos[0], seq);
    Db aln; open_db(a.pos[1], aln);
    const double t1 = now_ms();
    cdg_ctx *ctx = make_ctx();
    std::vector<uint32_t> lens;
    upload(ctx, seq, lens);
    cdg_damage dmg; char err[512];
    if (cdh_build_damage(a.get("--ancient-damage", "").c_str(), 0.01, &dmg, err, sizeof err) != 0) die(err);   // correction.cpp:153-156,196
    ck(ctx, cdg_set_damage(ctx, &dmg), "cdg_set_damage");
    std::vector<uint64_t> off;
    std::vector<AlnLine> all;   // targetKey replaced by the target's index
    parse_entries_parallel(seq.size(), host_threads(a), off, all, [&](size_t q, std::vector<AlnLine> &out) {
        long ai = aln.find(seq.keys[q]);
        const size_t before = out.size();
        if (ai >= 0) parse_aln_entry((const char *) aln.data.data() + aln.offsets[ai], out);
        for (size_t i = before; i < out.size(); i++) {
            long ti = seq.find(out[i].targetKey);
            if (ti < 0) die("Invalid database read for database data file=" + a.pos[0]);
            out[i].targetKey = (uint32_t) ti;
        }
    });
    const size_t A = all.size();
    std::vector<uint32_t> tgt(A); std::vector<float> sid(A); std::vector<int32_t> qs(A), qe(A), ds(A), de(A);
    for (size_t i = 0; i < A; i++) { tgt[i] = all[i].targetKey; sid[i] = all[i].seqId; qs[i] = all[i].qStart; qe[i] = all[i].qEnd; ds[i] = all[i].dbStart; de[i] = all[i].dbEnd; }
    const double t2 = now_ms();
    ck(ctx, cdg_set_alignments(ctx, off.data(), seq.size(), tgt.data(), sid.data(), qs.data(), qe.data(), ds.data(), de.data(), tgt.size()),
       "cdg_set_alignments");
    cdg_corr_params p;
    p.min_seqid_corr_reads = (float) a.num("--min-seq-id", 0.0);
    p.min_ryseq_id_corr_reads = (float) a.num("--min-ryseq-id-corr-reads", 0.99);
    std::vector<uint8_t> out(seq.data.size());
    uint64_t changed = 0;
    ck(ctx, cdg_correct(ctx, &p, out.data(), &changed), "cdg_correct");
    const double t3 = now_ms();
    write_entries_parallel(a.pos[2], seq.dbtype ? seq.dbtype : DBTYPE_NUCLEOTIDES, seq.size(), host_threads(a), [&](size_t q, std::string &entry, uint32_t &key, uint8_t &ext) {
        entry.assign((const char *) out.data() + seq.offsets[q], lens[q]);
        entry.push_back('\n');                                           // correction.cpp:465-470
        key = seq.keys[q]; ext = seq.ext[q];
    });
    fprintf(stderr, "carpedeam_gpu ancient_correction: %llu bases corrected; read DBs %.1f ms, parse aln text %.1f ms, upload + GPU + export %.1f ms, write DB %.1f ms\n",
            (unsigned long long) changed, t1 - t0, t2 - t1, t3 - t2, now_ms() - t3);
    cdg_destroy(ctx);
    return EXIT_SUCCESS;
}

// iteration <i:sequenceDB> <o:prefDB> <o:alnDB> <o:correctedDB>: the three hot-path stages of one assembler iteration in ONE process and
// one upload (cdg_iteration: no pref / aln text round trip between the stages on the device side), emitting the same three databases the
// separate stages write (binary in-process hand-off with DB emission for drop-in compatibility).  Flags: the union of the
// three stages' (kmermatcher's -k / --kmer-per-seq / --include-only-extendable ..., rescorediagonal's --min-seq-id / -e, the correction's
// --ancient-damage / --min-ryseq-id-corr-reads).
int iteration(int argc, char **argv) {
    Args a = parse_args(argc, argv, 2);
    if (a.pos.size() < 4) die("Usage: iteration <i:sequenceDB> <o:prefDB> <o:alnDB> <o:correctedDB> [options]");
    require_pinned(a, "iteration", {{"--sub-mat", "nucleotide.out"}, {"--alph-size", "nucl:5"}, {"--spaced-kmer-mode", "0"}, {"--adjust-kmer-len", "0"}, {"--mask", "0"},
                                    {"--mask-lower-case", "0"}, {"--ignore-multi-kmer", "1"}, {"--compressed", "0"}, {"-c", "0"}, {"--rescore-mode", "3"},
                                    {"--wrapped-scoring", "0"}, {"--filter-hits", "0"}, {"-a", "0"}, {"--min-aln-len", "0"}, {"--seq-id-mode", "0"},
                                    {"--add-self-matches", "0"}, {"--sort-results", "0"}});
    const double t0 = now_ms();
    Db seq; open_db(a.pos[0], seq);
    cdg_ctx *ctx = make_ctx();
    std::vector<uint32_t> lens;
    upload(ctx, seq, lens);
    cdg_damage dmg; char err[512];
    if (cdh_build_damage(a.get("--ancient-damage", "").c_str(), 0.01, &dmg, err, sizeof err) != 0) die(err);
    ck(ctx, cdg_set_damage(ctx, &dmg), "cdg_set_damage");
    cdg_iter_params ip;
    ip.kmer.k = (int) a.num("-k", 0);
    if (ip.kmer.k == 0) die("iteration: -k is required");
    ip.kmer.kmers_per_seq = (int) a.num("--kmer-per-seq", 21); ip.kmer.kmers_per_seq_scale = (float) a.num("--kmer-per-seq-scale", 0.2);
    ip.kmer.hash_shift = (uint64_t) a.num("--hash-shift", 67); ip.kmer.include_only_extendable = (int) a.num("--include-only-extendable", 0);
    ip.kmer.hash_lo = 0; ip.kmer.hash_hi = 65535; ip.kmer.split_passes = 0; ip.kmer.cov_thr = 0.f; ip.kmer.cov_mode = (int) a.num("--cov-mode", 0);
    ip.rescore.min_seq_id = (float) a.num("--min-seq-id", 0.0); ip.rescore.evalue_thr = a.num("-e", 0.001); ip.rescore.db_residues = 0;
    ip.corr.min_seqid_corr_reads = (float) a.num("--min-seq-id", 0.0); ip.corr.min_ryseq_id_corr_reads = (float) a.num("--min-ryseq-id-corr-reads", 0.99);
    cdg_hits h; cdg_alns r;
    std::vector<uint8_t> out(seq.data.size());
    uint64_t changed = 0;
    const double t1 = now_ms();
    ck(ctx, cdg_iteration(ctx, &ip, &h, &r, out.data(), &changed), "cdg_iteration");
    const double t2 = now_ms();
    const int threads = host_threads(a);
    std::vector<uint64_t> first(seq.size() + 1, 0);
    for (uint64_t i = 0; i < h.n_hits; i++) first[h.rep[i] + 1]++;
    for (size_t q = 0; q < seq.size(); q++) first[q + 1] += first[q];
    write_entries_parallel(a.pos[1], DBTYPE_PREFILTER_REV_RES, seq.size(), threads, [&](size_t q, std::string &entry, uint32_t &key, uint8_t &ext) {
        char buf[64];
        entry.append(buf, format_pref_line(buf, seq.keys[q], 0, 0));
        for (uint64_t i = first[q]; i < first[q + 1]; i++) entry.append(buf, format_pref_line(buf, seq.keys[h.target[i]], h.score[i], h.diag[i]));
        key = seq.keys[q]; ext = first[q + 1] > first[q] ? 0 : seq.ext[q];
    });
    write_entries_parallel(a.pos[2], DBTYPE_ALIGNMENT_RES, seq.size(), threads, [&](size_t q, std::string &entry, uint32_t &key, uint8_t &ext) {
        char buf[256];
        for (uint64_t l = r.line_off[q]; l < r.line_off[q + 1]; l++) {
            if (!r.keep[l]) continue;
            const uint32_t t = r.target[l];
            entry.append(buf, format_aln_line(buf, seq.keys[t], r.bits[l], r.seq_id[l], r.evalue[l], r.q_start[l], r.q_end[l], (int32_t) lens[q], r.db_start[l],
                                              r.db_end[l], (int32_t) lens[t]));
        }
        key = seq.keys[q]; ext = 0;
    });
    write_entries_parallel(a.pos[3], seq.dbtype ? seq.dbtype : DBTYPE_NUCLEOTIDES, seq.size(), threads, [&](size_t q, std::string &entry, uint32_t &key, uint8_t &ext) {
        entry.assign((const char *) out.data() + seq.offsets[q], lens[q]);
        entry.push_back('\n');
        key = seq.keys[q]; ext = seq.ext[q];
    });
    fprintf(stderr, "carpedeam_gpu iteration: %llu hits, %llu bases corrected; read DB + upload %.1f ms, GPU + export %.1f ms, format + write 3 DBs %.1f ms\n",
            (unsigned long long) h.n_hits, (unsigned long long) changed, t1 - t0, t2 - t1, now_ms() - t2);
    cdg_destroy(ctx);
    return EXIT_SUCCESS;
}

}   // namespace

int main(int argc, char **argv) {
    if (argc < 2) die("Usage: carpedeam_gpu <kmermatcher|rescorediagonal|ancient_correction|cyclecheck|createdb> ...");
    std::string cmd = argv[1];
    if (cmd == "kmermatcher") return kmermatcher(argc, argv);
    if (cmd == "rescorediagonal") return rescorediagonal(argc, argv);
    if (cmd == "ancient_correction") return ancient_correction(argc, argv);
    if (cmd == "iteration") return iteration(argc, argv);
    if (cmd == "cyclecheck") return cyclecheck(argc, argv);
    if (cmd == "createdb") return createdb(argc, argv);
    die("Invalid Command: " + cmd);
}

// carpedeam_gpu — drop-in stage binary: the three hot-path modules of the reference with the same positional
// database arguments and the flags the assembler workflow passes (data/nuclassemble.sh:105-126):
//
//   carpedeam_gpu kmermatcher        <seqDB> <prefDB>                 [flags]   (kmermatcher.cpp:786)
//   carpedeam_gpu rescorediagonal    <qDB> <tDB> <prefDB> <alnDB>     [flags]   (rescorediagonal.cpp:381)
//   carpedeam_gpu ancient_correction <seqDB> <alnDB> <outDB>          [flags]   (correction.cpp:492)
//
// Host side only: DBReader/DBWriter-compatible file access and the text formats; all arithmetic happens in
// libcarpedeam_gpu.so through the C ABI.  Errors print a message and exit(EXIT_FAILURE), as the reference does.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/carpedeam_gpu.h"
#include "formats.h"
#include "mmdb.h"

extern "C" int cdh_build_damage(const char *prefix, double seq_err, cdg_damage *out, char *errbuf, size_t errlen);

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
    Db seq; open_db(a.pos[0], seq);
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
    if (a.num("-c", 0.0) != 0.0) die("kmermatcher: only -c 0 (the assembler's setting) is supported");
    cdg_hits h;
    ck(ctx, cdg_kmermatch(ctx, &p, &h), "cdg_kmermatch");
    DbWriter w(a.pos[1], DBTYPE_PREFILTER_REV_RES);
    std::string entry; char buf[64];
    uint64_t hi = 0;
    for (size_t q = 0; q < seq.size(); q++) {
        entry.clear();
        entry.append(buf, format_pref_line(buf, seq.keys[q], 0, 0));   // `rep 0 0` / singleton `key 0 0` (kmermatcher.cpp:836, :710-730)
        bool isRep = false;
        while (hi < h.n_hits && h.rep[hi] == q) { entry.append(buf, format_pref_line(buf, seq.keys[h.target[hi]], h.score[hi], h.diag[hi])); hi++; isRep = true; }
        // only the singleton fill carries the ext flag (kmermatcher.cpp:727); entries of representatives are written without (:860,:923)
        w.write(entry.data(), entry.size(), seq.keys[q], isRep ? 0 : seq.ext[q]);
    }
    w.close();
    cdg_destroy(ctx);
    return EXIT_SUCCESS;
}

int rescorediagonal(int argc, char **argv) {
    Args a = parse_args(argc, argv, 2);
    if (a.pos.size() < 4) die("Usage: rescorediagonal <i:queryDB> <i:targetDB> <i:prefDB> <o:alnDB> [options]");
    if (a.pos[0] != a.pos[1]) die("rescorediagonal: query and target database must be the same (assembler usage)");
    if ((int) a.num("--rescore-mode", 0) != 3) die("rescorediagonal: only --rescore-mode 3 is implemented");
    Db seq; open_db(a.pos[0], seq);
    Db pref; open_db(a.pos[2], pref);
    cdg_ctx *ctx = make_ctx();
    std::vector<uint32_t> lens;
    upload(ctx, seq, lens);
    std::vector<uint64_t> lineOff(seq.size() + 1, 0);
    std::vector<uint32_t> tgt; std::vector<int32_t> score; std::vector<uint16_t> diag;
    std::vector<PrefLine> lines;
    for (size_t q = 0; q < seq.size(); q++) {
        long pi = pref.find(seq.keys[q]);
        lines.clear();
        if (pi >= 0) parse_pref_entry((const char *) pref.data.data() + pref.offsets[pi], lines);
        for (const PrefLine &l : lines) {
            long ti = seq.find(l.targetKey);
            if (ti < 0) die("Invalid database read for database data file=" + a.pos[1]);
            tgt.push_back((uint32_t) ti); score.push_back(l.score); diag.push_back(l.diag);
        }
        lineOff[q + 1] = tgt.size();
    }
    ck(ctx, cdg_set_pref_lines(ctx, lineOff.data(), seq.size(), tgt.data(), score.data(), diag.data(), tgt.size()), "cdg_set_pref_lines");
    cdg_rescore_params p;
    p.min_seq_id = (float) a.num("--min-seq-id", 0.0);
    p.evalue_thr = a.num("-e", 0.001);
    p.db_residues = 0;   // same DB: getAminoAcidDBSize() = sum of sequence lengths
    cdg_alns r;
    ck(ctx, cdg_rescore(ctx, &p, &r), "cdg_rescore");
    DbWriter w(a.pos[3], DBTYPE_ALIGNMENT_RES);
    std::string entry; char buf[256];
    for (size_t q = 0; q < seq.size(); q++) {
        entry.clear();
        for (uint64_t l = r.line_off[q]; l < r.line_off[q + 1]; l++) {
            if (!r.keep[l]) continue;
            uint32_t t = r.target[l];
            entry.append(buf, format_aln_line(buf, seq.keys[t], r.bits[l], r.seq_id[l], r.evalue[l], r.q_start[l], r.q_end[l], (int32_t) lens[q],
                                              r.db_start[l], r.db_end[l], (int32_t) lens[t]));
        }
        w.write(entry.data(), entry.size(), seq.keys[q], 0);
    }
    w.close();
    cdg_destroy(ctx);
    return EXIT_SUCCESS;
}

int ancient_correction(int argc, char **argv) {
    Args a = parse_args(argc, argv, 2);
    if (a.pos.size() < 3) die("Usage: ancient_correction <i:sequenceDB> <i:alnDB> <o:correctedDB> [options]");
    Db seq; open_db(a.pos[0], seq);
    Db aln; open_db(a.pos[1], aln);
    cdg_ctx *ctx = make_ctx();
    std::vector<uint32_t> lens;
    upload(ctx, seq, lens);
    cdg_damage dmg; char err[512];
    if (cdh_build_damage(a.get("--ancient-damage", "").c_str(), 0.01, &dmg, err, sizeof err) != 0) die(err);   // correction.cpp:153-156,196
    ck(ctx, cdg_set_damage(ctx, &dmg), "cdg_set_damage");
    std::vector<uint64_t> off(seq.size() + 1, 0);
    std::vector<uint32_t> tgt; std::vector<float> sid; std::vector<int32_t> qs, qe, ds, de;
    std::vector<AlnLine> lines;
    for (size_t q = 0; q < seq.size(); q++) {
        long ai = aln.find(seq.keys[q]);
        lines.clear();
        if (ai >= 0) parse_aln_entry((const char *) aln.data.data() + aln.offsets[ai], lines);
        for (const AlnLine &l : lines) {
            long ti = seq.find(l.targetKey);
            if (ti < 0) die("Invalid database read for database data file=" + a.pos[0]);
            tgt.push_back((uint32_t) ti); sid.push_back(l.seqId); qs.push_back(l.qStart); qe.push_back(l.qEnd); ds.push_back(l.dbStart); de.push_back(l.dbEnd);
        }
        off[q + 1] = tgt.size();
    }
    ck(ctx, cdg_set_alignments(ctx, off.data(), seq.size(), tgt.data(), sid.data(), qs.data(), qe.data(), ds.data(), de.data(), tgt.size()),
       "cdg_set_alignments");
    cdg_corr_params p;
    p.min_seqid_corr_reads = (float) a.num("--min-seq-id", 0.0);
    p.min_ryseq_id_corr_reads = (float) a.num("--min-ryseq-id-corr-reads", 0.99);
    std::vector<uint8_t> out(seq.data.size());
    uint64_t changed = 0;
    ck(ctx, cdg_correct(ctx, &p, out.data(), &changed), "cdg_correct");
    DbWriter w(a.pos[2], seq.dbtype ? seq.dbtype : DBTYPE_NUCLEOTIDES);
    std::string entry;
    for (size_t q = 0; q < seq.size(); q++) {
        entry.assign((const char *) out.data() + seq.offsets[q], lens[q]);
        entry.push_back('\n');                                           // correction.cpp:465-470
        w.write(entry.data(), entry.size(), seq.keys[q], seq.ext[q]);
    }
    w.close();
    fprintf(stderr, "ancient_correction: %llu bases corrected\n", (unsigned long long) changed);
    cdg_destroy(ctx);
    return EXIT_SUCCESS;
}

}   // namespace

int main(int argc, char **argv) {
    if (argc < 2) die("Usage: carpedeam_gpu <kmermatcher|rescorediagonal|ancient_correction> ...");
    std::string cmd = argv[1];
    if (cmd == "kmermatcher") return kmermatcher(argc, argv);
    if (cmd == "rescorediagonal") return rescorediagonal(argc, argv);
    if (cmd == "ancient_correction") return ancient_correction(argc, argv);
    die("Invalid Command: " + cmd);
}

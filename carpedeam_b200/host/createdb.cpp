// createdb on the host (FASTA / FASTQ [.gz] -> the sequence set the device path keeps resident).
//
// Follows lib/mmseqs/src/util/createdb.cpp (hard split mode, the assembler workflow's call: data/nuclassemble.sh:85) and the kseq
// reader it wraps (lib/mmseqs/lib/kseq/kseq.h): entry = name up to the first white space, comment = rest of the header line, sequence =
// the following lines up to a line starting with '>', '@' or '+', joined without line ends; a '+' line starts the qualities, which are
// read until they are as long as the sequence.  The sequence bytes are stored as they are (no case folding), entry `BASES\n\0`.
// --shuffle 1 (the default): entry i goes to split i % 32, the splits are concatenated and renumbered (createdb.cpp:60, :219,
// :277-280) — so key order = (i % 32, i / 32) order of the input, a fixed permutation.  --shuffle 0 keeps the input order.
//   cdh_createdb        parse the files -> flat set in DB layout (data, offsets, lengths; keys = 0..n-1) and, optionally, the header
//                       lines in the same order; with out_db != NULL also written as <out_db>{,.index,.dbtype}, <out_db>_h{,.index,.dbtype}
//                       <out_db>.source and (write_lookup) <out_db>.lookup
//   cdh_free_seqset     releases what cdh_createdb allocated
// The 2-bit packed arrays for cdg_upload_packed come from cdh_pack_seqs (pack.cpp) on this set.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/carpedeam_host.h"
#include "mmdb.h"

namespace {

struct GzReader {
    gzFile f = nullptr;
    unsigned char buf[1 << 16];
    int len = 0, pos = 0;
    bool eof = false;
    bool open(const char *path) { f = gzopen(path, "rb"); return f != nullptr; }
    void close() { if (f) gzclose(f); f = nullptr; }
    int getc() {
        if (pos >= len) {
            if (eof) return -1;
            len = gzread(f, buf, sizeof buf);
            pos = 0;
            if (len <= 0) { eof = true; len = 0; return -1; }
        }
        return buf[pos++];
    }
    // appends up to (not including) the delimiter: white space (space_sep) or end of line; returns the delimiter (-1 at end of file)
    int get_until(bool space_sep, std::string &out) {
        for (;;) {
            const int c = getc();
            if (c < 0) return -1;
            if (c == '\n' || (space_sep && (c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'))) return c;
            out.push_back((char) c);
        }
    }
};

struct Entry { std::string name, comment, seq; };

// kseq_read (kseq.h): returns 1 with an entry, 0 at end of file, -2 on truncated qualities
int read_entry(GzReader &r, int &last_char, Entry &e) {
    int c;
    if (last_char == 0) {
        while ((c = r.getc()) != -1 && c != '>' && c != '@') {}
        if (c == -1) return 0;
        last_char = c;
    }
    e.name.clear(); e.comment.clear(); e.seq.clear();
    c = r.get_until(true, e.name);
    if (c == -1 && e.name.empty()) return 0;
    if (c != '\n' && c != -1) { r.get_until(false, e.comment); }
    while (!e.comment.empty() && e.comment.back() == '\r') e.comment.pop_back();
    while ((c = r.getc()) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        e.seq.push_back((char) c);
        r.get_until(false, e.seq);
        while (!e.seq.empty() && e.seq.back() == '\r') e.seq.pop_back();
    }
    if (c == '>' || c == '@') last_char = c;
    if (c != '+') { if (c == -1) last_char = 0; return 1; }
    while ((c = r.getc()) != -1 && c != '\n') {}    // rest of the '+' line
    if (c == -1) return -2;
    size_t qual = 0;
    std::string q;
    while (qual < e.seq.size()) {
        q.clear();
        c = r.get_until(false, q);
        while (!q.empty() && q.back() == '\r') q.pop_back();
        qual += q.size();
        if (c == -1) break;
    }
    last_char = 0;
    return qual == e.seq.size() ? 1 : -2;
}

std::string base_name(const std::string &p) { const size_t s = p.find_last_of('/'); return s == std::string::npos ? p : p.substr(s + 1); }

}   // namespace

extern "C" void cdh_free_seqset(cdh_seqset *s) {
    if (!s) return;
    free(s->data); free(s->offsets); free(s->lengths); free(s->headers); free(s->header_offsets);
    memset(s, 0, sizeof *s);
}

// Util::parseFastaHeader (lib/mmseqs/src/commons/Util.cpp:173-256): the identifier inside the first token of a header line
static std::string fasta_identifier(const std::string &name) {
    if (name.empty()) return "";
    size_t offset = 0;
    if (name.size() >= 10 && name.compare(0, 10, "consensus_") == 0) offset = 10;
    struct Db { const char *prefix; unsigned length, bar; };
    static const Db dbs[] = {{"uc", 2, 0}, {"cl|", 3, 1}, {"sp|", 3, 1}, {"tr|", 3, 1}, {"gb|", 3, 1}, {"ref|", 4, 1}, {"pdb|", 4, 1}, {"bbs|", 4, 1},
                             {"lcl|", 4, 1}, {"pir||", 5, 1}, {"prf||", 5, 1}, {"gnl|", 4, 2}, {"pat|", 4, 2}, {"gi|", 3, 3}};
    for (const Db &d : dbs) {
        const size_t pl = strlen(d.prefix);
        if (name.size() < pl || name.compare(offset, pl, d.prefix) != 0) continue;   // Util::startWith(prefix, header, offset)
        size_t start = offset + d.length;
        for (unsigned j = 0; j + 1 < d.bar; j++) {
            const size_t end = name.find_first_of('|', start);
            if (end == std::string::npos) return "";
            start = end + 1;
        }
        size_t end = name.find_first_of('|', start);
        if (end == std::string::npos) end = name.find_first_of(" \n", start);
        if (end == std::string::npos) end = name.size();
        return name.substr(start, end - start);
    }
    size_t end = name.find_first_of(" \n", offset);
    if (end == std::string::npos) end = name.size();
    return name.substr(offset, end - offset);
}

extern "C" int cdh_createdb(const char *const *files, int n_files, int shuffle, int write_lookup, const char *out_db, cdh_seqset *out, char *errbuf,
                            size_t errlen) {
    auto fail = [&](const std::string &m) { if (errbuf && errlen) snprintf(errbuf, errlen, "%s", m.c_str()); return -1; };
    if (!files || n_files <= 0 || !out) return fail("cdh_createdb: no input");
    memset(out, 0, sizeof *out);
    std::vector<Entry> entries;
    std::vector<int> fileOf;
    for (int fi = 0; fi < n_files; fi++) {
        GzReader r;
        if (!r.open(files[fi])) return fail(std::string("Cannot open ") + files[fi]);
        int last = 0, rc;
        Entry e;
        while ((rc = read_entry(r, last, e)) == 1) {
            if (e.name.empty()) { r.close(); return fail("Fasta entry " + std::to_string(entries.size()) + " is invalid"); }   // createdb.cpp:153-156
            entries.push_back(e); fileOf.push_back(fi);
        }
        r.close();
        if (rc == -2) return fail(std::string("truncated quality string in ") + files[fi]);
    }
    const size_t n = entries.size();
    if (n == 0) return fail("The input files have no entry");   // createdb.cpp:267-274
    // storage order: split s holds the entries with index % splits == s (createdb.cpp:219), splits are concatenated and renumbered (:277-280)
    const size_t splits = shuffle ? 32 : 1;
    std::vector<size_t> order; order.reserve(n);
    for (size_t s = 0; s < splits; s++) for (size_t i = s; i < n; i += splits) order.push_back(i);
    size_t dataBytes = 0, hdrBytes = 0;
    for (const Entry &e : entries) { dataBytes += e.seq.size() + 2; hdrBytes += e.name.size() + (e.comment.empty() ? 0 : 1 + e.comment.size()) + 2; }
    out->data = (uint8_t *) malloc(dataBytes + 1); out->offsets = (uint64_t *) malloc((n + 1) * 8); out->lengths = (uint32_t *) malloc((n + 1) * 4);
    out->headers = (uint8_t *) malloc(hdrBytes + 1); out->header_offsets = (uint64_t *) malloc((n + 1) * 8);
    if (!out->data || !out->offsets || !out->lengths || !out->headers || !out->header_offsets) { cdh_free_seqset(out); return fail("out of memory"); }
    size_t off = 0, hoff = 0;
    for (size_t k = 0; k < n; k++) {
        const Entry &e = entries[order[k]];
        out->offsets[k] = off; out->lengths[k] = (uint32_t) e.seq.size();
        memcpy(out->data + off, e.seq.data(), e.seq.size()); off += e.seq.size();
        out->data[off++] = '\n'; out->data[off++] = '\0';
        out->header_offsets[k] = hoff;
        memcpy(out->headers + hoff, e.name.data(), e.name.size()); hoff += e.name.size();
        if (!e.comment.empty()) { out->headers[hoff++] = ' '; memcpy(out->headers + hoff, e.comment.data(), e.comment.size()); hoff += e.comment.size(); }
        out->headers[hoff++] = '\n'; out->headers[hoff++] = '\0';
    }
    out->offsets[n] = off; out->header_offsets[n] = hoff;
    out->n = n; out->data_bytes = off; out->header_bytes = hoff; out->total_residues = off - 2 * n;
    if (out_db) {
        cdh::DbWriter seq(out_db, cdh::DBTYPE_NUCLEOTIDES), hdr(std::string(out_db) + "_h", 12 /*DBTYPE_GENERIC_DB*/);
        for (size_t k = 0; k < n; k++) {
            seq.write((const char *) out->data + out->offsets[k], out->lengths[k] + 1, (uint32_t) k, 0);
            hdr.write((const char *) out->headers + out->header_offsets[k], (size_t) (out->header_offsets[k + 1] - out->header_offsets[k] - 1), (uint32_t) k, 0);
        }
        seq.close(); hdr.close();
        FILE *src = fopen((std::string(out_db) + ".source").c_str(), "w");
        if (!src) return fail(std::string("Cannot write ") + out_db + ".source");
        for (int fi = 0; fi < n_files; fi++) fprintf(src, "%d\t%s\n", fi, base_name(files[fi]).c_str());   // createdb.cpp:121-123
        fclose(src);
        if (write_lookup) {   // createdb.cpp:292-330: key, identifier of the header, number of the input file
            FILE *lk = fopen((std::string(out_db) + ".lookup").c_str(), "w");
            if (!lk) return fail(std::string("Cannot write ") + out_db + ".lookup");
            for (size_t k = 0; k < n; k++) fprintf(lk, "%zu\t%s\t%d\n", k, fasta_identifier(entries[order[k]].name).c_str(), fileOf[order[k]]);
            fclose(lk);
        }
    }
    return 0;
}

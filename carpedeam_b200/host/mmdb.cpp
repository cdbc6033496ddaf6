#include "mmdb.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/stat.h>

namespace cdh {

void die(const std::string &msg) {
    fprintf(stderr, "%s\n", msg.c_str());   // Debug(Debug::ERROR) << ...; EXIT(EXIT_FAILURE)
    exit(EXIT_FAILURE);
}

static bool file_exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

static void append_file(const std::string &p, std::vector<uint8_t> &out) {
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) die("Cannot open data file " + p);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    size_t old = out.size();
    out.resize(old + (size_t) sz);
    if (sz > 0 && fread(out.data() + old, 1, (size_t) sz, f) != (size_t) sz) die("Cannot read data file " + p);
    fclose(f);
}

long Db::find(uint32_t key) const {
    if (dense) return key < keys.size() ? (long) key : -1;   // keys 0..n-1 (createdb, and every DB the assembler derives from it): no search
    auto it = std::lower_bound(keys.begin(), keys.end(), key);
    if (it == keys.end() || *it != key) return -1;
    return (long) (it - keys.begin());
}

void open_db(const std::string &path, Db &db) {
    db = Db();
    if (file_exists(path)) append_file(path, db.data);
    else {
        // split data files written by per-thread DBWriters (<db>.0, <db>.1, ...): offsets run over their concatenation
        int i = 0;
        while (file_exists(path + "." + std::to_string(i))) { append_file(path + "." + std::to_string(i), db.data); i++; }
        if (i == 0) die("Could not open data file " + path + "!");
    }
    // the index is text, one `key\toffset\tlength[\text]` line per entry (DBReader.cpp:792-817): read whole and parsed by hand — sscanf per
    // line cost 0.4 us, a 10 M-entry index 4 s
    std::vector<uint8_t> idx;
    {
        FILE *f = fopen((path + ".index").c_str(), "rb");
        if (!f) die("Could not open index file " + path + ".index!");
        fclose(f);
        append_file(path + ".index", idx);
    }
    idx.push_back('\n'); idx.push_back('\0');
    struct Row { uint32_t key; uint64_t off, len; uint8_t ext; };
    auto number = [](const char *&q, unsigned long long &v) -> bool {   // blanks, then digits
        while (*q == ' ' || *q == '\t') q++;
        if (*q < '0' || *q > '9') return false;
        unsigned long long x = 0;
        while (*q >= '0' && *q <= '9') { x = x * 10 + (unsigned long long) (*q - '0'); q++; }
        v = x;
        return true;
    };
    // (one thread: splitting the text over the host threads measured slower — the per-thread row vectors cost more in first-touch page
    // faults than the parsing itself)
    std::vector<Row> rows;
    rows.reserve(idx.size() / 12);
    const char *p = (const char *) idx.data(), *end = p + idx.size() - 1;
    bool sorted = true;
    while (p < end) {
        unsigned long long a = 0, b = 0, c = 0, e = 0;
        const char *q = p;
        const bool ok = number(q, a) && number(q, b) && number(q, c);
        const bool hasExt = ok && number(q, e);
        while (*q != '\n') q++;
        p = q + 1;
        if (!ok) continue;
        Row r; r.key = (uint32_t) a; r.off = b; r.len = c; r.ext = hasExt ? (uint8_t) (e != 0) : 0;
        if (!rows.empty() && rows.back().key > r.key) sorted = false;
        rows.push_back(r);
    }
    if (!sorted) std::sort(rows.begin(), rows.end(), [](const Row &x, const Row &y) { return x.key < y.key; });
    db.keys.reserve(rows.size()); db.offsets.reserve(rows.size()); db.lengths.reserve(rows.size()); db.ext.reserve(rows.size());
    for (const Row &r : rows) {
        if (r.off + r.len > db.data.size()) die("Index of " + path + " points past the end of the data file");
        db.keys.push_back(r.key); db.offsets.push_back(r.off); db.lengths.push_back(r.len); db.ext.push_back(r.ext);
    }
    db.dense = true;
    for (size_t i = 0; i < db.keys.size() && db.dense; i++) db.dense = db.keys[i] == (uint32_t) i;
    FILE *t = fopen((path + ".dbtype").c_str(), "rb");
    if (t) { int32_t v = 0; if (fread(&v, 4, 1, t) == 1) db.dbtype = v & 0x7FFFFFFF; fclose(t); }
}

void DbWriter::write(const char *payload, size_t len, uint32_t key, uint8_t extFlag) {
    Entry e; e.key = key; e.off = data.size(); e.len = len + 1; e.ext = extFlag;
    data.insert(data.end(), (const uint8_t *) payload, (const uint8_t *) payload + len);
    data.push_back(0);
    entries.push_back(e);
}

void DbWriter::append(const DbWriter &other) {
    const uint64_t shift = data.size();
    data.insert(data.end(), other.data.begin(), other.data.end());
    for (Entry e : other.entries) { e.off += shift; entries.push_back(e); }
}

void DbWriter::close() { close_parts(nullptr, 0); }

// this writer's entries followed by those of `more[0..nMore)` (per-thread writers, in that order) as ONE database: the data of every
// part goes to the file as it is, only the index offsets are shifted (no concatenated copy of the payload in memory)
void DbWriter::close_parts(DbWriter *more, size_t nMore) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) die("Could not open " + path + " for writing!");
    if (!data.empty() && fwrite(data.data(), 1, data.size(), f) != data.size()) die("Could not write " + path);
    uint64_t shift = data.size();
    for (size_t m = 0; m < nMore; m++) {
        DbWriter &o = more[m];
        if (!o.data.empty() && fwrite(o.data.data(), 1, o.data.size(), f) != o.data.size()) die("Could not write " + path);
        entries.reserve(entries.size() + o.entries.size());
        for (Entry e : o.entries) { e.off += shift; entries.push_back(e); }
        shift += o.data.size();
        std::vector<uint8_t>().swap(o.data); std::vector<Entry>().swap(o.entries);
    }
    fclose(f);
    bool sorted = true;
    for (size_t i = 1; i < entries.size() && sorted; i++) sorted = entries[i - 1].key <= entries[i].key;
    if (!sorted) std::sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return a.key < b.key; });
    f = fopen((path + ".index").c_str(), "wb");
    if (!f) die("Could not open " + path + ".index for writing!");
    {   // `key\toffset\tlength\text\n` per entry, formatted into one buffer (fprintf per entry: 0.2 us)
        std::vector<char> text;
        text.resize(entries.size() * 56 + 16);
        char *o = text.data();
        auto put = [&](unsigned long long v) { char tmp[24]; int n = 0; do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v); while (n) *o++ = tmp[--n]; };
        for (const Entry &e : entries) { put(e.key); *o++ = '\t'; put(e.off); *o++ = '\t'; put(e.len); *o++ = '\t'; put(e.ext); *o++ = '\n'; }
        const size_t bytes = (size_t) (o - text.data());
        if (bytes && fwrite(text.data(), 1, bytes, f) != bytes) die("Could not write " + path + ".index");
    }
    fclose(f);
    f = fopen((path + ".dbtype").c_str(), "wb");
    if (!f) die("Could not open " + path + ".dbtype for writing!");
    int32_t v = dbtype;
    fwrite(&v, 4, 1, f);
    fclose(f);
}

}   // namespace cdh

// ---- C entry point for the CPU test-suite: read a database, write it again through DbWriter, return the entry count ----
extern "C" long cdh_db_roundtrip(const char *inPath, const char *outPath, int *dbtypeOut) {
    cdh::Db db;
    cdh::open_db(inPath, db);
    cdh::DbWriter w(outPath, db.dbtype);
    for (size_t i = 0; i < db.size(); i++)
        w.write((const char *) db.data.data() + db.offsets[i], db.lengths[i] ? db.lengths[i] - 1 : 0, db.keys[i], db.ext[i]);
    w.close();
    if (dbtypeOut) *dbtypeOut = db.dbtype;
    return (long) db.size();
}

// the same through `parts` writers over contiguous blocks of entries, closed into one database (what the stage binaries do per host thread)
extern "C" long cdh_db_roundtrip_parts(const char *inPath, const char *outPath, int parts) {
    cdh::Db db;
    cdh::open_db(inPath, db);
    if (parts < 1) parts = 1;
    std::vector<cdh::DbWriter> w;
    for (int t = 0; t < parts; t++) w.emplace_back(outPath, db.dbtype);
    const size_t n = db.size();
    for (int t = 0; t < parts; t++)
        for (size_t i = n * (size_t) t / (size_t) parts; i < n * (size_t) (t + 1) / (size_t) parts; i++)
            w[t].write((const char *) db.data.data() + db.offsets[i], db.lengths[i] ? db.lengths[i] - 1 : 0, db.keys[i], db.ext[i]);
    w[0].close_parts(parts > 1 ? &w[1] : nullptr, (size_t) (parts - 1));
    return (long) n;
}

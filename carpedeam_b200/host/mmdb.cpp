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
    FILE *f = fopen((path + ".index").c_str(), "r");
    if (!f) die("Could not open index file " + path + ".index!");
    struct Row { uint32_t key; uint64_t off, len; uint8_t ext; };
    std::vector<Row> rows;
    char line[512];
    while (fgets(line, sizeof line, f)) {
        Row r; unsigned long long a = 0, b = 0, c = 0, e = 0;
        int got = sscanf(line, "%llu\t%llu\t%llu\t%llu", &a, &b, &c, &e);
        if (got < 3) continue;
        r.key = (uint32_t) a; r.off = b; r.len = c; r.ext = got >= 4 ? (uint8_t) (e != 0) : 0;
        rows.push_back(r);
    }
    fclose(f);
    std::sort(rows.begin(), rows.end(), [](const Row &x, const Row &y) { return x.key < y.key; });
    for (const Row &r : rows) {
        if (r.off + r.len > db.data.size()) die("Index of " + path + " points past the end of the data file");
        db.keys.push_back(r.key); db.offsets.push_back(r.off); db.lengths.push_back(r.len); db.ext.push_back(r.ext);
    }
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

void DbWriter::close() {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) die("Could not open " + path + " for writing!");
    if (!data.empty() && fwrite(data.data(), 1, data.size(), f) != data.size()) die("Could not write " + path);
    fclose(f);
    std::sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return a.key < b.key; });
    f = fopen((path + ".index").c_str(), "w");
    if (!f) die("Could not open " + path + ".index for writing!");
    for (const Entry &e : entries) fprintf(f, "%u\t%llu\t%llu\t%u\n", e.key, (unsigned long long) e.off, (unsigned long long) e.len, (unsigned) e.ext);
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

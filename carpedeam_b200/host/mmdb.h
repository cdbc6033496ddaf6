// Minimal MMseqs2 database access for the drop-in stage drivers: the same files the reference's
// DBReader/DBWriter use (lib/mmseqs/src/commons/DBReader.cpp:792-817 index lines `key\toffset\tlength[\text]`,
// DBWriter.cpp:409-426; data entries end in '\0'; <db>.dbtype = 4-byte little-endian int).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace cdh {

enum { DBTYPE_NUCLEOTIDES = 1, DBTYPE_ALIGNMENT_RES = 5, DBTYPE_PREFILTER_REV_RES = 14 };   // Parameters.h

struct Db {
    std::vector<uint8_t> data;      // concatenation of <db> or <db>.0, <db>.1, ...
    std::vector<uint32_t> keys;     // ascending
    std::vector<uint64_t> offsets;
    std::vector<uint64_t> lengths;  // including the trailing '\0' (and '\n' for sequences)
    std::vector<uint8_t> ext;       // CarpeDeam's "was extended" column (0 when absent)
    int dbtype = 0;
    bool dense = false;             // keys[i] == i for every entry: find() needs no search
    size_t size() const { return keys.size(); }
    // index of `key` or -1
    long find(uint32_t key) const;
};

// Reads index + data (+dbtype).  On error prints the reference-style message and exits (EXIT_FAILURE),
// like every error path of the reference (MMseqsMPI.h:27-38).
void open_db(const std::string &path, Db &db);

struct DbWriter {
    std::string path;
    int dbtype;
    std::vector<uint8_t> data;
    struct Entry { uint32_t key; uint64_t off; uint64_t len; uint8_t ext; };
    std::vector<Entry> entries;
    DbWriter(const std::string &p, int type) : path(p), dbtype(type) {}
    // appends payload + '\0' (DBWriter::writeData adds the null byte, DBWriter.cpp:360-407)
    void write(const char *payload, size_t len, uint32_t key, uint8_t extFlag);
    void append(const DbWriter &other);   // entries of another (per-thread) writer behind this one's
    void close();   // data, .index (sorted by key, 4 columns), .dbtype
    void close_parts(DbWriter *more, size_t nMore);   // close() of this writer + the entries of further (per-thread) writers behind it, without copying their data
};

[[noreturn]] void die(const std::string &msg);

}   // namespace cdh

#include "formats.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace cdh {

size_t format_pref_line(char *buf, uint32_t targetKey, int32_t score, int16_t diag) {
    return (size_t) sprintf(buf, "%u\t%d\t%d\n", targetKey, score, (int) diag);
}

static const char *skip_line(const char *p) {
    while (*p != '\0' && *p != '\n') p++;
    return *p == '\n' ? p + 1 : p;
}

void parse_pref_entry(const char *data, std::vector<PrefLine> &out) {
    const char *p = data;
    while (*p != '\0') {
        char *e;
        PrefLine l;
        l.targetKey = (uint32_t) strtoul(p, &e, 10);
        l.score = (int32_t) strtol(e, &e, 10);
        l.diag = (uint16_t) (short) strtol(e, &e, 10);   // static_cast<unsigned short>(fast_atoi<short>)
        out.push_back(l);
        p = skip_line(p);
    }
}

// "1.00" for exactly 1, else "0." + optional zeros + (int)(seqId*1000) — the reference's own quirks included
static char *seqid_to_buffer(float seqId, char *b) {
    if (seqId == 1.0) { memcpy(b, "1.00", 4); return b + 4; }
    *b++ = '0'; *b++ = '.';
    if (seqId < 0.10) *b++ = '0';
    if (seqId < 0.01) *b++ = '0';
    return b + sprintf(b, "%d", (int) (seqId * 1000));
}

size_t format_aln_line(char *buf, uint32_t targetKey, int32_t bits, float seqId, double evalue, int32_t qStart, int32_t qEnd,
                       int32_t qLen, int32_t dbStart, int32_t dbEnd, int32_t dbLen) {
    char *p = buf;
    p += sprintf(p, "%u\t%d\t", targetKey, bits);
    p = seqid_to_buffer(seqId, p);
    p += sprintf(p, "\t%.3E\t%d\t%d\t%d\t%d\t%d\t%d\n", evalue, qStart, qEnd, qLen, dbStart, dbEnd, dbLen);
    return (size_t) (p - buf);
}

void parse_aln_entry(const char *data, std::vector<AlnLine> &out) {
    const char *p = data;
    while (*p != '\0') {
        char *e;
        AlnLine a;
        a.targetKey = (uint32_t) strtoul(p, &e, 10);
        strtol(e, &e, 10);                         // score
        a.seqId = (float) strtod(e, &e);           // double seqId = strtod(...) narrowed into result_t::seqId (float)
        strtod(e, &e);                             // evalue
        a.qStart = (int32_t) strtol(e, &e, 10);
        a.qEnd = (int32_t) strtol(e, &e, 10);
        strtol(e, &e, 10);                         // qLen
        a.dbStart = (int32_t) strtol(e, &e, 10);
        a.dbEnd = (int32_t) strtol(e, &e, 10);
        out.push_back(a);
        p = skip_line(p);
    }
}

}   // namespace cdh

// ---- C entry points for the CPU test-suite (ctypes): round trips of whole DB entries through the parsers / formatters ----
extern "C" size_t cdh_pref_roundtrip(const char *entry, char *out, size_t cap) {
    std::vector<cdh::PrefLine> lines;
    cdh::parse_pref_entry(entry, lines);
    size_t n = 0;
    char buf[64];
    for (const cdh::PrefLine &l : lines) {
        const size_t k = cdh::format_pref_line(buf, l.targetKey, l.score, (int16_t) l.diag);
        if (n + k >= cap) return 0;
        memcpy(out + n, buf, k); n += k;
    }
    out[n] = '\0';
    return n;
}
extern "C" size_t cdh_format_aln_line(char *buf, uint32_t targetKey, int32_t bits, float seqId, double evalue, int32_t qStart, int32_t qEnd,
                                      int32_t qLen, int32_t dbStart, int32_t dbEnd, int32_t dbLen) {
    return cdh::format_aln_line(buf, targetKey, bits, seqId, evalue, qStart, qEnd, qLen, dbStart, dbEnd, dbLen);
}
// fields of line `index` of an alignment entry as the correction stage reads them; returns the number of lines
extern "C" size_t cdh_parse_aln_entry(const char *entry, size_t index, uint32_t *targetKey, float *seqId, int32_t *coords4) {
    std::vector<cdh::AlnLine> lines;
    cdh::parse_aln_entry(entry, lines);
    if (index < lines.size()) {
        const cdh::AlnLine &a = lines[index];
        *targetKey = a.targetKey; *seqId = a.seqId; coords4[0] = a.qStart; coords4[1] = a.qEnd; coords4[2] = a.dbStart; coords4[3] = a.dbEnd;
    }
    return lines.size();
}

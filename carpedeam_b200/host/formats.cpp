#include "formats.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace cdh {

// ---- fast integer / %.3E formatting -------------------------------------------------------------------------------------------
// The formatters below produce exactly what the reference's Itoa / snprintf("%.3E") calls write (Matcher.cpp:356-401,
// QueryMatcher.h:114-126).  They exist because the text DBs are what a stage binary spends its time on once the kernels are fast:
// sprintf cost 0.9 us per alignment line (28 M lines per 10 M reads).  tests/test_host_formats.py checks them against snprintf /
// strtod on tens of millions of values (cdh_format_selftest).
static inline char *put_u32(char *p, uint32_t v) {
    char tmp[10]; int n = 0;
    do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
static inline char *put_i32(char *p, int32_t v) {
    if (v < 0) { *p++ = '-'; return put_u32(p, (uint32_t) (-(int64_t) v)); }
    return put_u32(p, (uint32_t) v);
}
// snprintf(p, "%.3E", v) for finite v: four significant digits, correctly rounded from the exact binary value (glibc rounds to nearest,
// ties to even on the exact value).  The scaled value v * 10^(3 - e10) is computed in long double (64-bit mantissa); whenever it lies
// within 1e-9 of a tie — or anything else looks unusual — snprintf decides.
static const long double POW10L[] = {1e0L, 1e1L, 1e2L, 1e3L, 1e4L, 1e5L, 1e6L, 1e7L, 1e8L, 1e9L, 1e10L, 1e11L, 1e12L, 1e13L, 1e14L, 1e15L, 1e16L, 1e17L, 1e18L,
                                     1e19L, 1e20L, 1e21L, 1e22L, 1e23L, 1e24L, 1e25L, 1e26L, 1e27L};
static long double pow10l_int(int e) {   // 10^e for |e| <= 350, from exactly representable pieces (10^27 is exact in a 64-bit mantissa)
    long double r = 1.0L;
    int a = e < 0 ? -e : e;
    while (a >= 27) { r *= POW10L[27]; a -= 27; }
    r *= POW10L[a];
    return e < 0 ? 1.0L / r : r;
}
char *put_e3(char *p, double v) {
    if (v == 0.0 && !std::signbit(v)) { memcpy(p, "0.000E+00", 9); return p + 9; }
    if (!(v > 0.0) || !std::isfinite(v) || v < 1e-300 || v > 1e300) return p + sprintf(p, "%.3E", v);
    int e10 = (int) std::floor(std::log10(v));
    long double scaled = (long double) v * pow10l_int(3 - e10);            // in [1000, 10000) up to the error of log10 / the powers
    if (scaled < 1000.0L) { e10--; scaled *= 10.0L; }
    else if (scaled >= 10000.0L) { e10++; scaled /= 10.0L; }
    const long double fl = std::floor(scaled), frac = scaled - fl;
    // only a value within the error bound of a tie is ambiguous (next to an integer both neighbours round to that integer)
    if (fl < 1000.0L || fl > 9999.0L || std::fabs((double) (frac - 0.5L)) < 1e-9) return p + sprintf(p, "%.3E", v);
    uint32_t d = (uint32_t) fl + (frac > 0.5L ? 1u : 0u);
    if (d == 10000u) { d = 1000u; e10++; }
    p[0] = (char) ('0' + d / 1000); p[1] = '.'; p[2] = (char) ('0' + d / 100 % 10); p[3] = (char) ('0' + d / 10 % 10); p[4] = (char) ('0' + d % 10);
    p[5] = 'E'; p[6] = e10 < 0 ? '-' : '+';
    const uint32_t ae = (uint32_t) (e10 < 0 ? -e10 : e10);
    p += 7;
    if (ae < 10) { *p++ = '0'; *p++ = (char) ('0' + ae); return p; }
    return put_u32(p, ae);
}

size_t format_pref_line(char *buf, uint32_t targetKey, int32_t score, int16_t diag) {
    char *p = put_u32(buf, targetKey);
    *p++ = '\t'; p = put_i32(p, score);
    *p++ = '\t'; p = put_i32(p, (int32_t) diag);
    *p++ = '\n'; *p = '\0';
    return (size_t) (p - buf);
}

static const char *skip_line(const char *p) {
    while (*p != '\0' && *p != '\n') p++;
    return *p == '\n' ? p + 1 : p;
}
// strtol / strtoul on the fields the reference writes (optional blanks, optional sign, digits)
static inline const char *get_i64(const char *p, int64_t &v) {
    while (*p == ' ' || *p == '\t') p++;
    bool neg = false;
    if (*p == '-') { neg = true; p++; } else if (*p == '+') p++;
    int64_t x = 0;
    while (*p >= '0' && *p <= '9') { x = x * 10 + (*p - '0'); p++; }
    v = neg ? -x : x;
    return p;
}
static inline const char *skip_field(const char *p) {
    while (*p == ' ' || *p == '\t') p++;
    while (*p != '\0' && *p != '\t' && *p != '\n' && *p != ' ') p++;
    return p;
}

void parse_pref_entry(const char *data, std::vector<PrefLine> &out) {
    const char *p = data;
    while (*p != '\0') {
        PrefLine l; int64_t v;
        const char *e = get_i64(p, v); l.targetKey = (uint32_t) v;
        e = get_i64(e, v); l.score = (int32_t) v;
        e = get_i64(e, v); l.diag = (uint16_t) (short) v;   // static_cast<unsigned short>(fast_atoi<short>)
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
    return put_i32(b, (int) (seqId * 1000));
}

size_t format_aln_line(char *buf, uint32_t targetKey, int32_t bits, float seqId, double evalue, int32_t qStart, int32_t qEnd,
                       int32_t qLen, int32_t dbStart, int32_t dbEnd, int32_t dbLen) {
    char *p = put_u32(buf, targetKey);
    *p++ = '\t'; p = put_i32(p, bits);
    *p++ = '\t'; p = seqid_to_buffer(seqId, p);
    *p++ = '\t'; p = put_e3(p, evalue);
    *p++ = '\t'; p = put_i32(p, qStart);
    *p++ = '\t'; p = put_i32(p, qEnd);
    *p++ = '\t'; p = put_i32(p, qLen);
    *p++ = '\t'; p = put_i32(p, dbStart);
    *p++ = '\t'; p = put_i32(p, dbEnd);
    *p++ = '\t'; p = put_i32(p, dbLen);
    *p++ = '\n'; *p = '\0';
    return (size_t) (p - buf);
}

// strtod of the identity field as the reference wrote it ("1.00", "0.957", "0.05"): plain decimal digits, value = D / 10^n with D and
// 10^n exact in a double, so the (correctly rounded) quotient IS the correctly rounded decimal; anything else goes to strtod
static const char *get_seqid(const char *p, float &out) {
    while (*p == ' ' || *p == '\t') p++;
    const char *q = p;
    uint64_t D = 0; int nd = 0, nfrac = 0; bool dot = false;
    for (;; q++) {
        if (*q >= '0' && *q <= '9') { D = D * 10 + (uint64_t) (*q - '0'); nd++; if (dot) nfrac++; }
        else if (*q == '.' && !dot) dot = true;
        else break;
    }
    static const double P10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    if (nd == 0 || nd > 15 || nfrac > 15 || (*q != '\t' && *q != '\n' && *q != '\0' && *q != ' ')) {
        char *e;
        out = (float) strtod(p, &e);   // exponents, signs, inf / nan: the general parser
        return e;
    }
    out = (float) ((double) D / P10[nfrac]);
    return q;
}

void parse_aln_entry(const char *data, std::vector<AlnLine> &out) {
    const char *p = data;
    while (*p != '\0') {
        AlnLine a; int64_t v;
        const char *e = get_i64(p, v); a.targetKey = (uint32_t) v;
        e = get_i64(e, v);                          // score
        e = get_seqid(e, a.seqId);                  // double seqId = strtod(...) narrowed into result_t::seqId (float)
        e = skip_field(e);                          // evalue
        e = get_i64(e, v); a.qStart = (int32_t) v;
        e = get_i64(e, v); a.qEnd = (int32_t) v;
        e = get_i64(e, v);                          // qLen
        e = get_i64(e, v); a.dbStart = (int32_t) v;
        e = get_i64(e, v); a.dbEnd = (int32_t) v;
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

// self test of the fast formatters / parsers against the C library on `n` pseudo-random values (seeded): returns the number of mismatches
extern "C" uint64_t cdh_format_selftest(uint64_t n, uint64_t seed) {
    uint64_t bad = 0, x = seed * 2862933555777941757ull + 3037000493ull;
    auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    char a[64], b[64];
    for (uint64_t i = 0; i < n; i++) {
        // E-values: random mantissa bits, exponents over the range the stages see and beyond; decimal ties (x.xxx5 exactly) now and then
        const uint64_t r = next();
        double v;
        if ((r & 15) == 0) v = (double) (1000 + next() % 9000) * 1e-3 + 0.0005;                         // near a rounding boundary
        else if ((r & 15) == 1) v = std::ldexp((double) (next() % 10000), (int) (next() % 200) - 100);    // few significant bits
        else if ((r & 15) == 2) v = (double) (next() % 100000) * std::pow(10.0, -(double) (next() % 45));  // short decimals (exact integers after scaling)
        else if ((r & 15) == 3) v = (double) (next() % 3 == 0 ? 0 : 1) * 1e-3;                              // 0 and the threshold itself
        else { v = std::ldexp(1.0 + (double) (next() >> 11) / 9007199254740992.0, (int) (next() % 400) - 330); }
        *cdh::put_e3(a, v) = '\0';
        sprintf(b, "%.3E", v);
        if (strcmp(a, b) != 0) bad++;
        // identities as the reference prints them
        const float sid = (float) (next() % 100001) / 100000.0f;
        char line[256], ref[256];
        cdh::format_aln_line(line, (uint32_t) next(), (int32_t) (next() % 400) - 10, sid, v, (int32_t) (next() % 70000), (int32_t) (next() % 70000), 70000,
                             (int32_t) (next() % 70000) - 1, (int32_t) (next() % 70000), 65000);
        std::vector<cdh::AlnLine> out;
        cdh::parse_aln_entry(line, out);
        // reference parse: strtoul / strtol / strtod field by field
        char *e; const char *p = line;
        const uint32_t key = (uint32_t) strtoul(p, &e, 10); strtol(e, &e, 10);
        const float sid2 = (float) strtod(e, &e); strtod(e, &e);
        const int32_t qs = (int32_t) strtol(e, &e, 10), qe = (int32_t) strtol(e, &e, 10); strtol(e, &e, 10);
        const int32_t ds = (int32_t) strtol(e, &e, 10), de = (int32_t) strtol(e, &e, 10);
        if (out.size() != 1 || out[0].targetKey != key || out[0].seqId != sid2 || out[0].qStart != qs || out[0].qEnd != qe || out[0].dbStart != ds || out[0].dbEnd != de) bad++;
        (void) ref;
    }
    return bad;
}

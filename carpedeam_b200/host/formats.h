// Text formats between the stages (the reference keeps them in DBs between processes).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace cdh {

// prefilter line `target\tscore\tdiag\n`  (QueryMatcher::prefilterHitToBuffer, QueryMatcher.h:114-126)
size_t format_pref_line(char *buf, uint32_t targetKey, int32_t score, int16_t diag);
// QueryMatcher::parsePrefilterHit (QueryMatcher.h:81-104): diagonal is read back as unsigned short
struct PrefLine { uint32_t targetKey; int32_t score; uint16_t diag; };
void parse_pref_entry(const char *data, std::vector<PrefLine> &out);

// snprintf("%.3E") without snprintf (exactly the same characters; returns the end)
char *put_e3(char *p, double v);
// alignment line (Matcher::resultToBuffer, Matcher.cpp:356-401; Util::fastSeqIdToBuffer, Util.cpp:278-307)
size_t format_aln_line(char *buf, uint32_t targetKey, int32_t bits, float seqId, double evalue, int32_t qStart, int32_t qEnd,
                       int32_t qLen, int32_t dbStart, int32_t dbEnd, int32_t dbLen);
// Matcher::parseAlignmentRecord (Matcher.cpp:274-353): the fields ancient_correction uses
struct AlnLine { uint32_t targetKey; float seqId; int32_t qStart, qEnd, dbStart, dbEnd; };
void parse_aln_entry(const char *data, std::vector<AlnLine> &out);

}   // namespace cdh

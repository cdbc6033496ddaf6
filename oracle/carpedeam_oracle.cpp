// TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of CarpeDeam's read-level hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
// product (carpedeam_b200/) never links, imports or executes it.
//
// Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py) against
// databases produced by the unmodified reference binary (oracle/_ref/carpedeam_ref, built by
// oracle/build_ref.py, fixtures by oracle/make_golden.py): pref/aln/correction DBs of config 1
// iterations 0, 1, 5, 9, and ALP E-value known answers (tests/golden/alp_kat.txt).
//
// Scalar, single-threaded, written for clarity; flat arrays in, flat arrays out (ctypes friendly).
// Reference citations are relative to /root/reference.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <string>
#include <vector>

typedef uint64_t u64;
typedef uint32_t u32;
static const u64 B63 = 1ull << 63;

// ---------------------------------------------------------------------------------------------
// a1: letter mapping  (lib/mmseqs/src/commons/NucleotideMatrix.cpp:17-61, alphabet nucleotide.out:4)
// codes: A0 C1 T2 G3 X4
static int map_letter(unsigned char c) {
    if (c >= 128) return 4;
    switch (toupper(c)) {
        case 'A': return 0;
        case 'C': return 1;
        case 'T': case 'U': case 'W': return 2;
        case 'G': case 'K': case 'B': case 'D': case 'V': case 'R': case 'S': return 3;
        case 'M': case 'Y': case 'H': return 1;
        default: return 4;
    }
}
static const char NUM2AA[5] = {'A', 'C', 'T', 'G', 'X'};
// NucleotideMatrix::reverseResidue: A<->T, C<->G, X stays
static int rev_code(int c) { return c == 4 ? 4 : (c ^ 2); }

// ---------------------------------------------------------------------------------------------
// a3: XXH64 of one little-endian u64 (lib/mmseqs/lib/xxhash v0.8.0; kmermatcher.cpp:33-38)
static inline u64 rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
static u64 xxh64_u64(u64 v, u64 seed) {
    const u64 P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
              P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
    u64 h = seed + P5 + 8;
    u64 k = rotl64(v * P2, 31) * P1;
    h ^= k;
    h = rotl64(h, 27) * P1 + P4;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}
extern "C" u64 orc_xxh64_u64(u64 v, u64 seed) { return xxh64_u64(v, seed); }

// a2: Util::revComplement (Util.cpp:601-638): reverse the k 2-bit groups, complement = xor 2
static u64 revcomp_idx(u64 idx, int k) {
    u64 r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | ((idx & 3) ^ 2); idx >>= 2; }
    return r;
}
extern "C" u64 orc_revcomp(u64 idx, int k) { return revcomp_idx(idx, k); }

// ---------------------------------------------------------------------------------------------
// kmermatcher  (lib/mmseqs/src/linclust/kmermatcher.cpp)
struct KmerRec { u64 kmer; u32 id; int seqLen; int pos; };
struct SeqPos { unsigned short score; u64 kmer; u32 pos; };

struct KmOut {
    std::vector<u32> rep, tgt; std::vector<int32_t> score; std::vector<int16_t> diag;
    std::vector<KmerRec> r1, r2;
    u64 nR1 = 0, nR2 = 0;
} g_km;

// a4: fillKmerPositionArray<NUCL> per sequence (kmermatcher.cpp:125-349)
static void extract_seq(const char *s, int L, u32 id, int k, int kmersPerSeq, float scale, u64 seed,
                        unsigned hashLo, unsigned hashHi, std::vector<KmerRec> &out) {
    std::vector<unsigned char> num(L);
    for (int i = 0; i < L; i++) num[i] = (unsigned char) map_letter((unsigned char) s[i]);
    // identity record: Util::hash (Util.h:337-346) then XXH64 (kmermatcher.cpp:133-137)
    u64 h = 0;
    for (int i = 0; i < L; i++) h = h * 31 + num[i];
    u64 seqHash = xxh64_u64(h, seed);
    if (L < k) {   // no window: only the identity record (:244); skips the 256 KB histogram below (sets with millions of such entries)
        if ((unsigned short) seqHash >= hashLo && (unsigned short) seqHash <= hashHi) { KmerRec r; r.kmer = seqHash; r.id = id; r.pos = 0; r.seqLen = L; out.push_back(r); }
        return;
    }
    std::vector<SeqPos> km;
    std::vector<unsigned> scoreDist(65536, 0);
    unsigned hier[128] = {0};
    for (int p = 0; p + k <= L; p++) {
        bool hasX = false; u64 idx = 0;
        for (int j = 0; j < k; j++) { if (num[p + j] == 4) hasX = true; idx = (idx << 2) | (num[p + j] & 3); }
        if (hasX) continue;                                       // :146
        u64 rc = revcomp_idx(idx, k);
        if (rc == idx) continue;                                  // :159
        bool useRc = rc < idx;
        u64 canon = useRc ? rc : idx;
        unsigned short sc = (unsigned short) xxh64_u64(canon, seed);   // :164
        SeqPos sp; sp.kmer = useRc ? canon : (canon | B63); sp.pos = useRc ? (u32) (L - p - k) : (u32) p; sp.score = sc;
        km.push_back(sp);
        scoreDist[sc]++; hier[sc >> 9]++;
    }
    size_t cnt = km.size();
    // :226  int - 1 + float * int, evaluated in float, then truncated to size_t
    float consF = (float) (kmersPerSeq - 1) + scale * (float) L;
    size_t considered = std::min((size_t) consF, cnt);
    unsigned threshold = 0; size_t inBins = 0;
    if (cnt > 0) {                                                // :228-241
        size_t ht = 0;
        for (ht = 0; ht < 128 && inBins < considered; ht++) inBins += hier[ht];
        ht -= (ht > 0) ? 1 : 0;
        inBins -= hier[ht];
        for (threshold = (unsigned) ht * 512; threshold <= 65535 && inBins < considered; threshold++) inBins += scoreDist[threshold];
    }
    int tooMuch = (int) (inBins - considered);
    if ((unsigned short) seqHash >= hashLo && (unsigned short) seqHash <= hashHi) {   // :244
        KmerRec r; r.kmer = seqHash; r.id = id; r.pos = 0; r.seqLen = L; out.push_back(r);
    }
    // :269-273 compareByScoreReverse (kmermatcher.h:29-46)
    std::sort(km.begin(), km.end(), [](const SeqPos &a, const SeqPos &b) {
        if (a.score != b.score) return a.score < b.score;
        u64 ka = a.kmer | B63, kb = b.kmer | B63;
        if (ka != kb) return ka < kb;
        if (a.pos != b.pos) return a.pos < b.pos;
        return (a.kmer & B63) > (b.kmer & B63);   // full tie in the reference; forward first here
    });
    size_t selected = 0;
    for (size_t i = 0; i < cnt && selected < considered; i++) {   // :275-347
        u64 kmer = km[i].kmer | B63;
        if (i + 1 < cnt) {
            u64 next = km[i + 1].kmer | B63;
            if (kmer == next) {
                while (kmer == next && i < cnt) {
                    i++;
                    if (i >= cnt) break;
                    next = km[i].kmer | B63;
                }
            }
        }
        if (i >= cnt) break;
        if (km[i].score < threshold) {
            if (km[i].score == threshold - 1 && tooMuch) {
                tooMuch--;
                threshold -= (tooMuch == 0) ? 1 : 0;
            }
            selected++;
            if (km[i].score >= hashLo && km[i].score <= hashHi) {
                KmerRec r; r.kmer = km[i].kmer; r.id = id; r.pos = (int) km[i].pos; r.seqLen = L; out.push_back(r);
            }
        }
    }
}

// -c / --cov-mode of the k-mer matcher (assignGroup's Util::canBeCovered, kmermatcher.cpp:526-529); the assembler passes -c 0
static bool can_be_covered(float covThr, int covMode, float q, float t);
static float g_covThr = 0.f; static int g_covMode = 1;
extern "C" void orc_set_kmermatch_cov(float covThr, int covMode) { g_covThr = covThr; g_covMode = covMode; }
// a5..a8.  ids are sequence indices; the caller guarantees keys ascend with the index, so
// ordering by index == ordering by dbKey (KmerPosition.id is the dbKey, kmermatcher.cpp:143).
// hashLo/hashHi: one split of the reference's split loop (kmermatcher.cpp:665-675): only records whose 16-bit hash lies in the range
// (:244, :315), grouped and written like a whole run (the first-run quirk applies to the split's first run)
extern "C" long orc_kmermatch_range(const char *blob, const u64 *off, const u32 *len, size_t n, int k, int kmersPerSeq,
                                    float scale, u64 seed, int onlyExtendable, int keepRecords, unsigned hashLo, unsigned hashHi);
extern "C" long orc_kmermatch(const char *blob, const u64 *off, const u32 *len, size_t n, int k, int kmersPerSeq,
                              float scale, u64 seed, int onlyExtendable, int keepRecords) {
    return orc_kmermatch_range(blob, off, len, n, k, kmersPerSeq, scale, seed, onlyExtendable, keepRecords, 0, 65535);
}
extern "C" long orc_kmermatch_range(const char *blob, const u64 *off, const u32 *len, size_t n, int k, int kmersPerSeq,
                                    float scale, u64 seed, int onlyExtendable, int keepRecords, unsigned hashLo, unsigned hashHi) {
    KmOut &o = g_km;
    o = KmOut();
    std::vector<KmerRec> recs;
    for (size_t i = 0; i < n; i++) extract_seq(blob + off[i], (int) len[i], (u32) i, k, kmersPerSeq, scale, seed, hashLo, hashHi, recs);
    o.nR1 = recs.size();
    // a5 sort 1: compareRepSequenceAndIdAndPosReverse (kmermatcher.h:76-96); bit 63 breaks full ties here
    std::sort(recs.begin(), recs.end(), [](const KmerRec &a, const KmerRec &b) {
        u64 ka = a.kmer | B63, kb = b.kmer | B63;
        if (ka != kb) return ka < kb;
        if (a.seqLen != b.seqLen) return a.seqLen > b.seqLen;
        if (a.id != b.id) return a.id < b.id;
        if (a.pos != b.pos) return a.pos < b.pos;
        return (a.kmer & B63) > (b.kmer & B63);
    });
    if (keepRecords) o.r1 = recs;
    // a6 assignGroup (kmermatcher.cpp:453-562)
    struct G { u32 rep; bool need; u32 id; int diag; };
    std::vector<G> grp;
    size_t N = recs.size(), i = 0; bool first = true;
    while (i < N) {
        size_t j = i;
        while (j < N && (recs[j].kmer | B63) == (recs[i].kmer | B63)) j++;
        if (j - i > 1) {
            const KmerRec &r = recs[i];
            bool repRev = first ? false : ((r.kmer & B63) == 0);   // :466 vs :545 (first-group quirk)
            for (size_t m = i; m < j; m++) {
                const KmerRec &t = recs[m];
                bool tRev = (t.kmer & B63) == 0;
                int q, tp; bool need;
                if (repRev && !tRev) { q = r.pos; tp = t.pos; need = true; }
                else if (repRev && tRev) { q = (r.seqLen - 1) - r.pos; tp = (t.seqLen - 1) - t.pos; need = false; }
                else if (!repRev && tRev) { q = (r.seqLen - 1) - r.pos; tp = (t.seqLen - 1) - t.pos; need = true; }
                else { q = r.pos; tp = t.pos; need = false; }
                int d = q - tp;
                bool canExtend = d < 0 || d > (r.seqLen - t.seqLen);
                const bool canBeCovered = can_be_covered(g_covThr, g_covMode, (float) r.seqLen, (float) t.seqLen);   // :527-529
                if ((!onlyExtendable && canBeCovered) || (canExtend && onlyExtendable)) { G g; g.rep = r.id; g.need = need; g.id = t.id; g.diag = d; grp.push_back(g); }
            }
        }
        first = false;
        i = j;
    }
    o.nR2 = grp.size();
    // a7 sort 2: compareRepSequenceAndIdAndDiagReverse (kmermatcher.h:98-114); full ties (unspecified in the
    // reference, whose comparator masks bit 63): needRev==false first
    std::sort(grp.begin(), grp.end(), [](const G &a, const G &b) {
        if (a.rep != b.rep) return a.rep < b.rep;
        if (a.id != b.id) return a.id < b.id;
        if (a.diag != b.diag) return a.diag < b.diag;
        return (int) a.need < (int) b.need;
    });
    if (keepRecords) { o.r2.clear(); for (auto &g : grp) { KmerRec r; r.kmer = g.rep | (g.need ? 0 : B63); r.id = g.id; r.seqLen = 0; r.pos = g.diag; o.r2.push_back(r); } }
    // a8 writeKmerMatcherResult (kmermatcher.cpp:815-930) incl. the rep-boundary bleed (:886-888)
    size_t M = grp.size(); i = 0;
    while (i < M) {
        u32 rep = grp[i].rep; size_t j = i;
        while (j < M && grp[j].rep == rep) j++;
        size_t m = i;
        while (m < j) {
            u32 tid = grp[m].id;
            size_t e = m; int dc = 0, mx = 0, best = grp[m].diag, prev = grp[m].diag, top = 0; bool brev = grp[m].need;
            while (e < M && grp[e].id == tid) {
                dc = (prev == grp[e].diag) ? dc + 1 : 1;
                if (dc >= mx) { best = grp[e].diag; mx = dc; brev = grp[e].need; }
                prev = grp[e].diag; e++; top++;
            }
            if (tid != rep) { o.rep.push_back(rep); o.tgt.push_back(tid); o.score.push_back(brev ? -top : top); o.diag.push_back((int16_t) best); }
            while (m < j && grp[m].id == tid) m++;
        }
        i = j;
    }
    return (long) o.rep.size();
}
extern "C" void orc_km_hits(u32 *rep, u32 *tgt, int32_t *score, int16_t *diag) {
    size_t H = g_km.rep.size();
    memcpy(rep, g_km.rep.data(), H * 4); memcpy(tgt, g_km.tgt.data(), H * 4);
    memcpy(score, g_km.score.data(), H * 4); memcpy(diag, g_km.diag.data(), H * 2);
}
extern "C" void orc_km_counts(u64 *r1, u64 *r2) { *r1 = g_km.nR1; *r2 = g_km.nR2; }
// debugging aid for kernel bring-up: records after sort 1 (stage 1) / after sort 2 (stage 2)
extern "C" void orc_km_records(int stage, u64 *kmer, u32 *id, int32_t *seqLen, int32_t *pos) {
    const std::vector<KmerRec> &v = stage == 1 ? g_km.r1 : g_km.r2;
    for (size_t i = 0; i < v.size(); i++) { kmer[i] = v[i].kmer; id[i] = v[i].id; seqLen[i] = v[i].seqLen; pos[i] = v[i].pos; }
}

// ---------------------------------------------------------------------------------------------
// E-values: Karlin-Altschul with ALP finite-size correction
// (lib/mmseqs/src/alignment/EvalueComputation.h:36-40; lib/mmseqs/lib/alp/sls_pvalues.cpp:366-545,
//  sls_alignment_evaluer.hpp:154-167, sls_basic.hpp:195-198).  Gumbel parameters for the ungapped
// nucleotide.out matrix are what ALP's initGapless computes at start-up (dumped with oracle/ref_probe.cpp).
static const double ALP_LAMBDA = 0.63373155264486880078, ALP_K = 0.40796623464181452912,
                    ALP_A = 0.69454686319701297581, ALP_ALPHA = 0.83333515157614945768,
                    ALP_SIGMA = 0.83333515157614945768, ALP_THR = 2.6299310744375534732;
static double alp_area(double y, double m, double n) {   // get_appr_tail_prob_with_cov_without_errors, blast_=false
    const double const_val = 1 / sqrt(2.0 * 3.1415926535897932384626433832795);
    double m_li_y = m - (ALP_A * y + 0.0);
    double vi_y = std::max(ALP_THR, ALP_ALPHA * y + 0.0);
    double sqrt_vi_y = sqrt(vi_y);
    double m_F = (sqrt_vi_y == 0.0) ? 1e100 : m_li_y / sqrt_vi_y;
    double P_m_F = 0.5 * erfc(-sqrt(0.5) * m_F);
    double E_m_F = -const_val * exp(-0.5 * m_F * m_F);
    double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;
    double n_lj_y = n - (ALP_A * y + 0.0);
    double vj_y = std::max(ALP_THR, ALP_ALPHA * y + 0.0);
    double sqrt_vj_y = sqrt(vj_y);
    double n_F = (sqrt_vj_y == 0.0) ? 1e100 : n_lj_y / sqrt_vj_y;
    double P_n_F = 0.5 * erfc(-sqrt(0.5) * n_F);
    double E_n_F = -const_val * exp(-0.5 * n_F * n_F);
    double p2 = n_lj_y * P_n_F - sqrt_vj_y * E_n_F;
    double c_y = std::max(ALP_THR, ALP_SIGMA * y + 0.0);
    return p1 * p2 + c_y * P_m_F * P_n_F;
}
extern "C" void orc_evalue(double score, double qLen, double dbRes, double *evalue, double *bits) {
    // EvalueComputation::computeEvalue: evaluePerArea(score) * area(score, seqLength, dbResCount);
    // AlignmentEvaluer::area passes (seqlen2_, seqlen1_) = (dbRes, qLen) as (m_, n_)
    double epa = ALP_K * exp(-ALP_LAMBDA * score);
    *evalue = epa * alp_area(score, dbRes, qLen);
    *bits = (ALP_LAMBDA * score - log(ALP_K)) / log(2.0);
}

// ---------------------------------------------------------------------------------------------
// rescorediagonal mode 3 (lib/mmseqs/src/alignment/rescorediagonal.cpp:194-334, DistanceCalculator.h:93-220)
static inline int sub_score(unsigned char a, unsigned char b) {   // createAsciiSubMat over nucleotide.out
    int x = map_letter(a), y = map_letter(b);
    return (x == y && x != 4) ? 2 : -3;
}
struct LocalAln { int startPos = -1, endPos = -1; unsigned score = 0; int diagonalLen = 0; unsigned dist = 0; int diagonal = 0; };
static void global_dist(const char *s1, const char *s2, unsigned length, LocalAln &res) {   // :204-220
    unsigned first = (s1[0] == '*' || s2[0] == '*') ? 1 : 0;
    unsigned last = length - 1;
    if (last > 0 && (s1[length - 1] == '*' || s2[length - 1] == '*')) last--;
    int64_t score = 0;
    for (unsigned p = first; p <= last; p++) score += sub_score((unsigned char) s1[p], (unsigned char) s2[p]);
    score = std::max(score, (int64_t) 0);
    res.startPos = (int) first; res.endPos = (int) last; res.score = (unsigned) score;
}
static LocalAln by_diagonal(const char *q, unsigned qLen, const char *t, unsigned tLen, int diagonal) {   // :115-175
    LocalAln res; unsigned ad = (unsigned) abs(diagonal);
    res.dist = ad; res.diagonal = diagonal;
    if (diagonal >= 0 && ad < qLen) {
        unsigned m = std::min(tLen, qLen - ad); res.diagonalLen = (int) m;
        global_dist(q + ad, t, m, res);
    } else if (diagonal < 0 && ad < tLen) {
        unsigned m = std::min(tLen - ad, qLen); res.diagonalLen = (int) m;
        global_dist(q, t + ad, m, res);
    }
    return res;
}
static LocalAln ungapped(const char *q, unsigned qLen, const char *t, unsigned tLen, unsigned short diag) {   // :93-113
    LocalAln mx;
    for (unsigned d = 1; d <= 1 + tLen / 32768; d++) {
        LocalAln tmp = by_diagonal(q, qLen, t, tLen, (int) (-d * 65536 + diag));
        if (tmp.score > mx.score) mx = tmp;
    }
    for (unsigned d = 0; d <= qLen / 65536; d++) {
        LocalAln tmp = by_diagonal(q, qLen, t, tLen, (int) (d * 65536 + diag));
        if (tmp.score > mx.score) mx = tmp;
    }
    return mx;
}
// One call per prefilter line (self lines `key 0 0` included), in pref order.  Outputs per line.
extern "C" void orc_rescore(const char *blob, const u64 *off, const u32 *len, size_t n,
                            const u32 *qIdx, const u32 *tIdx, const int32_t *prefScore, const uint16_t *diag16, size_t M,
                            double dbRes, double evalThr, float seqIdThr,
                            uint8_t *keep, int32_t *score, int32_t *bits, double *evalue, float *seqId,
                            int32_t *qS, int32_t *qE, int32_t *dS, int32_t *dE, int32_t *idCntOut, int32_t *alnLenOut) {
    (void) n;
    std::string rev;
    for (size_t h = 0; h < M; h++) {
        const char *q = blob + off[qIdx[h]]; int qLen = (int) len[qIdx[h]];
        const char *t = blob + off[tIdx[h]]; int tLen = (int) len[tIdx[h]];
        bool isRev = prefScore[h] < 0;
        const char *qa = q;
        if (isRev) {   // rescorediagonal.cpp:173-179: num2aa[reverseResidue(aa2num[c])]
            rev.resize(qLen);
            for (int p = qLen - 1; p > -1; p--) rev[(qLen - 1) - p] = NUM2AA[rev_code(map_letter((unsigned char) q[p]))];
            qa = rev.data();
        }
        bool isIdentity = qIdx[h] == tIdx[h];
        LocalAln a = ungapped(qa, (unsigned) qLen, t, (unsigned) tLen, diag16[h]);
        int distance = (int) a.score;
        double ev, bs; orc_evalue(distance, qLen, dbRes, &ev, &bs);
        int bitScore = (int) (bs + 0.5);
        int alnLen = (a.endPos - a.startPos) + 1;
        int qs, qe, ds, de;
        if (a.diagonal >= 0) { qs = a.startPos + (int) a.dist; qe = a.endPos + (int) a.dist; ds = a.startPos; de = a.endPos; }
        else { qs = a.startPos; qe = a.endPos; ds = a.startPos + (int) a.dist; de = a.endPos + (int) a.dist; }
        double sid = 0; int idc = 0;
        if (ev <= evalThr || isIdentity) {
            for (int i = qs; i <= qe; i++) {
                char ql = qa[i] & (unsigned char) (~0x20), tl = t[ds + (i - qs)] & (unsigned char) (~0x20);
                idc += (ql == tl) ? 1 : 0;
            }
            sid = (float) idc / (float) alnLen;   // Util::computeSeqId SEQ_ID_ALN_LEN (Util.cpp:588-598)
        }
        if (isRev) { qs = qLen - qs - 1; qe = qLen - qe - 1; }
        bool hasSeqId = sid >= (seqIdThr - FLT_EPSILON);
        bool hasEvalue = ev <= evalThr;
        keep[h] = (isIdentity || (hasSeqId && hasEvalue)) ? 1 : 0;   // cov-mode 1 / -c 0 / min-aln-len 0 always true
        score[h] = distance; bits[h] = bitScore; evalue[h] = ev; seqId[h] = (float) sid;
        qS[h] = qs; qE[h] = qe; dS[h] = ds; dE[h] = de; idCntOut[h] = idc; alnLenOut[h] = alnLen;
    }
}

// ---------------------------------------------------------------------------------------------
// damage profiles (src/assembler/nuclassembleUtil.h:53-102, nuclassembleUtil.cpp:49-65,821-1007)
struct DiNuc { long double p[4][4]; };
static std::vector<DiNuc> g_D(11), g_Drev(11);
static std::vector<std::vector<long double>> read_prof(const char *fn) {
    std::vector<std::vector<long double>> rows;
    FILE *f = fopen(fn, "r");
    if (!f) { fprintf(stderr, "oracle: cannot open %s\n", fn); exit(1); }
    char line[4096]; bool header = true;
    while (fgets(line, sizeof line, f)) {
        if (header) { header = false; continue; }
        std::vector<long double> r; char *p = line;
        for (int i = 0; i < 12; i++) { char *e; r.push_back(strtold(p, &e)); p = e; }
        rows.push_back(r);
    }
    fclose(f);
    return rows;
}
extern "C" void orc_load_damage(const char *p5, const char *p3) {
    std::vector<std::vector<long double>> s5, s3;
    if (p5 == NULL || p5[0] == 0) { s5.assign(5, std::vector<long double>(12, 0.0L)); s3 = s5; }   // :824-832
    else { s5 = read_prof(p5); s3 = read_prof(p3); }
    DiNuc dflt;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) dflt.p[i][j] = (i == j) ? 1.0f : 0.0f;
    if (!s5.empty()) { dflt.p[1][3] = s5.back()[5]; dflt.p[1][1] = 1 - s5.back()[5]; }
    if (!s3.empty()) { dflt.p[2][0] = s3.front()[6]; dflt.p[2][2] = 1 - s3.front()[6]; }
    double dCC = (double) dflt.p[1][1], dCT = (double) dflt.p[1][3], dGA = (double) dflt.p[2][0], dGG = (double) dflt.p[2][2];
    std::vector<DiNuc> sub;
    for (int which = 0; which < 2; which++) {
        for (auto &r : (which == 0 ? s5 : s3)) {
            DiNuc m; int idx = 0;
            for (int i = 0; i < 4; i++) {
                double sum = 0.0;
                for (int j = 0; j < 4; j++) { if (i == j) continue; m.p[i][j] = r[idx]; sum += r[idx]; idx++; }
                m.p[i][i] = 1.0 - sum;
            }
            if (which == 0) { m.p[2][0] = dGA; m.p[2][2] = dGG; }   // defaultGA into every 5p matrix
            else { m.p[1][1] = dCC; m.p[1][3] = dCT; }              // defaultCT into every 3p matrix
            sub.push_back(m);
        }
    }
    for (int i = 0; i < 11; i++) g_D[i] = dflt;
    for (int i = 0; i < 5; i++) { g_D[i] = sub[i]; g_D[6 + i] = sub[sub.size() - 5 + i]; }
    g_Drev = g_D;
    for (int i = 0; i < 11; i++) {
        const DiNuc &e = g_D[10 - i];
        g_Drev[i].p[1][3] = e.p[2][0]; g_Drev[i].p[1][1] = e.p[2][2];
        g_Drev[i].p[2][0] = e.p[1][3]; g_Drev[i].p[2][2] = e.p[1][1];
    }
}
extern "C" void orc_get_damage(double *D, double *Drev) {   // [11][4][4] as doubles (inspection / KATs)
    for (int l = 0; l < 11; l++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
        D[(l * 4 + i) * 4 + j] = (double) g_D[l].p[i][j]; Drev[(l * 4 + i) * 4 + j] = (double) g_Drev[l].p[i][j];
    }
}

// ---------------------------------------------------------------------------------------------
// ancient_correction (src/assembler/correction.cpp).  Base index order here is A0 C1 G2 T3 (:170-174).
static int nuc_map(char c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }   // unordered_map::operator[] -> 0
static int ry_map(char c) { return (c == 'C' || c == 'T') ? 1 : 0; }
static const double SMOOTH = 0.001;

// mostLikeliBaseRead (correction.cpp:7-123); liks (optional) receives the four accumulated sums
static int most_likely_base(int b, unsigned qIter, const int cnt[4][11], const int rv[4][11], const DiNuc &seqErr,
                            bool wasCorr, unsigned qLen, long double *liks, int *evaluated) {
    unsigned cov[4] = {0, 0, 0, 0};
    for (int t = 0; t < 4; t++) for (int l = 0; l < 11; l++) cov[t] += (unsigned) cnt[t][l];
    double ctRatio = (double) cov[3] / (cov[1] + cov[3] + cov[0] + cov[2]);
    double gaRatio = (double) cov[0] / (cov[1] + cov[3] + cov[0] + cov[2]);
    double logQ[4], logT[4];
    *evaluated = 0;
    for (unsigned q = 0; q < 4; q++) {
        logT[q] = (double) logl(seqErr.p[q][b]);
        if (wasCorr) logQ[q] = (double) logl(seqErr.p[q][b]);
        else {
            if (ctRatio >= 0.4 || gaRatio >= 0.4) return b;
            const DiNuc *d;
            if (qIter < 5) d = &g_D[qIter];
            else if (qIter >= qLen - 5) d = &g_D[11 - (qLen - qIter)];
            else d = &g_D[5];
            double deam = (double) d->p[q][b];
            logQ[q] = std::log(std::max(deam, SMOOTH));
        }
    }
    long double lik[4];
    for (int q = 0; q < 4; q++) {
        long double acc = 0;
        for (int t = 0; t < 4; t++) {
            if (cov[t] == 0) continue;
            for (int l = 0; l < 11; l++) {
                double dp = std::max((double) g_D[l].p[q][t], SMOOTH), dr = std::max((double) g_Drev[l].p[q][t], SMOOTH);
                int c = cnt[t][l], nr = rv[t][l];
                double ldp = std::log(dp), ldr = std::log(dr);
                if (c != 0) {
                    acc += (c - nr) * (logT[t] + logQ[q] + ldp);
                    acc += nr * (logT[t] + logQ[q] + ldr);
                }
            }
        }
        lik[q] = acc;
    }
    *evaluated = 1;
    if (liks) for (int q = 0; q < 4; q++) liks[q] = lik[q];
    int best = 0;
    for (int q = 1; q < 4; q++) if (lik[q] > lik[best]) best = q;   // std::max_element: first maximum
    return best;
}

// Alignments come per query in CSR form (alnOff[n+1]) exactly as parsed from the aln DB
// (Matcher::parseAlignmentRecord, Matcher.cpp:274-353): raw qStart/qEnd (qStart > qEnd = reverse).
// outBlob has the same layout as blob.  likOut (optional): [totalBases][4] doubles, NaN where not evaluated.
extern "C" void orc_correct(const char *blob, const u64 *off, const u32 *len, const uint8_t *ext, size_t n,
                            const u64 *alnOff, const u32 *tIdx, const float *seqId,
                            const int32_t *qS_, const int32_t *qE_, const int32_t *dS_, const int32_t *dE_,
                            float seqIdThr, float ryThr, double seqErrRate,
                            char *outBlob, const u64 *likOff, double *likOut) {
    DiNuc seqErr; long double err = (long double) seqErrRate;
    // correction.cpp:196 `long double seqErrCorrection = 0.01;` is a double literal widened: same as the cast above
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) seqErr.p[i][j] = (i == j) ? 1 - err : err / 3;
    struct A { u32 t; float sid; int qS, qE, dS, dE, dL, aln; bool rev; };
    std::vector<A> alns; std::string trev;
    for (size_t qi = 0; qi < n; qi++) {
        const char *q = blob + off[qi]; unsigned qLen = len[qi];
        bool qExt = ext[qi] != 0;
        alns.clear();
        float avCov = 0;
        for (u64 a = alnOff[qi]; a < alnOff[qi + 1]; a++) {   // pass 1 :220-245
            A x; x.t = tIdx[a]; x.sid = seqId[a]; x.qS = qS_[a]; x.qE = qE_[a]; x.dS = dS_[a]; x.dE = dE_[a]; x.dL = (int) len[x.t];
            x.aln = std::max(abs(x.qE - x.qS), abs(x.dE - x.dS)) + 1;   // Matcher.cpp:319
            if (x.qS > x.qE) { std::swap(x.qS, x.qE); int ds = x.dS; x.dS = x.dL - x.dE - 1; x.dE = x.dL - ds - 1; x.rev = true; }
            else x.rev = false;
            avCov += x.aln;
            alns.push_back(x);
        }
        avCov = (float) avCov / qLen;
        std::vector<int> cnt(qLen * 44, 0), rv(qLen * 44, 0), tot(qLen, 0);
        for (const A &x : alns) {   // passes 2+3 :273-392 (per target the two passes see the same data)
            if (ext[x.t]) continue;
            const char *ts = blob + off[x.t];
            if (x.rev) {   // getNuclRevFragment (nuclassembleUtil.cpp:67-76)
                trev.resize(x.dL);
                for (int p = x.dL - 1; p > -1; p--) { char r = NUM2AA[rev_code(map_letter((unsigned char) ts[p]))]; trev[(x.dL - 1) - p] = (r == 'X') ? 'N' : r; }
                ts = trev.data();
            }
            unsigned dist = 0;
            // getRYSeqId :78-92.  Its result only matters for aln >= 30 (:359); score-0 alignments (start = end = -1) make the
            // reference read the byte in front of the sequence here, which the restatement does not imitate.
            if (x.aln >= 30)
            for (int p = 0; p < x.aln; p++) if (ry_map(q[x.qS + p]) != ry_map(ts[x.dS + p])) dist++;
            float ry = (float) ((unsigned) x.aln - dist) / (float) x.aln;
            float thr = ryThr;
            if (x.aln <= 100) { thr = ((float) x.aln - 1) / (float) x.aln; thr = std::floor(thr * 1000) / 1000; }
            bool sel = (ry >= thr && x.dS == 0 && (unsigned) x.qE == qLen - 1) ||
                       (ry >= thr && x.qS == 0 && x.dE == x.dL - 1) || (ry >= thr && avCov < 50);
            if (!sel) continue;
            if (!(ry >= thr && x.sid >= seqIdThr && x.aln >= 30)) continue;
            for (int p = 0; p < x.aln; p++) {
                int tp = x.dS + p, qp = x.qS + p;
                int cl = (tp >= x.dL - 5) ? 6 + (tp - (x.dL - 5)) : (tp < 5 ? tp : 5);   // :361-371 (last loop wins)
                int b = nuc_map(ts[tp]);
                cnt[qp * 44 + b * 11 + cl] += 1; tot[qp] += 1; rv[qp * 44 + b * 11 + cl] += x.rev ? 1 : 0;
            }
        }
        char *out = outBlob + off[qi];
        for (unsigned qp = 0; qp < qLen; qp++) {   // pass 4 :397-463
            if (likOut) for (int x = 0; x < 4; x++) likOut[(likOff[qi] + qp) * 4 + x] = NAN;
            if (tot[qp] <= 1) { out[qp] = q[qp]; continue; }
            int c4[4][11], r4[4][11];
            for (int t = 0; t < 4; t++) for (int l = 0; l < 11; l++) { c4[t][l] = cnt[qp * 44 + t * 11 + l]; r4[t][l] = rv[qp * 44 + t * 11 + l]; }
            long double liks[4]; int evaluated;
            int nb = most_likely_base(nuc_map(q[qp]), qp, c4, r4, seqErr, qExt, qLen, liks, &evaluated);
            if (likOut && evaluated) for (int x = 0; x < 4; x++) likOut[(likOff[qi] + qp) * 4 + x] = (double) liks[x];
            out[qp] = "ACGT"[nb];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// damage-aware pair likelihood: calcLikelihood (nuclassembleUtil.cpp:95-200),
// calcLikelihoodConsensus (:203-374), called through r_s_pair (ancientReadsResults.cpp:48-70, D or Drev by
// isRevToAlignment).  Pinned by tests/golden/pairlik_kat.txt (outputs of the reference's own functions).
static double column_lik(const DiNuc &tp, const DiNuc &seqErr, int qBase, int tBase) {   // :150-165
    double lik = 0;
    for (int x = 0; x < 4; x++) {
        double match = (double) std::max((long double) SMOOTH, tp.p[qBase][x]);
        long double tErr = seqErr.p[x][tBase];
        lik += (tErr * match);
    }
    return lik;
}
static const DiNuc &class_of(const std::vector<DiNuc> &tab, unsigned pos, unsigned dbLen) {   // subdeam_lookup :126-136
    if (pos >= dbLen - 5) return tab[6 + (pos - (dbLen - 5))];   // the last loop overwrites the earlier ones
    return tab[pos < 5 ? pos : 5];
}
static void oriented_target(const char *t, unsigned dbLen, bool rev, std::string &out) {
    out.assign(t, dbLen);
    if (rev) for (unsigned p = 0; p < dbLen; p++) { char r = NUM2AA[rev_code(map_letter((unsigned char) t[dbLen - 1 - p]))]; out[p] = (r == 'X') ? 'N' : r; }
}
static void finish_pair(long double likMod, unsigned maxAln, unsigned counted, float randAlnPenal, float excessPenal, double *sLen, double *sRatio) {
    unsigned excess = maxAln - counted;                // unsigned wrap kept (:178, :331)
    // `log(float)` resolves to the float overload there, so both products are FLOAT products (unsigned * float),
    // pinned by the known answers: (float) maxAln * logf(0.85f) reproduces sRatio to the last bit
    likMod += (float) excess * logf(excessPenal);
    double randAln = (float) maxAln * logf(randAlnPenal);
    *sLen = (double) likMod;
    *sRatio = 1.0 / (1.0 + exp(randAln - (double) likMod));
}
// coordinates in the oriented frame; `target` is the stored sequence (reverse-complemented here when rev)
extern "C" void orc_pair_loglik(const char *query, const char *target, unsigned dbLen, int rev, int qStart, int dbStart,
                                unsigned alnLen, unsigned maxAln, double seqErrRate, float randAlnPenal, float excessPenal,
                                double *sLen, double *sRatio) {
    DiNuc seqErr; long double err = (long double) seqErrRate;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) seqErr.p[i][j] = (i == j) ? 1 - err : err / 3;
    const std::vector<DiNuc> &tab = rev ? g_Drev : g_D;
    std::string t; oriented_target(target, dbLen, rev != 0, t);
    long double likMod = 0.0;
    for (unsigned c = 0; c < alnLen; c++)
        likMod += log(column_lik(class_of(tab, dbStart + c, dbLen), seqErr, nuc_map(query[qStart + c]), nuc_map(t[dbStart + c])));
    finish_pair(likMod, maxAln, alnLen, randAlnPenal, excessPenal, sLen, sRatio);
}
extern "C" void orc_pair_loglik_consensus(const char *consensus, unsigned qLen, const char *target, unsigned dbLen, int rev,
                                          int qStart, int qEnd, int dbStart, int dbEnd, unsigned alnLen, unsigned maxAln,
                                          double seqErrRate, float randAlnPenal, float excessPenal, double *sLen, double *sRatio) {
    DiNuc seqErr; long double err = (long double) seqErrRate;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) seqErr.p[i][j] = (i == j) ? 1 - err : err / 3;
    const std::vector<DiNuc> &tab = rev ? g_Drev : g_D;
    std::string t; oriented_target(target, dbLen, rev != 0, t);
    const bool rightStart = dbStart == 0 && (unsigned) qEnd == qLen - 1;
    const bool leftStart = qStart == 0 && (unsigned) dbEnd == dbLen - 1;
    long double likMod = 0.0; unsigned counted = 0;
    if (leftStart || rightStart) {
        unsigned pad = qLen - (dbLen - alnLen);   // consensusStart :246,:289
        // left: the target sits at consensus[pad ..]; right: it ends pad columns before the end of the 3*qLen consensus
        unsigned base = leftStart ? pad : 3 * qLen - (dbLen + pad);
        unsigned tIdx = 0;
        for (unsigned j = 0; j < dbLen; j++) {
            if (t[j] != 'N') tIdx++;
            char c = consensus[base + j];
            if (c == 'N' || t[j] == 'N') continue;
            counted++;
            likMod += log(column_lik(class_of(tab, tIdx - 1, dbLen), seqErr, nuc_map(c), nuc_map(t[j])));
        }
    }
    finish_pair(likMod, maxAln, counted, randAlnPenal, excessPenal, sLen, sRatio);
}

// ---- scoring front half of ancient_read_assemble ------------------------------------------------------------
// Per query, over the alignment lines of its aln DB entry (Matcher::readAlignmentResults: alnLength = max(|qE-qS|,|dbE-dbS|)+1):
//   orientation                       src/assembler/ancientReadsResults.cpp:200-236
//   identity / RY identity recompute  :247-293   (target index compared with the query KEY, as the reference does)
//   extension candidates (notContig)  :294-315
//   consensusCaller                   src/assembler/nuclassembleUtil.cpp:570-702 (+ calculateConsensus :533-568)
//   updateSeqIdConsensusReads         :377-470
//   final filter + r_s_pair           ancientReadsResults.cpp:330-365, :48-70 -> calcLikelihoodConsensus
// flags out: bit 0 candidate, bit 1 isRevToAlignment, bit 2 scored.  Coordinates out: the oriented frame.
static void call_consensus(std::string &consensus, const std::vector<std::vector<unsigned>> &cov, unsigned minCov) {   // :533-568
    for (size_t i = 0; i < cov.size(); i++) {
        unsigned total = cov[i][0] + cov[i][1] + cov[i][2] + cov[i][3];
        if (total >= minCov) {
            unsigned mx = 0; char nuc = 'N'; int nMax = 0;
            for (int j = 0; j < 4; j++) {
                if (cov[i][j] > mx) { mx = cov[i][j]; nuc = "ACGT"[j]; nMax = 1; }
                else if (cov[i][j] == mx && mx > 0) nMax++;
            }
            if (nMax > 1) nuc = 'N';
            consensus[i] = nuc;
        }
    }
}
extern "C" void orc_extension(const char *blob, const u64 *off, const u32 *len, const uint8_t *ext, const u32 *keys, size_t n, const u64 *lineOff,
                              const u32 *tgt, const int32_t *qS, const int32_t *qE, const int32_t *dS, const int32_t *dE, const float *seqIdIn,
                              float seqIdThr, float rySeqIdThr, int unsafe, unsigned minCovSafe, double seqErrRate, float randAlnPenal, float excessPenal,
                              uint8_t *flags, int32_t *oqS, int32_t *oqE, int32_t *odS, int32_t *odE, float *seqIdOut, float *ryOut, double *sLenOut,
                              double *sRatioOut, u32 *maxLeftOut, u32 *maxRightOut, char *consOut /*3 * sum(len), query q at 3 * (sum of len before q)*/) {
    u64 consOff = 0;
    for (size_t q = 0; q < n; q++) {
        const unsigned qLen = len[q];
        const char *query = blob + off[q];
        const u64 l0 = lineOff[q], l1 = lineOff[q + 1];
        struct Aln { u32 t; int qs, qe, ds, de; unsigned dbLen, alnLen; float seqId, ry; bool rev; };
        std::vector<Aln> a(l1 - l0);
        for (u64 l = l0; l < l1; l++) {
            Aln &x = a[l - l0];
            x.t = tgt[l]; x.qs = qS[l]; x.qe = qE[l]; x.ds = dS[l]; x.de = dE[l]; x.dbLen = len[x.t];
            x.alnLen = (unsigned) (std::max(std::abs(x.qe - x.qs), std::abs(x.de - x.ds)) + 1);
            x.seqId = seqIdIn[l]; x.ry = 0.f; x.rev = false;
        }
        auto ends = [&](const Aln &x, bool &right, bool &left) {
            right = (unsigned) x.ds == 0 && (unsigned) x.qe == qLen - 1;
            left = (unsigned) x.qs == 0 && (unsigned) x.de == x.dbLen - 1;
        };
        for (Aln &x : a) {   // orientation
            bool r, l; ends(x, r, l);
            if (!r && !l) continue;
            if (x.qs > x.qe) { std::swap(x.qs, x.qe); int d0 = x.ds; x.ds = (int) x.dbLen - x.de - 1; x.de = (int) x.dbLen - d0 - 1; x.rev = true; }
            else x.rev = false;
        }
        std::string tSeq;
        for (Aln &x : a) {   // identities against the query
            bool r, l; ends(x, r, l);
            if (!r && !l) continue;
            if (x.t == keys[q]) continue;   // (sic) target INDEX against query KEY
            oriented_target(blob + off[x.t], x.dbLen, x.rev, tSeq);
            int idCnt = 0, idRy = 0;
            for (int i = x.qs; i <= x.qe; i++) {
                char qc = query[i], tc = tSeq[x.ds + (i - x.qs)];
                idCnt += qc == tc; idRy += ry_map(qc) == ry_map(tc);
            }
            x.seqId = (float) idCnt / x.alnLen; x.ry = (float) idRy / x.alnLen;
        }
        std::vector<size_t> cand;
        for (size_t i = 0; i < a.size(); i++) {   // candidates
            bool r, l; ends(a[i], r, l);
            if (!r && !l) continue;
            bool noOffset = (a[i].dbLen - a[i].alnLen) == 0;
            if (!ext[a[i].t] && a[i].alnLen >= 30 && a[i].seqId >= seqIdThr && !noOffset) cand.push_back(i);
        }
        std::string consensus(3 * (size_t) qLen, 'N');
        if (!unsafe) { for (unsigned p = 0; p < qLen; p++) consensus[qLen + p] = query[p]; }
        else {
            std::vector<std::vector<unsigned>> cov(3 * (size_t) qLen, std::vector<unsigned>(4, 0));
            for (size_t ci : cand) {
                const Aln &x = a[ci];
                const bool rightStart = x.ds == 0 && x.de != (int) x.dbLen - 1;
                const bool leftStart = x.qs == 0 && x.qe != (int) qLen - 1;
                const bool notId = keys[x.t] != keys[q];
                if (!((rightStart || leftStart) && notId)) continue;
                oriented_target(blob + off[x.t], x.dbLen, x.rev, tSeq);
                if ((unsigned) x.ds == 0 && (unsigned) x.qe == qLen - 1) { for (unsigned p = 0; p < x.dbLen; p++) cov[qLen + x.qs + p][nuc_map(tSeq[p])]++; }
                else if ((unsigned) x.qs == 0 && (unsigned) x.de == x.dbLen - 1) { for (unsigned p = 0; p < x.dbLen; p++) cov[qLen - (x.dbLen - x.alnLen) + p][nuc_map(tSeq[p])]++; }
            }
            call_consensus(consensus, cov, minCovSafe);
            for (unsigned p = 0; p < qLen; p++) consensus[qLen + p] = query[p];
        }
        unsigned maxLeft = 0, maxRight = 0;
        for (size_t ci : cand) {   // updateSeqIdConsensusReads
            Aln &x = a[ci];
            oriented_target(blob + off[x.t], x.dbLen, x.rev, tSeq);
            const bool rightStart = (unsigned) x.ds == 0 && (unsigned) x.qe == qLen - 1;
            const bool leftStart = (unsigned) x.qs == 0 && (unsigned) x.de == x.dbLen - 1;
            int idCnt = 0, idRy = 0; unsigned tot = 0;
            const unsigned pad = qLen - (x.dbLen - x.alnLen);
            if (leftStart) {
                for (unsigned i = 0; i < x.dbLen; i++) {   // the padded target starts at consensus[pad]
                    char c = consensus[pad + i], t = tSeq[i];
                    if (!(c == 'N' || t == 'N')) { idCnt += c == t; idRy += ry_map(c) == ry_map(t); tot++; }
                }
            } else if (rightStart) {
                const size_t base = consensus.size() - (x.dbLen + pad);
                for (unsigned i = 0; i < x.dbLen; i++) {
                    char c = consensus[base + i], t = tSeq[i];
                    if (!(c == 'N' || t == 'N')) { idCnt += c == t; idRy += ry_map(c) == ry_map(t); tot++; }
                }
            }
            if (tot != 0) { x.seqId = (float) idCnt / tot; x.ry = (float) idRy / tot; }
            if (leftStart && tot > maxLeft) maxLeft = tot;
            else if (rightStart && tot > maxRight) maxRight = tot;
        }
        maxLeftOut[q] = maxLeft; maxRightOut[q] = maxRight;
        memcpy(consOut + consOff, consensus.data(), consensus.size());
        for (u64 l = l0; l < l1; l++) {
            const Aln &x = a[l - l0];
            flags[l] = x.rev ? 2 : 0; oqS[l] = x.qs; oqE[l] = x.qe; odS[l] = x.ds; odE[l] = x.de; seqIdOut[l] = x.seqId; ryOut[l] = x.ry; sLenOut[l] = 0; sRatioOut[l] = 0;
        }
        for (size_t ci : cand) {   // scores
            const Aln &x = a[ci];
            const u64 l = l0 + ci;
            flags[l] |= 1;
            const bool notInside = x.dbLen != x.alnLen, rightStart = x.ds == 0, leftStart = x.qs == 0, notId = keys[x.t] != keys[q];
            if ((rightStart || leftStart) && notInside && notId && x.ry >= rySeqIdThr && x.seqId >= seqIdThr) {
                unsigned maxOverlap = maxRight;
                if ((unsigned) x.qs == 0 && (unsigned) x.de == x.dbLen - 1) maxOverlap = maxLeft;
                orc_pair_loglik_consensus(consensus.data(), qLen, blob + off[x.t], x.dbLen, x.rev ? 1 : 0, x.qs, x.qe, x.ds, x.de, x.alnLen, maxOverlap, seqErrRate,
                                          randAlnPenal, excessPenal, &sLenOut[l], &sRatioOut[l]);
                flags[l] |= 4;
            }
        }
        consOff += 3 * (u64) qLen;
    }
}

// ---- scoring front half of ancient_contig_merge --------------------------------------------------------------
// Per query, over the alignment lines of its aln DB entry (doNuclAssembly2, src/assembler/ancientContigsResults.cpp):
//   orientation + identity / RY identity against the query + merge candidates ("contigs")   :187-236
//   consensusCaller                   src/assembler/nuclassembleUtil.cpp:570-702 (the same function N1 uses)
//   updateSeqIdConsensus              :704-790   (identities over the non-N columns, alnLengthCons)
//   queue filter                      ancientContigsResults.cpp:246-262 (minAlnLen = 500, or min(500, 0.2*dbLen) for shorter overlaps)
//   ancientMatchCount / deamMatches   nuclassembleUtil.cpp:1050-1180, :1011-1047 (D or Drev by isRevToAlignment, interior class 5)
//   queue order                       CompareNuclResultByScoreContigs, ancientContigsResults.cpp:24-70 (std::priority_queue)
// Pinned by tests/golden/merge_kat_*.txt.gz (outputs of the reference's own functions, oracle/ref_merge.cpp).
// flags out: bit 0 merge candidate, bit 1 isRevToAlignment, bit 2 queued.  seqId0/ry0 = identities against the query.
static double deam_posterior(unsigned alnLength, unsigned scoreAln, double matchLik) {   // deamMatches :1011-1047
    const double logAdj = std::log(1.4e-9);
    const unsigned maxLength = 100000;
    auto logPower = [logAdj](unsigned length) { return logAdj - 3.0 * std::log(length); };
    const double logMin = logPower(10), logMax = logPower(maxLength);
    const double logLength = logPower(std::min(alnLength, maxLength));
    const double fractionLength = (std::abs(logLength) - std::abs(logMax)) / (std::abs(logMin) - std::abs(logMax));
    const double priorAln = 1 - fractionLength;
    const double pMatch = 0.5f * ((((static_cast<double>(scoreAln) + 3.0f * alnLength) / 5.0f) + 0.9f) / (alnLength + 1)) + 0.5f * priorAln;
    const double likNoMatch = 1 - pMatch;
    const double oddsRatio = likNoMatch / matchLik;
    const double odds = (1 - pMatch) / pMatch;
    return 1 / (1 + oddsRatio * odds);
}
extern "C" int orc_merge_less(unsigned alnLenCons1, float deam1, unsigned alnLenCons2, float deam2) {   // ancientContigsResults.cpp:24-70
    float mm1 = alnLenCons1 - deam1, mm2 = alnLenCons2 - deam2;
    float alpha1 = mm1 + 1, alpha2 = mm2 + 1, beta1 = deam1 + 1, beta2 = deam2 + 1;
    double log_c = (std::lgamma(beta1 + beta2) + std::lgamma(alpha1 + beta1)) - (std::lgamma(alpha1 + beta1 + beta2) + std::lgamma(beta1));
    double log_r = 0.0, p = 0.0;
    for (size_t idx = 0; idx < alpha2; idx++) {
        p += exp(log_r + log_c);
        log_r = log(alpha1 + idx) + log(beta2 + idx) - (log(idx + 1) + log(idx + alpha1 + beta1 + beta2)) + log_r;
    }
    if (p < 0.45) return 1;
    if (p > 0.55) return 0;
    if (alnLenCons1 < alnLenCons2) return 1;
    if (alnLenCons1 > alnLenCons2) return 0;
    return 1;
}
extern "C" void orc_merge(const char *blob, const u64 *off, const u32 *len, const u32 *keys, size_t n, const u64 *lineOff, const u32 *tgt,
                          const int32_t *qS, const int32_t *qE, const int32_t *dS, const int32_t *dE, float mergeSeqIdThr, float rySeqIdThr, int unsafe,
                          unsigned minCovSafe, uint8_t *flags, int32_t *oqS, int32_t *oqE, int32_t *odS, int32_t *odE, float *seqId0, float *ry0,
                          float *seqIdOut, float *ryOut, u32 *alnLenConsOut, float *deamOut, u64 *consHash /*per query, FNV-1a 64*/) {
    for (size_t q = 0; q < n; q++) {
        const unsigned qLen = len[q];
        const char *query = blob + off[q];
        const u64 l0 = lineOff[q], l1 = lineOff[q + 1];
        struct Aln { u32 t; int qs, qe, ds, de; unsigned dbLen, alnLen, alnLenCons; float seqId, ry; bool rev; };
        std::vector<Aln> a(l1 - l0);
        std::vector<size_t> cand;
        std::string tSeq;
        for (u64 l = l0; l < l1; l++) {
            Aln &x = a[l - l0];
            x.t = tgt[l]; x.qs = qS[l]; x.qe = qE[l]; x.ds = dS[l]; x.de = dE[l]; x.dbLen = len[x.t]; x.alnLenCons = 0;
            x.alnLen = (unsigned) (std::max(std::abs(x.qe - x.qs), std::abs(x.de - x.ds)) + 1);
            x.rev = x.qs > x.qe;
            if (x.rev) { std::swap(x.qs, x.qe); int d0 = x.ds; x.ds = (int) x.dbLen - x.de - 1; x.de = (int) x.dbLen - d0 - 1; }
            oriented_target(blob + off[x.t], x.dbLen, x.rev, tSeq);
            int idCnt = 0, idRy = 0;
            for (int i = x.qs; i <= x.qe; i++) {
                char qc = query[i], tc = tSeq[x.ds + (i - x.qs)];
                idCnt += qc == tc; idRy += ry_map(qc) == ry_map(tc);
            }
            x.seqId = (float) idCnt / x.alnLen; x.ry = (float) idRy / x.alnLen;
            seqId0[l] = x.seqId; ry0[l] = x.ry;
            if (x.seqId >= mergeSeqIdThr && x.ry >= rySeqIdThr && keys[q] != keys[x.t]) cand.push_back(l - l0);
        }
        std::string consensus(3 * (size_t) qLen, 'N');
        if (!unsafe) { for (unsigned p = 0; p < qLen; p++) consensus[qLen + p] = query[p]; }
        else {
            std::vector<std::vector<unsigned>> cov(3 * (size_t) qLen, std::vector<unsigned>(4, 0));
            for (size_t ci : cand) {
                const Aln &x = a[ci];
                const bool rightStart = x.ds == 0 && x.de != (int) x.dbLen - 1;
                const bool leftStart = x.qs == 0 && x.qe != (int) qLen - 1;
                if (!(rightStart || leftStart)) continue;
                oriented_target(blob + off[x.t], x.dbLen, x.rev, tSeq);
                if ((unsigned) x.ds == 0 && (unsigned) x.qe == qLen - 1) { for (unsigned p = 0; p < x.dbLen; p++) cov[qLen + x.qs + p][nuc_map(tSeq[p])]++; }
                else if ((unsigned) x.qs == 0 && (unsigned) x.de == x.dbLen - 1) { for (unsigned p = 0; p < x.dbLen; p++) cov[qLen - (x.dbLen - x.alnLen) + p][nuc_map(tSeq[p])]++; }
            }
            call_consensus(consensus, cov, minCovSafe);
            for (unsigned p = 0; p < qLen; p++) consensus[qLen + p] = query[p];
        }
        u64 h = 1469598103934665603ull;
        for (unsigned char c : consensus) { h ^= c; h *= 1099511628211ull; }
        consHash[q] = h;
        for (u64 l = l0; l < l1; l++) {
            const Aln &x = a[l - l0];
            flags[l] = x.rev ? 2 : 0; oqS[l] = x.qs; oqE[l] = x.qe; odS[l] = x.ds; odE[l] = x.de; alnLenConsOut[l] = 0; deamOut[l] = 0.f;
        }
        for (size_t ci : cand) {
            Aln &x = a[ci];
            const u64 l = l0 + ci;
            flags[l] |= 1;
            oriented_target(blob + off[x.t], x.dbLen, x.rev, tSeq);
            const bool rightStart = (unsigned) x.ds == 0 && (unsigned) x.qe == qLen - 1;
            const bool leftStart = (unsigned) x.qs == 0 && (unsigned) x.de == x.dbLen - 1;
            const unsigned pad = qLen - (x.dbLen - x.alnLen);
            // the padded target lies at consensus[base ..): left = pad, right = 3*qLen - (dbLen + pad)
            const size_t base = leftStart ? pad : consensus.size() - (x.dbLen + pad);
            int idCnt = 0, idRy = 0; unsigned tot = 0, nCT = 0, nGA = 0;
            if (leftStart || rightStart) {
                for (unsigned i = 0; i < x.dbLen; i++) {
                    char c = consensus[base + i], t = tSeq[i];
                    if (!(c == 'N' || t == 'N')) {
                        idCnt += c == t; idRy += ry_map(c) == ry_map(t); tot++;
                        const int dimer = 4 * nuc_map(c) + nuc_map(t);
                        nCT += dimer == 7; nGA += dimer == 8;
                    }
                }
            }
            if (tot != 0) { x.seqId = (float) idCnt / tot; x.ry = (float) idRy / tot; }   // updateSeqIdConsensus
            x.alnLenCons = tot;
            alnLenConsOut[l] = tot;
            unsigned minAlnLen = 500;
            minAlnLen = (x.alnLen < minAlnLen) ? std::min(minAlnLen, static_cast<unsigned>(0.2 * x.dbLen)) : minAlnLen;
            if (x.seqId >= mergeSeqIdThr && x.ry >= rySeqIdThr && x.alnLen >= minAlnLen) {   // ancientMatchCount
                const std::vector<DiNuc> &tab = x.rev ? g_Drev : g_D;
                float mCT = 0, mGA = 0;
                unsigned mmCons = (1 - x.seqId) * x.alnLenCons + 0.5;
                unsigned mCons = x.alnLenCons - mmCons;
                unsigned scoreAln = mCons * 2 + mmCons * (-3);
                const double likCT = tab[5].p[1][3], likGA = tab[5].p[2][0];
                if (likCT > 0) for (unsigned i = 0; i < nCT; i++) mCT += deam_posterior(x.alnLen, scoreAln, likCT);
                if (likGA > 0) for (unsigned i = 0; i < nGA; i++) mGA += deam_posterior(x.alnLen, scoreAln, likGA);
                deamOut[l] = ((static_cast<float>(scoreAln) + 3.0f * x.alnLenCons) / 5.0f) + mCT + mGA;
                flags[l] |= 4;
            }
        }
        for (u64 l = l0; l < l1; l++) { seqIdOut[l] = a[l - l0].seqId; ryOut[l] = a[l - l0].ry; }
    }
}
// the order in which std::priority_queue<…, CompareNuclResultByScoreContigs> hands out the queued lines (pushed in line order)
extern "C" void orc_merge_queue_order(const u32 *alnLenCons, const float *deam, size_t m, u32 *order) {
    struct Item { u32 i, alc; float d; };
    struct Less { bool operator()(const Item &a, const Item &b) const { return orc_merge_less(a.alc, a.d, b.alc, b.d) != 0; } };
    std::priority_queue<Item, std::vector<Item>, Less> pq;
    for (size_t i = 0; i < m; i++) pq.push(Item{(u32) i, alnLenCons[i], deam[i]});
    size_t k = 0;
    while (!pq.empty()) { order[k++] = pq.top().i; pq.pop(); }
}

// ---- rescorediagonal --rescore-mode 0 (Hamming) with and without --wrapped-scoring -------------------------------
// linclust's pre-clustering step (lib/mmseqs/data/workflow/linclust.sh:28-32, HAMMING_PAR of lib/mmseqs/src/workflow/Linclust.cpp:103-112).
//   per-line body                        lib/mmseqs/src/alignment/rescorediagonal.cpp:194-334 (HAMMING branch :251-254, output :323-329)
//   computeUngappedAlignment             lib/mmseqs/src/alignment/DistanceCalculator.h:93-113, ungappedAlignmentByDiagonal :115-175
//   computeUngappedWrappedAlignment      :57-91 (query doubled, every start of the circular query against the whole target)
//   computeInverseHammingDistance        :275-295 (number of equal bytes)
//   Util::canBeCovered / hasCoverage / computeSeqId   lib/mmseqs/src/commons/Util.cpp:533-567, :588-598
// Pinned by tests/golden/cfg1_lin_rescore_9.gz / cfg1_lin_rescore_wrapped_9.gz (outputs of the reference binary, oracle/make_hamming_golden.py).
static unsigned inv_hamming(const char *a, const char *b, unsigned n) { unsigned s = 0; for (unsigned i = 0; i < n; i++) s += a[i] == b[i]; return s; }
static LocalAln ham_by_diagonal(const char *q, unsigned qLen, const char *t, unsigned tLen, int diagonal) {
    LocalAln res; unsigned ad = (unsigned) abs(diagonal);
    res.dist = ad; res.diagonal = diagonal;
    if (diagonal >= 0 && ad < qLen) { unsigned m = std::min(tLen, qLen - ad); res.diagonalLen = (int) m; res.score = inv_hamming(q + ad, t, m); }
    else if (diagonal < 0 && ad < tLen) { unsigned m = std::min(tLen - ad, qLen); res.diagonalLen = (int) m; res.score = inv_hamming(q, t + ad, m); }
    return res;
}
static bool can_be_covered(float covThr, int covMode, float q, float t) {
    switch (covMode) {
        case 0: return (q / t >= covThr) && (t / q >= covThr);
        case 1: return (q / t) >= covThr;    // COV_MODE_TARGET
        case 2: return (t / q) >= covThr;    // COV_MODE_QUERY
        case 3: return ((t / q) >= covThr) && (t / q) <= 1.0;
        case 4: return ((q / t) >= covThr) && (q / t) <= 1.0;
        case 5: return (std::min(t, q) / std::max(t, q)) >= covThr;
        default: return true;
    }
}
static bool has_coverage(float covThr, int covMode, float qCov, float tCov) {
    switch (covMode) {
        case 0: return qCov >= covThr && tCov >= covThr;
        case 2: return qCov >= covThr;
        case 1: return tCov >= covThr;
        default: return true;
    }
}
// per prefilter line: keep (0 also for lines the reference skips), prefScore and diagonal as written (diagonal through unsigned short -> short)
extern "C" void orc_rescore_hamming(const char *blob, const u64 *off, const u32 *len, const u32 *qIdx, const u32 *tIdx, const int32_t *prefScore,
                                    const uint16_t *diag16, size_t M, int wrapped, float covThr, int covMode, float seqIdThr, int seqIdMode, int alnLenThr,
                                    uint8_t *keep, int32_t *scoreOut, int32_t *diagOut, float *seqIdOut) {
    std::string rev, dbl;
    for (size_t h = 0; h < M; h++) {
        const char *q = blob + off[qIdx[h]]; const int origQLen = (int) len[qIdx[h]];
        const char *t = blob + off[tIdx[h]]; const int dbLen = (int) len[tIdx[h]];
        keep[h] = 0; scoreOut[h] = 0; diagOut[h] = 0; seqIdOut[h] = 0.f;
        int qLen = origQLen;
        if (wrapped) { dbl.assign(q, origQLen); dbl += dbl; q = dbl.data(); qLen = 2 * origQLen; }
        const bool isRev = prefScore[h] < 0;
        if (isRev) {
            rev.resize(qLen);
            for (int p = qLen - 1; p > -1; p--) rev[(qLen - 1) - p] = NUM2AA[rev_code(map_letter((unsigned char) q[p]))];
            q = rev.data();
        }
        const bool isIdentity = qIdx[h] == tIdx[h];
        if (!can_be_covered(covThr, covMode, (float) origQLen, (float) dbLen)) continue;
        LocalAln a;
        if (wrapped) {
            if (dbLen > origQLen) continue;
            const unsigned half = (unsigned) qLen / 2, dbSeqLen = (unsigned) dbLen;
            for (unsigned d = 1; (-d * 65536 + diag16[h]) > -dbSeqLen; d++) {
                int realDiagonal = (int) (-d * 65536 + diag16[h]) + (int) half;
                LocalAln tmp = ham_by_diagonal(q + realDiagonal, half, t, dbSeqLen, 0);
                tmp.diagonal += realDiagonal; tmp.dist = (unsigned) abs(realDiagonal);
                if (tmp.score > a.score) a = tmp;
            }
            for (unsigned d = 0; (d * 65536 + diag16[h]) < half; d++) {
                int realDiagonal = (int) (d * 65536 + diag16[h]);
                LocalAln tmp = ham_by_diagonal(q + realDiagonal, half, t, dbSeqLen, 0);
                tmp.diagonal += realDiagonal; tmp.dist = (unsigned) abs(realDiagonal);
                if (tmp.score > a.score) a = tmp;
            }
            a.diagonalLen = (int) std::min(dbSeqLen, half);
        } else {
            for (unsigned d = 1; d <= 1 + (unsigned) dbLen / 32768; d++) {
                LocalAln tmp = ham_by_diagonal(q, (unsigned) qLen, t, (unsigned) dbLen, (int) (-d * 65536 + diag16[h]));
                if (tmp.score > a.score) a = tmp;
            }
            for (unsigned d = 0; d <= (unsigned) qLen / 65536; d++) {
                LocalAln tmp = ham_by_diagonal(q, (unsigned) qLen, t, (unsigned) dbLen, (int) (d * 65536 + diag16[h]));
                if (tmp.score > a.score) a = tmp;
            }
        }
        const int diagonalLen = a.diagonalLen, distance = (int) a.score;
        const float targetCov = (float) diagonalLen / (float) dbLen, queryCov = (float) diagonalLen / (float) origQLen;
        const int idCnt = (int) (float) distance;
        float sidf;   // Util::computeSeqId
        if (seqIdMode == 1) sidf = (float) idCnt / (float) std::min(origQLen, dbLen);
        else if (seqIdMode == 2) sidf = (float) idCnt / (float) std::max(origQLen, dbLen);
        else sidf = (float) idCnt / (float) diagonalLen;
        const double seqId = sidf;
        const int alnLen = diagonalLen;
        const bool hasCov = has_coverage(covThr, covMode, queryCov, targetCov);
        const bool hasSeqId = seqId >= (seqIdThr - FLT_EPSILON);
        const bool hasAlnLen = alnLen >= alnLenThr;
        if (isIdentity || (hasAlnLen && hasCov && hasSeqId)) {
            int ps = (int) (100 * seqId);
            keep[h] = 1; scoreOut[h] = isRev ? -ps : ps; diagOut[h] = (int) (short) (unsigned short) a.diagonal; seqIdOut[h] = sidf;
        }
    }
}

// ---- cyclecheck (src/assembler/cyclecheck.cpp:31-269) ---------------------------------------------------------------
// Per sequence: k-mers (k = 22, forward only, index = sum code_i * 4^i over the mapped letters, Indexer.h:21-84 — an X adds 4 * 4^i, so
// windows with X are NOT skipped) in three parts by position, matches front x back, front x middle, middle x back of equal k-mers (first
// occurrence of the k-mer in the earlier part against every occurrence in the later one) counted per diagonal >= L/3, and the first diagonal
// whose 1 % band holds more than 0.24 hits per possible k-mer is where the sequence wraps around.
// The part of a k-mer is decided with the iterator position BEFORE nextKmer() (:107-108: position - 1, UINT_MAX for the first k-mer, which
// therefore lands in the back part), the stored position is the real one.
// Pinned by tests/golden/cycle_kat.txt (chopped lengths written by the reference binary with --chop-cycle 1, oracle/make_cycle_golden.py).
extern "C" void orc_cyclecheck(const char *blob, const u64 *off, const u32 *len, size_t n, int k, unsigned maxSeqLen, u32 *splitOut) {
    struct KP { u64 kmer; unsigned pos; };
    auto cmp = [](const KP &a, const KP &b) { return a.kmer != b.kmer ? a.kmer < b.kmer : a.pos < b.pos; };
    std::vector<KP> front, middle, back;
    std::vector<unsigned> diagHits;
    for (size_t s = 0; s < n; s++) {
        splitOut[s] = 0;
        const unsigned seqLen = len[s];
        if (seqLen >= maxSeqLen) continue;
        const char *seq = blob + off[s];
        front.clear(); middle.clear(); back.clear();
        const unsigned third = seqLen / 3;
        for (unsigned p = 0; p + (unsigned) k <= seqLen; p++) {
            u64 idx = 0, pw = 1;
            for (int i = 0; i < k; i++) { idx += (u64) map_letter((unsigned char) seq[p + i]) * pw; pw *= 4; }
            const unsigned itPos = p - 1;   // getCurrentPosition() before nextKmer()
            if (itPos < third + 1) front.push_back(KP{idx, p});
            else if (itPos < 2 * third + 1) middle.push_back(KP{idx, p});
            else back.push_back(KP{idx, p});
        }
        std::sort(front.begin(), front.end(), cmp); std::sort(middle.begin(), middle.end(), cmp); std::sort(back.begin(), back.end(), cmp);
        unsigned kmermatches = 0;
        diagHits.assign(2 * (size_t) third + 1, 0);
        size_t idx = 0, jdx = 0, kdx = 0;
        while (idx < front.size() && (jdx < back.size() || kdx < middle.size())) {
            const u64 km = front[idx].kmer; const unsigned pos = front[idx].pos;
            while (jdx < back.size() && back[jdx].kmer < km) jdx++;
            while (kdx < middle.size() && middle[kdx].kmer < km) kdx++;
            while (jdx < back.size() && km == back[jdx].kmer) {
                int diag = (int) (back[jdx].pos - pos);
                if (diag >= (int) (seqLen / 3)) { diagHits[diag - seqLen / 3]++; kmermatches++; }
                jdx++;
            }
            while (kdx < middle.size() && km == middle[kdx].kmer) {
                int diag = (int) (middle[kdx].pos - pos);
                if (diag >= (int) (seqLen / 3)) { diagHits[diag - seqLen / 3]++; kmermatches++; }
                kdx++;
            }
            idx++;
            while (idx < front.size() && km == front[idx].kmer) idx++;
        }
        jdx = 0; kdx = 0;
        while (kdx < middle.size() && jdx < back.size()) {
            if (middle[kdx].kmer < back[jdx].kmer) kdx++;
            else if (middle[kdx].kmer > back[jdx].kmer) jdx++;
            else {
                const u64 km = middle[kdx].kmer; const unsigned pos = middle[kdx].pos;
                while (jdx < back.size() && km == back[jdx].kmer) {
                    int diag = (int) (back[jdx].pos - pos);
                    if (diag >= (int) (seqLen / 3)) { diagHits[diag - seqLen / 3]++; kmermatches++; }
                    jdx++;
                }
                while (kdx < middle.size() && km == middle[kdx].kmer) kdx++;
            }
        }
        unsigned split = 0;
        if (kmermatches > 0) {
            for (unsigned d = 0; d < 2 * third; d++) {
                if (diagHits[d] == 0) continue;
                const unsigned diag = d + third, diaglen = seqLen - diag;
                const unsigned gapwindow = diaglen * 0.01;
                const unsigned lower = std::max(0, static_cast<int>(d - gapwindow));
                const unsigned upper = std::min(d + gapwindow, 2 * third);
                unsigned band = 0;
                for (size_t i = lower; i <= upper; i++) if (diagHits[i] <= diagHits[d]) band += diagHits[i];
                const float rate = static_cast<float>(band) / (diaglen - (size_t) k + 1);
                if (rate > 0.24) { split = diag; break; }
            }
        }
        splitOut[s] = split;
    }
}

// Host side of the damage model: `--ancient-damage PFX` -> cdg_damage log tables for the device.
//
// Mirrors what every OpenMP thread of the reference builds at the top of doCorrection
// (src/assembler/correction.cpp:184-197): readNucSubstitionRatesFreq (nuclassembleUtil.h:53-102),
// initDeamProbabilities (nuclassembleUtil.cpp:821-1007), getSeqErrorProf (:49-65), and the logs
// mostLikeliBaseRead takes per column (correction.cpp:51-75, 93-107).  The numeric types follow the
// reference exactly, because they decide the last bit of every table entry: rates are long double,
// row sums are accumulated into a double, the defaults pass through double, every matrix entry is
// narrowed to double before max(.,1e-3) and log(double); the sequencing-error log is taken in long double.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/carpedeam_gpu.h"

namespace {

typedef long double ld;
struct Matrix4 { ld p[4][4]; };   // [from][to], base order A C G T
struct Rates { ld s[12]; };       // A>C A>G A>T C>A C>G C>T G>A G>C G>T T>A T>C T>G

bool read_profile(const std::string &path, std::vector<Rates> &rows, std::string &err) {
    FILE *f = fopen(path.c_str(), "r");
    if (!f) { err = "cannot open " + path; return false; }
    char line[8192];
    bool first = true;
    while (fgets(line, sizeof line, f)) {
        std::vector<std::string> fields;
        std::string cur;
        for (char *c = line; *c && *c != '\n' && *c != '\r'; c++) {
            if (*c == '\t') { fields.push_back(cur); cur.clear(); } else cur.push_back(*c);
        }
        fields.push_back(cur);
        if (fields.size() != 12) { fclose(f); err = "Profile not 12 fields: " + path; return false; }   // nuclassembleUtil.h:70,82
        if (first) { first = false; continue; }   // header line
        Rates r;
        for (int i = 0; i < 12; i++) r.s[i] = strtold(fields[i].c_str(), nullptr);
        rows.push_back(r);
    }
    fclose(f);
    if (first) { err = "empty profile " + path; return false; }
    return true;
}

// one position-specific matrix from a profile row; diagonal = 1 - (double) row sum
Matrix4 matrix_from_rates(const Rates &r) {
    Matrix4 m;
    int idx = 0;
    for (int from = 0; from < 4; from++) {
        double sum = 0.0;
        for (int to = 0; to < 4; to++) {
            if (to == from) continue;
            m.p[from][to] = r.s[idx];
            sum += r.s[idx];
            idx++;
        }
        m.p[from][from] = 1.0 - sum;
    }
    return m;
}

}   // namespace

// D[11] / Drev[11] of initDeamProbabilities (nuclassembleUtil.cpp:821-1007)
static int build_matrices(const char *prefix, Matrix4 *D, Matrix4 *Drev, char *errbuf, size_t errlen) {
    std::string pfx = prefix ? prefix : "";
    std::vector<Rates> p5, p3;
    std::string err;
    if (pfx.empty()) {   // "5p.prof"/"3p.prof": no damage (nuclassembleUtil.cpp:824-832)
        Rates z; for (int i = 0; i < 12; i++) z.s[i] = 0.0;
        p5.assign(5, z); p3.assign(5, z);
    } else if (!read_profile(pfx + "5p.prof", p5, err) || !read_profile(pfx + "3p.prof", p3, err)) {
        if (errbuf && errlen) snprintf(errbuf, errlen, "%s", err.c_str());
        return -1;
    }
    if (p5.size() + p3.size() < 5) {
        if (errbuf && errlen) snprintf(errbuf, errlen, "damage profiles have fewer than 5 rows");
        return -1;
    }
    // interior class: identity, except C>T := last 5' row, G>A := first 3' row (:847-858)
    Matrix4 interior;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) interior.p[i][j] = (i == j) ? 1.0f : 0.0f;
    if (!p5.empty()) { interior.p[1][3] = p5.back().s[5]; interior.p[1][1] = 1 - p5.back().s[5]; }
    if (!p3.empty()) { interior.p[2][0] = p3.front().s[6]; interior.p[2][2] = 1 - p3.front().s[6]; }
    const double defCC = (double) interior.p[1][1], defCT = (double) interior.p[1][3];
    const double defGA = (double) interior.p[2][0], defGG = (double) interior.p[2][2];
    std::vector<Matrix4> all;
    for (const Rates &r : p5) { Matrix4 m = matrix_from_rates(r); m.p[2][0] = defGA; m.p[2][2] = defGG; all.push_back(m); }
    for (const Rates &r : p3) { Matrix4 m = matrix_from_rates(r); m.p[1][1] = defCC; m.p[1][3] = defCT; all.push_back(m); }
    for (int l = 0; l < 11; l++) D[l] = interior;
    for (int l = 0; l < 5; l++) { D[l] = all[l]; D[6 + l] = all[all.size() - 5 + l]; }
    for (int l = 0; l < 11; l++) {   // :995-1005
        Drev[l] = D[l];
        const Matrix4 &e = D[10 - l];
        Drev[l].p[1][3] = e.p[2][0]; Drev[l].p[1][1] = e.p[2][2];
        Drev[l].p[2][0] = e.p[1][3]; Drev[l].p[2][2] = e.p[1][1];
    }
    return 0;
}

static const double SMOOTHING_VALUE = 0.001;   // correction.cpp:2, nuclassembleUtil.cpp:2

extern "C" int cdh_build_damage(const char *prefix, double seq_err, cdg_damage *out, char *errbuf, size_t errlen) {
    Matrix4 D[11], Drev[11];
    if (build_matrices(prefix, D, Drev, errbuf, errlen)) return -1;
    for (int l = 0; l < 11; l++)
        for (int q = 0; q < 4; q++)
            for (int t = 0; t < 4; t++) {
                out->log_d[l][q][t] = std::log(std::max((double) D[l].p[q][t], SMOOTHING_VALUE));
                out->log_drev[l][q][t] = std::log(std::max((double) Drev[l].p[q][t], SMOOTHING_VALUE));
            }
    ld e = (ld) seq_err;   // `long double seqErrCorrection = 0.01;` (correction.cpp:196): a double literal widened
    for (int q = 0; q < 4; q++)
        for (int b = 0; b < 4; b++) {
            ld p = (q == b) ? 1 - e : e / 3;
            out->log_seq_err[q][b] = (double) logl(p);
        }
    return 0;
}

// Pair-likelihood model: the per-column term of calcLikelihood / calcLikelihoodConsensus
// (nuclassembleUtil.cpp:150-168, :262-283) tabulated over (strand, class, qBase, tBase) in the reference's arithmetic:
// match = (double) max((long double) 1e-3, p[q][x]); lik (double) += seqErr.p[x][t] (long double) * match; log(lik) in double.
// seq_err = 0.001 in ancient_read_assemble (ancientReadsResults.cpp:172); the two penalties enter as logf(float).
extern "C" int cdh_build_pair_model(const char *prefix, double seq_err, float rand_align_penal, float excess_penal, cdg_pair_model *out,
                                    char *errbuf, size_t errlen) {
    Matrix4 D[11], Drev[11];
    if (build_matrices(prefix, D, Drev, errbuf, errlen)) return -1;
    ld e = (ld) seq_err;
    Matrix4 se;
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) se.p[a][b] = (a == b) ? 1 - e : e / 3;
    for (int rev = 0; rev < 2; rev++)
        for (int l = 0; l < 11; l++) {
            const Matrix4 &m = rev ? Drev[l] : D[l];
            for (int q = 0; q < 4; q++)
                for (int t = 0; t < 4; t++) {
                    double lik = 0;
                    for (int x = 0; x < 4; x++) {
                        double match = (double) std::max((ld) SMOOTHING_VALUE, m.p[q][x]);
                        lik += (se.p[x][t] * match);
                    }
                    out->log_lik[rev][l][q][t] = std::log(lik);
                }
        }
    out->log_excess = logf(excess_penal);
    out->log_rand = logf(rand_align_penal);
    return 0;
}

// D[11] / Drev[11] narrowed to double, [class][from][to] (inspection, and the entries the contig-merge scores read: host/merge.cpp)
extern "C" int cdh_damage_matrices(const char *prefix, double *D_out, double *Drev_out, char *errbuf, size_t errlen) {
    Matrix4 D[11], Drev[11];
    if (build_matrices(prefix, D, Drev, errbuf, errlen)) return -1;
    for (int l = 0; l < 11; l++)
        for (int q = 0; q < 4; q++)
            for (int t = 0; t < 4; t++) {
                D_out[l * 16 + q * 4 + t] = (double) D[l].p[q][t];
                Drev_out[l * 16 + q * 4 + t] = (double) Drev[l].p[q][t];
            }
    return 0;
}

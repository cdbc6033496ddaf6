// Host side of the contig-merge scores: the damage entries ancientMatchCount reads, and the ordering of the
// reference's priority queue, which stays on the host together with the greedy merge.
//   cdh_build_merge_model    D[5] / Drev[5] of initDeamProbabilities (nuclassembleUtil.cpp:821-1007) narrowed to double, the way
//                            `matchLik = subDeamDiNuc[5].p[qBase][tBase]` (:1117, :1151) reads them
//   cdh_merge_less           CompareNuclResultByScoreContigs (src/assembler/ancientContigsResults.cpp:24-70): Beta-binomial probability
//                            that the first overlap has the lower mismatch rate, from (alnLengthCons, deamMatch) of both
//   cdh_merge_queue_order    the order std::priority_queue hands out lines pushed in line order (:74, :256)
// The numeric types follow the reference because they decide the last bit: float alpha/beta, lgammaf through std::lgamma(float),
// double log / exp for the series.
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <queue>
#include <vector>

#include "../../include/carpedeam_gpu.h"

extern "C" int cdh_damage_matrices(const char *prefix, double *D /*[11][4][4]*/, double *Drev /*[11][4][4]*/, char *errbuf, size_t errlen);

extern "C" int cdh_build_merge_model(const char *prefix, cdg_merge_model *out, char *errbuf, size_t errlen) {
    std::vector<double> D(176), Drev(176);
    if (cdh_damage_matrices(prefix, D.data(), Drev.data(), errbuf, errlen)) return -1;
    for (int q = 0; q < 4; q++)
        for (int t = 0; t < 4; t++) {
            out->match_lik[0][q][t] = D[5 * 16 + q * 4 + t];
            out->match_lik[1][q][t] = Drev[5 * 16 + q * 4 + t];
        }
    return 0;
}

extern "C" int cdh_merge_less(uint32_t alnLenCons1, float deamMatch1, uint32_t alnLenCons2, float deamMatch2) {
    const float mm1 = alnLenCons1 - deamMatch1, mm2 = alnLenCons2 - deamMatch2;
    const float alpha1 = mm1 + 1, alpha2 = mm2 + 1;
    const float beta1 = deamMatch1 + 1, beta2 = deamMatch2 + 1;
    const double log_c = (std::lgamma(beta1 + beta2) + std::lgamma(alpha1 + beta1)) - (std::lgamma(alpha1 + beta1 + beta2) + std::lgamma(beta1));
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

extern "C" void cdh_merge_queue_order(const uint32_t *alnLenCons, const float *deamMatch, size_t m, uint32_t *order) {
    struct Item { uint32_t i, alc; float d; };
    struct Less { bool operator()(const Item &a, const Item &b) const { return cdh_merge_less(a.alc, a.d, b.alc, b.d) != 0; } };
    std::priority_queue<Item, std::vector<Item>, Less> pq;
    for (size_t i = 0; i < m; i++) pq.push(Item{(uint32_t) i, alnLenCons[i], deamMatch[i]});
    size_t k = 0;
    while (!pq.empty()) { order[k++] = pq.top().i; pq.pop(); }
}

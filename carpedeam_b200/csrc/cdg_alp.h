// E-value arithmetic shared by the CUDA kernels (rescore_kernel) and the host code (carpedeam_b200/host/expand.cpp), plain C++.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define CDG_ALP_HD __host__ __device__ __forceinline__
#else
#define CDG_ALP_HD static inline
#endif

// ---- E-values: Karlin-Altschul + ALP finite size correction ------------------------------------
// EvalueComputation.h:36-40 -> Sls::AlignmentEvaluer::area (sls_pvalues.cpp:366-545, blast_=false) with
// the Gumbel parameters ALP's initGapless derives for nucleotide.out (+2/-3, background 0.25): these
// are constants of the matrix, embedded here the way the reference embeds its gapped parameter sets
// (EvalueComputation.h:55-75).  Values dumped from the reference with oracle/ref_probe.cpp.
struct CdgAlp {
    static constexpr double lambda = 0.63373155264486880078;
    static constexpr double K = 0.40796623464181452912;
    static constexpr double a = 0.69454686319701297581;      // a_I = a_J, b = 0
    static constexpr double alpha = 0.83333515157614945768;  // alpha_I = alpha_J = sigma, beta = tau = 0
    static constexpr double thr = 2.6299310744375534732;     // vi_y_thr = vj_y_thr = c_y_thr
    static constexpr double logK = -0.89657086696128563;     // log(K), filled for reference; recomputed on host
};
CDG_ALP_HD double cdg_alp_area(double y, double m, double n) {
    const double const_val = 0.39894228040143267794;   // 1/sqrt(2*pi)
    double m_li_y = m - CdgAlp::a * y;
    double vi_y = fmax(CdgAlp::thr, CdgAlp::alpha * y);
    double sqrt_vi_y = sqrt(vi_y);
    double m_F = m_li_y / sqrt_vi_y;
    double P_m_F = 0.5 * erfc(-0.70710678118654752440 * m_F);
    double E_m_F = -const_val * exp(-0.5 * m_F * m_F);
    double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;
    double n_lj_y = n - CdgAlp::a * y;
    double n_F = n_lj_y / sqrt_vi_y;
    double P_n_F = 0.5 * erfc(-0.70710678118654752440 * n_F);
    double E_n_F = -const_val * exp(-0.5 * n_F * n_F);
    double p2 = n_lj_y * P_n_F - sqrt_vi_y * E_n_F;
    double c_y = vi_y;   // sigma == alpha and the same threshold
    return p1 * p2 + c_y * P_m_F * P_n_F;
}
CDG_ALP_HD double cdg_evalue(double score, double qLen, double dbRes) {
    return (CdgAlp::K * exp(-CdgAlp::lambda * score)) * cdg_alp_area(score, dbRes, qLen);
}
CDG_ALP_HD double cdg_bitscore(double score) {
    return (CdgAlp::lambda * score - log(CdgAlp::K)) / 0.69314718055994530942;
}


// RMP2 / RCCD / RCCSD drivers on device-resident data.  Internal.
#pragma once
#include "contract.h"
#include "transform.h"

namespace jues {

// Problem description after padding (every extent even; padded orbitals have zero coefficients and
// orbital energies far away from the physical ones, so every padded amplitude is exactly zero).
struct Problem {
    int64_t nao = 0, nocc = 0, nvir = 0;      // logical
    int64_t np = 0, o = 0, v = 0;             // padded
    DBuf Co, Cv;                              // (np x o), (np x v)
    DBuf eo, ev;                              // (o), (v)
};

void setup_problem(jues_ctx* ctx, Problem& P, int64_t nao, const double* Cao, int64_t nocc,
                   const double* Cav, int64_t nvir, const double* eps);

struct CCResult {
    double energy = 0.0;
    std::vector<double> e_hist;  // [maxit+1]
};

// RMP2 (RMP2.jl:11-45)
double rmp2_dev(jues_ctx* ctx, Problem& P, GaoSource& gao);

// RCCD (RCCD.jl:33-83) when singles == false, RCCSD (RCCSD.jl:33-116) when true.
// guess_mode (RCCD only): 0 = reference (ij|ab)/D, 1 = MP2.  T1_out/T2_out: HOST, unpadded, nullable.
CCResult cc_dev(jues_ctx* ctx, Problem& P, GaoSource& gao, bool singles, int maxit, int guess_mode,
                double* T1_out, double* T2_out, jues_b200_amp_cb cb, void* cb_user);

// Density-fitted variants.  pqP (nao,nao,naux) = (pq|P) and Jpqh (naux,naux) = (P|Q)^(-1/2) are HOST arrays, what
// DF.setup_df returns (DF.jl:30-51).  do_df_rmp2 (DF-RMP2.jl:1-46); do_df_rccd (DF-RCCD.jl:11-54): `maxit` sweeps
// from the MP2 guess with the density-fitted file's own ring intermediate (DF-RCCD.jl:248-258).
double df_rmp2_dev(jues_ctx* ctx, Problem& P, const double* pqP, int64_t naux, const double* Jpqh);
CCResult df_rccd_dev(jues_ctx* ctx, Problem& P, const double* pqP, int64_t naux, const double* Jpqh, int maxit,
                     double* T2_out, jues_b200_amp_cb cb, void* cb_user);

// AutoRCCSD.do_rccsd (AutoRCCSD.jl:193-301; options CoupledCluster.jl:36-43)
struct AutoOptions {
    int max_iter = 50;
    double e_conv = 1e-10, max_rms = 1e-10;
    bool do_pT = false;
};
struct AutoResult {
    double ecc = 0.0, ept = 0.0;
    bool has_pt = false, converged = false;
    int iterations = 0;
    std::vector<double> e_hist, rms_hist;   // [iterations + 1]; entry 0 = MP2 guess / 1.0
};
// foo (nocc,nocc), fov (nocc,nvir), fvv (nvir,nvir): HOST, off-diagonal Fock blocks (zero diagonals);
// P.eo / P.ev: the Fock diagonal.  T1_out/T2_out: HOST, unpadded, nullable.
AutoResult auto_rccsd_dev(jues_ctx* ctx, Problem& P, GaoSource& gao, const double* foo, const double* fov,
                          const double* fvv, const AutoOptions& opt, double* T1_out, double* T2_out,
                          jues_b200_amp_cb cb, void* cb_user);

// mRCCD.do_rccd (mRCCD.jl:37-120): RCCD from zero amplitudes with the reference's Float32 DIIS.
struct MrccdResult {
    double energy = 0.0;
    int iterations = 0;
    std::vector<double> rms_hist, e_hist;   // [iterations]: ||dT||_2 before, energy after the extrapolation
};
MrccdResult mrccd_dev(jues_ctx* ctx, Problem& P, GaoSource& gao, int maxit, double* T2_out,
                      jues_b200_amp_cb cb, void* cb_user);

// get_fock (IntegralTransformation.jl:119-141): f = C^T h C + 2 sum_k (pq|kk) - sum_k (pk|qk), k over the
// columns of Co.  hao (nao,nao), C (nao,nmo), Co (nao,nocc): HOST; f_out (nmo,nmo): HOST.
void fock_dev(jues_ctx* ctx, GaoSource& gao, const double* hao, const double* C, int64_t nmo,
              const double* Co, int64_t nocc, double* f_out);

}  // namespace jues

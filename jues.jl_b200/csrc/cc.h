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

}  // namespace jues

/*
 * jues_b200.h -- C ABI of libjues_b200.so, the B200 (sm_100a) implementation of the JuES.jl
 * hot path: 4-index AO->MO ERI transformation, RMP2 energy, RCCD / RCCSD amplitude iterations.
 *
 * The reference (mdav2/JuES.jl) has no FFI layer: its boundary is a set of Julia method
 * signatures.  Each entry point below sits *under* one of them and is reached with `ccall` from
 * a thin Julia shim (jues.jl_b200/julia/JuESB200.jl) or with ctypes (jues.jl_b200/__init__.py).
 * Citations are file:line in the reference tree.
 *
 * Conventions
 *   - every array is column-major IEEE FP64 exactly as Julia holds an Array{Float64,N}
 *     (element [i1,i2,i3,i4] of a (d1,d2,d3,d4) array at i1 + d1*(i2 + d2*(i3 + d3*i4)), 0-based);
 *     dims are int64_t (Julia Int).
 *   - pointers named *_host / plain `const double*` arguments are HOST pointers owned by the
 *     caller; the library never keeps them after the call returns.
 *   - calls are blocking and one-at-a-time per context (the reference entry points are
 *     single-task and blocking).
 *   - return 0 on success, a negative JUES_B200_E* code otherwise; the message is available
 *     from jues_b200_last_error().  No exception, exit() or signal crosses the ABI.
 *   - there is NO CPU fallback: without a CUDA device jues_b200_init fails with JUES_B200_ECUDA.
 */
#ifndef JUES_B200_H
#define JUES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JUES_B200_OK       0
#define JUES_B200_EINVAL  -1   /* bad shape / null pointer / unknown option            */
#define JUES_B200_ENOMEM  -2   /* device (or pinned host) capacity exceeded            */
#define JUES_B200_ECUDA   -3   /* CUDA runtime / driver error, or no device            */
#define JUES_B200_ENCCL   -4   /* NCCL error or NCCL library not loadable              */
#define JUES_B200_ESTATE  -5   /* call not valid in the current state of the handle    */

typedef struct jues_ctx jues_ctx;   /* opaque: device, stream, workspaces, communicator    */
typedef struct jues_t4  jues_t4;    /* opaque: device-resident rank-4 tensor (see below)    */

/* ---- context ------------------------------------------------------------------------- */
/* Create a context on CUDA device `device` (one context = one GPU = one process rank).     */
int  jues_b200_init(jues_ctx** ctx, int device);
void jues_b200_finalize(jues_ctx* ctx);
/* Message of the last failing call on this context ("" if none). ctx==NULL: global message
 * of a failed jues_b200_init.                                                              */
const char* jues_b200_last_error(jues_ctx* ctx);
/* Library / build identification: "jues_b200 <version> sm_100a".                           */
const char* jues_b200_version(void);

/* Multi-GPU (one process per GPU): attach rank/size and an NCCL unique id (128 bytes, made by
 * rank 0 with jues_b200_nccl_unique_id and broadcast by the host program, e.g. through
 * torch.distributed / MPI / a file).  Collectives used on the path: all-gather of the new T2
 * slab and all-reduce of energy partials.  The reference's only
 * multi-process design is ParCCD.jl:23-86,241-295 (Julia Distributed, slabs over virtual b).  */
/* Multi-GPU inside ONE process (what `Input.exec` / `com(JuWfn; ...)` needs, Input.jl:58-71: the reference's
 * entry points are called once, from one task): contexts on devices 0..ngpu-1 (ngpu <= 0: every visible
 * device), one NCCL communicator per device (ncclCommInitAll), returned as ONE handle -- the leader.  Every
 * entry point below that shards (rmp2, rccd, rccsd, auto_rccsd, mrccd and their _t4 forms with a generated
 * tensor) called on the leader runs on all devices, one host thread per GPU, and returns the leader's
 * results; the others run on the leader's GPU alone.  jues_b200_finalize(leader) releases everything.   */
int jues_b200_init_multi(jues_ctx** leader, int ngpu);
int jues_b200_group_size(jues_ctx* ctx);   /* devices behind this handle (1 for a plain context) */

int jues_b200_nccl_unique_id(unsigned char id_out[128]);
int jues_b200_init_dist(jues_ctx* ctx, int rank, int nranks, const unsigned char id[128]);

/* ---- BLAS.gemm! replacement ----------------------------------------------------------- */
/* C = alpha*op(A)*op(B) + beta*C, column-major, host pointers.  Replaces the explicit
 * `BLAS.gemm!` calls of src/CoupledCluster/mRCCD.jl:268-274,289-297,314-389,408-485 and is the
 * kernel every contraction of the path is built from (FP64 DMMA tiles fed by TMA).
 * transA/transB: 'N' or 'T'.                                                               */
int jues_b200_dgemm(jues_ctx* ctx, char transA, char transB, int64_t M, int64_t N, int64_t K,
                    double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                    double beta, double* C, int64_t ldc);

/* ---- Transformation.tei_transform (src/Backend/Transformation.jl:15-26, 39-93) ---------- */
/* out[i,a,j,b] = sum_{mu nu lam sig} C1[mu,i] C2[nu,a] C3[lam,j] C4[sig,b] gao[mu,nu,lam,sig]
 * gao: (nao,nao,nao,nao); Ck: (nao,dk); out: (d1,d2,d3,d4) in chemists' order, or, when
 * phys_order != 0, permuted [1,3,2,4] -> (d1,d3,d2,d4) as every caller of the reference does
 * (RCCD.jl:91-95, RCCSD.jl:118-132; IntegralTransformation.jl:96-98).                       */
int jues_b200_tei_transform(jues_ctx* ctx, const double* gao, int64_t nao,
                            const double* C1, int64_t d1, const double* C2, int64_t d2,
                            const double* C3, int64_t d3, const double* C4, int64_t d4,
                            int phys_order, double* out);

/* ---- MollerPlesset.do_rmp2 (src/MollerPlesset/RMP2.jl:11-45) ---------------------------- */
/* E = sum_{ijab} v_ijab (2 v_ijab - v_ijba) / (eps_i + eps_j - eps_a - eps_b), v = <ij|ab>.
 * eps: (nocc+nvir) orbital energies, occupied first (Wfn.epsa, Wavefunction.jl:85).         */
int jues_b200_rmp2(jues_ctx* ctx, const double* gao, int64_t nao,
                   const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                   const double* eps, double* e_mp2);

/* ---- CoupledCluster.RCCD.do_rccd (src/CoupledCluster/RCCD.jl:33-83) --------------------- */
/* Limit of every coupled-cluster entry point (rccd, rccsd, auto_rccsd, mrccd, df_rccd): nocc <= 158.  The
 * kernels that combine T2 with its (i,j)-transposed partner stage one nocc x nocc block in shared memory
 * (200 KB); a larger occupied space returns JUES_B200_EINVAL ("nocc too large ...").  The reference has no
 * such limit; BASELINE's largest configuration has nocc = 60.
 *
 * maxit Jacobi sweeps with no convergence test (reference: maxit = 40, RCCD.jl:34,55).
 * guess_mode 0 = reference guess T2 = (ij|ab)/D (RCCD.jl:45,145-160); 1 = MP2 guess.
 * e_hist (nullable): [maxit+1] energies, e_hist[0] = energy of the guess, e_hist[k] after
 * sweep k.  T2_out (nullable): (nocc,nocc,nvir,nvir) final amplitudes (return_T2, RCCD.jl:36,78).
 * Returns the correlation energy of the last sweep in *e_ccd (RCCD.jl:81).                  */
int jues_b200_rccd(jues_ctx* ctx, const double* gao, int64_t nao,
                   const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                   const double* eps, int maxit, int guess_mode,
                   double* e_ccd, double* e_hist, double* T2_out);

/* ---- CoupledCluster.RCCSD.do_rccsd (src/CoupledCluster/RCCSD.jl:33-116) ----------------- */
/* T1 = 0, T2 = <ij|ab>/D guess (RCCSD.jl:70,80); maxit sweeps (reference: 40, RCCSD.jl:36,88);
 * both updates use the old amplitudes (RCCSD.jl:160-169).  e_hist as above (the reference
 * prints it every sweep, RCCSD.jl:104).  T1_out (nocc,nvir), T2_out (nocc,nocc,nvir,nvir)
 * nullable (return_T, RCCSD.jl:111-115).                                                    */
int jues_b200_rccsd(jues_ctx* ctx, const double* gao, int64_t nao,
                    const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                    const double* eps, int maxit,
                    double* e_ccsd, double* e_hist, double* T1_out, double* T2_out);

/* Per-sweep amplitude capture for parity tests: called (on the host, from inside
 * jues_b200_rccd / _rccsd) after the guess (it = 0) and after every sweep with host copies of
 * the amplitudes in the caller-visible (unpadded) layout.  T1 is NULL for RCCD.  Set to NULL to
 * disable (default).                                                                        */
typedef void (*jues_b200_amp_cb)(void* user, int it, double energy,
                                 const double* T1, const double* T2);
int jues_b200_set_amplitude_callback(jues_ctx* ctx, jues_b200_amp_cb cb, void* user);

/* ---- device-resident rank-4 tensor: the DiskFourTensor replacement ---------------------- */
/* src/DiskTensors/DiskFourTensors.jl:5-95 keeps a rank-4 Float64 tensor in an HDF5 file with
 * slice getindex / setindex!, eltype and blockfill!.  Here the tensor lives in HBM (on the
 * context's GPU) behind the same surface; ranges are 0-based half-open [lo,hi) per index.    */
int jues_b200_t4_create(jues_ctx* ctx, int64_t d1, int64_t d2, int64_t d3, int64_t d4, jues_t4** out);
int jues_b200_t4_destroy(jues_t4* t);
int jues_b200_t4_dims(const jues_t4* t, int64_t dims_out[4]);
int jues_b200_t4_fill(jues_t4* t, double value);                            /* blockfill! :88-95 */
int jues_b200_t4_set_slice(jues_t4* t, const int64_t lo[4], const int64_t hi[4], const double* host); /* setindex! :57-80 */
int jues_b200_t4_get_slice(const jues_t4* t, const int64_t lo[4], const int64_t hi[4], double* host); /* getindex :43-53 */
/* Fill with the counter-based synthetic 8-fold-symmetric ERIs of SURVEY.md section 8d
 * (bit-identical to jues.jl_b200.synth.counter_eri on the host).                             */
int jues_b200_t4_synth_eri(jues_t4* t, uint64_t seed, double scale);
/* The same synthetic tensor WITHOUT storage: a read-only (nao,nao,nao,nao) handle whose sigma slabs
 * are generated on demand while a transformation streams through them -- for shapes whose dense
 * N^4 does not fit in HBM (BASELINE configs 4 and 5: 358-500 GB).                              */
int jues_b200_t4_create_synth(jues_ctx* ctx, int64_t nao, uint64_t seed, double scale, jues_t4** out);

/* Entry points taking a device-resident gao (Wfn.ao_eri::DiskFourTensor dispatch,
 * Wavefunction.jl:87; Transformation.jl:94-192).  Inputs are already in HBM when they start. */
int jues_b200_tei_transform_t4(jues_ctx* ctx, const jues_t4* gao,
                               const double* C1, int64_t d1, const double* C2, int64_t d2,
                               const double* C3, int64_t d3, const double* C4, int64_t d4,
                               int phys_order, jues_t4** out);
int jues_b200_rmp2_t4(jues_ctx* ctx, const jues_t4* gao,
                      const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                      const double* eps, double* e_mp2);
int jues_b200_rccd_t4(jues_ctx* ctx, const jues_t4* gao,
                      const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                      const double* eps, int maxit, int guess_mode,
                      double* e_ccd, double* e_hist, double* T2_out);
int jues_b200_rccsd_t4(jues_ctx* ctx, const jues_t4* gao,
                       const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                       const double* eps, int maxit,
                       double* e_ccsd, double* e_hist, double* T1_out, double* T2_out);

/* ======== SURVEY.md section 8f: the callers either side of the path ====================== */

/* ---- IntegralTransformation.get_fock (src/Backend/IntegralTransformation.jl:119-141) ------ */
/* f[p,q] = C[mu,p] C[nu,q] hao[mu,nu] + 2 C C Co Co gao[mu,nu,lam,sig] - C C Co Co gao[mu,lam,nu,sig]
 * hao (nao,nao) core Hamiltonian (Wfn.hao), C (nao,nmo) = Wfn.Ca, Co (nao,nocc) = Wfn.Cao;
 * f_out (nmo,nmo).  Two N^4 nocc transforms on the device + traces.                           */
int jues_b200_get_fock(jues_ctx* ctx, const double* gao, int64_t nao, const double* hao,
                       const double* C, int64_t nmo, const double* Co, int64_t nocc, double* f_out);
int jues_b200_get_fock_t4(jues_ctx* ctx, const jues_t4* gao, const double* hao,
                          const double* C, int64_t nmo, const double* Co, int64_t nocc, double* f_out);

/* ---- CoupledCluster.AutoRCCSD.do_rccsd (src/CoupledCluster/AutoRCCSD.jl:193-301) ---------- */
/* Options = CoupledCluster.defaults (src/CoupledCluster/CoupledCluster.jl:36-43).            */
typedef struct {
    int    cc_max_iter;   /* 50     maximum number of sweeps (AutoRCCSD.jl:271)                 */
    double cc_e_conv;     /* 1e-10  stop when |dE| <= cc_e_conv ...                             */
    double cc_max_rms;    /* 1e-10  ... and max(r1,r2) <= cc_max_rms, r = ||dT||_2/length(T) (:174-175,270) */
    int    do_pT;         /* 0      add the perturbative triples (:294-300)                     */
    int    fcn;           /* 0      number of frozen core orbitals (:216, get_eri fcn=)         */
    int    diis;          /* 0      accepted and ignored, as in the reference                   */
} jues_b200_cc_options;
int jues_b200_cc_default_options(jues_b200_cc_options* opt);

/* RCCSD for a possibly non-canonical RHF reference: the Fock matrix is built from hao and gao
 * (get_fock), its diagonal gives the resolvents, its off-diagonal blocks enter the amplitude
 * equations (AutoRCCSD.jl:219-231).  Ca (nao,nmo): all MO coefficients, ndocc doubly occupied first
 * (Wfn.Ca, Wfn.nalpha); frozen core = the first opt->fcn of them.  Guess T1 = f_ov/d, T2 = <ij|ab>/D.
 * Outputs: *e_cc correlation energy; *e_pt (T) correction (written only if opt->do_pT; nullable
 * otherwise); *iterations sweeps done; *converged 1/0 (:288).  e_hist / rms_hist (nullable):
 * [cc_max_iter+1], entry 0 = energy of the guess / 1.0, entry k = after sweep k (the table the
 * reference prints, :285).  T1_out (nact,nvir), T2_out (nact,nact,nvir,nvir) nullable,
 * nact = ndocc - fcn, nvir = nmo - ndocc.  opt == NULL: defaults.                              */
int jues_b200_auto_rccsd(jues_ctx* ctx, const double* gao, int64_t nao, const double* hao,
                         const double* Ca, int64_t nmo, int64_t ndocc,
                         const jues_b200_cc_options* opt,
                         double* e_cc, double* e_pt, int* iterations, int* converged,
                         double* e_hist, double* rms_hist, double* T1_out, double* T2_out);
int jues_b200_auto_rccsd_t4(jues_ctx* ctx, const jues_t4* gao, const double* hao,
                            const double* Ca, int64_t nmo, int64_t ndocc,
                            const jues_b200_cc_options* opt,
                            double* e_cc, double* e_pt, int* iterations, int* converged,
                            double* e_hist, double* rms_hist, double* T1_out, double* T2_out);

/* ---- PerturbativeTriples.compute_pT (src/CoupledCluster/PerturbativeTriples.jl:35-138) ---- */
/* Stand-alone (T) from host arrays in the reference's own argument layouts:
 * T1 (o,v), T2 (o,o,v,v), Vvvvo (v,v,v,o), Vvooo (v,o,o,o), Vvovo (v,o,v,o), fo (o), fv (v).      */
int jues_b200_compute_pt(jues_ctx* ctx, const double* T1, const double* T2, const double* Vvvvo,
                         const double* Vvooo, const double* Vvovo, const double* fo, const double* fv,
                         int64_t nocc, int64_t nvir, double* e_pt);

/* ---- CoupledCluster.mRCCD.do_rccd (src/CoupledCluster/mRCCD.jl:37-120, 143-207) ---------- */
/* RCCD from zero amplitudes with the reference's DIIS: at most 6 vectors kept in Float32, B
 * normalised by max|B|, Float32 solve, stop when ||T2new - T2old||_2 < 1e-7 or after maxit sweeps
 * (maxit IS an honoured keyword of this entry point, mRCCD.jl:37).  *iterations: sweeps done;
 * rms_hist / e_hist (nullable): [maxit] norm of the un-extrapolated change / energy after the
 * extrapolation, per sweep.  T2_out nullable (return_T2).  Amplitudes carry Float32 rounding by
 * construction (the extrapolated T2 is a combination of Float32 vectors): parity with the CPU path
 * is bounded by that, not by FP64 round-off.                                                  */
int jues_b200_mrccd(jues_ctx* ctx, const double* gao, int64_t nao,
                    const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                    const double* eps, int maxit,
                    double* e_ccd, int* iterations, double* rms_hist, double* e_hist, double* T2_out);
int jues_b200_mrccd_t4(jues_ctx* ctx, const jues_t4* gao,
                       const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                       const double* eps, int maxit,
                       double* e_ccd, int* iterations, double* rms_hist, double* e_hist, double* T2_out);

/* ---- density-fitted variants ------------------------------------------ */
/* MollerPlesset.do_df_rmp2 (src/MollerPlesset/DF-RMP2.jl:1-46) and CoupledCluster.DFRCCD.do_df_rccd
 * (src/CoupledCluster/DF-RCCD.jl:11-54).  pqP (nao,nao,naux) = (pq|P) and Jpqh (naux,naux) = (P|Q)^(-1/2)
 * are exactly what DF.setup_df returns (src/Backend/DF.jl:30-51; the reference obtains them from psi4).
 * b[p,q,Q] = pqP[p,q,P] Jpqh[P,Q] (DF.jl:52-58) is never formed in the AO basis: the orbital coefficients
 * are contracted first.
 * df_rmp2: E = sum (ia|jb) (2 (ia|jb) - (ib|ja)) / D with (ia|jb) = b[i,a,Q] b[j,b,Q].
 * df_rccd: `maxit` sweeps (this driver honours its keyword, DF-RCCD.jl:28) from the MP2 guess (:111-135)
 * with the integral classes built from b; the ring intermediate WmBeJ follows DF-RCCD.jl:248-258, which
 * contracts <mn|ef> where RCCD.jl:399-402 contracts <nm|ef>.  e_hist (nullable): [maxit+1]; T2_out
 * (nullable): (nocc,nocc,nvir,nvir).                                                                */
int jues_b200_df_rmp2(jues_ctx* ctx, const double* pqP, int64_t nao, int64_t naux, const double* Jpqh,
                      const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                      const double* eps, double* e_mp2);
int jues_b200_df_rccd(jues_ctx* ctx, const double* pqP, int64_t nao, int64_t naux, const double* Jpqh,
                      const double* Cao, int64_t nocc, const double* Cav, int64_t nvir,
                      const double* eps, int maxit, double* e_ccd, double* e_hist, double* T2_out);

/* ---- instrumentation (bench.py) ---------------------------------------------------------- */
/* Statistics of the last entry-point call on this context: CUDA-event milliseconds per phase
 * on the library's stream, FP64 flops issued by the GEMM kernels, kernel launch counts.
 * Keys: see DESIGN.md ("instrumentation").  Returns the number of phases written.            */
typedef struct {
    char   name[48];
    double ms;
} jues_b200_phase;
int jues_b200_get_phases(jues_ctx* ctx, jues_b200_phase* out, int cap);
/* Trace level of the following calls: 0 coarse phases; 1 also cc.part.*, cc.comm.*, tei.* regions
 * (the JUES_B200_TRACE=1 default); 2 also one CUDA-event pair around every DGEMM launch, reported as
 * phase "gemm MxNxKxbatch" (in-situ launch durations).  Tracing runs the sweeps eagerly (no graph replay). */
int jues_b200_set_trace(jues_ctx* ctx, int level);
int jues_b200_get_counters(jues_ctx* ctx, double* gemm_flops, int64_t* gemm_launches,
                           int64_t* aux_launches, int64_t* bytes_peak);
/* NCCL collectives issued by the last call and the bytes this rank received in them.           */
int jues_b200_get_comm_counters(jues_ctx* ctx, int64_t* collectives, double* bytes_received);
/* Kernel-level check of the packed (symmetric/antisymmetric) particle-particle ladder, the largest
 * contraction of the sweeps (form_T2 term T2[i,j,e,f]*Wabef, RCCD.jl:247; RCCSD.jl:268):
 * out[i,j,a,b] = sum_ef tau[i,j,e,f] W4[e,f,a,b] for W4[e,f,a,b] = <ef|ab> (= W4[f,e,b,a]), evaluated
 * exactly as `nslabs` ranks would: per slab of the last index its own [W+|W-] block, one batched GEMM,
 * the blocks side by side instead of all-gathered, every slab unpacked.  tau (nocc,nocc,nvir,nvir),
 * W4 (nvir,nvir,nvir,nvir), out (nocc,nocc,nvir,nvir): host.                                   */
int jues_b200_sa_ladder(jues_ctx* ctx, const double* tau, const double* W4, int64_t nocc, int64_t nvir,
                        int nslabs, double* out);
/* Time `reps` back-to-back launches of the DGEMM kernel on device-resident random operands
 * (no host traffic); returns average ms per launch.  Used for the roofline line.            */
int jues_b200_dgemm_bench(jues_ctx* ctx, char transA, char transB, int64_t M, int64_t N, int64_t K,
                          int reps, double* ms_avg);

/* Determinism stress of the DGEMM kernel: the product is launched reps+1 times on device-resident
 * pseudo-random operands and every repetition compared with the first ON THE DEVICE; *n_bad = number of
 * repetitions that differ (0 expected), *worst_sqdiff = largest sum of squared differences.  batch > 1:
 * the quarter-transform layout A[M,K,batch], B shared, C[M,N,batch] ('N','N').                */
int jues_b200_dgemm_stress(jues_ctx* ctx, char transA, char transB, int64_t M, int64_t N, int64_t K,
                           int64_t batch, int reps, int* n_bad, double* worst_sqdiff);

/* Determinism stress of the 4-index transform (test hook): reps+1 identical transforms inside one call,
 * the output of each of the four quarter transforms compared on the device with the first run's.
 * stats[reps][4 quarters][4] = {elements that differ (0 expected), first and last differing linear index
 * (-1: none), largest |difference|} per repetition and quarter (in execution order).           */
int jues_b200_transform_stress(jues_ctx* ctx, const jues_t4* gao,
                               const double* C1, int64_t d1, const double* C2, int64_t d2,
                               const double* C3, int64_t d3, const double* C4, int64_t d4,
                               int reps, double* stats);

#ifdef __cplusplus
}
#endif
#endif /* JUES_B200_H */

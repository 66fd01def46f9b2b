/*
 * neci_gpu.h -- C ABI of the B200-native FCIQMC walker-propagation engine.
 *
 * This is the drop-in boundary for the hot path of NECI's PerformFCIMCycPar
 * (reference: src/FciMCPar.F90:1177-1920).  The reference has no FFI seam for
 * this path; the replaceable unit is the single call
 *     call PerformFciMCycPar(iter_data_fciqmc, err)        (src/FciMCPar.F90:507)
 * plus the places the Fortran host reads/writes the walker arrays between
 * iterations.  A Fortran shim binds these symbols through
 * ISO_C_BINDING, following the conventions of src/lib/dSFMT_interface.F90:31-42
 * (scalars by value, arrays by reference, integer status return).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero on
 *     failure (mirrors `integer, intent(out) :: err`, src/FciMCPar.F90:1188).
 *   - the host owns every pointer it passes; the engine copies what it needs
 *     into HBM before returning and never retains host pointers.
 *   - one engine handle == one MPI rank ("node") of the reference == one GPU.
 *   - orbitals are 1-based spin orbitals exactly as in NECI (odd = beta,
 *     even = alpha, src/macros.h:16,21); bit (o-1)%64 of word (o-1)/64.
 *   - walker records use NECI's AoS layout ilut(0:NIfTot):
 *       words 0..nifd  : occupation bits            (src/BitReps.F90:164-306)
 *       word  nifd+1   : sign, fp64 bit-cast        (src/bit_rep_data.F90:152)
 *       word  nifd+2   : flags                      (src/bit_rep_data.F90:87-119)
 *     (the `neci` build: lenof_sign = 1, inum_runs = 1).
 */
#ifndef NECI_GPU_H
#define NECI_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- flag bits of the last ilut word (src/bit_rep_data.F90:87-119) ------- */
#define NECI_FLAG_REMOVED        0
#define NECI_FLAG_DETERM_PARENT  1
#define NECI_FLAG_TRIAL          2
#define NECI_FLAG_CONNECTED      3
#define NECI_FLAG_INITIATOR      13
#define NECI_FLAG_STATIC_INIT    16
#define NECI_FLAG_DETERMINISTIC  19

/* ---- system (Hamiltonian / excitation generator) selector ---------------- */
#define NECI_SYS_FCIDUMP_PCHB    1   /* sltcnd + PCHB doubles + uniform singles */
#define NECI_SYS_HUBBARD_RS      2   /* gen_excit_rs_hubbard                    */
#define NECI_SYS_HUBBARD_K       3   /* gen_excit_k_space_hub                   */

/* ---- per-iteration statistics returned by neci_gpu_iterate ---------------
 * The per-rank accumulators that communicate_estimates reduces
 * (src/fcimc_iter_utilities.F90:466-535,695-709), in a fixed order.          */
enum neci_stat_index {
    NECI_ST_NOBORN = 0,          /* NoBorn            fcimc_pointed_fns.F90:553 */
    NECI_ST_NODIED,              /* NoDied            fcimc_helper.F90:2336     */
    NECI_ST_ANNIHILATED,         /* Annihilated       Annihilation.F90:601,1117 */
    NECI_ST_NOABORTED,           /* NoAborted         Annihilation.F90:1092     */
    NECI_ST_NOREMOVED,           /* NoRemoved         Annihilation.F90:1374     */
    NECI_ST_SPAWNFROMSING,       /* SpawnFromSing                               */
    NECI_ST_ACCEPTANCES,         /* acceptances       fcimc_helper.F90:304      */
    NECI_ST_HFCYC,               /* HFCyc / NoatHF contribution (signed)        */
    NECI_ST_NOATDOUBS,           /* NoatDoubs                                   */
    NECI_ST_ENUMCYC,             /* ENumCyc           fcimc_helper.F90:757      */
    NECI_ST_ENUMCYCABS,          /* ENumCycAbs                                  */
    NECI_ST_INITSENUMCYC,        /* InitsENumCyc                                */
    NECI_ST_NOINITDETS,          /* NoInitDets        fcimc_helper.F90:1083     */
    NECI_ST_NONONINITDETS,
    NECI_ST_NOINITWALK,
    NECI_ST_NONONINITWALK,
    NECI_ST_NOADDEDINITIATORS,   /* net change this iteration (can be < 0)      */
    NECI_ST_NVALIDEXCITS,        /* nValidExcits      FciMCPar.F90:1649         */
    NECI_ST_NINVALIDEXCITS,      /* nInvalidExcits    FciMCPar.F90:1647         */
    NECI_ST_BLOOM_COUNT_1,       /* bloom_count(1)                              */
    NECI_ST_BLOOM_COUNT_2,
    NECI_ST_MAX_CYC_SPAWN,       /* max-reduced                                 */
    NECI_ST_BLOOM_SIZE_1,        /* max-reduced                                 */
    NECI_ST_BLOOM_SIZE_2,        /* max-reduced                                 */
    NECI_ST_TAU_GAMMA_SING,      /* tau search: max |H_ij|/(pgen/pSingles) this iteration, max-reduced
                                    (log_spawn_magnitude, tau/tau_search_conventional.F90:138-260)  */
    NECI_ST_TAU_GAMMA_DOUB,      /* doubles without the parallel bias, max-reduced */
    NECI_ST_TAU_GAMMA_PAR,       /* same-spin doubles, max-reduced              */
    NECI_ST_TAU_GAMMA_OPP,       /* opposite-spin doubles, max-reduced          */
    NECI_ST_TAU_MAX_DEATH_CPT,   /* max (K_ii - S) over the determinants that attempted death this iteration
                                    (log_death_magnitude, tau/tau_main.F90:198-207; fcimc_pointed_fns.F90:640), max-reduced;
                                    0 when the tau search is off                 */
    NECI_ST_TOTPARTS,            /* after CalcHashTableStats load_balancer.fpp:738 */
    NECI_ST_NORM_PSI_SQ,         /* norm_psi_squared                            */
    NECI_ST_NORM_SEMISTOCH_SQ,
    NECI_ST_INSTNOATHF,          /* InstNoatHF                                  */
    NECI_ST_TOTWALKERS,          /* list length incl. holes (TotWalkers)        */
    NECI_ST_HOLESINLIST,         /* HolesInList                                 */
    NECI_ST_NSPAWNED_SENT,       /* spawn records produced by this rank         */
    NECI_ST_NSPAWNED_RECV,       /* spawn records received by this rank         */
    NECI_ST_NSPAWNED_MERGED,     /* unique determinants after CompressSpawnedList */
    NECI_ST_NINSERTED,           /* determinants newly added (AddNewHashDet)    */
    NECI_ST_HIGHEST_POP,         /* iHighestPop, max-reduced                    */
    NECI_ST_TRIAL_NUMERATOR,     /* trial_numerator   fcimc_helper.F90:586-648  */
    NECI_ST_TRIAL_DENOM,         /* trial_denom                                 */
    NECI_ST_INIT_TRIAL_NUMERATOR,/* init_trial_numerator                        */
    NECI_ST_INIT_TRIAL_DENOM,    /* init_trial_denom                            */
    NECI_ST_TAU_CNT_SING,        /* spawns logged by log_spawn_magnitude per class (cnt_sing / cnt_doub / cnt_par / cnt_opp) */
    NECI_ST_TAU_CNT_DOUB,
    NECI_ST_TAU_CNT_PAR,
    NECI_ST_TAU_CNT_OPP,
    NECI_ST_ERR_FLAGS,           /* bit0 spawn overflow, bit1 list overflow,
                                    bit2 death prob > 2, bit3 hash overflow     */
    NECI_ST_TIME_SPAWN_MS,       /* device time of the spawn/death pass         */
    NECI_ST_TIME_COMM_MS,
    NECI_ST_TIME_ANNIHIL_MS,
    NECI_ST_TIME_DETERM_MS,
    NECI_ST_COUNT
};
#define NECI_ST_FIRST_MAX NECI_ST_MAX_CYC_SPAWN   /* [FIRST_MAX, LAST_MAX] are max-reduced */
#define NECI_ST_LAST_MAX  NECI_ST_TAU_MAX_DEATH_CPT

/* ---- configuration: the module-level globals the hot path reads ----------
 * (filled at the end of InitFCIMCCalcPar, src/FciMCPar.F90:256)              */
typedef struct neci_gpu_config {
    int32_t nel;                  /* SystemData::nel                           */
    int32_t nbasis;               /* spin orbitals (<= 128)                    */
    int32_t nifd;                 /* int(nbasis/64)   src/BitReps.F90:174      */
    int32_t niftot;               /* nifd + 2         src/BitReps.F90:214-286  */
    int32_t nocc_alpha, nocc_beta;
    int32_t nranks, rank;         /* nNodes, iProcIndex                        */
    int32_t device;               /* CUDA device ordinal for this rank         */
    int32_t balance_blocks;       /* load_balancer.fpp:86-99                   */
    int64_t max_walkers;          /* MaxWalkersPart fcimc_initialisation.fpp:1650 */
    int64_t max_spawned;          /* MaxSpawned     fcimc_initialisation.fpp:3678 */
    int32_t system_type;          /* NECI_SYS_*                                */
    int32_t t_trunc_initiator;    /* tTruncInitiator                           */
    int32_t t_all_real_coeff;     /* tAllRealCoeff                             */
    int32_t t_real_spawn_cutoff;  /* tRealSpawnCutoff                          */
    int32_t t_death_before_comms; /* fcimc_initialisation.fpp:1999-2001        */
    int32_t t_init_coherent_rule; /* tInitCoherentRule Annihilation.F90:579    */
    int32_t t_no_brillouin;       /* tNoBrillouin    matel_getter.F90:90       */
    int32_t t_exch;               /* tExch           sltcnd.fpp:611            */
    int32_t t_semi_stochastic;    /* tSemiStochastic                           */
    int32_t t_core_inits;         /* t_core_inits    Calc.F90:125              */
    int32_t t_tau_search;         /* tau_search_method /= OFF: log spawn magnitudes (fcimc_pointed_fns.F90:428-434) */
    int32_t t_consider_par_bias;  /* consider_par_bias  tau/tau_search_conventional.F90:66-117 */
    int32_t t_hphf;               /* tHPHF (even S): walkers live on HPHF functions, represented by the determinant
                                     IsAllowedHPHF accepts (src/DetBitOps.F90:693-718); generate_excitation =
                                     gen_hphf_excit (src/HPHFRandExcit.F90:175-476), matrix elements from
                                     src/HPHFIntegrals.fpp.  FCIDUMP/PCHB systems only.                       */
    int32_t reserved0;            /* keeps the doubles 8-byte aligned; must be 0 */
    double  initiator_walk_no;    /* InitiatorWalkNo                           */
    double  real_spawn_cutoff;    /* RealSpawnCutoff                           */
    double  occupied_thresh;      /* OccupiedThresh                            */
    double  av_mc_excits;         /* AvMCExcits                                */
    double  hii;                  /* Hii (reference energy, incl. ECore)       */
    double  ecore;                /* ECore                                     */
    uint64_t seed;                /* Philox key (replaces the dSFMT seed)      */
    const int32_t *random_orb_index;    /* RandomOrbIndex(nbasis)  fcimc_initialisation.fpp:862 */
    const int32_t *random_hash2;        /* RandomHash2(nbasis)                                  */
    const int32_t *load_balance_mapping;/* LoadBalanceMapping(balance_blocks), 0-based ranks    */
    const int64_t *ilut_ref;            /* iLutRef(0:nifd,1)                                    */
} neci_gpu_config;

typedef struct neci_gpu_engine neci_gpu_engine;   /* opaque */

/* ---- lifetime ------------------------------------------------------------ */
/* Called at the end of InitFCIMCCalcPar (src/FciMCPar.F90:256).  On failure (non-zero return) *out still holds a
 * handle: it carries the message for neci_gpu_last_error and whatever was allocated before the failure, accepts no
 * other call, and must be released with neci_gpu_finalize.                                                       */
int neci_gpu_init(const neci_gpu_config *cfg, neci_gpu_engine **out);
/* Called from DeallocFCIMCMemPar.                                            */
int neci_gpu_finalize(neci_gpu_engine *e);
/* Text of the last error on this handle (never NULL).                        */
const char *neci_gpu_last_error(const neci_gpu_engine *e);

/* ---- read-only system tables (uploaded once) ------------------------------ */
/* UMAT: packed 8-fold array indexed by UMatInd (src/UMatCache.F90:257-296),
 * 1-based index i stored at umat[i-1]; TMAT2D: dense nbasis x nbasis,
 * column-major as in Fortran, tmat[(i-1) + nbasis*(j-1)] (src/OneEInts.F90:221). */
int neci_gpu_set_system_fcidump(neci_gpu_engine *e, const double *umat, int64_t n_umat,
                                const double *tmat2d);
/* PCHB spatial-orbital tables built by the host
 * (src/gasci_pchb_doubles_spatorb_fastweighted.fpp:329-445) and the
 * single/double and parallel biases.  Sampler s in {0:SAME_SPIN,
 * 1:OPP_SPIN_NO_EXCH, 2:OPP_SPIN_EXCH}; entry (ij, s, ab) at
 * [((ij-1)*3 + s)*ab_max + (ab-1)].  alias is 1-based; a sampler whose weights
 * sum to zero has alias[..first entry..] == 0 (AliasSampler_t::sample returns
 * tgt = 0, src/aliasSampling.F90:439-443).  tgt_orbs[2*(ab-1) + {0,1}] are the
 * spatial orbitals of pair index ab (smaller first).
 * class_of_spinorb / class lists drive the uniform singles generator
 * (src/GenRandSymExcitNUMod.F90:1118-1286).                                   */
int neci_gpu_set_pchb(neci_gpu_engine *e, int32_t n_spat, int32_t ij_max, int32_t ab_max,
                      const double *probs, const double *bias, const int32_t *alias,
                      const double *p_exch, const int32_t *tgt_orbs,
                      double p_singles, double p_doubles, double p_parallel,
                      int32_t n_classes, const int32_t *class_of_spinorb);
/* PCHB particle selection (PCHB_ParticleSelection_t, src/gasci_pchb_doubles_select_particles.fpp:38-46): mode 0 =
 * UNIF-UNIF (pick_biased_elecs, the default after neci_gpu_set_pchb; the tables are ignored), mode 1 = FULL-FULL
 * (PC_FullyWeightedParticles_t, :330-438): the first particle I is drawn with p_first[I] and the second with
 * p_second[I][J] = p(J | I), both restricted to the occupied orbitals and renormalised (constrained sampling), and
 * p({I, J}) sums both orders.  p_first[nBasis] and p_second[nBasis * nBasis] (row I) are the normalised probabilities
 * of the selector's I_sampler / J_sampler (AliasSampler_t::get_prob), spin orbitals in NECI order.  mode 2 = UNIF-FULL
 * (PC_WeightedParticles_t, :440-506): the first particle uniformly among the electrons, the second as above, p = (p(J | I)
 * + p(I | J)) / nEl.  Call after neci_gpu_set_pchb; not available together with t_hphf.  (UNIF-FAST is not built.) */
int neci_gpu_set_pchb_particles(neci_gpu_engine *e, int32_t mode, const double *p_first, const double *p_second);

/* New excitation-class biases from the tau search (update_tau, src/tau/tau_search_conventional.F90:274-499 assigns
 * pSingles / pDoubles / pParallel); takes effect from the next iteration.  FCIDUMP/PCHB systems.               */
int neci_gpu_set_excit_probs(neci_gpu_engine *e, double p_singles, double p_doubles, double p_parallel);
/* Real-space Hubbard: spin-orbital neighbour lists in the order of
 * lat%get_spinorb_neighbors (src/real_space_hubbard.F90:1986), padded with 0;
 * TMAT2D holds the hopping (bhub) and uhub is U.                              */
int neci_gpu_set_system_hubbard_rs(neci_gpu_engine *e, int32_t max_neigh,
                                   const int32_t *neighbours, const double *tmat2d, double uhub);
/* k-space Hubbard: n_k k-points; spatial orbital s (1-based) has k index s-1;
 * ksum[k1*n_k + k2] = index of k1+k2, kdiff[k1*n_k+k2] = index of k1-k2
 * (what lat%get_orb_from_k_vec resolves, src/lattice_models_utils.F90:1881-1925);
 * eps_k = one-body energies; u_over_n = UMAT(1) = UHUB/OMEGA
 * (src/Integrals_neci.F90:643).                                               */
int neci_gpu_set_system_hubbard_k(neci_gpu_engine *e, int32_t n_k, const int32_t *ksum,
                                  const int32_t *kdiff, const double *eps_k, double u_over_n);
/* Semi-stochastic core space: CSR flattening of this rank's rows of sparse_core_ham
 * (src/fast_determ_hamil.F90:1421-1507; diagonal has Hii subtracted), 0-based
 * columns into the gathered vector.  core_iluts holds ALL core determinants
 * (sum(sizes) entries of nifd+1 words) in core-space order, i.e. rank-major:
 * rank r owns entries [displs[r], displs[r] + sizes[r]) -- the reference keeps
 * the whole core space on every rank too (core_space + its hash table,
 * src/core_space_util.F90:20-90) because is_core_state (src/semi_stoch_procs.F90:547)
 * must recognise core determinants owned by other ranks.  n_local must equal
 * sizes[rank] (anything else is an error).                                     */
int neci_gpu_set_core_space(neci_gpu_engine *e, int64_t n_local, const int64_t *row_ptr,
                            const int32_t *col, const double *val,
                            const int32_t *sizes, const int32_t *displs,
                            const int64_t *core_iluts);

/* The same hand-over with the sparse core Hamiltonian built ON THE DEVICE instead of by the host: replaces
 * calc_determ_hamil_sparse / calc_determ_hamil_opt (src/sparse_arrays.F90:426-572,
 * src/fast_determ_hamil.F90:892-1548) for this rank's rows.  sizes / displs / core_iluts as in
 * neci_gpu_set_core_space (the whole core space, rank-major, nifd+1 words per determinant); every row holds its
 * non-zero off-diagonal elements in ascending column order and H_ii - Hii as its last entry; Hii is
 * neci_gpu_config.hii.  The rows never cross the host link.  *nnz_out = number of stored elements of this rank.
 * The walker list must already hold this rank's core determinants (neci_gpu_upload_walkers).               */
int neci_gpu_build_core_space(neci_gpu_engine *e, const int32_t *sizes, const int32_t *displs,
                              const int64_t *core_iluts, int64_t *nnz_out);

/* Copies this rank's sparse core Hamiltonian out (write_core_space-style dumps, tests): row_ptr[n_local + 1],
 * then col / val with row_ptr[n_local] entries each (either may be NULL to fetch the row offsets only).     */
int neci_gpu_get_core_hamiltonian(neci_gpu_engine *e, int64_t *row_ptr, int32_t *col, double *val);

/* Trial-wavefunction estimator (init_trial_wf, src/trial_wf_gen.F90): the trial space with the trial vector
 * (trial_space / trial_wfs) and the connected space with con_space_vecs = sum_j H_ij psiT_j, i.e. the contents of
 * the two hash tables trial_ht / con_ht that hash_search_trial reads (src/searching.F90:182-223; ntrial_excits = 1).
 * iluts hold nifd+1 words per determinant.  Sets flag bits 2 (trial) / 3 (connected) and current_trial_amps for
 * the resident list and for every determinant inserted afterwards (src/load_balancer.fpp:586-611); each iteration
 * then returns trial_numerator / trial_denom (SumEContrib, src/fcimc_helper.F90:586-648).  A determinant present
 * in both spaces counts as trial (the trial table is searched first).                                           */
int neci_gpu_set_trial_space(neci_gpu_engine *e, int64_t n_trial, const int64_t *trial_iluts,
                             const double *trial_amps, int64_t n_con, const int64_t *con_iluts,
                             const double *con_amps);

/* ---- walker list transfer -------------------------------------------------- */
/* CurrentDets(0:NIfTot, 1:n) + global_determinant_data rows diagH (= H_ii - Hii)
 * and offdiagH (src/global_det_data.fpp:166-175).  If gdata_diag/offdiag are
 * NULL the engine recomputes them (get_diagonal_matel / get_off_diagonal_matel,
 * src/matel_getter.F90:30-105).                                               */
int neci_gpu_upload_walkers(neci_gpu_engine *e, const int64_t *current_dets, int64_t n,
                            const double *gdata_diag, const double *gdata_offdiag);
/* Returns the list length (TotWalkers, holes included) in *n; arrays must hold
 * max_walkers records.  Any output pointer may be NULL.                       */
int neci_gpu_download_walkers(neci_gpu_engine *e, int64_t *current_dets, int64_t *n,
                              double *gdata_diag, double *gdata_offdiag);

/* POPSFILE gather (WriteToPopsfileParOneArr / write_pops_det, src/Popsfile.F90:1622-1949,2054-2107;
 * write_walkers, src/hdf5_popsfile.F90:756): the occupied determinants with |sign| > min_weight
 * (binarypops_min_weight), compacted on the device in list order.  A record is det(0:NIfD), sign, flags -- the
 * body of one binary POPSFILE record -- followed in gdata_* by its global_determinant_data rows.  Call with
 * dets_out = NULL to get the count first.  Reading a POPSFILE back is neci_gpu_upload_walkers (it re-hashes and,
 * without gdata, recomputes H_ii / H_0i as ReadFromPopsfile's caller does).                                   */
int neci_gpu_download_occupied(neci_gpu_engine *e, double min_weight, int64_t *dets_out, int64_t *n,
                               double *gdata_diag, double *gdata_offdiag);

/* ---- the hot path ----------------------------------------------------------- */
/* One FCIQMC iteration == the body of PerformFCIMCycPar.  stats_out receives
 * NECI_ST_COUNT doubles: this rank's accumulators.                            */
int neci_gpu_iterate(neci_gpu_engine *e, double tau, double diag_sft, int64_t iter,
                     double *stats_out);
/* The same iteration for a host that keeps CurrentDets authoritative: uploads
 * the list (n records, gdata rows), iterates, and downloads the list back into
 * the same buffers (which must hold max_walkers records); *n is updated.       */
int neci_gpu_iterate_host(neci_gpu_engine *e, int64_t *current_dets, int64_t *n,
                          double *gdata_diag, double *gdata_offdiag,
                          double tau, double diag_sft, int64_t iter, double *stats_out);

/* Annihilation of a fixed spawned list (AoS records of niftot+1 words, already
 * on their owner rank): CompressSpawnedList + AnnihilateSpawnedParts +
 * CalcHashTableStats (src/Annihilation.F90:249-515,965-1352;
 * src/load_balancer.fpp:646-805) against the resident main list.              */
int neci_gpu_annihilate(neci_gpu_engine *e, const int64_t *spawned_parts, int64_t n_spawned,
                        int64_t iter, double *stats_out);

/* Multi-rank wiring: 128-byte NCCL unique id from rank 0 (broadcast by the
 * host with MPIBCast / torch.distributed), then collective communicator init. */
int neci_gpu_nccl_unique_id(uint8_t id_out[128]);
int neci_gpu_nccl_init(neci_gpu_engine *e, const uint8_t id[128]);

/* Spawn exchange over NVLink peer memory instead of NCCL send/recv (same result, no host round trip for the
 * counts): every rank creates its inbox and returns a 64-byte CUDA IPC handle; the host all-gathers the handles
 * (MPIAllGather / torch.distributed) and every rank opens its peers' inboxes.  One process per GPU, all on one
 * NVLink/NVSwitch box.  When this has been called neci_gpu_iterate / neci_gpu_rebalance use it for
 * SendProcNewParts (src/Annihilation.F90:150-247); otherwise they use the NCCL communicator.  In neci_gpu_iterate the
 * spawning kernels then route their spawns themselves (DetermineDetNode) and store them into the owners' inboxes
 * while they run; the exchange proper is a mailbox hand-shake.  A semi-stochastic run on several ranks needs
 * neci_gpu_nccl_init as well (the core vector is gathered with NCCL); nranks <= 64.                             */
int neci_gpu_p2p_handle(neci_gpu_engine *e, uint8_t handle_out[64]);
int neci_gpu_p2p_open(neci_gpu_engine *e, const uint8_t *handles /* nranks x 64 bytes, rank order */);

/* adjust_load_balance (src/load_balancer.fpp:178-351): install a new
 * LoadBalanceMapping and move the determinants of the listed blocks
 * (collective over all ranks).                                                */
int neci_gpu_rebalance(neci_gpu_engine *e, const int32_t *new_mapping);
/* Per-block walker counts on this rank (input of the greedy balancer,
 * src/load_balancer.fpp:216-235).                                             */
int neci_gpu_block_populations(neci_gpu_engine *e, double *block_parts);

/* ---- measurement helpers ----------------------------------------------------- */
/* CUDA events on the engine's own stream (torch.cuda.Event would only see
 * torch's current stream): start synchronises the stream and records; stop
 * records, waits and returns the device time between the two in ms.  This is
 * the device-side counterpart of the reference's set_timer/halt_timer pair
 * around the iteration (src/FciMCPar.F90:1226, src/lib/timing.F90:122-223).  */
int neci_gpu_timer_start(neci_gpu_engine *e);
int neci_gpu_timer_stop(neci_gpu_engine *e, double *ms_out);
/* Number of kernels this engine has launched since neci_gpu_init.              */
int64_t neci_gpu_launch_count(const neci_gpu_engine *e);
/* Page-locked host memory for CurrentDets / global_determinant_data when the
 * host keeps the list authoritative (neci_gpu_iterate_host): the Fortran host
 * maps it with c_f_pointer instead of ALLOCATE (fcimc_initialisation.fpp:1650-1763). */
int neci_gpu_alloc_host(int64_t bytes, void **out);
int neci_gpu_free_host(void *p);

/* Benchmark set-up ("frozen synthetic list"): replaces the resident list by this rank's share of
 * a global list of n_dets_total uniformly random determinants (n_alpha of nBasis/2 spatial orbitals for the alpha
 * electrons, n_beta for the beta electrons; signs +-round(1 + Exp(1))), generated on the device as a function of
 * (seed, candidate index) alone -- the global list does not depend on the number of ranks; ownership by
 * DetermineDetNode (src/load_balance_calcnodes.F90:25-117).  H_ii / H_0i are computed as AddNewHashDet does.  No
 * counterpart in the reference (its runs grow their list); lists of 1e8-1e9 walkers cannot be built on the host in
 * the time of a benchmark.  n_local_out: records taken in on this rank (duplicates of a determinant are holes). */
int neci_gpu_synthetic_list(neci_gpu_engine *e, int64_t n_dets_total, uint64_t seed, int64_t *n_local_out);

/* ---- batch probes (parity tests call the same device functions) ----------- */
/* get_det_block / DetermineDetNode (src/load_balance_calcnodes.F90:25-117):
 * block is 1-based, node 0-based.                                             */
int neci_gpu_probe_det_node(neci_gpu_engine *e, int64_t n, const int64_t *iluts,
                            int32_t *block_out, int32_t *node_out);
/* get_helement(nI, nJ) for arbitrary pairs (ic = 0,1,2; > 2 gives 0).         */
int neci_gpu_probe_helement(neci_gpu_engine *e, int64_t n, const int64_t *iluts_i,
                            const int64_t *iluts_j, double *hel_out);
/* generate_excitation + get_spawn_helement for (det, attempt) pairs with the
 * engine's counter-based RNG stream of iteration `iter`:
 * ilut_j (0 words => null excitation), ic, ex[4] = {src1,src2,tgt1,tgt2},
 * parity, pgen, hel.                                                          */
int neci_gpu_probe_gen_excit(neci_gpu_engine *e, int64_t n, const int64_t *iluts,
                             const int32_t *attempt, int64_t iter,
                             int64_t *ilut_j_out, int32_t *ic_out, int32_t *ex_out,
                             int32_t *parity_out, double *pgen_out, double *hel_out);

#ifdef __cplusplus
}
#endif
#endif /* NECI_GPU_H */

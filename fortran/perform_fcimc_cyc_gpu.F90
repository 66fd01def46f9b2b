! perform_fcimc_cyc_gpu.F90 -- the Fortran side of the drop-in: a replacement body for PerformFCIMCycPar
! (src/FciMCPar.F90:1177-1920) that runs the iteration on the B200 engine, plus the hand-over routines the host
! calls once after its own set-up.  Module variables are the reference's own (FciMCData, CalcData, SystemData, ...);
! nothing here re-implements host logic.  No Fortran compiler exists in the image this repository is built in, so
! this file is checked syntactically against include/neci_gpu.h by tests/test_fortran_interface_cpu.py (every
! neci_gpu_* call: symbol exists, argument count matches the interface) and is compiled for the first time on the
! maintainer's side:  add both files of this directory to src/ and link libneci_gpu.so.
!
! What PerformFCIMCycPar leaves behind for its callers, and where this file provides it:
!   per-rank accumulators of communicate_estimates  (src/fcimc_iter_utilities.F90:466-535)   -> gpu_scatter_stats
!   SumWalkersCyc / SumWalkersOut += TotParts       (end_iter_stats, src/fcimc_helper.F90:1489-1499,
!                                                    called BEFORE annihilation, src/FciMCPar.F90:1810)
!   iter_data%update_growth / %update_iters         (update_iter_data, src/fcimc_iter_utilities.F90:1442-1451,
!                                                    src/FciMCPar.F90:1895)
!   TotWalkers, TotParts, HolesInList, iHighestPop  (DirectAnnihilation / CalcHashTableStats)
!   err /= 0 on any overflow                        (src/FciMCPar.F90:1764,1851; the caller stop_all's at :508)
module perform_fcimc_cyc_gpu
    use, intrinsic :: iso_c_binding
    use neci_gpu_interface
    use constants, only: dp, int64, n_int, stdout
    use FciMCData, only: fcimc_iter_data, NoBorn, NoDied, Annihilated, NoAborted, NoRemoved, SpawnFromSing, &
                         Acceptances, HFCyc, HFOut, NoatHF, NoatDoubs, ENumCyc, ENumCycAbs, ENumOut, SumENum, SumNoatHF, &
                         InitsENumCyc, NoInitDets, NoNonInitDets, NoInitWalk, NoNonInitWalk, NoAddedInitiators, &
                         nValidExcits, nInvalidExcits, bloom_count, bloom_sizes, max_cyc_spawn, TotParts, TotWalkers, &
                         norm_psi_squared, norm_semistoch_squared, HolesInList, iHighestPop, SumWalkersCyc, &
                         SumWalkersOut, CurrentDets, MaxWalkersPart, MaxSpawned, iLutRef, Hii, Iter, trial_numerator, &
                         trial_denom, trial_num_inst, trial_denom_inst, init_trial_numerator, init_trial_denom, &
                         nspawned, iStartFreeSlot, iEndFreeSlot
    use CalcData, only: tau, DiagSft, tTruncInitiator, InitiatorWalkNo, tAllRealCoeff, tRealSpawnCutoff, &
                        RealSpawnCutoff, OccupiedThresh, AvMCExcits, tInitCoherentRule, tSemiStochastic, &
                        t_core_inits, tTrialWavefunction, tDeathBeforeComms
    use tau_main, only: tau_search_method, possible_tau_search_methods, max_death_cpt
    use tau_search_conventional, only: tau_search_stats
    implicit none
    private
    public :: PerformFCIMCycPar_gpu, gpu_engine_init, gpu_engine_finalize, gpu_upload_current_dets, &
              gpu_download_current_dets, gpu_handle

    type(c_ptr), save :: gpu_handle = c_null_ptr

contains

    ! ------------------------------------------------------------------------------------------------------------
    ! Drop-in for `call PerformFciMCycPar(iter_data_fciqmc, err)` (src/FciMCPar.F90:507): same signature.
    ! ------------------------------------------------------------------------------------------------------------
    subroutine PerformFCIMCycPar_gpu(iter_data, err)
        type(fcimc_iter_data), intent(inout) :: iter_data
        integer, intent(out) :: err
        real(c_double) :: st(0:NECI_ST_COUNT - 1)
        integer(c_int) :: rc

        ! end_iter_stats (src/fcimc_helper.F90:1489-1499) runs before annihilation in the reference, i.e. on the
        ! TotParts the walker loop of this iteration saw: the value still held by the module variable here.
        SumWalkersCyc = SumWalkersCyc + TotParts
        SumWalkersOut = SumWalkersOut + TotParts

        rc = neci_gpu_iterate(gpu_handle, real(tau, c_double), real(DiagSft(1), c_double), int(Iter, c_int64_t), st)
        err = int(rc)
        if (err /= 0) then
            call gpu_report_error('neci_gpu_iterate')
            return                                           ! FciMCPar.F90:508 stops the run
        end if
        call gpu_scatter_stats(st, iter_data)

        ! update_iter_data (src/fcimc_iter_utilities.F90:1442-1451): growth of this update cycle and its length,
        ! read by update_shift (:1156-1158) and by the growth check at :841
        iter_data%update_growth = iter_data%update_growth + iter_data%nborn - iter_data%ndied - iter_data%nannihil &
                                  - iter_data%naborted - iter_data%nremoved
        iter_data%update_iters = iter_data%update_iters + 1
    end subroutine PerformFCIMCycPar_gpu

    ! Per-rank accumulators -> the module variables communicate_estimates reduces (src/fcimc_iter_utilities.F90:466-535)
    subroutine gpu_scatter_stats(st, iter_data)
        real(c_double), intent(in) :: st(0:NECI_ST_COUNT - 1)
        type(fcimc_iter_data), intent(inout) :: iter_data

        NoBorn(1) = NoBorn(1) + st(NECI_ST_NOBORN)
        NoDied(1) = NoDied(1) + st(NECI_ST_NODIED)
        Annihilated(1) = Annihilated(1) + st(NECI_ST_ANNIHILATED)
        NoAborted(1) = NoAborted(1) + st(NECI_ST_NOABORTED)
        NoRemoved(1) = NoRemoved(1) + st(NECI_ST_NOREMOVED)
        SpawnFromSing(1) = SpawnFromSing(1) + st(NECI_ST_SPAWNFROMSING)
        Acceptances(1) = Acceptances(1) + st(NECI_ST_ACCEPTANCES)
        HFCyc(1) = HFCyc(1) + st(NECI_ST_HFCYC)
        HFOut(1) = HFOut(1) + st(NECI_ST_HFCYC)
        SumNoatHF(1) = SumNoatHF(1) + st(NECI_ST_HFCYC)
        NoatHF(1) = st(NECI_ST_INSTNOATHF)
        NoatDoubs(1) = NoatDoubs(1) + st(NECI_ST_NOATDOUBS)
        ENumCyc(1) = ENumCyc(1) + st(NECI_ST_ENUMCYC)
        ENumOut(1) = ENumOut(1) + st(NECI_ST_ENUMCYC)
        SumENum(1) = SumENum(1) + st(NECI_ST_ENUMCYC)
        ENumCycAbs(1) = ENumCycAbs(1) + st(NECI_ST_ENUMCYCABS)
        InitsENumCyc(1) = InitsENumCyc(1) + st(NECI_ST_INITSENUMCYC)
        NoInitDets(1) = int(st(NECI_ST_NOINITDETS), int64)
        NoNonInitDets(1) = int(st(NECI_ST_NONONINITDETS), int64)
        NoInitWalk(1) = st(NECI_ST_NOINITWALK)
        NoNonInitWalk(1) = st(NECI_ST_NONONINITWALK)
        NoAddedInitiators(1) = NoAddedInitiators(1) + int(st(NECI_ST_NOADDEDINITIATORS), int64)
        nValidExcits = nValidExcits + int(st(NECI_ST_NVALIDEXCITS), int64)
        nInvalidExcits = nInvalidExcits + int(st(NECI_ST_NINVALIDEXCITS), int64)
        bloom_count(1) = bloom_count(1) + int(st(NECI_ST_BLOOM_COUNT_1))
        bloom_count(2) = bloom_count(2) + int(st(NECI_ST_BLOOM_COUNT_2))
        bloom_sizes(1) = max(bloom_sizes(1), st(NECI_ST_BLOOM_SIZE_1))
        bloom_sizes(2) = max(bloom_sizes(2), st(NECI_ST_BLOOM_SIZE_2))
        max_cyc_spawn = max(max_cyc_spawn, st(NECI_ST_MAX_CYC_SPAWN))
        nspawned = nspawned + int(st(NECI_ST_NSPAWNED_SENT), int64)

        ! what DirectAnnihilation / CalcHashTableStats leave behind (src/load_balancer.fpp:646-805)
        TotParts(1) = st(NECI_ST_TOTPARTS)
        norm_psi_squared(1) = st(NECI_ST_NORM_PSI_SQ)
        norm_semistoch_squared(1) = st(NECI_ST_NORM_SEMISTOCH_SQ)
        TotWalkers = int(st(NECI_ST_TOTWALKERS), int64)
        HolesInList = int(st(NECI_ST_HOLESINLIST))
        iHighestPop = int(st(NECI_ST_HIGHEST_POP))
        ! the free-slot list lives on the device; the host's copy is kept empty
        iStartFreeSlot = 1
        iEndFreeSlot = 0

        if (tTrialWavefunction) then                          ! SumEContrib, src/fcimc_helper.F90:586-648
            trial_numerator(1) = trial_numerator(1) + st(NECI_ST_TRIAL_NUMERATOR)
            trial_denom(1) = trial_denom(1) + st(NECI_ST_TRIAL_DENOM)
            trial_num_inst(1) = st(NECI_ST_TRIAL_NUMERATOR)
            trial_denom_inst(1) = st(NECI_ST_TRIAL_DENOM)
            init_trial_numerator(1) = init_trial_numerator(1) + st(NECI_ST_INIT_TRIAL_NUMERATOR)
            init_trial_denom(1) = init_trial_denom(1) + st(NECI_ST_INIT_TRIAL_DENOM)
        end if

        if (tau_search_method /= possible_tau_search_methods%OFF) then
            ! log_spawn_magnitude (src/tau/tau_search_conventional.F90:138-260) and log_death_magnitude
            ! (src/tau/tau_main.F90:198-207); tau_search / update_tau then run unchanged on these
            associate(t_s => tau_search_stats)
                t_s%gamma_sing = max(t_s%gamma_sing, st(NECI_ST_TAU_GAMMA_SING))
                t_s%gamma_doub = max(t_s%gamma_doub, st(NECI_ST_TAU_GAMMA_DOUB))
                t_s%gamma_par = max(t_s%gamma_par, st(NECI_ST_TAU_GAMMA_PAR))
                t_s%gamma_opp = max(t_s%gamma_opp, st(NECI_ST_TAU_GAMMA_OPP))
                t_s%cnt_sing = t_s%cnt_sing + int(st(NECI_ST_TAU_CNT_SING))
                t_s%cnt_doub = t_s%cnt_doub + int(st(NECI_ST_TAU_CNT_DOUB))
                t_s%cnt_par = t_s%cnt_par + int(st(NECI_ST_TAU_CNT_PAR))
                t_s%cnt_opp = t_s%cnt_opp + int(st(NECI_ST_TAU_CNT_OPP))
                t_s%enough_sing = t_s%cnt_sing > 50
                t_s%enough_par = t_s%cnt_par > 50
                t_s%enough_opp = t_s%cnt_opp > 50
                t_s%enough_doub = (t_s%cnt_doub > 50) .or. (t_s%enough_par .and. t_s%enough_opp)
            end associate
            max_death_cpt = max(max_death_cpt, st(NECI_ST_TAU_MAX_DEATH_CPT))
        end if

        iter_data%nborn(1) = iter_data%nborn(1) + st(NECI_ST_NOBORN)
        iter_data%ndied(1) = iter_data%ndied(1) + st(NECI_ST_NODIED)
        iter_data%nannihil(1) = iter_data%nannihil(1) + st(NECI_ST_ANNIHILATED)
        iter_data%naborted(1) = iter_data%naborted(1) + st(NECI_ST_NOABORTED)
        iter_data%nremoved(1) = iter_data%nremoved(1) + st(NECI_ST_NOREMOVED)
    end subroutine gpu_scatter_stats

    ! ------------------------------------------------------------------------------------------------------------
    ! End of InitFCIMCCalcPar (src/FciMCPar.F90:256): configuration + system tables + the initial walker list.
    ! ------------------------------------------------------------------------------------------------------------
    subroutine gpu_engine_init(device, seed, random_orb_index, random_hash2, lb_mapping0, system_type, err)
        use SystemData, only: nel, nBasis, nOccAlpha, nOccBeta, tNoBrillouin, tExch, tHPHF, ECore
        use bit_rep_data, only: NIfD, NIfTot
        use Parallel_neci, only: nNodes, iProcIndex
        use load_balance_calcnodes, only: balance_blocks
        integer, intent(in) :: device, system_type
        integer(int64), intent(in) :: seed
        integer(c_int32_t), intent(in), target :: random_orb_index(:), random_hash2(:), lb_mapping0(:)   ! 0-based ranks
        integer, intent(out) :: err
        type(neci_gpu_config) :: cfg
        integer(c_int64_t), target :: ilut_ref_c(0:NIfD)

        ilut_ref_c(0:NIfD) = int(iLutRef(0:NIfD, 1), c_int64_t)
        cfg%nel = nel; cfg%nbasis = nBasis; cfg%nifd = NIfD; cfg%niftot = NIfTot
        cfg%nocc_alpha = nOccAlpha; cfg%nocc_beta = nOccBeta
        cfg%nranks = nNodes; cfg%rank = iProcIndex; cfg%device = device; cfg%balance_blocks = balance_blocks
        cfg%max_walkers = int(MaxWalkersPart, c_int64_t); cfg%max_spawned = int(MaxSpawned, c_int64_t)
        cfg%system_type = system_type
        cfg%t_trunc_initiator = merge(1, 0, tTruncInitiator)
        cfg%t_all_real_coeff = merge(1, 0, tAllRealCoeff)
        cfg%t_real_spawn_cutoff = merge(1, 0, tRealSpawnCutoff)
        cfg%t_death_before_comms = merge(1, 0, tDeathBeforeComms)
        cfg%t_init_coherent_rule = merge(1, 0, tInitCoherentRule)
        cfg%t_no_brillouin = merge(1, 0, tNoBrillouin)
        cfg%t_exch = merge(1, 0, tExch)
        cfg%t_semi_stochastic = merge(1, 0, tSemiStochastic)
        cfg%t_core_inits = merge(1, 0, t_core_inits)
        cfg%t_tau_search = merge(1, 0, tau_search_method /= possible_tau_search_methods%OFF)
        cfg%t_consider_par_bias = 0
        cfg%t_hphf = merge(1, 0, tHPHF)
        cfg%reserved0 = 0
        cfg%initiator_walk_no = InitiatorWalkNo; cfg%real_spawn_cutoff = RealSpawnCutoff
        cfg%occupied_thresh = OccupiedThresh; cfg%av_mc_excits = AvMCExcits
        cfg%hii = Hii; cfg%ecore = ECore
        cfg%seed = int(seed, c_int64_t)
        cfg%random_orb_index = c_loc(random_orb_index(1)); cfg%random_hash2 = c_loc(random_hash2(1))
        cfg%load_balance_mapping = c_loc(lb_mapping0(1)); cfg%ilut_ref = c_loc(ilut_ref_c(0))
        err = int(neci_gpu_init(cfg, gpu_handle))
        if (err /= 0) call gpu_report_error('neci_gpu_init')
    end subroutine gpu_engine_init

    ! After InitFCIMC_HF / ReadFromPopsfile: CurrentDets(0:NIfTot, 1:TotWalkers); H_ii - Hii and H_0i are recomputed
    ! on the device (get_diagonal_matel / get_off_diagonal_matel) when no global_determinant_data is passed.
    subroutine gpu_upload_current_dets(err)
        integer, intent(out) :: err
        integer(c_int64_t), pointer :: flat(:)
        call c_f_pointer(c_loc(CurrentDets), flat, [size(CurrentDets, kind=int64)])
        err = int(neci_gpu_upload_walkers(gpu_handle, flat, int(TotWalkers, c_int64_t), c_null_ptr, c_null_ptr))
        if (err /= 0) call gpu_report_error('neci_gpu_upload_walkers')
    end subroutine gpu_upload_current_dets

    ! Before WriteToPopsfileParOneArr, PrintHighPops, core-space re-selection, end of run.
    subroutine gpu_download_current_dets(err)
        integer, intent(out) :: err
        integer(c_int64_t) :: n
        err = int(neci_gpu_download_walkers(gpu_handle, c_loc(CurrentDets), n, c_null_ptr, c_null_ptr))
        if (err /= 0) then
            call gpu_report_error('neci_gpu_download_walkers')
            return
        end if
        TotWalkers = int(n, int64)
    end subroutine gpu_download_current_dets

    ! DeallocFCIMCMemPar
    subroutine gpu_engine_finalize()
        integer(c_int) :: rc
        if (c_associated(gpu_handle)) rc = neci_gpu_finalize(gpu_handle)
        gpu_handle = c_null_ptr
    end subroutine gpu_engine_finalize

    subroutine gpu_report_error(where)
        character(*), intent(in) :: where
        type(c_ptr) :: msg
        character(kind=c_char), pointer :: txt(:)
        integer :: n
        msg = neci_gpu_last_error(gpu_handle)
        if (.not. c_associated(msg)) return
        call c_f_pointer(msg, txt, [512])
        n = 1
        do while (n < 512 .and. txt(n) /= c_null_char)
            n = n + 1
        end do
        write(stdout, '(a,a,a,512a1)') ' GPU engine: ', where, ' failed: ', txt(1:n - 1)
    end subroutine gpu_report_error

end module perform_fcimc_cyc_gpu

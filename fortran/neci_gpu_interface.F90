! neci_gpu_interface.F90 -- ISO_C_BINDING image of include/neci_gpu.h, one `bind(c)` interface per exported symbol,
! following the reference's own convention for C code (src/lib/dSFMT_interface.F90:31-42,
! docs/pages/02_dev_doc/12_interfacing_C.md): scalars by value, arrays by reference, integer status returned.
!
! Every symbol of libneci_gpu.so is bound here.  tests/test_fortran_interface_cpu.py parses this file and
! include/neci_gpu.h and checks, symbol by symbol, the argument count, order, C type and by-value / by-reference
! passing, the field order of neci_gpu_config and the NECI_ST_* / NECI_FLAG_* / NECI_SYS_* constants.
!
! Arrays are declared assumed-size with intent; a C pointer that may be NULL (optional output) is declared
! `type(c_ptr), value` and the caller passes c_loc(array) or c_null_ptr.
module neci_gpu_interface
    use, intrinsic :: iso_c_binding
    implicit none

    ! ---- flag bits of the last ilut word (src/bit_rep_data.F90:87-119) ----
    integer(c_int), parameter :: NECI_FLAG_REMOVED = 0
    integer(c_int), parameter :: NECI_FLAG_DETERM_PARENT = 1
    integer(c_int), parameter :: NECI_FLAG_TRIAL = 2
    integer(c_int), parameter :: NECI_FLAG_CONNECTED = 3
    integer(c_int), parameter :: NECI_FLAG_INITIATOR = 13
    integer(c_int), parameter :: NECI_FLAG_STATIC_INIT = 16
    integer(c_int), parameter :: NECI_FLAG_DETERMINISTIC = 19

    ! ---- system selector ----
    integer(c_int), parameter :: NECI_SYS_FCIDUMP_PCHB = 1
    integer(c_int), parameter :: NECI_SYS_HUBBARD_RS = 2
    integer(c_int), parameter :: NECI_SYS_HUBBARD_K = 3

    ! ---- enum neci_stat_index (0-based positions in the statistics vector) ----
    integer(c_int), parameter :: NECI_ST_NOBORN = 0
    integer(c_int), parameter :: NECI_ST_NODIED = 1
    integer(c_int), parameter :: NECI_ST_ANNIHILATED = 2
    integer(c_int), parameter :: NECI_ST_NOABORTED = 3
    integer(c_int), parameter :: NECI_ST_NOREMOVED = 4
    integer(c_int), parameter :: NECI_ST_SPAWNFROMSING = 5
    integer(c_int), parameter :: NECI_ST_ACCEPTANCES = 6
    integer(c_int), parameter :: NECI_ST_HFCYC = 7
    integer(c_int), parameter :: NECI_ST_NOATDOUBS = 8
    integer(c_int), parameter :: NECI_ST_ENUMCYC = 9
    integer(c_int), parameter :: NECI_ST_ENUMCYCABS = 10
    integer(c_int), parameter :: NECI_ST_INITSENUMCYC = 11
    integer(c_int), parameter :: NECI_ST_NOINITDETS = 12
    integer(c_int), parameter :: NECI_ST_NONONINITDETS = 13
    integer(c_int), parameter :: NECI_ST_NOINITWALK = 14
    integer(c_int), parameter :: NECI_ST_NONONINITWALK = 15
    integer(c_int), parameter :: NECI_ST_NOADDEDINITIATORS = 16
    integer(c_int), parameter :: NECI_ST_NVALIDEXCITS = 17
    integer(c_int), parameter :: NECI_ST_NINVALIDEXCITS = 18
    integer(c_int), parameter :: NECI_ST_BLOOM_COUNT_1 = 19
    integer(c_int), parameter :: NECI_ST_BLOOM_COUNT_2 = 20
    integer(c_int), parameter :: NECI_ST_MAX_CYC_SPAWN = 21
    integer(c_int), parameter :: NECI_ST_BLOOM_SIZE_1 = 22
    integer(c_int), parameter :: NECI_ST_BLOOM_SIZE_2 = 23
    integer(c_int), parameter :: NECI_ST_TAU_GAMMA_SING = 24
    integer(c_int), parameter :: NECI_ST_TAU_GAMMA_DOUB = 25
    integer(c_int), parameter :: NECI_ST_TAU_GAMMA_PAR = 26
    integer(c_int), parameter :: NECI_ST_TAU_GAMMA_OPP = 27
    integer(c_int), parameter :: NECI_ST_TAU_MAX_DEATH_CPT = 28
    integer(c_int), parameter :: NECI_ST_TOTPARTS = 29
    integer(c_int), parameter :: NECI_ST_NORM_PSI_SQ = 30
    integer(c_int), parameter :: NECI_ST_NORM_SEMISTOCH_SQ = 31
    integer(c_int), parameter :: NECI_ST_INSTNOATHF = 32
    integer(c_int), parameter :: NECI_ST_TOTWALKERS = 33
    integer(c_int), parameter :: NECI_ST_HOLESINLIST = 34
    integer(c_int), parameter :: NECI_ST_NSPAWNED_SENT = 35
    integer(c_int), parameter :: NECI_ST_NSPAWNED_RECV = 36
    integer(c_int), parameter :: NECI_ST_NSPAWNED_MERGED = 37
    integer(c_int), parameter :: NECI_ST_NINSERTED = 38
    integer(c_int), parameter :: NECI_ST_HIGHEST_POP = 39
    integer(c_int), parameter :: NECI_ST_TRIAL_NUMERATOR = 40
    integer(c_int), parameter :: NECI_ST_TRIAL_DENOM = 41
    integer(c_int), parameter :: NECI_ST_INIT_TRIAL_NUMERATOR = 42
    integer(c_int), parameter :: NECI_ST_INIT_TRIAL_DENOM = 43
    integer(c_int), parameter :: NECI_ST_TAU_CNT_SING = 44
    integer(c_int), parameter :: NECI_ST_TAU_CNT_DOUB = 45
    integer(c_int), parameter :: NECI_ST_TAU_CNT_PAR = 46
    integer(c_int), parameter :: NECI_ST_TAU_CNT_OPP = 47
    integer(c_int), parameter :: NECI_ST_ERR_FLAGS = 48
    integer(c_int), parameter :: NECI_ST_TIME_SPAWN_MS = 49
    integer(c_int), parameter :: NECI_ST_TIME_COMM_MS = 50
    integer(c_int), parameter :: NECI_ST_TIME_ANNIHIL_MS = 51
    integer(c_int), parameter :: NECI_ST_TIME_DETERM_MS = 52
    integer(c_int), parameter :: NECI_ST_COUNT = 53

    ! ---- struct neci_gpu_config (same field order and types as the header) ----
    type, bind(c) :: neci_gpu_config
        integer(c_int32_t) :: nel
        integer(c_int32_t) :: nbasis
        integer(c_int32_t) :: nifd
        integer(c_int32_t) :: niftot
        integer(c_int32_t) :: nocc_alpha
        integer(c_int32_t) :: nocc_beta
        integer(c_int32_t) :: nranks
        integer(c_int32_t) :: rank
        integer(c_int32_t) :: device
        integer(c_int32_t) :: balance_blocks
        integer(c_int64_t) :: max_walkers
        integer(c_int64_t) :: max_spawned
        integer(c_int32_t) :: system_type
        integer(c_int32_t) :: t_trunc_initiator
        integer(c_int32_t) :: t_all_real_coeff
        integer(c_int32_t) :: t_real_spawn_cutoff
        integer(c_int32_t) :: t_death_before_comms
        integer(c_int32_t) :: t_init_coherent_rule
        integer(c_int32_t) :: t_no_brillouin
        integer(c_int32_t) :: t_exch
        integer(c_int32_t) :: t_semi_stochastic
        integer(c_int32_t) :: t_core_inits
        integer(c_int32_t) :: t_tau_search
        integer(c_int32_t) :: t_consider_par_bias
        integer(c_int32_t) :: t_hphf
        integer(c_int32_t) :: reserved0
        real(c_double) :: initiator_walk_no
        real(c_double) :: real_spawn_cutoff
        real(c_double) :: occupied_thresh
        real(c_double) :: av_mc_excits
        real(c_double) :: hii
        real(c_double) :: ecore
        integer(c_int64_t) :: seed
        type(c_ptr) :: random_orb_index
        type(c_ptr) :: random_hash2
        type(c_ptr) :: load_balance_mapping
        type(c_ptr) :: ilut_ref
    end type neci_gpu_config

    interface
        ! ---- lifetime ----------------------------------------------------------------------------------------
        function neci_gpu_init(cfg, handle) result(err) bind(c, name='neci_gpu_init')
            import :: c_int, c_ptr, neci_gpu_config
            type(neci_gpu_config), intent(in) :: cfg
            type(c_ptr), intent(out) :: handle
            integer(c_int) :: err
        end function
        function neci_gpu_finalize(handle) result(err) bind(c, name='neci_gpu_finalize')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
            integer(c_int) :: err
        end function
        function neci_gpu_last_error(handle) result(msg) bind(c, name='neci_gpu_last_error')
            import :: c_ptr
            type(c_ptr), value :: handle
            type(c_ptr) :: msg
        end function

        ! ---- read-only system tables ---------------------------------------------------------------------------
        function neci_gpu_set_system_fcidump(handle, umat, n_umat, tmat2d) result(err) bind(c, name='neci_gpu_set_system_fcidump')
            import :: c_int, c_ptr, c_double, c_int64_t
            type(c_ptr), value :: handle
            real(c_double), intent(in) :: umat(*)
            integer(c_int64_t), value :: n_umat
            real(c_double), intent(in) :: tmat2d(*)
            integer(c_int) :: err
        end function
        function neci_gpu_set_pchb(handle, n_spat, ij_max, ab_max, probs, bias, alias, p_exch, tgt_orbs, p_singles, &
                                   p_doubles, p_parallel, n_classes, class_of_spinorb) result(err) bind(c, name='neci_gpu_set_pchb')
            import :: c_int, c_ptr, c_double, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: n_spat
            integer(c_int32_t), value :: ij_max
            integer(c_int32_t), value :: ab_max
            real(c_double), intent(in) :: probs(*)
            real(c_double), intent(in) :: bias(*)
            integer(c_int32_t), intent(in) :: alias(*)
            real(c_double), intent(in) :: p_exch(*)
            integer(c_int32_t), intent(in) :: tgt_orbs(*)
            real(c_double), value :: p_singles
            real(c_double), value :: p_doubles
            real(c_double), value :: p_parallel
            integer(c_int32_t), value :: n_classes
            integer(c_int32_t), intent(in) :: class_of_spinorb(*)
            integer(c_int) :: err
        end function
        function neci_gpu_set_pchb_particles(handle, mode, p_first, p_second) result(err) bind(c, name='neci_gpu_set_pchb_particles')
            import :: c_int, c_ptr, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: mode
            real(c_double), intent(in) :: p_first(*)
            real(c_double), intent(in) :: p_second(*)
            integer(c_int) :: err
        end function
        function neci_gpu_set_excit_probs(handle, p_singles, p_doubles, p_parallel) result(err) bind(c, name='neci_gpu_set_excit_probs')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: p_singles
            real(c_double), value :: p_doubles
            real(c_double), value :: p_parallel
            integer(c_int) :: err
        end function
        function neci_gpu_set_system_hubbard_rs(handle, max_neigh, neighbours, tmat2d, uhub) result(err) &
                bind(c, name='neci_gpu_set_system_hubbard_rs')
            import :: c_int, c_ptr, c_double, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: max_neigh
            integer(c_int32_t), intent(in) :: neighbours(*)
            real(c_double), intent(in) :: tmat2d(*)
            real(c_double), value :: uhub
            integer(c_int) :: err
        end function
        function neci_gpu_set_system_hubbard_k(handle, n_k, ksum, kdiff, eps_k, u_over_n) result(err) &
                bind(c, name='neci_gpu_set_system_hubbard_k')
            import :: c_int, c_ptr, c_double, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: n_k
            integer(c_int32_t), intent(in) :: ksum(*)
            integer(c_int32_t), intent(in) :: kdiff(*)
            real(c_double), intent(in) :: eps_k(*)
            real(c_double), value :: u_over_n
            integer(c_int) :: err
        end function
        function neci_gpu_set_core_space(handle, n_local, row_ptr, col, val, sizes, displs, core_iluts) result(err) &
                bind(c, name='neci_gpu_set_core_space')
            import :: c_int, c_ptr, c_double, c_int32_t, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: n_local
            integer(c_int64_t), intent(in) :: row_ptr(*)
            integer(c_int32_t), intent(in) :: col(*)
            real(c_double), intent(in) :: val(*)
            integer(c_int32_t), intent(in) :: sizes(*)
            integer(c_int32_t), intent(in) :: displs(*)
            integer(c_int64_t), intent(in) :: core_iluts(*)
            integer(c_int) :: err
        end function
        function neci_gpu_build_core_space(handle, sizes, displs, core_iluts, nnz) result(err) bind(c, name='neci_gpu_build_core_space')
            import :: c_int, c_ptr, c_int32_t, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int32_t), intent(in) :: sizes(*)
            integer(c_int32_t), intent(in) :: displs(*)
            integer(c_int64_t), intent(in) :: core_iluts(*)
            integer(c_int64_t), intent(out) :: nnz
            integer(c_int) :: err
        end function
        function neci_gpu_get_core_hamiltonian(handle, row_ptr, col, val) result(err) bind(c, name='neci_gpu_get_core_hamiltonian')
            import :: c_int, c_ptr, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), intent(out) :: row_ptr(*)
            type(c_ptr), value :: col                      ! int32_t *, may be c_null_ptr
            type(c_ptr), value :: val                      ! double *, may be c_null_ptr
            integer(c_int) :: err
        end function
        function neci_gpu_set_trial_space(handle, n_trial, trial_iluts, trial_amps, n_con, con_iluts, con_amps) result(err) &
                bind(c, name='neci_gpu_set_trial_space')
            import :: c_int, c_ptr, c_double, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: n_trial
            integer(c_int64_t), intent(in) :: trial_iluts(*)
            real(c_double), intent(in) :: trial_amps(*)
            integer(c_int64_t), value :: n_con
            integer(c_int64_t), intent(in) :: con_iluts(*)
            real(c_double), intent(in) :: con_amps(*)
            integer(c_int) :: err
        end function

        ! ---- walker list transfer --------------------------------------------------------------------------------
        function neci_gpu_upload_walkers(handle, current_dets, n, gdata_diag, gdata_offdiag) result(err) &
                bind(c, name='neci_gpu_upload_walkers')
            import :: c_int, c_ptr, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), intent(in) :: current_dets(*)
            integer(c_int64_t), value :: n
            type(c_ptr), value :: gdata_diag               ! const double *, c_null_ptr = recompute on the device
            type(c_ptr), value :: gdata_offdiag            ! const double *, c_null_ptr = recompute on the device
            integer(c_int) :: err
        end function
        function neci_gpu_download_walkers(handle, current_dets, n, gdata_diag, gdata_offdiag) result(err) &
                bind(c, name='neci_gpu_download_walkers')
            import :: c_int, c_ptr, c_int64_t
            type(c_ptr), value :: handle
            type(c_ptr), value :: current_dets             ! int64_t *, may be c_null_ptr
            integer(c_int64_t), intent(out) :: n
            type(c_ptr), value :: gdata_diag               ! double *, may be c_null_ptr
            type(c_ptr), value :: gdata_offdiag            ! double *, may be c_null_ptr
            integer(c_int) :: err
        end function
        function neci_gpu_download_occupied(handle, min_weight, dets_out, n, gdata_diag, gdata_offdiag) result(err) &
                bind(c, name='neci_gpu_download_occupied')
            import :: c_int, c_ptr, c_double, c_int64_t
            type(c_ptr), value :: handle
            real(c_double), value :: min_weight
            type(c_ptr), value :: dets_out                 ! int64_t *, c_null_ptr = count only
            integer(c_int64_t), intent(out) :: n
            type(c_ptr), value :: gdata_diag               ! double *, may be c_null_ptr
            type(c_ptr), value :: gdata_offdiag            ! double *, may be c_null_ptr
            integer(c_int) :: err
        end function

        ! ---- the hot path --------------------------------------------------------------------------------------------
        function neci_gpu_iterate(handle, tau, diag_sft, iter, stats_out) result(err) bind(c, name='neci_gpu_iterate')
            import :: c_int, c_ptr, c_double, c_int64_t
            type(c_ptr), value :: handle
            real(c_double), value :: tau
            real(c_double), value :: diag_sft
            integer(c_int64_t), value :: iter
            real(c_double), intent(out) :: stats_out(*)
            integer(c_int) :: err
        end function
        function neci_gpu_iterate_host(handle, current_dets, n, gdata_diag, gdata_offdiag, tau, diag_sft, iter, stats_out) &
                result(err) bind(c, name='neci_gpu_iterate_host')
            import :: c_int, c_ptr, c_double, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), intent(inout) :: current_dets(*)
            integer(c_int64_t), intent(inout) :: n
            real(c_double), intent(inout) :: gdata_diag(*)
            real(c_double), intent(inout) :: gdata_offdiag(*)
            real(c_double), value :: tau
            real(c_double), value :: diag_sft
            integer(c_int64_t), value :: iter
            real(c_double), intent(out) :: stats_out(*)
            integer(c_int) :: err
        end function
        function neci_gpu_annihilate(handle, spawned_parts, n_spawned, iter, stats_out) result(err) bind(c, name='neci_gpu_annihilate')
            import :: c_int, c_ptr, c_double, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), intent(in) :: spawned_parts(*)
            integer(c_int64_t), value :: n_spawned
            integer(c_int64_t), value :: iter
            real(c_double), intent(out) :: stats_out(*)
            integer(c_int) :: err
        end function

        ! ---- multi-rank wiring -----------------------------------------------------------------------------------------
        function neci_gpu_nccl_unique_id(id_out) result(err) bind(c, name='neci_gpu_nccl_unique_id')
            import :: c_int, c_int8_t
            integer(c_int8_t), intent(out) :: id_out(128)
            integer(c_int) :: err
        end function
        function neci_gpu_nccl_init(handle, id) result(err) bind(c, name='neci_gpu_nccl_init')
            import :: c_int, c_ptr, c_int8_t
            type(c_ptr), value :: handle
            integer(c_int8_t), intent(in) :: id(128)
            integer(c_int) :: err
        end function
        function neci_gpu_p2p_handle(handle, handle_out) result(err) bind(c, name='neci_gpu_p2p_handle')
            import :: c_int, c_ptr, c_int8_t
            type(c_ptr), value :: handle
            integer(c_int8_t), intent(out) :: handle_out(64)
            integer(c_int) :: err
        end function
        function neci_gpu_p2p_open(handle, handles) result(err) bind(c, name='neci_gpu_p2p_open')
            import :: c_int, c_ptr, c_int8_t
            type(c_ptr), value :: handle
            integer(c_int8_t), intent(in) :: handles(*)    ! nranks x 64 bytes, rank order
            integer(c_int) :: err
        end function
        function neci_gpu_rebalance(handle, new_mapping) result(err) bind(c, name='neci_gpu_rebalance')
            import :: c_int, c_ptr, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), intent(in) :: new_mapping(*)
            integer(c_int) :: err
        end function
        function neci_gpu_block_populations(handle, block_parts) result(err) bind(c, name='neci_gpu_block_populations')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(out) :: block_parts(*)
            integer(c_int) :: err
        end function

        ! ---- measurement helpers ------------------------------------------------------------------------------------------
        function neci_gpu_timer_start(handle) result(err) bind(c, name='neci_gpu_timer_start')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
            integer(c_int) :: err
        end function
        function neci_gpu_timer_stop(handle, ms_out) result(err) bind(c, name='neci_gpu_timer_stop')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(out) :: ms_out
            integer(c_int) :: err
        end function
        function neci_gpu_launch_count(handle) result(n) bind(c, name='neci_gpu_launch_count')
            import :: c_ptr, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t) :: n
        end function
        function neci_gpu_alloc_host(bytes, ptr_out) result(err) bind(c, name='neci_gpu_alloc_host')
            import :: c_int, c_ptr, c_int64_t
            integer(c_int64_t), value :: bytes
            type(c_ptr), intent(out) :: ptr_out
            integer(c_int) :: err
        end function
        function neci_gpu_free_host(p) result(err) bind(c, name='neci_gpu_free_host')
            import :: c_int, c_ptr
            type(c_ptr), value :: p
            integer(c_int) :: err
        end function

        ! benchmark set-up: this rank's share of a frozen synthetic list generated on the device
        function neci_gpu_synthetic_list(handle, n_dets_total, seed, n_local_out) result(err) bind(c, name='neci_gpu_synthetic_list')
            import :: c_int, c_ptr, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: n_dets_total
            integer(c_int64_t), value :: seed
            integer(c_int64_t), intent(out) :: n_local_out
            integer(c_int) :: err
        end function

        ! ---- batch probes -------------------------------------------------------------------------------------------------
        function neci_gpu_probe_det_node(handle, n, iluts, block_out, node_out) result(err) bind(c, name='neci_gpu_probe_det_node')
            import :: c_int, c_ptr, c_int32_t, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: n
            integer(c_int64_t), intent(in) :: iluts(*)
            integer(c_int32_t), intent(out) :: block_out(*)
            integer(c_int32_t), intent(out) :: node_out(*)
            integer(c_int) :: err
        end function
        function neci_gpu_probe_helement(handle, n, iluts_i, iluts_j, hel_out) result(err) bind(c, name='neci_gpu_probe_helement')
            import :: c_int, c_ptr, c_double, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: n
            integer(c_int64_t), intent(in) :: iluts_i(*)
            integer(c_int64_t), intent(in) :: iluts_j(*)
            real(c_double), intent(out) :: hel_out(*)
            integer(c_int) :: err
        end function
        function neci_gpu_probe_gen_excit(handle, n, iluts, attempt, iter, ilut_j_out, ic_out, ex_out, parity_out, pgen_out, &
                                          hel_out) result(err) bind(c, name='neci_gpu_probe_gen_excit')
            import :: c_int, c_ptr, c_double, c_int32_t, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: n
            integer(c_int64_t), intent(in) :: iluts(*)
            integer(c_int32_t), intent(in) :: attempt(*)
            integer(c_int64_t), value :: iter
            integer(c_int64_t), intent(out) :: ilut_j_out(*)
            integer(c_int32_t), intent(out) :: ic_out(*)
            integer(c_int32_t), intent(out) :: ex_out(*)
            integer(c_int32_t), intent(out) :: parity_out(*)
            real(c_double), intent(out) :: pgen_out(*)
            real(c_double), intent(out) :: hel_out(*)
            integer(c_int) :: err
        end function
    end interface

end module neci_gpu_interface

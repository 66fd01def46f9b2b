"""Host-side mirror of what NECI's Fortran host prepares for the hot path.

Everything here produces inputs of exactly the shapes the Fortran host would
pass through the C ABI (include/neci_gpu.h): integral tables, PCHB alias tables,
lattice tables, the hashing tables, the reference determinant and the engine
configuration.  The heavy loops live in ``csrc/host/neci_host.cpp``
(``libneci_host.so``) and ``csrc/host/core_space.cpp`` (semi-stochastic set-up: core space, its sparse
Hamiltonian, DetermineDetNode); this module is the typed Python face of that library.
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB = os.path.join(HERE, "libneci_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB):
            from . import _build
            _build.build_host()
        _lib = C.CDLL(HOST_LIB)
        _lib.neci_host_umat_size.restype = C.c_int64
        _lib.neci_host_update_shift.restype = C.c_double
        _lib.neci_host_sd_space.restype = C.c_int64
        _lib.neci_host_core_ham_build.restype = C.c_void_p
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


@dataclass
class System:
    """A Hamiltonian + excitation generator in the engine's input format."""
    kind: int
    nel: int
    nbasis: int
    nocc_alpha: int
    nocc_beta: int
    ecore: float = 0.0
    t_exch: int = 1
    t_no_brillouin: int = 0
    tables: dict = field(default_factory=dict)
    ref_orbs: np.ndarray = None           # reference determinant, sorted spin orbitals (1-based)

    @property
    def nifd(self):
        return self.nbasis // 64           # src/BitReps.F90:174

    @property
    def nw(self):
        return self.nifd + 1

    @property
    def W(self):
        return self.nifd + 3

    def ilut(self, orbs):
        """EncodeBitDet: sorted orbital list -> occupation words (int64)."""
        w = [0] * self.nw
        for o in orbs:
            o = int(o)                     # numpy int32 would overflow in the shift
            w[(o - 1) // 64] |= 1 << ((o - 1) % 64)
        return np.array(w, dtype=np.uint64).view(np.int64)

    def apply(self, engine):
        """Upload this system's tables through the C ABI."""
        t = self.tables
        if self.kind == capi.SYS_FCIDUMP_PCHB:
            engine.set_system_fcidump(t["umat"], t["tmat"])
            engine.set_pchb(t["pchb"])
            sel = t["pchb"].get("particle_selection", "UNIF-UNIF")
            if sel != "UNIF-UNIF":
                engine.set_pchb_particles({"FULL-FULL": 1, "UNIF-FULL": 2}[sel], t["pchb"]["p_first"], t["pchb"]["p_second"])
        elif self.kind == capi.SYS_HUBBARD_RS:
            engine.set_system_hubbard_rs(t["max_neigh"], t["neighbours"], t["tmat"], t["uhub"])
        elif self.kind == capi.SYS_HUBBARD_K:
            engine.set_system_hubbard_k(t["n_k"], t["ksum"], t["kdiff"], t["eps_k"], t["u_over_n"])
        else:
            raise ValueError("unknown system kind")


# ---------------------------------------------------------------------------------------
def random_fcidump_system(n_spat, nel, sparse=1.0, sparse_t=1.0, seed=25, diag_shift=2.0,
                          p_singles=0.1, p_parallel=None, ms2=0, particle_selection="UNIF-UNIF"):
    """Synthetic FCIDUMP per generate_random_integrals (src/unit_test_helper_excitgen.F90:371-485),
    PCHB `MANUAL UNIF:UNIF UNIF-UNIF:FAST-FAST` spatial-orbital tables."""
    L = lib()
    nb = 2 * n_spat
    n_umat = L.neci_host_umat_size(C.c_int32(n_spat))
    umat = np.zeros(n_umat)
    tmat = np.zeros(nb * nb)
    L.neci_host_random_fcidump(C.c_int32(n_spat), C.c_double(sparse), C.c_double(sparse_t), C.c_uint64(seed),
                               C.c_double(diag_shift), _p(umat, C.c_double), _p(tmat, C.c_double))
    nalpha = (nel + ms2) // 2
    nbeta = nel - nalpha
    pchb = build_pchb(n_spat, umat, p_singles=p_singles, p_parallel=p_parallel, nalpha=nalpha, nbeta=nbeta,
                      particle_selection=particle_selection)
    # aufbau reference: lowest nbeta beta orbitals (odd) and nalpha alpha orbitals (even)
    ref = sorted([2 * i - 1 for i in range(1, nbeta + 1)] + [2 * i for i in range(1, nalpha + 1)])
    return System(kind=capi.SYS_FCIDUMP_PCHB, nel=nel, nbasis=nb, nocc_alpha=nalpha, nocc_beta=nbeta,
                  ecore=0.0, t_exch=1, t_no_brillouin=1,
                  tables=dict(umat=umat, tmat=tmat, pchb=pchb), ref_orbs=np.array(ref, dtype=np.int32))


def fcidump_system(norb, nelec, h1, eri, ecore=0.0, ms2=0, orbsym=None, eps=None, ref_spatial=None,
                   p_singles=0.1, p_parallel=None, particle_selection="UNIF-UNIF"):
    """A system from FCIDUMP data (what IntInit / readint hand over, src/readint.F90): h1 = [(i, j, value)], eri =
    [(i, j, k, l, value)] in the FCIDUMP's chemist order (ij|kl) with 1-based spatial orbitals.
    UMAT is the packed 8-fold array indexed by UMatInd (src/UMatCache.F90:257-296): <pr|qs> = (pq|rs) sits at
    tri(tri(p,q), tri(r,s)); TMAT2D is spin-orbital, h_pq between equal spins.  ORBSYM gives the (spin, irrep) classes
    of the uniform singles generator; the reference determinant is the aufbau filling by orbital energy."""
    L = lib()
    nb = 2 * norb

    def tri(a, b):
        return a * (a - 1) // 2 + b if a > b else b * (b - 1) // 2 + a
    umat = np.zeros(L.neci_host_umat_size(C.c_int32(norb)))
    for i, j, k, l, v in eri:
        umat[tri(tri(int(i), int(j)), tri(int(k), int(l))) - 1] = v
    tmat = np.zeros(nb * nb)
    for i, j, v in h1:
        for (a, b) in ((int(i), int(j)), (int(j), int(i))):
            for spin in (0, 1):                      # beta = odd spin orbital 2a-1, alpha = even 2a
                so_a, so_b = 2 * a - 1 + spin, 2 * b - 1 + spin
                tmat[(so_a - 1) + nb * (so_b - 1)] = v
    nalpha = (nelec + ms2) // 2
    nbeta = nelec - nalpha
    cls = None
    if orbsym is not None:
        labels = sorted(set(int(x) for x in orbsym))
        cls = np.zeros(nb, dtype=np.int32)
        for so in range(1, nb + 1):
            irr = labels.index(int(orbsym[(so + 1) // 2 - 1]))
            cls[so - 1] = 2 * irr + (0 if so % 2 else 1)
    pchb = build_pchb(norb, umat, p_singles=p_singles, p_parallel=p_parallel, nalpha=nalpha, nbeta=nbeta,
                      class_of_spinorb=cls, particle_selection=particle_selection)
    if ref_spatial is None:
        order = np.argsort(np.asarray(eps if eps is not None else [tmat[(2 * a - 2) + nb * (2 * a - 2)] for a in range(1, norb + 1)]),
                           kind="stable")
        ref = sorted([2 * int(a) + 1 for a in order[:nbeta]] + [2 * int(a) + 2 for a in order[:nalpha]])
    else:
        ref = sorted([2 * a - 1 for a in ref_spatial[:nbeta]] + [2 * a for a in ref_spatial[:nalpha]])
    return System(kind=capi.SYS_FCIDUMP_PCHB, nel=nelec, nbasis=nb, nocc_alpha=nalpha, nocc_beta=nbeta,
                  ecore=float(ecore), t_exch=1, t_no_brillouin=0,
                  tables=dict(umat=umat, tmat=tmat, pchb=pchb), ref_orbs=np.array(ref, dtype=np.int32))


def build_pchb(n_spat, umat, p_singles=0.1, p_parallel=None, nalpha=None, nbeta=None, class_of_spinorb=None,
               particle_selection="UNIF-UNIF"):
    """GAS_doubles_PCHB_compute_samplers (src/gasci_pchb_doubles_spatorb_fastweighted.fpp:329-445).
    particle_selection: "UNIF-UNIF" (pick_biased_elecs), "FULL-FULL" (PC_FullyWeightedParticles_t,
    src/gasci_pchb_doubles_select_particles.fpp:330-384) or "UNIF-FULL" (PC_WeightedParticles_t, :440-506); the
    weighted ones add the tables p_first / p_second."""
    L = lib()
    ij = C.c_int32(); ab = C.c_int32()
    L.neci_host_pchb_dims(C.c_int32(n_spat), C.byref(ij), C.byref(ab))
    ij_max, ab_max = ij.value, ab.value
    n = ij_max * 3 * ab_max
    probs = np.zeros(n); bias = np.zeros(n); alias = np.zeros(n, dtype=np.int32)
    p_exch = np.zeros(ij_max); tgt = np.zeros(2 * ab_max, dtype=np.int32)
    um = np.ascontiguousarray(umat, dtype=np.float64)
    L.neci_host_pchb_build(C.c_int32(n_spat), _p(um, C.c_double), _p(probs, C.c_double), _p(bias, C.c_double),
                           _p(alias, C.c_int32), _p(p_exch, C.c_double), _p(tgt, C.c_int32))
    if p_parallel is None:
        # fraction of same-spin electron pairs (what the tau-search would start from)
        par = nalpha * (nalpha - 1) // 2 + nbeta * (nbeta - 1) // 2
        opp = nalpha * nbeta
        p_parallel = par / float(par + opp) if par + opp else 0.0
    if class_of_spinorb is None:
        # ORBSYM all 1: one class per spin (0: beta = odd orbitals, 1: alpha = even orbitals)
        class_of_spinorb = np.array([0 if (o % 2) else 1 for o in range(1, 2 * n_spat + 1)], dtype=np.int32)
    out = dict(n_spat=n_spat, ij_max=ij_max, ab_max=ab_max, probs=probs, bias=bias, alias=alias, p_exch=p_exch,
               tgt_orbs=tgt, p_singles=float(p_singles), p_doubles=1.0 - float(p_singles),
               p_parallel=float(p_parallel), n_classes=int(class_of_spinorb.max()) + 1,
               class_of_spinorb=class_of_spinorb, particle_selection=particle_selection)
    if particle_selection in ("FULL-FULL", "UNIF-FULL"):
        nb = 2 * n_spat
        p_first = np.zeros(nb); p_second = np.zeros(nb * nb)
        L.neci_host_pchb_particle_probs(C.c_int32(n_spat), _p(um, C.c_double), _p(p_first, C.c_double), _p(p_second, C.c_double))
        out["p_first"] = p_first; out["p_second"] = p_second
    elif particle_selection != "UNIF-UNIF":
        raise ValueError("particle_selection must be UNIF-UNIF, FULL-FULL or UNIF-FULL")
    return out


def hubbard_rs_system(lx, ly, nel=None, U=4.0, t=1.0, pbc=True):
    """Real-space Hubbard on an lx x ly square lattice (gen_excit_rs_hubbard); half filling by default."""
    L = lib()
    ns = lx * ly
    nb = 2 * ns
    nel = ns if nel is None else nel
    neigh = np.zeros(nb * 4, dtype=np.int32)
    tmat = np.zeros(nb * nb)
    L.neci_host_hubbard_rs_setup(C.c_int32(lx), C.c_int32(ly), C.c_int32(1 if pbc else 0), C.c_double(t),
                                 _p(neigh, C.c_int32), _p(tmat, C.c_double))
    nalpha = (nel + 1) // 2
    nbeta = nel - nalpha
    # Neel-like reference: alpha on even-parity sites, beta on odd-parity sites, in site order
    sites_a = [s for s in range(ns) if ((s % lx) + (s // lx)) % 2 == 0]
    sites_b = [s for s in range(ns) if ((s % lx) + (s // lx)) % 2 == 1]
    rest = [s for s in range(ns)]
    a_sites = (sites_a + [s for s in rest if s not in sites_a])[:nalpha]
    b_sites = (sites_b + [s for s in rest if s not in sites_b])[:nbeta]
    ref = sorted([2 * (s + 1) for s in a_sites] + [2 * (s + 1) - 1 for s in b_sites])
    return System(kind=capi.SYS_HUBBARD_RS, nel=nel, nbasis=nb, nocc_alpha=nalpha, nocc_beta=nbeta, ecore=0.0,
                  t_exch=0, t_no_brillouin=1,     # real_space_hubbard.F90:151-160
                  tables=dict(max_neigh=4, neighbours=neigh, tmat=tmat, uhub=float(U)),
                  ref_orbs=np.array(ref, dtype=np.int32))


def hubbard_k_system(lx, ly, nel=None, U=4.0, t=1.0):
    """k-space Hubbard on an lx x ly periodic mesh (gen_excit_k_space_hub); half filling by default."""
    L = lib()
    nk = lx * ly
    nb = 2 * nk
    nel = nk if nel is None else nel
    ksum = np.zeros(nk * nk, dtype=np.int32); kdiff = np.zeros(nk * nk, dtype=np.int32); eps = np.zeros(nk)
    L.neci_host_hubbard_k_setup(C.c_int32(lx), C.c_int32(ly), C.c_double(t), _p(ksum, C.c_int32), _p(kdiff, C.c_int32),
                                _p(eps, C.c_double))
    nalpha = (nel + 1) // 2
    nbeta = nel - nalpha
    ref = sorted([2 * i for i in range(1, nalpha + 1)] + [2 * i - 1 for i in range(1, nbeta + 1)])
    return System(kind=capi.SYS_HUBBARD_K, nel=nel, nbasis=nb, nocc_alpha=nalpha, nocc_beta=nbeta, ecore=0.0,
                  t_exch=1, t_no_brillouin=0,     # k_space_hubbard.F90:391-393
                  tables=dict(n_k=nk, ksum=ksum, kdiff=kdiff, eps_k=eps, u_over_n=float(U) / nk),
                  ref_orbs=np.array(ref, dtype=np.int32))


def random_hash_tables(nbasis, seed=7):
    """RandomOrbIndex / RandomHash2 (src/fcimc_initialisation.fpp:862-942)."""
    roi = np.zeros(nbasis, dtype=np.int32); rh2 = np.zeros(nbasis, dtype=np.int32)
    lib().neci_host_random_hash_tables(C.c_int32(nbasis), C.c_uint64(seed), _p(roi, C.c_int32), _p(rh2, C.c_int32))
    return roi, rh2


def update_shift(diag_sft, sft_damp, tau, steps_sft, av_walkers, old_av_walkers):
    """update_shift (src/fcimc_iter_utilities.F90:1063-1072)."""
    return lib().neci_host_update_shift(C.c_double(diag_sft), C.c_double(sft_damp), C.c_double(tau),
                                        C.c_int32(steps_sft), C.c_double(av_walkers), C.c_double(old_av_walkers))


def make_params(system, hii, max_walkers, max_spawned, nranks=1, rank=0, device=0, seed=7, initiator=True,
                initiator_walk_no=3.0, all_real_coeff=False, real_spawn_cutoff=0.95, occupied_thresh=1.0,
                av_mc_excits=1.0, semi_stochastic=False, blocks_per_rank=1, hash_seed=7, mapping=None,
                death_before_comms=None, tau_search=False, consider_par_bias=None, hphf=False,
                random_orb_index=None):
    """The module-level globals of the reference that the engine needs (neci_gpu_config).
    Defaults follow src/Calc.F90:120-480; tDeathBeforeComms is .false. unless the walkers are integers
    (src/Calc.F90:475, src/fcimc_initialisation.fpp:1997-2001) or DEATH-BEFORE-COMMS is given."""
    if death_before_comms is None:
        death_before_comms = not all_real_coeff
    roi, rh2 = random_hash_tables(system.nbasis, hash_seed)
    if random_orb_index is not None:          # the host's own RandomOrbIndex (src/fcimc_initialisation.fpp:862-890)
        roi = np.asarray(random_orb_index, dtype=np.int32)
    balance_blocks = nranks * blocks_per_rank
    if mapping is None:
        # init_load_balance (src/load_balancer.fpp:72-99): LoadBalanceMapping(i) = int((i-1)/oversample_factor),
        # i.e. contiguous runs of blocks per rank
        mapping = np.array([b // blocks_per_rank for b in range(balance_blocks)], dtype=np.int32)
    return dict(
        nel=system.nel, nbasis=system.nbasis, nifd=system.nifd, niftot=system.nifd + 2,
        nocc_alpha=system.nocc_alpha, nocc_beta=system.nocc_beta, nranks=nranks, rank=rank, device=device,
        balance_blocks=balance_blocks, max_walkers=int(max_walkers), max_spawned=int(max_spawned),
        system_type=system.kind, t_trunc_initiator=int(initiator), t_all_real_coeff=int(all_real_coeff),
        t_real_spawn_cutoff=int(all_real_coeff), t_death_before_comms=int(bool(death_before_comms)), t_init_coherent_rule=1,
        t_no_brillouin=system.t_no_brillouin, t_exch=system.t_exch, t_semi_stochastic=int(semi_stochastic),
        t_core_inits=1, t_tau_search=int(bool(tau_search)),
        # consider_par_bias: true for the PCHB generator with uniform particle selection and for the k-space Hubbard
        # generator's parent class, false for the real-space lattice (tau/tau_search_conventional.F90:80-117)
        t_consider_par_bias=int((system.kind == 1) if consider_par_bias is None else bool(consider_par_bias)),
        t_hphf=int(bool(hphf)), reserved0=0,
        initiator_walk_no=float(initiator_walk_no), real_spawn_cutoff=float(real_spawn_cutoff),
        occupied_thresh=float(occupied_thresh), av_mc_excits=float(av_mc_excits), hii=float(hii),
        ecore=float(system.ecore), seed=int(seed), random_orb_index=roi, random_hash2=rh2,
        load_balance_mapping=np.asarray(mapping, dtype=np.int32), ilut_ref=system.ilut(system.ref_orbs))


def record(system, orbs, sign, flags=0):
    """One CurrentDets record ilut(0:NIfTot) as int64 words."""
    rec = np.zeros(system.W, dtype=np.int64)
    rec[:system.nw] = system.ilut(orbs)
    rec[system.nw] = np.array([sign], dtype=np.float64).view(np.int64)[0]
    rec[system.nw + 1] = flags
    return rec


def signs_of(dets, nw):
    return np.ascontiguousarray(dets[:, nw]).view(np.float64)


# ---- POPSFILE body (text form) ---------------------------------------------------------------
def popsfile_lines(dets, gd=None, go=None, nw=1):
    """The determinant lines of a text POPSFILE as write_pops_det formats them (src/Popsfile.F90:2095-2107):
    (i24) per orbital word, (f30.8) sign, (i24) flags, (f30.8) per gdata row."""
    dets = np.asarray(dets, dtype=np.int64).reshape(-1, nw + 2)
    sg = signs_of(dets, nw)
    out = []
    for j in range(dets.shape[0]):
        line = "".join("%24d" % int(dets[j, k]) for k in range(nw)) + "%30.8f" % sg[j] + "%24d" % int(dets[j, nw + 1])
        if gd is not None:
            line += "%30.8f" % gd[j] + "%30.8f" % go[j]
        out.append(line)
    return out


def parse_popsfile_lines(lines, nw=1):
    """Inverse of popsfile_lines: CurrentDets records (and gdata rows if present) for neci_gpu_upload_walkers."""
    dets = np.zeros((len(lines), nw + 2), dtype=np.int64)
    gd, go = [], []
    for j, ln in enumerate(lines):
        t = ln.split()
        for k in range(nw):
            dets[j, k] = int(t[k])
        dets[j, nw] = np.array([float(t[nw])], dtype=np.float64).view(np.int64)[0]
        dets[j, nw + 1] = int(t[nw + 1])
        if len(t) > nw + 2:
            gd.append(float(t[nw + 2])); go.append(float(t[nw + 3]))
    return dets, (np.array(gd) if gd else None), (np.array(go) if go else None)


# ---------------------------------------------------------------------------------------
# semi-stochastic set-up (what init_semi_stochastic hands to neci_gpu_set_core_space)
# ---------------------------------------------------------------------------------------
def _iluts(system, iluts):
    return np.ascontiguousarray(np.asarray(iluts, dtype=np.int64).reshape(-1, system.nw))


def _integral_tables(system):
    """UMAT / TMAT2D of a system for the host library: as they are for FCIDUMP systems; for the real-space Hubbard
    model the on-site repulsion <ii|ii> = U in UMatInd packing beside its hopping matrix."""
    if system.kind == capi.SYS_FCIDUMP_PCHB:
        return system.tables
    if system.kind == capi.SYS_HUBBARD_RS:
        if "umat" not in system.tables:
            ns = system.nbasis // 2
            umat = np.zeros(lib().neci_host_umat_size(C.c_int32(ns)))
            for i in range(1, ns + 1):
                p = i * (i - 1) // 2 + i                       # tri(i, i)
                umat[p * (p - 1) // 2 + p - 1] = system.tables["uhub"]
            system.tables["umat"] = umat
        return system.tables
    raise ValueError("no tabulated integrals for this system kind")


def get_helement(system, iluts_i, iluts_j, hphf=False):
    """get_helement (src/Determinants.F90:508-554) for pairs of determinants of an FCIDUMP system, on the host;
    hphf: between HPHF functions given by their allowed representatives (src/HPHFIntegrals.fpp)."""
    a, b = _iluts(system, iluts_i), _iluts(system, iluts_j)
    out = np.zeros(a.shape[0])
    if system.kind == capi.SYS_HUBBARD_K:
        if hphf:
            raise ValueError("host get_helement: no HPHF functions for the lattice models")
        t = system.tables
        ks, ek = np.ascontiguousarray(t["ksum"], dtype=np.int32), np.ascontiguousarray(t["eps_k"], dtype=np.float64)
        rc = lib().neci_host_get_helement_hubbard_k(C.c_int32(system.nel), C.c_int32(system.nbasis), C.c_int32(t["n_k"]),
                                                    _p(ks, C.c_int32), _p(ek, C.c_double), C.c_double(t["u_over_n"]),
                                                    _p(a, C.c_int64), _p(b, C.c_int64), C.c_int64(a.shape[0]),
                                                    _p(out, C.c_double))
        if rc:
            raise RuntimeError("neci_host_get_helement_hubbard_k failed (%d)" % rc)
        return out
    t = _integral_tables(system)
    fn = lib().neci_host_get_helement_hphf if hphf else lib().neci_host_get_helement
    rc = fn(C.c_int32(system.nel), C.c_int32(system.nbasis), _p(t["umat"], C.c_double),
                                      _p(t["tmat"], C.c_double), C.c_double(system.ecore), _p(a, C.c_int64),
                                      _p(b, C.c_int64), C.c_int64(a.shape[0]), _p(out, C.c_double))
    if rc:
        raise RuntimeError("neci_host_get_helement failed (%d)" % rc)
    return out


def sing_doub_space(system, ref_ilut=None, only_keep_conn=False, orbsym=None):
    """`doubles-core`: the reference determinant and its single and double excitations
    (generate_sing_doub_determinants, src/semi_stoch_gen.F90:537-604) as n x nw occupation words.
    orbsym: ORBSYM labels of the spatial orbitals (abelian groups); None = all irreps equal."""
    ref = _iluts(system, system.ilut(system.ref_orbs) if ref_ilut is None else ref_ilut)
    t = system.tables
    if system.kind == capi.SYS_HUBBARD_K:
        # enumerate_sing_doub_kpnt: spin- and momentum-conserving doubles only
        ks = np.ascontiguousarray(t["ksum"], dtype=np.int32)
        L = lib()
        L.neci_host_sd_space_hubbard_k.restype = C.c_int64
        need = -L.neci_host_sd_space_hubbard_k(C.c_int32(system.nbasis), C.c_int32(t["n_k"]), _p(ks, C.c_int32),
                                               _p(ref, C.c_int64), C.c_int64(0), None)
        out = np.zeros((max(need, 1), system.nw), dtype=np.int64)
        n = L.neci_host_sd_space_hubbard_k(C.c_int32(system.nbasis), C.c_int32(t["n_k"]), _p(ks, C.c_int32),
                                           _p(ref, C.c_int64), C.c_int64(out.shape[0]), _p(out, C.c_int64))
        if n <= 0:
            raise RuntimeError("neci_host_sd_space_hubbard_k failed (%d)" % n)
        return out[:n].copy()
    na, nb_ = system.nocc_alpha, system.nocc_beta
    va, vb = system.nbasis // 2 - na, system.nbasis // 2 - nb_
    cap = 1 + na * va + nb_ * vb + (na * (na - 1) // 2) * (va * (va - 1) // 2) \
        + (nb_ * (nb_ - 1) // 2) * (vb * (vb - 1) // 2) + na * va * nb_ * vb
    out = np.zeros((cap, system.nw), dtype=np.int64)
    sym = None if orbsym is None else np.ascontiguousarray(orbsym, dtype=np.int32)
    assert sym is None or sym.shape[0] == system.nbasis // 2
    n = lib().neci_host_sd_space(C.c_int32(system.nel), C.c_int32(system.nbasis), _p(t["umat"], C.c_double),
                                 _p(t["tmat"], C.c_double), _p(ref, C.c_int64), C.c_int32(int(only_keep_conn)),
                                 None if sym is None else _p(sym, C.c_int32), C.c_int64(cap), _p(out, C.c_int64))
    if n <= 0:
        raise RuntimeError("neci_host_sd_space failed (%d)" % n)
    return out[:n].copy()


def det_node(params, iluts, nw):
    """DetermineDetNode (src/load_balance_calcnodes.F90:25-117) on the host -> (block, node) per determinant."""
    il = np.ascontiguousarray(np.asarray(iluts, dtype=np.int64).reshape(-1, nw))
    blocks = np.zeros(il.shape[0], dtype=np.int32)
    nodes = np.zeros(il.shape[0], dtype=np.int32)
    roi = np.ascontiguousarray(params["random_orb_index"], dtype=np.int32)
    mapping = np.ascontiguousarray(params["load_balance_mapping"], dtype=np.int32)
    rc = lib().neci_host_det_node(C.c_int32(params["nbasis"]), _p(roi, C.c_int32), C.c_int32(params["balance_blocks"]),
                                  _p(mapping, C.c_int32), _p(il, C.c_int64), C.c_int64(il.shape[0]),
                                  _p(blocks, C.c_int32), _p(nodes, C.c_int32))
    if rc:
        raise RuntimeError("neci_host_det_node failed (%d)" % rc)
    return blocks, nodes


def _ilut_sort_keys(il):
    """Keys for np.lexsort reproducing ilut_lt (src/DetBitOps.F90:431-473): signed compare, word 0 most significant."""
    return tuple(il[:, k] for k in range(il.shape[1] - 1, -1, -1))


def layout_core_space(core_iluts, nodes, nranks):
    """Order of the whole core space as store_whole_core_space leaves it (src/semi_stoch_gen.F90:214-243): rank by
    rank, each rank's determinants sorted with ilut_lt.  Returns (ordered iluts, sizes, displs)."""
    il = np.asarray(core_iluts, dtype=np.int64)
    nodes = np.asarray(nodes)
    order = np.lexsort(_ilut_sort_keys(il) + (nodes,))
    sizes = np.bincount(nodes, minlength=nranks).astype(np.int32)
    displs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    return np.ascontiguousarray(il[order]), sizes, displs


def core_hamiltonian(system, core_iluts, hii, displ=0, n_local=None, threads=0, hphf=False):
    """Sparse core Hamiltonian rows of one rank (calc_determ_hamil_sparse / _hphf, src/sparse_arrays.F90:426-690; each row:
    the non-zero off-diagonal elements, then H_ii - Hii last as src/fast_determ_hamil.F90:1494-1507 leaves it).
    Returns dict(row_ptr int64, col int32, val float64) for neci_gpu_set_core_space."""
    il = _iluts(system, core_iluts)
    n_core = il.shape[0]
    n_local = n_core - displ if n_local is None else int(n_local)
    nnz = C.c_int64(0)
    L = lib()
    if system.kind == capi.SYS_HUBBARD_K:
        t = system.tables
        ks, ek = np.ascontiguousarray(t["ksum"], dtype=np.int32), np.ascontiguousarray(t["eps_k"], dtype=np.float64)
        L.neci_host_core_ham_build_hubbard_k.restype = C.c_void_p
        h = L.neci_host_core_ham_build_hubbard_k(C.c_int32(system.nel), C.c_int32(system.nbasis), C.c_int32(t["n_k"]),
                                                 _p(ks, C.c_int32), _p(ek, C.c_double), C.c_double(t["u_over_n"]),
                                                 C.c_double(hii), _p(il, C.c_int64), C.c_int64(n_core), C.c_int64(int(displ)),
                                                 C.c_int64(n_local), C.c_int32(int(threads)), C.byref(nnz))
    else:
        t = _integral_tables(system)
        h = L.neci_host_core_ham_build(C.c_int32(system.nel), C.c_int32(system.nbasis), _p(t["umat"], C.c_double),
                                       _p(t["tmat"], C.c_double), C.c_double(system.ecore), C.c_double(hii),
                                       _p(il, C.c_int64), C.c_int64(n_core), C.c_int64(int(displ)), C.c_int64(n_local),
                                       C.c_int32(int(threads)), C.c_int32(int(bool(hphf))), C.byref(nnz))
    if not h:
        raise RuntimeError("neci_host_core_ham_build failed")
    row_ptr = np.zeros(n_local + 1, dtype=np.int64)
    col = np.zeros(nnz.value, dtype=np.int32)
    val = np.zeros(nnz.value, dtype=np.float64)
    L.neci_host_core_ham_fetch(C.c_void_p(h), _p(row_ptr, C.c_int64), _p(col, C.c_int32), _p(val, C.c_double))
    return dict(row_ptr=row_ptr, col=col, val=val)


def ham_apply(system, rows, cols, vec, threads=0):
    """out_i = sum_j <row_i|H|col_j> vec_j on the host (FCIDUMP systems)."""
    if system.kind != capi.SYS_FCIDUMP_PCHB:
        raise ValueError("host ham_apply: FCIDUMP systems only")
    r, c = _iluts(system, rows), _iluts(system, cols)
    v = np.ascontiguousarray(vec, dtype=np.float64)
    assert v.shape[0] == c.shape[0]
    out = np.zeros(r.shape[0])
    t = system.tables
    rc = lib().neci_host_ham_apply(C.c_int32(system.nel), C.c_int32(system.nbasis), _p(t["umat"], C.c_double),
                                   _p(t["tmat"], C.c_double), C.c_double(system.ecore), _p(r, C.c_int64),
                                   C.c_int64(r.shape[0]), _p(c, C.c_int64), C.c_int64(c.shape[0]), _p(v, C.c_double),
                                   C.c_int32(int(threads)), _p(out, C.c_double))
    if rc:
        raise RuntimeError("neci_host_ham_apply failed (%d)" % rc)
    return out


def hphf_representative(iluts):
    """The allowed HPHF representative of every determinant (IsAllowedHPHF, src/DetBitOps.F90:693-718): the larger of
    the determinant and its spin-flipped partner in the signed word order of DetBitLT, word 0 first."""
    x = np.ascontiguousarray(iluts, dtype=np.int64).view(np.uint64)
    A, B = np.uint64(0xAAAAAAAAAAAAAAAA), np.uint64(0x5555555555555555)
    f = ((x & A) >> np.uint64(1)) | ((x & B) << np.uint64(1))
    xs, fs = x.view(np.int64), f.view(np.int64)
    take_x = np.ones(x.shape[0], dtype=bool)
    decided = np.zeros(x.shape[0], dtype=bool)
    for w in range(x.shape[1]):
        gt, lt = (xs[:, w] > fs[:, w]) & ~decided, (xs[:, w] < fs[:, w]) & ~decided
        take_x[lt] = False
        decided |= gt | lt
    return np.where(take_x[:, None], xs, fs)


def trial_space(system, trial_iluts, orbsym=None, hphf=False):
    """init_trial_wf (src/trial_wf_gen.F90) for a given trial space: the trial vector is the lowest eigenvector of H
    in that space; the connected space is every determinant outside it within two excitations of one of its members
    (generate_connected_space) with con_space_vecs_i = sum_j H_ij psiT_j != 0.
    orbsym: ORBSYM labels (symmetry-allowed excitations only, as GenExcitations3 enumerates them); None = all equal.
    hphf: trial_iluts are allowed HPHF representatives and H acts between HPHF functions.
    Returns (trial_iluts, trial_amps, con_iluts, con_amps, trial_energy) for neci_gpu_set_trial_space."""
    ti = _iluts(system, trial_iluts)
    nt = ti.shape[0]
    I = np.repeat(np.arange(nt), nt); J = np.tile(np.arange(nt), nt)
    Ht = get_helement(system, ti[I], ti[J], hphf=hphf).reshape(nt, nt)
    w, v = np.linalg.eigh(Ht)
    psi = v[:, 0].copy()
    con = np.concatenate([sing_doub_space(system, ref_ilut=ti[k], orbsym=orbsym)[1:] for k in range(nt)])
    if hphf:                                  # generate_connection_normal: excitations of the representative, mapped
        con = hphf_representative(con)        # to their own allowed representatives (enumerate_excitations.F90:310-316)
    con = np.unique(np.concatenate([ti, con]), axis=0)
    con = np.ascontiguousarray(con[~rows_in(con, ti)])            # without the trial determinants themselves
    if hphf:
        amps = np.zeros(con.shape[0])
        step = max(1, (1 << 22) // max(nt, 1))
        for lo in range(0, con.shape[0], step):
            blk = con[lo:lo + step]
            I = np.repeat(np.arange(blk.shape[0]), nt); J = np.tile(np.arange(nt), blk.shape[0])
            amps[lo:lo + step] = get_helement(system, blk[I], ti[J], hphf=True).reshape(blk.shape[0], nt) @ psi
    else:
        amps = ham_apply(system, con, ti, psi)
    keep = np.abs(amps) > 0
    return ti, psi, np.ascontiguousarray(con[keep]), amps[keep], float(w[0])


def rows_in(a, b):
    """Boolean mask: which rows of a occur in b (both n x nw int64)."""
    av = np.ascontiguousarray(a).view([("", a.dtype)] * a.shape[1]).ravel()
    bv = np.ascontiguousarray(b).view([("", b.dtype)] * b.shape[1]).ravel()
    return np.isin(av, bv)


def cas_space(system, eps, occ_cas, virt_cas, orbsym=None):
    """`cas-core OccCASOrbs VirtCASOrbs` / `cas-trial` (generate_cas, src/semi_stoch_gen.F90): the highest occ_cas
    occupied and the lowest virt_cas virtual SPIN orbitals of the energy-ordered reference are active, everything below
    stays doubly occupied; all determinants of the reference's Ms and (with orbsym) irrep.  Closed-shell references,
    even occ_cas / virt_cas.  Returns n x nw occupation words."""
    import itertools
    if system.nocc_alpha != system.nocc_beta or occ_cas % 2 or virt_cas % 2:
        raise ValueError("cas_space: closed-shell reference and even active-orbital counts only")
    order = [int(x) + 1 for x in np.argsort(np.asarray(eps, dtype=float), kind="stable")]
    nocc, na = system.nocc_alpha, occ_cas // 2
    core = order[:nocc - na]
    act = order[nocc - na: nocc + virt_cas // 2]
    irr = (lambda orbs: 0) if orbsym is None else \
        (lambda orbs: int(np.bitwise_xor.reduce([int(orbsym[(o + 1) // 2 - 1]) - 1 for o in orbs])))
    target = irr([int(x) for x in system.ref_orbs])
    out = []
    for a in itertools.combinations(act, na):
        for b in itertools.combinations(act, na):
            d = sorted([2 * c for c in core] + [2 * c - 1 for c in core] + [2 * x for x in a] + [2 * x - 1 for x in b])
            if irr(d) == target:
                out.append(system.ilut(d))
    return np.array(out, dtype=np.int64).reshape(len(out), system.nw)


def optimised_space(system, cutoff_num=None, cutoff_amp=None, orbsym=None):
    """`optimised-core` / `optimised-trial` (generate_optimised_space, src/semi_stoch_gen.F90:824-1029): starting from
    the reference, each loop takes the space connected to the current one (its determinants and their allowed singles
    and doubles), finds the ground state of H in it and keeps the cutoff_num[k] determinants of largest amplitude, or
    those with |amplitude| above cutoff_amp[k] (eigenvector normalised to 1).  Returns n x nw occupation words."""
    if (cutoff_num is None) == (cutoff_amp is None):
        raise ValueError("optimised_space: give cutoff_num or cutoff_amp")
    space = _iluts(system, system.ilut(system.ref_orbs))
    for k in range(len(cutoff_num if cutoff_num is not None else cutoff_amp)):
        con = np.unique(np.concatenate([space] + [sing_doub_space(system, ref_ilut=r, orbsym=orbsym) for r in space]), axis=0)
        n = con.shape[0]
        I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
        w, v = np.linalg.eigh(get_helement(system, con[I], con[J]).reshape(n, n))
        a = np.abs(v[:, 0])
        keep = np.argsort(-a, kind="stable")[:int(cutoff_num[k])] if cutoff_num is not None else np.nonzero(a > cutoff_amp[k])[0]
        space = np.ascontiguousarray(con[np.sort(keep)])
    return space


def ras_space(system, eps, ras1, ras2, ras3, min1, max3, orbsym=None):
    """`ras-core ras1 ras2 ras3 min1 max3` / `ras-trial` (generate_ras, src/semi_stoch_gen.F90): the energy-ordered
    spatial orbitals are split into RAS1 (the first ras1), RAS2 (the next ras2) and RAS3 (the next ras3); all
    determinants of the reference's Ms (and irrep, with orbsym) with at least min1 electrons in RAS1 and at most max3 in
    RAS3.  Plain enumeration of the Ms sector: for the small systems such spaces are used on."""
    import itertools
    ns = system.nbasis // 2
    if ras1 + ras2 + ras3 != ns:
        raise ValueError("ras_space: the three spaces must cover all %d spatial orbitals" % ns)
    order = [int(x) + 1 for x in np.argsort(np.asarray(eps, dtype=float), kind="stable")]
    r1, r3 = set(order[:ras1]), set(order[ras1 + ras2:])
    irr = (lambda orbs: 0) if orbsym is None else \
        (lambda orbs: int(np.bitwise_xor.reduce([int(orbsym[(o + 1) // 2 - 1]) - 1 for o in orbs])))
    target = irr([int(x) for x in system.ref_orbs])
    out = []
    for a in itertools.combinations(range(1, ns + 1), system.nocc_alpha):
        for b in itertools.combinations(range(1, ns + 1), system.nocc_beta):
            n1 = sum(1 for x in a if x in r1) + sum(1 for x in b if x in r1)
            n3 = sum(1 for x in a if x in r3) + sum(1 for x in b if x in r3)
            if n1 < min1 or n3 > max3:
                continue
            d = sorted([2 * x for x in a] + [2 * x - 1 for x in b])
            if irr(d) == target:
                out.append(system.ilut(d))
    return np.array(out, dtype=np.int64).reshape(len(out), system.nw)


def most_populated_space(dets, n, nw=1):
    """`pops-core n` / `pops-trial n` (generate_space_most_populated, src/semi_stoch_gen.F90:1031-1221): the n
    determinants of a walker list (CurrentDets records, e.g. from neci_gpu_download_occupied) with the largest |sign|;
    determinants below 1e-8 are never taken, fewer than n are returned if the list is shorter.  Returns (n x nw
    occupation words, their signs)."""
    d = np.asarray(dets, dtype=np.int64).reshape(-1, nw + 2)
    sg = signs_of(d, nw)
    ok = np.nonzero(np.abs(sg) >= 1.e-8)[0]
    pick = ok[np.argsort(-np.abs(sg[ok]), kind="stable")[:int(n)]]
    return np.ascontiguousarray(d[pick, :nw]), sg[pick].copy()

"""Generates tests/golden/hehe_ss_doubles.json from the reference's own regression case
test_suite/neci/parallel/HeHe_SS_Doubles: the FCIDUMP integrals (input data of that test) and the numbers the
reference's CPU run printed into its checked-in benchmark output (reference-determinant energy, energy of the
highest determinant, final projected energy with its error bar).  Run in the build container, where /root/reference
exists; the tests only read the JSON."""
import glob
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
CASE = os.path.join(REF, "test_suite", "neci", "parallel", "HeHe_SS_Doubles")


def main():
    txt = open(os.path.join(CASE, "FCIDUMP")).read()
    head, body = txt.split("&END")
    norb = int(re.search(r"NORB\s*=\s*(\d+)", head).group(1))
    nelec = int(re.search(r"NELEC\s*=\s*(\d+)", head).group(1))
    ms2 = int(re.search(r"MS2\s*=\s*(-?\d+)", head).group(1))
    orbsym = [int(x) for x in re.search(r"ORBSYM\s*=\s*([\d,\s]+)", head).group(1).replace("\n", "").split(",") if x.strip()]
    ecore, eps, h1, eri = 0.0, {}, [], []
    for ln in body.strip().splitlines():
        t = ln.split()
        if len(t) != 5:
            continue
        v = float(t[0]); i, j, k, l = (int(x) for x in t[1:])
        if i == 0:
            ecore = v
        elif j == 0:
            eps[i] = v
        elif k == 0:
            h1.append([i, j, v])
        else:
            eri.append([i, j, k, l, v])
    bench = open(glob.glob(os.path.join(CASE, "benchmark*"))[0]).read()
    ref_energy = float(re.search(r"Current reference energy\s+(-?[\d.]+)", bench).group(1))
    hi = re.search(r"Highest energy determinant is \(approximately\):\s+(-?[\d.Ee+-]+)", bench)
    hi_det = re.search(r"Highest energy determinant is:\s+([\d\s]+)\n", bench)
    tot = re.search(r"Total projected energy\s+(-?[\d.]+)\s*\+/-\s*([\d.Ee+-]+)", bench)
    ref_det = re.search(r"Generated reference determinants:\s*\n\(\s*([\d,\s]+)\)", bench)
    # `Total size of deterministic space` printed by this case (HPHF functions) and by the determinant-basis twin on
    # the same FCIDUMP, test_suite/neci/determ_and_trial_spaces/determ_doubles (semi-stochastic doubles-core)
    core_hphf = int(re.search(r"Total size of deterministic space:\s+(\d+)", bench).group(1))
    twin = os.path.join(REF, "test_suite", "neci", "determ_and_trial_spaces", "determ_doubles")
    assert open(os.path.join(twin, "FCIDUMP")).read() == txt
    bench2 = open(glob.glob(os.path.join(twin, "benchmark*"))[0]).read()
    core_dets = int(re.search(r"Total size of deterministic space:\s+(\d+)", bench2).group(1))
    # that run also prints the lowest eigenvalue of its core Hamiltonian (Davidson) and starts from the core ground
    # state scaled to `startsinglepart` walkers: the first line of the iteration table holds its weight on the
    # reference and on the doubles, and the projected energy of that state
    sd_counts = re.search(r"(\d+) double excitations, and\s+(\d+) single excitations found from reference", bench2)
    e_core = float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", bench2).group(1))
    table = bench2[bench2.index("Step    Shift"):]
    row1 = re.search(r"\n\s+1\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+", table)
    f = [float(x) for x in row1.groups()]
    rows = {}
    for k in (2, 3):
        m = re.search(r"\n\s+%d\s+(\S+)" % k + r"\s+(\S+)" * 11 + r"\s+", table)
        rows[k] = [float(x) for x in m.groups()]
    # the HPHF run itself: core correlation energy and the first two lines of its iteration table
    tab = bench[bench.index("Step    Shift"):]
    hrows = {}
    for k in (1, 2):
        m = re.search(r"\n\s+%d\s+(\S+)" % k + r"\s+(\S+)" * 11 + r"\s+", tab)
        hrows[k] = [float(x) for x in m.groups()]
    hphf_run = dict(core_correlation_energy=float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", bench).group(1)),
                    start_walkers=10.0, tau=0.001, diagshift=1.0,
                    step1_no_at_hf=hrows[1][10], step1_no_at_doubs=hrows[1][11], step2_no_at_hf=hrows[2][10])
    # `semi-stochastic fci-core` on the same FCIDUMP (test_suite/neci/rdm_singlerun/parallel/HeHe_determ): the whole
    # symmetry sector as core space, so the printed core correlation energy is the exact FCI correlation energy
    fci = os.path.join(REF, "test_suite", "neci", "rdm_singlerun", "parallel", "HeHe_determ")
    assert open(os.path.join(fci, "FCIDUMP")).read() == txt
    bench3 = open(glob.glob(os.path.join(fci, "benchmark*"))[0]).read()
    fci_core = dict(source="test_suite/neci/rdm_singlerun/parallel/HeHe_determ (same FCIDUMP; benchmark.out...)",
                    size=int(re.search(r"Total size of deterministic space:\s+(\d+)", bench3).group(1)),
                    correlation_energy=float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", bench3).group(1)))
    # `semi-stochastic read-core`: the determinants of the checked-in CORESPACE file (occupation words) and the lowest
    # eigenvalue the reference printed for them
    rd = os.path.join(REF, "test_suite", "neci", "determ_and_trial_spaces", "determ_read")
    assert open(os.path.join(rd, "FCIDUMP")).read() == txt
    bench4 = open(glob.glob(os.path.join(rd, "benchmark*"))[0]).read()
    read_core = dict(source="test_suite/neci/determ_and_trial_spaces/determ_read (same FCIDUMP; CORESPACE, benchmark.out...)",
                     iluts=[int(x) for x in open(os.path.join(rd, "CORESPACE")).read().split()],
                     size=int(re.search(r"Total size of deterministic space:\s+(\d+)", bench4).group(1)),
                     correlation_energy=float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", bench4).group(1)))
    # trial-wavefunction twins: `doubles-trial` and `read-trial` (TRIALSPACE = the CORESPACE file above)
    trial = {}
    for name in ("trial_doubles", "trial_read", "trial_cas", "trial_opt_num", "trial_opt_amp"):
        td = os.path.join(REF, "test_suite", "neci", "determ_and_trial_spaces", name)
        assert open(os.path.join(td, "FCIDUMP")).read() == txt
        b = open(glob.glob(os.path.join(td, "benchmark*"))[0]).read()
        ti = open(os.path.join(td, "neci.inp")).read()
        cut = re.search(r"optimised-trial-cutoff-(num|amp)\s+([\d.\s]+)\n", ti)
        trial[name] = dict(cutoff=None if not cut else [cut.group(1)] + [float(x) for x in cut.group(2).split()],
                           source="test_suite/neci/determ_and_trial_spaces/%s (same FCIDUMP; benchmark.out...)" % name,
                           trial_size=int(re.search(r"Total size of the trial space:\s+(\d+)", b).group(1)),
                           connected_size=int(re.search(r"Total size of connected space:\s+(\d+)", b).group(1)),
                           trial_energy=float(re.search(r"Energy eigenvalue\(s\) of the trial space:\s+(-?[\d.]+)", b).group(1)))
    assert open(os.path.join(REF, "test_suite", "neci", "determ_and_trial_spaces", "trial_read", "TRIALSPACE")).read() == \
        open(os.path.join(rd, "CORESPACE")).read()
    # `cas-core 2 6` / `cas-trial 2 6`
    cd = os.path.join(REF, "test_suite", "neci", "determ_and_trial_spaces", "determ_cas")
    assert open(os.path.join(cd, "FCIDUMP")).read() == txt
    b5 = open(glob.glob(os.path.join(cd, "benchmark*"))[0]).read()
    inp5 = open(os.path.join(cd, "neci.inp")).read()
    cas_core = dict(source="test_suite/neci/determ_and_trial_spaces/determ_cas (same FCIDUMP; benchmark.out...)",
                    cas=[int(x) for x in re.search(r"cas-core\s+(\d+)\s+(\d+)", inp5).groups()],
                    size=int(re.search(r"Total size of deterministic space:\s+(\d+)", b5).group(1)),
                    correlation_energy=float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", b5).group(1)))
    # `optimised-core` with cut-offs by number and by amplitude
    opt = {}
    for name in ("determ_opt_num", "determ_opt_amp"):
        od = os.path.join(REF, "test_suite", "neci", "determ_and_trial_spaces", name)
        assert open(os.path.join(od, "FCIDUMP")).read() == txt
        b = open(glob.glob(os.path.join(od, "benchmark*"))[0]).read()
        cut = re.search(r"optimised-core-cutoff-(num|amp)\s+([\d.\s]+)\n", open(os.path.join(od, "neci.inp")).read())
        opt[name] = dict(source="test_suite/neci/determ_and_trial_spaces/%s (same FCIDUMP; benchmark.out...)" % name,
                         cutoff=[cut.group(1)] + [float(x) for x in cut.group(2).split()],
                         size=int(re.search(r"Total size of deterministic space:\s+(\d+)", b).group(1)),
                         correlation_energy=float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", b).group(1)))
    # `ras-core 2 0 8 1 3`
    rsd = os.path.join(REF, "test_suite", "neci", "determ_and_trial_spaces", "determ_ras")
    assert open(os.path.join(rsd, "FCIDUMP")).read() == txt
    b6 = open(glob.glob(os.path.join(rsd, "benchmark*"))[0]).read()
    ras_core = dict(source="test_suite/neci/determ_and_trial_spaces/determ_ras (same FCIDUMP; benchmark.out...)",
                    ras=[int(x) for x in re.search(r"ras-core\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)", open(os.path.join(rsd, "neci.inp")).read()).groups()],
                    size=int(re.search(r"Total size of deterministic space:\s+(\d+)", b6).group(1)),
                    correlation_energy=float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", b6).group(1)))
    out = dict(
        ras_core=ras_core,
        optimised_core=opt,
        cas_core=cas_core,
        trial_runs=trial,
        read_core=read_core,
        fci_core=fci_core,
        hphf_run=hphf_run,
        determ_doubles=dict(source="test_suite/neci/determ_and_trial_spaces/determ_doubles (same FCIDUMP; benchmark.out...)",
                            n_doubles_from_reference=int(sd_counts.group(1)), n_singles_from_reference=int(sd_counts.group(2)),
                            core_correlation_energy=e_core, start_walkers=10000.0, tau=0.01,
                            projected_correlation_energy=float(re.search(r"Projected correlation energy\s+(-?[\d.]+)", bench2).group(1)),
                            projected_correlation_energy_error=float(re.search(r"Estimated error in Projected correlation energy\s+([\d.Ee+-]+)", bench2).group(1)),
                            nmcyc=400, total_walkers=10000, add_to_initiator=2.0, real_spawn_cutoff=0.01, steps_shift=1,
                            step1_proj_e=f[7], step1_no_at_hf=f[10], step1_no_at_doubs=f[11],
                            shift_damp=0.5, step2_shift=rows[2][0], step2_no_at_hf=rows[2][10], step2_no_at_doubs=rows[2][11],
                            step3_no_at_hf=rows[3][10]),
        doubles_core_size_hphf=core_hphf, doubles_core_size_determinants=core_dets,
        source="test_suite/neci/parallel/HeHe_SS_Doubles (FCIDUMP, neci.inp, benchmark.out...)",
        input=dict(hphf=True, allrealcoeff=True, realspawncutoff=0.01, tau=0.001, totalwalkers=1000, shiftdamp=0.1,
                   stepsshift=1, semi_stochastic="doubles-core", startsinglepart=10, diagshift=1.0, nmcyc=6000),
        norb=norb, nelec=nelec, ms2=ms2, orbsym=orbsym[:norb], ecore=ecore, eps=[eps[i] for i in range(1, norb + 1)],
        h1=h1, eri=eri,
        reference_det=[int(x) for x in ref_det.group(1).replace(",", " ").split()],
        reference_energy=ref_energy,
        highest_det=[int(x) for x in hi_det.group(1).split()], highest_det_energy=float(hi.group(1)),
        total_projected_energy=float(tot.group(1)), total_projected_energy_error=float(tot.group(2)),
    )
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hehe_ss_doubles.json")
    json.dump(out, open(dst, "w"))
    print("wrote", dst, "norb", norb, "nelec", nelec, "h1", len(h1), "eri", len(eri), out["reference_energy"],
          out["highest_det_energy"], out["total_projected_energy"], out["total_projected_energy_error"])


if __name__ == "__main__":
    main()

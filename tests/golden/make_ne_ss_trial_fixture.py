"""Generates tests/golden/ne_ss_trial.json: numbers the reference printed for its regression run
test_suite/neci/parallel/Ne_SS_Trial_Pops (first stage, neci-popsprint.inp: HPHF, `semi-stochastic mp1-core 50`,
`trial-wavefunction mp1-trial 200`, `freeze 2 0`).  Its FCIDUMP is the file of Ne_FciMCPar_pchb, whose frozen-core
integrals are already in tests/golden/ne_pchb.npz.  Run in the build container (/root/reference present)."""
import glob
import hashlib
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
CASE = os.path.join(REF, "test_suite", "neci", "parallel", "Ne_SS_Trial_Pops")
TWIN = os.path.join(REF, "test_suite", "neci", "parallel", "Ne_FciMCPar_pchb")


def main():
    md5 = lambda p: hashlib.md5(open(p, "rb").read()).hexdigest()
    assert md5(os.path.join(CASE, "FCIDUMP")) == md5(os.path.join(TWIN, "FCIDUMP"))
    bench = open(glob.glob(os.path.join(CASE, "benchmark*popsprint.inp"))[0]).read()
    out = dict(
        source="test_suite/neci/parallel/Ne_SS_Trial_Pops (benchmark.out...inp=neci-popsprint.inp)",
        input=dict(hphf=True, semi_stochastic="mp1-core 50", trial_wavefunction="mp1-trial 200", freeze=[2, 0]),
        reference_energy=float(re.search(r"Reference Energy set to:\s+(-?[\d.]+)", bench).group(1)),
        core_size=int(re.search(r"Total size of deterministic space:\s+(\d+)", bench).group(1)),
        core_correlation_energy=float(re.search(r"Deterministic subspace correlation energy:\s+(-?[\d.]+)", bench).group(1)),
        trial_size=int(re.search(r"Total size of the trial space:\s+(\d+)", bench).group(1)),
        trial_energy=float(re.search(r"Energy eigenvalue\(s\) of the trial space:\s+(-?[\d.]+)", bench).group(1)),
        connected_size=int(re.search(r"Total size of connected space:\s+(\d+)", bench).group(1)))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ne_ss_trial.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, out)


if __name__ == "__main__":
    main()

"""Generates tests/golden/ne_pchb.npz from the reference's regression case test_suite/neci/parallel/Ne_FciMCPar_pchb:
the FCIDUMP integrals with the case's `freeze 2 0` applied (the lowest orbital folded into ECore and the one-body
integrals, as src/Integrals_neci.F90 IntFreeze does) and the numbers the reference's CPU run printed (reference
determinant + energy, final projected energy with error).  Run in the build container (/root/reference present)."""
import glob
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
CASE = os.path.join(REF, "test_suite", "neci", "parallel", "Ne_FciMCPar_pchb")


def main():
    txt = open(os.path.join(CASE, "FCIDUMP")).read()
    head, body = txt.split("&END")
    norb = int(re.search(r"NORB\s*=\s*(\d+)", head).group(1))
    nelec = int(re.search(r"NELEC\s*=\s*(\d+)", head).group(1))
    orbsym = [int(x) for x in re.search(r"ORBSYM\s*=\s*([\d,\s]+)", head).group(1).replace("\n", "").split(",") if x.strip()][:norb]
    h = np.zeros((norb + 1, norb + 1)); eps = np.zeros(norb + 1); ecore = 0.0
    g = {}
    for ln in body.strip().splitlines():
        t = ln.split()
        if len(t) != 5:
            continue
        v = float(t[0]); i, j, k, l = (int(x) for x in t[1:])
        if i == 0: ecore = v
        elif j == 0: eps[i] = v
        elif k == 0: h[i, j] = h[j, i] = v
        else:
            for (a, b, c, d) in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k), (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
                g[(a, b, c, d)] = v

    def eri(a, b, c, d):                       # chemist (ab|cd)
        return g.get((a, b, c, d), 0.0)
    core = [int(np.argmin(eps[1:])) + 1]       # freeze 2 0: the two lowest spin orbitals = the lowest spatial orbital
    keep = [p for p in range(1, norb + 1) if p not in core]
    e2 = ecore
    for c in core:
        e2 += 2.0 * h[c, c]
        for c2 in core:
            e2 += 2.0 * eri(c, c, c2, c2) - eri(c, c2, c2, c)
    h1, eri_out = [], []
    for x, p in enumerate(keep, 1):
        for y, q in enumerate(keep, 1):
            if y > x: continue
            v = h[p, q] + sum(2.0 * eri(p, q, c, c) - eri(p, c, c, q) for c in core)
            if v != 0.0: h1.append((x, y, v))
    n = len(keep)
    for x in range(1, n + 1):
        for y in range(1, x + 1):
            for z in range(1, x + 1):
                for w in range(1, z + 1):
                    if (z, w) > (x, y) and z == x: continue
                    v = eri(keep[x - 1], keep[y - 1], keep[z - 1], keep[w - 1])
                    if v != 0.0: eri_out.append((x, y, z, w, v))
    bench = open(glob.glob(os.path.join(CASE, "benchmark*"))[0]).read()
    ref_det = [int(x) for x in re.search(r"Generated reference determinants:\s*\n\(\s*([\d,\s]+)\)", bench).group(1).replace(",", " ").split()]
    ref_energy = float(re.search(r"Reference Energy set to:\s+(-?[\d.]+)", bench).group(1))
    tot = re.search(r"Total projected energy\s+(-?[\d.]+)\s*\+/-\s*([\d.Ee+-]+)", bench)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ne_pchb.npz")
    np.savez_compressed(dst, norb=n, nelec=nelec - 2 * len(core), ecore=e2, orbsym=np.array([orbsym[p - 1] for p in keep]),
                        eps=np.array([eps[p] for p in keep]), h1=np.array(h1), eri=np.array(eri_out),
                        reference_det=np.array(ref_det), reference_energy=ref_energy,
                        total_projected_energy=float(tot.group(1)), total_projected_energy_error=float(tot.group(2)),
                        input_totalwalkers=20000, input_addtoinitiator=3.0, input_shiftdamp=0.03, input_stepsshift=25)
    print("wrote", dst, os.path.getsize(dst), "bytes; norb", n, "h1", len(h1), "eri", len(eri_out), ref_det, ref_energy, tot.group(1), tot.group(2))


if __name__ == "__main__":
    main()

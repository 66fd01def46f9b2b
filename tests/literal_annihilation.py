"""A second, structurally different restatement of the reference's annihilation step, in the reference's own form:
sort the spawned list and merge runs (CompressSpawnedList + FindResidualParticle, src/Annihilation.F90:249-634), then
walk the compressed list against the main list (AnnihilateSpawnedParts, :965-1352, stochRoundSpawn :1354-1418,
test_abort_spawn :1462-1479, AddNewHashDet src/load_balancer.fpp:514-629) and finish with CalcHashTableStats
(src/load_balancer.fpp:646-805).  Plain Python over small lists; the oracle (hash merge in C++) and the CUDA engine
(atomic hash merge) are checked against it.  Test infrastructure only.

Single run, lenof_sign = 1, no RDMs / GUGA / truncation / preconditioning: the `neci` defaults.
The main list is a dict det -> [sign, flags]: the reference's HashIndex + CurrentDets is an associative container and
nothing observable depends on slot numbers.
"""
import numpy as np

FLAG_INITIATOR = 13
FLAG_DETERMINISTIC = 19


def _signed(w):
    w = int(w)
    return w - (1 << 64) if w >= (1 << 63) else w


def compress_spawned_list(spawned, nw, t_trunc_initiator=True, t_init_coherent_rule=True):
    """spawned: list of (det tuple of nw ints, sign float, flags int).  Returns (compressed list, Annihilated)."""
    # call sort(SpawnedParts(...), ilut_lt, ilut_gt): signed integer compare, word 0 first (src/DetBitOps.F90:431-473)
    recs = sorted(spawned, key=lambda r: tuple(_signed(x) for x in r[0]))
    out = []
    annihilated = 0.0
    begin = 0
    n = len(recs)
    while begin < n:
        cur = begin + 1
        while cur < n and recs[cur][0] == recs[begin][0]:
            cur += 1
        end = cur - 1
        if end == begin:
            # a block of one entry is copied as it is, unless it carries no amplitude (:311-327)
            det, sgn, flg = recs[begin]
            if abs(sgn) >= 1.e-12:
                out.append((det, sgn, flg))
            begin = cur
            continue
        cum_sgn, cum_flg = 0.0, 0                        # cum_det = 0, orbital words copied (:379-380)
        for i in range(begin, end + 1):
            _, new_sgn, new_flg = recs[i]
            # FindResidualParticle (:551-634)
            new_init = (new_flg >> FLAG_INITIATOR) & 1
            if t_trunc_initiator:
                if t_init_coherent_rule:
                    if (abs(cum_sgn) > 1.e-12 and abs(new_sgn) > 1.e-12) or new_init:
                        cum_flg |= 1 << FLAG_INITIATOR
                elif new_init:
                    cum_flg |= 1 << FLAG_INITIATOR
            if cum_sgn * new_sgn < 0.0:
                annihilated += 2 * min(abs(cum_sgn), abs(new_sgn))
            cum_sgn = cum_sgn + new_sgn
        if abs(cum_sgn) > 1.e-12:                        # (:468-486)
            out.append((recs[begin][0], cum_sgn, cum_flg))
        begin = cur
    return out, annihilated


def annihilate_spawned_parts(main, compressed, draw_round, t_trunc_initiator=True, occupied_thresh=1.0):
    """main: dict det -> [sign, flags], modified in place.  draw_round(det) -> the uniform number stochRoundSpawn
    draws for that determinant.  Returns dict of the counters the routine updates."""
    st = dict(Annihilated=0.0, NoAborted=0.0, NoRemoved=0.0, NoBorn=0.0, inserted=0)
    for det, spawned_sign, flg in compressed:
        spawn_init = (flg >> FLAG_INITIATOR) & 1
        if det in main:                                   # hash_table_lookup: tSuccess
            cur = main[det]
            current_sign = cur[0]
            sign_prod = current_sign * spawned_sign
            determ = (cur[1] >> FLAG_DETERMINISTIC) & 1
            if abs(current_sign) >= 1.e-12 or determ:
                if abs(current_sign) < 1.e-12:            # is_run_unnocc: an empty core determinant
                    if t_trunc_initiator and not spawn_init and not determ:
                        st["NoAborted"] += abs(spawned_sign)
                        spawned_sign = 0.0
                if sign_prod < 0:
                    st["Annihilated"] += 2 * min(abs(current_sign), abs(spawned_sign))
                cur[0] = spawned_sign + current_sign
                if not determ and abs(cur[0]) < 1.e-12:
                    del main[det]                         # RemoveHashDet
            continue
        if t_trunc_initiator and not spawn_init:          # test_abort_spawn
            st["NoAborted"] += abs(spawned_sign)
            spawned_sign = 0.0
        if abs(spawned_sign) < 1.e-12:
            continue
        # stochRoundSpawn, scFVal = 1
        if 1.e-12 < abs(spawned_sign) < occupied_thresh:
            p_remove = 1.0 - abs(spawned_sign) / occupied_thresh
            if p_remove > draw_round(det):
                st["NoRemoved"] += abs(spawned_sign)
                spawned_sign = 0.0
            else:
                st["NoBorn"] += occupied_thresh - abs(spawned_sign)
                spawned_sign = float(np.copysign(occupied_thresh, spawned_sign))
        if abs(spawned_sign) >= 1.e-12:
            main[det] = [spawned_sign, flg & ~1]          # AddNewHashDet: the record as it is, flag_removed cleared
            st["inserted"] += 1
    return st


def calc_hash_table_stats(main, draw_prune, occupied_thresh=1.0):
    """The pass over the list after annihilation: stochastic removal below OccupiedThresh, TotParts, norms."""
    st = dict(NoRemoved=0.0, NoBorn=0.0, TotParts=0.0, norm_psi_squared=0.0, iHighestPop=0)
    for det in list(main.keys()):
        s, flg = main[det]
        determ = (flg >> FLAG_DETERMINISTIC) & 1
        if abs(s) < 1.e-12 and not determ:
            continue
        if not determ and 1.e-12 < abs(s) < occupied_thresh:
            p_remove = (occupied_thresh - abs(s)) / occupied_thresh
            if p_remove > draw_prune(det):
                st["NoRemoved"] += abs(s)
                del main[det]
                continue
            st["NoBorn"] += occupied_thresh - abs(s)
            s = float(np.copysign(occupied_thresh, s))
            main[det][0] = s
        st["TotParts"] += abs(s)
        st["norm_psi_squared"] += s * s
        if abs(s) > st["iHighestPop"]:
            st["iHighestPop"] = int(abs(s))
    return st

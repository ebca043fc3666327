"""ORACLE pinning — test infrastructure.

Checks the CPU oracle against every known-answer value the reference's own tests hold for the
hot path (/root/reference/test/runtests.jl:24-56, 165-169, 269-272, 446-450), reading the
reference's RASPA fixtures where they lie, and (re)generates the flattened input fixtures under
tests/golden/ that the GPU box uses (it has no /root/reference).

Run here (build container):   python oracle/pin_goldens.py [--write]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ceg_b200 as ceg  # noqa: E402
import oracle as orc  # noqa: E402

REF_RASPA = "/root/reference/test/raspa"
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# test/runtests.jl known-answer values (line numbers in the comments)
GOLDENS = {
    "na_cha_origin_vdw": -11083.13758653269,            # :44
    "na_cha_origin_coulomb": -1.8509402225092095e6,     # :45
    "ar_chana_min": -1841.0165092850448,                # :38  lattice index (52,50,91) of energy_grid(0.3)
    "two_co2_reciprocal": 1177.5215489122043 + 7.4133039820109055,  # :56
    "ar_cit7_point": -1043.35893781,                    # :169 (stale at ~2e-8)
    "na_cha_min": -1.9278944364761321e6,                # :50  (29,60,60); stale at 4e-5, rtol 1e-3 only
    "restart_baseline_co2": -14128.888030042883,        # :450 one CO2 in CHA+Na from the restart file (rtol 1e-3 there)
}


def tail_correction_pair(rule, cut):
    """interactions.jl:411-430, 646-652 — only needed to compare with the restart baseline, which includes the
    (out-of-scope) TailCorrection term of the Monte-Carlo setup."""
    from ceg_b200.forcefields import InteractionRuleSum, Kind
    if isinstance(rule, InteractionRuleSum):
        return sum(tail_correction_pair(r, cut) for r in rule.rules)
    if (not rule.tailcorrection) or rule.kind in (Kind.NoInteraction, Kind.HardSphere, Kind.CoulombEwaldDirect):
        return 0.0
    if rule.kind is Kind.LennardJones:
        e, sg = rule.params
        x3 = (sg / cut) ** 3
        return (4 / 3) * e * sg ** 3 * (x3 ** 3 / 3 - x3)
    if rule.kind is Kind.Buckingham:
        A, B, C = rule.params
        return A * np.exp(-B * cut) * (2.0 + B * cut * (2.0 + B * cut)) / B ** 3 - C / (3 * cut ** 3)
    raise ValueError(rule.kind)


def tail_correction_one_molecule(ff, fw, mol_names, cut=12.0):
    """tailcorrection.jl:23-79 for a framework plus ONE molecule of one species."""
    from ceg_b200.forcefields import get_atom_name
    n = len(ff.sdict)
    cnt = np.zeros(n, dtype=int)
    for sym in fw.atomic_symbol:
        cnt[ff.sdict[get_atom_name(str(sym))]] += 1
    lam = 2 * np.pi / abs(np.linalg.det(fw.mat))
    idx = [ff.sdict[a] for a in mol_names]
    present = [k for k in range(n) if cnt[k] > 0]
    ap = lambda i, j: tail_correction_pair(ff.interactions[i][j], cut) * lam   # noqa: E731
    value = sum(cnt[i] * cnt[j] * ap(i, j) for i in present for j in present)
    f = sum(ap(k1, k2) * cnt[k2] for k2 in present for k1 in idx)
    c = sum(ap(a, b) for a in idx for b in idx)
    return value + 2 * f + c


def save_inputs(name, inp, cset, extra=None):
    d = dict(mat=inp.mat, invmat=inp.invmat, ortho=inp.ortho, safemin2=inp.safemin2, positions=inp.positions,
             types=inp.types, charges=inp.charges if inp.charges is not None else np.zeros(0),
             n_types=inp.n_types, table=inp.table.view(np.uint8).reshape(inp.table.shape + (72,)),
             cutoff2=inp.cutoff2, want_coulomb=inp.want_coulomb, alpha=inp.alpha, supercell=np.array(inp.supercell),
             unit_mat=cset.cell.mat, spacing=cset.spacing)
    if extra:
        d.update(extra)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true", help="(re)write tests/golden/*.npz and goldens.json")
    args = ap.parse_args()
    ceg.setdir_RASPA(REF_RASPA)
    orc.build()
    results = {}
    ffname = "BoulfelfelSholl2021"
    ff = ceg.parse_forcefield_RASPA(ffname)

    # ---- Na+ in CHA_1.4_3b4eeb96 (runtests.jl:40-46) --------------------------------------------
    fw = ceg.load_framework_RASPA("CHA_1.4_3b4eeb96", ffname)
    oe = orc.OracleEwald(fw)
    inp = ceg.probe_inputs(fw, ff, atoms=("Na",), coulomb=True, alpha=oe.alpha)
    cset = ceg.GridCoordinatesSetup.from_cell(fw.mat, 0.15)
    na = ceg.raspa.default_system("Na", ffname)
    origin = np.zeros(3)
    vdw = orc.interpolate_sparse(inp, 0, cset, True, origin)
    direct = na.atomic_charge[0] * orc.interpolate_sparse(inp, 1, cset, False, origin)
    ctx = ceg.ewald_context(oe.as_framework(), ((na,),))
    recip = oe.reciprocal(origin[None, None, :], ctx.charges, ctx.energy_net_charges, ctx.static_contribution)[0]
    results["na_cha_origin_vdw"] = vdw
    results["na_cha_origin_coulomb"] = direct + recip
    print(f"Na@CHA origin: vdw {vdw!r}  direct {direct!r}  reciprocal {recip!r}")
    # minimum of energy_grid(0.3) at CartesianIndex(29,60,60) (runtests.jl:48-51)
    mat = fw.mat
    nums = [int(np.floor(np.linalg.norm(mat[:, c]) / 0.3)) + 1 for c in range(3)]
    p = (28 * mat[:, 0] / nums[0] + 59 * mat[:, 1] / nums[1]) + 59 * mat[:, 2] / nums[2]
    vmin = orc.interpolate_sparse(inp, 0, cset, True, p)
    dmin = na.atomic_charge[0] * orc.interpolate_sparse(inp, 1, cset, False, p)
    rmin = oe.reciprocal(p[None, None, :], ctx.charges, ctx.energy_net_charges, ctx.static_contribution)[0]
    results["na_cha_min"] = vmin + (dmin + rmin)
    if args.write:
        save_inputs("cha_na", inp, cset, dict(
            kmax=oe.kmax, krows=oe.rows, kfactors=oe.kfactors, S_re=oe.S.real, S_im=oe.S.imag, UIon=oe.UIon,
            net_charges_framework=oe.net_charges_framework, probe_charge=na.atomic_charge[0],
            site_frac=ceg.framework_sites(fw, oe.supercell)[0], site_charge=ceg.framework_sites(fw, oe.supercell)[1],
            super_invmat=oe.invmat))

    # ---- two CO2 reciprocal (runtests.jl:53-56) --------------------------------------------------
    co2 = ceg.load_molecule_RASPA("CO2", "TraPPE", ffname)
    m1 = co2.with_positions([[1.3, 2.9, 1.149], [1.3, 2.9, 0.0], [1.3, 2.9, -1.149]])
    m2 = co2.with_positions([[4.7, 10.1, 1.149], [4.7, 10.1, 0.0], [4.7, 10.1, -1.149]])
    results["two_co2_reciprocal"] = orc.compute_ewald(oe, ([m1, m2],))

    # ---- Ar in CHA_1.4_3b4eeb96_Na_11812 (runtests.jl:24-38) -------------------------------------
    fw2 = ceg.load_framework_RASPA("CHA_1.4_3b4eeb96_Na_11812", ffname)
    inp2 = ceg.probe_inputs(fw2, ff, atoms=("Ar",))
    cset2 = ceg.GridCoordinatesSetup.from_cell(fw2.mat, 0.15)
    mat2 = fw2.mat
    nums2 = [int(np.floor(np.linalg.norm(mat2[:, c]) / 0.3)) + 1 for c in range(3)]
    vals = {}
    for idx in [(52, 50, 91), (51, 51, 91), (56, 88, 49)]:
        p = ((idx[0] - 1) * mat2[:, 0] / nums2[0] + (idx[1] - 1) * mat2[:, 1] / nums2[1]) + (idx[2] - 1) * mat2[:, 2] / nums2[2]
        vals[idx] = orc.interpolate_sparse(inp2, 0, cset2, True, p)
    results["ar_chana_min"] = vals[(52, 50, 91)]
    print("Ar@CHA+Na minima:", vals)
    if args.write:
        save_inputs("chana_ar", inp2, cset2)

    # ---- restart baseline: one CO2 in CHA+Na (runtests.jl:446-450) -------------------------------
    rpos = np.array([[float(x) for x in line.split()[3:6]]
                     for line in open("/root/reference/test/CHA_1.4_3b4eeb96_Na_11812.restart")
                     if line.startswith("Adsorbate-atom-position:")])
    oe2 = orc.OracleEwald(fw2)
    names = [str(x) for x in co2.atomic_symbol]
    inp2c = ceg.probe_inputs(fw2, ff, atoms=("O_co2", "C_co2"), coulomb=True, alpha=oe2.alpha)
    gidx = {"O_co2": 0, "C_co2": 1}
    r_vdw = sum(orc.interpolate_sparse(inp2c, gidx[nm], cset2, True, p) for nm, p in zip(names, rpos))
    r_dir = sum(q * orc.interpolate_sparse(inp2c, 2, cset2, False, p) for q, p in zip(co2.atomic_charge, rpos))
    ctx2 = ceg.ewald_context(oe2.as_framework(), ((co2.with_positions(rpos),),))
    r_rec = oe2.reciprocal(rpos[None], ctx2.charges, ctx2.energy_net_charges, ctx2.static_contribution)[0]
    r_tail = tail_correction_one_molecule(ff, fw2, names)
    results["restart_baseline_co2"] = r_vdw + r_dir + r_rec + r_tail
    print(f"restart CO2: vdw {r_vdw!r} direct {r_dir!r} reciprocal {r_rec!r} tail correction {r_tail!r}")
    if args.write:
        save_inputs("chana_co2", inp2c, cset2, dict(
            site_charge=ceg.framework_sites(fw2, oe2.supercell)[1], restart_positions=rpos, co2_charges=np.asarray(co2.atomic_charge),
            oracle_vdw=r_vdw, oracle_direct=r_dir, oracle_reciprocal=r_rec, tail_correction=r_tail))

    # ---- Ar in triclinic CIT-7 (runtests.jl:165-169) ---------------------------------------------
    fw3 = ceg.load_framework_RASPA("CIT-7", ffname)
    assert ceg.find_supercell(fw3.mat, 12.0) == (2, 3, 3)
    inp3 = ceg.probe_inputs(fw3, ff, atoms=("Ar",))
    cset3 = ceg.GridCoordinatesSetup.from_cell(fw3.mat, 0.15)
    pt = np.array([-6.2437738548095165, 22.4072046579578092, 5.7557837121224642])
    results["ar_cit7_point"] = orc.interpolate_sparse(inp3, 0, cset3, True, pt)
    if args.write:
        fw3c = ceg.load_framework_RASPA("CIT-7", ffname)
        oe3 = orc.OracleEwald(fw3c)
        inp3c = ceg.probe_inputs(fw3c, ff, atoms=("Ar", "C_co2", "O_co2"), coulomb=True, alpha=oe3.alpha)
        save_inputs("cit7_multi", inp3c, cset3)

    # ---- blocking sphere across a periodic boundary (runtests.jl:269-272) -------------------------
    fw4 = ceg.load_framework_RASPA("CIT7block", ffname)
    cset4 = ceg.GridCoordinatesSetup.from_cell(fw4.mat, 0.15)
    blockpath = os.path.join(REF_RASPA, "structures", "block", "CIT7block.block")
    blk = ceg.BlockFile.from_mask(cset4, orc.parse_blockfile(blockpath, cset4))
    b1 = blk[np.array([12.5, 0.8, 0.3])]
    b2 = blk[np.array([20.175808361078516, 10.58027451750961, 9.185277912816744])]
    if args.write:
        ctr, r2 = ceg.coordinates.read_blockfile(blockpath, cset4)
        np.savez_compressed(os.path.join(GOLDEN_DIR, "cit7block.npz"), unit_mat=fw4.mat, spacing=0.15, centers=ctr, radius2=r2,
                            blocked_points=np.array([[12.5, 0.8, 0.3], [20.175808361078516, 10.58027451750961, 9.185277912816744]]))
    results["cit7block_blocked"] = [bool(b1), bool(b2)]
    assert b1 and b2, "blocking-sphere golden failed"

    ok = True
    for k, gold in GOLDENS.items():
        got = results[k]
        rel = abs(got - gold) / abs(gold)
        tol = {"na_cha_min": 1e-3, "ar_cit7_point": 1e-7, "restart_baseline_co2": 1e-6}.get(k, 1e-9)
        flag = "OK " if rel <= tol else "FAIL"
        ok &= rel <= tol
        print(f"{flag} {k:28s} oracle {got!r:26}  golden {gold!r:26}  rel {rel:.2e} (tol {tol:g})")
    if args.write:
        with open(os.path.join(GOLDEN_DIR, "goldens.json"), "w") as f:
            json.dump({"reference_values": GOLDENS, "oracle_values": {k: (float(v) if not isinstance(v, list) else v) for k, v in results.items()},
                       "source": "test/runtests.jl:24-56,165-169,269-272,446-450"}, f, indent=1)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Generate the committed golden fixtures from the reference's regression tests.

Run in the build container (needs /root/reference, scipy):
    python tests/golden/make_mesh_fixtures.py

For each regression case of the hot path it writes
  tests/golden/<case>.mesh.npz : the ExodusII (classic NetCDF CDF-2) mesh flattened to
       arrays -- coords, tet/tri connectivity (0-based, file order), element-block
       order, side sets (file-internal element ids + side ids, 0-based)
  tests/golden/<case>.diag.std : the reference's golden diagnostics file, verbatim
       (tests/regression/inciter/RieCG/<Case>/diag.std)
The .exo files cannot travel to the GPU box; these fixtures can.
"""
import os, shutil, sys
import numpy as np
from scipy.io import netcdf_file

REF = "/root/reference/tests/regression/inciter"
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "riecg_sod": ("RieCG/Sod/rectangle_01_1.5k.exo", "RieCG/Sod/diag.std"),
    "riecg_sedov": ("RieCG/Sedov/sedov_coarse.exo", "RieCG/Sedov/diag.std"),
    "riecg_taylor_green": ("RieCG/TaylorGreen/unitcube_1k.exo", "RieCG/TaylorGreen/diag.std"),
    "laxcg_bump": ("LaxCG/Bump/bump.exo", "LaxCG/Bump/diag.std"),
    "chocg_unitcube": ("ChoCG/Poisson/unitcube_01_1k.exo", None),
    "chocg_pidiv4": ("ChoCG/Poisson/unitcube_0pidiv4_1k.exo", None),
    "chocg_poiseuille": ("ChoCG/Poiseuille/poiseuille1tetz.exo", None),
    "sphere2_5k": ("ChoCG/Sphere/sphere2_5K.exo", None),
    "unitsquare_3_6k": ("ZalCG/SlotCyl/unitsquare_01_3.6k.exo", None),
    "riecg_canyon": ("RieCG/Canyon/canyon.exo", "RieCG/Canyon/diag.std"),
}
EXTRA_DIAG = {"laxcg_bump_hllc": "LaxCG/Bump/diag_hllc.std",
              "chocg_poisson_const": "ChoCG/Poisson/diag_poisson_const.std",
              "chocg_poisson_sine": "ChoCG/Poisson/diag_poisson_sine.std",
              "chocg_poisson_sine3": "ChoCG/Poisson/diag_poisson_sine3.std",
              "chocg_poisson_neumann": "ChoCG/Poisson/diag_poisson_neumann.std",
              "chocg_poiseuille_damp2": "ChoCG/Poiseuille/diag_poiseuille_damp2.std",
              "chocg_poiseuille_damp4": "ChoCG/Poiseuille/diag_poiseuille_damp4.std",
              "chocg_poiseuille_rk2": "ChoCG/Poiseuille/diag_poiseuille_rk2.std",
              "chocg_poiseuille_rk3": "ChoCG/Poiseuille/diag_poiseuille_rk3.std",
              "chocg_poiseuille_rk4": "ChoCG/Poiseuille/diag_poiseuille_rk4.std",
              "chocg_ldc": "ChoCG/Lid/diag_ldc.std",
              # vortical flow runs on the Taylor-Green case's mesh (same unitcube_1k.exo)
              "riecg_vortical_flow": "RieCG/VorticalFlow/diag.std",
              "riecg_vortical_flow_hllc": "RieCG/VorticalFlow/diag_hllc.std",
              "riecg_vortical_flow_stab2": "RieCG/VorticalFlow/diag_stab2.std",
              "riecg_vortical_flow_hllc_stab2": "RieCG/VorticalFlow/diag_hllc_stab2.std",
              "riecg_vortical_flow_steady": "RieCG/VorticalFlow/diag_steady.std",
              "kozcg_vortical_flow": "KozCG/VorticalFlow/diag.std",
              "riecg_nleg": "RieCG/NonlinearEnergyGrowth/diag.std",
              "riecg_rayleigh_taylor": "RieCG/RayleighTaylor/diag.std",
              "riecg_pipe": "RieCG/Pipe/diag.std",
              "lohcg_poiseuille_damp2": "LohCG/Poiseuille/diag_poiseuille_damp2.std",
              "lohcg_poiseuille_damp4": "LohCG/Poiseuille/diag_poiseuille_damp4.std",
              "lohcg_ldc": "LohCG/Lid/diag_ldc.std",
              "chocg_poiseuille_theta": "ChoCG/Poiseuille/diag_poiseuille_theta.std",
              "kozcg_nleg": "KozCG/NonlinearEnergyGrowth/diag.std",
              "kozcg_rayleigh_taylor": "KozCG/RayleighTaylor/diag.std",
              "zalcg_bump": "ZalCG/Bump/diag.std",
              "chocg_inviscid_sphere": "ChoCG/Sphere/diag_inviscid_sphere.std",
              "chocg_viscous_sphere": "ChoCG/Sphere/diag_sphere_chocg_viscous_test.std",
              "lohcg_viscous_sphere": "LohCG/Sphere/diag_sphere_lohcg_viscous_test.std",
              "chocg_sphere_point_src": "ChoCG/Sphere/diag_sphere_point_src.std",
              "riecg_rayleigh_taylor_st": "RieCG/RayleighTaylor/diag_st.std",
              "kozcg_rayleigh_taylor_st": "KozCG/RayleighTaylor/diag_st.std",
              "riecg_canyon_farfield": "RieCG/Canyon/diag_farfield.std",
              "riecg_slot_cyl": "RieCG/SlotCyl/diag.std",
              "zalcg_slot_cyl": "ZalCG/SlotCyl/diag.std",
              "kozcg_slot_cyl": "KozCG/SlotCyl/diag.std",
              "chocg_slot_cyl": "ChoCG/SlotCyl/diag.std",
              "chocg_slot_cyl_damp4": "ChoCG/SlotCyl/diag_damp4.std",
              "lohcg_slot_cyl": "LohCG/SlotCyl/diag.std",
              "lohcg_slot_cyl_damp4": "LohCG/SlotCyl/diag_damp4.std",
              "chocg_slot_cyl_damp4_freeze": "ChoCG/SlotCyl/diag_damp4_freeze.std",
              "riecg_sod_hist_range": "RieCG/Sod/diag_hist_range.std",
              "zalcg_bump_fctfreeze": "ZalCG/Bump/diag_fctfreeze.std"}


def flatten(exo):
    f = netcdf_file(exo, "r", mmap=False)
    v = f.variables
    coord = np.stack([v["coordx"][:], v["coordy"][:], v["coordz"][:]]).astype(np.float64)
    nblk = f.dimensions["num_el_blk"]
    tets, tris, btype, bn = [], [], [], []
    for b in range(1, nblk + 1):
        c = v[f"connect{b}"]
        et = c.elem_type.decode().upper()
        conn = np.asarray(c[:], dtype=np.int64) - 1
        if et.startswith("TET"):
            tets.append(conn); btype.append(1)
        elif et.startswith("TRI"):
            tris.append(conn); btype.append(0)
        else:
            raise RuntimeError("unsupported element type " + et)
        bn.append(conn.shape[0])
    tets = np.concatenate(tets) if tets else np.zeros((0, 4), np.int64)
    tris = np.concatenate(tris) if tris else np.zeros((0, 3), np.int64)
    ids = np.asarray(v["ss_prop1"][:], dtype=np.int32)
    off, elem, side = [0], [], []
    for s in range(1, len(ids) + 1):
        e = np.asarray(v[f"elem_ss{s}"][:], dtype=np.int64) - 1
        sd = np.asarray(v[f"side_ss{s}"][:], dtype=np.int64) - 1
        elem.append(e); side.append(sd); off.append(off[-1] + len(e))
    return dict(coord=coord, tets=tets.astype(np.uint64), tris=tris.astype(np.uint64),
                block_type=np.asarray(btype, np.int32), block_n=np.asarray(bn, np.uint64),
                set_id=ids, set_off=np.asarray(off, np.uint64),
                set_elem=np.concatenate(elem).astype(np.uint64),
                set_side=np.concatenate(side).astype(np.uint64))


def main():
    for name, (exo, diag) in CASES.items():
        m = flatten(os.path.join(REF, exo))
        np.savez_compressed(os.path.join(HERE, name + ".mesh.npz"), **m)
        if diag:
            shutil.copyfile(os.path.join(REF, diag), os.path.join(HERE, name + ".diag.std"))
            os.chmod(os.path.join(HERE, name + ".diag.std"), 0o644)
        print(name, m["coord"].shape[1], "nodes", len(m["tets"]), "tets", len(m["tris"]), "tris")
    # one small mesh file verbatim, for the test of the product's own ExodusII reader
    shutil.copyfile(os.path.join(REF, "RieCG/Sod/rectangle_01_1.5k.exo"), os.path.join(HERE, "riecg_sod.exo"))
    os.chmod(os.path.join(HERE, "riecg_sod.exo"), 0o644)
    for name, diag in EXTRA_DIAG.items():
        shutil.copyfile(os.path.join(REF, diag), os.path.join(HERE, name + ".diag.std"))
        os.chmod(os.path.join(HERE, name + ".diag.std"), 0o644)


if __name__ == "__main__":
    sys.exit(main())

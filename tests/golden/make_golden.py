"""Copies the reference's golden CSV numbers into tests/golden/reference_csv.json.
Run in the build container (reads /root/reference, which does not exist on the GPU box)."""
import csv
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_csv.json")


def rows(path):
    with open(os.path.join(REF, path)) as f:
        return [[float(v) for v in r] for r in list(csv.reader(f))[1:]]


def main():
    ip_rows = rows("StefanDG/InteriorPenalty/uncut_mesh_tests/IP_convergence.csv")
    ldg_rows = rows("StefanDG/LocalDG/uncut_mesh_tests/MD-LDG_convergence.csv")
    robin = rows("StefanRobinIC/DG1D/robin-interface-tests/robin-convergence.csv")
    cyl = rows("StefanDG/InteriorPenalty/cylindrical_bc_tests/IP_convergence.csv")
    cylr = rows("StefanDG/InteriorPenalty/robin_cylindrical_bc_tests/IP_convergence.csv")
    ldgcyl = rows("StefanDG/LocalDG/cylindrical_bc_tests/LDG_convergence.csv")
    cut = {
        "ip_plane": "StefanDG/InteriorPenalty/plane_interface_tests/IP_convergence.csv",
        "ip_curved": "StefanDG/InteriorPenalty/curved_interface_tests/IP_convergence.csv",
        "ldg_plane": "StefanDG/LocalDG/plane_interface_tests/LDG_convergence.csv",
        "ldg_curved": "StefanDG/LocalDG/curved_interface_tests/LDG_convergence.csv",
    }
    gold = {
        "ip_uncut": [[int(r[0]), r[1], r[2]] for r in ip_rows],
        "ldg_uncut": [[int(r[0])] + r[1:] for r in ldg_rows],
        "dg1d_robin_linear": [r[1] for r in robin],
        "dg1d_robin_quadratic": [r[2] for r in robin],
        # cut-cell studies: NElmts, linear, quadratic (relative L2 errors); parity at discretisation level only
        "ip_cylinder": [[int(r[0]), r[1], r[2]] for r in cyl],
        "ip_cylinder_robin": [[int(r[0]), r[1], r[3]] for r in cylr],
        # NElmts, then errT, errG1, errG2 (relative L2) for the linear and the quadratic basis
        "ldg_cylinder": [[int(r[0])] + r[1:] for r in ldgcyl],
        # cos(2 pi x) sin(2 pi y) on the unit square cut by a straight line (60 degrees through (0.8, 0)) / by the circle
        # of radius 0.5 around (1, 1), k1 = k2 = 1.  IP: NElmts, ABSOLUTE L2 error linear, quadratic.  LDG: NElmts, then
        # errT (relative), errG1, errG2 (absolute) for the linear and the quadratic basis
        **{k: [[int(r[0])] + r[1:] for r in rows(v)] for k, v in cut.items()},
        "sources": {
            **cut,
            "ldg_cylinder": "StefanDG/LocalDG/cylindrical_bc_tests/LDG_convergence.csv",
            "ip_cylinder": "StefanDG/InteriorPenalty/cylindrical_bc_tests/IP_convergence.csv",
            "ip_cylinder_robin": "StefanDG/InteriorPenalty/robin_cylindrical_bc_tests/IP_convergence.csv",
            "ip_uncut": "StefanDG/InteriorPenalty/uncut_mesh_tests/IP_convergence.csv",
            "ldg_uncut": "StefanDG/LocalDG/uncut_mesh_tests/MD-LDG_convergence.csv",
            "dg1d_robin": "StefanRobinIC/DG1D/robin-interface-tests/robin-convergence.csv",
        },
    }
    with open(OUT, "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

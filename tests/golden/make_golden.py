"""Generates tests/golden/*.json from the CPU oracle (oracle/nls_oracle.c).

The reference is pure Julia and `julia` is not installed in this image, so fixtures cannot be
produced by running the reference. What is committed here:
  * reference_kats.json  -- known answers DERIVED FROM THE REFERENCE'S FORMULAS by hand / in the
    survey; written as literals, not computed by the oracle.
  * oracle_cases.json    -- small-case oracle outputs (after the oracle passed the KATs and the
    independent torch-autograd check in tests/test_oracle.py), used as regression vectors for the
    oracle and as inputs/outputs for the GPU parity tests.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import __graft_entry__ as g
    nls = g.load_package()
    from oracle import oracle as orc
    from helpers import distorted, grid_problem

    kats = {
        "source": "SURVEY.md 8c; derived from src/materials/defaults.jl:2-26, exponential.jl:20-21, "
                  "shapefunctions.jl:23-81 (reference formulas), not from running Julia",
        "KAT3": {"L": 1.0, "sigma_sat": 1.0, "b": 2.0, "A": 1.0, "tip_force": 0.5, "rtol": 1e-10,
                 "tip": 0.34657359027997264, "iterations": 5, "reason": "CONVERGED_RELATIVE",
                 "hist": [0.5, 0.10653065971263342, 0.0088309880370516, 7.6191446373e-05, 5.80395731e-09],
                 "meshes": [[4, 1, 2], [3, 2, 3]]},
        "KAT4": {"p": 2, "center_node": 6.123233995736766e-17,
                 "weights": [0.5, -1.0000000000000002, 0.5000000000000001],
                 "dN_at_0": [-0.5000000000000001, 0.0, 0.5000000000000001]},
    }
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(kats, f, indent=1)

    cases = {}
    mat = nls.NeoHookean(E=1.0, nu=0.3)
    for name, dim, n, kind in [("q1_2d_n4_distorted", 2, 4, "distorted"), ("q1_3d_n3_distorted", 3, 3, "distorted"),
                               ("q1_2d_n5_affine", 2, 5, "affine")]:
        fem, om = grid_problem(nls, orc, n, dim, mat, compress=-0.2 / (n - 1))
        if kind == "distorted":
            X = distorted(fem.mesh)
            fem.mesh.coords[:] = X
            om.coords[:] = X
        U = nls.smooth_field(fem.mesh.coords, amp=0.02)
        rp, ci = om.pattern()
        Uo, R = om.residual(U)
        vals = om.jacobian_csr(U)
        cases[name] = dict(dim=dim, n=n, mu=mat.mu, lam=mat.lam, coords=fem.mesh.coords.ravel().tolist(),
                           conn=fem.mesh.conn.ravel().tolist(), dir_dofs=om.dir_dofs.tolist(),
                           dir_vals=om.dir_vals.tolist(), U=U.tolist(), U_out=Uo.tolist(), R=R.tolist(),
                           rowptr=rp.tolist(), colind=ci.tolist(), vals=vals.tolist())
    # 1-D bar, exponential law, p=3, q=4
    mesh = nls.bar_mesh(5, 3)
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=3, nq=4, mat_kind=orc.MAT_EXPONENTIAL, mat=(1.0, 2.0), area=0.7)
    om.set_dirichlet([0], [0.0]); om.set_neumann([mesh.nn - 1], [0.5])
    U = np.random.default_rng(7).uniform(-0.05, 0.05, mesh.nn)
    rp, ci = om.pattern()
    Uo, R = om.residual(U)
    cases["bar_exp_p3_q4"] = dict(dim=1, nel=5, p=3, q=4, area=0.7, U=U.tolist(), U_out=Uo.tolist(), R=R.tolist(),
                                  rowptr=rp.tolist(), colind=ci.tolist(), vals=om.jacobian_csr(U).tolist())
    with open(os.path.join(HERE, "oracle_cases.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()

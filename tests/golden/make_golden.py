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

    # dynamics and continuation: mass matrix, Newmark histories, arc-length path
    dyn = {}
    c = cases["q1_2d_n4_distorted"]
    om = orc.OracleModel(2, np.array(c["conn"], dtype=np.int32).reshape(-1, 4), np.array(c["coords"]).reshape(-1, 2),
                         mat_kind=orc.MAT_NEOHOOKEAN, mat=(c["mu"], c["lam"]))
    dyn["mass_q1_2d_n4_distorted"] = dict(case="q1_2d_n4_distorted", rho=1.7, vals=om.mass_csr(1.7).tolist())
    p, q, nel, rho, A = 2, 3, 4, 2.0, 0.5
    mesh = nls.bar_mesh(nel, p)
    om = orc.OracleModel(1, mesh.conn, mesh.coords, p=p, nq=q, mat_kind=orc.MAT_EXPONENTIAL, mat=(1.0, 2.0), area=A)
    om.set_dirichlet([0], [0.0]); om.set_neumann([mesh.nn - 1], [0.0])
    out = om.newmark(rho, 0.05, 0.3, area=A, bc_update=lambda t: (None, [0.2 * t]), rtol=1e-11, atol=1e-13, maxits=20)
    dyn["newmark_bar_exp_p2_q3"] = dict(nel=nel, p=p, q=q, rho=rho, area=A, dt=0.05, maxtime=0.3, tip_rate=0.2,
                                        U=out["U"].tolist(), Ud=out["Ud"].tolist(), Udd=out["Udd"].tolist(),
                                        num_its=out["num_its"].tolist())
    nx, ny, R, th0, t = 25, 3, 10.0, 0.25, 0.08
    m = nls.grid_mesh((nx, ny), 2)
    s_, yy = m.coords[:, 0], m.coords[:, 1] * (nx - 1) / (ny - 1)
    th, r = (2 * s_ - 1) * th0, R + (yy - 0.5) * t
    X = np.stack([r * np.sin(th), r * np.cos(th) - R * np.cos(th0)], axis=1)
    E, nu = 1000.0, 0.3
    mu, lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    om = orc.OracleModel(2, m.conn, X, mat_kind=orc.MAT_NEOHOOKEAN, mat=(mu, lam))
    ends = np.concatenate([np.arange(ny) * nx, np.arange(ny) * nx + nx - 1])
    dofs = (ends[:, None] * 2 + np.arange(2)[None, :]).ravel()
    om.set_dirichlet(dofs, np.zeros(len(dofs)))
    apex = (ny - 1) * nx + nx // 2
    F = np.zeros(om.ndof); F[2 * apex + 1] = -0.05
    out = orc.newtonarclength(om, F, b=0.5, da=0.4, rtol=1e-7, ftol=1e-7, maxsteps=20, maxits=30)
    dyn["arclength_arch_25x3"] = dict(nx=nx, ny=ny, R=R, th0=th0, t=t, E=E, nu=nu, load=-0.05, b=0.5, da=0.4, rtol=1e-7, ftol=1e-7,
                                      maxsteps=20, maxits=30, lam=out["lam"].tolist(), apex_v=out["d"][:, 2 * apex + 1].tolist(),
                                      num_its=out["num_its"].tolist())
    with open(os.path.join(HERE, "oracle_dynamics_cases.json"), "w") as f:
        json.dump(dyn, f)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()

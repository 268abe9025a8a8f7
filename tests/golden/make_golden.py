"""Writes tests/golden/*.npz: small inputs and the ORACLE's outputs on them.

The reference (R) cannot be run here and ships no golden vectors, so these
files do not pin the oracle to the reference; they pin the oracle -- and through it the CUDA
path -- against regressions over time.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dcar, glm, graphs, hcar  # noqa: E402
from tests import helpers  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # direct CAR, 10 x 10 torus, N = 120 (the vignette's geometry)
    odata, psi0, sd, rng = helpers.torus_case(10, 10, 120, seed=33)
    psis = psi0[None, :] * (1 + 0.04 * rng.standard_normal((6, 4)))
    np.savez_compressed(
        os.path.join(HERE, "dcar_torus10.npz"), y=odata["y"], X=odata["X"], psi0=psi0, simY=sd["simY"], t10=sd["t10"], t20=sd["t20"],
        psis=psis, lr_var=dcar.eval_batch_dCAR(psis, odata, sd, rho_cons=None),
        grad=np.stack([dcar.mc_gradlik_dCAR(p, odata, sd) for p in psis]),
        gradvar2=np.stack([dcar.mc_gradlik_dCAR(p, odata, sd, Evar=2)["grad.var"] for p in psis]),
        hess=np.stack([dcar.mc_Hesslik_dCAR(p, odata, sd) for p in psis]),
        mcmle_var=np.stack([dcar.MCMLE_var(p, odata, sd) for p in psis]))
    # binomial GLM, 6 x 6 torus
    rng = np.random.default_rng(7)
    W = graphs.torus_W(6, 6)
    X = np.column_stack([np.ones(36), rng.standard_normal(36) * 0.5])
    psi = np.array([0.15, 1.2, 0.4, -0.3])
    for fam in ("binom", "poisson"):
        data = glm.sim_glm_data(fam, W, psi, X, 10, rng)
        Zy = data["Z.true"][None, :] + 0.4 * rng.standard_normal((150, 36))
        mc = glm.add_prep(psi, data, Zy, fam)
        ps = psi[None, :] * (1 + 0.03 * rng.standard_normal((4, 4)))
        ew = np.linalg.eigvalsh(W.toarray())
        np.savez_compressed(
            os.path.join(HERE, f"glm_{fam}_torus6.npz"), y=data["y"], X=X, n_trial=10.0, psi0=psi, Zy=Zy, lHZy=mc["lHZy.psi"], psis=ps,
            lr_var=glm.eval_batch_glm(ps, mc, fam, clamp=False), grad=np.stack([glm.mcl_grad_glm(p, mc, fam, ew=ew) for p in ps]),
            hess=np.stack([glm.mcl_Hessian_glm(p, mc, fam, ew=ew) for p in ps]))
    # HCAR with W
    psi = np.array([0.1, 0.12, 1.0, 0.5, 1.0, 0.5])
    data = hcar.make_hcar_data(9, 3, psi, 2, rng, with_W=True)
    mc = hcar.sim_HCAR(psi, data, 150, rng)
    ps = psi[None, :] * (1 + 0.03 * rng.standard_normal((4, 6)))
    np.savez_compressed(
        os.path.join(HERE, "hcar_w9.npz"), y=data["y"], X=data["X"], W=data["W"].toarray(), M=data["M"].toarray(), Z=data["Z"],
        psi0=psi, u_y=mc["u.y"], lpsi=mc["lpsi"], const_psi=mc["const.psi"], psis=ps,
        lr_var=np.stack([hcar.mcl_HCAR(p, mc, data, Evar=True) for p in ps]), grad=np.stack([hcar.mcl_grad_HCAR(p, mc, data) for p in ps]),
        hess=np.stack([hcar.mcl_Hessian_HCAR(p, mc, data) for p in ps]))


if __name__ == "__main__":
    main()

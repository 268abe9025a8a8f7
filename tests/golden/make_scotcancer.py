"""Extracts the reference's only real-data fixture, data/ScotCancer.rda (Scottish lip cancer, 56
districts: response y, model matrix covX = [1, Paff], neighbourhood matrix W and the stored
parameter psi = (rho, sigma^2, beta)), into tests/golden/scotcancer.npz.  Runs only where
/root/reference exists (this container); the .npz travels with the repository.

    python tests/golden/make_scotcancer.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from rda_reader import load_rda  # noqa: E402

SRC = "/root/reference/mclcar_0.2-0/data/ScotCancer.rda"


def main():
    (name, val), = load_rda(SRC)
    assert name == "ScotCancer"
    fields = dict(zip(val["attr"]["names"], val["value"]))

    def arr(x):
        a = np.array(x["value"], dtype=np.float64)
        return a.reshape(x["attr"]["dim"], order="F")      # R matrices are column-major

    out = {"y": np.array(fields["y"], dtype=np.float64), "covX": arr(fields["covX"]), "W": arr(fields["W"]),
           "psi": np.array(fields["psi"], dtype=np.float64)}
    assert out["y"].shape == (56,) and out["covX"].shape == (56, 2) and out["W"].shape == (56, 56) and out["psi"].shape == (4,)
    np.savez_compressed(os.path.join(HERE, "scotcancer.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

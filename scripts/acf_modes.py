import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from mclcar_b200 import api
n1 = n2 = 96; n = n1*n2
ii, jj = np.divmod(np.arange(n), n2)
X = np.column_stack([np.ones(n), np.where((ii + jj) % 2 == 0, 1.0, -1.0)])
data = api.make_data(np.zeros(n), X=X, torus=(n1, n2), seed=77)
C = 512; E = 400
for om in (0.0, 1.25, 1.3, 1.0):
    sd = api.mcl_prep_dCAR([0.2, 1.0, 0.0, 0.0], C * E, data, {"n.chains": C, "keep": False, "thin": 1, "omega": om})
    d = api.sampler_diag(data, sd)
    q = sd.stats()
    m = q[:, [0, 1, 2, 3]].reshape(E, C, 4)
    m = m - m.mean(axis=(0, 1), keepdims=True)
    v = (m * m).mean(axis=(0, 1))
    acf = np.array([(m[k:] * m[:E - k]).mean(axis=(0, 1)) / v for k in range(1, 8)])
    print("omega", d["omega"], "cols: x'x, x'Wx, const mode, checker mode")
    print(np.round(acf, 4))

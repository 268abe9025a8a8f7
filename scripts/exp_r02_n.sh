set -u
O=gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2_pytest15.log 2>&1; tail -3 $O/r2_pytest15.log
for lin in 1 0; do MCLCAR_STATS_LINEAR=$lin python bench.py --config c2 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('c2 linear=$lin', '%.3f ms/step value %.4g' % (d['ms_per_step'], d['value']), d['kernels'])"; done
python - <<'PY'
# config 5 geometry with p = 2: statistics pass over 600 resident fp32 samples, both forms
import os, time, numpy as np, ctypes as C
from mclcar_b200 import api, _lib as L
n1 = n2 = 1000; n = n1 * n2
rng = np.random.default_rng(0)
X = np.column_stack([np.ones(n), rng.permutation(np.log(np.arange(1, n + 1)) / 5.0)])
y = rng.standard_normal(n)
for lin in ("1", "0"):
    os.environ["MCLCAR_STATS_LINEAR"] = lin
    d = api.make_data(y, X=X, torus=(n1, n2), seed=1)
    sd = api.mcl_prep_dCAR([0.2, 1.0, 1.0, 0.5], 600, d, {"n.chains": 100, "burn": 2, "thin": 1})
    ms, ln = np.zeros(8), np.zeros(8, dtype=np.int64)
    L.check(L.lib.mclcar_ctx_profile(d.ctx.h, 1), d.ctx.h)
    q = sd.stats()
    L.check(L.lib.mclcar_ctx_profile_read(d.ctx.h, L.dptr(ms), L.i64ptr(ln)), d.ctx.h)
    print("c5 p=2 statistics pass, linear=%s: %.3f ms for 600 samples = %.0f GB/s" % (lin, ms[2], 600 * n * 4 / ms[2] / 1e6), q[0, :3])
PY

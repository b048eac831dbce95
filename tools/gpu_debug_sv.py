import sys
sys.path.insert(0, ".")
import numpy as np
from bayesianvars_b200 import Handle
from bayesianvars_b200 import bvar as B
from tests.test_gpu_gibbs import setup
h = Handle(0, seed=20240229)
Y, A, pp, ps, prep = setup(4, 1, 120, "HS")
smp = B.Sampler(h, prep, pp, ps, draws=3000, burnin=1000)
done = 0
prev = None
CH = int(sys.argv[1]) if len(sys.argv) > 1 else 25
while done < 4000:
    try:
        done = smp.run(CH)
    except Exception as e:
        print("FAILED by sweep", done + CH, str(e)[:160])
        break
    st = smp.state()
    bad = {k: (not np.isfinite(v).all()) for k, v in st.items()}
    lv = st["logvar"]
    if any(bad.values()) or np.abs(lv).max() > 40 or np.abs(st["PHI"]).max() > 50:
        print("suspicious at sweep", done, bad, "logvar range", lv.min(), lv.max(), "PHI max", np.abs(st["PHI"]).max(), "sv_para", st["sv_para"])
        break
    prev = st
print("done", done)
if prev is not None:
    print("last good logvar range", prev["logvar"].min(), prev["logvar"].max(), "sv_para\n", prev["sv_para"], "\nfacload", prev["facload"].ravel())

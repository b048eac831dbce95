"""tests/golden/usmacro_growth.npz from the reference's dataset (data/usmacro_growth.rda, 247 x 21 quarterly growth
rates, FRED-QD), decoded with bayesianvars_b200.rdata.  /root/reference does not exist on the GPU box, so the decoded
matrix travels as a fixture; tests/test_rdata.py re-decodes the .rda here (when the reference is present) and checks
the fixture against it.   python tools/make_golden_usmacro.py [path/to/usmacro_growth.rda]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bayesianvars_b200.rdata import read_rda  # noqa: E402

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/usmacro_growth.rda")
m = read_rda(src)["usmacro_growth"]
assert m.values.shape == (247, 21) and not np.isnan(m.values).any()
np.savez_compressed(ROOT / "tests" / "golden" / "usmacro_growth.npz", values=m.values, colnames=np.array(m.colnames),
                    rownames=np.array(m.rownames))
print("wrote tests/golden/usmacro_growth.npz", m.values.shape)

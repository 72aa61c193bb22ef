"""CPU: what tests/test_gpu_zzz_full_size.py should measure on hardware, predicted by running the same checks
(tests/util.full_size_checks) over the rounding-aware emulation of the kernel wrappers (tests/emu.py rounds to bf16 exactly
where the kernels store bf16; it reproduced smoke()'s hardware figure of 4.45e-3 to three digits).

    PYTHONPATH=. python tools/predict_full_size.py bond     # C2: 10 k molecules, ~3 min on 8 cores
    PYTHONPATH=. python tools/predict_full_size.py atom     # C4: 10 k reaction graphs, oracle included, ~10 min
"""
import sys
import time

import pytest

from tests import emu
from tests.util import full_size_checks

kind = sys.argv[1] if len(sys.argv) > 1 else "bond"
with pytest.MonkeyPatch.context() as m:
    emu.patch_engine(m)
    t0 = time.time()
    if kind == "bond":
        out = full_size_checks("bond", 10000, "cpu", gen_kw=dict(seed=1, mean_atoms=25.0, pool=30000), grad_tol=1.0)
    else:
        out = full_size_checks("atom", 10000, "cpu", gen_kw=dict(seed=1, cgr=True), grad_tol=1.0)
    print(kind, f"{time.time() - t0:.0f} s", out)

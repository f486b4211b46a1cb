"""A fixed-seed slice of the randomised parity sweep (tools/fuzz_parity.py) inside the GPU suite: shapes drawn across
the kernels' whole envelope (tile / query-tile edges, odd dims, k 1..128, masks, doc-id bases, f16/f32 sparse inputs,
every impl request), every entry path that must agree, the device BM25 scorer and the training-time scorer - each
checked against its CPU oracle with the assertions of tests/parity.py.  The full sweeps (thousands of cases) are run
with the tool itself; their results are under profiles/."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fuzz_parity as F  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,seed,n_cases", [("kernels", 101, 80), ("kernels", 102, 80), ("api", 103, 10),
                                               ("bm25", 104, 20), ("train", 105, 60)])
def test_random_shapes_against_the_oracle(mode, seed, n_cases):
    draw, run = F.MODES[mode]
    rng = np.random.RandomState(seed)
    for n in range(n_cases):
        case = draw(rng)
        try:
            run(case)
        except Exception as e:  # noqa: BLE001
            raise AssertionError(f"{mode} case {n} {case}: {type(e).__name__}: {e}") from e

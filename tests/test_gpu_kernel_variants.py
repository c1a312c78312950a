"""Every kernel variant the library can select (environment switches read once per process) runs the hierarchy / V-cycle / PCG
parity tests against the oracle in its own process: the element-form, cp.async neighbour-form and TMA-staged neighbour-form
level-0 smoothers on short and long rows, the per-colour fallback, the persistent small-level sweep, and the V-cycle with a separate residual kernel (the default takes the residual from the pre-smoothing sweep)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELECT = "test_hierarchy_operators or test_vcycle_matches_oracle or test_pcg_parity or test_masked_operators or test_masked_pcg"

VARIANTS = {
    "per_colour_level0_smoother": {"VF_GS_ROWS": "0"},
    "rows_element_form": {"VF_GS_ROWS": "2", "VF_GS_NB": "0"},
    "rows_element_form_general_table": {"VF_GS_ROWS": "2", "VF_GS_NB": "0", "VF_GS_ISO": "0"},
    "rows_neighbour_form_cp_async": {"VF_GS_ROWS": "2", "VF_GS_TMA": "0"},
    "rows_neighbour_form_tma": {"VF_GS_ROWS": "2"},
    "persistent_small_level_sweep": {"VF_SWEEP_FUSED_NODES": "100000"},
    "no_programmatic_dependent_launch": {"VF_PDL": "0"},
    "dense_level0_apply": {"VF_L0_DENSE": "1"},
    "one_shot_galerkin_coarsening": {"VF_COARSEN_ONESHOT": "1"},
    "dense_coarse_factorization": {"VF_COARSE_DENSE": "1"},
    "separate_residual_kernel": {"VF_GS_RESIDUAL": "0"},
    "level1_coarsening_with_table_loads": {"VF_COARSEN_PARAMK": "0"},
    "tile_coordinates_without_position_table": {"VF_ST_POSTAB": "0"},
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_parity(name):
    env = dict(os.environ); env.update(VARIANTS[name])
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-k", SELECT],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    assert r.returncode == 0, "%s %s\n%s" % (name, VARIANTS[name], tail)
    assert " passed" in tail and "failed" not in tail, tail

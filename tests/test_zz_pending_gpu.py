"""GPU legs of the compositions written after round 1's GPU minutes were spent (Heisenberg-picture path, inner between different
bond dimensions, observables over a Steiner tree, multi-site density matrices, single-layer contraction, the plain-C demo).

STATUS: the host code composes entry points that tests/test_gpu_parity.py covers one by one, and its Python side runs on the CPU
against an oracle-backed stand-in (tests/test_host_logic.py), but these compositions have not run on a B200 yet.  Therefore:
  * every body (tests/zz_bodies_gpu.py) runs in a CHILD process with a time-out: a crash or a hang cannot take the verified suite down;
  * the wrappers are non-strict xfail: a pass shows up as XPASS in the round-end log, a failure does not turn the suite red;
  * the file sorts last.
Round 2: drop the xfail mark of every XPASS and move the body into tests/test_gpu_parity.py."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BODIES = ["test_heisenberg_example_matches_golden", "test_inner_between_different_bond_dimensions_on_a_tree",
          "test_multi_site_observables_over_the_steiner_tree", "test_two_and_three_site_reduced_density_matrices",
          "test_plain_c_program_through_the_abi", "test_single_layer_network_contraction"]


def test_every_body_is_listed():
    import ast
    tree = ast.parse(open(os.path.join(HERE, "zz_bodies_gpu.py")).read())
    assert sorted(n.name for n in tree.body if isinstance(n, ast.FunctionDef) and n.name.startswith("test_")) == sorted(BODIES)


@pytest.mark.gpu
@pytest.mark.xfail(reason="composition first written after the round's GPU budget was spent; not yet run on a B200", strict=False)
@pytest.mark.parametrize("name", BODIES)
def test_pending_gpu_body_in_a_child_process(name):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "zz_bodies_gpu.py") + "::" + name, "-q", "-x", "-m", "gpu",
                        "-p", "no:cacheprovider"], cwd=os.path.dirname(HERE), capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, (r.stdout[-3000:] + r.stderr[-2000:])

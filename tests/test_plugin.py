"""The acados plugin (acados_b200/plugin/ocp_qp_cuipm.c, plain C): CPU -- it builds against the reference headers and
exports the full qp_solver vtable; GPU -- the C test driver runs the reference's xcond solver with HPIPM and with the
cuipm plugin swapped into the same slot (mass-spring fixture, N2 in {15, 5, 3}, batch entry, getters)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUG = os.path.join(ROOT, "acados_b200", "plugin")
VTABLE = ["opts_calculate_size", "opts_assign", "opts_initialize_default", "opts_update", "opts_set", "opts_get",
          "memory_calculate_size", "memory_assign", "memory_get", "workspace_calculate_size", "memory_reset", "solver_get",
          "eval_forw_sens", "eval_adj_sens", "terminate", "config_initialize_default", "batch_solve"]


def test_plugin_exports_vtable(built):
    so = os.path.join(PLUG, "libocp_qp_cuipm.so")
    if not os.path.exists(so):
        pytest.skip("plugin not built (needs the reference headers)")
    out = subprocess.run(["nm", "-D", "--defined-only", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    for name in VTABLE:
        assert f" T ocp_qp_cuipm_{name}" in out, name
    assert " T ocp_qp_cuipm\n" in out      # evaluate
    # host code stays plain C: no C++ runtime, no CUDA runtime in the plugin itself
    ldd = subprocess.run(["ldd", so], stdout=subprocess.PIPE, text=True).stdout
    assert "libstdc++" not in ldd.split("libcuipm")[0]


@pytest.mark.gpu
def test_plugin_driver_matches_hpipm_plugin(built):
    exe = os.path.join(PLUG, "plugin_test_driver")
    if not os.path.exists(exe):
        pytest.skip("plugin test driver did not travel")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and "PLUGIN TEST PASSED" in r.stdout, r.stdout

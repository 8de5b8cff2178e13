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


INTEG = os.path.join(ROOT, "integration")


def test_registration_patch_applies_and_library_exports_the_plugin(built):
    """integration/acados_cuipm.patch is a real unified diff against the reference (enum entry, case, name string, CMake
    option, Python whitelist); where /root/reference exists it is applied to a scratch copy and a libacados with
    ACADOS_WITH_CUIPM is built (integration/Makefile); elsewhere the prebuilt library is inspected."""
    patch = open(os.path.join(INTEG, "acados_cuipm.patch")).read()
    for needle in ("PARTIAL_CONDENSING_CUIPM,", "case PARTIAL_CONDENSING_CUIPM:", '"PARTIAL_CONDENSING_CUIPM"', "ACADOS_WITH_CUIPM",
                   "ocp_qp_cuipm_config_initialize_default(solver_config->qp_solver);", "'PARTIAL_CONDENSING_CUIPM'"):
        assert needle in patch, needle
    if os.path.isdir("/root/reference"):
        r = subprocess.run(["make", "-C", INTEG], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
        for f in ("interfaces/acados_c/ocp_qp_interface.h", "interfaces/acados_c/ocp_qp_interface.c"):
            assert "PARTIAL_CONDENSING_CUIPM" in open(os.path.join(INTEG, "_build", "src", f)).read()
    so = os.path.join(INTEG, "_build", "libacados_cuipm.so")
    if not os.path.exists(so):
        pytest.skip("patched libacados not built (needs the reference)")
    out = subprocess.run(["nm", "-D", "--defined-only", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    for name in ("ocp_qp_xcond_solver_config_create_from_name", "ocp_qp_cuipm_config_initialize_default", "ocp_qp_cuipm", "ocp_qp_hpipm"):
        assert f" T {name}\n" in out, name


@pytest.mark.gpu
def test_registered_solver_passes_the_reference_acceptance_test(built):
    """ocp_qp_xcond_solver_config_create_from_name("PARTIAL_CONDENSING_CUIPM") on the patched libacados: the reference's
    acceptance test (test/ocp_qp/test_qpsolvers.cpp:117-268: mass-spring, N2 in {15, 5, 3}) against PARTIAL_CONDENSING_HPIPM of
    the same library, plus getters / sensitivities / batch entry of the driver."""
    exe = os.path.join(INTEG, "_build", "registered_test")
    if not os.path.exists(exe):
        pytest.skip("registered_test did not travel")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and "PLUGIN TEST PASSED" in r.stdout, r.stdout


@pytest.mark.gpu
def test_plugin_batch_entry_on_reference_structs(built):
    """ocp_qp_cuipm_batch_solve on n panel-major ocp_qp_in objects (integration/plugin_bench.c): solutions as the oracle's on the
    same QPs (iteration counts equal, |du| <= 1e-10), statuses ACADOS_SUCCESS, second call reuses the page-locked staging."""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    from acados_b200 import problems
    from acados_b200.binding import default_opts
    from integration import plugin_bench as pb
    from oracle import oracle_binding as ob
    if not pb.available():
        pytest.skip("libplugin_bench.so did not travel")
    b = problems.chain_mass(96, N=20, seed=4)
    o = default_opts()
    p = pb.PluginBatch(b, o)
    st, sec = p.run(2)
    sol, it, status = p.solutions()
    p.close()
    osol, oinfo = ob.oracle_solve(b, o, nthreads=8)
    assert st == 0 and (status == 0).all()
    assert np.array_equal(it, oinfo["iter"])
    assert np.max(np.abs(b.layout.u_traj(sol) - b.layout.u_traj(osol))) <= 1e-10

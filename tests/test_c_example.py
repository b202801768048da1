"""The C examples compile against the host-layer headers on CPU and run on the GPU box; in the dev container the
reference's OWN example program (src/example.c) is compiled, unmodified, against include/freesasa.h as well."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "example_coord")


def _compile(source=os.path.join(ROOT, "examples", "example_coord.c"), exe=EXE):
    import __graft_entry__ as g

    g.build()
    csrc = os.path.join(ROOT, "freesasa_b200", "csrc")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=gnu99", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), source, "-L", csrc,
                    "-lfreesasa_b200_host", "-lfsb200", f"-Wl,-rpath,{csrc}", "-lm", "-o", exe], check=True)


def test_c_example_compiles_and_links():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_c_example_runs():
    _compile()
    out = subprocess.run([EXE], check=True, capture_output=True, text=True).stdout
    assert "Lee & Richards, 2000 slices" in out and "Shrake & Rupley, 5000 points" in out
    total = float(out.splitlines()[0].split("total")[1].split()[0])
    import math
    from tests import analytic

    exact = analytic.surface_two_spheres([0, 0, 0], [2, 0, 0], 1.0, 2.0, 1.4) + 4 * math.pi * 2.9**2
    assert analytic.rel_err(exact, total) < 1e-4


EXE_S = os.path.join(ROOT, "examples", "example_structure")
REF_EXAMPLE = "/root/reference/src/example.c"


def test_structure_example_compiles_and_links():
    _compile(os.path.join(ROOT, "examples", "example_structure.c"), EXE_S)
    assert os.path.exists(EXE_S)


def test_models_example_compiles_and_links(tmp_path):
    """examples/example_models.c: freesasa_structure_array() + the additive freesasa_calc_tree_batch()."""
    exe = os.path.join(tmp_path, "example_models")
    _compile(os.path.join(ROOT, "examples", "example_models.c"), exe)
    assert os.path.exists(exe)


@pytest.mark.skipif(not os.path.exists(REF_EXAMPLE), reason="reference tree not present")
def test_reference_example_program_compiles_unmodified(tmp_path):
    """The reference's own src/example.c (freesasa_structure_from_pdb -> freesasa_calc_structure ->
    freesasa_result_classes) builds against include/freesasa.h and links the B200-backed libraries as is."""
    exe = os.path.join(tmp_path, "ref_example")
    _compile(REF_EXAMPLE, exe)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_structure_example_runs(tmp_path):
    """PDB text on stdin -> totals; they equal the serial class sums of the per-atom values the Python binding gets
    for the same text, and the written PDB carries the same per-atom numbers."""
    import ctypes

    import numpy as np

    import freesasa_b200 as fs
    from freesasa_b200 import structure as st
    from freesasa_b200 import workloads as w

    _compile(os.path.join(ROOT, "examples", "example_structure.c"), EXE_S)
    text = w.pdb_text(2000, seed=9, chains=2).encode()
    out_pdb = os.path.join(tmp_path, "out.pdb")
    out = subprocess.run([EXE_S, out_pdb], input=text, check=True, capture_output=True).stdout.decode().splitlines()
    api = st.api()
    s = api.from_pdb(text)
    sasa, total = s.calc(fs.default_parameters())
    assert out[0] == f"atoms  : {s.n}"
    assert out[1] == "Total  : %f A2" % total
    assert [ln.split()[1] for ln in out if ln.startswith("CHAIN")] == ["A", "B"]
    written = [ln for ln in open(out_pdb).read().splitlines() if ln.startswith("ATOM")]
    assert len(written) == s.n
    assert [ln[60:66] for ln in written] == ["%6.2f" % v for v in sasa]


def test_multi_gpu_example_compiles_and_links(tmp_path):
    """examples/example_multi.c: fsb200_lr_multi() + fsb200_get_multi_stats() straight from C, engine library only."""
    import __graft_entry__ as g

    g.build()
    csrc = os.path.join(ROOT, "freesasa_b200", "csrc")
    exe = os.path.join(tmp_path, "example_multi")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=gnu99", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "example_multi.c"), "-L", csrc, "-lfsb200", f"-Wl,-rpath,{csrc}", "-lm", "-o", exe],
                   check=True)
    assert os.path.exists(exe)


"""The C example (examples/example_coord.c) compiles against the host-layer header on CPU and runs on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "example_coord")


def _compile():
    import __graft_entry__ as g

    g.build()
    csrc = os.path.join(ROOT, "freesasa_b200", "csrc")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=gnu99", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "example_coord.c"), "-L", csrc, "-lfreesasa_b200_host", "-lfsb200",
                    f"-Wl,-rpath,{csrc}", "-lm", "-o", EXE], check=True)


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

"""One device-resident C2 (or C3) call pattern for an ncu capture of the integration kernel:
    ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 3 -c 1 -o gpurun_out/x python tests/tools/ncu_one.py [lr|sr]
(FSB200_ENGINE_LIB selects a variant build).  Five calls; skip the first three."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import freesasa_b200 as fs  # noqa: E402

alg, res = (1, 1000) if "sr" in sys.argv[1:] else (0, 100)
x, r = fs.workloads.globule(100000)
dev = torch.device("cuda", 0)
dx, dr = torch.tensor(x, device=dev), torch.tensor(r, device=dev)
eng = fs.Engine(0)
for _ in range(5):
    eng.calc_device(alg, dx, dr, 1.4, res)
print(eng.stats())

import sys; sys.path.insert(0,".")
import numpy as np
import freesasa_b200 as fs
structs = fs.workloads.batch(int(sys.argv[1]) if len(sys.argv)>1 else 512, 4000, 6000, seed=0)
e = fs.Engine(0)
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 8):
    outs = e.calc_batch(0, structs, 1.4, 50)
    print(it, e.stats()["n_certified"], e.stats()["n_overflow"], float(sum(o.sum() for o in outs)), flush=True)

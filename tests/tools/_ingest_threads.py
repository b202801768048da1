import os, sys, time, subprocess
sys.path.insert(0, os.getcwd())
code = r'''
import os, sys, time
sys.path.insert(0, os.getcwd())
from freesasa_b200 import structure as st, workloads as w
api = st.api(); api.lib.freesasa_set_verbosity(1)
for n in (100000, 1000000):
    text = w.pdb_text(n, seed=5, chains=8).encode()
    path = "/dev/shm/_fsb_%d.pdb" % n
    open(path, "wb").write(text)
    best = 1e9
    for _ in range(5):
        t = time.perf_counter(); s = api.from_pdb_path(path); dt = time.perf_counter() - t; s.free(); best = min(best, dt)
    os.remove(path)
    print("threads", os.environ.get("FREESASA_B200_THREADS"), n, "atoms", round(best * 1e3, 2), "ms", flush=True)
'''
print("cpus", os.cpu_count())
for t in (1, 2, 4, 8, 16):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FREESASA_B200_THREADS=str(t)))

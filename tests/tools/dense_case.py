import sys, time; sys.path.insert(0,".")
import numpy as np
import freesasa_b200 as fs
e=fs.Engine(0)
x,r=fs.workloads.globule(100000); x=x*0.78
for i in range(3):
    t=time.perf_counter(); e.calc(0,x,r,1.4,100); print("dense 100k LR100 wall ms", (time.perf_counter()-t)*1e3, e.stats())

import sys; sys.path.insert(0,".")
import numpy as np, time
import freesasa_b200 as fs
x,r=fs.workloads.globule(100000)
e=fs.Engine(0)
for i in range(6):
    t=time.perf_counter(); e.calc(0,x,r,1.4,100); dt=time.perf_counter()-t
    s=e.stats(); print("wall %.3f ms | lib total %.3f stage %.3f device %.3f integrate %.3f" % (dt*1e3, s["host_total_ms"], s["host_stage_ms"], s["device_ms"], s["integrate_ms"]))
p=fs.Parameters(0,1.4,100,100,1)
for i in range(4):
    t=time.perf_counter(); fs.calc_coord(x,r,p); print("calc_coord wall %.3f ms" % ((time.perf_counter()-t)*1e3))

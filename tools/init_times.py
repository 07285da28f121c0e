import time, sys, os
sys.path.insert(0, "/root/repo")
t0=time.perf_counter()
import numpy as np
import telomeric_identifier_b200 as T
t1=time.perf_counter()
ctx=T.Context(0)
t2=time.perf_counter()
seq=np.frombuffer(b"ACGT"*2500000, dtype=np.uint8).copy(); off=np.array([0,seq.size],dtype=np.uint64)
b=ctx.upload(seq,off)
t3=time.perf_counter()
f,r=ctx.window_counts_resident(b,10000,[b"AACCT"])
t4=time.perf_counter()
f,r=ctx.window_counts_resident(b,10000,[b"AACCT"])
t5=time.perf_counter()
runs=ctx.explore_runs_resident(b,0.01,5,12,100)
t6=time.perf_counter()
print(f"import {t1-t0:.3f}  create {t2-t1:.3f}  upload10MB {t3-t2:.3f}  first count {t4-t3:.3f}  second {t5-t4:.4f}  first explore {t6-t5:.3f}")

import sys; sys.path.insert(0,'/root/repo')
import __graft_entry__ as G, numpy as np, ctypes as C
pkg=G.build(); ctx=pkg.Context(0)
for glen,k in ((2000000,127),(2000000,65)):
    nreads=int(glen*30/150)//2*2
    seq,offs=pkg.synth_reads(2,glen,nreads,150,True,700,10000)
    c=pkg.serial_mem(k,pkg.CANONICAL,ctx=ctx)((seq,offs))
    print(glen,k,len(c),c.n_instances,flush=True)
    c.free()

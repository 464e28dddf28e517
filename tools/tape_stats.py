"""Per-level statistics of the K5 tapes of one config-5 agent (host only)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcora_b200 import _lib, graph as pg, synthetic

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
A = int(sys.argv[2]) if len(sys.argv) > 2 else 256
agent = int(sys.argv[3]) if len(sys.argv) > 3 else A // 2
t0 = time.time()
meas, gt, idx = synthetic.grid3d(S, 0.1, seed=0)
gp = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, A))
diag, cross = gp.split(agents=[agent])
rowptr, col, val = diag[agent]
k = len(rowptr) - 1
print(f"agent {agent}: n={gp.n_a[agent]} l={gp.l_a[agent]} k={k} nnz={len(val)} ({time.time()-t0:.1f}s)")
np.savez("/tmp/agent_q.npz", rowptr=rowptr, col=col, val=val, n=gp.n_a[agent], l=gp.l_a[agent])

import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, gemma_b200
from gemma_b200 import synth
from test_mvlmm_core import _problem
n=1100; pb=_problem(n,2,9); bed,G=synth.make_bed(n,96,seed=321,miss_rate=0.01)
ctx=gemma_b200.Context(0)
ctx.mvlmm_setup(pb["U"],pb["ev"],pb["U"]@pb["UtW"],pb["U"]@pb["UtY"]); ctx.mvlmm_null()
a=ctx.mvlmm_batch_bed(bed,n,a_mode=4); b=ctx.mvlmm_batch_geno(np.where(G<0,np.nan,G),4)
print("max rel diff bed vs geno", float(np.max(np.abs(a-b)/np.maximum(np.abs(b),1e-300))))

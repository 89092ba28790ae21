import sys, json, time, faulthandler
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, astar_pairwise_aligner_b200 as A
G=json.load(open('/root/repo/tests/golden/reference_vectors.json'))
pairs=[(p['a'].encode(),p['b'].encode()) for p in G['pairs']]
which=sys.argv[1]
if which=='kat':
    e2=A.Engine(0); a,_=A.generate_pair(256,0,0,1); b,_=A.generate_pair(64,0,0,2); print(e2.block_compute(a,b)[0], flush=True)
t=time.time()
c,cg=A.AstarPa2(0,True).align_batch(pairs)
print("stream first", c[:4], time.time()-t, flush=True)

import sys,time,hashlib; sys.path.insert(0,'/root/repo')
import spand_public_b200 as S, numpy as np
n,d,L=128,3,16
A=S.neglapl(n,d); X=S.linspace_nd(n,d); G=S.symmetric_graph(A)
hs=set()
for rep in range(int(sys.argv[1])):
    t=S.Tree(L); t.set_use_geo(True); t.set_Xcoo(X); t.partition(G)
    p=t.get_assembly_perm(); ids=np.stack(t.partition_ids())
    h=hashlib.md5(p.tobytes()).hexdigest()[:8]+hashlib.md5(np.ascontiguousarray(ids).tobytes()).hexdigest()[:8]
    hs.add(h)
print("distinct results:", len(hs), hs)

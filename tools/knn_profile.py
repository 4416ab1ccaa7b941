import sys, numpy as np, torch
sys.path.insert(0,'.')
from morb_slam_b200 import capi, synth
ex = capi.ORBextractor(1000)
nq, ndb = 1200, 1250000
q = torch.from_numpy(synth.random_descriptors(10, nq)).cuda()
db = torch.randint(0,256,(ndb,32),dtype=torch.uint8,device='cuda')
oi = torch.empty((nq,2),dtype=torch.int32,device='cuda'); od = torch.empty_like(oi)
fl = capi.ORB_SRC_DEVICE|capi.ORB_DST_DEVICE
for _ in range(3):
    capi.hamming_knn2(ex, q.data_ptr(), db.data_ptr(), 0, fl, ndb=ndb, nq=nq, out=(oi.data_ptr(), od.data_ptr()))
print('done')

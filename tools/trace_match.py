import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from object_detection_torch2_b200 import ops, synth, _lib
N=32; dev=torch.device("cuda"); lib=_lib.load(); lib.ssdh_debug_set_loss_trace.argtypes=[ctypes.c_void_p]
priors=ops.default_boxes(dev); o,t=synth.make_batch(N,0,"D1"); o,t=o.to(dev),t.to(dev)
for _ in range(3): ops.multibox_loss_raw(o,t,priors)
CL=ops.device_info()["loss_cluster_size"]
trace=torch.zeros(N*CL,64,dtype=torch.int64,device=dev)
torch.cuda.synchronize(); lib.ssdh_debug_set_loss_trace(trace.data_ptr())
ops.multibox_loss_raw(o,t,priors); torch.cuda.synchronize(); lib.ssdh_debug_set_loss_trace(None)
tr=trace.cpu().numpy()
for cta in (0,1,5,40):
    base=tr[cta,2]
    print('cta',cta,'t2',0,'t3',tr[cta,3]-base)
    print('  warp0 g stamps', (tr[cta,16:36]-base).tolist())
    print('  warp11 g stamps', (tr[cta,40:60]-base).tolist())
gt=(t[:, :, 2]>0).sum(1).cpu().tolist(); print('gt counts', gt[:12])

import sys, ctypes, torch, math
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import ops_util as ou
from image2video_synthesis_using_cinns_b200 import lib
L=lib.load()
def run(name, B,C,T,H,W,Cout):
    x=torch.randn(B,T,H,W,C,device='cuda'); w=torch.randn(27,Cout,C,device='cuda')*0.02; b=torch.zeros(Cout,device='cuda')
    ncta=2048
    buf=torch.zeros(ncta*8,dtype=torch.int64,device='cuda')
    for rep in range(2):
        lib.check(L.i2v_debug_conv_tc_timestamps(ctypes.c_void_p(buf.data_ptr()), ncta))
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        # note: op includes split kernels; time only informative
        y=ou.conv_tc(x,w,b,None,(3,3,3),variant=2)
        torch.cuda.synchronize()
    lib.check(L.i2v_debug_conv_tc_timestamps(None, 0))
    t=buf.cpu().view(ncta,8).double()
    t=t[t[:,6]>0]
    t0=t[:,0:1]
    d=(t-t0)/1000.0
    import numpy as np
    print(name, 'ctas',len(t))
    names=['start','prologue','first_stage','last_mma_issued','acc_complete','epi_stores','end']
    for i,n in enumerate(names): print(f"   {n:16s} median {d[:,i].median():9.2f} us   p90 {d[:,i].quantile(0.9):9.2f}")
    # CTA start times relative to kernel start: wave structure
    st=(t[:,0]-t[:,0].min())/1000.0
    print('   start time quantiles us', [round(float(st.quantile(q)),1) for q in (0.1,0.3,0.5,0.7,0.9,1.0)], 'kernel span', float((t[:,6].max()-t[:,0].min())/1000))
run('g3_conv1_like', 4,128,16,64,64,128)
run('g4_conv1_like', 4,64,16,64,64,64)
run('g3_conv0_like', 4,256,16,64,64,128)
run('g4_conv0_like', 4,128,16,64,64,64)

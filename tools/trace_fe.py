import sys, time, os, numpy as np
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import common
from common import drr, synth_wad
path, gm = common.wad('e1m1')
W,H,n=320,200,4096
views=np.array(synth_wad.walk_viewpoints(gm,n),np.float32)
ctx=drr.Context(W,H,0,n); scene=drr.Scene(path,'E1M1',W,H); scene.upload_assets(ctx)
for it in range(6):
    if it==5: ctx.set_knob('fe_trace', 1)  # (the DRR_* variables are read once, at drr_ctx_create)
    t0=time.perf_counter(); ctx.reset(); t1=time.perf_counter(); scene.emit_views_device(ctx, views, 0.0, 3); t2=time.perf_counter(); ctx.draw(); t3=time.perf_counter(); c=ctx.read_checksums(0,n); t4=time.perf_counter()
    print('reset %.0f emit %.0f draw %.0f read %.0f total %.0f us' % ((t1-t0)*1e6,(t2-t1)*1e6,(t3-t2)*1e6,(t4-t3)*1e6,(t4-t0)*1e6))

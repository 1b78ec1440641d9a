import sys; sys.path.insert(0,'dendro-kt_b200')
import dkt, torch, time
for lvl,w in ((9,1/32),(9,1/16),(9,0.25),(8,0.5)):
    t=time.time()
    x,l = dkt.trees.moving_ball_tree(4, lvl, 12, use_torch=True, t0=0.5-w/2, t1=0.5+w/2)
    torch.cuda.synchronize()
    print(lvl, w, x.shape[0], time.time()-t, torch.cuda.max_memory_allocated()/1e9)
    del x,l

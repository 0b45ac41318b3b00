import sys; sys.path.insert(0,'/root/repo')
import numpy as np, cv2
import oracle as O, trackingbench_slam_b200 as tb
left = cv2.imread('tests/golden/data/left.png',0)
ctx = tb.Context(0)
nl, sc, tgt, a, b = 8, 1/1.2, 2000, 20, 7
sf, lw, lh = tb.pyramid_geometry(left.shape[1], left.shape[0], nl, sc)
lv,_ = O.pyramid(left, nl, sc)
fe = tb.Frontend(ctx, lw, lh, 1, tgt)
fe.upload_frames(0, left[None]); fe.build_pyramid(1); fe.extract(1, sf, tgt, a, b); ctx.sync()
for l in range(nl):
    co = O.cell_candidates(lv[l], a, b); cg = fe.get_candidates(0, l)
    so = set((int(x),int(y),int(r)) for x,y,r in zip(co['x'],co['y'],co['response']))
    sg = set((int(x),int(y),int(r)) for x,y,r in zip(cg['x'],cg['y'],cg['response']))
    miss = sorted(so-sg); extra = sorted(sg-so)
    W,H = lv[l].shape[1], lv[l].shape[0]
    width=W-32; height=H-32; nC=int(width/30); nR=int(height/30); wC=-(-width//nC); hC=-(-height//nR)
    print('L',l,W,H,'wCell',wC,'hCell',hC,'missing',len(miss),'extra',len(extra))
    for (x,y,r) in miss[:40]: print('   miss',x,y,r,'cell',(x-3)//wC,(y-3)//hC,'in-cell',(x-3)%wC,(y-3)%hC)
    for (x,y,r) in extra[:10]: print('   extra',x,y,r,'cell',(x-3)//wC,(y-3)//hC)

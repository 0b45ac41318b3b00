"""Developer parity sweep on a GPU box: CUDA path (through the C ABI) vs the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import oracle as O
import trackingbench_slam_b200 as tb

left = cv2.imread('trackingbench_slam_b200/data/left.png', 0)
right = cv2.imread('trackingbench_slam_b200/data/right.png', 0)
ctx = tb.Context(0)
fails = 0
def check(name, ok, extra=''):
    global fails
    print(('OK   ' if ok else 'FAIL ') + name, extra)
    if not ok: fails += 1

def kp_equal(a, b):
    return len(a) == len(b) and a.tobytes() == b.tobytes()

for name, img in [('left', left), ('right', right), ('rand', np.random.default_rng(1).integers(0, 256, (300, 400), dtype=np.uint8))]:
    for (nl, sc, tgt, a, b) in [(8, 1/1.2, 2000, 20, 7), (5, 0.8, 1000, 80, 30)]:
        sf, lw, lh = tb.pyramid_geometry(img.shape[1], img.shape[0], nl, sc)
        lv, sf_o = O.pyramid(img, nl, sc)
        check('%s geometry %d' % (name, nl), np.array_equal(sf, sf_o) and all((l.shape[1], l.shape[0]) == (w, h) for l, w, h in zip(lv, lw, lh)))
        fe = tb.Frontend(ctx, lw, lh, max_batch=2, max_target=tgt)
        fe.upload_frames(0, np.stack([img, img]))
        fe.build_pyramid(2)
        for l in range(nl):
            for f in range(2):
                got = fe.download_level(f, l)
                check('%s pyr L%d f%d' % (name, l, f), np.array_equal(got, lv[l]), '' if np.array_equal(got, lv[l]) else 'maxdiff %d' % np.abs(got.astype(int) - lv[l]).max())
        fe.extract(2, sf, tgt, a, b)
        ctx.sync()
        ko, do, npl = O.orb_extract(lv, sf, tgt, a, b)
        for l in range(nl):
            co = O.cell_candidates(lv[l], a, b)
            cg = fe.get_candidates(1, l)
            ok = len(co) == len(cg) and np.array_equal(co['x'], cg['x']) and np.array_equal(co['y'], cg['y']) and np.array_equal(co['response'], cg['response'])
            check('%s cand L%d (%d vs %d)' % (name, l, len(co), len(cg)), ok)
            bo = O.gauss7x7(lv[l]); bg = fe.download_level(1, l, blurred=True)
            check('%s blur L%d' % (name, l), np.array_equal(bo, bg))
        print('  level counts', fe.get_level_counts(0), list(npl))
        for f in range(2):
            kg, dg = fe.get_results(f)
            okk = kp_equal(kg, ko)
            check('%s extract f%d kps (%d vs %d)' % (name, f, len(kg), len(ko)), okk)
            if not okk and len(kg) == len(ko):
                for fld in kg.dtype.names: print('    ', fld, int((kg[fld] != ko[fld]).sum()))
            okd = dg.shape == do.shape and np.array_equal(dg, do)
            check('%s extract f%d desc' % (name, f), okd, '' if okd or dg.shape != do.shape else 'rows differing %d' % int((dg != do).any(1).sum()))
        fe.close()
print('FAILS', fails)
sys.exit(1 if fails else 0)

#!/usr/bin/env python3
"""bench.py — headline benchmark: frames/s of pyramid + FAST/ORB extract + frame-to-frame Hamming
match on a synthetic 1080p stream (BASELINE.json configs[2]: 2000 kpts/frame, batch 256, 8 levels,
scale 1.2, FAST 20/7, searchByBF(k, k+1)).

  python bench.py --gpus N --steps K --warmup W            # ours (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on host cores

One "step" = one pass of the hot path over one batch of B frames per GPU (weak scaling: every rank
owns its own B-frame stream, no data-path collective).  `value` times device-resident inputs with
CUDA events on the launching stream; `e2e` times the same step through the C ABI with HOST buffers
(pinned H2D of the frames + D2H of keypoints/descriptors/matches inside the timed region).
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'frames/s FAST+ORB extract+match at 1080p'
PSET = dict(n_levels=8, scale=1.0 / 1.2, target=2000, initTh=20.0, minTh=7.0, ratio=10.0, match_minTh=30.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--width', type=int, default=1920)
    ap.add_argument('--height', type=int, default=1080)
    ap.add_argument('--cpu-sample', type=int, default=0, help='frames in the CPU-baseline sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--match-impl', default='tc', choices=['popc', 'tc'], help='brute-force matcher kernel: LOP3+POPC or tcgen05 int8')
    ap.add_argument('--fast-th', default='', help='developer knob: initTh,minTh override (not the benchmark config)')
    ap.add_argument('--profile-out', default='', help='write the per-kernel time table (JSON) here')
    ap.add_argument('--no-extras', action='store_true', help='skip the sustained run, config 4 and the other-config blocks')
    ap.add_argument('--knn-n', type=int, default=200000, help='rows of the config-4 kNN problem')
    return ap.parse_args()


def env_rank():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def workload_name(a):
    return 'synthetic %dx%d frame stream, %d kpts/frame, extract + frame-to-frame match, batch %d' % (
        a.width, a.height, PSET['target'], a.batch)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self._stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.rows)}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(frames, n_threads, keep_results=False):
    """The reference's CPU path on `frames` (consecutive synthetic frames), frames farmed over host threads:
      * pyramid = chained cv2.resize, the call Frame::ComputePyramid makes (src/types/Frame.cpp:423), cv2 4.13 single-threaded;
      * ORBExtractor::operator() = oracle/_ref, the reference's own ORBextractor.cpp compiled unmodified — its cv::FAST /
        cv::GaussianBlur calls land in the plain-C shim restatements (oracle/shim), NOT in OpenCV's SIMD kernels, because the
        cv2 Python module exports no C++ symbols to link against (else the C port of the whole extractor);
      * searchByBF(k, k+1) = cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match (matcher.cpp:20,207) + the reference's filter.
    Returns a dict: fps, kind, primitives and (keep_results) the per-frame (kps, desc), cross-check and filtered lists."""
    import oracle as O
    from concurrent.futures import ThreadPoolExecutor
    kind = 'reference' if O.ref_available() else 'port'
    nl, sc, tgt, a, b = PSET['n_levels'], PSET['scale'], PSET['target'], PSET['initTh'], PSET['minTh']
    H, W = frames[0].shape
    sf = O.scale_factors(nl, sc)
    sizes = O.level_sizes(W, H, sf)
    try:
        import cv2
        cv2.setNumThreads(1)
        prim = 'cv2 %s resize + BFMatcher (1 thread per frame); naive C shim FAST/GaussianBlur inside the extractor' % cv2.__version__
    except Exception:
        cv2 = None
        prim = 'naive C restatements of resize/FAST/GaussianBlur/BFMatcher (cv2 not importable)'
    local = threading.local()

    def extract(img):
        if cv2 is not None:
            lv = [img]
            for i in range(1, nl):
                lv.append(cv2.resize(lv[-1], sizes[i]))
        else:
            lv, _ = O.pyramid(img, nl, sc)
        if kind == 'reference':
            if not hasattr(local, 'ex'):
                local.ex = O.RefExtractor()
            return local.ex(lv, sf, tgt, a, b)
        k, d, _ = O.orb_extract(lv, sf, tgt, a, b)
        return k, d

    def match(pair):
        d1, d2 = pair
        if cv2 is not None and len(d1) and len(d2):
            ms = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(d1, d2)
            cross = np.array([(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in ms], dtype=O.MATCH_DTYPE)
        else:
            cross = O.bf_cross(d1, d2)
        if len(cross) == 0:
            return cross, cross
        thr = min(np.float32(PSET['ratio']) * np.float32(cross['distance'].min()), np.float32(PSET['match_minTh']))   # matcher.cpp:209-218
        return cross, cross[cross['distance'] < thr]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:
        res = list(ex.map(extract, list(frames)))
        mt = list(ex.map(match, [(res[i][1], res[i + 1][1]) for i in range(len(res) - 1)]))
    dt = time.perf_counter() - t0
    out = {'fps': len(frames) / dt, 'kind': kind, 'primitives': prim}
    if keep_results:
        out.update(results=res, cross=[m[0] for m in mt], filtered=[m[1] for m in mt])
    return out


def run_reference(a):
    rank, local_rank, world = env_rank()
    if rank != 0:
        return
    from trackingbench_slam_b200.synth import synth_stream
    cores = os.cpu_count() or 1
    n = a.cpu_sample or max(4, min(2 * cores, 32))
    frames = synth_stream(n, a.width, a.height)
    vals = []
    for i in range(a.warmup + a.steps):
        r = cpu_reference_run(frames, cores)
        v, kind = r['fps'], r['kind']
        if i >= a.warmup:
            vals.append(v)
        if i == 0 and n / v > 40:  # keep the whole run within a few minutes on slow hosts
            a.warmup, a.steps = min(a.warmup, 1), min(a.steps, 2)
            if len(vals) >= a.steps:
                break
    value = float(np.mean(vals)) if vals else v
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': a.gpus, 'steps': len(vals) or 1,
            'warmup': a.warmup, 'ms_per_step': 1000.0 * n / value, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': workload_name(a), 'levels': 8, 'scale': 1.2, 'fast_th': [20, 7], 'frames_per_gpu': a.batch,
                       'sharding': 'CPU: frames farmed over %d host threads' % cores},
            'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': kind, 'primitives': r['primitives'],
                             'sample': '%d consecutive frames per step (extract all + searchByBF of the %d pairs)' % (n, n - 1)},
            'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ ours: supplementary blocks
def config4_block(torch, dist, ctx, world, rank, n, steps=3):
    """BASELINE config 4: brute-force Hamming kNN-2 on n x n 256-bit descriptors, query rows sharded over the ranks, one
    all_gather of (argbest, best, second) — plus the cross-check variant (transposed problem + second gather).
    Every rank must call this. Times on the device (CUDA events), max over ranks; kernel and gather are split."""
    from trackingbench_slam_b200.sharding import sharded_knn2, sharded_knn2_cross
    from trackingbench_slam_b200.synth import synth_descriptors
    Q, T = synth_descriptors(n, seed=42)
    q, t = torch.from_numpy(Q).cuda(), torch.from_numpy(T).cuda()

    def reduce_max(x):
        if world > 1:
            v = torch.tensor([x], dtype=torch.float64, device='cuda')
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            return float(v.item())
        return x

    for _ in range(2):
        sharded_knn2(ctx, q, t)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tot, ker, gat = [], [], []
    for _ in range(steps):
        ev = []
        idx, best, second = sharded_knn2(ctx, q, t, None, ev)
        torch.cuda.synchronize()
        tot.append(ev[0].elapsed_time(ev[2])); ker.append(ev[0].elapsed_time(ev[1])); gat.append(ev[1].elapsed_time(ev[2]))
    ms, kms, gms = reduce_max(float(np.mean(tot))), reduce_max(float(np.mean(ker))), reduce_max(float(np.mean(gat)))
    ev = []
    sharded_knn2_cross(ctx, q, t, None, [])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    i12, b12, s12, mutual = sharded_knn2_cross(ctx, q, t, None, ev)
    e_end = torch.cuda.Event(enable_timing=True)
    e_end.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    cross_ms = reduce_max(ev[0].elapsed_time(e_end))
    out = None
    if rank == 0:
        b, s_, i = best.cpu().numpy(), second.cpu().numpy(), idx.cpu().numpy()
        spot = np.unpackbits(Q[:2000] ^ T[i[:2000]], axis=1).sum(1)
        out = {'workload': 'brute-force Hamming kNN-2 %dx%d 256-bit descriptors, ratio test, query rows sharded' % (n, n),
               'n_ranks': world, 'ms': ms, 'Tpairs_per_s': float(n) * n / (ms * 1e-3) / 1e12, 'kernel_ms': kms, 'gather_us': gms * 1e3,
               'gather': 'none (1 rank)' if world == 1 else 'all_gather of (argbest, best, second): %d B per rank' % (12 * -(-n // world)),
               'ratio_test_pass': int((b.astype(np.float32) < np.float32(0.7) * s_.astype(np.float32)).sum()),
               'planted_found': int((b[:n // 2] < 60).sum()), 'spot_check_exact': bool(np.array_equal(spot, b[:2000])),
               'cross_check': {'ms': cross_ms, 'mutual_matches': int(mutual.sum().item()),
                               'what': 'both directions sharded by rows + two gathers + idx21[idx12[q]] == q'}}
        try:   # kernel time against the measured int8 tensor peak of the GPUs used (512 int8 ops per pair at K = 256)
            pk = json.load(open(os.path.join(ROOT, 'profiles', 'r02_int8_peak.json')))['int8_dense_Pop_per_s']
            out['kernel_frac_of_int8_peak'] = float(n) * n * 512 / (kms * 1e-3) / 1e15 / (pk * world)
        except Exception:
            pass
    del q, t
    return out


def extras_block(torch, dist, tb, local_rank, world, rank, a):
    """The other BASELINE configs as short device-resident measurements (CUDA events, after warm-up): config 1 and 2 on the
    KITTI pair, the 1241x376 synthetic stream, config 5 (4K streams sharded over the ranks). Rank 0 reports."""
    from trackingbench_slam_b200.synth import load_gray, synth_frame, synth_stream
    st = torch.cuda.current_stream()
    ctx = tb.Context(local_rank, st.cuda_stream)
    ctx.set_match_impl(1 if a.match_impl == 'tc' else 0)
    out = {}

    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(iters):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    if rank == 0:
        left, right, disp = load_gray('left.png'), load_gray('right.png'), load_gray('disparity.png')
        H, W = left.shape
        # ---- config 1: ORB extract 1000 kpts on left/right (the reference's own call-site parameters 5 x 0.8, 80/30) + searchByBF
        nl, sc, tgt = 5, 0.8, 1000
        sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
        fe = tb.Frontend(ctx, lw, lh, max_batch=2, max_target=tgt)
        cap = fe.kp_cap
        fe.upload_frames(0, np.stack([left, right]))
        m = torch.zeros((1, cap, 4), dtype=torch.int32, device='cuda')
        c = torch.zeros(1, dtype=torch.int32, device='cuda')

        def c1():
            fe.build_pyramid(2)
            fe.extract(2, sf, tgt, 80.0, 30.0)
            fe.search_by_bf([0], [1], 10.0, 30.0, m.data_ptr(), cap, c.data_ptr())
        ms = timed(c1, 50)
        _, _, cnt = fe.get_results_batch(2)
        out['config1'] = {'workload': 'left.png/right.png 1241x376: pyramid + ORB extract 1000 kpts (5 levels, 0.8, FAST 80/30) + searchByBF',
                          'ms_per_pair': ms, 'pairs_per_s': 1e3 / ms, 'keypoints': [int(x) for x in cnt], 'matches': int(c.item())}
        fe.close()
        # ---- config 2: stereo pairs, extract 2000 + disparity-guided match = searchByProjection(R, L) with map points from the
        #      left keypoints at depth b*fx/disparity (test_vo.cpp:414,496-498); NP stereo pairs per step
        NP, tgt = 32, 2000
        fe = tb.Frontend(ctx, lw, lh, max_batch=2 * NP, max_target=tgt)
        cap = fe.kp_cap
        fe.upload_frames(0, np.stack([left, right] * NP))
        fe.build_pyramid(2 * NP)
        fe.extract(2 * NP, sf, tgt, 80.0, 30.0)
        kps, _, cnt = fe.get_results_batch(2 * NP)
        fx, fy, cx, cy, base = 718.856, 718.856, 607.1928, 185.2157, 0.573
        valid = np.zeros((2 * NP, cap), np.uint8)
        xyz = np.zeros((2 * NP, cap, 3), np.float32)
        for f in range(0, 2 * NP, 2):
            k = kps[f, :cnt[f]]
            u = np.clip(k['x'].astype(np.int32), 0, W - 1)
            v = np.clip(k['y'].astype(np.int32), 0, H - 1)
            d = disp[v, u].astype(np.float32)
            depth = np.where(d > 0, np.float32(base * fx) / np.maximum(d, 1), 0).astype(np.float32)
            valid[f, :cnt[f]] = d > 0
            xyz[f, :cnt[f]] = np.stack([(k['x'] - np.float32(cx)) / np.float32(fx) * depth, (k['y'] - np.float32(cy)) / np.float32(fy) * depth, depth], 1)
        dv, dx = torch.from_numpy(valid).cuda(), torch.from_numpy(xyz).cuda()
        T = np.tile(np.eye(4, dtype=np.float32), (NP, 1, 1))
        T[:, 0, 3] = -base
        cam = tb.Pinhole(fx, fy, cx, cy, W, H, (tb.capi.C.c_float * 5)(0, 0, 0, 0, 0))
        pp = tb.MatcherParams(30, 50, 30, 1, 30.0)
        f1, f2 = np.arange(1, 2 * NP, 2, dtype=np.int32), np.arange(0, 2 * NP, 2, dtype=np.int32)
        m = torch.zeros((NP, cap, 4), dtype=torch.int32, device='cuda')
        c = torch.zeros(NP, dtype=torch.int32, device='cuda')

        def c2():
            fe.build_pyramid(2 * NP)
            fe.extract(2 * NP, sf, tgt, 80.0, 30.0)
            fe.search_by_projection(pp, cam, f1, f2, T, None, dv.data_ptr(), dx.data_ptr(), sf, m.data_ptr(), cap, c.data_ptr())
        ms = timed(c2, 20)
        out['config2'] = {'workload': 'stereo left/right 1241x376: pyramid + ORB extract 2000 kpts + disparity-guided searchByProjection(R, L), %d pairs per step' % NP,
                          'ms_per_step': ms, 'stereo_pairs_per_s': NP * 1e3 / ms, 'frames_per_s': 2 * NP * 1e3 / ms,
                          'projection_matches_per_pair': int(c[0].item())}
        fe.close()
        # ---- synthetic 1241x376 stream (8 levels x 1/1.2, 2000 kpts, FAST 20/7), batch 256 + searchByBF(k, k+1)
        Bk, Wk, Hk = 256, 1241, 376
        sfk, lwk, lhk = tb.pyramid_geometry(Wk, Hk, PSET['n_levels'], PSET['scale'])
        fe = tb.Frontend(ctx, lwk, lhk, max_batch=Bk, max_target=PSET['target'])
        cap = fe.kp_cap
        fe.upload_frames(0, synth_stream(Bk, Wk, Hk))
        qf, tf = np.arange(0, Bk - 1, dtype=np.int32), np.arange(1, Bk, dtype=np.int32)
        m = torch.zeros((Bk - 1, cap, 4), dtype=torch.int32, device='cuda')
        c = torch.zeros(Bk - 1, dtype=torch.int32, device='cuda')

        def ck():
            fe.build_pyramid(Bk)
            fe.extract(Bk, sfk, PSET['target'], PSET['initTh'], PSET['minTh'])
            fe.search_by_bf(qf, tf, PSET['ratio'], PSET['match_minTh'], m.data_ptr(), cap, c.data_ptr())
        ms = timed(ck, 10)
        out['kitti_size_stream'] = {'workload': 'synthetic 1241x376 frame stream, 2000 kpts/frame, extract + frame-to-frame match, batch 256',
                                    'ms_per_step': ms, 'frames_per_s': Bk * 1e3 / ms}
        # ---- the same stream through Matcher::searchByOPFlow(F_{k+1}, F_k, pts, equalized = true, reject = true), the call of
        #      test_vo.cpp:716 / LocalBA.cpp:54: CLAHE + pyramidal LK + IsInFrame filter + rejectWithF (RANSAC), 255 pairs per step
        pts = torch.zeros((Bk - 1, cap, 2), dtype=torch.float32, device='cuda')
        stt = torch.zeros((Bk - 1, cap), dtype=torch.uint8, device='cuda')
        itr = torch.zeros(Bk - 1, dtype=torch.int32, device='cuda')
        lkp = tb.LKParams()

        def cf():
            fe.search_by_opflow(lkp, Bk, tf, qf, 1, pts.data_ptr(), stt.data_ptr())
            fe.reject_with_f(qf, pts.data_ptr(), stt.data_ptr(), None, itr.data_ptr())
        ms = timed(cf, 10)
        out['opflow_reject'] = {'workload': 'searchByOPFlow(equalized, reject = true) over the %d consecutive pairs of that stream: CLAHE + LK 21x21 / 3 levels + IsInFrame + rejectWithF' % (Bk - 1),
                                'ms_per_step': ms, 'pairs_per_s': (Bk - 1) * 1e3 / ms, 'tracks_kept_per_pair': float(stt.float().sum().item() / (Bk - 1)),
                                'mean_ransac_iterations': float(itr.float().mean().item())}
        del pts, stt, itr
        fe.close()
        del m, c, dv, dx
    # ---- config 5: 4K, 8 streams x 4 frames, extract 2000 + projection-guided match to the previous frame of the stream;
    #      streams sharded over the ranks, no collective
    W5, H5, NS, NF = 3840, 2160, 8, 4
    mine = [s_ for s_ in range(NS) if s_ % world == rank]
    B5 = len(mine) * NF
    sf5, lw5, lh5 = tb.pyramid_geometry(W5, H5, PSET['n_levels'], PSET['scale'])
    fe = tb.Frontend(ctx, lw5, lh5, max_batch=B5, max_target=PSET['target'])
    cap = fe.kp_cap
    fe.upload_frames(0, np.stack([synth_frame(k, W5, H5, seed0=2000 + 1000 * s_) for s_ in mine for k in range(NF)]))
    pairs = [(i * NF + k, i * NF + k - 1) for i in range(len(mine)) for k in range(1, NF)]
    f1 = np.array([p_[0] for p_ in pairs], np.int32)
    f2 = np.array([p_[1] for p_ in pairs], np.int32)
    fx5 = 2000.0
    T5 = np.tile(np.eye(4, dtype=np.float32), (len(pairs), 1, 1))
    T5[:, 0, 3] = -10.0 / fx5
    cam5 = tb.Pinhole(fx5, fx5, W5 / 2.0, H5 / 2.0, W5, H5, (tb.capi.C.c_float * 5)(0, 0, 0, 0, 0))
    p5 = tb.MatcherParams(30, 50, 30, 1, 30.0)
    kps_d, _, _ = fe.results_device()

    class _Arr:
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {'shape': shape, 'typestr': '<f4', 'data': (ptr, False), 'version': 2}
    kt = torch.as_tensor(_Arr(kps_d, (B5, cap, 7)), device='cuda')
    m5 = torch.zeros((len(pairs), cap, 4), dtype=torch.int32, device='cuda')
    c5 = torch.zeros(len(pairs), dtype=torch.int32, device='cuda')
    v5 = torch.ones((B5, cap), dtype=torch.uint8, device='cuda')
    x5 = torch.zeros((B5, cap, 3), dtype=torch.float32, device='cuda')

    def c5fn():
        fe.build_pyramid(B5)
        fe.extract(B5, sf5, PSET['target'], PSET['initTh'], PSET['minTh'])
        x5[:, :, 0] = (kt[:, :, 0] - W5 / 2.0) / fx5 * 10.0     # map points = keypoints back-projected at 10 m (device plumbing)
        x5[:, :, 1] = (kt[:, :, 1] - H5 / 2.0) / fx5 * 10.0
        x5[:, :, 2] = 10.0
        fe.search_by_projection(p5, cam5, f1, f2, T5, None, v5.data_ptr(), x5.data_ptr(), sf5, m5.data_ptr(), cap, c5.data_ptr())
    ms5 = timed(c5fn, 5, warm=2)
    if world > 1:
        v = torch.tensor([ms5], dtype=torch.float64, device='cuda')
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms5 = float(v.item())
    if rank == 0:
        out['config5'] = {'workload': 'synthetic 3840x2160, %d streams x %d frames: extract 2000 kpts + projection-guided match to the previous frame; streams sharded over %d rank(s)' % (NS, NF, world),
                          'ms_per_step': ms5, 'frames_per_s': NS * NF * 1e3 / ms5, 'projection_matches_rank0': int(c5.sum().item())}
    fe.close()
    ctx.close()
    return out if rank == 0 else None


def parity_block(tb, ctx, cpu, n, kps, desc, counts, desc_dptr, cap, gpu_filtered):
    """Compares the GPU results of the first n frames of the timed configuration (the batch-256 code path: direct launches,
    per-level octree, three streams, tcgen05 matcher) with what the CPU reference leg computed for the very same frames."""
    kp_mis = desc_flip = match_mis = cross_mis = cross_total = 0
    for f in range(n):
        ko, do = cpu['results'][f]
        if int(counts[f]) != len(ko):
            kp_mis += abs(int(counts[f]) - len(ko)) + 1
            continue
        a_, b_ = kps[f, :len(ko)].view(np.uint8).reshape(len(ko), -1), np.ascontiguousarray(ko).view(np.uint8).reshape(len(ko), -1)
        kp_mis += int((a_ != b_).any(axis=1).sum())
        desc_flip += int((desc[f, :len(ko)] != do).any(axis=1).sum())
    for f in range(n - 1):
        exp = cpu['filtered'][f]
        if gpu_filtered[f].tobytes() != np.ascontiguousarray(exp).tobytes():
            match_mis += 1
        got = ctx.match_bf_cross_device(desc_dptr + f * cap * 32, int(counts[f]), desc_dptr + (f + 1) * cap * 32, int(counts[f + 1]))
        cross_total += len(cpu['cross'][f])
        if got.tobytes() != np.ascontiguousarray(cpu['cross'][f]).tobytes():
            cross_mis += 1
    return {'frames': n, 'pairs': n - 1, 'kp_mismatch': kp_mis, 'desc_rows_flipped': desc_flip, 'match_mismatch': match_mis,
            'cross_check_matches': cross_total, 'cross_check_mismatch': cross_mis,
            'checker': 'CPU reference leg (reference ORBextractor.cpp build + cv2 resize/BFMatcher) on the same frames, bytes compared'}


# ------------------------------------------------------------------------------------------ ours
def algorithmic_bytes(lw, lh, B, n_kp_total):
    """compulsory u8 traffic per stage for one batch."""
    px = [int(w) * int(h) for w, h in zip(lw, lh)]
    sparse = n_kp_total * (749 + 512 + 32 + 28)
    return {'pyr_resize': B * (sum(px[:-1]) + sum(px[1:])), 'fast9_cells': B * sum(px), 'gauss7x7_u8': 2 * B * sum(px),
            'orb_describe': sparse, 'orb_describe_fused': sparse}


def run_ours(a):
    if a.fast_th:
        PSET['initTh'], PSET['minTh'] = [float(x) for x in a.fast_th.split(',')]
    import torch
    import torch.distributed as dist
    import trackingbench_slam_b200 as tb
    from trackingbench_slam_b200.synth import synth_stream
    rank, local_rank, world = env_rank()
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback)'
    # multi-GPU end-to-end runs are bound by pinned uploads: keep this rank (and the pinned pages it is about to allocate)
    # on the CPUs next to its GPU (TB_BIND_NUMA=0 disables); the single-rank run keeps every core for the CPU baseline leg
    from trackingbench_slam_b200.hostbind import bind_to_gpu
    bound_cpus = bind_to_gpu(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    B, W, H = a.batch, a.width, a.height
    nl, sc, tgt = PSET['n_levels'], PSET['scale'], PSET['target']
    stream = torch.cuda.Stream()   # a real (non-NULL) stream handle: the ctx launches on it, the events time it
    torch.cuda.set_stream(stream)
    ctx = tb.Context(local_rank, stream.cuda_stream)
    ctx.set_match_impl(1 if a.match_impl == 'tc' else 0)
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=B, max_target=tgt)
    cap = fe.kp_cap
    # inputs: every rank owns its own stream of B frames (weak scaling), pinned on the host
    host = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)
    frames = host.numpy()
    synth_stream(B, W, H, seed0=1234 + rank * B, out=frames)
    P = B - 1
    qf, tf = np.arange(0, P, dtype=np.int32), np.arange(1, B, dtype=np.int32)
    matches = torch.zeros((max(P, 1), cap, 4), dtype=torch.int32, device='cuda')
    mcounts = torch.zeros(max(P, 1), dtype=torch.int32, device='cuda')
    h_kps = torch.empty((B, cap, 7), dtype=torch.float32, pin_memory=True)
    h_desc = torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True)
    h_counts = torch.empty(B, dtype=torch.int32, pin_memory=True)
    h_matches = torch.empty((max(P, 1), cap, 4), dtype=torch.int32, pin_memory=True)
    h_mcounts = torch.empty(max(P, 1), dtype=torch.int32, pin_memory=True)

    def step():
        fe.build_pyramid(B)
        fe.extract(B, sf, tgt, PSET['initTh'], PSET['minTh'])
        if P > 0:
            fe.search_by_bf(qf, tf, PSET['ratio'], PSET['match_minTh'], matches.data_ptr(), cap, mcounts.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    fe.upload_frames_ptr(0, B, host.data_ptr(), W)
    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = B * world * a.steps / (ms / 1000.0)

    # end to end through the C ABI with HOST buffers: every step uploads its frames from pinned memory, runs the path
    # and reads keypoints/descriptors/matches back, all inside the timed region.  NSLOT ctx/stream/frontend sets rotate
    # between steps so that the upload of one step (the PCIe link is the end-to-end bound: 531 MB per step at the
    # measured 55.6 GB/s takes 9.55 ms, as long as the kernels) overlaps the kernels of the previous one.  The timed
    # region starts with an empty pipeline, so K steps cost about K uploads + one step of kernels + one readback.
    # TB_E2E_CHUNKS > 1 additionally splits a step into chunks of the batch (each chunk carries the first frame of the
    # next one as its last frame so the k <-> k+1 matches across the boundary are still produced).  Two chunks are the default
    # since the kernels of a step (6 ms) fit well inside its upload (10 ms): the smaller launches cost kernel time that stays
    # hidden, and only half a step of kernels is left behind the last upload (25.7 k against 25.4 k frames/s at 20 steps,
    # 24.3 k against 23.1 k at 5; four chunks: 24.5 k).
    NCH = max(1, min(int(os.environ.get('TB_E2E_CHUNKS', '2')), B // 16))
    NSLOT = int(os.environ.get('TB_E2E_SLOTS', '3'))
    bounds = [(c * B // NCH, (c + 1) * B // NCH) for c in range(NCH)]
    chunk_slots = []
    for k in range(NSLOT * NCH):
        c = k % NCH
        f_lo, f_hi = bounds[c]
        n_out = f_hi - f_lo                                   # frames whose results this chunk reports
        n_in = n_out + (1 if c < NCH - 1 else 0)              # + the next chunk's first frame (match partner)
        n_pairs = n_in - 1
        st = torch.cuda.Stream()
        c2 = tb.Context(local_rank, st.cuda_stream)
        c2.set_match_impl(1 if a.match_impl == 'tc' else 0)
        f2 = tb.Frontend(c2, lw, lh, max_batch=n_in, max_target=tgt)
        chunk_slots.append(dict(stream=st, ctx=c2, fe=f2, f_lo=f_lo, n_in=n_in, n_out=n_out, n_pairs=n_pairs,
                                qf=np.arange(0, n_pairs, dtype=np.int32), tf=np.arange(1, n_pairs + 1, dtype=np.int32),
                                matches=torch.zeros((max(n_pairs, 1), cap, 4), dtype=torch.int32, device='cuda'),
                                mcounts=torch.zeros(max(n_pairs, 1), dtype=torch.int32, device='cuda'),
                                h_kps=torch.empty((n_out, cap, 7), dtype=torch.float32, pin_memory=True),
                                h_desc=torch.empty((n_out, cap, 32), dtype=torch.uint8, pin_memory=True),
                                h_counts=torch.empty(n_out, dtype=torch.int32, pin_memory=True),
                                h_matches=torch.empty((max(n_pairs, 1), cap, 4), dtype=torch.int32, pin_memory=True),
                                h_mcounts=torch.empty(max(n_pairs, 1), dtype=torch.int32, pin_memory=True)))

    compute_done = [None]

    def chunk_e2e(s):
        fe2 = s['fe']
        fe2.upload_frames_ptr(0, s['n_in'], host.data_ptr() + s['f_lo'] * W * H, W)
        # kernels of consecutive chunks run back to back (not interleaved), so that this chunk's upload
        # overlaps the previous chunk's kernels and the next chunk's upload overlaps these
        if compute_done[0] is not None:
            s['stream'].wait_event(compute_done[0])
        fe2.build_pyramid(s['n_in'])
        fe2.extract(s['n_in'], sf, tgt, PSET['initTh'], PSET['minTh'])
        if s['n_pairs'] > 0:
            fe2.search_by_bf(s['qf'], s['tf'], PSET['ratio'], PSET['match_minTh'], s['matches'].data_ptr(), cap, s['mcounts'].data_ptr())
        ev = torch.cuda.Event()
        ev.record(s['stream'])
        compute_done[0] = ev
        fe2.get_results_batch_async(s['n_out'], s['h_kps'].numpy().view(tb.KP_DTYPE).reshape(s['n_out'], cap), s['h_desc'].numpy(),
                                    s['h_counts'].numpy())
        with torch.cuda.stream(s['stream']):
            s['h_matches'].copy_(s['matches'], non_blocking=True)
            s['h_mcounts'].copy_(s['mcounts'], non_blocking=True)

    def step_e2e(i, reuse):
        for c in range(NCH):
            s = chunk_slots[(i % NSLOT) * NCH + c]
            if reuse:
                s['fe'].status()      # the slot's previous results are complete (and read) before it is reused
            chunk_e2e(s)

    for i in range(NSLOT):
        step_e2e(i, False)
    for s in chunk_slots:
        s['fe'].status()
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        step_e2e(i, i >= NSLOT)
    for s in chunk_slots:
        s['fe'].status()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1000.0)
    e2e_value = B * world * a.steps / (e2e_ms / 1000.0)
    h2d = sum(s['n_in'] for s in chunk_slots[:NCH]) * W * H
    d2h = sum(s['n_out'] * (cap * (28 + 32) + 4) + max(s['n_pairs'], 1) * (cap * 16 + 4) for s in chunk_slots[:NCH])
    li = ((a.steps - 1) % NSLOT) * NCH if a.steps > 0 else 0
    last = chunk_slots[li:li + NCH]
    counts = np.concatenate([s['h_counts'].numpy() for s in last])
    n_matches = int(sum(int(s['h_mcounts'].numpy()[:s['n_pairs']].sum()) for s in last)) if P > 0 else 0
    # the pipelined results are the same bytes as the device-resident run's
    fe.get_results_batch(B, h_kps.numpy().view(tb.KP_DTYPE).reshape(B, cap), h_desc.numpy(), h_counts.numpy())
    h_mcounts.copy_(mcounts)
    torch.cuda.synchronize()
    assert np.array_equal(h_counts.numpy(), counts)
    assert np.array_equal(h_desc.numpy(), np.concatenate([s['h_desc'].numpy() for s in last]))
    if P > 0:
        assert np.array_equal(h_mcounts.numpy()[:P], np.concatenate([s['h_mcounts'].numpy()[:s['n_pairs']] for s in last]))
    for s in chunk_slots:
        s['fe'].close()
        s['ctx'].close()

    # per-kernel times of one step (CUDA events around every launch) for the roofline
    roof, table, hamming = None, {}, None
    if rank == 0:
        ctx.profile_enable(True)
        step()
        prof = ctx.profile_report()
        ctx.profile_enable(False)
        tot = sum(v[1] for v in prof.values())
        alg = algorithmic_bytes(lw, lh, B, int(counts.sum()))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak, peak_src = (peaks['hbm_gbs'], 'measured') if 'hbm_gbs' in peaks else (6650.0, 'fallback')
        for name, (cnt, kms) in prof.items():
            table[name] = {'launches': cnt, 'ms': kms, 'share': kms / tot if tot else 0.0}
            if name in alg and kms > 0:
                table[name]['algorithmic_bytes'] = alg[name]
                table[name]['GBps'] = alg[name] / (kms * 1e-3) / 1e9
        dom = max((n for n in table if 'GBps' in table[n]), key=lambda n: table[n]['ms'])
        traffic = None
        try:  # dram bytes per launch of the dominant kernel from the committed ncu capture, if any
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json'))).get(dom)
        except Exception:
            pass
        # whole step against the same peak: SURVEY 8(d)'s per-frame figure (pyramid + FAST + blur + sparse; the blur term is kept
        # although the dense Gaussian no longer runs, so the number stays comparable with round 1)
        step_bytes = alg['pyr_resize'] + alg['fast9_cells'] + alg['gauss7x7_u8'] + alg['orb_describe']
        roof = {'kernel': dom, 'bound': 'hbm', 'achieved': table[dom]['GBps'], 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                'frac': table[dom]['GBps'] / peak, 'traffic': traffic,
                'launches_per_step': table[dom]['launches'], 'step_share': table[dom]['share'],
                'whole_step': {'algorithmic_bytes': step_bytes, 'achieved': step_bytes / (ms / a.steps * 1e-3) / 1e9,
                               'frac': step_bytes / (ms / a.steps * 1e-3) / 1e9 / peak},
                'note': 'fast9_cells is bound by instruction issue on the integer pipes (ncu: issue active 75 %, ALU pipe 71 %, DRAM 5 %), '
                        'not by HBM: profiles/r02_fast_ncu.md'}
        # Hamming throughput of the matcher kernel alone (both cross-check directions are evaluated)
        mk = 'hamming_i8mma' if a.match_impl == 'tc' else 'hamming_knn'
        pair_evals = 2.0 * float((counts[:-1].astype(np.float64) * counts[1:].astype(np.float64)).sum())
        hamming = {'kernel': mk, 'pair_evaluations_per_step': pair_evals,
                   'Gpairs_per_s': pair_evals / (table[mk]['ms'] * 1e-3) / 1e9 if mk in table and table[mk]['ms'] > 0 else None}
        try:   # tensor roofline of the int8 matcher: 512 int8 ops per pair (K = 256) against the measured tcgen05 int8 peak
            pk = json.load(open(os.path.join(ROOT, 'profiles', 'r02_int8_peak.json')))
            if mk == 'hamming_i8mma' and hamming['Gpairs_per_s']:
                ach = hamming['Gpairs_per_s'] * 512 / 1e6
                hamming['roofline'] = {'bound': 'tensor', 'achieved': ach, 'peak': pk['int8_dense_Pop_per_s'], 'unit': 'Pop/s (int8, dense)',
                                       'frac': ach / pk['int8_dense_Pop_per_s'], 'peak_source': 'measured: scripts/imma_probe.cu, profiles/r02_int8_peak.json',
                                       'note': 'the kernel multiplies K = 288 (the tie-break key rides in a ninth K step): tensor-pipe occupancy is frac * 288/256'}
        except Exception:
            pass
        if a.profile_out:
            json.dump({'per_kernel': table, 'step_ms_profiled': tot}, open(a.profile_out, 'w'), indent=1)

    # sustained figure: the same step repeated for >= 1 s with the clocks sampled (the K-step number above is a burst)
    sustained = None
    if not a.no_extras:
        n_sus = max(a.steps, int(1200.0 / max(ms / a.steps, 1e-3)) + 1)
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        barrier()
        e0.record(stream)
        for _ in range(n_sus):
            step()
        e1.record(stream)
        barrier()
        ms_sus = max_over_ranks(e0.elapsed_time(e1))
        ck2 = sampler2.stop() if rank == 0 else None
        sustained = {'steps': n_sus, 'seconds': ms_sus / 1e3, 'ms_per_step': ms_sus / n_sus, 'value': B * world * n_sus / (ms_sus / 1e3), 'clocks': ck2}

    cpu, parity = None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = a.cpu_sample or max(4, min(2 * cores, 32))
        n = min(n, B)
        cpu_reference_run(frames[:min(n, cores)], cores)          # warm-up (first-touch of the libraries, thread pool)
        r = cpu_reference_run(frames[:n], cores, keep_results=True)
        cpu = {'value': r['fps'], 'unit': 'frames/s', 'cores': cores, 'kind': r['kind'], 'primitives': r['primitives'],
               'sample': 'first %d frames of the same stream: extract all + searchByBF of the %d consecutive pairs (second of two runs)' % (n, n - 1)}
        # parity of the timed configuration against the CPU leg's results for the same frames
        step()
        gk, gd, gc = fe.get_results_batch(B)
        hm, hc = matches.cpu().numpy(), mcounts.cpu().numpy()
        gpu_filtered = [hm[f, :hc[f]].copy().view(tb.MATCH_DTYPE).reshape(-1) for f in range(n - 1)]
        _, desc_dptr, _ = fe.results_device()
        parity = parity_block(tb, ctx, r, n, gk, gd, gc, desc_dptr, cap, gpu_filtered)

    config4 = extras = None
    if not a.no_extras:
        config4 = config4_block(torch, dist, ctx, world, rank, a.knn_n)
        extras = extras_block(torch, dist, tb, local_rank, world, rank, a)

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
                'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
                'data': 'synthetic',
                'config': {'workload': workload_name(a), 'levels': 8, 'scale': 1.2, 'fast_th': [20, 7], 'frames_per_gpu': B,
                           'sharding': 'frame streams, %d rank(s), no collective' % world,
                           'l2': 'inputs (%.0f MB of frames per step) exceed the 126 MB L2' % (B * W * H / 1e6),
                           'mean_kpts_per_frame': float(counts.mean()), 'matches_per_step': n_matches, 'match_impl': a.match_impl,
                           'e2e_pipeline': '%d chunk(s) per step x %d rotating slots; bound: PCIe upload of the frames' % (NCH, NSLOT),
                           'host_binding': ('rank 0 bound to %d CPUs local to its GPU' % len(bound_cpus)) if bound_cpus else 'none'},
                'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
                'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'hamming': hamming,
                'per_kernel_ms': {k: round(v['ms'], 3) for k, v in table.items()},
                'parity': parity, 'sustained': sustained, 'config4': config4, 'other_configs': extras}
        print(json.dumps(line))
    fe.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)

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
def cpu_reference_throughput(frames, n_threads):
    """The reference's CPU path on `frames` (consecutive synthetic frames): per frame the pyramid
    (oracle restatement of Frame::ComputePyramid) + ORBExtractor::operator() (oracle/_ref = the
    reference's own ORBextractor.cpp compiled unmodified, else the C port), then searchByBF(k,k+1).
    Frames are farmed over host threads (ctypes releases the GIL). Returns (frames/s, kind)."""
    import oracle as O
    from concurrent.futures import ThreadPoolExecutor
    kind = 'reference' if O.ref_available() else 'port'
    nl, sc, tgt, a, b = PSET['n_levels'], PSET['scale'], PSET['target'], PSET['initTh'], PSET['minTh']
    local = threading.local()

    def extract(img):
        lv, sf = O.pyramid(img, nl, sc)
        if kind == 'reference':
            if not hasattr(local, 'ex'):
                local.ex = O.RefExtractor()
            return local.ex(lv, sf, tgt, a, b)
        k, d, _ = O.orb_extract(lv, sf, tgt, a, b)
        return k, d

    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:
        res = list(ex.map(extract, list(frames)))
        list(ex.map(lambda i: O.search_by_bf(res[i][1], res[i + 1][1], PSET['ratio'], PSET['match_minTh']), range(len(res) - 1)))
    dt = time.perf_counter() - t0
    return len(frames) / dt, kind


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
        v, kind = cpu_reference_throughput(frames, cores)
        if i >= a.warmup:
            vals.append(v)
        if i == 0 and n / v > 40:  # keep the whole run within a few minutes on slow hosts
            a.warmup, a.steps = min(a.warmup, 1), min(a.steps, 2)
            if len(vals) >= a.steps:
                break
    value = float(np.mean(vals)) if vals else v
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': a.gpus, 'steps': len(vals) or 1,
            'warmup': a.warmup, 'ms_per_step': 1000.0 * n / value, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8', 'data': 'synthetic', 'config': {'workload': workload_name(a), 'levels': 8, 'scale': 1.2, 'fast_th': [20, 7]},
            'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
                             'sample': '%d consecutive frames per step (extract all + searchByBF of the %d pairs)' % (n, n - 1)},
            'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ ours
def algorithmic_bytes(lw, lh, B, n_kp_total):
    """compulsory u8 traffic per stage for one batch."""
    px = [int(w) * int(h) for w, h in zip(lw, lh)]
    return {'pyr_resize': B * (sum(px[:-1]) + sum(px[1:])), 'fast9_cells': B * sum(px), 'gauss7x7_u8': 2 * B * sum(px),
            'orb_describe': n_kp_total * (749 + 512 + 32 + 28)}


def run_ours(a):
    if a.fast_th:
        PSET['initTh'], PSET['minTh'] = [float(x) for x in a.fast_th.split(',')]
    import torch
    import torch.distributed as dist
    import trackingbench_slam_b200 as tb
    from trackingbench_slam_b200.synth import synth_stream
    rank, local_rank, world = env_rank()
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback)'
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
    # next one as its last frame so the k <-> k+1 matches across the boundary are still produced); measured slower
    # (smaller launches are less efficient: 64-frame batches cost +15 % kernel time), so the default is 1.
    NCH = max(1, min(int(os.environ.get('TB_E2E_CHUNKS', '1')), B // 16))
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
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json'))).get(dom)
        except Exception:
            pass
        roof = {'kernel': dom, 'bound': 'hbm', 'achieved': table[dom]['GBps'], 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                'frac': table[dom]['GBps'] / peak, 'traffic': traffic,
                'launches_per_step': table[dom]['launches'], 'step_share': table[dom]['share']}
        # Hamming throughput of the matcher kernel alone (both cross-check directions are evaluated)
        mk = 'hamming_i8mma' if a.match_impl == 'tc' else 'hamming_knn'
        pair_evals = 2.0 * float((counts[:-1].astype(np.float64) * counts[1:].astype(np.float64)).sum())
        hamming = {'kernel': mk, 'pair_evaluations_per_step': pair_evals,
                   'Gpairs_per_s': pair_evals / (table[mk]['ms'] * 1e-3) / 1e9 if mk in table and table[mk]['ms'] > 0 else None}
        if a.profile_out:
            json.dump({'per_kernel': table, 'step_ms_profiled': tot}, open(a.profile_out, 'w'), indent=1)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = a.cpu_sample or max(4, min(2 * cores, 32))
        v, kind = cpu_reference_throughput(frames[:n], cores)
        cpu = {'value': v, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
               'sample': 'first %d frames of the same stream: extract all + searchByBF of the %d consecutive pairs' % (n, n - 1)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
                'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
                'data': 'synthetic',
                'config': {'workload': workload_name(a), 'levels': 8, 'scale': 1.2, 'fast_th': [20, 7], 'frames_per_gpu': B,
                           'sharding': 'frame streams, %d rank(s), no collective' % world,
                           'l2': 'inputs (%.0f MB of frames per step) exceed the 126 MB L2' % (B * W * H / 1e6),
                           'mean_kpts_per_frame': float(counts.mean()), 'matches_per_step': n_matches, 'match_impl': a.match_impl,
                           'e2e_pipeline': '%d chunk(s) per step x %d rotating slots; bound: PCIe upload of the frames' % (NCH, NSLOT)},
                'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
                'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'hamming': hamming,
                'per_kernel_ms': {k: round(v['ms'], 3) for k, v in table.items()}}
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

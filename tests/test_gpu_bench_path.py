"""Parity of the code path bench.py times (BASELINE config 3): batches of >= 32 frames take direct launches
(no CUDA graph), one octree launch per level on a side stream, three streams inside tb_frontend_extract and the tcgen05
matcher for the k <-> k+1 searchByBF — none of which the small-batch tests reach.  1080p, 8 levels x 1/1.2, 2000
keypoints, FAST 20/7; a sample of the frames and every match list between consecutive frames is compared with the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NL, SC, TGT, A, B = 8, 1.0 / 1.2, 2000, 20.0, 7.0


@pytest.mark.parametrize('batch', [40, 64])
def test_large_batch_1080p_extract_and_match(tb, oracle, batch):
    from trackingbench_slam_b200.synth import synth_stream
    W, H = 1920, 1080
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = tb.Context(0, stream.cuda_stream)
        ctx.set_match_impl(1)                       # tcgen05 int8 matcher, as in bench.py
        sf, lw, lh = tb.pyramid_geometry(W, H, NL, SC)
        fe = tb.Frontend(ctx, lw, lh, max_batch=batch, max_target=TGT)
        try:
            frames = synth_stream(batch, W, H, seed0=77 + batch)
            fe.upload_frames(0, frames)
            cap = fe.kp_cap
            P = batch - 1
            m = torch.zeros((P, cap, 4), dtype=torch.int32, device='cuda')
            c = torch.zeros(P, dtype=torch.int32, device='cuda')
            for _ in range(2):                      # twice: buffers are reused between steps exactly as in the benchmark loop
                fe.build_pyramid(batch)
                fe.extract(batch, sf, TGT, A, B)
                fe.search_by_bf(np.arange(0, P, dtype=np.int32), np.arange(1, batch, dtype=np.int32), 10.0, 30.0, m.data_ptr(), cap, c.data_ptr())
            ctx.sync()
            kps, desc, counts = fe.get_results_batch(batch)
            hm, hc = m.cpu().numpy(), c.cpu().numpy()
            _, desc_dptr, _ = fe.results_device()
            # runs of consecutive frames at the start, across the 32-frame mark and at the end of the batch
            sample = sorted(set([0, 1, 2, 3] + [30, 31, 32, 33, 34] + [batch - 3, batch - 2, batch - 1]))
            ref = {}
            for f in sample:
                lv, sfo = oracle.pyramid(frames[f], NL, SC)
                ko, do, _ = oracle.orb_extract(lv, sfo, TGT, A, B)
                ref[f] = (ko, do)
                assert counts[f] == len(ko), (f, counts[f], len(ko))
                assert kps[f, :counts[f]].tobytes() == ko.tobytes(), 'keypoints of frame %d' % f
                assert np.array_equal(desc[f, :counts[f]], do), 'descriptors of frame %d' % f
            assert len(sample) >= 8
            n_pairs = 0
            for f in sample:
                if f + 1 not in ref:
                    continue
                exp = oracle.search_by_bf(ref[f][1], ref[f + 1][1], 10.0, 30.0)
                got = hm[f, :hc[f]].copy().view(tb.MATCH_DTYPE).reshape(-1)
                assert hc[f] == len(exp) and got.tobytes() == exp.tobytes(), 'searchByBF(%d,%d)' % (f, f + 1)
                # the filter keeps nothing on this stream (min distance 0): compare the cross-check list itself as well
                cross = ctx.match_bf_cross_device(desc_dptr + f * cap * 32, int(counts[f]), desc_dptr + (f + 1) * cap * 32, int(counts[f + 1]))
                expc = oracle.bf_cross(ref[f][1], ref[f + 1][1])
                assert len(expc) > 500 and cross.tobytes() == expc.tobytes(), 'cross-check list (%d,%d)' % (f, f + 1)
                n_pairs += 1
            assert n_pairs >= 7
        finally:
            fe.close()
            ctx.close()


@pytest.mark.parametrize('impl', [0, 1])
def test_sharded_knn2_world1_and_cross_check(tb, oracle, impl):
    """BASELINE config 4 entry (sharding.sharded_knn2 / sharded_knn2_cross) on one rank against the oracle."""
    from trackingbench_slam_b200.sharding import sharded_knn2, sharded_knn2_cross
    from trackingbench_slam_b200.synth import synth_descriptors
    # the sharding helpers do their plumbing (slices, gathers, the cross-check) with torch on the CURRENT stream: the ctx must
    # launch on that stream too, exactly as bench.py sets it up
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = tb.Context(0, stream.cuda_stream)
        ctx.set_match_impl(impl)
        try:
            for n in (1, 257, 3001):
                Q, T = synth_descriptors(n, seed=n)
                q, t = torch.from_numpy(Q).cuda(), torch.from_numpy(T).cuda()
                ev = []
                idx, best, second = sharded_knn2(ctx, q, t, None, ev)
                ctx.sync()
                assert len(ev) == 3
                i0, b0, s0 = oracle.knn2(Q, T)
                assert np.array_equal(idx.cpu().numpy(), i0) and np.array_equal(best.cpu().numpy(), b0) and np.array_equal(second.cpu().numpy(), s0)
                i12, b12, s12, mutual = sharded_knn2_cross(ctx, q, t)
                ctx.sync()
                bf = oracle.bf_cross(Q, T)
                mu = mutual.cpu().numpy()
                assert np.array_equal(np.flatnonzero(mu), bf['queryIdx'])
                assert np.array_equal(i12.cpu().numpy()[mu], bf['trainIdx']) and np.array_equal(b12.cpu().numpy()[mu].astype(np.float32), bf['distance'])
        finally:
            ctx.close()


def test_side_stream_pyramid_is_joined_by_every_reader(tb, oracle):
    """Batches of >= 32 frames build their pyramid on a side stream (FAST of level l only waits for level l). Everything
    else has to wait for the whole pyramid: reading the levels right after build_pyramid, overwriting level 0 with the next
    batch while the coarse levels may still be read from it, a second frontend on the same ctx in between."""
    from trackingbench_slam_b200.synth import synth_stream
    W, H, nl, sc, tgt = 640, 360, 6, 1.0 / 1.2, 500
    batch = 36
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = tb.Context(0, stream.cuda_stream)
        sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
        fe = tb.Frontend(ctx, lw, lh, max_batch=batch, max_target=tgt)
        sf2, lw2, lh2 = tb.pyramid_geometry(320, 200, 3, 0.8)
        fe2 = tb.Frontend(ctx, lw2, lh2, max_batch=batch, max_target=tgt)
        try:
            fa, fb = synth_stream(batch, W, H, seed0=5), synth_stream(batch, W, H, seed0=900)
            small = synth_stream(batch, 320, 200, seed0=31)
            # (1) levels read straight after the build
            fe.upload_frames(0, fa)
            fe.build_pyramid(batch)
            for f in (0, batch - 1):
                lv, _ = oracle.pyramid(fa[f], nl, sc)
                for l in range(nl):
                    assert np.array_equal(fe.download_level(f, l), lv[l]), ('levels after build', f, l)
            # (2) the next batch overwrites level 0 right behind a build; a second frontend's build and extract in between
            for rep in range(3):
                fe.upload_frames(0, fa)
                fe.build_pyramid(batch)
                fe2.upload_frames(0, small)
                fe2.build_pyramid(batch)
                fe.upload_frames(0, fb)
                fe.build_pyramid(batch)
                fe2.extract(batch, sf2, tgt, 20.0, 7.0)
                fe.extract(batch, sf, tgt, 20.0, 7.0)
            ctx.sync()
            kps, desc, counts = fe.get_results_batch(batch)
            for f in (0, 17, batch - 1):
                lv, sfo = oracle.pyramid(fb[f], nl, sc)
                ko, do, _ = oracle.orb_extract(lv, sfo, tgt, 20.0, 7.0)
                assert counts[f] == len(ko) and kps[f, :counts[f]].tobytes() == ko.tobytes() and np.array_equal(desc[f, :counts[f]], do), f
            k2, d2, c2 = fe2.get_results_batch(batch)
            for f in (0, batch - 1):
                lv, sfo = oracle.pyramid(small[f], 3, 0.8)
                ko, do, _ = oracle.orb_extract(lv, sfo, tgt, 20.0, 7.0)
                assert c2[f] == len(ko) and k2[f, :c2[f]].tobytes() == ko.tobytes() and np.array_equal(d2[f, :c2[f]], do), ('second frontend', f)
        finally:
            fe.close()
            fe2.close()
            ctx.close()

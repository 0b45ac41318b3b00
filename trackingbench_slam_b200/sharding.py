"""Multi-GPU partitioning of the path: one process per GPU.

* frame streams / batches shard by FRAME with no data-path collective (extract + frame-to-frame or
  projection matching are independent per frame pair);
* large brute-force matching (BASELINE config 4) shards QUERY ROWS, the train set is replicated and the
  per-row (argbest, best, second) triples are assembled with one all_gather — the only collective.
Works with any torch.distributed backend: nccl on GPUs, gloo in the CPU tests.
The plumbing here (slices, the gather, the cross-check) is torch on the CURRENT stream: create the tb.Context on that stream
(tb.Context(device, torch.cuda.current_stream().cuda_stream)) so that kernels and plumbing are ordered.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def frame_shards(n_frames, world):
    return [shard_range(n_frames, r, world) for r in range(world)]


def gather_rows(local, n_total, group=None):
    """all_gather of row-sharded tensors (shard_range layout) into the full [n_total, ...] tensor."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [hi - lo for lo, hi in (shard_range(n_total, r, world) for r in range(world))]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def sharded_knn2(ctx, q_dev, t_dev, group=None, events=None):
    """Config 4 on GPUs: every rank holds the full Q and T on its device, computes best / second /
    argbest for its query rows with the CUDA kernel and gathers. Returns full (idx, best, second).
    events (optional list) receives three CUDA events recorded on the current stream: before the kernel,
    after the kernel, after the gather — so the caller can split kernel time from collective time."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = q_dev.shape[0]
    lo, hi = shard_range(n, rank, world)
    m = hi - lo
    buf = torch.empty((3, m), dtype=torch.int32, device=q_dev.device)   # rows: argbest, best, second — gathered as one block
    idx, best, second = buf[0], buf[1], buf[2]

    def mark():
        if events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream())
            events.append(e)

    mark()
    if m:
        ctx.knn2_device(q_dev[lo:hi].data_ptr(), m, t_dev.data_ptr(), t_dev.shape[0], idx.data_ptr(), best.data_ptr(), second.data_ptr())
    mark()
    if world == 1:
        mark()
        return idx, best, second
    if n % world == 0:
        # equal shards: ONE collective straight into the result block and one re-ordering copy (the padded list form below
        # costs a dozen small launches, which at 8 GPUs was a quarter of config 4's time)
        out = torch.empty((world, 3, m), dtype=torch.int32, device=q_dev.device)
        dist.all_gather_into_tensor(out, buf, group=group)
        full = out.permute(1, 0, 2).reshape(3, n)
        mark()
        return full[0], full[1], full[2]
    full = gather_rows(buf.t().contiguous(), n, group)
    mark()
    return full[:, 0].contiguous(), full[:, 1].contiguous(), full[:, 2].contiguous()


def cross_check(idx12, idx21):
    """cv::BFMatcher(crossCheck=true) on gathered argbest arrays: query q keeps its match t = idx12[q] iff idx21[t] == q
    (lowest index on ties in both directions, as the kernels produce them)."""
    n = idx12.shape[0]
    return idx21[idx12.long()] == torch.arange(n, dtype=idx21.dtype, device=idx12.device)


def sharded_knn2_cross(ctx, q_dev, t_dev, group=None, events=None):
    """SURVEY.md §8(e) cross-check variant of config 4: the row-sharded kNN-2 of Q against T, the transposed problem
    (T's rows sharded, against Q) and a second gather; returns (idx, best, second, mutual) for the query rows."""
    i12, b12, s12 = sharded_knn2(ctx, q_dev, t_dev, group, events)
    i21, _, _ = sharded_knn2(ctx, t_dev, q_dev, group, events)
    return i12, b12, s12, cross_check(i12, i21)

"""Sharded single-frame mode (BASELINE config 4): ONE tracking problem whose candidate list is
split over the ranks of a torch.distributed group, with one all-reduce of the 32 normal-equation
sums per Gauss-Newton sweep (NCCL over NVLink on GPUs) -- the only collective of the hot path.

The loop is written against a small backend protocol so the host-side logic (partitioning,
reduction, identical termination on every rank) is also testable on CPU with gloo:
    begin(rank, nranks) ; accumulate() -> 32-element float64 tensor ; update(tensor) -> done ;
    result()
"""


def partition(n, rank, nranks):
    """Contiguous range [lo, hi) of `n` candidate rows owned by `rank` (same rule as the
    kernels: lo = n*rank/nranks)."""
    return (n * rank) // nranks, (n * (rank + 1)) // nranks


class TrackerShardBackend:
    """Protocol adapter over uw_slam_b200.Tracker (C ABI uwt_shard_*)."""

    def __init__(self, tracker, prev_slot, cur_slot, init_pose=None):
        import torch
        self.t, self.prev, self.cur, self.init = tracker, prev_slot, cur_slot, init_pose
        self.dev = torch.device("cuda", tracker.cfg.device)
        self.sums = torch.zeros(32, dtype=torch.float64, device=self.dev)
        self.stream = torch.cuda.ExternalStream(tracker.stream_ptr(), device=self.dev)

    def begin(self, rank, nranks):
        self.t.ShardBegin(self.prev, self.cur, rank, nranks, self.init)

    def accumulate(self):
        self.t.ShardAccumulate(self.sums.data_ptr())
        return self.sums

    def update(self, sums):
        return self.t.ShardUpdate(sums.data_ptr())

    def result(self):
        return self.t.ShardResult()


def estimate_pose_sharded(backend, group=None, max_sweeps=10_000):
    """Runs the sharded Gauss-Newton loop on this rank.  Every rank returns the same pose."""
    import contextlib
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    backend.begin(rank, world)
    ctx = contextlib.nullcontext()
    stream = getattr(backend, "stream", None)
    if stream is not None:
        import torch
        ctx = torch.cuda.stream(stream)  # NCCL is ordered with the library's own stream
    sweeps = 0
    with ctx:
        while True:
            sums = backend.accumulate()
            if world > 1:
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            sweeps += 1
            if backend.update(sums) or sweeps >= max_sweeps:
                break
    pose, stats = backend.result()
    return pose, stats, sweeps


def connect_fused(tracker, group=None):
    """One-time setup of the fused (in-kernel all-reduce) sharded mode: exchange the CUDA IPC
    handles of the per-rank mailboxes with an all-gather and map the peers' mailboxes."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mine = tracker.ShardIpcExport()
    if world == 1:
        tracker.ShardIpcConnect(0, 1, mine)
        return rank, world
    dev = torch.device("cuda", tracker.cfg.device)
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    handles = b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)
    tracker.ShardIpcConnect(rank, world, handles)
    dist.barrier(group=group)  # nobody starts a sweep before every mailbox is mapped + zeroed
    return rank, world


def estimate_pose_sharded_fused(tracker, prev_slot, cur_slot, init_pose=None, grid=0):
    """Whole sharded Gauss-Newton loop in ONE kernel per rank; the per-sweep all-reduce happens
    inside the kernel through peer-mapped mailboxes.  Call on every rank of the group."""
    tracker.ShardEstimateFusedAsync(prev_slot, cur_slot, init_pose, grid)
    return tracker.ShardEstimateFusedWait()

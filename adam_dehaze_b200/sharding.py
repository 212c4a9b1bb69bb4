"""Image sharding across the GPUs of one box: every image is independent in eval mode, so rank r of W
takes a contiguous slice of the batch, classifies, buckets and dehazes it locally and writes its slice of the output.
No data-path collective exists at inference."""


def shard_bounds(total, rank, world):
    """[lo, hi) of rank's contiguous shard; the first total % world ranks carry one extra image."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total, world):
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


# ----------------------------------------------------------------------------- optional re-balance on predicted class
# A shard heavy in Complex images costs ~13x a shard of Light ones (8.5 vs 0.66 ms per image at 1024x2048), so image-count
# sharding alone leaves the step time at the mercy of the mix ("the only scaling risk").  After HDEN has
# classified the local shard, the ranks all-gather their class ids (a few hundred bytes) and trade whole images so that
# every rank holds an equal share of EVERY class; the images travel by one NCCL all-to-all over NVLink (25 MB each,
# ~10-20 ms for half a 256-image shard), are dehazed where they land, and travel back the same way.
def rebalance_plan(labels_by_rank, n_classes=3):
    """labels_by_rank[r]: class id of every image rank r holds.  Returns send[r][d] = ascending local indices rank r sends to
    rank d (d == r: the images it keeps), such that every rank ends with its equal share (+-1) of every class and an image
    moves only when its rank holds more than its share."""
    world = len(labels_by_rank)
    send = [[[] for _ in range(world)] for _ in range(world)]
    for k in range(n_classes):
        have = [[i for i, v in enumerate(lab) if int(v) == k] for lab in labels_by_rank]
        total = sum(len(h) for h in have)
        base, extra = divmod(total, world)
        # the +1 remainders rotate with the class index so that no rank collects all of them
        target = [base + (1 if ((r - k) % world) < extra else 0) for r in range(world)]
        surplus = []
        for r in range(world):
            keep = min(len(have[r]), target[r])
            send[r][r] += have[r][:keep]
            surplus += [(r, i) for i in have[r][keep:]]
        pos = 0
        for d in range(world):
            need = target[d] - min(len(have[d]), target[d])
            for r, i in surplus[pos:pos + need]:
                send[r][d].append(i)
            pos += need
    for r in range(world):
        for d in range(world):
            send[r][d].sort()
    return send


def unrouted_stay(labels_by_rank, send, n_classes=3):
    """Images whose class id is outside [0, n_classes) are not part of the plan: they stay where they are."""
    for r, lab in enumerate(labels_by_rank):
        send[r][r] = sorted(set(send[r][r]) | {i for i, v in enumerate(lab) if not (0 <= int(v) < n_classes)})
    return send


class Exchange:
    """One re-balance round trip for rank `rank`: forward(x) -> the rows this rank processes (grouped by source rank, each
    group in the sender's ascending order) and their labels; backward(y) -> results returned to the rank that owns them, in
    that rank's original row order."""

    def __init__(self, labels_by_rank, rank, group=None, n_classes=3):
        import torch
        self.rank, self.group = rank, group
        self.world = len(labels_by_rank)
        send = unrouted_stay(labels_by_rank, rebalance_plan(labels_by_rank, n_classes), n_classes)
        self.in_splits = [len(send[rank][d]) for d in range(self.world)]
        self.out_splits = [len(send[s][rank]) for s in range(self.world)]
        self.order = torch.tensor([i for d in range(self.world) for i in send[rank][d]], dtype=torch.int64)
        self.labels = torch.tensor([int(labels_by_rank[s][i]) for s in range(self.world) for i in send[s][rank]], dtype=torch.int64)
        self.moved = sum(self.in_splits) - self.in_splits[rank]

    def forward(self, x):
        import torch
        import torch.distributed as dist
        order = self.order.to(x.device)
        packed = x.index_select(0, order)
        out = torch.empty((sum(self.out_splits),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_to_all_single(out, packed, self.out_splits, self.in_splits, group=self.group)
        return out, self.labels.to(x.device)

    def backward(self, y):
        import torch
        import torch.distributed as dist
        back = torch.empty((sum(self.in_splits),) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        dist.all_to_all_single(back, y.contiguous(), self.in_splits, self.out_splits, group=self.group)
        out = torch.empty_like(back)
        out.index_copy_(0, self.order.to(y.device), back)
        return out

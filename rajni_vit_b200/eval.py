"""evaluate_model — drop-in for rajni/eval.py:6-75, plus data-parallel sharding.

Same signature, same return value ``(top-1 accuracy in %, images/s)`` and the same
timing contract: warm-up batches first, then for every batch the host->device copy
happens OUTSIDE the timed region and only ``model(images)`` is timed between two
device synchronisations (eval.py:48-59).

Differences, all deliberate:
  * the device is synchronised whenever it is a CUDA device, whether it was passed
    as the string "cuda" or as a ``torch.device`` (the reference only syncs for the
    exact string, so its CLI timings are unsynchronised);
  * when ``torch.distributed`` is initialised with world_size > 1 the batch is
    sharded contiguously across ranks (images are independent); the
    only collective is one all-reduce of (correct, total, images) and one MAX of
    the elapsed time at the very end, so throughput = all images / slowest rank.
    A foreign loader is sharded by slicing each batch after loading; a loader from
    ``rajni_vit_b200.data.sharded_loader`` (``rajni_sharded = True``) already yields
    this rank's shard - 1/N of the decode work - and is used as it comes.
"""
from __future__ import annotations

import time

import torch

try:                                    # progress bar is cosmetic; never a hard dependency
    from tqdm import tqdm
except Exception:                       # pragma: no cover
    tqdm = None


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of a batch for this rank (first ranks get the remainder)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _is_cuda(device) -> bool:
    return torch.device(device).type == "cuda"


@torch.no_grad()
def evaluate_model(model, dataloader, device="cuda", max_batches=None, warmup=5, shard=None, progress=True):
    dist = _dist()
    if shard is None:
        shard = dist is not None
    rank, world = (dist.get_rank(), dist.get_world_size()) if (dist is not None and shard) else (0, 1)
    cuda = _is_cuda(device)

    model.eval()
    model.to(device)

    presharded = bool(getattr(dataloader, "rajni_sharded", False))      # data.sharded_loader: batches are this rank's already

    def take(t):
        if world == 1 or presharded:
            return t
        lo, hi = shard_bounds(t.shape[0], rank, world)
        return t[lo:hi]

    # ---- warm-up (eval.py:17-26): cycles the loader if it is shorter than `warmup`
    if rank == 0:
        print(f"Warming up {warmup} batches")
    it = iter(dataloader)
    for _ in range(warmup):
        try:
            x, _ = next(it)
        except StopIteration:
            it = iter(dataloader)
            x, _ = next(it)
        x = take(x)
        if x.shape[0]:
            model(x.to(device))
    if cuda:
        torch.cuda.synchronize(device)

    correct = total = total_images = 0
    total_time = 0.0
    try:
        n_batches = max_batches if max_batches is not None else len(dataloader)
    except TypeError:
        n_batches = None
    bar = None
    batches = dataloader
    if progress and tqdm is not None and rank == 0:
        bar = batches = tqdm(dataloader, desc="Evaluating", total=n_batches, leave=False)

    for i, (images, labels) in enumerate(batches):
        if max_batches is not None and i >= max_batches:
            break
        images = take(images).to(device)          # H2D outside the timed region (eval.py:48-49)
        labels = take(labels).to(device)
        if images.shape[0] == 0:
            continue
        if cuda:
            torch.cuda.synchronize(device)
        start = time.time()
        logits = model(images)
        if cuda:
            torch.cuda.synchronize(device)
        total_time += time.time() - start

        correct += (logits.argmax(dim=1) == labels).sum().item()
        total += labels.size(0)
        total_images += images.size(0)
        if bar is not None and total > 0:
            bar.set_postfix(acc=f"{100.0 * correct / total:.2f}%",
                            imgs_per_s=f"{total_images / max(total_time, 1e-6):.1f}")
    if bar is not None:
        bar.close()

    if world > 1:
        red_dev = device if cuda else "cpu"
        counts = torch.tensor([correct, total, total_images], dtype=torch.float64, device=red_dev)
        elapsed = torch.tensor([total_time], dtype=torch.float64, device=red_dev)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        correct, total, total_images = (int(v) for v in counts.tolist())
        total_time = float(elapsed.item())

    acc = 100.0 * correct / max(total, 1)
    throughput = total_images / max(total_time, 1e-6)
    return acc, throughput

"""Host-side helpers the engine needs (restated from /root/reference/utils/__init__.py; no arithmetic on
the hot path lives here)."""
from __future__ import annotations

import datetime
import math
import os
import time
from collections import defaultdict, deque

import numpy as np
import torch
import torch.distributed as dist


def is_dist_avail_and_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_world_size() -> int:
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process() -> bool:
    return get_rank() == 0


def init_distributed_mode(args=None, backend: str = "nccl"):
    """Single-box torchrun rendezvous (replaces utils/__init__.py:482-547 + the XLA variant 26-96).
    Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment."""
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ or int(os.environ["WORLD_SIZE"]) <= 1:
        if args is not None:
            args.distributed, args.rank, args.world_size, args.gpu = False, 0, 1, 0
        return False
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(minutes=30))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world, timeout=datetime.timedelta(minutes=30))
    dist.barrier()
    if args is not None:
        args.distributed, args.rank, args.world_size, args.gpu = True, rank, world, local
    return True


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0,
                     warmup_steps=-1):
    """Per-iteration cosine schedule with linear warm-up (/root/reference/utils/__init__.py:667-684)."""
    warmup_schedule = np.array([])
    warmup_iters = warmup_epochs * niter_per_ep
    if warmup_steps > 0:
        warmup_iters = warmup_steps
    if warmup_epochs > 0:
        warmup_schedule = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    schedule = np.array(
        [final_value + 0.5 * (base_value - final_value) * (1 + math.cos(math.pi * i / (len(iters)))) for i in iters])
    schedule = np.concatenate((warmup_schedule, schedule))
    assert len(schedule) == epochs * niter_per_ep
    return schedule


class SmoothedValue:
    """Window median + global average (/root/reference/utils/__init__.py:103-160)."""

    def __init__(self, window_size=20, fmt=None):
        self.deque = deque(maxlen=window_size)
        self.total = 0.0
        self.count = 0
        self.fmt = fmt or "{median:.4f} ({global_avg:.4f})"

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        if not is_dist_avail_and_initialized():
            return
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device=dev)
        dist.barrier()
        dist.all_reduce(t)
        t = t.tolist()
        self.count = int(t[0])
        self.total = t[1]

    @property
    def median(self):
        return float(np.median(list(self.deque))) if self.deque else float("nan")

    @property
    def avg(self):
        return float(np.mean(list(self.deque))) if self.deque else float("nan")

    @property
    def global_avg(self):
        return self.total / max(self.count, 1)

    @property
    def max(self):
        return max(self.deque)

    @property
    def value(self):
        return self.deque[-1]

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg,
                               max=self.max if self.deque else float("nan"),
                               value=self.value if self.deque else float("nan"))


class MetricLogger:
    def __init__(self, delimiter="\t"):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                v = v.item()
            self.meters[k].update(float(v))

    def synchronize_between_processes(self):
        for meter in self.meters.values():
            meter.synchronize_between_processes()

    def __str__(self):
        return self.delimiter.join(f"{name}: {meter}" for name, meter in self.meters.items())

    def log_every(self, iterable, print_freq, header=None, quiet=False):
        i = 0
        header = header or ""
        start = time.time()
        for obj in iterable:
            yield obj
            if not quiet and print_freq and i % print_freq == 0 and is_main_process():
                print(f"{header} [{i}] {self} elapsed: {time.time() - start:.1f}s", flush=True)
            i += 1


class DevicePrefetcher:
    """Host -> device input path of the training loop (replaces torch_xla's MpDeviceLoader, /root/reference/main.py:1017):
    batch i+1 is copied from (pinned) host memory on a side stream while batch i is being trained on, so the H2D
    copy (155 MB per 256 x 3 x 224 x 224 fp32 batch) never sits on the compute stream's critical path."""

    def __init__(self, loader, device: torch.device):
        self.loader = loader
        self.device = device
        self.stream = torch.cuda.Stream(device)

    def __len__(self):
        return len(self.loader)

    def _fetch(self, it):
        try:
            batch = next(it)
        except StopIteration:
            return None
        with torch.cuda.stream(self.stream):
            return tuple(t.to(self.device, non_blocking=True) if torch.is_tensor(t) else t for t in batch)

    def __iter__(self):
        it = iter(self.loader)
        nxt = self._fetch(it)
        while nxt is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_stream(self.stream)          # batch i has landed
            for t in nxt:
                if torch.is_tensor(t):
                    t.record_stream(cur)          # keep its memory until the compute stream is done with it
            batch = nxt
            nxt = self._fetch(it)                 # batch i+1 starts copying now, overlapping the step on batch i
            yield batch


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """timm.utils.accuracy: top-k accuracy in percent."""
    maxk = min(max(topk), output.size(1))
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    return [correct[: min(k, maxk)].reshape(-1).float().sum(0) * 100.0 / batch_size for k in topk]

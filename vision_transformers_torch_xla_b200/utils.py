"""Host-side helpers the engine needs (restated from /root/reference/utils/__init__.py; no arithmetic on
the hot path lives here)."""
from __future__ import annotations

import datetime
import math
import os
import time
from collections import defaultdict, deque

import numpy as np
from typing import Optional

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
    """Per-iteration schedule: linear warm-up, then half a cosine from ``base_value`` to ``final_value``.

    Same values as /root/reference/utils/__init__.py:667-684, element for element (pinned bit-for-bit by
    tests/test_ref_fixtures.py against arrays the reference function itself produced), computed as one vector
    expression.  The reference's two conventions are kept: ``warmup_steps > 0`` overrides the warm-up LENGTH, but the
    ramp is only emitted when ``warmup_epochs > 0`` (otherwise the lengths no longer add up and the assert fires)."""
    total = int(epochs * niter_per_ep)
    n_warm = int(warmup_steps if warmup_steps > 0 else warmup_epochs * niter_per_ep)
    ramp = np.linspace(start_warmup_value, base_value, n_warm) if warmup_epochs > 0 else np.empty(0)
    n_cos = total - n_warm
    phase = math.pi * np.arange(n_cos) / max(n_cos, 1)
    tail = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(phase))
    schedule = np.concatenate((ramp, tail))
    assert len(schedule) == total
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
    batch i+1 is copied from (pinned) host memory on a side stream while batch i is being trained on.  The copies land in
    two persistent device buffer sets used alternately (allocating a fresh 155 MB tensor per step on the side stream makes
    the caching allocator cudaMalloc / cudaFree every step, which costs more than the copy it was meant to hide)."""

    def __init__(self, loader, device: torch.device):
        self.loader = loader
        self.device = device
        self.stream = torch.cuda.Stream(device)
        self._bufs = [None, None]

    def __len__(self):
        return len(self.loader)

    def _fetch(self, it, slot: int):
        try:
            batch = next(it)
        except StopIteration:
            return None
        cur = torch.cuda.current_stream(self.device)
        bufs = self._bufs[slot]
        if bufs is None or len(bufs) != len(batch) or any(
                torch.is_tensor(t) and (b is None or b.shape != t.shape or b.dtype != t.dtype) for t, b in zip(batch, bufs)):
            bufs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) if torch.is_tensor(t) else None for t in batch]
            self._bufs[slot] = bufs
        self.stream.wait_stream(cur)   # the step that last read this slot (two batches ago) has been enqueued: wait for it
        with torch.cuda.stream(self.stream):
            out = []
            for t, b in zip(batch, bufs):
                if torch.is_tensor(t):
                    b.copy_(t, non_blocking=True)
                    out.append(b)
                else:
                    out.append(t)
        return tuple(out)

    def __iter__(self):
        it = iter(self.loader)
        slot = 0
        nxt = self._fetch(it, slot)
        while nxt is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)   # batch i has landed
            batch = nxt
            slot ^= 1
            nxt = self._fetch(it, slot)   # batch i+1 starts copying now, overlapping the step on batch i
            yield batch


# ------------------------------------------------------------------------------------------------
# checkpoint save / resume with the reference's on-disk format (/root/reference/utils/__init__.py:686-770):
# {'model', 'optimizer', 'epoch', 'scaler', 'args'[, 'model_ema']} in <output_dir>/checkpoint-<epoch>.pth
# ------------------------------------------------------------------------------------------------
def save_on_master(*args, **kwargs):
    if is_main_process():
        torch.save(*args, **kwargs)


def ema_state_dict(model: torch.nn.Module, optimizer) -> Optional[dict]:
    """EMA weights kept by ``FusedAdamW.enable_ema`` (one flat fp32 buffer updated inside the AdamW launch) as a
    state_dict with the model's own keys, i.e. what ``timm.utils.get_state_dict(model_ema)`` returns in the reference."""
    ema = getattr(optimizer, "ema", None)
    if ema is None or getattr(optimizer, "_plan", None) is None:
        return None
    st = optimizer._plan[0]["store"]
    out = {}
    for name, p in model.named_parameters():
        if id(p) in st.offsets:
            o, n = st.offsets[id(p)]
            out[name] = ema[o:o + n].view(p.shape).detach().cpu().clone()
    for name, b in model.named_buffers():
        out[name] = b.detach().cpu().clone()
    return out


def load_ema_state_dict(model: torch.nn.Module, optimizer, state: dict) -> None:
    ema = getattr(optimizer, "ema", None)
    if ema is None:
        raise RuntimeError("call optimizer.enable_ema(decay) before loading EMA weights")
    st = optimizer._plan[0]["store"]
    for name, p in model.named_parameters():
        if name in state and id(p) in st.offsets:
            o, n = st.offsets[id(p)]
            ema[o:o + n].copy_(state[name].reshape(-1).to(ema.device, ema.dtype))


def save_model(args, epoch, model, model_without_ddp, optimizer, loss_scaler, model_ema=None):
    """Same call and same file as the reference.  ``model_ema`` may be a timm-style object with ``.ema`` / ``.module``,
    or ``True`` / the optimizer itself to save the fused EMA buffer."""
    from pathlib import Path
    output_dir = Path(args.output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    checkpoint_path = output_dir / ("checkpoint-%s.pth" % str(epoch))
    to_save = {
        "model": {k: v.detach().cpu().clone() for k, v in model_without_ddp.state_dict().items()},
        "optimizer": optimizer.state_dict(),
        "epoch": epoch,
        "scaler": loss_scaler.state_dict() if loss_scaler is not None else None,
        "args": args,
    }
    if model_ema is not None and model_ema is not False:
        inner = getattr(model_ema, "ema", None)
        if isinstance(inner, torch.nn.Module):
            to_save["model_ema"] = {k: v.detach().cpu().clone() for k, v in inner.state_dict().items()}
        elif isinstance(getattr(model_ema, "module", None), torch.nn.Module):
            to_save["model_ema"] = {k: v.detach().cpu().clone() for k, v in model_ema.module.state_dict().items()}
        else:
            sd = ema_state_dict(model_without_ddp, optimizer)
            if sd is not None:
                to_save["model_ema"] = sd
    save_on_master(to_save, checkpoint_path)
    if is_main_process() and isinstance(epoch, int) and hasattr(args, "save_ckpt_num") and hasattr(args, "save_ckpt_freq"):
        to_del = epoch - args.save_ckpt_num * args.save_ckpt_freq
        old_ckpt = output_dir / ("checkpoint-%s.pth" % to_del)
        if os.path.exists(old_ckpt):
            os.remove(old_ckpt)
    return checkpoint_path


def auto_load_model(args, model, model_without_ddp, optimizer, loss_scaler, model_ema=None):
    """Resume from ``args.resume`` or (``args.auto_resume``) the newest ``checkpoint-<int>.pth`` in ``args.output_dir``;
    sets ``args.start_epoch`` like the reference."""
    import glob
    output_dir = str(args.output_dir)
    if getattr(args, "auto_resume", False) and len(getattr(args, "resume", "") or "") == 0:
        latest = -1
        for ckpt in glob.glob(os.path.join(output_dir, "checkpoint-*.pth")):
            t = ckpt.split("-")[-1].split(".")[0]
            if t.isdigit():
                latest = max(int(t), latest)
        if latest >= 0:
            args.resume = os.path.join(output_dir, "checkpoint-%d.pth" % latest)
    if not getattr(args, "resume", ""):
        return None
    if str(args.resume).startswith("https"):
        raise NotImplementedError("remote checkpoints are not fetched (no network dependency on the training path)")
    checkpoint = torch.load(args.resume, map_location="cpu", weights_only=False)
    model_without_ddp.load_state_dict(checkpoint["model"])
    if next(model_without_ddp.parameters()).is_cuda:
        # the reference's resume order is create model -> create optimizer -> auto_load_model, before any forward: make
        # sure the flat parameter store exists so that the optimizer state lands in its flat moment buffers directly
        from .parallel import DataParallel
        from .store import get_store

        get_store(DataParallel._find_root(model_without_ddp))
    if "optimizer" in checkpoint and "epoch" in checkpoint:
        optimizer.load_state_dict(checkpoint["optimizer"])
        if not isinstance(checkpoint["epoch"], str):
            args.start_epoch = checkpoint["epoch"] + 1
        else:
            assert getattr(args, "eval", False), "Does not support resuming with checkpoint-best"
        if model_ema is not None and model_ema is not False:
            ema_sd = checkpoint.get("model_ema", checkpoint["model"])
            inner = getattr(model_ema, "ema", None)
            if isinstance(inner, torch.nn.Module):
                inner.load_state_dict(ema_sd)
            elif getattr(optimizer, "ema", None) is not None:
                load_ema_state_dict(model_without_ddp, optimizer, ema_sd)
        if "scaler" in checkpoint and loss_scaler is not None and checkpoint["scaler"] is not None:
            loss_scaler.load_state_dict(checkpoint["scaler"])
    return checkpoint


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """timm.utils.accuracy: top-k accuracy in percent."""
    maxk = min(max(topk), output.size(1))
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    return [correct[: min(k, maxk)].reshape(-1).float().sum(0) * 100.0 / batch_size for k in topk]

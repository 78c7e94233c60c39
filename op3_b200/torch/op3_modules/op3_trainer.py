"""Mirror of op3/torch/op3_modules/op3_trainer.py: TrainingScheduler (:19-111) and OP3Trainer (:116-270).

The step body (:166-197: zero_grad -> forward -> .mean() -> backward -> clip_grad_norm_(5.0) -> Adam -> 4x .item())
keeps its order; clip + Adam run as one fused kernel over the flat parameter buffer (op3_adam_clip), and when
torch.distributed is initialised (one process per GPU) the flat gradient is all-reduced over NCCL first, which
reproduces nn.DataParallel's mean over equal shards (train_op3.py:99-100)."""
import ctypes as C
import time
from collections import OrderedDict, defaultdict
from os import path as osp

import numpy as np
import torch

from op3_b200 import _lib
from op3_b200._lib import check, ptr
from op3_b200.torch import pytorch_util as ptu


class TrainingScheduler:
    """Schedules of 0 (refine) / 1 (dynamics) / 2 (next-step refine) codes; same types and curriculum constants."""

    def __init__(self, seed_steps, schedule_type, max_T, T):
        self.seed_steps, self.schedule_type, self.max_T, self.T = seed_steps, schedule_type, max_T, T
        self.curriculum_offset = 100
        self.curriculum_increase = 10
        self.rprp_curriculum_len = 50

    def get_schedule(self, epoch, is_train):
        kind, T, seed = self.schedule_type, self.T, self.seed_steps
        if kind in ("single_step_physics", "multi_step_physics"):
            schedule = np.ones((T,))
            schedule[:seed] = 0
        elif kind == "random_alternating":
            schedule = np.random.randint(0, 2, (T,)) if is_train else np.ones((T,))
            schedule[:seed] = 0
        elif kind == "curriculum":
            longest = (epoch - self.curriculum_offset) // self.curriculum_increase + 1 if epoch > self.curriculum_offset else 1
            longest = min(longest, T)
            rollout = np.random.randint(longest) + 1 if is_train else longest     # at least one physics step
            schedule = np.zeros(seed + rollout + 1)
            schedule[seed:seed + rollout] = 1                                      # e.g. [0,0,0,0,1,1,1,0]
        elif kind == "rprp_curriculum":
            base = epoch // self.rprp_curriculum_len + 1
            rollout = np.random.randint(base) + 1 if is_train else base
            schedule = np.zeros(seed + 2 * (rollout + 1))
            schedule[seed:] = 1
            schedule[seed - 1::rollout + 1] = 0
        elif kind == "static_iodine":
            schedule = np.zeros((T,))
        elif kind == "rprp":
            schedule = np.zeros(seed + (T - 1) * 2)
            schedule[seed::2] = 1
        elif kind == "next_step":
            return np.ones(T) * 2
        else:
            raise ValueError("{}".format(self.schedule_type))
        if self.max_T is not None:              # at most max_T - 1 physics steps
            schedule = np.where(np.cumsum(schedule) <= self.max_T - 1, schedule, 0)
        return schedule

    def get_loss_schedule(self, schedule):
        return np.arange(1, len(schedule) + 1)

    def should_save_model(self, epoch):
        if self.schedule_type == "curriculum":
            k = (epoch - self.curriculum_offset) // self.curriculum_increase
            if epoch >= self.curriculum_offset and (epoch - self.curriculum_offset) % self.curriculum_increase == 0 and k <= self.T:
                return "{}_physics_steps".format(k)
        elif epoch % 50 == 49:
            return "{}_epoch_save".format(epoch)
        return None


class FusedClipAdam:
    """clip_grad_norm_(max_norm) + Adam(lr, defaults) over the engine's flat parameter buffer in two launches."""

    def __init__(self, model, lr, max_norm=5.0, betas=(0.9, 0.999), eps=1e-8):
        self.model, self.lr, self.max_norm, self.betas, self.eps = model, lr, max_norm, betas, eps
        self.step_count = 0
        self.exp_avg = self.exp_avg_sq = self.scratch = None
        self.last_grad_norm = None

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None

    def _flat_grad(self):
        eng = self.model.engine
        flat = eng.flat_params()
        g = eng.last_flat_grad
        params = list(self.model.parameters())
        if g is not None and all(p.grad is not None and p.grad.data_ptr() == g.data_ptr() + 4 * o
                                 for p, o in zip(params, eng._offsets)):
            return flat, g
        g = torch.zeros_like(flat)              # gradients came from elsewhere (e.g. accumulated twice): gather them
        for p, o in zip(params, eng._offsets):
            if p.grad is not None:
                g[o:o + p.numel()].copy_(p.grad.reshape(-1))
        return flat, g

    def step(self, grad_scale=1.0, flat_grad=None):
        """`flat_grad`: an already gathered (e.g. all-reduced) flat gradient buffer; default: gather from p.grad."""
        lib = _lib.load()
        if flat_grad is None:
            flat, g = self._flat_grad()
        else:
            flat, g = self.model.engine.flat_params(), flat_grad
            if g.numel() != flat.numel() or g.device != flat.device:
                raise ValueError("flat_grad must match the flat parameter buffer")
        if self.exp_avg is None or self.exp_avg.data_ptr() == 0 or self.exp_avg.numel() != flat.numel():
            self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
            self.scratch = torch.zeros(2048, device=flat.device)
        self.step_count += 1
        st = torch.cuda.current_stream(flat.device).cuda_stream
        check(lib.op3_adam_clip(flat.numel(), ptr(flat), ptr(g), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.step_count,
                                self.lr, self.betas[0], self.betas[1], self.eps, self.max_norm, grad_scale,
                                ptr(self.scratch), st), "op3_adam_clip")
        self.model.engine.manual_version += 1        # derived weights must be refreshed
        self.last_grad_norm = self.scratch[0]

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr}

    def load_state_dict(self, sd):
        self.step_count, self.exp_avg, self.exp_avg_sq, self.lr = sd["step"], sd["exp_avg"], sd["exp_avg_sq"], sd["lr"]
        if self.exp_avg is not None:
            self.scratch = torch.zeros(2048, device=self.exp_avg.device)


def frames_to_float(frames_u8):
    """uint8 CUDA frames (..., 3, D, D) or channels-last (..., D, D, 3) -> float32 (..., 3, D, D) in [0, 1]."""
    if frames_u8.device.type != "cuda":
        raise RuntimeError("op3_b200: frames_to_float needs a CUDA tensor (no CPU fallback)")
    x = frames_u8.contiguous()
    channels_last = x.shape[-1] == 3 and x.shape[-3] != 3
    D = x.shape[-2]
    lead = tuple(x.shape[:-3])
    n = int(np.prod(lead)) if lead else 1
    out = torch.empty(lead + (3, D, D), device=x.device, dtype=torch.float32)
    st = torch.cuda.current_stream(x.device).cuda_stream
    check(_lib.load().op3_prepare_frames(n, D, 1 if channels_last else 0, ptr(x), ptr(out), st), "op3_prepare_frames")
    return out


class ScalarReadback:
    """Per-step scalars on the host without stalling the step loop.

    The reference's train_epoch reads the loss with .item() after every batch (op3_trainer.py:203-213): the GPU then idles
    while the host prepares and launches the next step.  Here every step's scalars are copied into a pinned ring
    (non_blocking, one event per row) and the host picks a row up when its event has completed -- at the latest `lag` steps
    later, i.e. while the next step is already running.  lag = 0 is the reference's blocking read."""

    def __init__(self, device, n=4, lag=1, slots=8):
        self.ring = torch.empty(max(slots, lag + 2), n, dtype=torch.float32).pin_memory()
        self.events = [torch.cuda.Event() for _ in range(self.ring.shape[0])]
        self.device, self.lag, self.pending, self.k = device, int(lag), [], 0

    def push(self, scalars):
        """Queue one step's scalars (device tensors); returns the rows (lists of floats) that became readable, in step order."""
        slot = self.k % self.ring.shape[0]
        self.k += 1
        self.ring[slot].copy_(torch.stack([s.detach().reshape(()) for s in scalars]), non_blocking=True)
        self.events[slot].record(torch.cuda.current_stream(self.device))
        self.pending.append(slot)
        out = []
        while len(self.pending) > self.lag:
            out.append(self._read(self.pending.pop(0)))
        return out

    def drain(self):
        out = [self._read(s) for s in self.pending]
        self.pending = []
        return out

    def _read(self, slot):
        self.events[slot].synchronize()
        return self.ring[slot].tolist()


class OP3Trainer:
    def __init__(self, train_dataset, test_dataset, model, scheduler_class, batch_size, lr=1e-3):
        self.batch_size = batch_size
        if ptu.device is not None:
            model.to(ptu.device)
        self.model = model
        self.scheduler_class = scheduler_class
        self.lr = lr
        self.optimizer = FusedClipAdam(self._core(), lr=lr, max_norm=5.0)
        self.train_dataset, self.test_dataset = train_dataset, test_dataset
        self.vae_logger_stats_for_rl = {}
        self._extra_stats_to_log = None
        self.timing_info = defaultdict(list)
        self.use_cuda_graph = False         # train_step replays one captured CUDA graph per (shapes, schedule) when set
        self._graphs = {}

    def _core(self):
        return self.model.module if hasattr(self.model, "module") else self.model

    def save_model(self, prefix="", directory=None):
        """Same file name / content as op3_trainer.py:145-146 (`<prefix>_params.pkl`, a state_dict).  Rank 0 only under
        torch.distributed.  Keys carry `module.` like the DataParallel-saved reference checkpoints."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
            return
        if directory is None:
            from op3.core import logger          # the reference logger owns the snapshot dir (op3/core, reused as is)
            directory = logger.get_snapshot_dir()
        sd = OrderedDict(("module." + k, v.detach().cpu().clone()) for k, v in self._core().state_dict().items())
        torch.save(sd, open(osp.join(directory, "{}_params.pkl".format(prefix)), "wb"))

    def prepare_tensors(self, tensors):
        """op3_trainer.py:148-153.  Float frames take the reference's `/ 255.0`; uint8 frames (host pinned or device
        resident, (B,T,3,D,D) or channels-last (B,T,D,D,3)) are moved as bytes and converted by op3_prepare_frames."""
        dev = ptu.device if ptu.device is not None else next(self._core().parameters()).device
        frames = tensors[0]
        if frames.dtype == torch.uint8:
            imgs = frames_to_float(frames.to(dev, non_blocking=True))
        else:
            imgs = frames.to(dev) / 255.0
        if len(tensors) == 2 and tensors[1] is not None:
            return imgs, tensors[1].to(dev, non_blocking=True)
        return imgs, None

    # ------------------------------------------------------------------ CUDA-graph step
    def _graph_entry(self, true_images, actions, schedule, loss_schedule):
        """The captured forward+backward of one (shapes, schedule, precision) combination, or None while it is still
        warming up eagerly (lazy one-time initialisations must not happen inside a capture)."""
        import op3_b200
        core = self._core()
        if not (isinstance(true_images, torch.Tensor) and true_images.is_cuda and true_images.dtype == torch.float32):
            return None
        key = (tuple(true_images.shape), None if actions is None else (tuple(actions.shape), actions.dtype),
               tuple(float(x) for x in schedule), tuple(float(x) for x in loss_schedule), true_images.device.index,
               op3_b200.get_precision(), core.K, float(core.beta))
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._graphs[key] = {"eager": 0}
        if "graph" in ent:
            return ent
        if ent["eager"] < 2:
            ent["eager"] += 1
            return None
        eng = core.engine
        dev = true_images.device
        ent["images"] = true_images.clone()
        ent["actions"] = None if actions is None else actions.clone()
        sched, lw = np.asarray(schedule), np.asarray(loss_schedule)
        torch.cuda.synchronize(dev)
        eng._prep_key = None                      # op3_prepare must be part of the graph: the weights change every step
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(dev), torch.cuda.graph(graph):
            out, tape = eng.forward(ent["images"], ent["actions"], None, sched, lw, core._noise, True)
            gflat = eng.backward(tape, 1.0, 0.0, 0.0)
            ent["scalars"] = torch.stack([out["total_loss"], out["clog_prob"], out["kle_loss"], out["mse"]])
        del tape, out
        ent["graph"], ent["gflat"] = graph, gflat
        ent["grad_views"] = eng.grad_views(gflat)
        return ent

    def _graph_step(self, ent, true_images, actions):
        import torch.distributed as dist
        core = self._core()
        ent["images"].copy_(true_images)
        if ent["actions"] is not None:
            ent["actions"].copy_(actions)
        ent["graph"].replay()
        core.engine.last_flat_grad = ent["gflat"]
        for p, v in zip(core.parameters(), ent["grad_views"]):
            p.grad = v
        scale, g = 1.0, None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            g = ent["gflat"]
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            scale = 1.0 / dist.get_world_size()
        self.optimizer.step(grad_scale=scale, flat_grad=g if g is not None else ent["gflat"])
        vals = ent["scalars"].clone()             # the graph's own buffer is overwritten by the next replay
        return vals[0], vals[1], vals[2], vals[3]

    stat_lag = 1     # train_epoch reads a step's scalars this many steps later (0 = blocking read every step, like the reference)

    def train_step(self, true_images, actions, schedule, loss_schedule):
        """One optimisation step (op3_trainer.py:166-197).  Returns the four logged scalars as device tensors.

        With `use_cuda_graph` the forward + backward of a recurring (batch shape, schedule) combination is captured once
        (after two eager steps) and replayed: ~1000 kernel launches become one graph launch.  Gradient all-reduce and
        the fused clip+Adam stay outside the graph (Adam's bias correction takes the step count from the host)."""
        import torch.distributed as dist
        if self.use_cuda_graph and torch.is_grad_enabled():
            ent = self._graph_entry(true_images, actions, schedule, loss_schedule)
            if ent is not None:
                return self._graph_step(ent, true_images, actions)
        self.optimizer.zero_grad()
        _, _, _, total_loss, total_kle_loss, total_clog_prob, mse, _ = self.model.forward(
            true_images, actions, initial_hidden_state=None, schedule=schedule, loss_schedule=loss_schedule)
        total_loss, total_clog_prob, total_kle_loss, mse = total_loss.mean(), total_clog_prob.mean(), total_kle_loss.mean(), mse.mean()
        total_loss.backward()
        scale, g = 1.0, None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            _, g = self.optimizer._flat_grad()                # gathered ONCE: the reduced buffer is what step() consumes
            dist.all_reduce(g, op=dist.ReduceOp.SUM)          # one flat 3.5-3.8 MB buffer, NCCL over NVLink
            scale = 1.0 / dist.get_world_size()
        self.optimizer.step(grad_scale=scale, flat_grad=g)
        return total_loss.detach(), total_clog_prob.detach(), total_kle_loss.detach(), mse.detach()

    def train_epoch(self, epoch):
        t_start = time.time()
        name = self.scheduler_class.should_save_model(epoch)
        if name:
            self.save_model(name)
        self.model.train()
        losses, log_probs, kles, mses, timings = [], [], [], [], []
        rb = None

        def log(rows):
            for vals in rows:
                losses.append(vals[0]); log_probs.append(vals[1]); kles.append(vals[2]); mses.append(vals[3])
        for batch_idx, tensors in enumerate(self.train_dataset.dataloader):
            true_images, actions = self.prepare_tensors(tensors)
            schedule = self.scheduler_class.get_schedule(epoch, is_train=True)
            loss_schedule = self.scheduler_class.get_loss_schedule(schedule)
            t0 = time.time()
            scalars = self.train_step(true_images, actions, schedule, loss_schedule)
            if rb is None:
                rb = ScalarReadback(scalars[0].device, 4, self.stat_lag)
            log(rb.push(scalars))                  # stat_lag = 1: step i's scalars are read while step i + 1 runs
            timings.append(time.time() - t0)
        if rb is not None:
            log(rb.drain())
        return OrderedDict([
            ("train/epoch", epoch), ("train/Log Prob", np.mean(log_probs)), ("train/KL", np.mean(kles)),
            ("train/loss", np.mean(losses)), ("train/mse", np.mean(mses)), ("train/time_step", float(np.sum(timings))),
            ("train/time_total", time.time() - t_start)])

    def test_epoch(self, epoch, save_reconstruction=True, train=True, batches=1):
        """op3_trainer.py:220-271 (uses train-mode schedules like the reference, trap T10).  No image dumps."""
        schedule = self.scheduler_class.get_schedule(epoch, is_train=True)
        loss_schedule = self.scheduler_class.get_loss_schedule(schedule)
        self.model.eval()
        losses, log_probs, kles, mses = [], [], [], []
        loader = self.train_dataset.dataloader if train else self.test_dataset.dataloader
        for batch_idx, tensors in enumerate(loader):
            true_images, actions = self.prepare_tensors(tensors)
            with torch.no_grad():
                _, _, _, total_loss, kle, clog, mse, _ = self.model.forward(true_images, actions, initial_hidden_state=None,
                                                                          schedule=schedule, loss_schedule=loss_schedule)
            vals = torch.stack([total_loss.mean(), clog.mean(), kle.mean(), mse.mean()]).tolist()
            losses.append(vals[0]); log_probs.append(vals[1]); kles.append(vals[2]); mses.append(vals[3])
            if batch_idx >= batches - 1:
                break
        return OrderedDict([("test/Log Prob", np.mean(log_probs)), ("test/KL", np.mean(kles)),
                            ("test/loss", np.mean(losses)), ("test/mse", np.mean(mses))])

"""Reference-facing training routine: the drop-in for ``fullbatch.training.train`` (reference
fullbatch/training/training.py:50-340) restricted to the hot path -- the full-batch closure with the per-microbatch
finite-difference regulariser -- plus the SGD sanity branch (:241-286) through the same kernels.

Structure mirrors the reference:

    while step < cfg.hyp.steps:
        def closure():                                   # training.py:226-234
            loss = _accumulate_full_gradient()           # :121-185 -> engine (CUDA graph per microbatch, one all-reduce)
            _modify_gradient_params()                    # :187-215 (global-norm clip)
            return loss
        optimizer.step(closure); scheduler.step()        # :237-238 (stock torch.optim.SGD, as in the reference)

What is different on purpose: microbatch gradients never leave the device, the running mean lives in one flat fp32
buffer whose views are ``param.grad`` (the reference does the same in its distributed branch,
training/utils.py:36-41), and with several processes each rank owns a contiguous range of the single-process microbatch
list and the result is the exact global mean (the reference's own multi-process weighting is not a mean).
"""
import logging
import math
import os
import time
from collections import defaultdict

import torch

from .engine import FullBatchEngine
from .modules import GradRegularizer, LabelSmoothCrossEntropyLoss
from .optim import FlatSGD, S_GNORM, S_PNORM
from .schedulers import build_scheduler

log = logging.getLogger("fullbatch_b200")


def get_loss_fn(cfg_hyp, batch_size=None):
    """training.py:391-413: label_smoothing not in [None, ""] selects LabelSmoothCrossEntropyLoss (also for 0.0)."""
    if cfg_hyp.loss_modification is not None:
        raise ValueError(f"Loss modification {cfg_hyp.loss_modification} is not on the B200 path.")
    smoothing = cfg_hyp.label_smoothing if cfg_hyp.label_smoothing not in [None, ""] else 0.0
    return LabelSmoothCrossEntropyLoss(smoothing=smoothing)


def optim_interface(model, cfg_hyp, fused=False):
    """optimizers.py:10-93 for the configuration the path uses: ``Gradient Descent`` = torch.optim.SGD
    (:25-28) and the reference's scheduler objects (schedulers.build_scheduler: stock torch schedulers behind the linear
    warm-up wrapper), so ``scheduler.state_dict()`` in a checkpoint is interchangeable with the reference's."""
    if cfg_hyp.optim.name != "Gradient Descent" or cfg_hyp.optim.get("line_search", "none") != "none":
        raise ValueError(f"Optimizer {cfg_hyp.optim.name} / line search is not on the B200 path (closure optimizers "
                         "from the reference can be passed to train_with_optimizer).")
    if cfg_hyp.optim_modification.name != "none":
        raise ValueError("optim_modification is not on the B200 path")
    params = {k: v for k, v in cfg_hyp.optim.items() if k not in ("name", "line_search")}
    if fused:
        optimizer = FlatSGD(model.parameters(), **params)  # clip + SGD + param norm in one device sweep
    else:
        optimizer = torch.optim.SGD(model.parameters(), **params)
    return optimizer, build_scheduler(optimizer, cfg_hyp)


@torch.no_grad()
def save_checkpoint(model, optimizer, scheduler, step, file):
    """training/utils.py:43-51: the reference's 5-list [optim_state, model_state, scheduler_state, scaler_state, step]
    (scaler_state is None: no fp16 grad scaling on this path), loadable by the reference's `_load_from_checkpoint` and
    `hubconf.py:37-40`: state_dict keys, the optimizer state layout (torch SGD) and the scheduler state (torch schedulers
    / the warm-up wrapper's attribute names) are those of the reference."""
    model_state = {k: v.detach().clone() for k, v in model.state_dict().items()}  # params are views of the flat buffer
    os.makedirs(os.path.dirname(os.path.abspath(file)), exist_ok=True)
    torch.save([optimizer.state_dict(), model_state, scheduler.state_dict(), None, step], file)


@torch.no_grad()
def load_checkpoint(model, optimizer, scheduler, max_steps, device=None, file="checkpoints/fb.pth"):
    """training/utils.py:53-70: returns the step to continue from (0 if no checkpoint exists); raises ValueError if the
    checkpoint already reached max_steps."""
    try:
        optim_state, model_state, scheduler_state, _, step = torch.load(file, map_location=device)
    except FileNotFoundError:
        log.info("No existing checkpoint found. Starting to train from step 0.")
        return 0
    model.load_state_dict(model_state)  # copies into the existing (flat-buffer backed) parameters
    optimizer.load_state_dict(optim_state)
    scheduler.load_state_dict(scheduler_state)
    if step >= max_steps:
        raise ValueError("Maximum step size reached. Terminating computations.")
    log.info(f"Existing checkpoint loaded successfully. Continuing to train from step {step}.")
    return step


def shard_range(rank, world, num_microbatches):
    """Contiguous range [k0, k1) of the single-process microbatch list owned by `rank` (sizes differ by at most 1)."""
    return (rank * num_microbatches) // world, ((rank + 1) * num_microbatches) // world


def _resident_dataset(loader, device):
    """If the loader iterates a TensorDataset sequentially, keep the whole dataset in HBM (50k CIFAR images = 614 MB)."""
    ds = getattr(loader, "dataset", None)
    sampler = getattr(loader, "sampler", None)
    ok_sampler = isinstance(sampler, (torch.utils.data.SequentialSampler, torch.utils.data.RandomSampler))
    if isinstance(ds, torch.utils.data.TensorDataset) and ok_sampler:
        X, Y = ds.tensors
        if X.dtype == torch.uint8:  # raw HWC images: normalisation / augmentation happen inside the stem kernel
            return X.to(device=device).contiguous(), Y.to(device=device, dtype=torch.long).contiguous()
        return X.to(device=device, dtype=torch.float32).contiguous(), Y.to(device=device, dtype=torch.long).contiguous()
    return None


S_PNORM_PREV = 9  # scal slot: sum theta^2 at gradient-evaluation time (copied from S_PNORM before the update)


class Trainer:
    """State of one ``train`` call; ``step()`` is one iteration of the reference's main loop (training.py:219-339)."""

    def __init__(self, model, trainloader, validloader, setup, cfg):
        model.train()
        self.model, self.trainloader, self.validloader, self.setup, self.cfg = model, trainloader, validloader, setup, cfg
        hyp = cfg.hyp
        self.stochastic = bool(hyp.train_stochastic)
        # impl.fused_optimizer (default on): FlatSGD = clip + SGD + param norm as one sweep over the flat buffers;
        # off: stock torch.optim.SGD and the torch clip of training.py:198-211
        self.fused_opt = bool(cfg.impl.get("fused_optimizer", True)) and not self.stochastic and \
            (hyp.grad_clip is None or float(hyp.grad_clip_norm) == 2.0)
        self.optimizer, self.scheduler = optim_interface(model, hyp, fused=self.fused_opt)
        self.stats = defaultdict(list)
        self.device = torch.device(setup["device"])
        if str(cfg.impl.accumulation_dtype) not in ("float", "float32") or \
                setup.get("dtype", torch.float32) != torch.float32:
            raise ValueError("the B200 path keeps fp32 master weights and fp32 accumulation "
                             "(impl.dtype / impl.accumulation_dtype must be float)")
        if cfg.impl.mixed_precision:
            raise ValueError("impl.mixed_precision is not on the B200 path; use impl.precision = split | bf16")
        # reference options that are not on the B200 path are refused, never silently ignored
        if hyp.norm_bias.strength > 0:
            raise ValueError("hyp.norm_bias is not on the B200 path")
        if hyp.get("evaluate_ema", False):
            raise ValueError("hyp.evaluate_ema is not on the B200 path")
        noise = hyp.get("grad_noise", None) or {}
        if noise.get("additive") is not None or noise.get("multiplicative") is not None:
            raise ValueError("hyp.grad_noise is not on the B200 path")
        if hyp.get("only_linear_layers_weight_decay", False):
            raise ValueError("hyp.only_linear_layers_weight_decay is not on the B200 path")
        if hyp.get("train_switch_stochastic", None) is not None or hyp.get("train_semi_stochastic", False):
            raise ValueError("hyp.train_switch_stochastic / train_semi_stochastic are not on the B200 path")
        if hyp.batch_clip is not None and float(hyp.grad_clip_norm) != 2.0:
            raise ValueError("batch_clip is implemented for the 2-norm only")
        # the stochastic sanity branch differentiates whole loader blocks (training.py:252-262), the full-batch
        # branch chunks of hyp.sub_batch (training.py:155-156)
        self.mb = cfg.data.batch_size if self.stochastic else min(cfg.data.batch_size, hyp.sub_batch)
        self.num_blocks = len(trainloader)
        self.num_chunks = 1 if self.stochastic else max(cfg.data.batch_size // hyp.sub_batch, 1)
        self.dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.rank = torch.distributed.get_rank() if self.dist else 0
        self.world = torch.distributed.get_world_size() if self.dist else 1
        loss_fn = get_loss_fn(hyp, cfg.data.batch_size)
        self.engine = FullBatchEngine(model, self.mb, precision=cfg.impl.get("precision", "split"),
                                      label_smoothing=loss_fn.smoothing, device=self.device,
                                      groups=1 if self.stochastic else cfg.impl.get("groups", None),
                                      policy_groups=1 if self.stochastic else None,
                                      lanes=1 if self.stochastic else cfg.impl.get("lanes", None))
        self.gradreg = GradRegularizer(model, self.optimizer, loss_fn, **hyp.grad_reg, mixed_precision=False,
                                       engine=self.engine)
        self.bs, self.eps = self.gradreg.block_strength, self.gradreg.eps
        self.acc = float(hyp.grad_reg.acc_strength)
        self.impl = hyp.grad_reg.implementation if (self.bs != 0 or self.acc != 0) else "forward-differences"
        if self.impl == "forward-differences-legacy":
            self.acc = 0.0  # the reference's legacy path disregards pre_grads (modules.py:243-264)
        if self.fused_opt:
            self.optimizer.bind(self.engine, hyp.grad_clip)
        self.resident = _resident_dataset(trainloader, self.device) if cfg.impl.get("resident_dataset", True) and \
            not self.stochastic else None
        # microbatches per full-batch pass (training.py:65-66,146).  A resident dataset is sharded here by contiguous
        # microbatch ranges; a streamed loader is taken as this rank's shard already (impl.setup.sharded_loader).
        local = self.num_blocks * self.num_chunks
        if self.resident is not None:
            n = self.resident[0].shape[0]
            if getattr(trainloader, "batch_size", cfg.data.batch_size) != cfg.data.batch_size:
                raise ValueError(f"the loader's batch_size {trainloader.batch_size} disagrees with "
                                 f"cfg.data.batch_size {cfg.data.batch_size}")
            if local * self.mb > n:  # drop_last=False with a ragged last block
                raise ValueError(f"{local} microbatches of {self.mb} exceed the {n} samples of the dataset: the "
                                 "full-batch path needs drop_last=True (data_preparation.py:68)")
        if self.resident is not None or self.world == 1:
            self.K = local
            self.k0, self.k1 = shard_range(self.rank, self.world, local)
        else:
            if not cfg.impl.setup.get("sharded_loader", False):
                raise RuntimeError("multi-process training needs a TensorDataset loader (sharded on the device) or "
                                   "impl.setup.sharded_loader=True with one loader per rank")
            counts = torch.zeros(self.world, device=self.device, dtype=torch.int64)
            counts[self.rank] = local
            torch.distributed.all_reduce(counts)
            self.K = int(counts.sum())
            self.k0 = int(counts[:self.rank].sum())
            self.k1 = self.k0 + local
        self.step_count = 0
        if cfg.impl.checkpoint.name is not None:  # training.py:60-63
            self.checkpoint_file = os.path.join(cfg.original_cwd, "checkpoints", cfg.impl.checkpoint.name)
            self.step_count = load_checkpoint(model, self.optimizer, self.scheduler, hyp.steps, device=self.device,
                                              file=self.checkpoint_file)
        else:
            self.checkpoint_file = None
        # sum theta^2 for _record_stats stays on the device: initialised here (also right after a resume), refreshed by
        # every FlatSGD step
        self.engine.scal[S_PNORM] = self.engine.theta.double().pow(2).sum().float()
        # device-side data pipeline: raw uint8 dataset + crop/flip/normalise fused into the stem,
        # hyp.shuffle as a per-step device permutation (data_preparation.py:53-54)
        self.data_gen = torch.Generator(device=self.device)
        self.data_gen.manual_seed(int(cfg.seed) if cfg.get("seed") is not None else 0)
        self.perm = None
        self.augment = False
        if self.resident is not None and self.resident[0].dtype == torch.uint8:
            data = cfg.data
            if data.get("normalize", True):
                self.engine.set_normalization(data.mean, data.std)
            aug = data.get("augmentations_train") or {}
            unknown = set(aug.keys()) - {"RandomCrop", "RandomHorizontalFlip"}
            if unknown:
                raise ValueError(f"augmentations {sorted(unknown)} are not on the B200 path")
            self.augment = len(aug) > 0
            self.crop_pad = int(aug["RandomCrop"][1]) if "RandomCrop" in aug else 0
            self.flip_p = float(aug.get("RandomHorizontalFlip", 0.0))

    def _prepare_epoch(self):
        if self.resident is None:
            return
        n = self.resident[0].shape[0]
        if self.cfg.hyp.shuffle:
            if self.perm is None:
                self.perm = torch.zeros(n, device=self.device, dtype=torch.int64)
            self.perm.copy_(torch.randperm(n, device=self.device, generator=self.data_gen))
        if self.augment:
            self.engine.draw_augmentation(n, self.data_gen, self.crop_pad, self.flip_p)

    def _reduce_pre(self, pre, local):
        """training.py:139-140 across processes: every rank holds the mean raw gradient over ITS microbatches; the
        weighted all-reduce makes it the mean over all of them (exactly the single-process pre_grads)."""
        if self.dist:
            self.engine.all_reduce_flat(pre, local, self.K)

    # training.py:121-185
    def _accumulate_full_gradient(self):
        """Enqueues the whole full-batch gradient evaluation (no host synchronisation) and returns the mean loss as a
        device scalar; the statistics are read by _record_stats after the optimizer step."""
        self._t0 = time.time()
        eng, cfg = self.engine, self.cfg
        self._lr = self.optimizer.param_groups[0]["lr"]
        if not self.fused_opt:  # stock optimizer: nobody refreshes sum theta^2 on the device
            eng.scal[S_PNORM] = eng.theta.double().pow(2).sum().float()
        self._prepare_epoch()
        if self.resident is not None:
            local = self.k1 - self.k0
            eng.accumulate_resident(self.resident[0], self.resident[1], self._lr, self.bs, self.eps,
                                    first=self.k0 * self.mb, count=local, num_norms=self.K,
                                    norm_offset=self.k0, implementation=self.impl, acc_strength=self.acc,
                                    batch_clip=cfg.hyp.batch_clip, perm=self.perm,
                                    reduce_pre=lambda pre: self._reduce_pre(pre, local))
        else:
            if self.acc != 0 or cfg.hyp.batch_clip is not None or self.impl == "central-differences":
                raise RuntimeError("acc_strength / batch_clip / central-differences need a device-resident dataset "
                                   "(TensorDataset loader)")
            local = eng.accumulate_stream(self.trainloader, self._lr, self.bs, self.eps, self.K, norm_offset=self.k0)
        if self.dist:
            eng.all_reduce_mean(local, self.K)
        for p, g in zip(self.model.parameters(), eng.grads_list(eng.avg)):
            p.grad = g  # training.py:183 / training/utils.py:40
        eng.scal[S_PNORM_PREV] = eng.scal[S_PNORM]  # sum theta^2 of the parameters the gradient was evaluated at
        return eng.scal[0] / self.K

    def _record_stats(self):
        """training.py:85-119 from ONE host read of the device scalars (the reference syncs several times)."""
        eng, cfg, stats = self.engine, self.cfg, self.stats
        res = eng.results(self.K)
        lr = self._lr
        for idx, entry in enumerate(res["grad_norms"].sqrt().tolist()):
            stats[f"grad_norm_train_{idx}"] += [entry]
        param_norm = res["scal"][S_PNORM_PREV]
        full_grad_norm = float(res["grad_norms"].mean())
        full_loss = res["loss"] + 0.5 * cfg.hyp.optim.get("weight_decay", 0.0) * param_norm
        if self.bs != 0:
            full_loss += lr / 4 * self.bs * full_grad_norm
        if self.acc != 0:  # training.py:99-102
            full_loss += lr / 4 * self.acc * float(eng.pre.double().pow(2).sum())
        if cfg.hyp.batch_clip is not None:  # training.py:117-119 (a NameError in the reference as shipped)
            stats["clipped_batches"] += [res["clipped_batches"]]
        stats["train_loss"] += [res["loss"]]
        stats["train_acc"] += [res["correct"] / (self.K * self.mb)]
        stats["train_time"] += [time.time() - self._t0]
        stats["param_norm"] += [param_norm]
        stats["grad_norm"] += [math.sqrt(full_grad_norm)]
        stats["full_loss"] += [full_loss]
        if self.fused_opt and cfg.hyp.grad_clip is not None:
            gnorm = math.sqrt(res["scal"][S_GNORM])
            stats["preclip_gradnorm"] += [gnorm]
            stats["clipped_step"] += [1 if gnorm > cfg.hyp.grad_clip else 0]

    @torch.no_grad()
    def _modify_gradient_params(self):
        """training.py:187-215 (global clip only; the fused optimizer does it on the device inside its sweep)."""
        cfg, eng, stats = self.cfg, self.engine, self.stats
        if self.fused_opt:
            return  # FlatSGD.step clips on the device; the statistics are read after the step
        if cfg.hyp.grad_clip is not None:
            norm_type = float(cfg.hyp.grad_clip_norm)
            grad_norm = eng.avg.abs().max() if norm_type == float("inf") else torch.norm(eng.avg, norm_type)
            stats["preclip_gradnorm"] += [grad_norm.item()]
            if grad_norm > cfg.hyp.grad_clip:
                eng.avg.mul_(cfg.hyp.grad_clip / (grad_norm + 1e-6))
                stats["clipped_step"] += [1]
            else:
                stats["clipped_step"] += [0]

    def _sgd_epoch(self):
        """training.py:241-286: one optimizer step per loader block through the same kernels (raw gradient of the whole
        block, regulariser per block).  Loss / accuracy / gradient norms accumulate on the device; one host read per
        epoch.  Across processes the block gradients are AVERAGED: the reference sums them without dividing
        (training.py:268-270 via training/utils.py:35), which SURVEY.md 8e lists as a defect not to reproduce."""
        eng, cfg, stats = self.engine, self.cfg, self.stats
        t0 = time.time()
        acc = torch.zeros(2, device=self.device)  # sum of block losses, sum of correct predictions
        grad_norms = torch.zeros(max(self.num_blocks, 1), device=self.device)
        datapoints = 0
        for block, (inputs, labels) in enumerate(self.trainloader):
            if inputs.shape[0] != self.mb:
                raise RuntimeError(f"block of {inputs.shape[0]} samples, engine built for data.batch_size = {self.mb} "
                                   "(the stochastic branch needs drop_last=True)")
            inputs = inputs.to(device=self.device, dtype=torch.float32, non_blocking=True)
            labels = labels.to(device=self.device, dtype=torch.long, non_blocking=True)
            datapoints += labels.shape[0]

            def closure():
                loss, correct = eng.microbatch_gradient(inputs, labels)
                acc.add_(torch.stack([loss, correct]))
                grad_norms[block] = eng.grad_norms[0]  # training.py:262, squared norm of the raw block gradient
                if self.bs != 0:
                    eng.regularize(inputs, labels, self.optimizer.param_groups[0]["lr"], self.bs, self.eps, self.impl)
                if self.dist:
                    eng.all_reduce_flat(eng.g, 1, self.world)
                for p, g in zip(self.model.parameters(), eng.grads_list(eng.g)):
                    p.grad = g
                if cfg.hyp.grad_clip is not None:  # training.py:271-272
                    torch.nn.utils.clip_grad_norm_(self.model.parameters(), cfg.hyp.grad_clip, norm_type=2.0)
                return loss

            self.optimizer.step(closure)
        loss_sum, preds = acc.tolist()
        blocks = max(self.num_blocks, 1)
        param_norm = float(eng.theta.double().pow(2).sum())
        full_grad_norm = float(grad_norms.mean())
        train_loss = loss_sum / blocks
        full_loss = train_loss + 0.5 * cfg.hyp.optim.get("weight_decay", 0.0) * param_norm
        if self.bs != 0:
            full_loss += self.optimizer.param_groups[0]["lr"] / 4 * self.bs * full_grad_norm
        for idx, entry in enumerate(grad_norms.sqrt().tolist()):
            stats[f"grad_norm_train_{idx}"] += [entry]
        stats["train_loss"] += [train_loss]
        stats["train_acc"] += [preds / max(datapoints, 1)]
        stats["train_time"] += [time.time() - t0]
        stats["param_norm"] += [param_norm]
        stats["grad_norm"] += [math.sqrt(full_grad_norm)]
        stats["full_loss"] += [full_loss]

    def step(self, validate=True):
        cfg = self.cfg
        self.model.train()
        if not self.stochastic:
            def gradient_evaluation():  # training.py:226-234
                loss = self._accumulate_full_gradient()
                self._modify_gradient_params()
                return loss

            self.optimizer.step(gradient_evaluation)
            self._record_stats()
            self.scheduler.step()
        else:
            self._sgd_epoch()
            self.scheduler.step()
        self.step_count += 1
        self.engine.sync_bn_counters()
        step = self.step_count
        # training.py:303-304
        if validate and self.validloader is not None and ((step - 1) % cfg.impl.validate_every_nth_step == 0
                                                          or step >= cfg.hyp.steps or cfg.dryrun):
            evaluate(self.model, self.validloader, self.stats, self.setup, cfg.impl, cfg.hyp, dryrun=cfg.dryrun,
                     engine=self.engine if getattr(cfg.impl, "kernel_evaluate", True) else None)
        if self.rank == 0:
            log.info(status_message(self.optimizer, self.stats, step))
        return self.stats["train_loss"][-1]

    def maybe_checkpoint(self):
        """training.py:330-335 (after the divergence / full-accuracy checks of the main loop)"""
        step, cfg = self.step_count, self.cfg
        if self.rank == 0 and self.checkpoint_file is not None and \
                ((step - 1) % cfg.impl.checkpoint.save_every_nth_step == 0 or step >= cfg.hyp.steps):
            save_checkpoint(self.model, self.optimizer, self.scheduler, step, self.checkpoint_file)


def train(model, trainloader, validloader, setup, cfg):
    """Drop-in for ``fullbatch.training.train`` (training.py:50): trains `model` (built by ``construct_model``) and
    returns ``stats`` with the reference's keys."""
    trainer = Trainer(model, trainloader, validloader, setup, cfg)
    while trainer.step_count < cfg.hyp.steps:
        loss = trainer.step()
        if not math.isfinite(loss):  # training.py:314-317
            log.info("Terminating iterations due to divergence of loss...")
            break
        n_full = cfg.hyp.stop_at_full_training_accuracy  # training.py:319-328
        if n_full > 0 and min(trainer.stats["train_acc"][-n_full:]) == 1:
            log.info("Terminating training after fitting all datapoints.")
            if validloader is not None:
                evaluate(model, validloader, trainer.stats, setup, cfg.impl, cfg.hyp, dryrun=cfg.dryrun,
                         engine=trainer.engine if getattr(cfg.impl, "kernel_evaluate", True) else None)
            break
        trainer.maybe_checkpoint()
        if cfg.dryrun:
            break
    return trainer.stats


def measure_implementation_noise(model, trainloader, validloader, setup, cfg):
    """Drop-in for ``_measure_implementation_noise`` (training.py:429-600, the body of
    measure_floating_point_accuracy.py:22-31): evaluate the full-batch gradient twice from the same state (parameters
    AND BatchNorm buffers restored in between) and report the norms of the gradient and of the difference
    (training.py:586-598).  The reference measures cuDNN / atomics nondeterminism on a GPU (0.0 on CPU); every reduction
    on this path has a fixed order, so the expected drift is exactly 0."""
    trainer = Trainer(model, trainloader, validloader, setup, cfg)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    trainer._accumulate_full_gradient()
    first = trainer.engine.avg.clone()
    model.load_state_dict(state)
    trainer._accumulate_full_gradient()
    second = trainer.engine.avg
    diff = (first - second).double()
    ref = first.double()
    report = dict(grad_linf=float(ref.abs().max()), grad_l2=float(ref.norm()), grad_l1=float(ref.abs().sum()),
                  diff_linf=float(diff.abs().max()), diff_l2=float(diff.norm()), diff_l1=float(diff.abs().sum()))
    log.info("Gradient  : Linf %(grad_linf).4e L2 %(grad_l2).4e L1 %(grad_l1).4e" % report)
    log.info("Difference: Linf %(diff_linf).4e L2 %(diff_l2).4e L1 %(diff_l1).4e" % report)
    return report


@torch.no_grad()
def evaluate(model, dataloader, stats, setup, cfg_impl, cfg_hyp, dryrun=False, engine=None):
    """training.py:343-388: eval-mode forward over the validation loader.  With `engine` (the trainer's
    FullBatchEngine) the forward runs on the CUDA kernels (engine.forward_eval, in chunks of the microbatch size);
    without it through the torch modules, exactly as the reference."""
    loss_fn = torch.nn.CrossEntropyLoss()
    model.eval()
    if engine is not None:
        def forward(x):
            return torch.cat([engine.forward_eval(x[i:i + engine.mb].contiguous())
                              for i in range(0, x.shape[0], engine.mb)])
    else:
        forward = model
    device = torch.device(setup["device"])
    if cfg_impl.setup.dist and torch.distributed.is_available() and torch.distributed.is_initialized():
        # training.py:347-357: ranks see different microbatches, so BN running statistics are averaged first
        bufs = [b for b in model.buffers()]
        if bufs:
            flat = torch.cat([b.data.reshape(-1).float() for b in bufs])
            torch.distributed.all_reduce(flat)
            flat /= cfg_impl.setup.world_size
            off = 0
            for b in bufs:
                b.data.copy_(flat[off:off + b.numel()].view_as(b).to(b.dtype))
                off += b.numel()
    if stats is None:
        stats = defaultdict(list)
    step_loss, step_preds, datapoints = 0.0, 0.0, 0
    for inputs, labels in dataloader:
        inputs = inputs.to(device=device, dtype=torch.float32)
        labels = labels.to(device=device, dtype=torch.long)
        datapoints += labels.shape[0]
        if cfg_hyp.test_time_flips:  # training.py:370-373
            outputs = forward(inputs).softmax(dim=1) + forward(torch.flip(inputs, [3])).softmax(dim=1)
        else:
            outputs = forward(inputs)
        step_loss += loss_fn(outputs, labels).item() * labels.shape[0]
        step_preds += (outputs.argmax(dim=-1) == labels).float().sum().item()
        if dryrun:
            break
    stats["valid_loss"] += [step_loss / max(datapoints, 1)]
    stats["valid_acc"] += [step_preds / max(datapoints, 1)]
    model.train()
    return stats


def status_message(optimizer, stats, step):
    """training.py:416-426."""
    def last(key):
        return stats[key][-1] if len(stats[key]) > 0 else float("nan")

    return (f'Step: {step:<4}| lr: {optimizer.param_groups[0]["lr"]:.4f} | Time: {last("train_time"):4.2f}s |'
            f'TRAIN loss {last("train_loss"):7.4f} | TRAIN Acc: {last("train_acc"):7.2%} |'
            f'VAL loss {last("valid_loss"):7.4f} | VAL Acc: {last("valid_acc"):7.2%} |')

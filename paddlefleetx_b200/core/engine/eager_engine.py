"""``EagerEngine`` — the trainer: fit / evaluate / predict / save / load / export / inference.

Control flow and externally visible behaviour follow ppfleetx/core/engine/eager_engine.py:47-925 (SURVEY §3.1,
§7.4): epoch loop -> step loop (resume skips consumed batches, LR stepped in samples under ``use_increments``,
logging every ``logging_freq`` with device-synchronised timestamps, in-loop eval every ``eval_freq`` but never on
the first step, save every ``save_steps``, step-mode runs steps 0..max_steps inclusive), micro-batch gradient
accumulation, bf16 = static scale 1.0 / fp16 = dynamic scaler, reference checkpoint layout.

B200-first differences underneath:
  * inputs are staged through pinned host buffers and copied on a side stream one step ahead,
  * the loss stays on the device between logging boundaries (no per-step ``.item()``), step time is also taken
    with CUDA events so the log carries device time,
  * gradient traffic (DP all-reduce / ZeRO reduce-scatter + param all-gather) belongs to the flat optimizer,
    the pipeline schedule to ``parallel/pipeline.py``; the engine only sequences them.
"""
from __future__ import annotations

import os

import torch

from ...distributed.apis import amp as amp_api
from ...distributed.apis import env
from ...distributed.apis import io as ckpt_io
from ...distributed.apis.strategy import wrap_with_fleet
from ...optims import build_lr_scheduler, build_optimizer
from ...optims.lr_scheduler import LRScheduler
from ...parallel.rng import get_rng_state_tracker
from ...parallel import comm_ops as C
from ...parallel.tp_layers import allreduce_sequence_parallel_grads
from ...utils.log import get_timestamp, logger
from ..module.basic_module import BasicModule
from .basic_engine import BasicEngine


def _to_device(obj, device, non_blocking=True):
    if isinstance(obj, torch.Tensor):
        return obj.to(device, non_blocking=non_blocking)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(o, device, non_blocking) for o in obj)
    if isinstance(obj, dict):
        return {k: _to_device(v, device, non_blocking) for k, v in obj.items()}
    return obj


def _split_micro(batch, n: int):
    """Split every tensor of a (nested) batch into ``n`` micro-batches along dim 0."""
    if n == 1:
        return [batch]
    if isinstance(batch, torch.Tensor):
        return list(batch.chunk(n, dim=0))
    if isinstance(batch, (list, tuple)):
        parts = [_split_micro(b, n) for b in batch]
        return [type(batch)(p[i] for p in parts) for i in range(n)]
    if isinstance(batch, dict):
        parts = {k: _split_micro(v, n) for k, v in batch.items()}
        return [{k: parts[k][i] for k in batch} for i in range(n)]
    return [batch] * n


def _flatten_tensors(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _flatten_tensors(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _flatten_tensors(v)


class _Prefetcher:
    """Wraps a host dataloader: batch i+1 is copied H2D on a side stream while step i computes."""

    def __init__(self, loader, device: torch.device):
        self.loader, self.device = loader, device
        self.stream = torch.cuda.Stream() if device.type == "cuda" else None

    def __iter__(self):
        # DataLoader.__iter__ draws its worker base seed from the global host generator; fork the generator around it so that creating (or, on
        # resume, re-creating) an iterator never shifts the stream the model's own noise comes from
        with torch.random.fork_rng(devices=[]):
            it = iter(self.loader)
        nxt = self._load(it)
        while nxt is not None:
            if self.stream is not None:
                torch.cuda.current_stream().wait_stream(self.stream)
            cur = nxt
            if self.stream is not None:
                # allocated on the side stream, consumed on the compute stream: tell the caching allocator, otherwise the side stream may
                # recycle the block for a later batch while a backward kernel (embedding scatter) still reads it
                compute = torch.cuda.current_stream()
                for t in _flatten_tensors(cur):
                    if t.is_cuda:
                        t.record_stream(compute)
            nxt = self._load(it)
            yield cur

    def _load(self, it):
        try:
            batch = next(it)
        except StopIteration:
            return None
        if self.stream is None:
            return _to_device(batch, self.device, False)
        with torch.cuda.stream(self.stream):
            return _to_device(batch, self.device, True)

    def __len__(self):
        return len(self.loader)


class EagerEngine(BasicEngine):
    def __init__(self, configs, module: BasicModule, optimizer=None, lr=None, mode: str = "train"):
        super().__init__()
        if not isinstance(module, BasicModule):
            raise TypeError("'module' must be sub classes of `BasicModule`, but got: {}".format(module.__class__.__name__))
        self._configs = configs
        self._module = module
        self.mode = mode
        g, eng, d = configs.Global, configs.Engine, configs.Distributed
        self._device = torch.device("cuda", torch.cuda.current_device()) if str(g.get("device", "gpu")) == "gpu" and torch.cuda.is_available() \
            else torch.device("cpu")
        if module.model is not None and not any(True for _ in module.model.parameters()):
            logger.warning("module.model has no parameters")

        self._run_mode = eng.get("run_mode", "step")
        assert self._run_mode in ("epoch", "step"), "run_mode must be epoch or step"
        self._max_steps = eng.max_steps
        self._eval_freq = eng.eval_freq
        self._eval_iters = eng.eval_iters
        self._test_iters = eng.test_iters
        self._logging_freq = eng.logging_freq
        self._num_train_epochs = eng.num_train_epochs
        self._accumulate_steps = eng.accumulate_steps
        amp = eng.mix_precision
        self._use_pure_fp16 = bool(amp.enable) and mode != "export"
        self._amp_dtype = str(amp.get("dtype", "float16"))
        self._amp_level = str(amp.get("level", "O2"))
        self._scale_loss = amp.scale_loss
        self._save_steps = eng.save_load.save_steps
        self._save_epoch = eng.save_load.save_epoch
        self._output_dir = eng.save_load.output_dir
        self._ckpt_dir = eng.save_load.ckpt_dir
        self._resume_replay = bool(eng.save_load.get("resume_replay", False))
        self._global_batch_size = g.global_batch_size
        self._local_batch_size = g.local_batch_size
        self._micro_batch_size = g.micro_batch_size
        self._dist = d
        self._dp_degree, self._mp_degree, self._pp_degree = d.dp_degree, d.mp_degree, d.pp_degree
        self._sharding_degree = d.sharding.sharding_degree
        self._sharding_stage = d.sharding.sharding_stage
        self._hcg = env.get_hcg()
        self._dp_rank = self._hcg.get_data_parallel_rank()
        if self._sharding_stage in (2, 3) and self._sharding_degree > 1 and self._pp_degree > 1 and self._sharding_stage == 3:
            raise AssertionError("sharding stage 3 is not combined with pipeline parallel")

        # compression hook (prune / QAT) — mirrors eager_engine.py:757-774
        self._quant_mode = False
        if "Compress" in configs and configs.Compress:
            from ...utils.compression_helper import compress_model

            self._module.model, self._quant_mode = compress_model(self._module.model, configs.Compress, self._device)

        # parameter-efficient fine-tuning (PEFT: {method: lora | prefix, ...}): adapters go in before the optimizer is built, so only they
        # (and PEFT.train_modules) get optimizer state
        self._peft = None
        if configs.get("PEFT") and configs.PEFT.get("method"):
            from ...utils.peft import apply_peft

            self._peft = apply_peft(self._module.model, configs.PEFT)
            logger.info("PEFT {method}: {trainable:,} trainable of {total:,} parameters".format(**self._peft))

        if self._device.type == "cuda":
            self._module.to(self._device)

        # scaler
        if self._use_pure_fp16 and mode == "train":
            if self._amp_dtype == "float16":
                self._scaler = amp_api.GradScaler(True, self._scale_loss, True, hcg=self._hcg)
            else:
                self._scaler = amp_api.GradScaler(False, 1.0, False)
        else:
            self._scaler = None

        # optimizer + lr
        self._lr_scheduler, self._optimizer = None, None
        if mode == "train":
            self._use_increments = bool(configs.Optimizer.get("lr", {}).get("use_increments", False)) if isinstance(configs.Optimizer.get("lr"), dict) else False
            self._lr_scheduler_mode = "step"
            lr_cfg = configs.Optimizer.get("lr")
            if isinstance(lr_cfg, dict):
                # per-epoch schedules say so with ``run_mode: epoch`` (reference moco recipes) — or just ``update_unit: epoch``
                self._lr_scheduler_mode = lr_cfg.pop("run_mode", None) or ("epoch" if lr_cfg.get("update_unit") == "epoch" else "step")
                lr_cfg_clean = {k: v for k, v in lr_cfg.items() if k not in ("_scaled_by_batch",)}
            else:
                lr_cfg_clean = lr_cfg
            self._lr_scheduler = build_lr_scheduler(lr_cfg_clean) if lr is None else lr
            if self._sharding_stage == 3 and self._sharding_degree > 1 and self._pp_degree == 1 and env.world_size() > 1:
                from ...distributed.apis.strategy import broadcast_initial_parameters
                from ...parallel.sharding import GroupShardedStage3

                broadcast_initial_parameters(self._module.model, self._hcg)
                self._module.model = GroupShardedStage3(self._module.model, self._hcg)
            self._optimizer = build_optimizer(configs.Optimizer, self._module.model, self._lr_scheduler, hcg=self._hcg,
                                              dist_config=d, amp_config=amp) if optimizer is None else optimizer

        # distributed wrappers
        if env.world_size() > 1 and not hasattr(self._module.model, "optimizer_named_parameters"):
            self._module.model, self._optimizer, self._scaler = wrap_with_fleet(d, self._module.model, self._optimizer, self._scaler)

        if mode == "train" and hasattr(self._optimizer, "install_forward_hooks"):
            self._optimizer.install_forward_hooks(self._module.model)      # overlapped ZeRO parameter all-gather (no-op otherwise)

        self._load_recovery = {"step": 0, "epoch": 0, "rng_state": None}
        self._rng_restore_pending = False
        self._profiler = None
        if configs.get("Profiler", {}).get("enable", False) and mode == "train":
            from ...utils.profiler import StepProfiler

            self._profiler = StepProfiler(configs.Profiler)
        # failure detection (utils/watchdog.py): per-rank heartbeat files + stall monitor, SIGTERM/SIGUSR1 emergency checkpoint,
        # PFX_FAULT test hook.  Engine.watchdog: {enable, timeout, dir}
        from ...utils import watchdog as _wd

        self._wd = _wd
        self._fault = _wd.FaultInjector(rank=env.global_rank())
        self._stop_requested = False
        self._last_saved_step = -1
        self._heartbeat = None
        wcfg = (configs.Engine.get("watchdog") or {}) if mode == "train" else {}
        if wcfg.get("enable", False):
            self._heartbeat = _wd.Heartbeat(wcfg.get("dir") or os.path.join(self._output_dir or ".", "heartbeat"),
                                            int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
                                            timeout_s=float(wcfg.get("timeout", 600))).start()
        if mode == "train":
            _wd.install_signal_checkpoint()
        self._inference_engine = None
        self._train_tokens = None

    # ---------------------------------------------------------------------------------------- helpers
    def _amp_ctx(self):
        return amp_api.autocast_context(self._use_pure_fp16, self._amp_dtype, self._amp_level, self._device.type)

    @property
    def optimizer(self):
        return self._optimizer

    @property
    def module(self):
        return self._module

    # ---------------------------------------------------------------------------------------- fit
    def fit(self, epoch: int = 1, train_data_loader=None, valid_data_loader=None):
        self._module.model.train()
        train_cost = 0.0
        start_epoch = self._load_recovery["epoch"]
        if self._run_mode == "epoch" and train_data_loader is not None and 0 < len(train_data_loader) <= self._load_recovery["step"]:
            # an end-of-epoch checkpoint: that epoch (its evaluation, LR epoch step and save included) is done — continue with the next one
            start_epoch += 1
            self._load_recovery.update(epoch=start_epoch, step=0)
        self._restore_rng()
        for epoch_index in range(start_epoch, epoch):
            t0 = get_timestamp()
            self._train_one_epoch(epoch_index, train_data_loader, valid_data_loader)
            if self._stop_requested:
                break
            train_cost = get_timestamp() - t0
            self._module.training_epoch_end({"epoch": epoch_index, "train_cost": train_cost})
            if self._run_mode == "epoch":
                if self._lr_scheduler_mode == "epoch" and isinstance(self._lr_scheduler, LRScheduler):
                    self._lr_scheduler.step()
                if valid_data_loader is not None and self._eval_freq and epoch_index % max(self._eval_freq, 1) == 0 and self._eval_freq > 0:
                    e0 = get_timestamp()
                    self._evaluate_one_epoch(epoch_index, valid_data_loader)
                    self._module.validation_epoch_end({"epoch": epoch_index, "eval_cost": get_timestamp() - e0})   # metric report (accuracy, F1, ...)
                    self._module.model.train()
                if self._save_epoch and (epoch_index + 1) % self._save_epoch == 0:
                    self.save(epoch=epoch_index, step=len(train_data_loader) if train_data_loader is not None else 0)
        if self._profiler is not None:
            self._profiler.finish()
        if self._heartbeat is not None:
            self._heartbeat.stop()

    def _restore_rng(self) -> None:
        """Random-number streams of the checkpoint (device generator, host generator, the named tensor-parallel streams)."""
        ckpt_io.restore_rng(self._load_recovery)

    def _train_one_epoch(self, epoch_index: int, train_data_loader, valid_data_loader):
        self._module.model.train()
        device = self._device
        loader = _Prefetcher(train_data_loader, device)
        total_steps = self._max_steps if self._run_mode == "step" else len(train_data_loader)
        losses = []
        skip_first = True
        train_start = get_timestamp()
        ev0 = torch.cuda.Event(enable_timing=True) if device.type == "cuda" else None
        if ev0 is not None:
            ev0.record()
        resume_step = self._load_recovery["step"] if epoch_index == self._load_recovery["epoch"] else 0
        last_step = None
        first = 0
        sampler = getattr(train_data_loader, "batch_sampler", None)
        if hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch_index)      # a new shuffle every epoch (paddle.io.DistributedBatchSampler advances its epoch per pass)
        if resume_step and not self._resume_replay and hasattr(sampler, "skip_batches"):
            # resume: the sampler starts ``resume_step`` batches in, so consumed data is never read (the reference replays and discards it,
            # eager_engine.py:347-349 — ``Engine.save_load.resume_replay: True`` keeps that, e.g. to reproduce host-side augmentation draws)
            sampler.skip_batches = first = resume_step
        for step, batch in enumerate(loader, first):
            if step < resume_step:
                continue          # replayed batch of a resumed run
            if self._rng_restore_pending:
                # creating loader iterators and replaying consumed batches drew from the host generator (DataLoader base seed): put the streams
                # back to the checkpointed state right before the first step that trains — in whichever epoch that is — so noise (dropout, MoE
                # routing) continues where it stopped
                self._restore_rng()
                self._rng_restore_pending = False
            loss = self._fit_impl(batch)
            loss = self._fault.maybe_fire(step, loss)
            losses.append(loss)
            if self._heartbeat is not None:
                self._heartbeat.beat(step)
            found_inf = self._scaler.found_inf if (self._scaler is not None and self._amp_dtype == "float16") else False
            if self._lr_scheduler_mode == "step" and isinstance(self._lr_scheduler, LRScheduler) and not found_inf:
                self._lr_scheduler.step(epoch=self._global_batch_size if self._use_increments else None)
            if self._wd.emergency_requested():
                if device.type == "cuda":
                    torch.cuda.synchronize()
                logger.warning(f"emergency checkpoint at epoch {epoch_index} step {step + 1}, then stopping")
                self._optimizer.clear_grad()
                self.save(epoch=epoch_index, step=step + 1)
                self._stop_requested = True
                return

            if (step + 1) % self._logging_freq == 0:
                now = get_timestamp()
                train_cost = (now - train_start) / self._logging_freq
                dev_ms = None
                if ev0 is not None:
                    ev1 = torch.cuda.Event(enable_timing=True)
                    ev1.record(); ev1.synchronize()
                    dev_ms = ev0.elapsed_time(ev1) / self._logging_freq
                    ev0 = ev1
                vals = [float(l) for l in losses]
                log = {"epoch": epoch_index, "total_epoch": self._num_train_epochs, "batch": step, "total_step": total_steps,
                       "total_batch": total_steps, "train_cost": train_cost, "device_ms": dev_ms, "loss": sum(vals) / len(vals),
                       "lr": self._optimizer.get_lr(), "found_inf": float(found_inf),
                       "loss_scale": self._scaler.get_scale() if (self._scaler is not None and self._amp_dtype == "float16") else None}
                self._module.training_step_end(log)
                losses = []
                train_start = get_timestamp()

            self._optimizer.clear_grad()

            if self._run_mode == "step" and not skip_first:
                if self._eval_freq and self._eval_freq > 0 and step % self._eval_freq == 0 and valid_data_loader is not None:
                    self._module.model.eval()
                    eval_losses, t_eval = [], get_timestamp()
                    for eval_step, ebatch in enumerate(_Prefetcher(valid_data_loader, device)):
                        eval_losses.append(self._evaluate_impl(ebatch))
                        if eval_step >= self._eval_iters - 1:
                            break
                    ecost = (get_timestamp() - t_eval) / self._logging_freq    # (sic) reference divides by logging_freq
                    ev = [float(l) for l in eval_losses]
                    self._module.validation_step_end({"loss": sum(ev) / max(len(ev), 1), "epoch": epoch_index, "batch": eval_step,
                                                      "total_batch": total_steps, "eval_cost": ecost})
                    self._module.model.train()
                    train_start = get_timestamp()
                # A checkpoint is named by the number of COMPLETED steps: ``epoch_0_step_3`` holds the state after batches 0, 1, 2, and a resumed
                # run (which skips ``step < 3``) continues with batch 3 — bit-identical to the uninterrupted run.  (The reference saves after
                # training batch index ``save_steps`` under the same name, so its resumed run trains that batch a second time.)
                if self._save_steps and self._save_steps > 0 and (step + 1) % self._save_steps == 0:
                    if device.type == "cuda":
                        torch.cuda.synchronize()
                    self.save(epoch=epoch_index, step=step + 1)
                    self._last_saved_step = step + 1
            else:
                skip_first = False

            if self._profiler is not None:
                self._profiler.step()
            last_step = step
            if self._run_mode == "step" and step >= self._max_steps:
                break
        # step mode: the loader is sized to max_steps batches, so the periodic `step % save_steps` rule never fires on the last one;
        # leave a final checkpoint (named by the number of consumed batches) when periodic saving is on
        if (self._run_mode == "step" and self._save_steps and 0 < self._save_steps < (1 << 30) and last_step is not None
                and self._last_saved_step != last_step + 1 and not self._stop_requested):
            if device.type == "cuda":
                torch.cuda.synchronize()
            self.save(epoch=epoch_index, step=last_step + 1)

    # ---------------------------------------------------------------------------------------- one optimizer step
    def train_step(self, host_batch):
        """Public single-step API: ``host_batch`` is what the dataloader yields (pinned host tensors).  Copies it to
        the device, runs forward/backward over all micro-batches, the optimizer step, the LR step and clears the
        gradients; returns the (device) loss tensor."""
        batch = _to_device(host_batch, self._device, True)
        loss = self._fit_impl(batch)
        if self._lr_scheduler_mode == "step" and isinstance(self._lr_scheduler, LRScheduler):
            self._lr_scheduler.step(epoch=self._global_batch_size if self._use_increments else None)
        self._optimizer.clear_grad()
        return loss

    def _fit_impl(self, batch):
        self._module.model.train()
        batch = self._module.pretreating_batch(batch)
        if self._pp_degree == 1:
            loss = self._model_forward_backward(batch)
        else:
            with self._amp_ctx():
                self._module.model._prepare_training(batch, self._optimizer, self._lr_scheduler)
                loss = self._module.model.forward_backward_pipeline(batch, self._scaler)
            if self._mp_degree > 1 and self._configs.Model.get("sequence_parallel", False):
                allreduce_sequence_parallel_grads(self._module.model)      # LayerNorm / row-bias grads are sequence-partial on every stage
        self._optim_update_params()
        return loss

    def _model_forward_backward(self, batch):
        n = self._accumulate_steps
        if n > 1 and isinstance(batch, (list, tuple)) and len(batch) == n and all(isinstance(b, (list, tuple)) for b in batch):
            micro_batches = list(batch)          # collate already produced per-micro-batch lists (ErnieCollateData)
        else:
            micro_batches = _split_micro(batch, n)
        total = None
        for i, mb in enumerate(micro_batches):
            last = i == n - 1
            sync_ctx = self._optimizer.no_sync() if (not last and hasattr(self._optimizer, "no_sync")) else _Null()
            with sync_ctx:
                with self._amp_ctx():
                    loss = self._module.training_step(mb)
                loss_bw = self._scaler.scale(loss) if (self._scaler is not None and self._amp_dtype == "float16") else loss
                if n > 1:
                    loss_bw = loss_bw / n
                bw_ctx = self._module.model.backward_phase() if hasattr(self._module.model, "backward_phase") else _Null()
                C.run_pre_backward()       # leftover side-stream parameter updates (modules that did not run this step) finish first
                with bw_ctx:
                    self._module.backward(loss_bw)
            d = loss.detach()
            total = d if total is None else total + d
        if self._mp_degree > 1 and self._configs.Model.get("sequence_parallel", False):
            allreduce_sequence_parallel_grads(self._module.model)
        return total / n if n > 1 else total

    def _optim_update_params(self):
        if self._scaler is not None and self._amp_dtype == "float16":
            self._scaler.step(self._optimizer)
            self._scaler.update()
        else:
            self._optimizer.step()
        if hasattr(self._module.model, "after_optimizer_step"):
            self._module.model.after_optimizer_step()

    # ---------------------------------------------------------------------------------------- evaluate / predict
    @torch.no_grad()
    def evaluate(self, epoch: int = 1, valid_data_loader=None):
        self._module.model.eval()
        for e in range(epoch):
            t0 = get_timestamp()
            self._evaluate_one_epoch(e, valid_data_loader)
            self._module.validation_epoch_end({"epoch": e, "eval_cost": get_timestamp() - t0})

    @torch.no_grad()
    def _evaluate_one_epoch(self, epoch_index: int, valid_data_loader):
        self._module.model.eval()
        t0 = get_timestamp()
        losses = []
        total = len(valid_data_loader) if hasattr(valid_data_loader, "__len__") else -1
        for step, batch in enumerate(_Prefetcher(valid_data_loader, self._device)):
            losses.append(self._evaluate_impl(batch))
            if (step + 1) % self._logging_freq == 0:
                cost = (get_timestamp() - t0) / self._logging_freq
                vals = [float(l) for l in losses]
                self._module.validation_step_end({"loss": sum(vals) / len(vals), "epoch": epoch_index, "batch": step, "total_batch": total,
                                                  "eval_cost": cost})
                t0, losses = get_timestamp(), []
            if self._run_mode == "step" and self._eval_iters and step >= self._eval_iters - 1:
                break
        if losses:        # batches after the last full logging window (or an evaluation set shorter than one window)
            vals = [float(l) for l in losses]
            self._module.validation_step_end({"loss": sum(vals) / len(vals), "epoch": epoch_index, "batch": step, "total_batch": total,
                                              "eval_cost": (get_timestamp() - t0) / len(vals)})

    @torch.no_grad()
    def _evaluate_impl(self, batch):
        batch = self._module.pretreating_batch(batch)
        with self._amp_ctx():
            if self._pp_degree == 1:
                loss = self._module.validation_step(batch)
            else:
                loss = self._module.model.eval_batch(batch, compute_loss=True)
        return loss.detach() if isinstance(loss, torch.Tensor) else loss

    @torch.no_grad()
    def predict(self, epoch: int = 1, test_data_loader=None):
        self._module.model.eval()
        for e in range(epoch):
            t0 = get_timestamp()
            losses = []
            for step, batch in enumerate(_Prefetcher(test_data_loader, self._device)):
                losses.append(self._predict_impl(batch))
                if (step + 1) % self._logging_freq == 0:
                    cost = (get_timestamp() - t0) / self._logging_freq
                    vals = [float(l) for l in losses]
                    self._module.test_step_end({"loss": sum(vals) / len(vals), "epoch": e, "batch": step, "test_cost": cost})
                    t0, losses = get_timestamp(), []
                if self._run_mode == "step" and self._test_iters and step >= self._test_iters - 1:
                    break

    @torch.no_grad()
    def _predict_impl(self, batch):
        batch = self._module.pretreating_batch(batch)
        with self._amp_ctx():
            if self._pp_degree == 1:
                loss = self._module.test_step(batch)
            else:
                loss = self._module.model.eval_batch(batch, compute_loss=True)
        return loss.detach() if isinstance(loss, torch.Tensor) else loss

    # ---------------------------------------------------------------------------------------- checkpoint
    def save(self, epoch: int = 0, step: int = 0):
        if self._output_dir is None:
            return
        if hasattr(self._optimizer, "finish_param_sync"):
            self._optimizer.finish_param_sync()
        model = self._module.model
        if self._sharding_stage == 3 and self._sharding_degree > 1 and hasattr(model, "get_all_parameters"):
            model.get_all_parameters()
        d = ckpt_io.save(self._output_dir, model, self._optimizer, step=step, epoch=epoch, scaler=self._scaler)
        if self._peft is not None and d is not None:
            from ...utils.peft import adapter_state_dict

            torch.save({k: v.detach().cpu() for k, v in adapter_state_dict(model).items()}, os.path.join(d, "adapter.pdparams"))
        if self._configs.Engine.save_load.get("save_auto_inference", False):
            # layout-annotated copy of the weights for serving on a different tensor-parallel degree (reference: always on, eager_engine.py:750)
            ckpt_io.save_for_auto_inference(os.path.join(self._output_dir, "auto_infer", "auto"), model)

    def load(self):
        if not self._ckpt_dir:
            logger.warning("`load` requires Engine.save_load.ckpt_dir; skipping")
            return
        rec = ckpt_io.load(self._ckpt_dir, self._module.model, self._optimizer if self.mode == "train" else None,
                           mode="train" if self.mode == "train" else "eval", scaler=self._scaler)
        if rec:
            self._load_recovery.update(rec)
            self._rng_restore_pending = self.mode == "train"

    # ---------------------------------------------------------------------------------------- export / inference
    def export(self):
        from ...utils.export import export_inference_model

        self._module.model.eval()
        if self._peft is not None and self._peft["method"] == "lora" and self._configs.PEFT.get("merge_on_export", True):
            from ...utils.peft import merge_lora

            merge_lora(self._module.model)          # W += (alpha / r) B A: the exported model has the plain architecture
        save_dir = os.path.join(self._output_dir, f"rank_{self._dp_rank}")
        sq = (self._configs.get("Compress") or {}).get("SmoothQuant") or {}
        smooth = None
        if sq.get("enable", False):
            # post-training Shift-SmoothQuant (W8A8): calibrate per-channel activation ranges, migrate them into the weights, quantise
            from ...utils import smoothquant as SQ

            model = self._module.model
            core = next((m for m in model.modules() if hasattr(getattr(m, "decoder", None), "layers")), model)
            dev = next(model.parameters()).device
            batches = SQ.calibration_batches(self._configs, int(sq.get("calib_batches", 8)), int(sq.get("calib_batch_size", 4)), sq.get("calib_seq_len"))
            # hooks go on the exported model's own module names; the forward runs the transformer trunk, where all the linears live
            stats = SQ.calibrate(model, batches, forward_fn=lambda _m, b: core(b[0].to(dev), b[1].to(dev)))
            SQ.smooth_and_quantize(model, stats, alpha=float(sq.get("alpha", 0.5)), shift=bool(sq.get("shift", True)))
            smooth = {"alpha": float(sq.get("alpha", 0.5)), "shift": bool(sq.get("shift", True)), "layers": model.smooth_quant_layers}
            logger.info(f"Shift-SmoothQuant: {len(model.smooth_quant_layers)} linear layers quantised to W8A8 from {len(batches)} calibration batches")
        export_inference_model(self._module.model, self._module.input_spec(), save_dir, "model", configs=self._configs,
                               quant=self._quant_mode, smooth_quant=smooth)
        logger.info(f"export inference model saved in {save_dir}")

    def inference(self, data):
        if self._inference_engine is None:
            from .inference_engine import InferenceEngine

            inf = self._configs.Inference
            self._inference_engine = InferenceEngine(inf.model_dir, inf.mp_degree)
        return self._inference_engine.predict(data)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

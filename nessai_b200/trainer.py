"""Host side of the fused training kernels (``csrc/train.cuh``).

Replaces the autograd loop of ``FlowModel._train`` / ``_validate``
(/root/reference/src/nessai/flowmodel/base.py:365-523): one C-ABI call runs a
whole epoch of forward + backward + clip + optimiser steps on the flat parameter
buffers of a :class:`~nessai_b200.flowmodel.B200Flow`.  The torch optimiser object
the reference API exposes (``FlowModel._optimiser``) is kept as the holder of the
hyper-parameters; when it is one the kernels implement (AdamW / Adam / plain SGD)
the update runs in the kernel, otherwise the kernel produces the clipped gradient
and ``optimiser.step()`` applies it.  There is no autograd and no CPU path.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .train_plan import build_train_plan, param_mask


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class FusedTrainer:
    def __init__(self, model):
        self.model = model
        self.device = model.device
        plan, itab, reduce_idx = build_train_plan(model.spec, model.ints)
        self._plan, self._itab, self._reduce = plan, itab, reduce_idx
        self._handle = C.c_void_p()
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(
                lib.nb200_trainer_create(
                    C.byref(self._handle),
                    plan.ctypes.data_as(C.c_void_p), int(plan.size),
                    itab.ctypes.data_as(C.c_void_p), int(itab.size),
                    reduce_idx.ctypes.data_as(C.c_void_p), int(reduce_idx.size),
                ),
                "nb200_trainer_create",
            )
        n = model.spec.n_params
        segs = param_mask(model.spec)
        if segs:
            # MADE: masks are float buffers of the flow (theta_b)
            tb = model.theta_b.detach().cpu().numpy()
            pm = np.ones(n, dtype=np.float32)
            for w_off, m_off, size in segs:
                pm[w_off : w_off + size] = tb[m_off : m_off + size]
            with torch.cuda.device(self.device):
                _lib.check(lib.nb200_trainer_set_param_mask(self._handle, pm.ctypes.data_as(C.c_void_p)),
                           "nb200_trainer_set_param_mask")
        self.m = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.v = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.step = 0
        self._opt_id = None
        self._loss = torch.zeros(2, device=self.device, dtype=torch.float32)

    def rebind(self, model) -> bool:
        """Point the trainer at another flow of the same architecture (same plan, same index
        tables): fresh optimiser state, same kernels and workspaces.  False if it cannot."""
        if model.spec.n_params != self.model.spec.n_params or model.device != self.device:
            return False
        if param_mask(model.spec):
            return False  # MADE masks are uploaded per trainer
        plan, itab, reduce_idx = build_train_plan(model.spec, model.ints)
        if not (np.array_equal(plan, self._plan) and np.array_equal(reduce_idx, self._reduce)
                and itab.shape == self._itab.shape):
            return False
        if not np.array_equal(itab, self._itab):
            # the new flow's own permutations / index lists (reset_permutations draws new ones)
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().nb200_trainer_set_itab(self._handle, itab.ctypes.data_as(C.c_void_p),
                                                              int(itab.size), st), "nb200_trainer_set_itab")
            self._itab = itab
        self.model = model
        self.m.zero_()
        self.v.zero_()
        self.step = 0
        self._opt_id = None
        return True

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                _lib.load().nb200_trainer_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------- optimiser
    def _bind(self, optimiser) -> None:
        """A new optimiser object means fresh optimiser state (the reference builds
        a new one in ``initialise`` and ``reset_model``, flowmodel/base.py:771)."""
        if self._opt_id != id(optimiser):
            self.m.zero_()
            self.v.zero_()
            self.step = 0
            self._opt_id = id(optimiser)

    @staticmethod
    def _kernel_optimiser(optimiser):
        """(kind, lr, beta1, beta2, eps, weight_decay) if the kernel implements this
        optimiser configuration, else None."""
        if optimiser is None or len(optimiser.param_groups) != 1:
            return None
        g = optimiser.param_groups[0]
        if g.get("maximize") or g.get("amsgrad") or g.get("capturable") or g.get("differentiable"):
            return None
        if isinstance(optimiser, torch.optim.AdamW) or (
            isinstance(optimiser, torch.optim.Adam) and g.get("decoupled_weight_decay", False)
        ):
            kind = 0
        elif isinstance(optimiser, torch.optim.Adam):
            kind = 1
        elif isinstance(optimiser, torch.optim.SGD):
            if g.get("momentum", 0) or g.get("dampening", 0) or g.get("nesterov", False):
                return None
            return (2, float(g["lr"]), 0.0, 0.0, 0.0, float(g.get("weight_decay", 0.0)))
        else:
            return None
        if isinstance(g["lr"], torch.Tensor):
            return None
        b1, b2 = g["betas"]
        return (kind, float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]))

    # ----------------------------------------------------------------- epoch
    def epoch(self, x_all, w_all, perm, batch_size, optimiser, clip) -> torch.Tensor:
        """All optimisation steps of one epoch; returns the device scalar holding the
        SUM of the batch losses (no host synchronisation)."""
        model = self.model
        self._bind(optimiser)
        n_rows = int(x_all.shape[0])
        cfg = self._kernel_optimiser(optimiser)
        clip = float(clip) if clip else 0.0
        lib = _lib.load()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        loss_sum = self._loss[:1]
        loss_sum.zero_()
        with torch.cuda.device(self.device):
            if cfg is not None:
                kind, lr, b1, b2, eps, wd = cfg
                _lib.check(
                    lib.nb200_train_epoch(
                        self._handle, _ptr(model.theta_p), _ptr(model.theta_b), _ptr(self.m), _ptr(self.v),
                        _ptr(x_all), _ptr(w_all), _ptr(perm), n_rows, int(batch_size),
                        kind, lr, b1, b2, eps, wd, clip, self.step, _ptr(loss_sum), None, st,
                    ),
                    "nb200_train_epoch",
                )
                self.step += (n_rows + batch_size - 1) // batch_size
            else:
                # gradient from the kernels, update by the caller's torch optimiser
                for i0 in range(0, n_rows, batch_size):
                    n_b = min(batch_size, n_rows - i0)
                    self._grad_step(x_all, w_all, perm, i0, n_b, clip, loss_sum, None, st)
                    model.theta_p.grad = self.grad()
                    optimiser.step()
                    self.step += 1
        return loss_sum

    # ------------------------------------------------------------ epoch runs
    MAX_CHUNK = 64  # TR_MAX_CHUNK of csrc/train.cuh

    def begin_run(self):
        """Loop-control block and best-weights snapshot of one ``train()`` call
        (flowmodel/base.py:608-618: ``best_val_loss = inf``, ``best_epoch = 0``, initial weights)."""
        model = self.model
        ctl = np.zeros(4, dtype=np.int32)
        ctl[:1].view(np.float32)[0] = np.inf
        self._ctl = torch.from_numpy(ctl).to(self.device)
        self._best_p = model.theta_p.detach().clone()
        self._best_b = model.theta_b.detach().clone()
        n = 2 * self.MAX_CHUNK + 4
        if getattr(self, "_run_dev", None) is None:
            self._run_dev = torch.zeros(n, device=self.device, dtype=torch.float32)
            self._run_host = torch.zeros(n, dtype=torch.float32, pin_memory=True)

    def run(self, x, w, perms, batch_size, xv, wv, optimiser, clip, lrs, epoch0, validate, patience):
        """Up to ``MAX_CHUNK`` epochs -- optimisation steps, validation loss, best-weights
        snapshot and patience -- in one cooperative launch (``nb200_train_run``), then ONE
        read-back.  ``perms``: ``(n_epochs, n_rows)`` int64 device tensor, ``lrs``: the learning
        rate of every epoch.  Returns ``(epochs_done, stop, best_epoch, hist)`` with
        ``epochs_done`` the total number of epochs finished and ``hist`` the float32
        ``(n_run, 2)`` array of {sum of batch losses, validation loss} of the epochs that ran."""
        model = self.model
        self._bind(optimiser)
        cfg = self._kernel_optimiser(optimiser)
        if cfg is None:
            raise RuntimeError("optimiser not implemented by the training kernels")
        kind, _, b1, b2, eps, wd = cfg
        n_epochs, n_rows = int(perms.shape[0]), int(x.shape[0])
        assert 1 <= n_epochs <= self.MAX_CHUNK and perms.shape[1] == n_rows and perms.dtype == torch.int64
        lrs = np.ascontiguousarray(lrs, dtype=np.float64)
        assert lrs.shape == (n_epochs,)
        n_val = 0 if xv is None else int(xv.shape[0])
        hist, ctl = self._run_dev[: 2 * self.MAX_CHUNK], self._ctl
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.load().nb200_train_run(
                    self._handle, _ptr(model.theta_p), _ptr(model.theta_b), int(model.theta_b.numel()),
                    _ptr(self.m), _ptr(self.v), _ptr(x), _ptr(w), _ptr(perms), n_rows, int(batch_size),
                    _ptr(xv) if n_val else None, _ptr(wv) if n_val else None, n_val, n_epochs, int(epoch0),
                    int(bool(validate)), int(patience), kind, lrs.ctypes.data_as(C.c_void_p), b1, b2, eps, wd,
                    float(clip) if clip else 0.0, self.step, _ptr(hist), _ptr(ctl), _ptr(self._best_p),
                    _ptr(self._best_b), st,
                ),
                "nb200_train_run",
            )
            # the single host synchronisation of the run: history + loop control in one read
            self._run_dev[2 * self.MAX_CHUNK :].copy_(ctl.view(torch.float32))
            self._run_host.copy_(self._run_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        out = self._run_host.numpy()
        c = out[2 * self.MAX_CHUNK :].view(np.int32)
        epochs_done, stop, best_epoch = int(c[2]), bool(c[3]), int(c[1])
        n_run = epochs_done - int(epoch0)
        self.step += n_run * ((n_rows + batch_size - 1) // batch_size)
        return epochs_done, stop, best_epoch, out[: 2 * n_run].reshape(n_run, 2).copy()

    def restore_best(self):
        """Best-epoch weights back into the model (flowmodel/base.py:664-667)."""
        with torch.no_grad():
            self.model.theta_p.copy_(self._best_p)
            self.model.theta_b.copy_(self._best_b)

    def _grad_step(self, x_all, w_all, perm, i0, n_b, clip, loss_sum, info, st):
        """Loss + clipped gradient of the batch ``perm[i0 : i0 + n_b]`` (no update)."""
        if perm is not None:
            x, w, p = x_all, w_all, perm[i0 : i0 + n_b]
        else:
            x = x_all[i0 : i0 + n_b]
            w = None if w_all is None else w_all[i0 : i0 + n_b]
            p = None
        _lib.check(
            _lib.load().nb200_train_epoch(
                self._handle, _ptr(self.model.theta_p), _ptr(self.model.theta_b), None, None,
                _ptr(x), _ptr(w), _ptr(p), n_b, n_b, -1, 0.0, 0.0, 0.0, 0.0, 0.0, float(clip),
                self.step, _ptr(loss_sum), _ptr(info), st,
            ),
            "nb200_train_epoch",
        )

    def grad(self) -> torch.Tensor:
        """Gradient of the last step (after clipping) as a new device tensor."""
        out = torch.empty(self.model.spec.n_params, device=self.device, dtype=torch.float32)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().nb200_trainer_copy_grad(self._handle, _ptr(out), st), "nb200_trainer_copy_grad")
        return out

    def loss_and_grad(self, x, w=None, clip=0.0):
        """One gradient-only step on the rows of ``x`` (tests / diagnostics):
        returns ``(loss, grad)`` as device tensors; parameters are not updated, the
        BatchNorm running statistics are (train mode)."""
        info = torch.zeros(2, device=self.device, dtype=torch.float32)
        loss_sum = torch.zeros(1, device=self.device, dtype=torch.float32)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        with torch.cuda.device(self.device):
            self._grad_step(x, w, None, 0, int(x.shape[0]), clip, loss_sum, info, st)
        return loss_sum[0], self.grad(), info

    # ------------------------------------------------------------ validation
    def eval_loss(self, x, w=None) -> torch.Tensor:
        """Eval-mode loss (running statistics) of the current parameters: device scalar."""
        out = self._loss[1:2]
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.load().nb200_eval_loss(
                    self._handle, _ptr(self.model.theta_p), _ptr(self.model.theta_b), _ptr(x), _ptr(w),
                    int(x.shape[0]), _ptr(out), None, st,
                ),
                "nb200_eval_loss",
            )
        return out

    def log_prob_unfolded(self, x) -> torch.Tensor:
        """Per-row eval-mode log_prob straight from the unfolded parameters."""
        lp = torch.empty(int(x.shape[0]), device=self.device, dtype=torch.float32)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.load().nb200_eval_loss(
                    self._handle, _ptr(self.model.theta_p), _ptr(self.model.theta_b), _ptr(x), None,
                    int(x.shape[0]), None, _ptr(lp), st,
                ),
                "nb200_eval_loss",
            )
        return lp


def cosine_annealing_lr(base_lr: float, epoch: int, t_max: int, eta_min: float = 0.0) -> float:
    """Closed form of ``torch.optim.lr_scheduler.CosineAnnealingLR`` after ``epoch`` steps."""
    return eta_min + (base_lr - eta_min) * (1 + np.cos(np.pi * epoch / t_max)) / 2

"""``B200FlowModel``: drop-in for ``nessai.flowmodel.FlowModel`` (plugin point P2).

Mirrors /root/reference/src/nessai/flowmodel/base.py:25-957 -- same constructor,
methods, argument meaning, return types (fresh float64 numpy arrays), error
behaviour and config handling -- with every flow evaluation executed by the CUDA
library behind ``include/nessai_b200.h``.  ``BaseFlowProposal`` instantiates the
class named by its ``_FlowModelClass`` attribute
(/root/reference/src/nessai/proposal/flowproposal/base.py:101,383-389) and only
duck-types it, so this class does not need nessai to be importable.
"""

from __future__ import annotations

import copy
import ctypes as C
import json
import logging
import os
import shutil
from collections import OrderedDict
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .spec import FlowSpec, get_n_neurons

logger = logging.getLogger(__name__)

_FLOW_DEFAULTS = dict(
    n_inputs=None,
    n_neurons=None,
    n_blocks=4,
    n_layers=2,
    ftype="RealNVP",
    flow=None,
    distribution=None,
    distribution_kwargs=None,
)
_TRAINING_DEFAULTS = dict(
    device_tag="cuda",  # the reference default is "cpu" (flowmodel/config.py:32)
    inference_device_tag=None,
    lr=0.001,
    annealing=False,
    clip_grad_norm=5.0,
    batch_size=1000,
    val_size=0.1,
    max_epochs=500,
    patience=20,
    noise_type=None,
    noise_scale=None,
    use_dataloader=False,
    optimiser="adamw",
    optimiser_kwargs=None,
)


def update_flow_config(cfg):
    """Mirror of flowmodel/utils.py:12-43."""
    default = copy.deepcopy(_FLOW_DEFAULTS)
    if cfg is None:
        return default
    if not isinstance(cfg, dict):
        raise TypeError("Must pass a dictionary to update the default model config")
    default.update(copy.deepcopy(cfg))
    default["n_neurons"] = get_n_neurons(
        n_neurons=default.get("n_neurons"), n_inputs=default.get("n_inputs")
    )
    return default


def update_training_config(cfg):
    """Mirror of flowmodel/utils.py:46-67."""
    default = copy.deepcopy(_TRAINING_DEFAULTS)
    if cfg is None:
        return default
    if not isinstance(cfg, dict):
        raise TypeError("Must pass a dictionary to update the default model config")
    default.update(copy.deepcopy(cfg))
    if default["noise_type"] is not None and default["noise_scale"] is None:
        raise RuntimeError("`noise_scale` must be specified when `noise_type` is given.")
    if isinstance(default["noise_scale"], float):
        if default["noise_type"] is None:
            default["noise_type"] = "constant"
    elif default["noise_scale"] is not None:
        raise TypeError(
            "`noise_scale` must be a float. "
            f"'Got type: {type(default['noise_scale'])}"
        )
    return default


def update_config(flow_config, training_config=None):
    return update_flow_config(flow_config), update_training_config(training_config)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


_PINNED_STAGE = [None]


def _pinned_stage(nbytes: int) -> torch.Tensor:
    """A page-locked uint8 staging buffer of at least ``nbytes`` (module-wide, grow-only)."""
    buf = _PINNED_STAGE[0]
    if buf is None or buf.numel() < nbytes:
        size = max(int(1.5 * nbytes), 8 << 20)
        _PINNED_STAGE[0] = None  # release the old block before page-locking the new one
        buf = torch.empty(size, dtype=torch.uint8, pin_memory=True)
        _PINNED_STAGE[0] = buf
    return buf


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class B200Flow(torch.nn.Module):
    """Device-resident flow: flat parameters + the C-ABI flow handle.

    Presents the surface of the reference's ``BaseFlow``
    (/root/reference/src/nessai/flows/base.py:11-167) on torch tensors; the
    ``state_dict`` uses the reference key layout so ``model.pt`` files are meant to
    interchange with ``nessai.flows.RealNVP`` / ``NeuralSplineFlow``.  That is VERIFIED against
    the reference running on the restated nflows layer (``oracle/shims/glasflow``, both
    directions, ``tests/test_gpu_nessai_plugin.py``); against a real ``glasflow`` install it is
    checked by ``tests/test_spec.py::test_interchange_with_real_glasflow`` wherever one exists
    (this image has none).
    """

    def __init__(self, spec: FlowSpec, device: torch.device):
        super().__init__()
        self.spec = spec
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(
                f"nessai_b200 flows live on a CUDA device; got {self.device}. "
                "There is no CPU fallback."
            )
        if not torch.cuda.is_available():
            raise RuntimeError("nessai_b200 requires a CUDA device (B200, sm_100a)")
        theta, ints = spec.init_state()
        self.ints = ints
        self.ints_version = 0
        self.theta_p = torch.nn.Parameter(
            torch.from_numpy(theta[: spec.n_params].copy()).to(self.device)
        )
        self.register_buffer(
            "theta_b", torch.from_numpy(theta[spec.n_params :].copy()).to(self.device)
        )
        self._handle = C.c_void_p()
        self._dirty = True
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(
                lib.nb200_flow_create(C.byref(self._handle), spec.D, spec.H, spec.activation),
                "nb200_flow_create",
            )
            if spec.base_var != 1.0:
                _lib.check(lib.nb200_flow_set_base_variance(self._handle, float(spec.base_var)),
                           "nb200_flow_set_base_variance")

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                _lib.load().nb200_flow_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ---------------------------------------------------------------- params
    def theta_numpy(self) -> np.ndarray:
        return np.concatenate(
            [self.theta_p.detach().cpu().numpy(), self.theta_b.detach().cpu().numpy()]
        ).astype(np.float32)

    def set_theta_numpy(self, theta: np.ndarray) -> None:
        n = self.spec.n_params
        with torch.no_grad():
            self.theta_p.copy_(torch.from_numpy(np.ascontiguousarray(theta[:n])))
            self.theta_b.copy_(torch.from_numpy(np.ascontiguousarray(theta[n:])))
        self._dirty = True

    def state_dict(self, *args, **kwargs):
        sd = self.spec.state_dict_numpy(self.theta_numpy(), self.ints)
        return OrderedDict((k, torch.from_numpy(v)) for k, v in sd.items())

    def load_state_dict(self, state_dict, strict=True, **kwargs):
        theta = self.theta_numpy()
        sd = {
            k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v))
            for k, v in state_dict.items()
        }
        self.spec.load_state_dict_numpy(sd, theta, self.ints, strict=strict)
        self.ints_version += 1
        self.set_theta_numpy(theta)

    def to(self, device=None, *args, **kwargs):
        if device is not None and torch.device(device).type != "cuda":
            raise RuntimeError("nessai_b200 flows cannot be moved off the GPU")
        return self

    def train(self, mode: bool = True):
        if mode:
            self._dirty = True  # like nflows' LU cache invalidation on train()
        return super().train(mode)

    def mark_dirty(self):
        self._dirty = True

    # --------------------------------------------------------------- folding
    def finalise(self):
        """Fold the eval-mode flow and upload both programs (replaces the LU
        cache rebuild of flowmodel/base.py:680-690)."""
        ff = self.spec.fold(self.theta_numpy(), self.ints)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            for direction, inverse in ((0, False), (1, True)):
                prog = ff.program(inverse)
                ops = np.ascontiguousarray(prog.ops, dtype=np.int32)
                blob = np.ascontiguousarray(prog.blob, dtype=np.float32)
                _lib.check(
                    lib.nb200_flow_set_program(
                        self._handle, direction,
                        ops.ctypes.data_as(C.c_void_p), int(ops.shape[0]),
                        blob.ctypes.data_as(C.c_void_p), int(blob.size),
                        int(prog.final_buf), float(prog.const_logdet),
                    ),
                    "nb200_flow_set_program",
                )
        self._dirty = False

    def end_iteration(self):
        pass

    def _ready(self):
        if self._dirty:
            self.finalise()

    def _prep(self, t: torch.Tensor) -> torch.Tensor:
        if t.dim() != 2 or t.shape[1] != self.spec.D:
            raise ValueError(
                f"Expected features = {self.spec.D}, got {tuple(t.shape)}."
            )
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    # ---------------------------------------------------------- flow surface
    def forward(self, x, context=None):
        """x -> (z, log|J|)  (flows/base.py:209-214)."""
        z, logj, _ = self._forward(x, want_logp=False)
        return z, logj

    def _forward(self, x, want_logp=True, want_z=True, want_logj=True):
        self._ready()
        x = self._prep(x)
        n = x.shape[0]
        # (outputs nobody asked for are not written: log_prob alone saves 4 D + 4 bytes per row)
        z = torch.empty_like(x) if want_z else None
        logj = torch.empty(n, device=self.device, dtype=torch.float32) if want_logj else None
        logp = torch.empty(n, device=self.device, dtype=torch.float32) if want_logp else None
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.load().nb200_flow_forward(
                    self._handle, _ptr(x), _ptr(z), _ptr(logj), _ptr(logp), n, _stream()
                ),
                "nb200_flow_forward",
            )
        return z, logj, logp

    def inverse(self, z, context=None):
        """z -> (x, log|J|)  (flows/base.py:216-221)."""
        x, logj, _ = self._inverse(z, want_logq=False)
        return x, logj

    def _inverse(self, z, want_logq=True):
        self._ready()
        z = self._prep(z)
        n = z.shape[0]
        x = torch.empty_like(z)
        logj = torch.empty(n, device=self.device, dtype=torch.float32)
        logq = torch.empty(n, device=self.device, dtype=torch.float32) if want_logq else None
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.load().nb200_flow_inverse(
                    self._handle, _ptr(z), _ptr(x), _ptr(logj), _ptr(logq), n, _stream()
                ),
                "nb200_flow_inverse",
            )
        return x, logj, logq

    def log_prob(self, x, context=None):
        return self._forward(x, want_z=False, want_logj=False)[2]

    def forward_and_log_prob(self, x, context=None):
        z, _, logp = self._forward(x)
        return z, logp

    def sample_latent_distribution(self, n, context=None):
        if context is not None:
            raise NotImplementedError
        n = int(n)
        z = torch.empty((n, self.spec.D), device=self.device, dtype=torch.float32)
        seed, offset = self._next_rng(n)
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.load().nb200_sample_latent(_ptr(z), n, self.spec.D, seed, offset,
                                                float(np.sqrt(self.spec.base_var)), _stream()),
                "nb200_sample_latent",
            )
        return z

    def base_distribution_log_prob(self, z, context=None):
        """Log-pdf of the base distribution N(0, var I) (flows/base.py:248-257,
        flows/distributions.py:45-56; var = 1: StandardNormal); tiny, torch on device."""
        z = z.to(self.device)
        var = self.spec.base_var
        return -(0.5 / var) * torch.sum(z * z, dim=-1) - 0.5 * self.spec.D * np.log(2 * np.pi * var)

    def sample_and_log_prob(self, N, context=None):
        z = self.sample_latent_distribution(int(N))
        x, _, logq = self._inverse(z)
        return x, logq

    def sample(self, num_samples, context=None):
        return self.sample_and_log_prob(int(num_samples))[0]

    def freeze_transform(self):
        self.theta_p.requires_grad_(False)

    def unfreeze_transform(self):
        self.theta_p.requires_grad_(True)

    # ------------------------------------------------------------------- rng
    _rng_seed = None
    _rng_rows = 0

    def _next_rng(self, n_rows: int):
        """Philox (seed, row offset): the seed is drawn once from torch's CPU
        generator -- the stream the reference seeds with ``torch.manual_seed``
        (/root/reference/src/nessai/samplers/base.py:186-222) -- and rows are
        counted so successive draws never overlap."""
        if self._rng_seed is None:
            self._rng_seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())
        off = self._rng_rows
        self._rng_rows += int(n_rows)
        return C.c_uint64(self._rng_seed), C.c_uint64(off)


class B200FlowModel:
    """See module docstring.  API of ``nessai.flowmodel.FlowModel``."""

    noise_scale = None
    noise_type = None
    model: B200Flow = None

    def __init__(
        self,
        flow_config: Union[dict, None] = None,
        training_config: Union[dict, None] = None,
        output: Union[str, None] = None,
        rng: Optional[np.random.Generator] = None,
    ) -> None:
        if output is None:
            output = os.getcwd()
        self.model = None
        if rng is None:
            logger.debug("No rng specified, using the default rng.")
            rng = np.random.default_rng()
        self.rng = rng
        self.initialised = False
        self.output = output
        os.makedirs(self.output, exist_ok=True)
        self.setup_from_input_dict(flow_config=flow_config, training_config=training_config)
        self.weights_file = None
        self._batch_size = None
        self._optimiser = None

    def setup_from_input_dict(self, flow_config, training_config):
        self.flow_config, self.training_config = update_config(
            flow_config=flow_config, training_config=training_config
        )
        if str(self.training_config.get("device_tag", "cuda")).startswith("cpu"):
            # the reference default; this class only runs on the GPU
            self.training_config["device_tag"] = "cuda"
        self.noise_type = self.training_config.get("noise_type")
        self.noise_scale = self.training_config.get("noise_scale")
        for name, cfg in (("flow_config", self.flow_config), ("training_config", self.training_config)):
            with open(os.path.join(self.output, f"{name}.json"), "w") as fh:
                json.dump(cfg, fh, indent=4, default=str)

    def update_mask(self):
        pass

    # ------------------------------------------------------------- optimiser
    @property
    def optimiser_kwargs(self) -> dict:
        kwds = self.training_config.get("optimiser_kwargs")
        return {} if kwds is None else kwds

    @property
    def optimiser(self) -> str:
        return self.training_config["optimiser"]

    def get_optimiser(self, optimiser=None, **kwargs):
        """flowmodel/base.py:104-135: adam (weight_decay 1e-6) / adamw / sgd with
        torch defaults; the flat parameter buffer is one parameter group, which is
        elementwise identical to the reference's per-tensor groups."""
        optimisers = {
            "adam": (torch.optim.Adam, {"weight_decay": 1e-6}),
            "adamw": (torch.optim.AdamW, {}),
            "sgd": (torch.optim.SGD, {}),
        }
        if self.model is None:
            raise RuntimeError("Cannot initialise optimiser before model")
        if optimiser is None:
            optimiser = self.optimiser
        optim, default_kwargs = optimisers.get(optimiser.lower())
        default_kwargs["lr"] = self.training_config["lr"]
        default_kwargs.update(self.optimiser_kwargs)
        default_kwargs.update(kwargs)
        return optim(self.model.parameters(), **default_kwargs)

    def initialise(self):
        self.update_mask()
        cfg = dict(self.flow_config)
        spec = FlowSpec(cfg)
        self.device = torch.device(self.training_config.get("device_tag", "cuda"))
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.model = B200Flow(spec, self.device)
        self.inference_device = self.device
        self._optimiser = self.get_optimiser()
        self.initialised = True

    def move_to(self, device, update_default=False):
        device = torch.device(device)
        self.model.to(device)
        if update_default:
            self.device = device

    # ------------------------------------------------------------ data prep
    @staticmethod
    def check_batch_size(x, batch_size, min_fraction=0.1):
        """flowmodel/base.py:194-236 (known answers pinned by
        /root/reference/tests/test_flowmodel/test_flowmodel_base.py:145-183)."""
        if batch_size == 1:
            raise ValueError("Cannot use a batch size of 1!")
        n = len(x)
        smallest_tail = int(min_fraction * batch_size)

        def tail_ok(b):
            tail = n % b
            return tail == 0 or tail >= smallest_tail

        if tail_ok(batch_size):
            return batch_size
        # shrink until the last batch is empty or large enough; below the threshold a tail of
        # more than one row is accepted with a warning
        for b in range(batch_size - 1, 1, -1):
            if tail_ok(b):
                return b
            if b <= smallest_tail and n % b > 1:
                logger.warning(f"Batch size is less than {smallest_tail} but valid. Setting batch size to: {b}")
                return b
        raise RuntimeError("Could not find a valid batch size")

    def prep_data(self, samples, val_size, batch_size, weights=None, use_dataloader=False, conditional=None):
        """flowmodel/base.py:238-352; tensors live on the GPU (no DataLoader)."""
        if not self.initialised:
            self.initialise()
        if not np.isfinite(samples).all():
            raise ValueError("Cannot train with non-finite samples!")
        if conditional is not None:
            raise NotImplementedError("nessai_b200: conditional flows are not implemented")
        idx = self.rng.permutation(samples.shape[0])
        samples = samples[idx]
        if weights is not None:
            if not np.isfinite(weights).all():
                raise ValueError("Weights contain non-finite values!")
            weights = weights[idx]
        if val_size is None:
            val_size = 0
        n = int((1 - val_size) * samples.shape[0])
        x_train, x_val = samples[:n], samples[n:]
        if isinstance(batch_size, bool) or not isinstance(batch_size, int):
            if batch_size == "all" or batch_size is None:
                batch_size = x_train.shape[0]
            else:
                raise RuntimeError(f"Unknown batch size: {batch_size}")
        batch_size = self.check_batch_size(x_train, batch_size)
        self._batch_size = batch_size
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(torch.float32).to(self.device)  # noqa: E731
        train_data, val_data = to(x_train), to(x_val)
        if weights is not None:
            return (train_data, to(weights[:n])), (val_data, to(weights[n:])), batch_size
        return train_data, val_data, batch_size

    def end_iteration(self):
        self.model.end_iteration()

    # --------------------------------------------------------------- training
    def _trainer(self):
        """The fused training kernels bound to the current model (rebuilt when the
        permutations / masks change)."""
        model = self.model
        key = (id(model), model.ints_version)
        if getattr(self, "_fused", None) is None or self._fused_key != key:
            from .trainer import FusedTrainer

            # a new flow of the SAME architecture (every level of the importance sampler adds one)
            # reuses the trainer -- plan and workspaces only depend on the architecture; the new
            # flow's permutations are uploaded into it
            arch = (json.dumps(self.flow_config, sort_keys=True, default=str), str(model.device))
            old = getattr(self, "_fused", None)
            if old is not None and getattr(self, "_fused_arch", None) == arch and old.rebind(model):
                pass
            else:
                self._fused = FusedTrainer(model)
            self._fused_arch = arch
            self._fused_key = key
        return self._fused

    def _train(self, train_data, noise_scale=0.0, is_dataloader=False, weighted=False, is_conditional=False):
        """One epoch (flowmodel/base.py:365-452) as ONE call into the fused kernels:
        every batch runs forward + backward + clip + optimiser step on the device.
        The batch permutation comes from torch's CPU generator like the reference's
        ``torch.randperm``; the mean batch loss is read back once per epoch."""
        model = self.model
        model.train()
        x_all, w_all = (train_data if weighted else (train_data, None))
        perm = torch.randperm(x_all.shape[0]).to(self.device)
        if noise_scale:
            x_all = x_all + noise_scale * torch.randn_like(x_all)
        n_batches = -(-x_all.shape[0] // self._batch_size)
        total = self._trainer().epoch(
            x_all, w_all, perm, self._batch_size, self._optimiser, self.training_config["clip_grad_norm"]
        )
        self.end_iteration()
        if self.training_config["annealing"]:
            self._epochs_done += 1
            from .trainer import cosine_annealing_lr

            for g in self._optimiser.param_groups:
                g["lr"] = cosine_annealing_lr(self._base_lr, self._epochs_done, self._t_max)
        self._pending_train_loss = (total, n_batches)
        return None

    def _train_on_device(self, fused, train_data, val_data, weighted, max_epochs, patience, validate, history,
                         first_chunk=8):
        """The epoch loop of flowmodel/base.py:620-662 in chunks of epochs, one cooperative launch
        and one host synchronisation per chunk.  Every epoch's row order is ``torch.randperm``
        from torch's CPU generator, drawn in the reference's order; when the device stops on
        patience inside a chunk the generator is put back to where the reference's would be
        (the permutations of epochs that never ran are not consumed).  Returns the last epoch;
        appends to ``history``; the model holds the best weights when ``validate``."""
        x_all, w_all = (train_data if weighted else (train_data, None))
        x_val, w_val = (val_data if weighted else (val_data, None))
        if x_val is not None and not len(x_val):
            x_val = w_val = None
        n_rows = int(x_all.shape[0])
        n_batches = -(-n_rows // self._batch_size)
        annealing = self.training_config["annealing"]
        from .trainer import cosine_annealing_lr

        lr_now = float(self._optimiser.param_groups[0]["lr"])
        clip = self.training_config["clip_grad_norm"]
        self.model.train()
        fused.begin_run()
        epoch, chunk = 0, int(first_chunk)
        while epoch < max_epochs:
            n = min(chunk, max_epochs - epoch, fused.MAX_CHUNK)
            chunk = min(2 * chunk, fused.MAX_CHUNK)
            perms, states = [], []
            for _ in range(n):
                perms.append(torch.randperm(n_rows))
                states.append(torch.get_rng_state())
            perms = torch.stack(perms).to(self.device, non_blocking=True)
            lrs = [cosine_annealing_lr(self._base_lr, epoch + k, self._t_max) if annealing else lr_now
                   for k in range(n)]
            done, stop, _, hist = fused.run(x_all, w_all, perms, self._batch_size, x_val, w_val, self._optimiser,
                                            clip, lrs, epoch, validate, patience)
            n_run = done - epoch
            history["loss"].extend((hist[:, 0].astype(np.float64) / n_batches).tolist())
            history["val_loss"].extend(hist[:, 1].astype(np.float64).tolist())
            if n_run < n:
                torch.set_rng_state(states[n_run - 1])
            for _ in range(n_run):
                self.end_iteration()
            epoch = done
            if annealing:
                self._epochs_done = epoch
                for g in self._optimiser.param_groups:
                    g["lr"] = cosine_annealing_lr(self._base_lr, epoch, self._t_max)
            if stop:
                logger.debug(f"Epoch {epoch}: Reached patience")
                break
        if validate:
            fused.restore_best()
        return epoch

    def _validate(self, val_data, is_dataloader=False, weighted=False, is_conditional=False):
        """flowmodel/base.py:454-523: eval mode (running statistics); returns the
        device scalar (read back together with the training loss)."""
        x, w = (val_data if weighted else (val_data, None))
        if not len(x):
            return None
        self.model.eval()
        return self._trainer().eval_loss(x, w)

    def finalise(self):
        self.model.finalise()

    def train(self, samples, weights=None, conditional=None, max_epochs=None, patience=None,
              output=None, val_size=None, plot=True):
        """flowmodel/base.py:530-696; returns ``history = {loss, val_loss}``."""
        if not self.initialised:
            self.initialise()
        samples = np.asarray(samples)
        if not np.isfinite(samples).all():
            raise ValueError("Training data is not finite")
        if output is None:
            output = self.output
        else:
            os.makedirs(output, exist_ok=True)
        if val_size is None:
            val_size = self.training_config["val_size"]
        validate = not (val_size == 0.0)
        if self.noise_type == "adaptive":
            from scipy.spatial.distance import cdist

            d = cdist(samples, samples)
            d[d == 0] = np.inf
            noise_scale = self.noise_scale * np.mean(d.min(axis=1))
        elif self.noise_type == "constant":
            noise_scale = self.noise_scale
        else:
            noise_scale = None
        weighted = weights is not None
        train_data, val_data, _ = self.prep_data(
            samples, val_size=val_size, batch_size=self.training_config["batch_size"],
            weights=weights, conditional=conditional,
        )
        if max_epochs is None:
            max_epochs = self.training_config["max_epochs"]
        if self.training_config["annealing"]:
            # torch.optim.lr_scheduler.CosineAnnealingLR(optimiser, max_epochs), closed form
            self._base_lr = float(self._optimiser.param_groups[0]["lr"])
            self._t_max = max_epochs
            self._epochs_done = 0
        if patience is None:
            patience = self.training_config["patience"]
        best_epoch = 0
        best_val_loss = np.inf
        model = self.model
        best = (model.theta_p.detach().clone(), model.theta_b.clone())
        history = dict(loss=[], val_loss=[])
        current_weights_file = os.path.join(output, "model.pt")
        epoch = 0
        fused = self._trainer()
        on_device = (fused._kernel_optimiser(self._optimiser) is not None and not noise_scale
                     and getattr(self, "_device_loop", True))
        if on_device:
            # the whole loop on the device: optimisation steps, validation loss, best-weights
            # snapshot and patience inside one persistent kernel per chunk of epochs
            epoch = self._train_on_device(fused, train_data, val_data, weighted, max_epochs, patience, validate,
                                          history)
        else:
            # input noise / an optimiser the kernels do not implement: epoch by epoch from the host
            for epoch in range(1, max_epochs + 1):
                self._train(train_data, noise_scale=noise_scale, weighted=weighted)
                val_dev = self._validate(val_data, weighted=weighted)
                # the single host synchronisation of the epoch: both losses in one read
                total, n_batches = self._pending_train_loss
                both = self._fused._loss.cpu().numpy()
                loss = float(both[0]) / n_batches
                val_loss = float(both[1]) if val_dev is not None else np.nan
                history["loss"].append(loss)
                history["val_loss"].append(val_loss)
                if validate and (val_loss < best_val_loss):
                    best_epoch = epoch
                    best_val_loss = val_loss
                    best = (model.theta_p.detach().clone(), model.theta_b.clone())
                if validate and (epoch - best_epoch > patience):
                    logger.debug(f"Epoch {epoch}: Reached patience")
                    break
        model.train()
        model.eval()
        if validate and not on_device:
            with torch.no_grad():
                model.theta_p.copy_(best[0])
                model.theta_b.copy_(best[1])
        model.mark_dirty()
        self.finalise()
        self.save_weights(current_weights_file)
        self.model.eval()
        if plot:
            try:
                from nessai.plot import plot_loss

                plot_loss(epoch, history, filename=os.path.join(output, "loss.png"))
            except Exception as e:  # plotting is outside the hot path
                logger.debug(f"Skipping loss plot: {e}")
        return history

    # ---------------------------------------------------------------- weights
    def save_weights(self, weights_file):
        if os.path.exists(weights_file):
            shutil.move(weights_file, weights_file + ".old")
        torch.save(self.model.state_dict(), weights_file)
        self.weights_file = weights_file

    def load_weights(self, weights_file):
        if not self.initialised:
            self.initialise()
        self.model.load_state_dict(torch.load(weights_file, weights_only=True))
        self.model.eval()
        self.weights_file = weights_file

    def reload_weights(self, weights_file):
        if weights_file is None:
            weights_file = self.weights_file
        logger.debug(f"Reloading weights from {weights_file}")
        self.load_weights(weights_file)

    def reset_model(self, weights=True, permutations=False):
        """flowmodel/base.py:745-772."""
        if not any([weights, permutations]):
            logger.debug("Nothing to reset")
            return
        model = self.model
        if weights and permutations:
            theta, ints = model.spec.init_state()
            model.ints = ints
        else:
            theta = model.theta_numpy()
            if weights:
                model.spec.reset_weights(theta)
            else:
                model.spec.reset_permutations(theta, model.ints)
        model.ints_version += 1
        model.set_theta_numpy(theta)
        self._optimiser = self.get_optimiser()

    # -------------------------------------------------------------- inference
    def numpy_array_to_tensor(self, array: np.ndarray, /) -> torch.Tensor:
        """flowmodel/base.py:774-780."""
        # the rows cross PCIe in the caller's dtype and are rounded to fp32 on the device (the same
        # round-to-nearest as the reference's host-side ``.type(torch.float32)``, without a host pass)
        return torch.from_numpy(np.ascontiguousarray(array)).to(self.model.device).to(torch.float32)

    @staticmethod
    def _to_numpy(t: torch.Tensor) -> np.ndarray:
        """Device tensor -> a fresh float64 numpy array (what the reference returns): widened on the
        device, copied once into a page-locked staging buffer that is kept and grown geometrically
        (page-locking a new block per call costs more than the copy), then one host memcpy."""
        t = t.detach()
        if t.device.type != "cuda" or t.numel() < 4096:
            return t.cpu().numpy().astype(np.float64)
        t = t.to(torch.float64).contiguous()
        view = _pinned_stage(t.numel() * 8)[: t.numel() * 8].view(torch.float64).view(t.shape)
        view.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return view.numpy().copy()

    def forward_and_log_prob(self, x: np.ndarray, conditional=None) -> Tuple[np.ndarray, np.ndarray]:
        if conditional is not None:
            raise NotImplementedError("nessai_b200: conditional flows are not implemented")
        x = self.numpy_array_to_tensor(x)
        self.model.eval()
        z, log_prob = self.model.forward_and_log_prob(x)
        return self._to_numpy(z), self._to_numpy(log_prob)

    def inverse(self, z: np.ndarray, conditional=None) -> Tuple[np.ndarray, np.ndarray]:
        if conditional is not None:
            raise NotImplementedError("nessai_b200: conditional flows are not implemented")
        z = self.numpy_array_to_tensor(z)
        self.model.eval()
        x, log_j = self.model.inverse(z)
        return self._to_numpy(x), self._to_numpy(log_j)

    def log_prob(self, x: np.ndarray, conditional=None) -> np.ndarray:
        if conditional is not None:
            raise NotImplementedError("nessai_b200: conditional flows are not implemented")
        x = self.numpy_array_to_tensor(x)
        self.model.eval()
        return self._to_numpy(self.model.log_prob(x))

    def sample(self, n: int = 1, conditional=None) -> np.ndarray:
        return self._to_numpy(self.model.sample(int(n)))

    def sample_latent_distribution(self, n: int = 1) -> np.ndarray:
        return self._to_numpy(self.model.sample_latent_distribution(n))

    def sample_and_log_prob(self, N=1, z=None, conditional=None):
        """flowmodel/base.py:906-948."""
        if self.model is None:
            raise RuntimeError("Model is not initialised yet!")
        if self.model.training:
            self.model.eval()
        if z is None:
            x, log_prob = self.model.sample_and_log_prob(int(N))
        else:
            if isinstance(z, np.ndarray):
                z = self.numpy_array_to_tensor(z)
            x, _, log_prob = self.model._inverse(z)
        return self._to_numpy(x), self._to_numpy(log_prob)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["initialised"] = False
        state.pop("_optimiser", None)
        state.pop("model", None)
        state.pop("flow_config", None)
        state.pop("scheduler", None)
        for k in ("_fused", "_fused_key", "_fused_arch", "_pending_train_loss"):
            state.pop(k, None)
        return state

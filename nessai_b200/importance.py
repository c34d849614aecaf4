"""``ImportanceFlowModel`` on the B200 kernels (importance nested sampling, config 5).

Mirror of /root/reference/src/nessai/flowmodel/importance.py:22-239: a list of
flows, one per level of the importance nested sampler; ``log_prob_all`` evaluates
every stored flow on the same points (the meta-proposal density,
/root/reference/src/nessai/proposal/importance.py:425-440).  Here the points go to
the device ONCE, every flow's ``forward + log_prob`` kernel reads them from there
and the ``(N, K)`` result comes back in one copy; with ``torch.distributed``
initialised the rows are sharded over the ranks and all-gathered.
"""

from __future__ import annotations

import glob
import logging
import os
from typing import List, Optional

import numpy as np
import torch

from .flowmodel import B200Flow, B200FlowModel, update_flow_config
from .spec import FlowSpec

logger = logging.getLogger(__name__)


class B200ImportanceFlowModel(B200FlowModel):
    """Flow model that holds multiple flows for the importance sampler."""

    _resume_n_models: Optional[int] = None

    def __init__(self, flow_config=None, training_config=None, output=None, rng=None):
        self.models: List[B200Flow] = []
        super().__init__(flow_config=flow_config, training_config=training_config, output=output, rng=rng)
        self.weights_files = []

    # -------------------------------------------------------------- the list
    @property
    def model(self):
        """The current flow; ``None`` if no flow has been added."""
        if self.models:
            return self.models[-1]
        return None

    @model.setter
    def model(self, model):
        if model is not None:
            self.models.append(model)

    @property
    def n_models(self) -> int:
        return len(self.models) if self.models else 0

    def initialise(self) -> None:
        self.initialised = True

    def reset_optimiser(self) -> None:
        self._optimiser = self.get_optimiser()

    def _new_flow(self) -> B200Flow:
        self.device = torch.device(self.training_config.get("device_tag", "cuda"))
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.inference_device = self.device
        return B200Flow(FlowSpec(dict(self.flow_config)), self.device)

    def add_new_flow(self, reset=False):
        """importance.py:80-99: a fresh flow, or a copy of the current one."""
        if reset or not self.models:
            new_flow = self._new_flow()
        else:
            cur = self.model
            new_flow = self._new_flow()
            new_flow.ints = {k: np.array(v, copy=True) for k, v in cur.ints.items()}
            new_flow.ints_version += 1
            new_flow.set_theta_numpy(cur.theta_numpy())
        for m in self.models:
            m.eval()
        self.models.append(new_flow)
        self.reset_optimiser()

    # ------------------------------------------------------------- inference
    def _device_rows(self, x: np.ndarray, model: B200Flow) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(x)).to(model.device).to(torch.float32)

    def log_prob_ith(self, x, i):
        """importance.py:101-113."""
        m = self.models[i]
        if m.training:
            m.eval()
        return self._to_numpy(m.log_prob(self._device_rows(x, m)))

    def log_prob_all(self, x, shard=True):
        """importance.py:115-129: ``(N, n_models)`` float64.  One H2D copy of the
        points, one forward kernel per flow, one D2H copy of the whole matrix; rows are
        sharded over ``torch.distributed`` ranks when a process group is initialised."""
        import torch.distributed as dist

        n_models = self.n_models
        x = np.asarray(x)
        N = x.shape[0]
        world = dist.get_world_size() if (shard and dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank() if world > 1 else 0
        per = -(-N // world) if world > 1 else N
        lo, hi = min(rank * per, N), min((rank + 1) * per, N)
        dev = self.model.device
        out = torch.empty((n_models, per), device=dev, dtype=torch.float32)
        if hi > lo:
            xd = self._device_rows(x[lo:hi], self.model)
            models = self.models[:n_models]
            for m in models:
                if m.training:
                    m.eval()
                m._ready()  # fold + upload outside the fan-out below
            # The K flows are independent and each leaves most SMs idle on a level's worth of rows:
            # fan the K forward kernels out over a few streams (x stays resident, one weight image
            # per flow), then join -- the kernels of different flows overlap on the device.
            main = torch.cuda.current_stream(dev)
            streams = self._lp_streams(dev, min(n_models, 8))
            fork = torch.cuda.Event()
            fork.record(main)
            for i, m in enumerate(models):
                st = streams[i % len(streams)]
                if i < len(streams):
                    st.wait_event(fork)
                with torch.cuda.stream(st):
                    out[i, : hi - lo] = m._forward(xd, want_z=False, want_logj=False)[2]
            for st in streams:
                main.wait_stream(st)
            xd.record_stream(main)
        if world > 1:
            full = torch.empty((world, n_models, per), device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(full, out)
            out = full.permute(1, 0, 2).reshape(n_models, world * per)[:, :N]
        return self._to_numpy(out.t())

    def _lp_streams(self, dev, n):
        cur = getattr(self, "_streams", None)
        if cur is None or len(cur) < n or cur[0].device != dev:
            self._streams = cur = [torch.cuda.Stream(device=dev) for _ in range(n)]
        return cur[:n]

    def sample_ith(self, i, N=1):
        """importance.py:131-142."""
        if self.models is None:
            raise RuntimeError("Models are not initialised yet!")
        m = self.models[i]
        if m.training:
            m.eval()
        return self._to_numpy(m.sample(int(N)))

    # --------------------------------------------------------------- weights
    def save_weights(self, weights_file) -> None:
        super().save_weights(weights_file)
        self.weights_files.append(self.weights_file)

    def load_all_weights(self) -> None:
        """importance.py:149-165: rebuild every flow from its weights file."""
        self.models = []
        logger.debug(f"Loading weights from {self.weights_files}")
        for wf in self.weights_files:
            new_flow = self._new_flow()
            new_flow.load_state_dict(torch.load(wf, weights_only=True))
            new_flow.eval()
            self.models.append(new_flow)

    def update_weights_path(self, weights_path: str, n: Optional[int] = None) -> None:
        """importance.py:167-206."""
        all_weights_files = glob.glob(os.path.join(weights_path, "", "level_*", "model.pt"))
        if n is None:
            if self.n_models:
                n = self.n_models
            else:
                raise RuntimeError("n is None and no models are defined, cannot update weights path.")
        if len(all_weights_files) < n:
            raise RuntimeError(f"Cannot use weights from: {weights_path}. Not enough files.")
        elif len(all_weights_files) > n:
            logger.warning("More weights files than expected. Some files will be skipped.")
        self.weights_files = [os.path.join(weights_path, f"level_{i}", "model.pt") for i in range(n)]

    def resume(self, flow_config: dict, training_config: Optional[dict] = None,
               weights_path: Optional[str] = None) -> None:
        """importance.py:208-226."""
        self.flow_config = update_flow_config(flow_config)
        if training_config is not None:
            self.training_config = training_config
        if weights_path is None:
            weights_path = self.output
        self.update_weights_path(weights_path, n=self._resume_n_models)
        self.load_all_weights()
        self.initialise()

    def __getstate__(self):
        d = self.__dict__
        exclude = {"models", "_optimiser", "flow_config", "_fused", "_fused_key", "_fused_arch", "_pending_train_loss", "scheduler", "_streams"}
        state = {k: d[k] for k in d.keys() - exclude}
        state["initialised"] = False
        state["models"] = None
        state["_resume_n_models"] = len(d["models"]) if d.get("models") else 0
        return state

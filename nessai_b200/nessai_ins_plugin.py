"""Bindings for nessai's importance nested sampler (config 5, SURVEY.md 8b "(INS)").

``ImportanceFlowProposal`` builds its flow model inline and
``ImportanceNestedSampler.get_proposal`` builds the proposal inline -- there is no
class hook -- so both are subclassed with the construction restated and the class
swapped.  Everything else (level construction, meta-proposal weights, evidence) is
the reference's own code, unmodified.
"""

from __future__ import annotations

import logging
import os

from nessai.proposal.importance import ImportanceFlowProposal
from nessai.samplers.importancesampler import ImportanceNestedSampler

from .importance import B200ImportanceFlowModel

logger = logging.getLogger(__name__)


class B200ImportanceFlowProposal(ImportanceFlowProposal):
    """``ImportanceFlowProposal`` whose flows -- one per level -- live on a B200:
    training (fused kernels), ``sample_ith`` and the meta-proposal density
    ``log_prob_all`` go through :class:`B200ImportanceFlowModel`
    (/root/reference/src/nessai/proposal/importance.py:153-168 restated)."""

    def initialise(self):
        self._check_fields()
        if self.initialised:
            logger.debug("Proposal already initialised")
            return
        self.verify_rescaling()
        self.flow = B200ImportanceFlowModel(
            flow_config=self.flow_config,
            training_config=self.training_config,
            output=self.output,
        )
        self.flow.initialise()
        # Proposal.initialise (the grandparent): mark as initialised
        super(ImportanceFlowProposal, self).initialise()


class B200ImportanceNestedSampler(ImportanceNestedSampler):
    """``ImportanceNestedSampler`` with the proposal above
    (/root/reference/src/nessai/samplers/importancesampler.py:684-688)."""

    def get_proposal(self, subdir: str = "levels", **kwargs):
        output = os.path.join(self.output, subdir, "")
        return B200ImportanceFlowProposal(self.model, output, **kwargs)

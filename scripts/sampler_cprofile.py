"""cProfile of a whole FlowSampler run (the reference's sampler, unmodified) with the B200 proposal on an
8-D Gaussian: which share of the wall time is ours (nessai_b200/*), and where inside it.
needs baseline/_ref (the installed reference) + oracle/shims on the path, as the tests set them up."""
import cProfile, io, os, pstats, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import conftest  # noqa: F401  (puts baseline/_ref and the glasflow shim on sys.path)
conftest.reference_or_skip()
import numpy as np
from nessai.flowsampler import FlowSampler
from nessai.model import Model
from nessai_b200.nessai_plugin import B200NessaiFlowProposal

D = int(os.environ.get("NB200_PROFILE_D", 8))
class Gaussian(Model):
    def __init__(self):
        self.names = [f"x{i}" for i in range(D)]
        self.bounds = {n: [-10.0, 10.0] for n in self.names}
    def log_prior(self, x):
        return np.log(self.in_bounds(x), dtype="float") - D * np.log(20.0)
    def log_likelihood(self, x):
        return -0.5 * np.sum(self.unstructured_view(x) ** 2, axis=-1) - 0.5 * D * np.log(2 * np.pi)

fs = FlowSampler(Gaussian(), output=tempfile.mkdtemp(), resume=False, seed=1234, nlive=1000, plot=False,
                 flow_proposal_class=B200NessaiFlowProposal, checkpointing=False, max_iteration=int(sys.argv[1]) if len(sys.argv) > 1 else 6000,
                 **({"drawsize": int(os.environ["NB200_PROFILE_DRAWSIZE"])} if os.environ.get("NB200_PROFILE_DRAWSIZE") else {}))
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable(); fs.run(plot=False, save=False); pr.disable()
wall = time.perf_counter() - t0
prop = fs.ns._flow_proposal
print(f"wall {wall:.2f} s; trainings {prop.training_count}, populates {prop.populated_count}, population_time {prop.population_time.total_seconds():.3f} s, training_time {fs.ns.training_time.total_seconds():.3f} s, logZ {fs.ns.log_evidence:.3f}")
s = io.StringIO(); st = pstats.Stats(pr, stream=s); st.sort_stats("cumulative").print_stats("nessai_b200|flowproposal|flowmodel", 30); print(s.getvalue()[:7000])
